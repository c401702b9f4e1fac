"""The estimator node shim's host-side message synchronisation (lf-vio_b200/host/nodes.cpp: imu_callback, feature_callback,
getMeasurements, the IMU hand-over of process() with interpolation at stamp + td) on the CPU against the Python restatement
(oracle/nodes.py).  No device: the node is built without a device window and is fed fewer than WINDOW_SIZE + 1 images, which
only fills the window (Estimator::processIMU buffers samples, processImage adds features)."""
import ctypes as C
import os

import numpy as np
import pytest

from conftest import ROOT
from oracle.nodes import EstimatorNodeOracle


@pytest.fixture(scope="module")
def host():
    lib = C.CDLL(os.path.join(ROOT, "lf-vio_b200", "liblfvio_host.so"))
    lib.lfvio_host_estimator_node_create.restype = C.c_void_p
    lib.lfvio_host_estimator_node_create.argtypes = [C.c_int, C.c_int, C.c_int, C.c_char_p]
    lib.lfvio_host_estimator_node_destroy.argtypes = [C.c_void_p]
    lib.lfvio_host_estimator_node_imu.argtypes = [C.c_void_p, C.c_double, C.c_void_p, C.c_void_p]
    lib.lfvio_host_estimator_node_feature.argtypes = [C.c_void_p, C.c_double, C.c_void_p, C.c_int]
    lib.lfvio_host_estimator_node_process.argtypes = [C.c_void_p]
    lib.lfvio_host_estimator_node_imu_buf.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
    lib.lfvio_host_estimator_node_set_td.argtypes = [C.c_void_p, C.c_double]
    lib.lfvio_host_estimator_node_restart.argtypes = [C.c_void_p]
    return lib


def _feat(rng, n=30):
    b = rng.normal(size=(n, 3))
    b /= np.linalg.norm(b, axis=1, keepdims=True)
    return np.ascontiguousarray(np.hstack([b, np.arange(n)[:, None], rng.uniform(1, 900, (n, 2)), rng.normal(0, 0.1, (n, 3))]), np.float32)


def _slot(host, node, j):
    cap = 256
    dt, acc, gyr, info = np.zeros(cap), np.zeros((cap, 3)), np.zeros((cap, 3)), np.zeros(4, np.int32)
    n = host.lfvio_host_estimator_node_imu_buf(node, j, dt.ctypes.data, acc.ctypes.data, gyr.ctypes.data, cap, info.ctypes.data)
    return n, dt[:max(n, 0)], acc[:max(n, 0)], gyr[:max(n, 0)], info


@pytest.mark.parametrize("td", [0.0, -0.008, 0.0031])
def test_measurement_synchronisation(host, td):
    rng = np.random.default_rng(3)
    node = host.lfvio_host_estimator_node_create(10, 200, -1, None)
    assert node
    host.lfvio_host_estimator_node_set_td(node, td)
    ref = EstimatorNodeOracle(td)
    t0, imu_dt, n_img = 50.0, 0.005, 9
    img_times = [t0 + 0.0312 + 0.1 * k for k in range(n_img + 1)]       # images fall between IMU samples
    t_imu = t0 + 0.02                                                    # the IMU starts after the first image: that image is thrown
    processed = ref_processed = 0
    k_img = 0
    for step in range(int(1.05 / imu_dt)):
        t = t_imu + step * imu_dt
        acc, gyr = rng.normal([0, 0, 9.8], 0.3), rng.normal(0, 0.1, 3)
        host.lfvio_host_estimator_node_imu(node, t, acc.ctypes.data, gyr.ctypes.data)
        ref.imu_callback(t, acc, gyr)
        if step == 40:                                                   # a sample that arrives out of order is dropped
            host.lfvio_host_estimator_node_imu(node, t - 0.02, acc.ctypes.data, gyr.ctypes.data)
            ref.imu_callback(t - 0.02, acc, gyr)
        while k_img <= n_img and img_times[k_img] <= t - 0.012:          # images arrive ~12 ms late, as on a real system
            f = _feat(rng)
            host.lfvio_host_estimator_node_feature(node, img_times[k_img], f.ctypes.data, len(f))
            ref.feature_callback(img_times[k_img])
            k_img += 1
        processed += host.lfvio_host_estimator_node_process(node)
        ref_processed += ref.process()
        assert processed == ref_processed, step                          # the same images are released at the same moments
    assert processed == len(ref.frames)
    assert 6 <= processed <= 9
    # slot j of the window holds the samples handed over before image j (slot 0: none, estimator.cpp:100-104)
    n0, *_ = _slot(host, node, 0)
    assert n0 == 0
    for j in range(1, processed):
        n, dt, acc, gyr, info = _slot(host, node, j)
        calls = ref.frames[j][1]
        assert n == len(calls), j
        assert np.allclose(dt, [c[0] for c in calls], rtol=0, atol=1e-12)
        assert np.allclose(acc, [c[1] for c in calls], atol=1e-12) and np.allclose(gyr, [c[2] for c in calls], atol=1e-12)
        assert abs(dt.sum() - 0.1) < 1e-9                                # every interval spans exactly one image period
        if td != 0.0:
            assert 0 < dt[-1] < imu_dt                                   # the interpolated sample at stamp + td
    *_, info = _slot(host, node, 0)
    assert info[0] == processed and info[1] == ref.sum_of_wait
    assert info[2] == len(ref.imu_buf) and info[3] == len(ref.feature_buf)
    host.lfvio_host_estimator_node_restart(node)                         # restart_callback: buffers cleared, estimator reset
    *_, info = _slot(host, node, 0)
    assert info[0] == 0 and info[2] == 0 and info[3] == 0
    host.lfvio_host_estimator_node_destroy(node)


def _freq_control_reference(stamps, freq):
    """feature_tracker_node.cpp:27-62, 110-112, 173-178 restated: per image -1 (nothing published), -2 (stream restart) or 0
    (a message is published; it is empty here because there is no device)."""
    first_flag, first_time, last_time, pub_count, init_pub = True, 0.0, 0.0, 1, False
    out = []
    for t in stamps:
        if first_flag:
            first_flag, first_time, last_time = False, t, t
            out.append(-1)
            continue
        if t - last_time > 1.0 or t < last_time:
            first_flag, last_time, pub_count = True, 0.0, 1
            out.append(-2)
            continue
        last_time = t
        r = 1.0 * pub_count / (t - first_time)
        pub = np.floor(r + 0.5) <= freq                       # C round(): half away from zero (r > 0)
        if pub and abs(r - freq) < 0.01 * freq:
            first_time, pub_count = t, 0
        if not pub:
            out.append(-1)
            continue
        pub_count += 1
        out.append(0 if init_pub else -1)
        init_pub = True
    return out


@pytest.mark.parametrize("rate,freq", [(15.0, 10), (30.0, 10), (20.0, 20), (10.0, 30)])
def test_tracker_node_frequency_control(host, rate, freq):
    host.lfvio_host_tracker_node_create.restype = C.c_void_p
    host.lfvio_host_tracker_node_destroy.argtypes = [C.c_void_p]
    host.lfvio_host_tracker_node_image.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_double, C.c_void_p, C.c_int]
    host.lfvio_host_tracker_node_set_freq.argtypes = [C.c_int]
    host.lfvio_host_tracker_node_set_freq(freq)
    node = host.lfvio_host_tracker_node_create()
    stamps = [200.0 + k / rate for k in range(60)]
    stamps += [stamps[-1] + 1.5 + k / rate for k in range(25)]            # a gap of more than a second: restart
    stamps += [stamps[-1] - 0.3] + [stamps[-1] - 0.3 + (k + 1) / rate for k in range(25)]   # time goes backwards: restart
    img = np.zeros((4, 4), np.uint8)
    out = np.zeros((8, 9), np.float32)
    got = [host.lfvio_host_tracker_node_image(node, img.ctypes.data, 4, 4, 4, t, out.ctypes.data, 8) for t in stamps]
    host.lfvio_host_tracker_node_destroy(node)
    want = _freq_control_reference(stamps, freq)
    assert got == want
    assert want.count(-2) == 2
    published = want.count(0)
    if rate > freq:                                                       # throttled to about FREQ messages per second
        span = 60 / rate + 25 / rate + 26 / rate
        assert abs(published / span - freq) < 0.25 * freq
    else:                                                                 # slower camera than FREQ: every frame goes out
        assert published >= len(stamps) - 2 * 3 - 2
