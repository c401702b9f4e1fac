"""The ROS-free node shims (lf-vio_b200/host/nodes.{h,cpp}): feature_tracker_node's img_callback (frequency control,
first-publish suppression, PointCloud wire format) and estimator_node's imu/feature callbacks + process()
(measurement synchronisation, IMU interpolation at the image time, cold-start initialisation, odometry / TUM output)."""
import ctypes as C
import os

import numpy as np
import pytest

from conftest import ROOT
from lfvio import synth, synthwin, violoop
from lfvio.config import PAL_UP, tracker_defaults

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def host():
    lib = C.CDLL(os.path.join(ROOT, "lf-vio_b200", "liblfvio_host.so"))
    lib.lfvio_host_tracker_node_create.restype = C.c_void_p
    lib.lfvio_host_tracker_node_destroy.argtypes = [C.c_void_p]
    lib.lfvio_host_tracker_node_configure.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_uint64]
    lib.lfvio_host_tracker_node_image.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_double, C.c_void_p, C.c_int]
    lib.lfvio_host_estimator_node_create.restype = C.c_void_p
    lib.lfvio_host_estimator_node_create.argtypes = [C.c_int, C.c_int, C.c_int, C.c_char_p]
    lib.lfvio_host_estimator_node_destroy.argtypes = [C.c_void_p]
    lib.lfvio_host_estimator_node_estimator.restype = C.c_void_p
    lib.lfvio_host_estimator_node_estimator.argtypes = [C.c_void_p]
    lib.lfvio_host_estimator_node_imu.argtypes = [C.c_void_p, C.c_double, C.c_void_p, C.c_void_p]
    lib.lfvio_host_estimator_node_feature.argtypes = [C.c_void_p, C.c_double, C.c_void_p, C.c_int]
    lib.lfvio_host_estimator_node_process.argtypes = [C.c_void_p]
    lib.lfvio_host_estimator_node_latest.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.lfvio_host_estimator_set_params.argtypes = [C.c_void_p] + [C.c_void_p] * 4 + [C.c_double, C.c_double, C.c_int, C.c_void_p, C.c_void_p]
    lib.lfvio_host_estimator_enable_initialization.argtypes = [C.c_void_p, C.c_uint64, C.c_int]
    lib.lfvio_host_estimator_init_info.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    lib.lfvio_host_estimator_get.argtypes = [C.c_void_p] + [C.c_void_p] * 7
    return lib


def test_tracker_node_frequency_control_and_wire_format(lfvio_mod, gpu, host):
    n, rate, freq = 12, 15.0, 10
    src = synth.FrameSource(960, 1280, seed=3, rot_per_frame=0.004, shift_per_frame=(6.5, -4.2))
    frames = [np.ascontiguousarray(src.frame(k)) for k in range(n)]
    stamps = [100.0 + k / rate for k in range(n)]
    cfg = tracker_defaults()
    cam = lfvio_mod.Ocam.from_dict(PAL_UP)
    node = host.lfvio_host_tracker_node_create()
    assert host.lfvio_host_tracker_node_configure(node, C.byref(cam), 960, 1280, cfg["max_cnt"], cfg["min_dist"], freq, 1, 11) == 0
    cap = 512
    out = np.zeros((cap, 9), np.float32)
    # the reference's control flow restated (feature_tracker_node.cpp:29-62, 173-178)
    first_flag, first_time, pub_count, init_pub = True, 0.0, 1, False
    expect, got, msgs = [], [], {}
    for k in range(n):
        t = stamps[k]
        if first_flag:
            first_flag, first_time = False, t
            e = -1
        else:
            pub = round(1.0 * pub_count / (t - first_time)) <= freq
            if pub and abs(1.0 * pub_count / (t - first_time) - freq) < 0.01 * freq:
                first_time, pub_count = t, 0
            if pub:
                pub_count += 1
                e = 1 if init_pub else -1
                init_pub = True
            else:
                e = -1
        expect.append(e)
        r = host.lfvio_host_tracker_node_image(node, frames[k].ctypes.data, 960, 1280, 1280, t, out.ctypes.data, cap)
        got.append(1 if r >= 0 else r)
        if r >= 0:
            msgs[k] = out[:r].copy()
    host.lfvio_host_tracker_node_destroy(node)
    assert got == expect, (got, expect)
    assert sum(1 for e in expect if e == 1) >= 5
    last = msgs[max(msgs)]
    assert len(last) > 50
    assert np.allclose(np.linalg.norm(last[:, :3], axis=1), 1.0, atol=1e-5)       # unit bearings
    ids = last[:, 3]
    assert np.array_equal(ids, np.round(ids)) and len(set(ids.tolist())) == len(ids)  # integral float32 ids, unique
    assert (last[:, 4] >= 1).all() and (last[:, 4] < 1279).all() and (last[:, 5] >= 1).all() and (last[:, 5] < 959).all()
    assert np.abs(last[:, 6:9]).max() > 0                                          # bearing velocities present
    # a feature published in consecutive messages keeps its id and moves smoothly
    ks = sorted(msgs)
    a, b = msgs[ks[-2]], msgs[ks[-1]]
    common = set(a[:, 3].tolist()) & set(b[:, 3].tolist())
    assert len(common) > 40


def test_estimator_node_cold_start_stream(lfvio_mod, gpu, host, tmp_path):
    """IMU at 200 Hz and feature messages at 10 Hz fed through the callbacks; process() synchronises them as
    getMeasurements does, the estimator initialises itself (initialStructure) and tracks; odometry and the TUM file agree."""
    E = synthwin.ESTIMATOR
    old_td = E["td"]
    E["td"] = 0.0
    try:
        world = violoop.World(seed=2, n_frames=36, pixel_noise=0.1)
    finally:
        E["td"] = old_td
    result = str(tmp_path / "vins_result.txt")
    node = host.lfvio_host_estimator_node_create(10, 1000, gpu, result.encode())
    assert node, lfvio_mod.lib.lfvio_last_error()
    est = host.lfvio_host_estimator_node_estimator(node)
    vp = lambda a: np.ascontiguousarray(a, np.float64).ctypes.data_as(C.c_void_p)
    keep = [np.ascontiguousarray(x, np.float64) for x in (world.noise, world.G, world.tic, world.ric, np.zeros(3), np.zeros(3))]
    assert host.lfvio_host_estimator_set_params(est, vp(keep[0]), vp(keep[1]), vp(keep[2]), vp(keep[3]), 0.0, 0.0, 0, vp(keep[4]), vp(keep[5])) == 0
    assert host.lfvio_host_estimator_enable_initialization(est, 9, 0) == 0
    processed = 0
    a0, g0 = world.imu[1][0], world.imu[1][1]
    host.lfvio_host_estimator_node_imu(node, float(world.times[0]), vp(a0), vp(g0))
    n_sub = len(world.imu[1][2])
    for k in range(world.n_frames):
        if k > 0:
            _, _, dt, acc, gyr = world.imu[k]
            for j in range(n_sub):
                stamp = float(world.times[k]) if j == n_sub - 1 else float(world.times[k - 1] + world.frame_dt * (j + 1) / n_sub)
                host.lfvio_host_estimator_node_imu(node, stamp, vp(acc[j]), vp(gyr[j]))
        obs = world.frames[k]
        feat = np.array([[*b, fid, 0.0, row, *vel] for fid, (b, vel, row) in obs.items()], np.float32)
        host.lfvio_host_estimator_node_feature(node, float(world.times[k]), feat.ctypes.data, len(feat))
        processed += host.lfvio_host_estimator_node_process(node)
    # the first feature message is dropped (estimator_node.cpp:164-171) and the newest waits for a later IMU sample
    assert processed == world.n_frames - 2
    F = 11
    P, R, V = np.zeros((F, 3)), np.zeros((F, 9)), np.zeros((F, 3))
    ba, bg, td, info, ii, g = np.zeros((F, 3)), np.zeros((F, 3)), np.zeros(1), np.zeros(8, np.int32), np.zeros(8, np.int32), np.zeros(3)
    host.lfvio_host_estimator_get(est, vp(P), vp(R), vp(V), vp(ba), vp(bg), vp(td), info.ctypes.data_as(C.c_void_p))
    host.lfvio_host_estimator_init_info(est, ii.ctypes.data_as(C.c_void_p), vp(g))
    lp, lq, lv = np.zeros(3), np.zeros(4), np.zeros(3)
    n_odo = host.lfvio_host_estimator_node_latest(node, vp(lp), vp(lq), vp(lv))
    host.lfvio_host_estimator_node_destroy(node)
    assert ii[0] == 1 and ii[1] == 1 and info[4] == 0                      # initialised at the first attempt, no failure
    # frames 1..11 fill the window; odometry from the initialisation frame on
    assert n_odo == processed - 10
    lines = open(result).read().strip().split("\n")
    assert len(lines) == n_odo
    tum = np.array([[float(x) for x in ln.split()] for ln in lines])
    assert all(len(ln.split()) == 8 and len(ln.split()[1].split(".")[1]) == 12 for ln in lines)   # fixed, 12 decimals
    ks = [int(round((t - world.times[0]) / world.frame_dt)) for t in tum[:, 0]]
    assert ks == list(range(11, 11 + n_odo))
    truth = np.array([world.truth(k)[0] for k in ks])
    ate = violoop.ate_rmse(tum[:, 1:4], truth)
    assert np.allclose(np.linalg.norm(tum[:, 4:8], axis=1), 1.0, atol=1e-9)
    print(f"estimator node cold start: {n_odo} odometry messages, ATE {ate:.4f} m")
    assert ate < 0.05, ate
    # the IMU-rate prediction (update()/predict()) sits at the newest processed state propagated by the buffered samples
    assert np.linalg.norm(lp - P[F - 1]) < 0.5 and abs(np.linalg.norm(lq) - 1) < 1e-3
