"""examples/vio_from_files.cpp (the ROS-free driver of the node shims) compiles and links against the in-tree libraries,
prints its usage, and fails loudly — not silently on a CPU path — when there is no CUDA device."""
import os
import subprocess

import pytest

from conftest import ROOT


@pytest.fixture(scope="module")
def exe(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("ex") / "vio_from_files")
    lib = os.path.join(ROOT, "lf-vio_b200")
    subprocess.run(["g++", "-O1", "-std=c++17", "-Wall", "-Werror", "-I" + os.path.join(ROOT, "include"), "-I" + os.path.join(lib, "host"),
                    os.path.join(ROOT, "examples", "vio_from_files.cpp"), "-L" + lib, "-llfvio_host", "-llfvio", "-Wl,-rpath," + lib, "-o", out],
                   check=True)
    return out


def test_example_usage(exe):
    r = subprocess.run([exe, "--help"], capture_output=True, text=True)
    assert r.returncode == 0 and "usage: vio_from_files" in r.stdout
    assert subprocess.run([exe], capture_output=True).returncode == 2


def test_example_needs_a_device(exe, tmp_path, lfvio_mod):
    if lfvio_mod.lib.lfvio_device_count() > 0:
        pytest.skip("a CUDA device is present")
    (tmp_path / "imu.txt").write_text("0.0 0 0 9.8 0 0 0\n")
    (tmp_path / "feat.txt").write_text("F 0.0 0\n")
    r = subprocess.run([exe, str(tmp_path / "imu.txt"), str(tmp_path / "feat.txt"), str(tmp_path / "out.txt")], capture_output=True, text=True)
    assert r.returncode == 1 and "no estimator" in r.stderr


@pytest.mark.gpu
def test_example_runs_a_stream_both_ways(exe, tmp_path, gpu):
    """The driver on a 40-frame synthetic stream (cold start included), FeatureManager on the device (its default) and on the
    host: both write a trajectory that follows the ground truth, and the two agree to within the solve's run-to-run spread."""
    import sys

    import numpy as np

    from lfvio import violoop

    d = str(tmp_path)
    subprocess.run([sys.executable, os.path.join(ROOT, "tools", "make_example_stream.py"), d, "--frames", "40"], check=True)
    out = {}
    for name, host_list in (("device", "0"), ("host", "1")):
        res = os.path.join(d, f"result_{name}.txt")
        r = subprocess.run([exe, os.path.join(d, "imu.txt"), os.path.join(d, "features.txt"), res, os.path.join(d, "calib.txt"), "0.0", "9", host_list],
                           capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
        out[name] = np.loadtxt(res)
    truth = np.loadtxt(os.path.join(d, "truth_tum.txt"))
    assert out["device"].shape == out["host"].shape and len(out["device"]) >= 25
    assert np.abs(out["device"][:, 1:4] - out["host"][:, 1:4]).max() < 1e-4
    by_stamp = {round(t, 6): p for t, p in zip(truth[:, 0], truth[:, 1:4])}
    tru = np.array([by_stamp[round(t, 6)] for t in out["device"][:, 0]])
    assert violoop.ate_rmse(out["device"][:, 1:4], tru) < 0.05
