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
