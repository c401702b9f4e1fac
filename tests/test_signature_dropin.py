"""The host mirrors carry the reference's own signatures where its third-party types exist (missing #8):
readImage(const cv::Mat &, double), fisheye_mask / cur_img, Camera::liftProjective / spaceToPlane on Eigen vectors,
processImage(const map<...Eigen::Matrix<double, 8, 1>...> &, const std_msgs::Header &) — switched on by __has_include.  This image
has neither OpenCV nor Eigen nor ROS, so the translation unit that calls them the way feature_tracker_node.cpp /
estimator_node.cpp do is compiled against stand-in headers (tests/native/stubs, oracle/ref_stubs).  CPU only."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
INC = ["-I" + os.path.join(ROOT, "include"), "-I" + os.path.join(ROOT, "lf-vio_b200", "host")]
STUBS = ["-I" + os.path.join(ROOT, "tests", "native", "stubs"), "-I" + os.path.join(ROOT, "oracle", "ref_stubs")]
SRC = os.path.join(ROOT, "tests", "native", "signature_dropin.cpp")


def _cc(extra):
    return subprocess.run(["g++", "-std=c++17", "-fsyntax-only", "-Wall", *INC, *extra, SRC], capture_output=True, text=True)


def test_reference_signatures_compile_where_the_third_party_headers_exist():
    r = _cc(STUBS)
    assert r.returncode == 0, r.stderr


def test_overloads_are_off_without_the_headers_and_the_mirrors_still_build():
    r = _cc([])  # no OpenCV / Eigen / ROS on the include path: the #error in the test source fires, nothing else
    assert r.returncode != 0 and "guarded overloads were not enabled" in r.stderr, r.stderr
    hdr_only = subprocess.run(["g++", "-std=c++17", "-fsyntax-only", *INC, "-include", "estimator.h", "-include", "feature_tracker.h",
                               "-x", "c++", os.devnull], capture_output=True, text=True)
    assert hdr_only.returncode == 0, hdr_only.stderr


def test_space_to_plane_of_the_host_mirror():
    """OCAMCamera::spaceToPlane (ScaramuzzaCamera.cc:653-674) of the host mirror: the numpy statement of the oracle to 1e-9 px and, where oracle/_ref
    is there, the reference's own code bit for bit; lift -> project is the identity on the annulus."""
    import sys

    sys.path.insert(0, os.path.join(ROOT, "lf-vio_b200"))
    sys.path.insert(0, ROOT)
    import lfvio
    from lfvio.config import PAL_UP
    from oracle import camera as OC
    from oracle import ref

    path = os.path.join(ROOT, "lf-vio_b200", "liblfvio_host.so")
    if not os.path.exists(path):
        pytest.skip("host mirror library not built")
    host = C.CDLL(path)
    host.lfvio_host_ocam_space_to_plane.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
    rng = np.random.default_rng(3)
    ang = rng.uniform(0, 2 * np.pi, 2000)
    rad = rng.uniform(160, 500, 2000)
    uv = np.stack([PAL_UP["cx"] + rad * np.cos(ang), PAL_UP["cy"] + rad * np.sin(ang)], 1)
    P = np.ascontiguousarray(OC.lift_projective(PAL_UP, uv) * rng.uniform(0.5, 20.0, (2000, 1)))
    out = np.zeros((2000, 2))
    ocam = lfvio.Ocam.from_dict(PAL_UP)
    assert host.lfvio_host_ocam_space_to_plane(C.byref(ocam), P.ctypes.data, out.ctypes.data, len(P)) == 0
    want = OC.space_to_plane(PAL_UP, P)
    assert np.abs(out - want).max() < 1e-9  # (numpy's arctan2 and libm's atan2 may differ in the last bit)
    if ref.available(build=False):
        assert np.array_equal(out, ref.space_to_plane(PAL_UP, P))  # same libm, same operation order: bit for bit
    assert np.abs(out - uv).max() < 0.1  # the calibration's inverse polynomial is a fit: pixels, not ulps
