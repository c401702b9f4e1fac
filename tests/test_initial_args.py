"""Argument validation of the initialisation entries (no GPU needed: every check happens before the first CUDA call)."""
import ctypes as C

import numpy as np

from lfvio import initial as I

ERR_ARG = -2
_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)


def _p(a, t=C.c_double):
    return a.ctypes.data_as(C.POINTER(t))


def test_epnp_rejects_short_and_unanchored_problems(lfvio_mod):
    lib = lfvio_mod.lib
    pw, us = np.zeros((3, 3)), np.zeros((3, 3))
    R, T = np.zeros(9), np.zeros(3)
    first = np.array([0, 3], np.int32)
    assert lib.lfvio_epnp(_p(pw), _p(us), _p(first, C.c_int), 1, _p(R), _p(T), None, 0) == ERR_ARG
    assert b"at least 4" in lib.lfvio_last_error()
    first = np.array([1, 6], np.int32)
    assert lib.lfvio_epnp(_p(pw), _p(us), _p(first, C.c_int), 1, _p(R), _p(T), None, 0) == ERR_ARG
    assert lib.lfvio_epnp(None, None, None, 0, None, None, None, 0) == 0          # nothing to do


def test_global_sfm_rejects_malformed_tracks(lfvio_mod):
    lib = lfvio_mod.lib
    s = I.SfmSummary()
    rR, rT = np.eye(3).ravel().copy(), np.zeros(3)
    q, T = np.zeros((11, 4)), np.zeros((11, 3))
    pts = np.zeros((4, 3))

    def call(F, l, of, ofr):
        of, ofr = np.array(of, np.int32), np.array(ofr, np.int32)
        return lib.lfvio_global_sfm(F, l, _p(rR), _p(rT), _p(of, C.c_int), _p(ofr, C.c_int), _p(pts), len(of) - 1, 1, _p(q), _p(T), None, None,
                                    C.byref(s), 0)

    assert call(11, 0, [0, 2, 4], [0, 1, 3, 2]) == ERR_ARG and b"ascending" in lib.lfvio_last_error()
    assert call(11, 0, [0, 2, 4], [0, 1, 2, 11]) == ERR_ARG and b"outside the window" in lib.lfvio_last_error()
    assert call(11, 0, [1, 2, 4], [0, 1, 2, 3]) == ERR_ARG
    assert call(11, 10, [0, 2, 4], [0, 1, 2, 3]) == ERR_ARG                         # l must precede the newest frame
    assert call(64, 0, [0, 2, 4], [0, 1, 2, 3]) == ERR_ARG                          # beyond LFVIO_SFM_MAX_FRAMES


def test_alignment_and_relative_pose_argument_checks(lfvio_mod):
    lib = lfvio_mod.lib
    z = np.zeros(64)
    ok = C.c_int(0)
    assert lib.lfvio_linear_alignment(_p(z), _p(z), _p(z), _p(z), _p(z), 1, _p(z), 9.81, _p(z), _p(z), C.byref(ok), 0) == ERR_ARG
    assert lib.lfvio_linear_alignment(_p(z), _p(z), _p(z), _p(z), _p(z), 1000, _p(z), 9.81, _p(z), _p(z), C.byref(ok), 0) == ERR_ARG
    assert lib.lfvio_solve_gyro_bias(_p(z), _p(z), _p(z), 1, _p(z), 0) == ERR_ARG
    cnt = C.c_int(7)
    # fewer than 15 correspondences: the reference returns false without computing anything (solve_5pts.cpp:540)
    assert lib.lfvio_solve_relative_rt(_p(z), 10, 1, 0, _p(z), _p(z), C.byref(cnt), None, 0) == 0 and cnt.value == 0
