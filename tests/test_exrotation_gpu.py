"""B200 parity tests of the online extrinsic-rotation calibration (InitialEXRotation, ESTIMATE_EXTRINSIC == 2) through the
C-ABI and the C++ host mirror against oracle/initial.py."""
import ctypes as C
import os

import numpy as np
import pytest

from conftest import ROOT
from exrot_data import make_stream
from initial_data import InitScenario
from oracle import initial as OI

pytestmark = pytest.mark.gpu


def test_solve_relative_r(lfvio_mod, gpu):
    sc = InitScenario(seed=1, F=11)
    for k in (1, 4, 9):
        c = sc.corres(k - 1, k)
        got = lfvio_mod.solve_relative_r(c, 3, k)
        ref = OI.solve_relative_r(c, 3, k)
        assert np.allclose(got, ref, atol=1e-8)
        assert np.allclose(got, sc.Rc[k - 1].T @ sc.Rc[k], atol=1e-6)
    assert np.array_equal(lfvio_mod.solve_relative_r(sc.corres(0, 1)[:8], 3, 1), np.eye(3))     # below 9 correspondences


def test_calibrate_ex_rotation(lfvio_mod, gpu):
    ric_true, stream = make_stream(seed=2)
    Rc = [OI.solve_relative_r(c, 3, k) for k, (c, _) in enumerate(stream, 1)]
    Rimu = [OI._qrot(dq) for _, dq in stream]
    guess = np.eye(3)
    Rc_g = [guess.T @ Ri @ guess for Ri in Rimu]
    for n in (1, 3, len(stream)):
        ric, sv = lfvio_mod.calibrate_ex_rotation(Rc[:n], Rimu[:n], Rc_g[:n])
        ric_o, sv_o = OI.ex_rotation_solve(Rc[:n], Rimu[:n], Rc_g[:n])
        assert np.allclose(sv, sv_o, rtol=1e-9, atol=1e-12)
        if n > 1:
            assert np.allclose(ric, ric_o, atol=1e-8)
    assert np.abs(ric - ric_true).max() < 5e-3


def test_initial_ex_rotation_class(lfvio_mod, gpu):
    """The C++ mirror of InitialEXRotation over a rotation-rich stream.  The very first solve (one rotation pair) has a
    two-dimensional null space, so its ric — which enters Rc_g of the next pair and, through the Huber weight, every later
    solve — is an arbitrary member of that space in any implementation (Eigen, LAPACK, the device's Jacobi): the histories of
    two implementations are therefore compared on what is well defined — the first system's singular values, the success
    decision (frame_count >= WINDOW_SIZE and the second-smallest singular value > 0.25) and the calibrated rotation."""
    host = C.CDLL(os.path.join(ROOT, "lf-vio_b200", "liblfvio_host.so"))
    host.lfvio_host_ex_rotation_create.restype = C.c_void_p
    host.lfvio_host_ex_rotation_create.argtypes = [C.c_int, C.c_uint64, C.c_int]
    host.lfvio_host_ex_rotation_destroy.argtypes = [C.c_void_p]
    host.lfvio_host_ex_rotation_step.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
    ric_true, stream = make_stream(seed=0)
    h = host.lfvio_host_ex_rotation_create(10, 3, gpu)
    ref = OI.InitialEXRotationOracle(10)
    first_ok = None
    for k, (c, dq) in enumerate(stream, 1):
        c = np.ascontiguousarray(c)
        dq = np.ascontiguousarray(dq)
        ric, sv = np.zeros((3, 3)), np.zeros(4)
        rc = host.lfvio_host_ex_rotation_step(h, c.ctypes.data, len(c), dq.ctypes.data, ric.ctypes.data, sv.ctypes.data)
        assert rc >= 0, lfvio_mod.lib.lfvio_last_error()
        ok_o, ric_o = ref.calibrate(c, dq, 3)
        assert bool(rc) == ok_o, k
        if k == 1:
            _, sv_o = OI.ex_rotation_solve(ref.Rc, ref.Rimu, ref.Rc_g)
            assert np.allclose(sv, sv_o, rtol=1e-7, atol=1e-10)
        assert sv[0] >= sv[1] >= sv[2] >= sv[3] >= 0
        if rc:
            first_ok = first_ok or k
            assert np.abs(ric - ric_true).max() < 1e-2 and np.abs(ric - ric_o).max() < 1e-2
            assert np.allclose(ric @ ric.T, np.eye(3), atol=1e-12)
    host.lfvio_host_ex_rotation_destroy(h)
    assert first_ok == 10
