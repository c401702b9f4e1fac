"""The __host__ __device__ cores of the initialisation kernels (lf-vio_b200/csrc/init_math.cuh), compiled for the host by
tests/native/init_math_shim.cpp, against the numpy oracle (oracle/initial.py).  CPU only: this checks the arithmetic the
kernels run; the kernels themselves are compared with the same oracle in tests/test_initial_gpu.py."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from initial_data import InitScenario
from oracle import initial as OI

HERE = os.path.dirname(os.path.abspath(__file__))
_dp = C.POINTER(C.c_double)


def _p(a):
    return a.ctypes.data_as(_dp)


@pytest.fixture(scope="module")
def shim():
    out = os.path.join(HERE, "native", "_build")
    os.makedirs(out, exist_ok=True)
    so = os.path.join(out, "libinit_math_shim.so")
    subprocess.run(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-o", so, os.path.join(HERE, "native", "init_math_shim.cpp")],
                   check=True)
    lib = C.CDLL(so)
    lib.shim_epnp.restype = C.c_double
    return lib


@pytest.fixture(scope="module")
def sc():
    return InitScenario(seed=1, F=11)


def test_triangulate_point(shim, sc):
    rng = np.random.default_rng(0)
    for _ in range(20):
        i, j = sorted(rng.choice(11, 2, replace=False))
        P0 = np.hstack([sc.Rc[i].T, (-sc.Rc[i].T @ sc.Pc[i]).reshape(3, 1)])
        P1 = np.hstack([sc.Rc[j].T, (-sc.Rc[j].T @ sc.Pc[j]).reshape(3, 1)])
        X = rng.uniform(-5, 5, 3) + sc.Pc[i]
        b0 = P0[:, :3] @ X + P0[:, 3]
        b1 = P1[:, :3] @ X + P1[:, 3]
        b0, b1 = b0 / np.linalg.norm(b0), b1 / np.linalg.norm(b1)
        out = np.zeros(3)
        shim.shim_triangulate_point(_p(np.ascontiguousarray(P0)), _p(np.ascontiguousarray(P1)), _p(b0), _p(b1), _p(out))
        ref = OI.triangulate_point(P0, P1, b0, b1)
        assert np.allclose(out, ref, rtol=1e-9, atol=1e-9)
        assert np.allclose(out, X, rtol=1e-8, atol=1e-8)


def test_svd3_and_recover_pose(shim, sc):
    rng = np.random.default_rng(1)
    M = rng.normal(size=(3, 3))
    U, s, V = np.zeros((3, 3)), np.zeros(3), np.zeros((3, 3))
    shim.shim_svd3(_p(M), _p(U), _p(s), _p(V))
    assert np.allclose(U @ np.diag(s) @ V.T, M, atol=1e-13)
    assert np.allclose(s, np.linalg.svd(M)[1], atol=1e-13)
    c = sc.corres(0, 10)
    ok, Rr, Tr, cnt, E = OI.solve_relative_rt(c, 7, 0)
    b1, b2 = np.ascontiguousarray(c[:, :3]), np.ascontiguousarray(c[:, 3:])
    Ro, to, cnt_o, counts_o = OI.recover_pose(E, b1, b2)
    R, t, counts = np.zeros((3, 3)), np.zeros(3), np.zeros(4, np.int32)
    shim.shim_recover_pose(_p(np.ascontiguousarray(E)), _p(b1), _p(b2), len(c), _p(R), _p(t), counts.ctypes.data_as(C.POINTER(C.c_int)))
    assert sorted(counts.tolist()) == sorted(counts_o)       # the SVD sign freedom permutes the four candidates
    assert np.allclose(R, Ro, atol=1e-9) and np.allclose(t, to, atol=1e-9)


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_epnp(shim, sc, seed):
    rng = np.random.default_rng(seed)
    n = 12 + 40 * seed
    pw = rng.uniform(-4, 4, (n, 3)) + [0, 0, 1]
    Rt = sc.Rc[3 + seed].T
    Tt = -Rt @ sc.Pc[3 + seed] * 0.1
    pc = pw @ Rt.T + Tt
    us = pc / np.linalg.norm(pc, axis=1, keepdims=True)
    us = us + rng.normal(0, 1e-3 * seed, us.shape)          # seed 0: noise free
    us = us / np.linalg.norm(us, axis=1, keepdims=True)
    Ro, To, eo = OI.epnp_compute_pose(pw, us)
    R, T = np.zeros((3, 3)), np.zeros(3)
    e = shim.shim_epnp(_p(np.ascontiguousarray(pw)), _p(np.ascontiguousarray(us)), n, _p(R), _p(T))
    assert np.allclose(R, Ro, atol=1e-8) and np.allclose(T, To, atol=1e-8)
    assert abs(e - eo) <= 1e-8 * max(1.0, eo)
    if seed == 0:
        assert np.allclose(R, Rt, atol=1e-9) and np.allclose(T, Tt, atol=1e-9)


def test_alignment_system_and_gyro_bias(shim, sc):
    l = 0
    Rs = np.ascontiguousarray([sc.Rc[l].T @ sc.Rb[k] for k in range(11)])
    Ts = np.ascontiguousarray([sc.Rc[l].T @ (sc.Pc[k] - sc.Pc[l]) for k in range(11)]) / 2.5
    pre = sc.preintegrations()
    dq = np.ascontiguousarray([p.delta_q for p in pre])
    Jrg = np.ascontiguousarray([p.jacobian[3:6, 12:15] for p in pre])
    A3, b3 = np.zeros((3, 3)), np.zeros(3)
    shim.shim_gyro_bias(_p(Rs), _p(dq), _p(Jrg), 11, _p(A3), _p(b3))
    assert np.allclose(np.linalg.solve(A3, b3), OI.solve_gyroscope_bias(Rs, dq, Jrg), rtol=1e-9, atol=1e-12)
    dp = np.ascontiguousarray([p.delta_p for p in pre])
    dv = np.ascontiguousarray([p.delta_v for p in pre])
    dts = np.ascontiguousarray([p.sum_dt for p in pre])
    tic = np.ascontiguousarray(sc.tic)
    for gcols, gfixed in ((np.eye(3), np.zeros(3)), (OI.tangent_basis(np.array([0.3, -0.2, 9.7])), np.array([0.3, -0.2, 9.7]))):
        k = gcols.shape[1]
        ns = 33 + k + 1
        A, b = np.zeros((ns, ns)), np.zeros(ns)
        shim.shim_alignment_system(_p(Rs), _p(Ts), _p(dp), _p(dv), _p(dts), 11, _p(tic), _p(np.ascontiguousarray(gcols)), k,
                                   _p(np.ascontiguousarray(gfixed)), _p(A), _p(b))
        Ao, bo = OI._alignment_system(Rs, Ts, dp, dv, dts, sc.tic, gcols, gfixed)
        assert np.allclose(A, Ao, rtol=1e-12, atol=1e-9) and np.allclose(b, bo, rtol=1e-12, atol=1e-9)
    bc = np.zeros((3, 2))
    shim.shim_tangent_basis(_p(np.array([0.3, -0.2, 9.7])), _p(bc))
    assert np.allclose(bc, OI.tangent_basis(np.array([0.3, -0.2, 9.7])), atol=1e-15)


def test_ba_residual_jacobian(shim):
    rng = np.random.default_rng(3)
    for _ in range(5):
        q = rng.normal(size=4)
        q /= np.linalg.norm(q)
        t, p = rng.normal(size=3), rng.normal(size=3) * 3
        obs = rng.normal(size=3)
        obs /= np.linalg.norm(obs)
        r, Jr, Jt, Jp = np.zeros(3), np.zeros((3, 3)), np.zeros((3, 3)), np.zeros((3, 3))
        shim.shim_ba_residual_jac(_p(q), _p(t), _p(p), _p(obs), _p(r), _p(Jr), _p(Jt), _p(Jp))
        ro, Jro, Jto, Jpo = OI._ba_residual_jac(q, t, p, obs)
        assert np.allclose(r, ro, atol=1e-14) and np.allclose(Jr, Jro, atol=1e-13)
        assert np.allclose(Jt, Jto, atol=1e-13) and np.allclose(Jp, Jpo, atol=1e-13)
        # finite differences through the Ceres parameterisation (this also pins the oracle's analytic Jacobian)
        eps = 1e-6
        for k in range(3):
            d = np.zeros(3)
            d[k] = eps
            qp = np.zeros(4)
            shim.shim_quat_plus(_p(q), _p(d), _p(qp))
            assert np.allclose(qp, OI._quat_plus(q, d), atol=1e-15)
            fd = (OI._ba_residual_jac(qp, t, p, obs, False)[0] - ro) / eps
            assert np.allclose(fd, Jr[:, k], atol=1e-5)
            fd = (OI._ba_residual_jac(q, t + d, p, obs, False)[0] - ro) / eps
            assert np.allclose(fd, Jt[:, k], atol=1e-5)
            fd = (OI._ba_residual_jac(q, t, p + d, obs, False)[0] - ro) / eps
            assert np.allclose(fd, Jp[:, k], atol=1e-5)


def test_ex_rotation_solve(shim):
    """CalibrationExRotation's stacked system (initial_ex_rotation.cpp:20-60): device core vs numpy, and the known answer."""
    sc = InitScenario(seed=1, F=14)
    pre = sc.preintegrations(bg=sc.world.bg)
    Rc = np.ascontiguousarray([sc.Rc[k - 1].T @ sc.Rc[k] for k in range(1, 14)])          # exact camera rotations
    Rimu = np.ascontiguousarray([OI._qrot(np.asarray(p.delta_q)) for p in pre])
    ric_guess = np.eye(3)
    Rc_g = np.ascontiguousarray([ric_guess.T @ Rimu[i] @ ric_guess for i in range(13)])
    for N in (1, 5, 13):
        ric, sv = np.zeros((3, 3)), np.zeros(4)
        shim.shim_ex_rotation_solve(N, _p(Rc), _p(Rimu), _p(Rc_g), _p(ric), _p(sv))
        ric_o, sv_o = OI.ex_rotation_solve(Rc[:N], Rimu[:N], Rc_g[:N])
        assert np.allclose(sv, sv_o, rtol=1e-9, atol=1e-12)
        if N > 1:
            assert np.allclose(ric, ric_o, atol=1e-7)
    assert np.abs(ric - sc.ric).max() < 0.02                                               # gyro noise / bias limited


def test_cores_terminate_on_degenerate_input(shim):
    """The fixed-size cores must come back (possibly with non-finite numbers, as Eigen would) on degenerate input — they run
    inside single-CTA kernels where a non-terminating loop would hang the device: coplanar world points (singular
    control-point basis), a bearing on the image plane's horizon (division by z = 0), an all-zero essential matrix."""
    rng = np.random.default_rng(0)
    pw = np.ascontiguousarray(np.hstack([rng.uniform(-1, 1, (20, 2)), np.zeros((20, 1))]))        # coplanar
    pc = pw + [0, 0, 4.0]
    us = np.ascontiguousarray(pc / np.linalg.norm(pc, axis=1, keepdims=True))
    R, T = np.zeros((3, 3)), np.zeros(3)
    shim.shim_epnp(_p(pw), _p(us), 20, _p(R), _p(T))
    us2 = us.copy()
    us2[3] = [1.0, 0.0, 0.0]                                                                       # z = 0
    shim.shim_epnp(_p(np.ascontiguousarray(pw + rng.normal(0, 0.3, pw.shape))), _p(us2), 20, _p(R), _p(T))
    E = np.zeros((3, 3))
    b = np.ascontiguousarray(us)
    Rr, t, counts = np.zeros((3, 3)), np.zeros(3), np.zeros(4, np.int32)
    shim.shim_recover_pose(_p(E), _p(b), _p(b), 20, _p(Rr), _p(t), counts.ctypes.data_as(C.POINTER(C.c_int)))
    X = np.zeros(3)
    P0 = np.ascontiguousarray(np.hstack([np.eye(3), np.zeros((3, 1))]))
    shim.shim_triangulate_point(_p(P0), _p(P0), _p(b[0]), _p(b[0]), _p(X))                         # zero baseline
    U, s, V = np.zeros((3, 3)), np.zeros(3), np.zeros((3, 3))
    shim.shim_svd3(_p(np.zeros((3, 3))), _p(U), _p(s), _p(V))
    assert np.all(s == 0)


def test_ba_control_flow_terminates_on_non_finite_cost(shim):
    """A point sitting on a camera centre makes the residual non-finite; the LM loop (ba_begin_iteration / ba_decide — the same
    functions the kernel runs) must still stop, at the iteration cap at the latest, and report non-convergence."""
    sc = InitScenario(seed=2, F=6, max_feat=40)
    ok, Rr, Tr, _, _ = OI.solve_relative_rt(sc.corres(0, 5), 7, 0)
    res = OI.global_sfm(6, 0, Rr, Tr, sc.obs_first, sc.obs_frame, sc.obs_pts, ba=False)
    cr, ct, pts = [np.ascontiguousarray(a) for a in res["pre_ba"]]
    state = res["state"].astype(np.uint8)
    j = int(np.argmax(state))
    pts[j] = -OI._qrot(cr[sc.obs_frame[sc.obs_first[j]]] * [1, -1, -1, -1]) @ ct[sc.obs_frame[sc.obs_first[j]]]   # camera centre
    costs, ns, cv = np.zeros(51), C.c_int(), C.c_int()
    of, ofr = np.ascontiguousarray(sc.obs_first, np.int32), np.ascontiguousarray(sc.obs_frame, np.int32)
    ip = C.POINTER(C.c_int)
    it = shim.shim_ba(6, len(state), 0, of.ctypes.data_as(ip), ofr.ctypes.data_as(ip), _p(sc.obs_pts), state.ctypes.data_as(C.POINTER(C.c_ubyte)),
                      _p(cr), _p(ct), _p(pts), _p(costs), C.byref(ns), C.byref(cv))
    assert 0 <= it <= 50
