"""B200 parity tests of the estimator initialisation through the C-ABI against oracle/initial.py.

Tolerances: everything is fp64 on both sides; differences come from summation order (block reductions, atomics) and from
Jacobi vs LAPACK decompositions, so direct solves agree to ~1e-9 relative; iterated results (BA) are compared on the cost
sequence (1e-6 relative) and on the converged state."""
import numpy as np
import pytest

from initial_data import InitScenario
from oracle import initial as OI

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def sc():
    return InitScenario(seed=1, F=11)


@pytest.fixture(scope="module")
def sc_noisy():
    return InitScenario(seed=3, F=11, pixel_noise=0.3)


def _alignment_inputs(sc, l=0, s_true=2.5, bg=None):
    Rs = np.array([sc.Rc[l].T @ sc.Rb[k] for k in range(sc.F)])
    Ts = np.array([sc.Rc[l].T @ (sc.Pc[k] - sc.Pc[l]) for k in range(sc.F)]) / s_true
    return Rs, Ts, sc.preintegrations(bg=bg)


def test_gyro_bias(lfvio_mod, gpu, sc):
    Rs, _, pre = _alignment_inputs(sc)
    dq = [p.delta_q for p in pre]
    J = [p.jacobian[3:6, 12:15] for p in pre]
    got = lfvio_mod.solve_gyro_bias(Rs, dq, J)
    ref = OI.solve_gyroscope_bias(Rs, dq, J)
    assert np.allclose(got, ref, rtol=1e-9, atol=1e-13)
    assert np.allclose(got, sc.world.bg, atol=3e-3)


@pytest.mark.parametrize("F", [11, 21, 40])
def test_linear_alignment(lfvio_mod, gpu, F):
    sc = InitScenario(seed=2, F=F)
    Rs, Ts, pre0 = _alignment_inputs(sc)
    dbg = OI.solve_gyroscope_bias(Rs, [p.delta_q for p in pre0], [p.jacobian[3:6, 12:15] for p in pre0])
    pre = sc.preintegrations(bg=dbg)
    args = (Rs, Ts, [p.delta_p for p in pre], [p.delta_v for p in pre], [p.sum_dt for p in pre], sc.tic, sc.g_norm)
    ok, g, x = lfvio_mod.linear_alignment(*args)
    ok_o, g_o, x_o = OI.linear_alignment(*args)
    assert ok == ok_o and ok
    assert np.allclose(g, g_o, rtol=1e-8, atol=1e-8)
    assert np.allclose(x, x_o, rtol=1e-7, atol=1e-8)
    assert abs(x[-1] - 2.5) < 0.15                                   # the scale that was divided out
    assert abs(np.linalg.norm(g) - sc.g_norm) < 1e-9


def test_linear_alignment_rejects(lfvio_mod, gpu, sc):
    """Wrong-handed input (negated translations) gives a negative scale: both arms return false (:186-194)."""
    Rs, Ts, pre = _alignment_inputs(sc)
    args = (Rs, -Ts, [p.delta_p for p in pre], [p.delta_v for p in pre], [p.sum_dt for p in pre], sc.tic, sc.g_norm)
    ok, g, x = lfvio_mod.linear_alignment(*args)
    ok_o, g_o, x_o = OI.linear_alignment(*args)
    assert ok == ok_o and not ok
    assert np.allclose(g, g_o, rtol=1e-8, atol=1e-8)


@pytest.mark.parametrize("l", [0, 4, 8])
def test_solve_relative_rt(lfvio_mod, gpu, sc, l):
    c = sc.corres(l, sc.F - 1)
    ok, Rm, T, cnt, E = lfvio_mod.solve_relative_rt(c, 7, l)
    ok_o, R_o, T_o, cnt_o, E_o = OI.solve_relative_rt(c, 7, l)
    assert ok == ok_o and cnt == cnt_o
    s = np.sign(np.sum(E * E_o))                                      # a null vector's sign is free
    assert np.allclose(E, s * E_o, atol=1e-8)
    assert np.allclose(Rm, R_o, atol=1e-8) and np.allclose(T, T_o, atol=1e-8)
    Rtrue = sc.Rc[l].T @ sc.Rc[-1]
    Ttrue = sc.Rc[l].T @ (sc.Pc[-1] - sc.Pc[l])
    assert np.allclose(Rm, Rtrue, atol=1e-6)
    assert np.allclose(T / np.linalg.norm(T), Ttrue / np.linalg.norm(Ttrue), atol=1e-5)


def test_solve_relative_rt_too_few(lfvio_mod, gpu, sc):
    ok, _, _, cnt, _ = lfvio_mod.solve_relative_rt(sc.corres(0, 10)[:14], 7, 0)
    assert not ok and cnt == 0


def test_epnp_batched(lfvio_mod, gpu, sc):
    rng = np.random.default_rng(5)
    problems, truth = [], []
    for k, n in enumerate([6, 10, 37, 150, 400]):
        pw = rng.uniform(-4, 4, (n, 3)) + [0, 0, 1]
        Rt = sc.Rc[k + 2].T
        Tt = -Rt @ sc.Pc[k + 2] * 0.1
        pc = pw @ Rt.T + Tt
        us = pc / np.linalg.norm(pc, axis=1, keepdims=True)
        if k % 2:
            us = us + rng.normal(0, 1e-3, us.shape)
            us /= np.linalg.norm(us, axis=1, keepdims=True)
        problems.append((pw, us))
        truth.append((Rt, Tt))
    Rm, T, err = lfvio_mod.epnp(problems)
    for k, (pw, us) in enumerate(problems):
        R_o, T_o, e_o = OI.epnp_compute_pose(pw, us)
        assert np.allclose(Rm[k], R_o, atol=1e-8) and np.allclose(T[k], T_o, atol=1e-8), k
        assert abs(err[k] - e_o) <= 1e-8 * max(1.0, e_o)
        if k % 2 == 0:
            assert np.allclose(Rm[k], truth[k][0], atol=1e-8) and np.allclose(T[k], truth[k][1], atol=1e-8)


@pytest.mark.parametrize("noise", [0.0, 0.3])
def test_bundle_adjust(lfvio_mod, gpu, noise):
    sc = InitScenario(seed=2, F=11, pixel_noise=noise)
    l = 0
    ok, Rr, Tr, _, _ = OI.solve_relative_rt(sc.corres(l, 10), 7, 0)
    res = OI.global_sfm(11, l, Rr, Tr, sc.obs_first, sc.obs_frame, sc.obs_pts, ba=False)
    cr, ct, pts = [np.ascontiguousarray(a) for a in res["pre_ba"]]
    pts += np.random.default_rng(0).normal(0, 0.02, pts.shape)
    cr2, ct2, pts2 = cr.copy(), ct.copy(), pts.copy()
    so = OI.bundle_adjust(cr, ct, pts, res["state"], sc.obs_first, sc.obs_frame, sc.obs_pts, l)
    s = lfvio_mod.sfm_bundle_adjust(cr2, ct2, pts2, res["state"], sc.obs_first, sc.obs_frame, sc.obs_pts, l)
    assert s.ba_iterations == so["iterations"] and bool(s.ba_converged) == so["converged"]
    assert np.allclose(s.costs(), so["costs"], rtol=1e-6, atol=1e-18)
    st = res["state"]
    assert np.allclose(cr2, cr, atol=1e-7) and np.allclose(ct2, ct, atol=1e-7)
    assert np.allclose(pts2[st], pts[st], atol=1e-4 if noise else 1e-8)


@pytest.mark.parametrize("which", ["clean", "noisy"])
@pytest.mark.parametrize("l", [0, 5])
def test_global_sfm(lfvio_mod, gpu, sc, sc_noisy, which, l):
    s = sc if which == "clean" else sc_noisy
    ok, Rr, Tr, cnt, _ = OI.solve_relative_rt(s.corres(l, 10), 11, 0)
    assert ok
    ref = OI.global_sfm(11, l, Rr, Tr, s.obs_first, s.obs_frame, s.obs_pts)
    got = lfvio_mod.global_sfm(11, l, Rr, Tr, s.obs_first, s.obs_frame, s.obs_pts)
    assert got["ok"] == ref["ok"]
    if not ref["ok"]:
        return
    assert np.array_equal(got["state"], ref["state"])
    tol = 1e-7 if which == "clean" else 1e-5
    assert np.allclose(got["q"], ref["q"], atol=tol) and np.allclose(got["T"], ref["T"], atol=tol)
    st = ref["state"]
    assert np.allclose(got["points"][st], ref["points"][st], rtol=1e-4, atol=1e-4)
    if which == "clean":                                              # and against the ground truth, up to scale
        sc_ = np.linalg.norm(s.Pc[10] - s.Pc[l]) / np.linalg.norm(Tr)
        for i in range(11):
            assert np.allclose(OI._qrot(got["q"][i]), s.Rc[l].T @ s.Rc[i], atol=1e-6)
            assert np.allclose(got["T"][i] * sc_, s.Rc[l].T @ (s.Pc[i] - s.Pc[l]), atol=1e-5)


def test_global_sfm_before_ba(lfvio_mod, gpu, sc_noisy):
    """The PnP / triangulation chain alone (run_ba = 0) against the oracle's pre-BA state."""
    s, l = sc_noisy, 3
    ok, Rr, Tr, _, _ = OI.solve_relative_rt(s.corres(l, 10), 11, 0)
    ref = OI.global_sfm(11, l, Rr, Tr, s.obs_first, s.obs_frame, s.obs_pts, ba=False)
    got = lfvio_mod.global_sfm(11, l, Rr, Tr, s.obs_first, s.obs_frame, s.obs_pts, ba=False)
    assert got["ok"] and ref["ok"] and np.array_equal(got["state"], ref["state"])
    assert np.allclose(got["q"], ref["q"], atol=1e-7) and np.allclose(got["T"], ref["T"], atol=1e-7)
    st = ref["state"]
    assert np.allclose(got["points"][st], ref["points"][st], rtol=1e-6, atol=1e-6)


def test_global_sfm_pnp_failure(lfvio_mod, gpu, sc):
    """Fewer than 10 triangulated points in a frame: construct() returns false (initial_sfm.cpp:50-55)."""
    keep = 8
    of = sc.obs_first[:keep + 1]
    ok, Rr, Tr, _, _ = OI.solve_relative_rt(sc.corres(0, 10), 11, 0)
    ref = OI.global_sfm(11, 0, Rr, Tr, of, sc.obs_frame[:of[-1]], sc.obs_pts[:of[-1]])
    got = lfvio_mod.global_sfm(11, 0, Rr, Tr, of, sc.obs_frame[:of[-1]], sc.obs_pts[:of[-1]])
    assert not ref["ok"] and not got["ok"]
    assert got["summary"].fail_stage == 1 and got["summary"].fail_frame == ref["frame"]


def test_global_sfm_config4_window(lfvio_mod, gpu):
    """A 21-frame window (BASELINE config 4's WINDOW_SIZE = 20: reduced camera system 126 x 126 in shared memory)."""
    s, l = InitScenario(seed=3, F=21, pixel_noise=0.05), 12
    ok, Rr, Tr, cnt, _ = OI.solve_relative_rt(s.corres(l, 20), 11, 0)
    assert ok
    ref = OI.global_sfm(21, l, Rr, Tr, s.obs_first, s.obs_frame, s.obs_pts)
    got = lfvio_mod.global_sfm(21, l, Rr, Tr, s.obs_first, s.obs_frame, s.obs_pts)
    assert ref["ok"] and got["ok"] and np.array_equal(got["state"], ref["state"])
    assert got["summary"].ba_iterations == ref["ba"]["iterations"]
    assert np.allclose(got["summary"].costs(), ref["ba"]["costs"], rtol=1e-5, atol=1e-15)
    assert np.allclose(got["q"], ref["q"], atol=1e-6) and np.allclose(got["T"], ref["T"], atol=1e-6)
