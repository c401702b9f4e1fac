"""Pins oracle/initial.py (the numpy restatement of the estimator initialisation) on synthetic
geometry with known answers: the reference has no tests or vectors for these functions and cannot be built here."""
import numpy as np
import pytest

from initial_data import InitScenario
from oracle import initial as OI
from oracle.prestep import _qrot


@pytest.fixture(scope="module")
def sc():
    return InitScenario(seed=1, F=11)


def test_relative_pose_recovers_the_true_motion(sc):
    for l in (0, 6):
        c = sc.corres(l, 10)
        ok, Rm, T, cnt, E = OI.solve_relative_rt(c, 7, l)
        Rtrue = sc.Rc[l].T @ sc.Rc[10]
        Ttrue = sc.Rc[l].T @ (sc.Pc[10] - sc.Pc[l])
        assert ok and cnt == len(c)
        assert np.allclose(Rm, Rtrue, atol=1e-7)
        assert np.allclose(T / np.linalg.norm(T), Ttrue / np.linalg.norm(Ttrue), atol=1e-6)
        # epipolar constraint of the refit essential matrix: b2^T E b1 = 0
        assert np.abs(np.einsum("ni,ij,nj->n", c[:, 3:], E, c[:, :3])).max() < 1e-7


def test_decompose_essential_candidates_are_rotations(sc):
    _, _, _, _, E = OI.solve_relative_rt(sc.corres(0, 10), 7, 0)
    R1, R2, t = OI.decompose_essential(E)
    for Rm in (R1, R2):
        assert np.allclose(Rm @ Rm.T, np.eye(3), atol=1e-12) and abs(np.linalg.det(Rm) - 1) < 1e-12
    assert abs(np.linalg.norm(t) - 1) < 1e-12


def test_triangulate_point_exact(sc):
    rng = np.random.default_rng(0)
    P0 = np.hstack([sc.Rc[2].T, (-sc.Rc[2].T @ sc.Pc[2]).reshape(3, 1)])
    P1 = np.hstack([sc.Rc[9].T, (-sc.Rc[9].T @ sc.Pc[9]).reshape(3, 1)])
    for _ in range(10):
        X = rng.uniform(-5, 5, 3) + sc.Pc[2]
        b0, b1 = P0[:, :3] @ X + P0[:, 3], P1[:, :3] @ X + P1[:, 3]
        got = OI.triangulate_point(P0, P1, b0 / np.linalg.norm(b0), b1 / np.linalg.norm(b1))
        assert np.allclose(got, X, atol=1e-9)


def test_epnp_exact_on_noise_free_bearings_including_the_negative_half_space(sc):
    rng = np.random.default_rng(1)
    pw = rng.uniform(-4, 4, (60, 3))
    Rt, Tt = sc.Rc[4].T, np.array([0.3, -0.2, 0.1])
    pc = pw @ Rt.T + Tt
    us = pc / np.linalg.norm(pc, axis=1, keepdims=True)
    assert (us[:, 2] < 0).sum() > 10                                   # PAL bearings below the horizon
    Rm, T, err = OI.epnp_compute_pose(pw, us)
    assert np.allclose(Rm, Rt, atol=1e-9) and np.allclose(T, Tt, atol=1e-9)


def test_global_sfm_recovers_structure_up_to_scale(sc):
    l = 0
    ok, Rr, Tr, _, _ = OI.solve_relative_rt(sc.corres(l, 10), 7, 0)
    res = OI.global_sfm(11, l, Rr, Tr, sc.obs_first, sc.obs_frame, sc.obs_pts)
    assert res["ok"] and res["ba"]["converged"]
    s = np.linalg.norm(sc.Pc[10] - sc.Pc[l]) / np.linalg.norm(Tr)
    for i in range(11):
        assert np.allclose(_qrot(res["q"][i]), sc.Rc[l].T @ sc.Rc[i], atol=1e-6)
        assert np.allclose(res["T"][i] * s, sc.Rc[l].T @ (sc.Pc[i] - sc.Pc[l]), atol=1e-5)
    assert res["state"].sum() >= 150


def test_bundle_adjust_reduces_a_perturbed_problem_to_zero(sc):
    l = 0
    ok, Rr, Tr, _, _ = OI.solve_relative_rt(sc.corres(l, 10), 7, 0)
    res = OI.global_sfm(11, l, Rr, Tr, sc.obs_first, sc.obs_frame, sc.obs_pts, ba=False)
    cr, ct, pts = [a.copy() for a in res["pre_ba"]]
    pts += np.random.default_rng(0).normal(0, 0.02, pts.shape)
    s = OI.bundle_adjust(cr, ct, pts, res["state"], sc.obs_first, sc.obs_frame, sc.obs_pts, l)
    assert s["converged"] and s["costs"][0] > 1e-3 and s["final_cost"] < 1e-10
    assert all(b < a for a, b in zip(s["costs"], s["costs"][1:]))      # monotone: only successful steps are recorded
    assert np.allclose(cr[l], [1, 0, 0, 0]) and np.allclose(ct[l], 0)  # the constant blocks did not move
    assert np.allclose(ct[10], res["pre_ba"][1][10])


def test_visual_imu_alignment_recovers_scale_gravity_and_bias(sc):
    l, s_true = 0, 2.5
    Rs = np.array([sc.Rc[l].T @ sc.Rb[k] for k in range(11)])
    Ts = np.array([sc.Rc[l].T @ (sc.Pc[k] - sc.Pc[l]) for k in range(11)]) / s_true
    pre = sc.preintegrations()
    dbg = OI.solve_gyroscope_bias(Rs, [p.delta_q for p in pre], [p.jacobian[3:6, 12:15] for p in pre])
    assert np.allclose(dbg, sc.world.bg, atol=3e-3)                    # gyroscope noise 0.004 rad/s/sqrt(Hz) over one second
    pre = sc.preintegrations(bg=dbg)
    ok, g, x = OI.linear_alignment(Rs, Ts, [p.delta_p for p in pre], [p.delta_v for p in pre], [p.sum_dt for p in pre], sc.tic, sc.g_norm)
    assert ok and abs(x[-1] - s_true) < 0.1
    assert abs(np.linalg.norm(g) - sc.g_norm) < 1e-9
    assert np.allclose(g, sc.Rc[l].T @ sc.world.G, atol=0.15)
    assert np.allclose(Rs[0] @ x[0:3], sc.Rc[l].T @ sc.V[0], atol=0.05)


def test_initial_structure_end_to_end(sc):
    r = OI.initial_structure(11, sc.obs_first, sc.obs_frame, sc.obs_pts, sc.corres, lambda bg: sc.preintegrations(bg=bg), sc.ric,
                             sc.tic, sc.g_norm, np.zeros(3), seed=5)
    assert r is not None
    Rz = OI._yaw_rot(-OI._R2ypr_yaw_deg(sc.Rb[0]))
    assert np.abs(r["Ps"] - (sc.P - sc.P[0]) @ Rz.T).max() < 0.02      # metres, over a 1 s / ~1 m window
    assert np.abs(r["Vs"] - sc.V @ Rz.T).max() < 0.05
    assert np.allclose(r["g"], [0, 0, sc.g_norm], atol=1e-9)
    assert abs(OI._R2ypr_yaw_deg(r["Rs"][0])) < 1e-9                   # yaw of frame 0 removed


def test_extrinsic_rotation_calibration_recovers_ric():
    """InitialEXRotation restatement on a rotation-rich synthetic stream: succeeds exactly when frame_count reaches
    WINDOW_SIZE (the singular-value test is already satisfied by then) and returns the true camera-IMU rotation."""
    from exrot_data import make_stream

    ric_true, stream = make_stream(seed=0)
    ex = OI.InitialEXRotationOracle(10)
    oks = []
    for c, dq in stream:
        ok, ric = ex.calibrate(c, dq, 3)
        oks.append(ok)
    assert oks == [False] * 9 + [True] * (len(stream) - 9)
    assert np.abs(ric - ric_true).max() < 5e-3
    # the camera rotation of one pair, against the construction
    c, dq = stream[4]
    Rc = OI.solve_relative_r(c, 3, 5)
    Rimu = _qrot(np.asarray(dq))
    assert np.abs(ric_true @ Rc @ ric_true.T - Rimu).max() < 5e-3          # R_imu = ric R_c ric^T up to the injected gyro error
    assert np.array_equal(OI.solve_relative_r(c[:8], 3, 5), np.eye(3))     # below 9 correspondences: identity
