"""CPU checks that pin oracle/prestep.py (no reference vectors exist for these functions — see its header):
closed-form pre-integration, finite differences of the bias Jacobians, covariance properties, agreement with the
independent restatement used by the synthetic-window generator, and noise-free triangulation geometry."""
import numpy as np

from lfvio import synthwin
from oracle import prestep as O
from prestep_data import G, NOISE, imu_stream, tracks


def _integrate(stream, i, ba=None, bg=None):
    first, count, a0, g0 = stream["intervals"][i]
    ib = O.IntegrationOracle(a0, g0, stream["ba"] if ba is None else ba, stream["bg"] if bg is None else bg, NOISE)
    for k in range(first, first + count):
        ib.push_back(stream["dt"][k], stream["acc"][k], stream["gyr"][k])
    return ib


def test_constant_acceleration_closed_form():
    ib = O.IntegrationOracle([1.0, -2.0, 0.5], [0, 0, 0], np.zeros(3), np.zeros(3), NOISE)
    for _ in range(20):
        ib.push_back(0.005, [1.0, -2.0, 0.5], [0, 0, 0])
    T = 0.1
    assert abs(ib.sum_dt - T) < 1e-15
    assert np.allclose(ib.delta_v, np.array([1.0, -2.0, 0.5]) * T, atol=1e-14)
    assert np.allclose(ib.delta_p, 0.5 * np.array([1.0, -2.0, 0.5]) * T * T, atol=1e-14)
    assert np.allclose(ib.delta_q, [1, 0, 0, 0], atol=1e-15)
    # d delta_p / d ba = -T^2/2 I, d delta_v / d ba = -T I for a non-rotating body
    assert np.allclose(ib.jacobian[0:3, 9:12], -0.5 * T * T * np.eye(3), atol=1e-14)
    assert np.allclose(ib.jacobian[6:9, 9:12], -T * np.eye(3), atol=1e-14)


def test_constant_rotation_rate():
    w = np.array([0.0, 0.0, 0.7])
    ib = O.IntegrationOracle([0, 0, 0], w, np.zeros(3), np.zeros(3), NOISE)
    for _ in range(40):
        ib.push_back(0.0025, [0, 0, 0], w)
    ang = 2 * np.arctan2(np.linalg.norm(ib.delta_q[1:]), ib.delta_q[0])
    assert abs(ang - 0.07) < 1e-7 and abs(np.linalg.norm(ib.delta_q) - 1) < 1e-15  # first-order quaternion steps
    Jq = ib.jacobian[3:6, 12:15]  # -int R^T dt: exact about the spin axis, a small rotation in the plane
    assert abs(Jq[2, 2] + 0.1) < 1e-12 and abs(Jq[0, 0] + 0.1) < 1e-3 and abs(Jq[0, 1] + Jq[1, 0]) < 1e-12


def test_bias_jacobians_by_finite_differences():
    s = imu_stream(3)
    ib = _integrate(s, 2)
    eps = 1e-6
    for a in range(3):
        e = np.zeros(3)
        e[a] = eps
        pa, pg = _integrate(s, 2, ba=s["ba"] + e), _integrate(s, 2, bg=s["bg"] + e)
        assert np.allclose((pa.delta_p - ib.delta_p) / eps, ib.jacobian[0:3, 9 + a], atol=1e-6)
        assert np.allclose((pa.delta_v - ib.delta_v) / eps, ib.jacobian[6:9, 9 + a], atol=1e-6)
        assert np.allclose((pg.delta_p - ib.delta_p) / eps, ib.jacobian[0:3, 12 + a], atol=1e-5)
        # (the reference's discrete d/dbg blocks drop second-order terms of the mid-point rule: agreement to ~4e-5)
        assert np.allclose((pg.delta_v - ib.delta_v) / eps, ib.jacobian[6:9, 12 + a], atol=2e-4)
        dq = O._qmul(ib.delta_q * [1, -1, -1, -1], pg.delta_q)  # delta_q^-1 * delta_q(bg + e)
        assert np.allclose(2 * dq[1:] / eps, ib.jacobian[3:6, 12 + a], atol=2e-4)


def test_covariance_symmetric_psd_and_growing():
    s = imu_stream(4)
    ib = _integrate(s, 0)
    P = ib.covariance
    assert np.abs(P - P.T).max() < 1e-18 + 1e-12 * np.abs(P).max()
    ev = np.linalg.eigvalsh(0.5 * (P + P.T))
    assert ev.min() > -1e-18 and ev.max() > 0
    first, count, a0, g0 = s["intervals"][0]
    half = O.IntegrationOracle(a0, g0, s["ba"], s["bg"], NOISE)
    for k in range(first, first + count // 2):
        half.push_back(s["dt"][k], s["acc"][k], s["gyr"][k])
    assert np.all(np.diag(P) >= np.diag(half.covariance))


def test_matches_the_generator_restatement():
    s = imu_stream(5)
    first, count, a0, g0 = s["intervals"][1]
    ib = _integrate(s, 1)
    other = synthwin.PreIntegration(a0, g0, s["ba"], s["bg"], NOISE)
    for k in range(first, first + count):
        other.push(s["dt"][k], s["acc"][k], s["gyr"][k])
    assert np.allclose(ib.jacobian, other.J, rtol=1e-13, atol=1e-15) and np.allclose(ib.covariance, other.P, rtol=1e-12, atol=1e-22)
    assert np.allclose(ib.delta_p, other.dp, atol=1e-15) and np.allclose(ib.delta_q, other.dq, atol=1e-15)


def test_preintegration_predicts_the_trajectory():
    """delta_p/v/q of noise-free data reproduce the relative motion (mid-point error only)."""
    s = imu_stream(6, noisy=False)
    tr, t = s["traj"], s["times"]
    ib = _integrate(s, 0)
    R0, T = tr.rot(t[0]), ib.sum_dt
    dp_true = R0.T @ (tr.pos(t[1]) - tr.pos(t[0]) - tr.vel(t[0]) * T + 0.5 * G * T * T)
    dv_true = R0.T @ (tr.vel(t[1]) - tr.vel(t[0]) + G * T)
    assert np.abs(ib.delta_p - dp_true).max() < 2e-4 and np.abs(ib.delta_v - dv_true).max() < 2e-3
    P, V, R = O.process_imu(tr.pos(t[0]), tr.vel(t[0]), R0, s["ba"], s["bg"], G, s["intervals"][0][2], s["intervals"][0][3],
                            [(s["dt"][k], s["acc"][k], s["gyr"][k]) for k in range(s["intervals"][0][1])])
    assert np.abs(P - tr.pos(t[1])).max() < 2e-4 and np.abs(V - tr.vel(t[1])).max() < 2e-3
    assert np.abs(R - tr.rot(t[1])).max() < 1e-4


def test_triangulate_noise_free_depth_and_rules():
    d = tracks(1, noise=0.0)
    depth0 = -np.ones(len(d["start_frame"]))
    depth0[::7] = 3.25  # already triangulated: untouched
    out = O.triangulate(d["Ps"], d["Rs"], d["tic"], d["ric"], d["start_frame"], d["obs_first"], d["obs_pts"], depth0)
    K = np.diff(d["obs_first"])
    ws = len(d["Ps"]) - 1
    eligible = (K >= 2) & (d["start_frame"] < ws - 2) & (depth0 <= 0)
    assert np.array_equal(out[~eligible], depth0[~eligible])
    rel = np.abs(out[eligible] - d["true_depth"][eligible]) / d["true_depth"][eligible]
    assert eligible.sum() > 50 and np.median(rel) < 1e-5  # float32 bearings limit the accuracy


def test_triangulate_negative_depth_becomes_init_depth():
    d = tracks(2, n_feat=60)
    pts = d["obs_pts"].copy()
    for l in range(0, 60, 3):  # flip the host bearing: the triangulated point lies behind it (depth = X . pts_0 < 0)
        pts[d["obs_first"][l]] *= -1
    out = O.triangulate(d["Ps"], d["Rs"], d["tic"], d["ric"], d["start_frame"], d["obs_first"], pts, -np.ones(60), 5.0)
    K = np.diff(d["obs_first"])
    sel = [l for l in range(0, 60, 3) if K[l] >= 2 and d["start_frame"][l] < len(d["Ps"]) - 3]
    assert len(sel) > 3 and all(out[l] == 5.0 for l in sel)
