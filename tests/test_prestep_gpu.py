"""IMU pre-integration and triangulation on device vs oracle/prestep.py — reference integration_base.h:31-157,
estimator.cpp:86-120, feature_manager.cpp:199-253.

Tolerances (fp64 on both sides; the reference's Eigen products and JacobiSVD accumulate in another order, so bit parity
is not defined): pre-integrated deltas 1e-13 absolute, Jacobian 1e-12 relative to its largest entry, covariance 1e-11
relative to its largest entry, predicted state 1e-12; triangulated depth 1e-9 relative (one-sided Jacobi vs LAPACK SVD
of the same matrix), skip / INIT_DEPTH decisions identical.
"""
import numpy as np
import pytest

from lfvio import prestep as PS
from oracle import prestep as O
from prestep_data import G, NOISE, imu_stream, tracks

pytestmark = pytest.mark.gpu


def _oracle_interval(s, i):
    first, count, a0, g0 = s["intervals"][i]
    ib = O.IntegrationOracle(a0, g0, s["ba"], s["bg"], NOISE)
    for k in range(first, first + count):
        ib.push_back(s["dt"][k], s["acc"][k], s["gyr"][k])
    return ib


@pytest.mark.parametrize("seed,n_int,rate", [(1, 10, 200.0), (2, 20, 200.0), (3, 3, 1000.0)])
def test_preintegrate_matches_oracle(lfvio_mod, gpu, seed, n_int, rate):
    s = imu_stream(seed, n_intervals=n_int, rate=rate)
    tr, t = s["traj"], s["times"]
    ivs = [PS.make_interval(f, c, i, a0, g0, s["ba"], s["bg"], tr.pos(t[i]), tr.vel(t[i]), tr.rot(t[i]))
           for i, (f, c, a0, g0) in enumerate(s["intervals"])]
    out, state = lfvio_mod.preintegrate(ivs, s["dt"], s["acc"], s["gyr"], NOISE, G, device=gpu)
    for i in range(n_int):
        ib, f = _oracle_interval(s, i), out[i]
        assert f.frame_i == i and abs(f.sum_dt - ib.sum_dt) < 1e-15
        assert np.abs(np.array(f.delta_p) - ib.delta_p).max() < 1e-13
        assert np.abs(np.array(f.delta_v) - ib.delta_v).max() < 1e-13
        assert np.abs(np.array(f.delta_q) - ib.delta_q).max() < 1e-14
        assert np.array_equal(np.array(f.linearized_ba), s["ba"]) and np.array_equal(np.array(f.linearized_bg), s["bg"])
        J, P = np.array(f.jacobian).reshape(15, 15), np.array(f.covariance).reshape(15, 15)
        assert np.abs(J - ib.jacobian).max() < 1e-12 * np.abs(ib.jacobian).max()
        assert np.abs(P - ib.covariance).max() < 1e-11 * np.abs(ib.covariance).max()
        first, count, a0, g0 = s["intervals"][i]
        Po, Vo, Ro = O.process_imu(tr.pos(t[i]), tr.vel(t[i]), tr.rot(t[i]), s["ba"], s["bg"], G, a0, g0,
                                   [(s["dt"][k], s["acc"][k], s["gyr"][k]) for k in range(first, first + count)])
        assert np.abs(state[i, 0:3] - Po).max() < 1e-12 and np.abs(state[i, 3:6] - Vo).max() < 1e-12
        assert np.abs(state[i, 6:15].reshape(3, 3) - Ro).max() < 1e-13


def test_preintegrate_empty_and_single_sample(lfvio_mod, gpu):
    out, state = lfvio_mod.preintegrate([], np.zeros(0), np.zeros((0, 3)), np.zeros((0, 3)), NOISE, G, device=gpu)
    assert out == []
    iv = PS.make_interval(0, 0, 4, [0.1, 0.2, 9.8], [0.01, 0, 0], np.zeros(3), np.zeros(3))
    out, _ = lfvio_mod.preintegrate([iv], np.zeros(1), np.zeros((1, 3)), np.zeros((1, 3)), NOISE, G, device=gpu)
    f = out[0]  # IntegrationBase right after its constructor (integration_base.h:13-29)
    assert f.frame_i == 4 and f.sum_dt == 0.0 and list(f.delta_q) == [1.0, 0.0, 0.0, 0.0]
    assert np.array_equal(np.array(f.jacobian).reshape(15, 15), np.eye(15)) and not np.any(np.array(f.covariance))
    iv = PS.make_interval(0, 1, 0, [0.1, 0.2, 9.8], [0.01, 0, 0], [0.01, 0, 0], [0, 0.001, 0])
    out, _ = lfvio_mod.preintegrate([iv], [0.005], [[0.12, 0.18, 9.79]], [[0.012, 0.001, -0.002]], NOISE, G, device=gpu)
    ib = O.IntegrationOracle([0.1, 0.2, 9.8], [0.01, 0, 0], [0.01, 0, 0], [0, 0.001, 0], NOISE)
    ib.push_back(0.005, [0.12, 0.18, 9.79], [0.012, 0.001, -0.002])
    assert np.abs(np.array(out[0].covariance).reshape(15, 15) - ib.covariance).max() < 1e-12 * np.abs(ib.covariance).max()


def test_preintegrated_factor_feeds_the_imu_factor(lfvio_mod, gpu):
    """The device record is the IMUFactor input: a window built from it evaluates like one built from the oracle record."""
    from lfvio import synthwin
    from oracle import backend as ob

    w = synthwin.make_window(seed=3, opts=lfvio_mod.solver_default_opts(), perturb=0.3)
    s = imu_stream(3, n_intervals=w.n_imu)
    ivs = [PS.make_interval(f, c, i, a0, g0, s["ba"], s["bg"]) for i, (f, c, a0, g0) in enumerate(s["intervals"])]
    out, _ = lfvio_mod.preintegrate(ivs, s["dt"], s["acc"], s["gyr"], NOISE, G, device=gpu)
    a, b = w.clone(), w.clone()
    for i in range(w.n_imu):
        ib = _oracle_interval(s, i)
        a.imu[i] = out[i]
        fo = type(out[i])()
        fo.frame_i, fo.sum_dt = i, ib.sum_dt
        fo.delta_p[:], fo.delta_q[:], fo.delta_v[:] = ib.delta_p.tolist(), ib.delta_q.tolist(), ib.delta_v.tolist()
        fo.linearized_ba[:], fo.linearized_bg[:] = s["ba"].tolist(), s["bg"].tolist()
        fo.jacobian[:], fo.covariance[:] = ib.jacobian.ravel().tolist(), ib.covariance.ravel().tolist()
        b.imu[i] = fo
    ra, rb = ob.evaluate(a), ob.evaluate(b)
    assert np.abs(ra.res_imu - rb.res_imu).max() < 1e-7 * max(1.0, np.abs(rb.res_imu).max())


@pytest.mark.parametrize("seed,n_frames,n_feat,noise", [(1, 11, 300, 0.0), (2, 11, 500, 1e-3), (3, 21, 2000, 5e-4)])
def test_triangulate_matches_oracle(lfvio_mod, gpu, seed, n_frames, n_feat, noise):
    d = tracks(seed, n_frames=n_frames, n_feat=n_feat, noise=noise)
    depth0 = -np.ones(n_feat)
    depth0[::5] = 2.5
    pts = d["obs_pts"].copy()
    for l in range(3, n_feat, 11):
        pts[d["obs_first"][l]] *= -1  # behind the host camera -> INIT_DEPTH
    want = O.triangulate(d["Ps"], d["Rs"], d["tic"], d["ric"], d["start_frame"], d["obs_first"], pts, depth0, 5.0)
    got = lfvio_mod.triangulate(d["Ps"], d["Rs"], d["tic"], d["ric"], d["start_frame"], d["obs_first"], pts, depth0, 5.0, device=gpu)
    same = (want == depth0) | (want == 5.0)
    assert np.array_equal(got[same], want[same])  # untouched / skipped / INIT_DEPTH decisions
    assert (~same).sum() > n_feat // 4
    rel = np.abs(got[~same] - want[~same]) / np.abs(want[~same])
    assert rel.max() < 1e-9, rel.max()


def test_triangulate_empty(lfvio_mod, gpu):
    d = tracks(1, n_feat=4)
    got = lfvio_mod.triangulate(d["Ps"], d["Rs"], d["tic"], d["ric"], np.zeros(0, np.int32), np.zeros(1, np.int32), np.zeros((0, 3)),
                                np.zeros(0), device=gpu)
    assert got.shape == (0,)
