"""FeatureManager on the device (lfvio_fm_*, include/lfvio.h) against the Python restatement of the reference's bookkeeping
(oracle/feature_manager.py, itself pinned to the reference's own FeatureManager in tests/test_oracle_ref.py): key-frame
decision, the three ways the track list slides, setDepth / getDepthVector, the factor list of Estimator::optimization()
record by record, the device triangulation, and the closed loop whose factor list never exists on the host."""
import numpy as np
import pytest

from lfvio import violoop
from oracle.feature_manager import FeatureManagerOracle

pytestmark = pytest.mark.gpu
W = 10


def _frame(world, k, td):
    obs = world.frames[k]
    ids = np.array(sorted(obs), np.int32)
    f8 = np.zeros((len(ids), 8))
    image = {}
    for r, i in enumerate(ids):
        b, vel, row = obs[int(i)]
        u = 100.0 + 3.0 * r
        f8[r] = [b[0], b[1], b[2], u, row, vel[0], vel[1], vel[2]]
        image[int(i)] = (b, (vel, (u, row), td))
    return ids, f8, image


def _check_factors(store, ref, k):
    proj, lm, feat = store.build_factors()
    want, first, lam = ref.factor_list()
    assert len(proj) == len(want) and np.array_equal(lm, first), k
    assert np.array_equal(feat, lam), k
    for p, (fi, fj, l, pi, pj, ai, aj) in zip(proj, want):
        assert (p.frame_i, p.frame_j, p.landmark) == (fi, fj, l), k
        assert np.array_equal(p.pts_i[:], pi) and np.array_equal(p.pts_j[:], pj), k
        assert np.array_equal(p.vel_i[:], ai[0]) and np.array_equal(p.vel_j[:], aj[0]), k
        assert (p.row_i, p.row_j, p.td_i, p.td_j) == (ai[1][1], aj[1][1], ai[2], aj[2]), k


@pytest.mark.parametrize("min_parallax", [0.0, 0.45, 1e9])
def test_feature_store_stream(lfvio_mod, gpu, min_parallax):
    world = violoop.World(seed=6, n_frames=34, max_feat=60)
    store = lfvio_mod.FeatureStore(W, capacity=256, max_proj=4096, min_parallax=min_parallax)
    ref = FeatureManagerOracle(W, min_parallax)
    rng = np.random.default_rng(0)
    frame_count, kinds = 0, []
    for k in range(world.n_frames):
        td = 0.001 * k
        ids, f8, image = _frame(world, k, td)
        key, tracked = store.add_frame(frame_count, ids, f8, td)
        key_o = ref.add_feature_check_parallax(frame_count, image)
        assert key == key_o and tracked == ref.last_track_num, k
        assert [s[:3] for s in store.snapshot()] == [s[:3] for s in ref.snapshot()], k
        n_sel = ref.get_feature_count()
        if n_sel:  # setDepth over the selected features, getDepthVector back, then the factor list
            x = rng.uniform(0.1, 0.5, n_sel)
            assert store.set_depth(x) == n_sel
            ref.set_depth(x)
            assert np.array_equal(store.depth_vector(), ref.depth_vector()), k
        _check_factors(store, ref, k)
        if frame_count < W:
            frame_count += 1
            continue
        kinds.append(key)
        if key and k % 2 == 0:
            store.remove_back()
            ref.remove_back()
        elif key:
            mR, nR = world.truth(k - W)[1], world.truth(k - W + 1)[1]
            mP, nP = world.truth(k - W)[0], world.truth(k - W + 1)[0]
            store.remove_back_shift_depth(mR, mP, nR, nP)
            ref.remove_back_shift_depth(mR, mP, nR, nP)
        else:
            store.remove_front(frame_count)
            ref.remove_front(frame_count)
        store.remove_failures()  # LF-VIO never sets solve_flag = 2 (feature_manager.cpp:151-157): a no-op, as in the reference
        snap, rs = store.snapshot(), ref.snapshot()
        assert [s[:3] for s in snap] == [s[:3] for s in rs], k
        assert np.allclose([s[3] for s in snap], [s[3] for s in rs], rtol=1e-13, atol=0), k
        # carry the device's depths over so that fused-multiply-add rounding in the shift does not accumulate in the comparison
        for f, s in zip(ref.feature, snap):
            f.depth = s[3]
    store.close()
    if min_parallax == 0.0:
        assert all(kinds)
    else:
        assert any(kinds) and not all(kinds), kinds


def test_feature_store_capacity_is_an_error_not_a_crash(lfvio_mod, gpu):
    store = lfvio_mod.FeatureStore(W, capacity=8, max_proj=64)
    f8 = np.zeros((8, 8))
    f8[:, 2] = 1.0
    store.add_frame(0, np.arange(8), f8)
    with pytest.raises(Exception, match="more live features than slots"):
        store.add_frame(1, np.arange(4, 12), f8)
    store.clear()  # the error is sticky until clearState
    assert store.add_frame(0, np.arange(8), f8) == (True, 0)
    with pytest.raises(Exception, match="ascend"):
        store.add_frame(1, np.array([3, 2, 1]), f8[:3])


def test_feature_store_triangulate_matches_packed_kernel(lfvio_mod, gpu):
    """FeatureManager::triangulate over the store's slots == lfvio_triangulate over the same tracks packed on the host."""
    world = violoop.World(seed=3, n_frames=W + 1, max_feat=120)
    store = lfvio_mod.FeatureStore(W, capacity=1024, max_proj=16384)
    ref = FeatureManagerOracle(W)
    for k in range(W + 1):
        ids, f8, image = _frame(world, k, 0.0)
        store.add_frame(k, ids, f8)
        ref.add_feature_check_parallax(k, image)
    Ps = np.array([world.truth(k)[0] for k in range(W + 1)])
    Rs = np.array([world.truth(k)[1] for k in range(W + 1)])
    store.triangulate(Ps, Rs, world.tic, world.ric)
    start = np.array([f.start for f in ref.feature], np.int32)
    first = np.concatenate([[0], np.cumsum([len(f.pts) for f in ref.feature])]).astype(np.int32)
    pts = np.array([p for f in ref.feature for p in f.pts])
    want = lfvio_mod.triangulate(Ps, Rs, world.tic, world.ric, start, first, pts, np.full(len(start), -1.0), 5.0)
    got = np.array([s[3] for s in store.snapshot()])
    assert np.array_equal(got, want)
    assert (got > 0).sum() > 100 and (got == -1.0).sum() > 0  # selected ones triangulated, the others untouched


def test_window_consumes_device_built_list(lfvio_mod, gpu):
    """lfvio_window_optimize_full_fm against lfvio_window_optimize_full on the host-built list of the same tracks, step by step
    through 20 frames of one stream: same iteration counts, same track list, states / depths / next prior equal to within the
    solve's own run-to-run spread.  (The Schur accumulation adds with atomics, so two runs of the SAME path differ by up to
    ~1e-7 in the speeds after a few frames; measured here: host-vs-host 1.2e-7, store-vs-host 5.6e-7.  The inputs are the same
    bits — test_feature_store_stream compares the device-built records one by one.)"""
    world = violoop.World(seed=2, n_frames=40, max_feat=150)
    opts = lfvio_mod.solver_default_opts()
    opts.row = 960
    opts.estimate_td = 1
    win_a, win_b = lfvio_mod.Window(W, 1024, 16384), lfvio_mod.Window(W, 1024, 16384)
    store = lfvio_mod.FeatureStore(W, capacity=2048, max_proj=16384)
    host = violoop.VioLoop(world, violoop.device_ops(lfvio_mod, win_a), opts, F=W + 1)
    dev = violoop.VioLoopStore(world, win_b, store, opts, F=W + 1)
    TOL = 5e-6
    for step in range(20):
        more_a, more_b = host.step(), dev.step()
        assert more_a == more_b
        assert host.iterations[-1] == dev.iterations[-1], step
        assert np.abs(host.pose - dev.pose).max() < TOL and np.abs(host.sb - dev.sb).max() < TOL and abs(host.td - dev.td) < TOL, step
        assert host.prior.n == dev.prior.n and host.prior.n_blocks == dev.prior.n_blocks, step
        nb = host.prior.n_blocks
        assert np.array_equal(host.prior.block_id[:nb], dev.prior.block_id[:nb]) and np.array_equal(host.prior.block_idx[:nb], dev.prior.block_idx[:nb])
        Ha, ba = host.prior.information()
        Hb, bb = dev.prior.information()
        assert np.abs(Ha - Hb).max() < 1e-5 * np.abs(Ha).max() and np.abs(ba - bb).max() < 1e-5 * max(np.abs(ba).max(), 1.0), step
        snap = store.snapshot()
        assert [(s[0], s[1], s[2]) for s in snap] == [(f.fid, f.start, len(f.obs)) for f in host.features], step
        assert np.allclose([s[3] for s in snap], [f.depth for f in host.features], rtol=1e-4), step


def test_closed_loop_with_device_store_tracks_truth(lfvio_mod, gpu):
    """A whole stream (50 solves): the loop over the device store ends where the loop over the host-side list ends."""
    world = violoop.World(seed=1, n_frames=60, max_feat=150)
    opts = lfvio_mod.solver_default_opts()
    opts.row = 960
    opts.estimate_td = 1
    host = violoop.VioLoop(world, violoop.device_ops(lfvio_mod, lfvio_mod.Window(W, 1024, 16384)), opts, F=W + 1).run()
    loop = violoop.VioLoopStore(world, lfvio_mod.Window(W, 1024, 16384), lfvio_mod.FeatureStore(W, 2048, 16384), opts, F=W + 1).run()
    ate_h, ks_h = violoop.trajectory_error(host)
    ate_d, ks_d = violoop.trajectory_error(loop)
    assert ks_h == ks_d == list(range(10, 60))
    assert ate_d < 0.12 and abs(ate_d - ate_h) < 1e-5, (ate_d, ate_h)
    assert abs(loop.td - host.td) < 1e-6


def test_device_store_three_steps(lfvio_mod, gpu):
    """Three frames of the loop over the device store (add-frame, triangulate, build, solve from the device-built list, setDepth on
    the device, removeBackShiftDepth): the short case tools/sanitize.sh runs under compute-sanitizer."""
    world = violoop.World(seed=4, n_frames=16, max_feat=80)
    opts = lfvio_mod.solver_default_opts()
    opts.row = 960
    opts.estimate_td = 1
    loop = violoop.VioLoopStore(world, lfvio_mod.Window(W, 512, 8192), lfvio_mod.FeatureStore(W, 1024, 8192), opts, F=W + 1)
    for _ in range(3):
        assert loop.step()
    assert loop.iterations == [8, 8, 8] or all(0 < i <= 8 for i in loop.iterations)
    assert np.isfinite(loop.pose).all() and len(loop.store.snapshot()) > 50
