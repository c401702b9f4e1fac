"""Pins the self-written oracle to the REFERENCE'S OWN CODE.

oracle/_ref/libref_factors.so is /root/reference/vins_estimator/src/{factor/*, utility/utility.cpp, feature_manager.cpp,
initial/initial_aligment.cpp} and ScaramuzzaCamera.cc:602-674 compiled unmodified (oracle/Makefile, target `ref`) against
stand-in Eigen / Ceres-interface / ROS headers (oracle/ref_stubs). These tests compare, on the same inputs:

  oracle/backend/backend.cpp   vs  ProjectionTdFactor / ProjectionFactor / IMUFactor / MarginalizationFactor::Evaluate,
                                   ResidualBlockInfo::Evaluate (Cauchy correction), PoseLocalParameterization::Plus,
                                   MarginalizationInfo::marginalize (both flags)
  oracle/prestep.py            vs  IntegrationBase::push_back, FeatureManager::triangulate
  oracle/camera.py             vs  OCAMCamera::liftProjective / liftSphere / spaceToPlane   (bit-exact)
  oracle/initial.py            vs  solveGyroscopeBias, LinearAlignment + RefineGravity
  oracle/feature_manager.py and the C++ host mirror  vs  class FeatureManager

CPU only. On a box without /root/reference the prebuilt library is used; without either the tests skip and the committed
fixtures made from it (tests/golden/ref_backend.npz, tests/test_ref_golden.py) still hold the pin.
"""
import ctypes as C
import os

import numpy as np
import pytest

import lfvio
from conftest import ROOT
from initial_data import InitScenario
from lfvio import backend as B
from lfvio import synthwin, violoop
from oracle import backend as ob
from oracle import camera as ocam
from oracle import initial as OI
from oracle import prestep as OP
from oracle import ref
from oracle.feature_manager import FeatureManagerOracle
from prestep_data import G, NOISE, imu_stream, tracks

pytestmark = pytest.mark.skipif(not ref.available(), reason="oracle/_ref not built and /root/reference not mounted")


def rel(a, b):
    return float(np.abs(np.asarray(a) - np.asarray(b)).max() / max(np.abs(np.asarray(b)).max(), 1e-300))


@pytest.fixture(scope="module")
def window():
    w = synthwin.make_window_with_prior(ob.optimize, ob.marginalize, seed=1, opts=ob.default_opts(), perturb=0.3)
    return w


@pytest.mark.parametrize("robust", [False, True])
@pytest.mark.parametrize("estimate_td", [1, 0])
def test_factor_residuals_and_jacobians(window, robust, estimate_td):
    w = window.clone()
    w.opts = ob.default_opts()
    w.opts.estimate_td = estimate_td
    a, b = ob.evaluate(w, robust), ref.evaluate(w, robust)
    assert rel(a.res_proj, b.res_proj) < 1e-11 and rel(a.jac_proj, b.jac_proj) < 1e-11
    assert rel(a.res_imu, b.res_imu) < 1e-12 and rel(a.jac_imu, b.jac_imu) < 1e-12
    assert rel(a.res_prior, b.res_prior) < 1e-12
    assert abs(a.cost.value - b.cost.value) < 1e-12 * b.cost.value
    # the prior's Jacobian blocks are column slices of linearized_jacobians with a zero 7th column (:368-378)
    assert np.array_equal(b.jac_prior, w.prior.jacobian())


@pytest.mark.parametrize("row,tr", [(960, 0.0), (960, 0.03), (961, 0.03), (2881, 0.011)])
def test_rolling_shutter_term_uses_floating_row_half(window, row, tr):
    """ROW is a double in the reference (parameters.h:39): row_i = _row_i - ROW / 2 (projection_td_factor.cpp:20-21)."""
    w = window.clone()
    w.opts = ob.default_opts()
    w.opts.row, w.opts.tr = row, tr
    a, b = ob.evaluate(w, False), ref.evaluate(w, False)
    assert rel(a.res_proj, b.res_proj) < 1e-11 and rel(a.jac_proj, b.jac_proj) < 1e-11


def test_extreme_geometry(window):
    """bearings exactly on the tangent-basis switch (a == (0,0,1)), in the negative half space, and a long baseline."""
    w = window.clone()
    w.opts = ob.default_opts()
    for k, pj in enumerate([(0, 0, 1.0), (0, 0, -1.0), (0.6, 0.0, -0.8), (1.0, 0, 0)]):
        w.proj[k].pts_j[:] = pj
    a, b = ob.evaluate(w, True), ref.evaluate(w, True)
    # pts_j = (0, 0, -1) makes the reference's tangent basis 0 / 0 (projection_td_factor.cpp:25-30 only guards +z): NaN on both sides
    assert np.isnan(b.res_proj[2:4]).all() and np.array_equal(np.isnan(a.res_proj[:8]), np.isnan(b.res_proj[:8]))
    assert np.array_equal(np.isnan(a.jac_proj[:4]), np.isnan(b.jac_proj[:4]))
    keep = [0, 2, 3]
    assert rel(a.res_proj[:8].reshape(4, 2)[keep], b.res_proj[:8].reshape(4, 2)[keep]) < 1e-12
    assert rel(a.jac_proj[:4][keep], b.jac_proj[:4][keep]) < 1e-11


def test_pose_plus():
    rng = np.random.default_rng(0)
    for _ in range(50):
        q = rng.normal(size=4)
        x = np.concatenate([rng.normal(size=3), q / np.linalg.norm(q)])
        d = rng.normal(scale=0.3, size=6)
        assert np.abs(ob.pose_plus(x, d) - ref.pose_plus(x, d)).max() < 1e-15


@pytest.mark.parametrize("flag", [B.MARGIN_OLD, B.MARGIN_SECOND_NEW])
@pytest.mark.parametrize("seed", [2, 5])
def test_marginalization_prior(flag, seed):
    """MarginalizationInfo::marginalize (marginalization_factor.cpp:174-297) incl. the eigen-decomposition pseudo-inverse of
    A_mm and the eps = 1e-8 cut of the second decomposition; compared through J0^T J0, J0^T r0 and |r0|^2 because (J0, r0)
    is unique only up to the eigenvector signs and the block order only up to the unordered_map."""
    w = synthwin.make_window_with_prior(ob.optimize, ob.marginalize, seed=seed, opts=ob.default_opts(), perturb=0.3)
    ob.optimize(w)
    po = ob.marginalize(w, flag)
    pr, m = ref.marginalize(w, flag)
    assert po.n == pr.n and po.n_blocks == pr.n_blocks
    assert m == (6 if flag == B.MARGIN_SECOND_NEW else 15 + sum(1 for l in {f.landmark for f in w.proj[: w.n_proj] if f.frame_i == 0}))
    Ao, go, xo = ref.canonical_information(po, w.F)
    Ar, gr, xr = ref.canonical_information(pr, w.F)
    assert sorted(xo) == sorted(xr) and all(np.array_equal(xo[k], xr[k]) for k in xo)
    assert rel(Ao, Ar) < 2e-9 and rel(go, gr) < 2e-9
    # |r0|^2 = b'^T A'^+ b' only shifts the cost by a constant; it feels the 1e-8 eigenvalue cut in the gauge directions
    assert abs(po.residual() @ po.residual() - pr.residual() @ pr.residual()) < 1e-5 * (pr.residual() @ pr.residual())


def test_marginalization_without_prior_and_too_long_imu():
    w = synthwin.make_window(seed=3, opts=ob.default_opts(), perturb=0.3)
    po, (pr, m) = ob.marginalize(w, B.MARGIN_OLD), ref.marginalize(w, B.MARGIN_OLD)
    Ao, go, _ = ref.canonical_information(po, w.F)
    Ar, gr, _ = ref.canonical_information(pr, w.F)
    assert po.n == pr.n and rel(Ao, Ar) < 2e-9 and rel(go, gr) < 2e-9
    assert ref.marginalize(w, B.MARGIN_SECOND_NEW)[0].n == 0 and ob.marginalize(w, B.MARGIN_SECOND_NEW).n == 0
    w.imu[0].sum_dt = 10.5  # estimator.cpp:857: the IMU factor is left out
    po, (pr, m) = ob.marginalize(w, B.MARGIN_OLD), ref.marginalize(w, B.MARGIN_OLD)
    assert po.n == pr.n == 67 and m == pr.n * 0 + m  # the speed-bias of frame 1 is no longer connected
    Ao, go, _ = ref.canonical_information(po, w.F)
    Ar, gr, _ = ref.canonical_information(pr, w.F)
    assert rel(Ao, Ar) < 2e-9 and rel(go, gr) < 2e-9


def test_preintegration():
    """IntegrationBase::push_back / midPointIntegration (integration_base.h:30-157)."""
    s = imu_stream(4, n_intervals=4)
    from lfvio import prestep as PS

    L = ref.lib()
    ivs = [PS.make_interval(f, c, i, a0, g0, s["ba"], s["bg"]) for i, (f, c, a0, g0) in enumerate(s["intervals"])]
    arr = (PS.ImuInterval * len(ivs))(*ivs)
    out = (B.ImuFactor * len(ivs))()
    nz = PS.ImuNoise(*NOISE)
    dp = C.POINTER(C.c_double)
    L.ref_preintegrate.argtypes = [C.POINTER(PS.ImuInterval), C.c_int, dp, dp, dp, C.c_int, C.POINTER(PS.ImuNoise), dp, C.POINTER(B.ImuFactor)]
    assert L.ref_preintegrate(arr, len(ivs), s["dt"].ctypes.data_as(dp), s["acc"].ctypes.data_as(dp), s["gyr"].ctypes.data_as(dp),
                              len(s["dt"]), C.byref(nz), G.ctypes.data_as(dp), out) == 0
    for i, (first, count, a0, g0) in enumerate(s["intervals"]):
        ib = OP.IntegrationOracle(a0, g0, s["ba"], s["bg"], NOISE)
        for k in range(first, first + count):
            ib.push_back(s["dt"][k], s["acc"][k], s["gyr"][k])
        f = out[i]
        assert abs(f.sum_dt - ib.sum_dt) < 1e-15
        assert np.abs(np.array(f.delta_p) - ib.delta_p).max() < 1e-15 and np.abs(np.array(f.delta_v) - ib.delta_v).max() < 1e-15
        assert np.abs(np.array(f.delta_q) - ib.delta_q).max() < 1e-15
        assert rel(np.array(f.jacobian).reshape(15, 15), ib.jacobian) < 1e-13
        assert rel(np.array(f.covariance).reshape(15, 15), ib.covariance) < 1e-13


def test_lift_projective_bit_exact():
    rng = np.random.default_rng(0)
    uv = np.stack([rng.uniform(0, 1280, 100003), rng.uniform(0, 960, 100003)], 1)
    uv[:3] = [[645.107791, 486.025172], [1100, 700], [200, 100]]  # the centre (phi = 0) and two z < 0 rays
    for cam in (lfvio.PAL_UP, lfvio.scaled_camera(3.0)):
        u = uv * (cam["image_width"] / 1280.0)
        assert np.array_equal(ocam.lift_projective(cam, u).view(np.uint64), ref.lift(cam, u, False).view(np.uint64))
        assert np.array_equal(ocam.lift_sphere(cam, u).view(np.uint64), ref.lift(cam, u, True).view(np.uint64))
    P = rng.normal(size=(5000, 3))
    assert np.abs(ocam.space_to_plane(lfvio.PAL_UP, P) - ref.space_to_plane(lfvio.PAL_UP, P)).max() < 1e-9


def test_utility_rotations():
    rng = np.random.default_rng(1)
    for _ in range(20):
        g = rng.normal(size=3) * 9.8
        assert np.abs(OI.g2R(g) - ref.g2R(g)).max() < 1e-13
        ypr = rng.uniform(-170, 170, 3) * [1, 0.5, 1]
        R = ref.ypr2R(ypr)
        assert np.abs(R @ R.T - np.eye(3)).max() < 1e-14 and np.abs(ref.R2ypr(R) - ypr).max() < 1e-10


@pytest.mark.parametrize("seed,noise", [(1, 0.0), (2, 0.3)])
def test_visual_imu_alignment(seed, noise):
    """solveGyroscopeBias + LinearAlignment + RefineGravity (initial_aligment.cpp). RefineGravity keeps accumulating into A and
    b across its four passes (zeroed once, scaled by 1000 every pass): reproduced by the oracle and the device kernel."""
    sc = InitScenario(seed=seed, F=11, pixel_noise=noise)
    w = sc.world
    pre = sc.preintegrations()
    Rs, Ts = sc.Rb, sc.Pc * 0.37
    imu = [w.imu[k] for k in range(1, sc.F)]
    r0 = ref.visual_imu_alignment(Rs, Ts, imu, w.noise, sc.tic, sc.g_norm, mode=0)
    dbg = OI.solve_gyroscope_bias(Rs, [p.delta_q for p in pre], [p.jacobian[3:6, 12:15] for p in pre])
    assert np.abs(r0["bg"] - dbg).max() < 1e-13
    r1 = ref.visual_imu_alignment(Rs, Ts, imu, w.noise, sc.tic, sc.g_norm, mode=1)
    ok, g, x = OI.linear_alignment(Rs, Ts, [p.delta_p for p in pre], [p.delta_v for p in pre], [p.sum_dt for p in pre], sc.tic, sc.g_norm)
    assert r1["ok"] == ok and np.abs(r1["g"] - g).max() < 1e-9 and np.abs(r1["x"] - x).max() < 1e-8
    assert abs(1.0 / x[-1] - 0.37) < 0.02  # and the answer is the true scale


def test_visual_imu_alignment_rejects_a_wrong_gravity():
    sc = InitScenario(seed=1, F=11)
    w, pre = sc.world, sc.preintegrations()
    Rs, Ts = sc.Rb, sc.Pc * 0.37
    imu = [w.imu[k] for k in range(1, sc.F)]
    r1 = ref.visual_imu_alignment(Rs, Ts, imu, w.noise, sc.tic, 3.0, mode=1)
    ok, g, x = OI.linear_alignment(Rs, Ts, [p.delta_p for p in pre], [p.delta_v for p in pre], [p.sum_dt for p in pre], sc.tic, 3.0)
    assert (not ok) and (not r1["ok"]) and np.abs(r1["g"] - g).max() < 1e-9 and np.abs(r1["x"] - x).max() < 1e-8


def test_triangulate():
    """FeatureManager::triangulate (feature_manager.cpp:199-253) incl. tracks behind the camera (-> INIT_DEPTH)."""
    t = tracks(2, n_frames=11, n_feat=120, noise=0.002, behind_frac=0.1)
    fm = ref.RefFeatureManager(0.0)
    n = len(t["start_frame"])
    for fr in range(11):  # feed frame by frame (ids in list order = feature order because ids ascend with first appearance)
        ids, pts = [], []
        for l in range(n):
            k = fr - t["start_frame"][l]
            if 0 <= k < t["obs_first"][l + 1] - t["obs_first"][l]:
                ids.append(l)
                p = np.zeros(8)
                p[:3] = t["obs_pts"][t["obs_first"][l] + k]
                pts.append(p)
        fm.add(fr, ids, np.array(pts).reshape(-1, 8))
    meta, depth, _ = fm.snapshot()
    order = meta[:, 0]
    fm.triangulate(t["Ps"], t["Rs"], t["tic"], t["ric"])
    _, depth_ref, _ = fm.snapshot()
    want = OP.triangulate(t["Ps"], t["Rs"], t["tic"], t["ric"], t["start_frame"], t["obs_first"], t["obs_pts"], np.full(n, -1.0))
    assert np.array_equal(np.sort(order), np.arange(n))
    got = np.empty(n)
    got[order] = depth_ref
    assert (got == 5.0).sum() > 3 and np.array_equal(got == 5.0, want == 5.0)
    assert rel(got, want) < 1e-9


@pytest.mark.parametrize("n,seed", [(150, 1), (60, 2), (9, 3), (8, 5), (400, 4)])
def test_reject_with_f_ransac(n, seed):
    """compute_E_21 / check_inliers / myfindFundamentalMat (feature_tracker.cpp:186-327) — the reference's own code, fed from the
    injected random source — against oracle/ransac.py: same inlier decisions, same refit essential matrix up to sign."""
    from oracle import ransac as R
    from ransac_data import two_view

    cam, cfg = lfvio.PAL_UP, lfvio.tracker_defaults()
    cur, forw, _ = two_view(cam, cfg, n, seed)
    b1, b2 = R.lift_pair(cam, cur, forw)
    for frame in (0, 11):
        Eo, so, _, _ = R.find_fundamental(b1, b2, 7 + seed, frame)
        Er, sr = ref.find_fundamental(b1, b2, 7 + seed, frame)
        assert np.array_equal(so, sr)
        assert np.abs(Eo - np.sign((Eo * Er).sum()) * Er).max() < 1e-12


@pytest.fixture(scope="module")
def host():
    lib = C.CDLL(os.path.join(ROOT, "lf-vio_b200", "liblfvio_host.so"))
    lib.lfvio_host_fm_create.restype = C.c_void_p
    lib.lfvio_host_fm_create.argtypes = [C.c_int, C.c_double]
    lib.lfvio_host_fm_destroy.argtypes = [C.c_void_p]
    lib.lfvio_host_fm_add.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_double]
    lib.lfvio_host_fm_corresponding.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int]
    lib.lfvio_host_fm_slide.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p]
    lib.lfvio_host_fm_snapshot.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
    return lib


@pytest.mark.parametrize("min_parallax", [0.0, 0.45, 1e9])
def test_feature_manager_stream(host, min_parallax):
    """The reference's FeatureManager, the Python restatement and the C++ host mirror on one stream with key frames,
    non-key frames and all three ways the track list slides."""
    W = 10
    world = violoop.World(seed=6, n_frames=34, max_feat=60)
    h = host.lfvio_host_fm_create(W, min_parallax)
    py = FeatureManagerOracle(W, min_parallax)
    rf = ref.RefFeatureManager(min_parallax)
    rng = np.random.default_rng(0)
    frame_count = 0
    for k in range(world.n_frames):
        obs = world.frames[k]
        ids = np.array(sorted(obs), np.int32)
        b = np.ascontiguousarray([obs[i][0] for i in ids])
        p8 = np.zeros((len(ids), 8))
        p8[:, :3] = b
        key_h = bool(host.lfvio_host_fm_add(h, frame_count, ids.ctypes.data, b.ctypes.data, len(ids), 0.0))
        key_p = py.add_feature_check_parallax(frame_count, {int(i): b[j] for j, i in enumerate(ids)})
        key_r = rf.add(frame_count, ids, p8)
        assert key_r == key_p == key_h, k
        meta, depth, cnt = rf.snapshot()
        snap = py.snapshot()
        assert [tuple(m[:3]) for m in meta.tolist()] == [s[:3] for s in snap] and cnt == py.get_feature_count(), k
        if frame_count >= 1:
            assert np.array_equal(rf.corresponding(frame_count - 1, frame_count), py.get_corresponding(frame_count - 1, frame_count))
        if frame_count < W:
            frame_count += 1
            continue
        if key_r and k % 2 == 0:
            host.lfvio_host_fm_slide(h, 0, 0, None)
            py.remove_back()
            rf.slide(0)
        elif key_r:
            d = rng.uniform(2.0, 9.0, len(py.feature))
            rf.snapshot(set_depth=d)
            hm, hd = np.zeros((4096, 3), np.int32), np.zeros(4096)
            host.lfvio_host_fm_snapshot(h, hm.ctypes.data, hd.ctypes.data, 4096, np.ascontiguousarray(d).ctypes.data, None)
            for f, dd in zip(py.feature, d):
                f.depth = float(dd)
            mR, nR = world.truth(k - W)[1], world.truth(k - W + 1)[1]
            mP, nP = world.truth(k - W)[0], world.truth(k - W + 1)[0]
            pose = np.concatenate([mR.ravel(), mP, nR.ravel(), nP])
            host.lfvio_host_fm_slide(h, 2, 0, pose.ctypes.data)
            py.remove_back_shift_depth(mR, mP, nR, nP)
            rf.slide(2, 0, pose)
        else:
            host.lfvio_host_fm_slide(h, 1, frame_count, None)
            py.remove_front(frame_count)
            rf.slide(1, frame_count)
        meta, depth, cnt = rf.snapshot()
        snap = py.snapshot()
        hm, hd, hc = np.zeros((4096, 3), np.int32), np.zeros(4096), C.c_int(0)
        nh = host.lfvio_host_fm_snapshot(h, hm.ctypes.data, hd.ctypes.data, 4096, None, C.byref(hc))
        assert [tuple(m[:3]) for m in meta.tolist()] == [s[:3] for s in snap] == [tuple(r) for r in hm[:nh].tolist()], k
        assert np.abs(depth - np.array([s[3] for s in snap])).max() < 1e-12 and np.abs(depth - hd[:nh]).max() < 1e-12, k
        assert cnt == hc.value
        if cnt:  # setDepth / getDepthVector over the selected features (the restatement's set_depth / depth_vector)
            x = rng.uniform(0.1, 0.5, cnt)
            rf.set_depth(x)
            py.set_depth(x)
            assert np.array_equal(rf.depth_vector(), py.depth_vector()), k
            assert np.array_equal(rf.snapshot()[1], np.array([s[3] for s in py.snapshot()])), k
            rf.snapshot(set_depth=depth)  # back to the depths of the stream
            for f, dd in zip(py.feature, depth):
                f.depth = float(dd)
    host.lfvio_host_fm_destroy(h)
