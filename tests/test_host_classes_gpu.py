"""The C++ host mirrors of the reference classes (lf-vio_b200/host/) driven through their ctypes test hooks:
FeatureTracker (readImage + the node's updateID loop) and Estimator::optimization() (vector2double -> solve ->
double2vector gauge fix -> next prior) must reproduce what the C-ABI + oracle give on the same inputs."""
import ctypes as C
import os

import numpy as np
import pytest

from conftest import ROOT
from lfvio import backend as B
from lfvio import synth, synthwin
from lfvio.config import PAL_UP, tracker_defaults
from oracle import backend as ob

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def host():
    lib = C.CDLL(os.path.join(ROOT, "lf-vio_b200", "liblfvio_host.so"))
    return lib


def test_host_library_exists_on_cpu_too():
    assert os.path.exists(os.path.join(ROOT, "lf-vio_b200", "liblfvio_host.so"))


def test_feature_tracker_class_matches_golden(lfvio_mod, gpu, host, golden_dir):
    g = np.load(os.path.join(golden_dir, "frontend_seq_fast.npz"))
    n = int(g["n_frames"])
    src = synth.FrameSource(960, 1280, seed=int(g["seed"]), rot_per_frame=0.004, shift_per_frame=(6.5, -4.2))
    frames = np.ascontiguousarray(np.stack([src.frame(k) for k in range(n)]))
    pub = np.ascontiguousarray(g["pub"], np.int32)
    cfg = tracker_defaults()
    cam = lfvio_mod.Ocam.from_dict(PAL_UP)
    cap = 512
    pts, ids, cnt = np.zeros((cap, 2), np.float32), np.zeros(cap, np.int32), np.zeros(cap, np.int32)
    b, v = np.zeros((cap, 3), np.float32), np.zeros((cap, 3), np.float32)
    n_out, n_id = C.c_int(0), C.c_int(0)
    p = lambda a, t: a.ctypes.data_as(C.POINTER(t))
    host.lfvio_host_tracker_run.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_double, C.c_double,
                                            C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int,
                                            C.c_void_p, C.c_void_p]
    rc = host.lfvio_host_tracker_run(C.byref(cam), frames.ctypes.data, n, 960, 1280, pub.ctypes.data, 100.0, float(g["dt"]),
                                     cfg["max_cnt"], cfg["min_dist"], pts.ctypes.data, ids.ctypes.data, cnt.ctypes.data,
                                     b.ctypes.data, v.ctypes.data, cap, C.byref(n_out), C.byref(n_id))
    assert rc == 0
    m, k = n_out.value, n - 1
    assert np.array_equal(ids[:m], g[f"f{k}_ids"]) and np.array_equal(cnt[:m], g[f"f{k}_track_cnt"])
    assert np.abs(pts[:m] - g[f"f{k}_cur_pts"]).max() < 5e-3
    assert np.abs(b[:m] - g[f"f{k}_cur_un_pts_3d"]).max() < 2e-5
    assert n_id.value == int(g[f"f{k}_ids"].max()) + 1


def _gauge_fix(w0, w):
    """numpy statement of Estimator::double2vector's re-anchoring (estimator.cpp:532-583)."""
    def ypr(q):
        R = synthwin.quat_to_rot(np.array([q[3], q[0], q[1], q[2]]))
        return np.degrees(np.arctan2(R[1, 0], R[0, 0])), R
    y0, _ = ypr(w0.pose[0, 3:7])
    y1, _ = ypr(w.pose[0, 3:7])
    a = np.radians(y0 - y1)
    Rd = np.array([[np.cos(a), -np.sin(a), 0], [np.sin(a), np.cos(a), 0], [0, 0, 1]])
    P = (Rd @ (w.pose[:, :3] - w.pose[0, :3]).T).T + w0.pose[0, :3]
    V = (Rd @ w.sb[:, :3].T).T
    return P, V, Rd


def test_estimator_class_optimization_roundtrip(lfvio_mod, gpu, host):
    w0 = synthwin.make_window(seed=3, opts=lfvio_mod.solver_default_opts(), perturb=0.3)
    ref = w0.clone()
    ob.optimize(ref)
    P_ref, V_ref, Rd = _gauge_fix(w0, ref)
    F, L = w0.F, w0.L
    pose, sb, ex, feat, td = np.zeros((F, 7)), np.zeros((F, 9)), np.zeros(7), np.zeros(L), np.zeros(1)
    prior = B.PriorData()
    it = C.c_int(0)
    st, fl = w0.state(), w0.factors()
    host.lfvio_host_estimator_roundtrip.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int] + [C.c_void_p] * 5 + [C.c_void_p, C.c_void_p]
    rc = host.lfvio_host_estimator_roundtrip(C.byref(st), C.byref(fl), 1, 0, pose.ctypes.data, sb.ctypes.data, ex.ctypes.data,
                                             feat.ctypes.data, td.ctypes.data, C.byref(prior.c), C.byref(it))
    assert rc == 0, lfvio_mod.lib.lfvio_last_error()
    assert it.value > 0
    assert np.abs(pose[:, :3] - P_ref).max() < 1e-6          # positions re-anchored at frame 0
    assert np.abs(sb[:, :3] - V_ref).max() < 1e-5
    assert np.allclose(pose[0, :3], w0.pose[0, :3], atol=1e-9)
    assert np.abs(feat - ref.feat).max() < 1e-6 and abs(td[0] - ref.td[0]) < 1e-7
    # the prior was built at the re-anchored state: compare with the oracle's marginalisation at that state
    g = w0.clone()
    g.pose[:], g.sb[:], g.ex[:], g.feat[:], g.td[:] = pose, sb, ex, feat, td
    po = ob.marginalize(g, B.MARGIN_OLD)
    assert prior.n == po.n == 6 * 10 + 9 + 6 + 1
    Ag, bg = prior.information()
    Ao, bo = po.information()
    assert np.abs(Ag - Ao).max() < 1e-6 * np.abs(Ao).max()


def test_fused_optimize_full_equals_host_class_path(lfvio_mod, gpu, host):
    """lfvio_window_optimize_full (gauge fix on device) == Estimator::optimization() of the host mirror
    (solve, double2vector on the host, marginalise at the re-anchored state)."""
    w0 = synthwin.make_window(seed=4, opts=lfvio_mod.solver_default_opts(), perturb=0.3)
    # make frame 0 itself move during the solve so that the re-anchoring is not a no-op
    w0.pose[0, :3] += 0.02
    F, L = w0.F, w0.L
    pose, sb, ex, feat, td = np.zeros((F, 7)), np.zeros((F, 9)), np.zeros(7), np.zeros(L), np.zeros(1)
    prior_h = B.PriorData()
    it = C.c_int(0)
    st, fl = w0.state(), w0.factors()
    host.lfvio_host_estimator_roundtrip.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int] + [C.c_void_p] * 5 + [C.c_void_p, C.c_void_p]
    assert host.lfvio_host_estimator_roundtrip(C.byref(st), C.byref(fl), 1, 0, pose.ctypes.data, sb.ctypes.data, ex.ctypes.data,
                                               feat.ctypes.data, td.ctypes.data, C.byref(prior_h.c), C.byref(it)) == 0
    win = lfvio_mod.Window(10, 1000, 16384, device=gpu)
    w = w0.clone()
    summ, prior_d = win.optimize_full(w, B.MARGIN_OLD)
    assert summ.iterations == it.value
    assert np.abs(w.pose - pose).max() < 1e-7 and np.abs(w.sb - sb).max() < 1e-6
    assert np.allclose(w.pose[0, :3], w0.pose[0, :3], atol=1e-9)  # frame 0 keeps its position
    assert np.abs(w.feat - feat).max() < 1e-7 and abs(w.td[0] - td[0]) < 1e-9
    assert prior_d.n == prior_h.n and list(prior_d.block_id[: prior_d.n_blocks]) == list(prior_h.block_id[: prior_h.n_blocks])
    Ad, bd = prior_d.information()
    Ah, bh = prior_h.information()
    assert np.abs(Ad - Ah).max() < 1e-6 * np.abs(Ah).max()
    assert np.abs(bd - bh).max() < 1e-5 * max(np.abs(bh).max(), 1.0)
    win.close()


def test_estimator_process_imu_and_triangulate_mirrors(lfvio_mod, gpu, host):
    """Estimator::processIMU (+ flushIMU) and FeatureManager::triangulate of the C++ host mirror give what the C-ABI
    entry points give (which tests/test_prestep_gpu.py holds against the oracle)."""
    import sys

    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from lfvio import prestep as PS
    from prestep_data import G, NOISE, imu_stream, tracks

    s = imu_stream(7, n_intervals=3)
    tr, t = s["traj"], s["times"]
    i = 1
    first, count, a0, g0 = s["intervals"][i]
    P, V, R = tr.pos(t[i]), tr.vel(t[i]), tr.rot(t[i])
    iv = PS.make_interval(first, count, i, a0, g0, s["ba"], s["bg"], P, V, R)
    want, wstate = lfvio_mod.preintegrate([iv], s["dt"], s["acc"], s["gyr"], NOISE, G, device=gpu)
    out = B.ImuFactor()
    state = np.zeros(15)
    dp = lambda a: np.ascontiguousarray(a, np.float64).ctypes.data_as(C.c_void_p)
    host.lfvio_host_process_imu.argtypes = [C.c_int] + [C.c_void_p] * 12 + [C.c_int, C.c_void_p, C.c_void_p]
    sl = slice(first, first + count)
    keep = [np.ascontiguousarray(x, np.float64) for x in (P, V, R, s["ba"], s["bg"], G, NOISE, a0, g0, s["dt"][sl], s["acc"][sl], s["gyr"][sl])]
    rc = host.lfvio_host_process_imu(i + 1, *[k.ctypes.data_as(C.c_void_p) for k in keep], count, C.byref(out), state.ctypes.data_as(C.c_void_p))
    assert rc == 0, lfvio_mod.lib.lfvio_last_error()
    assert out.frame_i == i and out.sum_dt == want[0].sum_dt
    assert np.array_equal(np.array(out.jacobian), np.array(want[0].jacobian))
    assert np.array_equal(np.array(out.covariance), np.array(want[0].covariance))
    assert np.array_equal(np.array(out.delta_p), np.array(want[0].delta_p)) and np.array_equal(state, wstate[0])

    d = tracks(4, n_feat=120)
    depth0 = -np.ones(120)
    want_d = lfvio_mod.triangulate(d["Ps"], d["Rs"], d["tic"], d["ric"], d["start_frame"], d["obs_first"], d["obs_pts"], depth0, 5.0, device=gpu)
    got_d = depth0.copy()
    host.lfvio_host_triangulate.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
    arrs = [np.ascontiguousarray(d["Ps"], np.float64), np.ascontiguousarray(d["Rs"], np.float64), np.ascontiguousarray(d["tic"], np.float64),
            np.ascontiguousarray(d["ric"], np.float64), d["start_frame"], d["obs_first"], np.ascontiguousarray(d["obs_pts"], np.float64)]
    rc = host.lfvio_host_triangulate(arrs[0].ctypes.data, arrs[1].ctypes.data, len(d["Ps"]), arrs[2].ctypes.data, arrs[3].ctypes.data,
                                     arrs[4].ctypes.data, arrs[5].ctypes.data, arrs[6].ctypes.data, 120, got_d.ctypes.data)
    assert rc == 0 and np.array_equal(got_d, want_d) and (got_d > 0).sum() > 40


def test_estimator_class_closed_loop_matches_python_orchestration(lfvio_mod, gpu, host):
    """The C++ host mirror run as an estimator core — processIMU, processImage (addFeatureCheckParallax, triangulate,
    optimization, slideWindow / removeBackShiftDepth) — over the synthetic world of lfvio.violoop, against the Python
    restatement of the same control flow with the same device numeric steps.  MIN_PARALLAX = 0 makes every frame a
    keyframe, which is what the Python loop assumes; td is fixed (the world cuts IMU intervals at the stamps)."""
    from lfvio import violoop

    world = violoop.World(seed=4, n_frames=30)
    win = lfvio_mod.Window(10, 1000, 16384, device=gpu)
    dopts = lfvio_mod.solver_default_opts()
    dopts.row, dopts.estimate_td = 960, 0
    py = violoop.VioLoop(world, violoop.device_ops(lfvio_mod, win, gpu), dopts)
    py.td = 0.0
    for f in py.features:  # (bootstrap observations were stamped with the constructor's td)
        f.obs = [(o[0], o[1], 0.0, o[3]) for o in f.obs]
    boot_pose, boot_sb = py.pose.copy(), py.sb.copy()
    py.run()

    F = 11
    vp = lambda a: np.ascontiguousarray(a, np.float64).ctypes.data_as(C.c_void_p)
    host.lfvio_host_estimator_create.restype = C.c_void_p
    host.lfvio_host_estimator_create.argtypes = [C.c_int, C.c_int, C.c_int]
    host.lfvio_host_estimator_destroy.argtypes = [C.c_void_p]
    host.lfvio_host_estimator_set_params.argtypes = [C.c_void_p] + [C.c_void_p] * 4 + [C.c_double, C.c_double, C.c_int, C.c_void_p, C.c_void_p]
    host.lfvio_host_estimator_process_imu.argtypes = [C.c_void_p] + [C.c_void_p] * 5 + [C.c_int]
    host.lfvio_host_estimator_set_states.argtypes = [C.c_void_p] + [C.c_void_p] * 3 + [C.c_int]
    host.lfvio_host_estimator_process_image.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_double]
    host.lfvio_host_estimator_get.argtypes = [C.c_void_p] + [C.c_void_p] * 7
    est = host.lfvio_host_estimator_create(10, 1000, gpu)
    assert est, lfvio_mod.lib.lfvio_last_error()
    keep = [np.ascontiguousarray(x, np.float64) for x in (world.noise, world.G, world.tic, world.ric, boot_sb[0, 3:6], boot_sb[0, 6:9])]
    assert host.lfvio_host_estimator_set_params(est, vp(keep[0]), vp(keep[1]), vp(keep[2]), vp(keep[3]), 0.0, 0.0, 0, vp(keep[4]), vp(keep[5])) == 0
    P, R, V = np.zeros((F, 3)), np.zeros((F, 9)), np.zeros((F, 3))
    ba, bg, td, info = np.zeros((F, 3)), np.zeros((F, 3)), np.zeros(1), np.zeros(8, np.int32)
    worst = 0.0
    for k in range(world.n_frames):
        if k > 0:
            a0, g0, dt, acc, gyr = world.imu[k]
            arrs = [np.ascontiguousarray(x, np.float64) for x in (a0, g0, dt, acc, gyr)]
            assert host.lfvio_host_estimator_process_imu(est, *[vp(a) for a in arrs], len(dt)) == 0
        if k == F - 1:
            Rb = np.array([violoop._quat_to_rot(boot_pose[f, 3:]).ravel() for f in range(F)])
            assert host.lfvio_host_estimator_set_states(est, vp(boot_pose[:, :3].copy()), vp(Rb), vp(boot_sb[:, :3].copy()), 1) == 0, \
                lfvio_mod.lib.lfvio_last_error()
        obs = world.frames[k]
        ids = np.array(list(obs.keys()), np.int32)
        feat = np.array([[*b, 0.0, row, *vel] for (b, vel, row) in obs.values()])
        rc = host.lfvio_host_estimator_process_image(est, ids.ctypes.data_as(C.c_void_p), vp(feat), len(ids), float(world.times[k]))
        assert rc == 0, lfvio_mod.lib.lfvio_last_error()
        host.lfvio_host_estimator_get(est, vp(P), vp(R), vp(V), vp(ba), vp(bg), vp(td), info.ctypes.data_as(C.c_void_p))
        if k >= F - 1:
            assert info[0] == 0 and info[4] == 0 and info[2] > 0          # keyframe, no failure, the solver ran
            worst = max(worst, np.abs(P[F - 1] - py.estimates[k][0]).max())
    host.lfvio_host_estimator_destroy(est)
    win.close()
    assert info[5] == world.n_frames - (F - 1) and info[6] == 0            # one MARGIN_OLD slide per frame
    assert worst < 1e-4, worst
    ate = violoop.ate_rmse(np.array([py.estimates[k][0] for k in sorted(py.estimates)]),
                           np.array([world.truth(k)[0] for k in sorted(py.estimates)]))
    print(f"C++ estimator closed loop: max |p_cpp - p_python| = {worst:.3e} m, ATE {ate:.4f} m")


def test_estimator_class_non_keyframe_path(lfvio_mod, gpu, host):
    """MARGIN_SECOND_NEW through the C++ mirror: with a huge MIN_PARALLAX every frame after the bootstrap is a
    non-keyframe (estimator.cpp:125-128): the second-newest frame is dropped (removeFront), its IMU samples are merged
    into the previous interval, the prior is marginalised with the SECOND_NEW rule — and the newest pose keeps tracking."""
    from lfvio import violoop

    world = violoop.World(seed=5, n_frames=22)
    F = 11
    rng = np.random.default_rng(3)
    vp = lambda a: np.ascontiguousarray(a, np.float64).ctypes.data_as(C.c_void_p)
    host.lfvio_host_estimator_create.restype = C.c_void_p
    host.lfvio_host_estimator_create.argtypes = [C.c_int, C.c_int, C.c_int]
    host.lfvio_host_estimator_destroy.argtypes = [C.c_void_p]
    host.lfvio_host_estimator_set_params.argtypes = [C.c_void_p] + [C.c_void_p] * 4 + [C.c_double, C.c_double, C.c_int, C.c_void_p, C.c_void_p]
    host.lfvio_host_estimator_process_imu.argtypes = [C.c_void_p] + [C.c_void_p] * 5 + [C.c_int]
    host.lfvio_host_estimator_set_states.argtypes = [C.c_void_p] + [C.c_void_p] * 3 + [C.c_int]
    host.lfvio_host_estimator_process_image.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_double]
    host.lfvio_host_estimator_get.argtypes = [C.c_void_p] + [C.c_void_p] * 7
    est = host.lfvio_host_estimator_create(10, 1000, gpu)
    assert est
    keep = [np.ascontiguousarray(x, np.float64) for x in (world.noise, world.G, world.tic, world.ric, world.ba + 0.005, world.bg + 0.0005)]
    assert host.lfvio_host_estimator_set_params(est, vp(keep[0]), vp(keep[1]), vp(keep[2]), vp(keep[3]), 0.0, 1e9, 0, vp(keep[4]), vp(keep[5])) == 0
    P, R, V = np.zeros((F, 3)), np.zeros((F, 9)), np.zeros((F, 3))
    ba, bg, td, info = np.zeros((F, 3)), np.zeros((F, 3)), np.zeros(1), np.zeros(8, np.int32)
    errs = []
    for k in range(world.n_frames):
        if k > 0:
            a0, g0, dt, acc, gyr = world.imu[k]
            arrs = [np.ascontiguousarray(x, np.float64) for x in (a0, g0, dt, acc, gyr)]
            assert host.lfvio_host_estimator_process_imu(est, *[vp(a) for a in arrs], len(dt)) == 0
        if k == F - 1:
            Pt = np.array([world.truth(f)[0] for f in range(F)]) + rng.normal(0, 0.005, (F, 3))
            Rt = np.array([world.truth(f)[1].ravel() for f in range(F)])
            Vt = np.array([world.truth(f)[2] for f in range(F)]) + rng.normal(0, 0.01, (F, 3))
            assert host.lfvio_host_estimator_set_states(est, vp(Pt), vp(Rt), vp(Vt), 1) == 0
        obs = world.frames[k]
        ids = np.array(list(obs.keys()), np.int32)
        feat = np.array([[*b, 0.0, row, *vel] for (b, vel, row) in obs.values()])
        rc = host.lfvio_host_estimator_process_image(est, ids.ctypes.data_as(C.c_void_p), vp(feat), len(ids), float(world.times[k]))
        assert rc == 0, lfvio_mod.lib.lfvio_last_error()
        host.lfvio_host_estimator_get(est, vp(P), vp(R), vp(V), vp(ba), vp(bg), vp(td), info.ctypes.data_as(C.c_void_p))
        if k >= F:
            assert info[0] == 1 and info[4] == 0 and info[2] > 0   # non-keyframe, no failure, the solver ran
            errs.append(np.linalg.norm(P[F - 1] - world.truth(k)[0]))
    host.lfvio_host_estimator_destroy(est)
    assert info[6] == world.n_frames - (F - 1) and info[5] == 0  # SECOND_NEW from the first solved frame on
    # frames 0..8 of the window never move on and one ever longer IMU interval bridges to the newest frame: the
    # error grows (0.14 -> 0.37 m here) but stays bounded; the point of the test is the code path, not accuracy
    assert max(errs) < 1.0, errs


def _bind_estimator(host):
    host.lfvio_host_estimator_create.restype = C.c_void_p
    host.lfvio_host_estimator_create.argtypes = [C.c_int, C.c_int, C.c_int]
    host.lfvio_host_estimator_destroy.argtypes = [C.c_void_p]
    host.lfvio_host_estimator_set_params.argtypes = [C.c_void_p] + [C.c_void_p] * 4 + [C.c_double, C.c_double, C.c_int, C.c_void_p, C.c_void_p]
    host.lfvio_host_estimator_process_imu.argtypes = [C.c_void_p] + [C.c_void_p] * 5 + [C.c_int]
    host.lfvio_host_estimator_process_image.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_double]
    host.lfvio_host_estimator_get.argtypes = [C.c_void_p] + [C.c_void_p] * 7
    host.lfvio_host_estimator_enable_initialization.argtypes = [C.c_void_p, C.c_uint64, C.c_int]
    host.lfvio_host_estimator_init_info.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]


def _feed(host, est, world, k, vp):
    if k > 0:
        a0, g0, dt, acc, gyr = world.imu[k]
        arrs = [np.ascontiguousarray(x, np.float64) for x in (a0, g0, dt, acc, gyr)]
        assert host.lfvio_host_estimator_process_imu(est, *[vp(a) for a in arrs], len(dt)) == 0
    obs = world.frames[k]
    ids = np.array(list(obs.keys()), np.int32)
    feat = np.array([[*b, 0.0, row, *vel] for (b, vel, row) in obs.values()])
    return host.lfvio_host_estimator_process_image(est, ids.ctypes.data_as(C.c_void_p), vp(feat), len(ids), float(world.times[k]))


@pytest.mark.parametrize("seed,noise", [(1, 0.0), (2, 0.05), (4, 0.1)])
def test_estimator_class_initial_structure_matches_oracle(lfvio_mod, gpu, host, seed, noise):
    """Estimator::initialStructure + visualInitialAlign of the C++ mirror (device numeric steps) against the numpy
    restatement of the same chain (oracle/initial.py): relative pose, global SfM + BA, gyroscope bias, gravity / scale
    alignment, state rescaling and gravity alignment.  Every frame is a key frame (MIN_PARALLAX = 0)."""
    from initial_data import InitScenario
    from oracle import initial as OI

    F = 11
    sc = InitScenario(seed=seed, F=F, pixel_noise=noise)
    world = sc.world
    ref = OI.initial_structure(F, sc.obs_first, sc.obs_frame, sc.obs_pts, sc.corres, lambda bg: sc.preintegrations(bg=bg), sc.ric,
                               sc.tic, sc.g_norm, np.zeros(3), seed=5)
    assert ref is not None
    vp = lambda a: np.ascontiguousarray(a, np.float64).ctypes.data_as(C.c_void_p)
    _bind_estimator(host)
    est = host.lfvio_host_estimator_create(10, 1000, gpu)
    assert est, lfvio_mod.lib.lfvio_last_error()
    keep = [np.ascontiguousarray(x, np.float64) for x in (world.noise, world.G, world.tic, world.ric, np.zeros(3), np.zeros(3))]
    assert host.lfvio_host_estimator_set_params(est, vp(keep[0]), vp(keep[1]), vp(keep[2]), vp(keep[3]), 0.0, 0.0, 0, vp(keep[4]), vp(keep[5])) == 0
    assert host.lfvio_host_estimator_enable_initialization(est, 5, 1) == 0
    for k in range(F):
        assert _feed(host, est, world, k, vp) == 0, lfvio_mod.lib.lfvio_last_error()
    P, R, V = np.zeros((F, 3)), np.zeros((F, 9)), np.zeros((F, 3))
    ba, bg, td, info, ii, g = np.zeros((F, 3)), np.zeros((F, 3)), np.zeros(1), np.zeros(8, np.int32), np.zeros(8, np.int32), np.zeros(3)
    host.lfvio_host_estimator_get(est, vp(P), vp(R), vp(V), vp(ba), vp(bg), vp(td), info.ctypes.data_as(C.c_void_p))
    host.lfvio_host_estimator_init_info(est, ii.ctypes.data_as(C.c_void_p), vp(g))
    host.lfvio_host_estimator_destroy(est)
    assert ii[0] == 1 and ii[1] == 1 and ii[2] == 0 and ii[3] == 1, ii    # NON_LINEAR after the first attempt
    tol = 1e-6 if noise == 0.0 else 1e-4
    assert np.allclose(g, ref["g"], atol=tol) and np.allclose(bg, ref["Bgs"], atol=tol)
    assert np.allclose(P, ref["Ps"], atol=tol) and np.allclose(R.reshape(F, 3, 3), ref["Rs"], atol=tol)
    assert np.allclose(V, ref["Vs"], atol=10 * tol)
    # and against the ground truth (frame 0 at the origin with zero yaw)
    Rz = OI._yaw_rot(-OI._R2ypr_yaw_deg(sc.Rb[0]))
    assert np.abs(P - (sc.P - sc.P[0]) @ Rz.T).max() < 0.02
    assert np.abs(V - sc.V @ Rz.T).max() < 0.05
    print(f"initialStructure seed {seed}: max |P - oracle| = {np.abs(P - ref['Ps']).max():.2e}, max |P - truth| = "
          f"{np.abs(P - (sc.P - sc.P[0]) @ Rz.T).max():.4f} m, scale {ref['s']:.4f}")


def test_estimator_class_closed_loop_from_cold_start(lfvio_mod, gpu, host):
    """No injected initial state: the C++ mirror initialises itself (initialStructure) when the window is full, then runs
    the sliding-window optimisation; the trajectory is compared with the ground truth after a rigid alignment (ATE)."""
    from lfvio import synthwin, violoop

    E = synthwin.ESTIMATOR
    old_td = E["td"]
    E["td"] = 0.0
    try:
        world = violoop.World(seed=2, n_frames=40, pixel_noise=0.1)
    finally:
        E["td"] = old_td
    F = 11
    vp = lambda a: np.ascontiguousarray(a, np.float64).ctypes.data_as(C.c_void_p)
    _bind_estimator(host)
    est = host.lfvio_host_estimator_create(10, 1000, gpu)
    assert est, lfvio_mod.lib.lfvio_last_error()
    keep = [np.ascontiguousarray(x, np.float64) for x in (world.noise, world.G, world.tic, world.ric, np.zeros(3), np.zeros(3))]
    assert host.lfvio_host_estimator_set_params(est, vp(keep[0]), vp(keep[1]), vp(keep[2]), vp(keep[3]), 0.0, 0.0, 0, vp(keep[4]), vp(keep[5])) == 0
    assert host.lfvio_host_estimator_enable_initialization(est, 9, 0) == 0
    P, R, V = np.zeros((F, 3)), np.zeros((F, 9)), np.zeros((F, 3))
    ba, bg, td, info, ii, g = np.zeros((F, 3)), np.zeros((F, 3)), np.zeros(1), np.zeros(8, np.int32), np.zeros(8, np.int32), np.zeros(3)
    est_pos, true_pos = [], []
    for k in range(world.n_frames):
        assert _feed(host, est, world, k, vp) == 0, lfvio_mod.lib.lfvio_last_error()
        host.lfvio_host_estimator_get(est, vp(P), vp(R), vp(V), vp(ba), vp(bg), vp(td), info.ctypes.data_as(C.c_void_p))
        host.lfvio_host_estimator_init_info(est, ii.ctypes.data_as(C.c_void_p), vp(g))
        if k >= F - 1:
            assert ii[0] == 1 and ii[1] == 1, (k, ii)                     # initialised at the first full window
            assert info[4] == 0                                            # no failure detected
            est_pos.append(P[F - 1].copy())                                # (after the slide the newest state sits in the last slot)
            true_pos.append(world.truth(k)[0])
    host.lfvio_host_estimator_destroy(est)
    assert abs(g[2] - 9.81007) < 1e-3 and np.abs(g[:2]).max() < 1e-6      # gravity aligned with z
    ate = violoop.ate_rmse(np.array(est_pos), np.array(true_pos))
    print(f"cold start: ATE over {len(est_pos)} frames = {ate:.4f} m")
    assert ate < 0.05, ate


@pytest.mark.parametrize("device_store", [0, 1])
def test_estimator_class_failure_reboot_and_reinitialisation(lfvio_mod, gpu, host, device_store):
    """The restart path (estimator.cpp:196-204): failureDetection() trips (here: last_P moved by 10 m), the
    estimator reboots (clearState + setParameter), the window refills and initialStructure() runs a second time.  With the
    FeatureManager on the device the reboot goes back to the host-side list for the initialisation and loads the store again."""
    from lfvio import synthwin, violoop

    E = synthwin.ESTIMATOR
    old_td = E["td"]
    E["td"] = 0.0
    try:
        world = violoop.World(seed=4, n_frames=50, pixel_noise=0.1)
    finally:
        E["td"] = old_td
    F = 11
    vp = lambda a: np.ascontiguousarray(a, np.float64).ctypes.data_as(C.c_void_p)
    _bind_estimator(host)
    host.lfvio_host_estimator_poke_last_p.argtypes = [C.c_void_p, C.c_double]
    est = host.lfvio_host_estimator_create(10, 1000, gpu)
    assert est, lfvio_mod.lib.lfvio_last_error()
    keep = [np.ascontiguousarray(x, np.float64) for x in (world.noise, world.G, world.tic, world.ric, np.zeros(3), np.zeros(3))]
    assert host.lfvio_host_estimator_set_params(est, vp(keep[0]), vp(keep[1]), vp(keep[2]), vp(keep[3]), 0.0, 0.0, 0, vp(keep[4]), vp(keep[5])) == 0
    assert host.lfvio_host_estimator_enable_initialization(est, 9, 0) == 0
    host.lfvio_host_estimator_use_device_store.argtypes = [C.c_void_p, C.c_int]
    host.lfvio_host_estimator_use_device_store(est, device_store)
    P, R, V = np.zeros((F, 3)), np.zeros((F, 9)), np.zeros((F, 3))
    ba, bg, td, info, ii, g = np.zeros((F, 3)), np.zeros((F, 3)), np.zeros(1), np.zeros(8, np.int32), np.zeros(8, np.int32), np.zeros(3)
    flags, est_pos, true_pos = [], [], []
    k_fail = 20
    for k in range(world.n_frames):
        if k == k_fail:
            host.lfvio_host_estimator_poke_last_p(est, 10.0)
        assert _feed(host, est, world, k, vp) == 0, lfvio_mod.lib.lfvio_last_error()
        host.lfvio_host_estimator_get(est, vp(P), vp(R), vp(V), vp(ba), vp(bg), vp(td), info.ctypes.data_as(C.c_void_p))
        host.lfvio_host_estimator_init_info(est, ii.ctypes.data_as(C.c_void_p), vp(g))
        flags.append((int(ii[0]), int(ii[6]), int(info[3])))
        if k >= k_fail + F and ii[0] == 1:
            est_pos.append(P[F - 1].copy())
            true_pos.append(world.truth(k)[0])
    assert host.lfvio_host_estimator_use_device_store(est, device_store) == device_store   # the store is live again after the reboot
    host.lfvio_host_estimator_destroy(est)
    assert all(f[0] == 1 and f[1] == 0 for f in flags[F - 1:k_fail])          # running, no reboot yet
    assert flags[k_fail] == (0, 1, 0)                                         # rebooted: INITIAL, frame_count 0
    assert all(f[0] == 0 for f in flags[k_fail:k_fail + F])                   # the window refills (frames k_fail+1 .. k_fail+F)
    assert flags[k_fail + F][0] == 1 and ii[1] == 2 and ii[6] == 1            # second initialisation at the first full window
    ate = violoop.ate_rmse(np.array(est_pos), np.array(true_pos))
    print(f"reboot at frame {k_fail}: re-initialised at frame {k_fail + F}, ATE of the {len(est_pos)} frames after it = {ate:.4f} m")
    assert len(est_pos) == world.n_frames - (k_fail + F) and ate < 0.05, ate


@pytest.mark.parametrize("device_store", [0, 1])
def test_estimator_class_survives_a_failed_solve(lfvio_mod, gpu, host, device_store):
    """A solve that fails numerically is not fatal (the reference's ceres::Solve leaves the state at the last accepted point
    and processImage goes on to failureDetection / slideWindow, estimator.cpp:183-215): a pre-integration with a singular
    covariance makes lfvio_window_optimize return LFVIO_ERR_NUMERIC for as long as it stays in the window; the window must keep
    sliding (frame_count stays WINDOW_SIZE, tracks do not outgrow it) and the estimator must track again afterwards."""
    from lfvio import synthwin, violoop

    E = synthwin.ESTIMATOR
    old_td = E["td"]
    E["td"] = 0.0
    try:
        world = violoop.World(seed=4, n_frames=45, pixel_noise=0.1)
    finally:
        E["td"] = old_td
    F = 11
    vp = lambda a: np.ascontiguousarray(a, np.float64).ctypes.data_as(C.c_void_p)
    _bind_estimator(host)
    host.lfvio_host_estimator_poke_covariance.argtypes = [C.c_void_p, C.c_int]
    est = host.lfvio_host_estimator_create(10, 1000, gpu)
    assert est, lfvio_mod.lib.lfvio_last_error()
    keep = [np.ascontiguousarray(x, np.float64) for x in (world.noise, world.G, world.tic, world.ric, np.zeros(3), np.zeros(3))]
    assert host.lfvio_host_estimator_set_params(est, vp(keep[0]), vp(keep[1]), vp(keep[2]), vp(keep[3]), 0.0, 0.0, 0, vp(keep[4]), vp(keep[5])) == 0
    assert host.lfvio_host_estimator_enable_initialization(est, 9, 0) == 0
    host.lfvio_host_estimator_use_device_store.argtypes = [C.c_void_p, C.c_int]
    host.lfvio_host_estimator_use_device_store(est, device_store)  # (a failed solve then leaves the store's depths as they were)
    P, R, V = np.zeros((F, 3)), np.zeros((F, 9)), np.zeros((F, 3))
    ba, bg, td, info = np.zeros((F, 3)), np.zeros((F, 3)), np.zeros(1), np.zeros(8, np.int32)
    k_poke, failures, est_pos, true_pos = 20, [], [], []
    for k in range(world.n_frames):
        if k == k_poke:
            host.lfvio_host_estimator_poke_covariance(est, 3)
        assert _feed(host, est, world, k, vp) == 0, lfvio_mod.lib.lfvio_last_error()   # never a fatal status
        failures.append(host.lfvio_host_estimator_poke_covariance(est, -1))
        host.lfvio_host_estimator_get(est, vp(P), vp(R), vp(V), vp(ba), vp(bg), vp(td), info.ctypes.data_as(C.c_void_p))
        if k >= F - 1:
            assert info[3] == F - 1, (k, info[3])                                          # the window kept sliding
        if k >= k_poke + 8:
            est_pos.append(P[F - 1].copy())
            true_pos.append(world.truth(k)[0])
    host.lfvio_host_estimator_destroy(est)
    assert failures[k_poke - 1] == 0 and failures[k_poke] >= 1                             # the poisoned solve failed and was survived
    assert failures[-1] == failures[k_poke + 6]                                            # and stopped failing once the interval left the window
    ate = violoop.ate_rmse(np.array(est_pos), np.array(true_pos))
    print(f"{failures[-1]} failed solves survived; ATE of the {len(est_pos)} frames afterwards = {ate:.4f} m")
    assert np.isfinite(P).all() and ate < 0.5, ate


def _run_class_stream(lfvio_mod, gpu, host, world, min_parallax, device_store, cold_start, estimate_td=0):
    """One pass of a synthetic stream through the C++ Estimator mirror; returns per solved frame (P, V, ba, bg, td, info)."""
    F = 11
    rng = np.random.default_rng(3)
    vp = lambda a: np.ascontiguousarray(a, np.float64).ctypes.data_as(C.c_void_p)
    _bind_estimator(host)
    host.lfvio_host_estimator_set_states.argtypes = [C.c_void_p] + [C.c_void_p] * 3 + [C.c_int]
    host.lfvio_host_estimator_use_device_store.argtypes = [C.c_void_p, C.c_int]
    est = host.lfvio_host_estimator_create(10, 1000, gpu)
    assert est, lfvio_mod.lib.lfvio_last_error()
    b0 = (np.zeros(3), np.zeros(3)) if cold_start else (world.ba + 0.005, world.bg + 0.0005)
    keep = [np.ascontiguousarray(x, np.float64) for x in (world.noise, world.G, world.tic, world.ric, *b0)]
    assert host.lfvio_host_estimator_set_params(est, vp(keep[0]), vp(keep[1]), vp(keep[2]), vp(keep[3]), 0.0, min_parallax, estimate_td, vp(keep[4]), vp(keep[5])) == 0
    if cold_start:
        assert host.lfvio_host_estimator_enable_initialization(est, 9, 0) == 0
    assert host.lfvio_host_estimator_use_device_store(est, 1 if device_store else 0) == 0
    P, R, V = np.zeros((F, 3)), np.zeros((F, 9)), np.zeros((F, 3))
    ba, bg, td, info = np.zeros((F, 3)), np.zeros((F, 3)), np.zeros(1), np.zeros(8, np.int32)
    out = []
    for k in range(world.n_frames):
        if not cold_start and k == F - 1:
            if k > 0:
                a0, g0, dt, acc, gyr = world.imu[k]
                arrs = [np.ascontiguousarray(x, np.float64) for x in (a0, g0, dt, acc, gyr)]
                assert host.lfvio_host_estimator_process_imu(est, *[vp(a) for a in arrs], len(dt)) == 0
            Pt = np.array([world.truth(f)[0] for f in range(F)]) + rng.normal(0, 0.005, (F, 3))
            Rt = np.array([world.truth(f)[1].ravel() for f in range(F)])
            Vt = np.array([world.truth(f)[2] for f in range(F)]) + rng.normal(0, 0.01, (F, 3))
            assert host.lfvio_host_estimator_set_states(est, vp(Pt), vp(Rt), vp(Vt), 1) == 0
            obs = world.frames[k]
            ids = np.array(list(obs.keys()), np.int32)
            feat = np.array([[*b, 0.0, row, *vel] for (b, vel, row) in obs.values()])
            rc = host.lfvio_host_estimator_process_image(est, ids.ctypes.data_as(C.c_void_p), vp(feat), len(ids), float(world.times[k]))
        else:
            rc = _feed(host, est, world, k, vp)
        assert rc == 0, lfvio_mod.lib.lfvio_last_error()
        host.lfvio_host_estimator_get(est, vp(P), vp(R), vp(V), vp(ba), vp(bg), vp(td), info.ctypes.data_as(C.c_void_p))
        if k >= F - 1:
            out.append((P.copy(), V.copy(), ba.copy(), bg.copy(), float(td[0]), info.copy()))
    live = host.lfvio_host_estimator_use_device_store(est, 1 if device_store else 0)
    host.lfvio_host_estimator_destroy(est)
    return out, live


@pytest.mark.parametrize("min_parallax,cold_start", [(0.0, False), (0.45, False), (1e9, False), (0.0, True)])
def test_estimator_class_with_device_feature_store(lfvio_mod, gpu, host, min_parallax, cold_start):
    """The C++ Estimator with its FeatureManager on the device (use_device_store) against the same class with the host-side
    list, frame by frame: same key-frame decisions, same list length, same iteration counts, states equal to within the solve's
    run-to-run spread (tests/test_feature_store_gpu.py measures it).  Key frames only, a mixture, non-key frames only, and the
    cold start where the list moves to the device right after initialStructure()."""
    from lfvio import synthwin, violoop

    if cold_start:
        E = synthwin.ESTIMATOR
        old_td = E["td"]
        E["td"] = 0.0
        try:
            world = violoop.World(seed=2, n_frames=40, pixel_noise=0.1)
        finally:
            E["td"] = old_td
    else:
        world = violoop.World(seed=6, n_frames=34, max_feat=60)
    a, live_a = _run_class_stream(lfvio_mod, gpu, host, world, min_parallax, False, cold_start)
    b, live_b = _run_class_stream(lfvio_mod, gpu, host, world, min_parallax, True, cold_start)
    assert live_a == 0 and live_b == 1 and len(a) == len(b)
    flags = []
    for i, (x, y) in enumerate(zip(a, b)):
        assert x[5][0] == y[5][0] and x[5][1] == y[5][1] and x[5][2] == y[5][2], (i, x[5], y[5])  # flag, list length, iterations
        assert x[5][5] == y[5][5] and x[5][6] == y[5][6]                                           # sum_of_back / sum_of_front
        for u, v in zip(x[:4], y[:4]):
            assert np.abs(u - v).max() < 2e-5, (i, np.abs(u - v).max())
        flags.append(int(x[5][0]))
    if min_parallax == 0.45:
        assert 0 in flags[1:] and 1 in flags[1:], flags
