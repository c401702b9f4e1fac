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
