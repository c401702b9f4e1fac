"""Closed loop THROUGH PIXELS (north_star "ATE ... inside the reference's own run-to-run band"):
a textured 3-D room with a few independently moving objects is rendered through the Scaramuzza model along a hand-held
trajectory (lfvio/panoworld.py); the images go through FeatureTracker::readImage (CLAHE -> LK -> rejectWithF -> setMask ->
Shi-Tomasi -> lift), the published tracks (track_cnt > 1: id, bearing, bearing velocity, pixel row) through the sliding-window
estimator (pre-integration -> triangulate -> optimization -> marginalisation -> slideWindow), frame after frame.

  device arm : the tracker and every numeric estimator step through the C-ABI on the GPU
  CPU arm    : cv2 (the library the reference calls) + the reference's own RANSAC code (oracle/_ref) or its numpy restatement
               + the fp64 oracle back end, same orchestration
  band       : the CPU arm run with >= 5 RANSAC seeds (the reference seeds its RANSAC from std::random_device)

and, separately, the drop-in path: FeatureTrackerNode::img_callback -> PointCloud message -> EstimatorNode::process() of the C++
host mirrors (cold start: the estimator initialises itself with initialStructure).
"""
import ctypes as C
import os

import numpy as np
import pytest

from conftest import ROOT
from lfvio import panoworld, violoop
from oracle import backend as ob
from oracle import frontend as fe
from oracle import loop_ops
from oracle import ransac as R
from oracle import ref

pytestmark = pytest.mark.gpu
SEEDS = (1, 2, 3, 4, 5)


def cpu_ransac_hook(seed):
    use_ref = ref.available(build=False)

    def hook(tr, cur, forw):
        if len(forw) < 8:
            return np.ones(len(forw), np.uint8)
        if use_ref:  # compute_E_21 / check_inliers / myfindFundamentalMat of feature_tracker.cpp, compiled unmodified
            b1, b2 = R.lift_pair(tr.cam, cur, forw)
            return ref.find_fundamental(b1, b2, seed, tr.frames_seen)[1]
        return R.reject_with_f(tr.cam, cur, forw, seed, tr.frames_seen)

    return hook


@pytest.fixture(scope="module")
def world(lfvio_mod):
    cam, tcfg = lfvio_mod.PAL_UP, dict(lfvio_mod.tracker_defaults(), max_cnt=150)
    return panoworld.PixelWorld(cam, tcfg, seed=1, n_frames=40)


def test_device_ate_inside_the_cpu_seed_band(lfvio_mod, gpu, world):
    cam, tcfg = world.cam, world.tcfg
    o_cpu = ob.default_opts()
    o_cpu.row = 960
    cpu = []
    for s in SEEDS:
        world.track(fe.FeatureTrackerOracle(cam, tcfg, "cv2", ransac=cpu_ransac_hook(s)))
        loop = violoop.VioLoop(world, loop_ops.oracle_ops(), o_cpu).run()
        cpu.append(violoop.trajectory_error(loop)[0])
    cpu_tracks = list(world.n_tracked)
    o_dev = lfvio_mod.solver_default_opts()
    o_dev.row = 960
    win = lfvio_mod.Window(10, 1000, 16384, device=gpu)
    dev = []
    for s in SEEDS:
        tr = lfvio_mod.FeatureTracker(cam, dict(tcfg, ransac=1, ransac_seed=s), device=gpu)
        world.track(tr)
        tr.close()
        loop = violoop.VioLoop(world, violoop.device_ops(lfvio_mod, win, gpu), o_dev).run()
        dev.append(violoop.trajectory_error(loop)[0])
    win.close()
    assert min(world.n_tracked[2:]) > 100 and min(cpu_tracks[2:]) > 100   # both trackers hold the feature budget
    lo, hi = min(cpu), max(cpu)
    # the band of this scene is narrow (clear-cut outliers: every seed converges to the same inlier set), so it is widened by
    # 5 % of the CPU arm's ATE: sub-pixel LK differences (<= 2e-3 px) move the estimate by that order
    tol = 0.05 * float(np.mean(cpu))
    assert all(lo - tol <= a <= hi + tol for a in dev), (dev, cpu)
    assert max(dev) < 0.25  # and the estimate is a trajectory, not a drift (5.5 m path)


def test_node_path_from_pixels_cold_start(lfvio_mod, gpu, world, tmp_path):
    """FeatureTrackerNode::img_callback -> PointCloudMsg -> EstimatorNode::process(), no injected state."""
    host = C.CDLL(os.path.join(ROOT, "lf-vio_b200", "liblfvio_host.so"))
    host.lfvio_host_tracker_node_create.restype = C.c_void_p
    host.lfvio_host_tracker_node_destroy.argtypes = [C.c_void_p]
    host.lfvio_host_tracker_node_configure.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_uint64]
    host.lfvio_host_tracker_node_image.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_double, C.c_void_p, C.c_int]
    host.lfvio_host_estimator_node_create.restype = C.c_void_p
    host.lfvio_host_estimator_node_create.argtypes = [C.c_int, C.c_int, C.c_int, C.c_char_p]
    host.lfvio_host_estimator_node_destroy.argtypes = [C.c_void_p]
    host.lfvio_host_estimator_node_estimator.restype = C.c_void_p
    host.lfvio_host_estimator_node_estimator.argtypes = [C.c_void_p]
    host.lfvio_host_estimator_node_imu.argtypes = [C.c_void_p, C.c_double, C.c_void_p, C.c_void_p]
    host.lfvio_host_estimator_node_feature.argtypes = [C.c_void_p, C.c_double, C.c_void_p, C.c_int]
    host.lfvio_host_estimator_node_process.argtypes = [C.c_void_p]
    host.lfvio_host_estimator_set_params.argtypes = [C.c_void_p] + [C.c_void_p] * 4 + [C.c_double, C.c_double, C.c_int, C.c_void_p, C.c_void_p]
    host.lfvio_host_estimator_enable_initialization.argtypes = [C.c_void_p, C.c_uint64, C.c_int]
    w = world
    result = str(tmp_path / "traj.txt")
    tnode = host.lfvio_host_tracker_node_create()
    ocam = lfvio_mod.Ocam.from_dict(w.cam)
    # FREQ = 100: the stream is already at the publish rate, every frame is published (first one suppressed, :173-178)
    assert host.lfvio_host_tracker_node_configure(tnode, C.byref(ocam), 960, 1280, 150, w.tcfg["min_dist"], 100, 1, 3) == 0
    enode = host.lfvio_host_estimator_node_create(10, 1000, gpu, result.encode())
    assert enode, lfvio_mod.lib.lfvio_last_error()
    est = host.lfvio_host_estimator_node_estimator(enode)
    vp = lambda a: np.ascontiguousarray(a, np.float64).ctypes.data_as(C.c_void_p)
    keep = [np.ascontiguousarray(x, np.float64) for x in (w.noise, w.G, w.tic, w.ric, np.zeros(3), np.zeros(3))]
    assert host.lfvio_host_estimator_set_params(est, vp(keep[0]), vp(keep[1]), vp(keep[2]), vp(keep[3]), 0.0, 0.0, 1, vp(keep[4]), vp(keep[5])) == 0
    assert host.lfvio_host_estimator_enable_initialization(est, 9, 0) == 0
    out = np.zeros((512, 9), np.float32)
    host.lfvio_host_estimator_node_imu(enode, float(w.times[0]), vp(w.imu[1][0]), vp(w.imu[1][1]))
    n_sub, published, processed = len(w.imu[1][2]), 0, 0
    for k in range(w.n_frames):
        if k > 0:
            _, _, dt, acc, gyr = w.imu[k]
            for j in range(n_sub):
                stamp = float(w.times[k]) if j == n_sub - 1 else float(w.times[k - 1] + w.frame_dt * (j + 1) / n_sub)
                host.lfvio_host_estimator_node_imu(enode, stamp, vp(acc[j]), vp(gyr[j]))
        img = np.ascontiguousarray(w.images[k])
        r = host.lfvio_host_tracker_node_image(tnode, img.ctypes.data, 960, 1280, 1280, float(w.times[k]), out.ctypes.data, 512)
        if r >= 0:
            published += 1
            host.lfvio_host_estimator_node_feature(enode, float(w.times[k]), out.ctypes.data, r)
        processed += host.lfvio_host_estimator_node_process(enode)
    host.lfvio_host_tracker_node_destroy(tnode)
    host.lfvio_host_estimator_node_destroy(enode)
    assert published >= w.n_frames - 3 and processed >= published - 2
    tum = np.array([[float(x) for x in ln.split()] for ln in open(result).read().strip().split("\n")])
    assert len(tum) >= 15
    ks = [int(round((t - w.times[0]) / w.frame_dt)) for t in tum[:, 0]]
    truth = np.array([w.truth(k)[0] for k in ks])
    ate = violoop.ate_rmse(tum[:, 1:4], truth)
    assert ate < 0.25, ate
