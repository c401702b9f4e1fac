#!/usr/bin/env python
"""bench.py — VIO frames/sec (track + optimize, 10-keyframe window) on B200, next to the CPU path.

One "step" = one camera frame through the hot path BASELINE.json's north_star names:
  front end  FeatureTracker::readImage (CLAHE -> pyramidal LK -> border filter -> rejectWithF -> setMask -> Shi-Tomasi -> lift)
  back end   Estimator::optimization   (8 dogleg iterations on an 11-frame window with the steady-state marginalisation
             prior, DENSE_SCHUR semantics, then the MARGIN_OLD prior for the next frame)
Workload (config.workload): BASELINE.json configs[1] shape — synthetic 1280x960 panoramic frames @15 Hz published
at 10 Hz, 150 features, 200 Hz IMU, 10-keyframe window, estimate_td = 1.  One independent stream per GPU (weak
scaling, no data-path collective; NCCL only reduces the timing counters).

JSON keys follow the driver contract:
  value        frames/s with the inputs resident in HBM; the two halves run on their own CUDA streams, enqueued by two
               host threads (the reference's two nodes run as two processes: a frame's tracking overlaps the previous
               frame's optimisation); each thread closes its stream's CUDA-event bracket itself; max over ranks.
               `conventions` gives both ways of counting for BOTH arms: pipelined = 1 / max(front, back) (what `value`,
               `e2e` and the reference arm's `value` are), serial = 1 / (front + back)
  e2e          the same metric through the blocking C-ABI calls with HOST buffers (pinned frame H2D, problem upload,
               results D2H inside the timed region), two host threads; node_ms_per_frame / single_thread_ms_per_frame
  roofline     dominant stage by live CUDA-event time; achieved = algorithmic bytes / average launch time against the
               measured HBM peak, `latency_bound` + `fp64` give the yardstick that fits a one-CTA dependent chain;
               traffic from the committed ncu capture; per_stage lists every stage with a byte count
  cpu_baseline bounded sample of the CPU arm on this box's cores (N = 1 only): cv2 front end + rejectWithF (the
               reference's own RANSAC code from oracle/_ref when that library travelled, else the numpy restatement) +
               fp64 oracle back end, run the way the reference runs (two threads) and serially
  ate          closed loop THROUGH PIXELS (lfvio/panoworld.py: a textured room rendered through the Scaramuzza model):
               tracker -> published tracks -> sliding-window estimator, device arm next to the CPU arm over 5 RANSAC seeds
  closed_loop  the C++ node mirrors on the same rendered stream (FeatureTrackerNode -> PointCloud -> EstimatorNode, cold
               start): frames/s of the whole drop-in path where every solved window is the one the tracker just fed
  clocks, gpu_launches, host_enqueue_ms_per_step, per_rank
`--impl reference` times the reference's CPU path in a process that never maps the product library (LFVIO_TYPES_ONLY):
cv2 with the reference's arguments + its RANSAC + the fp64 C++ oracle of the back end (the reference itself —
ROS/OpenCV-C++/Ceres/Eigen — cannot be built in this image).
`--config` selects BASELINE configs 3 / 4 (parity-test sizes; their lines are kept under profiles/).
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "lf-vio_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

METRIC = "VIO frames/sec (track+optimize, 10-kf window)"
UNIT = "frames/s"
CONFIGS = {
    # name: (scale, max_cnt, window_size, n_landmarks for the synthetic window, camera Hz)
    "palvio_id01_1280x960_150feat_10kf": (1.0, 150, 10, 260, 15.0),
    "synthetic_1920x1440_500feat_10kf": (1.5, 500, 10, 900, 30.0),
    "synthetic_3840x2880_2000feat_20kf": (3.0, 2000, 20, 3600, 60.0),
}
ATE_SCENARIO = ("closed-loop sliding window (processIMU -> triangulate -> optimization -> MARGIN_OLD -> slideWindow) on a synthetic "
                "hand-held trajectory: 150 bearings/frame at 0.3 px (f = 160), 10 Hz keyframes, 200 Hz IMU with the PAL noise values, "
                "estimate_td = 1, bootstrap from perturbed ground truth; PALVIO itself is not available offline")
PUB_PATTERN = (1, 1, 0)  # 15 Hz camera published at 10 Hz (feature_tracker_node.cpp:50-62)


def measured_peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))), "measured"
    except Exception:
        return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0}, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md clocks line).  The sampler is
    started before the warm-up (nvidia-smi needs some hundred ms to produce its first line) and only the samples whose
    timestamp falls inside [mark_begin, mark_end] are used; if the region is shorter than the sampling period the
    samples taken under load since the warm-up are used instead and "window" says so."""

    Q = ("timestamp,index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device: int):
        self.path = tempfile.mktemp(prefix="lfvio_clocks_", suffix=".csv")
        self.proc = None
        self.device = device
        self.t0 = self.t1 = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.device)], stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def wait_ready(self, timeout=5.0):
        t_end = time.time() + timeout
        while self.proc is not None and time.time() < t_end:
            try:
                if os.path.getsize(self.path) > 0:
                    return
            except OSError:
                pass
            time.sleep(0.02)

    def mark_begin(self):
        self.t0 = time.time()

    def mark_end(self):
        self.t1 = time.time()

    def stop(self) -> dict:
        import datetime

        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        rows = []
        try:
            for line in open(self.path):
                f = [x.strip() for x in line.split(",")]
                if len(f) < 10:
                    continue
                try:
                    ts = datetime.datetime.strptime(f[0], "%Y/%m/%d %H:%M:%S.%f").timestamp()
                    rows.append((ts, float(f[2]), float(f[3]), f[6:10]))
                except ValueError:
                    continue
            os.unlink(self.path)
        except Exception:
            pass
        inside = [r for r in rows if self.t0 is not None and self.t1 is not None and self.t0 <= r[0] <= self.t1]
        window = "timed region"
        if not inside:
            inside = [r for r in rows if self.t1 is None or r[0] <= self.t1 + 0.05][-25:]
            window = "warm-up + timed region (timed region shorter than the sampling period)"
        reasons = set()
        for r in inside:
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if inside:
            out.update(sm_mhz=float(np.median([r[1] for r in inside])), sm_max_mhz=float(max(r[2] for r in inside)),
                       reasons=sorted(reasons), samples=len(inside), window=window)
        return out


def build_workload(cfg_name: str, rank: int, n_frames_pool: int):
    from lfvio import synth, synthwin
    from lfvio.config import scaled_camera, tracker_defaults

    scale, max_cnt, wsz, n_lm, hz = CONFIGS[cfg_name]
    cam = scaled_camera(scale)
    # rejectWithF runs on every published frame in the reference (feature_tracker.cpp:149); the device arm does the same.
    # FeatureTrackerOracle (CPU arm) ignores these two keys: it has no native RANSAC port, so the CPU baseline does LESS work.
    tcfg = dict(tracker_defaults(scale, max_cnt), ransac=1, ransac_seed=rank + 1)
    rows, cols = cam["image_height"], cam["image_width"]
    src = synth.FrameSource(rows, cols, seed=rank + 1)  # lfvio.multi.stream_seed: one independent stream per rank
    frames = np.stack([src.frame(k) for k in range(n_frames_pool)])
    return dict(cam=cam, tcfg=tcfg, rows=rows, cols=cols, frames=frames, hz=hz, window_size=wsz, n_landmarks=n_lm,
                max_cnt=max_cnt, synthwin=synthwin)


def make_window_problem(wl, opts, rank, optimize, marginalize):
    """Steady-state Estimator::optimization() input: projection + IMU factors AND the marginalisation prior left by the
    previous frame (n = 76 at F = 11), produced by `optimize` / `marginalize` of the arm that is being timed."""
    # Weak scaling = the SAME work on every GPU: every rank solves the window of seed 1 (the image streams differ per rank).
    # Round 1 seeded the window with the rank: windows whose trust-region steps are all accepted factor the reduced system 8
    # times, others re-use it after a rejected step, and the max over ranks then measured that lottery (step_kernel 51 us per
    # launch on two ranks, 88 us on the other two — the whole of round 1's "0.65 scaling efficiency").
    del rank
    return wl["synthwin"].make_window_with_prior(optimize, marginalize, seed=1, n_frames=wl["window_size"] + 1,
                                                 n_landmarks=wl["n_landmarks"], feats_per_frame=wl["max_cnt"], opts=opts,
                                                 perturb=0.3)


# ---------------------------------------------------------------------------------------------------
# CPU arm (reference path): cv2 front end + fp64 oracle back end
# ---------------------------------------------------------------------------------------------------
def cpu_ransac_hook(seed):
    """rejectWithF for the CPU arm: the reference's own compute_E_21 / check_inliers / myfindFundamentalMat
    (feature_tracker.cpp:186-327, oracle/_ref) when that library is there, else the numpy restatement."""
    from oracle import ransac as R
    from oracle import ref

    use_ref = ref.available(build=False)

    def hook(tr, cur, forw):
        if len(forw) < 8:
            return np.ones(len(forw), np.uint8)
        if use_ref:
            b1, b2 = R.lift_pair(tr.cam, cur, forw)
            return ref.find_fundamental(b1, b2, seed, tr.frames_seen)[1]
        return R.reject_with_f(tr.cam, cur, forw, seed, tr.frames_seen)

    return hook, ("reference code (oracle/_ref, stand-in Eigen)" if use_ref else "numpy restatement (oracle/_ref missing: slow)")


def cpu_arm(wl, steps: int, warmup: int, rank: int, with_marg: bool):
    """The CPU path on this box's cores.  Front end: cv2 (all its threads) + rejectWithF + the restated control flow; back end:
    fp64 C++ oracle, 1 thread solve + 4 threads marginalisation (as the reference).  Timed twice: the two halves in two threads
    (the reference's two nodes; cv2 and the oracle both release the GIL) and one after the other."""
    import threading

    import cv2

    from oracle import backend as ob
    from oracle import frontend as fe

    opts = ob.default_opts()
    opts.fixed_iterations = 1
    opts.row = wl["rows"]
    w = make_window_problem(wl, opts, rank, ob.optimize, ob.marginalize)
    hook, ransac_impl = cpu_ransac_hook(rank + 1)
    tr = fe.FeatureTrackerOracle(wl["cam"], wl["tcfg"], "cv2", ransac=hook)
    frames, hz = wl["frames"], wl["hz"]
    n = len(frames)
    k0 = [0]

    def front(count):
        t0 = time.perf_counter()
        for _ in range(count):
            k = k0[0]
            tr.read_image(frames[k % n], k / hz, bool(PUB_PATTERN[k % len(PUB_PATTERN)]))
            tr.update_ids()
            k0[0] += 1
        return time.perf_counter() - t0

    def back(count):
        return ob.time_backend(w, with_marg, 0, 4, count) * count

    front(warmup)
    back(max(1, warmup // 2))
    t_front, t_back = front(steps), back(steps)  # serial: one after the other
    res = {}
    th = [threading.Thread(target=lambda: res.__setitem__("f", front(steps))), threading.Thread(target=lambda: res.__setitem__("b", back(steps)))]
    t0 = time.perf_counter()
    for t in th:
        t.start()
    for t in th:
        t.join()
    t_pipe = time.perf_counter() - t0
    return dict(fps=steps / t_pipe, fps_serial=steps / (t_front + t_back), ms_front=1e3 * t_front / steps, ms_back=1e3 * t_back / steps,
                ms_front_pipelined=1e3 * res["f"] / steps, ms_back_pipelined=1e3 * res["b"] / steps,
                cores=os.cpu_count(), cv2_threads=cv2.getNumThreads(), n_proj=w.n_proj, n_landmarks=w.L,
                prior_n=int(w.prior.n) if w.prior is not None else 0, ransac_impl=ransac_impl)


def workload_config(cfg_name, with_marg, n_proj, n_landmarks, prior_n):
    """The `config` object: identical for both arms (it names the workload, nothing arm-specific)."""
    scale, max_cnt, wsz, n_lm, hz = CONFIGS[cfg_name]
    rows, cols = int(round(960 * scale)), int(round(1280 * scale))
    return {"workload": cfg_name, "image": f"{cols}x{rows}", "features": max_cnt, "window_frames": wsz + 1, "n_proj_factors": n_proj,
            "n_landmarks": n_landmarks, "prior_n": prior_n, "iterations": 8, "marginalize": with_marg, "pub_pattern": list(PUB_PATTERN),
            "reject_with_f": "on, every published frame (both arms)", "camera_hz": hz}


PIXEL_SCENARIO = ("closed loop through pixels: a textured 12 x 12 x 5 m room with 16 independently moving objects rendered through the "
                  "Scaramuzza model (PAL-up calibration, annulus 160-500 px, sensor noise 1.5 grey levels) along a hand-held trajectory, "
                  "10 Hz frames, 200 Hz IMU with the PAL noise values; readImage (CLAHE, LK, rejectWithF, setMask, Shi-Tomasi, lift) -> "
                  "tracks with track_cnt > 1 -> processIMU / triangulate / optimization / MARGIN_OLD / slideWindow; bootstrap from "
                  "perturbed ground truth; one run per RANSAC seed; PALVIO itself is not available offline")


def pixel_world(n_frames=36):
    from lfvio import panoworld
    from lfvio.config import PAL_UP, tracker_defaults

    return panoworld.PixelWorld(PAL_UP, dict(tracker_defaults(), max_cnt=150), seed=1, n_frames=n_frames)


def pixel_ate_cpu(world, seeds=(1, 2, 3, 4, 5)):
    from lfvio import violoop
    from oracle import backend as ob
    from oracle import frontend as fe
    from oracle import loop_ops

    o = ob.default_opts()
    o.row = 960
    out = []
    for sd in seeds:
        world.track(fe.FeatureTrackerOracle(world.cam, world.tcfg, "cv2", ransac=cpu_ransac_hook(sd)[0]))
        out.append(violoop.trajectory_error(violoop.VioLoop(world, loop_ops.oracle_ops(), o).run())[0])
    return out


def closed_loop_ate(ops, opts, n_frames=50, seed=1):
    """Second half of BASELINE.json's metric ("ATE vs CPU ref"): a closed-loop sliding-window run (lfvio.violoop) over a
    synthetic 150-feature / 200 Hz IMU stream with the arm's own numeric steps; ATE = RMSE after SE(3) alignment."""
    from lfvio import violoop

    world = violoop.World(seed=seed, n_frames=n_frames)
    loop = violoop.VioLoop(world, ops, opts).run()
    ate, ks = violoop.trajectory_error(loop)
    path = float(sum(np.linalg.norm(world.truth(k + 1)[0] - world.truth(k)[0]) for k in ks[:-1]))
    return loop, {"ate_rmse_m": ate, "frames": len(ks), "path_m": path}


# ---------------------------------------------------------------------------------------------------
def initialisation_leg(lfvio, device):
    """Relative pose (8-point RANSAC + recoverPose), GlobalSFM::construct (PnP / triangulation chain + full BA) and the
    visual-inertial alignment on an 11-frame synthetic start-up window: device entries (host buffers, synchronous, median
    of 5 after one warm-up call) and the numpy port (one run)."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from initial_data import InitScenario
    from oracle import initial as OI

    sc = InitScenario(seed=1, F=11, pixel_noise=0.05)
    c = sc.corres(0, 10)
    l = 0

    def timed(fn, reps):
        ts, out = [], None
        for _ in range(reps):
            t0 = time.perf_counter()
            out = fn()
            ts.append(1e3 * (time.perf_counter() - t0))
        return float(np.median(ts)), out

    dev, cpu = {}, {}
    lfvio.solve_relative_rt(c, 5, 64, device=device)  # warm-up (module load, first allocations)
    dev["relative_rt"], (ok, Rr, Tr, cnt, _) = timed(lambda: lfvio.solve_relative_rt(c, 5, 64, device=device), 5)
    cpu["relative_rt"], (ok_o, Rr_o, Tr_o, cnt_o, _) = timed(lambda: OI.solve_relative_rt(c, 5, 64), 1)
    lfvio.global_sfm(11, l, Rr, Tr, sc.obs_first, sc.obs_frame, sc.obs_pts, device=device)
    dev["global_sfm"], sfm = timed(lambda: lfvio.global_sfm(11, l, Rr, Tr, sc.obs_first, sc.obs_frame, sc.obs_pts, device=device), 5)
    cpu["global_sfm"], sfm_o = timed(lambda: OI.global_sfm(11, l, Rr_o, Tr_o, sc.obs_first, sc.obs_frame, sc.obs_pts), 1)
    chain_ms, _ = timed(lambda: lfvio.global_sfm(11, l, Rr, Tr, sc.obs_first, sc.obs_frame, sc.obs_pts, ba=False, device=device), 5)
    fR = np.array([OI._qrot(sfm_o["q"][i]) @ sc.ric.T for i in range(11)])
    pre = sc.preintegrations()
    args = (fR, sfm_o["T"], [p.delta_p for p in pre], [p.delta_v for p in pre], [p.sum_dt for p in pre], sc.tic, sc.g_norm)
    lfvio.linear_alignment(*args, device=device)
    dev["alignment"], (ok_a, g, x) = timed(lambda: lfvio.linear_alignment(*args, device=device), 5)
    cpu["alignment"], (ok_ao, g_o, x_o) = timed(lambda: OI.linear_alignment(*args), 1)
    dev["total"], cpu["total"] = sum(dev.values()), sum(cpu.values())
    return {"scenario": "11-frame start-up window, %d tracks / %d observations, 0.05 px bearing noise" % (len(sc.ids), len(sc.obs_frame)),
            "device_ms": dev, "cpu_port_ms": cpu, "ba_iterations": int(sfm["summary"].ba_iterations),
            "global_sfm_without_ba_device_ms": chain_ms,
            "agreement": {"relative_R": float(np.abs(Rr - Rr_o).max()), "sfm_T": float(np.abs(sfm["T"] - sfm_o["T"]).max()),
                          "scale": [float(x[-1]), float(x_o[-1])], "gravity": float(np.abs(g - g_o).max())},
            "note": "one-shot work at start-up / restart; the SfM entries reuse a per-process device scratch, the other stage-level entries allocate theirs per call"}


def node_path_closed_loop(lfvio, world, device, device_store=True):
    """FeatureTrackerNode::img_callback -> PointCloud message -> EstimatorNode::process() (the C++ host mirrors of the reference's two
    nodes, lf-vio_b200/host/nodes.cpp) on the rendered stream, host images in, odometry out, cold start (initialStructure runs
    inside).  Wall time per frame from the frame at which the estimator is initialised; ATE of the TUM trajectory it writes."""
    from lfvio import violoop

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
    host.lfvio_host_estimator_init_info.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    w = world
    result = tempfile.mktemp(prefix="lfvio_traj_", suffix=".txt")
    tnode = host.lfvio_host_tracker_node_create()
    ocam = lfvio.Ocam.from_dict(w.cam)
    if host.lfvio_host_tracker_node_configure(tnode, C.byref(ocam), 960, 1280, 150, w.tcfg["min_dist"], 100, 1, 3) != 0:
        return None
    enode = host.lfvio_host_estimator_node_create(10, 1000, device, result.encode())
    if not enode:
        return None
    est = host.lfvio_host_estimator_node_estimator(enode)
    vp = lambda a_: np.ascontiguousarray(a_, np.float64).ctypes.data_as(C.c_void_p)  # noqa: E731
    keep = [np.ascontiguousarray(x, np.float64) for x in (w.noise, w.G, w.tic, w.ric, np.zeros(3), np.zeros(3))]
    host.lfvio_host_estimator_set_params(est, vp(keep[0]), vp(keep[1]), vp(keep[2]), vp(keep[3]), 0.0, 0.0, 1, vp(keep[4]), vp(keep[5]))
    host.lfvio_host_estimator_enable_initialization(est, 9, 0)
    host.lfvio_host_estimator_use_device_store.argtypes = [C.c_void_p, C.c_int]
    host.lfvio_host_estimator_use_device_store(est, 1 if device_store else 0)  # FeatureManager on the device once initialised
    out = np.zeros((512, 9), np.float32)
    imgs = [np.ascontiguousarray(im) for im in w.images]
    host.lfvio_host_estimator_node_imu(enode, float(w.times[0]), vp(w.imu[1][0]), vp(w.imu[1][1]))
    n_sub = len(w.imu[1][2])
    # timed: the frames after the one in which initialStructure() succeeded (the window fills for 10 frames, then every failed
    # attempt costs a global SfM; counting those made round 2's first number, 104 frames/s, a statement about the start-up)
    t_steady, n_steady, steady, k_init = 0.0, 0, False, None
    ii_, gg_ = np.zeros(8, np.int32), np.zeros(3)
    for k in range(w.n_frames):
        t0 = time.perf_counter()
        if k > 0:
            _, _, dt, acc, gyr = w.imu[k]
            for j in range(n_sub):
                stamp = float(w.times[k]) if j == n_sub - 1 else float(w.times[k - 1] + w.frame_dt * (j + 1) / n_sub)
                host.lfvio_host_estimator_node_imu(enode, stamp, vp(acc[j]), vp(gyr[j]))
        r = host.lfvio_host_tracker_node_image(tnode, imgs[k].ctypes.data, 960, 1280, 1280, float(w.times[k]), out.ctypes.data, 512)
        if r >= 0:
            host.lfvio_host_estimator_node_feature(enode, float(w.times[k]), out.ctypes.data, r)
        done = host.lfvio_host_estimator_node_process(enode)
        if steady:
            t_steady += time.perf_counter() - t0
            n_steady += done
        elif k >= 10:
            host.lfvio_host_estimator_init_info(est, ii_.ctypes.data, gg_.ctypes.data)
            if ii_[0] == 1:
                steady, k_init = True, k
    host.lfvio_host_tracker_node_destroy(tnode)
    host.lfvio_host_estimator_node_destroy(enode)
    try:
        tum = np.array([[float(x) for x in ln.split()] for ln in open(result).read().strip().split("\n")])
        os.unlink(result)
        ks = [int(round((t - w.times[0]) / w.frame_dt)) for t in tum[:, 0]]
        ate = violoop.ate_rmse(tum[:, 1:4], np.array([w.truth(k)[0] for k in ks]))
    except Exception:
        return None
    return {"frames_per_s": n_steady / t_steady if t_steady > 0 else None, "ms_per_frame": 1e3 * t_steady / max(n_steady, 1),
            "frames_timed": n_steady, "initialised_at_frame": k_init, "odometry_messages": int(len(tum)), "ate_rmse_m": ate,
            "feature_manager": "device store (lfvio_fm_*)" if device_store else "host list",
            "path": "FeatureTrackerNode::img_callback -> PointCloud -> EstimatorNode::process() (C++ host mirrors over the C-ABI), one host "
                    "thread, host images in, 200 Hz IMU messages in, odometry out; FeatureManager / slideWindow included (on the device "
                    "once initialStructure has succeeded: no projection-factor list is built or uploaded by the host); cold start (the "
                    "frames up to and including the one in which initialStructure succeeds are not timed)"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=300)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="palvio_id01_1280x960_150feat_10kf", choices=list(CONFIGS))
    ap.add_argument("--pool", type=int, default=0, help="distinct frames cycled (0 => enough to exceed L2)")
    ap.add_argument("--cpu-sample", type=int, default=60, help="frames of the bounded CPU-baseline sample")
    ap.add_argument("--no-marginalize", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the ATE / closed-loop / initialisation legs")
    args = ap.parse_args()
    assert args.warmup >= 3 or args.impl == "reference", "W >= 3 warm-up steps required"

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    with_marg = not args.no_marginalize

    if args.impl == "reference":
        if rank != 0:
            return 0
        os.environ["LFVIO_TYPES_ONLY"] = "1"  # interface structs + workload generators only: liblfvio.so is never mapped here
        wl = build_workload(args.config, 0, 32)
        steps = min(args.steps, 100)
        r = cpu_arm(wl, steps, min(args.warmup, 5), 0, with_marg)
        ate = None
        if not args.no_extras and args.config == "palvio_id01_1280x960_150feat_10kf":
            cpu_ates = pixel_ate_cpu(pixel_world())
            ate = {"cpu_ref": {"ate_rmse_m": float(np.mean(cpu_ates)), "seeds": 5, "band_m": [min(cpu_ates), max(cpu_ates)], "all": cpu_ates},
                   "scenario": PIXEL_SCENARIO}
        line = {
            "impl": "reference", "metric": METRIC, "value": r["fps"], "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
            "warmup": min(args.warmup, 5), "ms_per_step": 1e3 / r["fps"], "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u8/f32 front end, f64 back end", "data": "synthetic",
            "config": workload_config(args.config, with_marg, r["n_proj"], r["n_landmarks"], r["prior_n"]),
            "conventions": {"pipelined_frames_per_s": r["fps"], "serial_frames_per_s": r["fps_serial"],
                            "note": "value = pipelined: front end and back end in two host threads, as the reference's two nodes"},
            "cpu_baseline": {"value": r["fps"], "unit": UNIT, "cores": r["cores"], "kind": "port",
                             "sample": f"{steps} frames, two threads; front end = cv2 (OpenCV {r['cv2_threads']} threads, the library the "
                                       f"reference calls) + rejectWithF [{r['ransac_impl']}] + restated control flow; back end = fp64 C++ "
                                       "oracle, 1 thread solve + 4 thread marginalisation (as the reference)",
                             "ms_front": r["ms_front"], "ms_back": r["ms_back"], "ms_front_pipelined": r["ms_front_pipelined"],
                             "ms_back_pipelined": r["ms_back_pipelined"]},
            "e2e": {"value": r["fps"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "ate": ate,
            "native_product_library_mapped": any("liblfvio" in ln for ln in open("/proc/self/maps")),
        }
        print(json.dumps(line))
        return 0

    import threading

    # One process per GPU, NUMA-local: bind this rank to the CPUs next to its GPU BEFORE the CUDA context and the pinned buffers
    # exist (torchrun does not bind).  Recorded in the line as per_rank[..].cpu_affinity.
    numa_note = "unbound"
    try:
        import pynvml

        pynvml.nvmlInit()
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        phys = int(vis.split(",")[local_rank]) if vis else local_rank
        hnd = pynvml.nvmlDeviceGetHandleByIndex(phys)
        words = pynvml.nvmlDeviceGetCpuAffinity(hnd, (os.cpu_count() + 63) // 64)
        cpus = {64 * i + b for i, wd in enumerate(words) for b in range(64) if (wd >> b) & 1}
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            numa_note = f"{min(cpus)}-{max(cpus)} ({len(cpus)} cpus)"
    except Exception as e:  # no NVML / not permitted: run unbound
        numa_note = f"unbound ({type(e).__name__})"

    import torch

    import lfvio
    from lfvio import backend as B
    from lfvio import multi

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        import torch.distributed as dist

        if os.environ.get("NCCL_DEBUG", "").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"  # NCCL prints its version banner on stdout; stdout carries the one JSON line
        dist.init_process_group("nccl", device_id=dev)
    peaks, peak_kind = measured_peaks()

    scale, max_cnt, wsz, n_lm, hz = CONFIGS[args.config]
    rows, cols = int(round(960 * scale)), int(round(1280 * scale))
    l2_bytes = 126 * 2 ** 20
    pool = args.pool or max(16, int(np.ceil(1.25 * l2_bytes / (rows * cols))))
    wl = build_workload(args.config, rank, pool)
    frames_host = wl["frames"]
    frames_dev = torch.from_numpy(frames_host).to(dev)
    pinned = torch.from_numpy(frames_host).pin_memory() if frames_host.nbytes < 2 ** 31 else None
    torch.cuda.synchronize()

    tracker = lfvio.FeatureTracker(wl["cam"], wl["tcfg"], device=local_rank)
    opts = lfvio.solver_default_opts()
    opts.fixed_iterations = 1
    opts.row = rows
    window = lfvio.Window(wsz, max(1000, 2 * n_lm), max(16384, 24 * n_lm), device=local_rank)
    wprob = make_window_problem(wl, opts, rank, window.optimize, window.marginalize)
    window.upload(wprob)
    t_stream = torch.cuda.ExternalStream(tracker.stream, device=dev)
    w_stream = torch.cuda.ExternalStream(window.stream, device=dev)
    frame_bytes = rows * cols
    px = rows * cols

    def step_dev(k):
        tracker.read_image_dev(frames_dev[k % pool].data_ptr(), cols, k / hz, bool(PUB_PATTERN[k % len(PUB_PATTERN)]))
        window.optimize_dev(with_marg, B.MARGIN_OLD)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    # ---- value: inputs resident in HBM, no host synchronisation inside the timed region -------------
    clocks = ClockSampler(local_rank) if rank == 0 else None  # one poller per job: nvidia-smi takes a driver-wide lock
    if clocks:
        clocks.start()
        clocks.wait_ready()
    # The path is a chain of latency-sized kernels on one or a few SMs: its speed follows the SM clock, and a GPU that has been
    # idle sits at a low clock for the first tenths of a second of such a light load (measured at N = 4: two ranks ran their
    # solver stream 1.5x slower than the other two — which turned out to be the rank-seeded window, not the clock: all four read
    # 1965 MHz).  Every rank still runs the path for half a second before the W warm-up steps, and the SM clock of every rank is
    # read (NVML, one call) right after the timed region and reported in per_rank, so that a cold rank would show.
    def sm_clock_mhz():
        try:
            import pynvml

            pynvml.nvmlInit()
            h = pynvml.nvmlDeviceGetHandleByIndex(torch.cuda.current_device() if not os.environ.get("CUDA_VISIBLE_DEVICES") else
                                                  int(os.environ["CUDA_VISIBLE_DEVICES"].split(",")[torch.cuda.current_device()]))
            return float(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM))
        except Exception:
            return float("nan")

    t_spin = time.perf_counter()
    k_spin = 0
    while time.perf_counter() - t_spin < 0.5:  # the path itself, half a second, before the W warm-up steps (own kernels only)
        for _ in range(32):
            step_dev(k_spin)
            k_spin += 1
        tracker.sync()
    for k in range(args.warmup):
        step_dev(k)
    tracker.sync()
    barrier()
    if clocks:
        clocks.mark_begin()
    launches0 = lfvio.lib.lfvio_launch_count()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    ev[0].record(t_stream)
    ev[1].record(w_stream)
    # Two host threads, as in the reference (feature_tracker_node / estimator_node are two processes): each enqueues its
    # own CUDA stream and closes its own event bracket; nothing inside the loops synchronises with the device.
    enq = [0.0, 0.0]

    def enqueue_tracker():
        t0 = time.perf_counter()
        for k in range(args.warmup, args.warmup + args.steps):
            tracker.read_image_dev(frames_dev[k % pool].data_ptr(), cols, k / hz, bool(PUB_PATTERN[k % len(PUB_PATTERN)]))
        ev[2].record(t_stream)
        enq[0] = time.perf_counter() - t0

    def enqueue_window():
        t0 = time.perf_counter()
        for k in range(args.warmup, args.warmup + args.steps):
            window.optimize_dev(with_marg, B.MARGIN_OLD)
        ev[3].record(w_stream)
        enq[1] = time.perf_counter() - t0

    t_wall0 = time.perf_counter()
    th = [threading.Thread(target=enqueue_tracker), threading.Thread(target=enqueue_window)]
    for t in th:
        t.start()
    for t in th:
        t.join()
    torch.cuda.synchronize()
    t_wall = time.perf_counter() - t_wall0
    my_sm_mhz = sm_clock_mhz()
    if clocks:
        clocks.mark_end()
    launches = lfvio.lib.lfvio_launch_count() - launches0
    clk = clocks.stop() if clocks else None
    # both streams start together after the barrier; the step is done when both are
    ms_tracker, ms_window = ev[0].elapsed_time(ev[2]), ev[1].elapsed_time(ev[3])
    ms_total = max(ms_tracker, ms_window, ev[0].elapsed_time(ev[3]), ev[1].elapsed_time(ev[2]), 0.0)
    tracker.sync()
    n_tracks = len(tracker.ids)
    summ = window.download(wprob.clone())
    my_ms_total = ms_total
    ms_total, frames_total = multi.reduce_counters(ms_total, args.steps, dev)  # MAX of times, SUM of frames (NCCL)
    # the same two streams one after the other (serial convention): tracker alone, then window alone
    barrier()
    ev2 = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    n_ser = min(args.steps, 200)
    ev2[0].record(t_stream)
    for k in range(n_ser):
        tracker.read_image_dev(frames_dev[k % pool].data_ptr(), cols, (args.warmup + args.steps + k) / hz, bool(PUB_PATTERN[k % len(PUB_PATTERN)]))
    ev2[1].record(t_stream)
    torch.cuda.synchronize()
    ev2[2].record(w_stream)
    for k in range(n_ser):
        window.optimize_dev(with_marg, B.MARGIN_OLD)
    ev2[3].record(w_stream)
    torch.cuda.synchronize()
    ms_tracker_alone, ms_window_alone = ev2[0].elapsed_time(ev2[1]) / n_ser, ev2[2].elapsed_time(ev2[3]) / n_ser
    tracker.sync()
    barrier()

    # ---- e2e: host buffers through the C-ABI, H2D + D2H inside the timed region ------------------------
    e2e_steps = max(10, min(args.steps, 200))
    work = wprob.clone()
    init = (wprob.pose.copy(), wprob.sb.copy(), wprob.ex.copy(), wprob.feat.copy(), wprob.td.copy())
    src_frames = pinned.numpy() if pinned is not None else frames_host
    prior_out = B.PriorData()

    # The reference runs the two halves as two ROS nodes (two processes, feature_tracker_node.cpp / estimator_node.cpp)
    # connected by a topic: a frame's tracking overlaps the previous frame's optimisation.  The e2e pass mirrors that
    # with two host threads, each making the blocking C-ABI calls of its node (ctypes releases the GIL in the calls).
    node_s = {"tracker": 0.0, "estimator": 0.0}
    k_e2e = args.warmup + args.steps + n_ser

    def tracker_node(k0, n):
        t_node = time.perf_counter()
        for k in range(k0, k0 + n):
            tracker.read_image(src_frames[k % pool], (k_e2e + k) / hz, bool(PUB_PATTERN[k % len(PUB_PATTERN)]))
            tracker.update_ids()
        node_s["tracker"] = time.perf_counter() - t_node

    # (the argument structs of the call are built once: they only hold pointers to `work`'s arrays, which are refilled per step)
    full_call = window.prepared_optimize_full(work, B.MARGIN_OLD, prior_out) if with_marg else None
    state_pairs = list(zip((work.pose, work.sb, work.ex, work.feat, work.td), init))

    def estimator_node(k0, n):
        t_node = time.perf_counter()
        for k in range(k0, k0 + n):
            for dst, s_ in state_pairs:
                np.copyto(dst, s_)
            if with_marg:  # Estimator::optimization() in one call: solve, double2vector gauge fix, next prior (all on device)
                full_call()
            else:
                window.optimize(work)
        node_s["estimator"] = time.perf_counter() - t_node

    def run_nodes(k0, n):
        th_ = [threading.Thread(target=tracker_node, args=(k0, n)), threading.Thread(target=estimator_node, args=(k0, n))]
        t_start = time.perf_counter()
        for t in th_:
            t.start()
        for t in th_:
            t.join()
        torch.cuda.synchronize()
        return time.perf_counter() - t_start

    run_nodes(0, 3)
    barrier()
    e2e_s = run_nodes(3, e2e_steps)
    e2e_node_ms = {k: 1e3 * v / e2e_steps for k, v in node_s.items()}
    # the same calls made back to back by ONE thread (latency of a frame through both halves = the serial convention)
    t0 = time.perf_counter()
    n_seq = max(5, e2e_steps // 4)
    for k in range(n_seq):
        tracker_node(3 + e2e_steps + k, 1)
        estimator_node(0, 1)
    torch.cuda.synchronize()
    e2e_seq_ms = 1e3 * (time.perf_counter() - t0) / n_seq
    my_e2e_ms = 1e3 * e2e_s
    e2e_ms, e2e_frames = multi.reduce_counters(1e3 * e2e_s, e2e_steps, dev)
    e2e_s = e2e_ms * 1e-3
    blob_bytes = wprob.n_proj * C.sizeof(B.ProjFactor) + wprob.n_imu * C.sizeof(B.ImuFactor) + 8 * (16 * (wsz + 1) + 8 + wprob.L)
    if wprob.prior is not None:  # MarginalizationInfo travels with the problem: J0 (n x n), r0, x0, block tables
        blob_bytes += 8 * (int(wprob.prior.n) ** 2 + 3 * int(wprob.prior.n) + 64)
    h2d = frame_bytes + blob_bytes
    d2h = max_cnt * (8 + 4 + 4 + 12 + 12) + 8 + 8 * (16 * (wsz + 1) + 8 + wprob.L) + 2048

    # per-rank record (every N): which rank is the slow one, and whether its host or its device was
    mine = torch.tensor([my_ms_total / args.steps, ms_tracker / args.steps, ms_window / args.steps, 1e3 * enq[0] / args.steps,
                         1e3 * enq[1] / args.steps, my_e2e_ms / e2e_steps, e2e_node_ms["tracker"], e2e_node_ms["estimator"], my_sm_mhz],
                        dtype=torch.float64, device=dev)
    if world > 1:
        allr = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(allr, mine)
    else:
        allr = [mine]
    keys = ("ms_per_step", "ms_tracker_stream", "ms_window_stream", "host_enqueue_ms_tracker", "host_enqueue_ms_window",
            "e2e_ms_per_step", "e2e_tracker_node_ms", "e2e_estimator_node_ms", "sm_mhz_after_timed_region")
    per_rank = [{k: round(float(v), 5) for k, v in zip(keys, r_.tolist())} for r_ in allr]
    if world > 1:
        notes = [None] * world
        dist.all_gather_object(notes, numa_note)
    else:
        notes = [numa_note]
    for pr_, nt_ in zip(per_rank, notes):
        pr_["cpu_affinity"] = nt_

    # ---- per-kernel durations, live (CUDA events on the launching streams), same workload ---------------
    prof_steps = min(args.steps, 200)
    nt, nw = lfvio.lib.lfvio_tracker_stage_count(), lfvio.lib.lfvio_window_stage_count()
    lfvio.check(lfvio.lib.lfvio_tracker_profile_begin(tracker.handle, prof_steps), "profile_begin")
    lfvio.check(lfvio.lib.lfvio_window_profile_begin(window._h, prof_steps), "profile_begin")
    window.upload(wprob)
    for k in range(prof_steps):
        step_dev(k_e2e + e2e_steps + n_seq + 8 + k)
    t_ms, t_cnt = (C.c_float * nt)(), (C.c_int * nt)()
    w_ms, w_cnt = (C.c_float * nw)(), (C.c_int * nw)()
    lfvio.check(lfvio.lib.lfvio_tracker_profile_end(tracker.handle, t_ms, t_cnt), "profile_end")
    lfvio.check(lfvio.lib.lfvio_window_profile_end(window._h, w_ms, w_cnt), "profile_end")
    tracker.sync()
    stages = {}
    for i in range(nt):
        if t_cnt[i]:
            stages[lfvio.lib.lfvio_tracker_stage_name(i).decode()] = (t_ms[i] / prof_steps, t_ms[i] / t_cnt[i], t_cnt[i] / prof_steps)
    for i in range(nw):
        if w_cnt[i]:
            stages[lfvio.lib.lfvio_window_stage_name(i).decode()] = (w_ms[i] / prof_steps, w_ms[i] / w_cnt[i], w_cnt[i] / prof_steps)
    # per-rank kernel times of the solver chain (us per launch): tells a slow rank's slow KERNELS from gaps between them
    chain = ("step_kernel", "evaluate_kernel(candidate)", "accumulate_kernel", "marginalize", "lk_kernel")
    mine_k = torch.tensor([1e3 * stages[k_][1] if k_ in stages else float("nan") for k_ in chain], dtype=torch.float64, device=dev)
    if world > 1:
        allk = [torch.zeros_like(mine_k) for _ in range(world)]
        dist.all_gather(allk, mine_k)
    else:
        allk = [mine_k]
    for pr_, kk_ in zip(per_rank, allk):
        pr_["us_per_launch"] = {k_: round(float(v_), 2) for k_, v_ in zip(chain, kk_.tolist())}
    # algorithmic bytes per launch group (SURVEY.md §8d / DESIGN.md)
    n_lk = max(n_tracks, 1)
    D = 15 * (wsz + 1) + 7
    pn = int(wprob.prior.n) if wprob.prior is not None else 0
    algo_bytes = {
        "clahe(3 kernels)": 3.0 * px,
        "pyr_down(3 kernels)": px * (1 + 0.25 + 0.0625 + 0.015625) + px * (0.25 + 0.0625),
        "lk_kernel": n_lk * 4 * 2 * 43 * 43 + n_lk * 17,
        "min_eig_kernel": 5.0 * px + px,
        "candidates_kernel": 5.0 * px,
        "mask_raster(copy+kernel)": 2.0 * px,
        "evaluate_kernel(candidate)": wprob.n_proj * (144 + 336) + wprob.n_imu * (3720 + 3720) + 8.0 * pn * pn,
        "accumulate_kernel": wprob.n_proj * 336 + wprob.n_imu * 3720 + 8.0 * D * D + 8.0 * pn * pn,
        "step_kernel": wprob.L * 13 * 8 * 2 + wprob.n_proj * 6 * 8 * 2 + 8.0 * D * (D + 1) / 2 + 6 * 8.0 * (D + wprob.L),
    }
    dom = max(stages.items(), key=lambda kv: kv[1][0])
    dom_name, (dom_ms_step, dom_ms_launch, dom_per_step) = dom[0], dom[1]
    ab = algo_bytes.get(dom_name)
    achieved = (ab / (dom_ms_launch * 1e-3) / 1e9) if ab else None
    # dram__bytes_read + dram__bytes_write per launch of the dominant kernel from the committed `ncu --set full` capture
    # of this command at the default workload; null for other workloads / kernels
    traffic = None
    try:
        if args.config == "palvio_id01_1280x960_150feat_10kf":
            for name in ("r2_traffic.json", "r1_traffic.json"):
                fp_ = os.path.join(ROOT, "profiles", name)
                if os.path.exists(fp_):
                    traffic = json.load(open(fp_))["dram_bytes_per_launch"].get(dom_name)
                    break
    except Exception:
        traffic = None
    # fp64 work of the step kernel: skyline Cholesky of the D x D reduced system (D^3 / 3 multiply-adds over the measured
    # 3.6x (F = 11) / 4.6x (F = 21) skyline saving), two triangular solves, landmark back-substitution
    sky = 3.6 if wsz <= 10 else 4.6
    step_flop = 2.0 * (D ** 3 / 3.0 / sky + 2.0 * D * D / 2 + 13.0 * wprob.L + 6.0 * wprob.n_proj)
    roofline = {
        "bound": "hbm", "kernel": dom_name, "achieved": achieved, "peak": peaks["hbm_gbs"], "unit": "GB/s",
        "frac": (achieved / peaks["hbm_gbs"]) if achieved else None, "traffic": traffic, "peak_kind": peak_kind,
        "algorithmic_bytes_per_launch": ab, "avg_launch_ms": dom_ms_launch,
        "latency_bound": True,
        "fp64": {"kernel": "step_kernel", "flop_per_launch": step_flop,
                 "achieved_gflops": (step_flop / (stages["step_kernel"][1] * 1e-3) / 1e9) if "step_kernel" in stages else None,
                 "note": "one CTA, a chain of 15 F + 7 dependent pivots (per pivot one DFMA 9.4 + one shuffle 26 + one reciprocal 40 cycles on "
                         "sm_100, 113 measured: tools/probe/prim_probe.cu) plus the block hand-overs: the yardstick is the length of that "
                         "chain, not HBM or the fp64 pipe; the figure is an average over launches that factor (~85 us) and launches that "
                         "re-use the factor after a rejected step (~5 us)"},
        "stage_ms_per_step": {k: round(v[0], 5) for k, v in sorted(stages.items(), key=lambda kv: -kv[1][0])},
        "note": "configs 1-3 are launch/latency bound (a whole 1280x960 frame is ~3 us of HBM time, SURVEY.md §7); stage timers put "
                "event records between the kernels, so graph replay is off in that pass and the stages add up to more than ms_per_step",
        # every stage that has an algorithmic byte count: [avg launch-group us, achieved GB/s, fraction of the measured HBM peak]
        "per_stage": {k: [round(1e3 * v[1], 2), round(algo_bytes[k] / (v[1] * 1e-3) / 1e9, 1),
                          round(algo_bytes[k] / (v[1] * 1e-3) / 1e9 / peaks["hbm_gbs"], 4)]
                      for k, v in stages.items() if k in algo_bytes and v[1] > 0},
    }

    # ---- CPU baseline on this box's host cores (rank 0, N = 1 only) -----------------------------------
    cpu = None
    cpu_conv = None
    if rank == 0 and world == 1:
        r = cpu_arm(wl, args.cpu_sample, 3, 0, with_marg)
        cpu = {"value": r["fps"], "unit": UNIT, "cores": r["cores"], "kind": "port",
               "sample": f"{args.cpu_sample} frames of the same workload, two threads; cv2 front end ({r['cv2_threads']} threads) + rejectWithF "
                         f"[{r['ransac_impl']}] + fp64 C++ oracle back end (1 thread solve, 4 thread marginalisation)",
               "ms_front": r["ms_front"], "ms_back": r["ms_back"], "serial_frames_per_s": r["fps_serial"]}
        cpu_conv = {"pipelined_frames_per_s": r["fps"], "serial_frames_per_s": r["fps_serial"]}

    # ---- ATE through pixels + the node path on the same stream (rank 0, N = 1 only) ---------------------
    ate = closed = init_leg = None
    if rank == 0 and world == 1 and not args.no_extras and args.config == "palvio_id01_1280x960_150feat_10kf":
        from lfvio import violoop

        pw = pixel_world()
        cpu_ates = pixel_ate_cpu(pw)
        d_opts = lfvio.solver_default_opts()
        d_opts.row = 960
        win2 = lfvio.Window(10, 1000, 16384, device=local_rank)
        dev_ates = []
        for sd in (1, 2, 3, 4, 5):
            tr2 = lfvio.FeatureTracker(pw.cam, dict(pw.tcfg, ransac=1, ransac_seed=sd), device=local_rank)
            pw.track(tr2)
            tr2.close()
            dev_ates.append(violoop.trajectory_error(violoop.VioLoop(pw, violoop.device_ops(lfvio, win2, local_rank), d_opts).run())[0])
        win2.close()
        path_m = float(sum(np.linalg.norm(pw.truth(k + 1)[0] - pw.truth(k)[0]) for k in range(10, pw.n_frames - 1)))
        ate = {"device": {"ate_rmse_m": float(np.mean(dev_ates)), "seeds": 5, "band_m": [min(dev_ates), max(dev_ates)], "all": dev_ates},
               "cpu_ref": {"ate_rmse_m": float(np.mean(cpu_ates)), "seeds": 5, "band_m": [min(cpu_ates), max(cpu_ates)], "all": cpu_ates},
               "device_inside_cpu_band": bool(min(cpu_ates) - 0.05 * np.mean(cpu_ates) <= min(dev_ates) and max(dev_ates) <= max(cpu_ates) + 0.05 * np.mean(cpu_ates)),
               "max_abs_difference_same_seed_m": float(max(abs(a - b) for a, b in zip(dev_ates, cpu_ates))),
               "frames": pw.n_frames - 10, "path_m": path_m, "tracked_min": int(min(pw.n_tracked[2:])), "scenario": PIXEL_SCENARIO}
        closed = node_path_closed_loop(lfvio, pw, local_rank)
        init_leg = initialisation_leg(lfvio, local_rank)

    if rank == 0:
        fps = frames_total / (ms_total * 1e-3)
        line = {
            "metric": METRIC, "value": fps, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u8/f32 front end, f64 back end", "data": "synthetic",
            "config": workload_config(args.config, with_marg, wprob.n_proj, wprob.L, int(wprob.prior.n) if wprob.prior is not None else 0),
            "run": {"tracked": n_tracks, "iterations": int(summ.iterations),
                    "l2_policy": f"inputs larger than L2: {pool}-frame pool ({pool * px / 2**20:.0f} MiB) cycled",
                    "streams": "one independent stream per GPU; front end and back end on separate CUDA streams, each frame / solve one CUDA graph launch"},
            "conventions": {"pipelined_frames_per_s": fps,
                            "serial_frames_per_s": world * 1e3 / (ms_tracker_alone + ms_window_alone),
                            "ms_tracker_alone": ms_tracker_alone, "ms_window_alone": ms_window_alone,
                            "e2e_pipelined_frames_per_s": e2e_frames / e2e_s, "e2e_serial_frames_per_s": world * 1e3 / e2e_seq_ms,
                            "cpu_arm": cpu_conv,
                            "note": "value / e2e = pipelined (two host threads = the reference's two nodes); the reference arm's value is "
                                    "pipelined too; serial = one thread, front end then back end"},
            "ms_tracker_stream": ms_tracker / args.steps, "ms_window_stream": ms_window / args.steps,
            "wall_ms_per_step": 1e3 * t_wall / args.steps,
            "host_enqueue_ms_per_step": {"tracker_thread": 1e3 * enq[0] / args.steps, "window_thread": 1e3 * enq[1] / args.steps},
            "per_rank": per_rank,
            "e2e": {"value": e2e_frames / e2e_s, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                    "steps": e2e_steps, "host_threads": 2, "single_thread_ms_per_frame": e2e_seq_ms, "node_ms_per_frame": e2e_node_ms,
                    "note": "two host threads = the reference's two ROS nodes; each makes blocking C-ABI calls with host buffers"},
            "gpu_launches": int(launches),
            "roofline": roofline,
            "cpu_baseline": cpu,
            "ate": ate,
            "closed_loop": closed,
            "initialisation": init_leg,
            "clocks": clk,
        }
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
