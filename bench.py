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
               host threads (the reference's two nodes); timed with CUDA events on those streams, max over ranks
  e2e          the same metric through the blocking C-ABI calls with HOST buffers (pinned frame H2D, problem upload,
               results D2H inside the timed region), two host threads; node_ms_per_frame / single_thread_ms_per_frame
  roofline     dominant stage by live CUDA-event time; achieved = algorithmic bytes / average launch time; traffic from
               the committed ncu capture (profiles/r1_traffic.json); per_stage lists every stage with a byte count
  cpu_baseline bounded sample of the CPU arm on this box's cores (N = 1 only)
  ate          closed-loop sliding-window run of both arms on one synthetic world (the second half of the metric)
  clocks, gpu_launches, host_enqueue_ms_per_step
`--impl reference` times the reference's CPU path (cv2 with the reference's arguments + the fp64 C++ oracle of the
back end; the reference itself — ROS/OpenCV-C++/Ceres/Eigen — cannot be built in this image).
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
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "20",
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
    return wl["synthwin"].make_window_with_prior(optimize, marginalize, seed=rank + 1, n_frames=wl["window_size"] + 1,
                                                 n_landmarks=wl["n_landmarks"], feats_per_frame=wl["max_cnt"], opts=opts,
                                                 perturb=0.3)


# ---------------------------------------------------------------------------------------------------
# CPU arm (reference path): cv2 front end + fp64 oracle back end
# ---------------------------------------------------------------------------------------------------
def cpu_arm(wl, steps: int, warmup: int, rank: int, with_marg: bool):
    import cv2

    from oracle import backend as ob
    from oracle import frontend as fe

    opts = ob.default_opts()
    opts.fixed_iterations = 1
    opts.row = wl["rows"]
    w = make_window_problem(wl, opts, rank, ob.optimize, ob.marginalize)
    tr = fe.FeatureTrackerOracle(wl["cam"], wl["tcfg"], "cv2")
    frames, hz = wl["frames"], wl["hz"]
    n = len(frames)
    t_front = t_back = 0.0
    for k in range(warmup + steps):
        img = frames[k % n]
        t0 = time.perf_counter()
        tr.read_image(img, k / hz, bool(PUB_PATTERN[k % len(PUB_PATTERN)]))
        tr.update_ids()
        t1 = time.perf_counter()
        tb = ob.time_backend(w, with_marg, 0, 4, 1)
        if k >= warmup:
            t_front += t1 - t0
            t_back += tb
    total = t_front + t_back
    return dict(fps=steps / total, ms_front=1e3 * t_front / steps, ms_back=1e3 * t_back / steps,
                cores=os.cpu_count(), cv2_threads=cv2.getNumThreads(), n_proj=w.n_proj, n_landmarks=w.L,
                prior_n=int(w.prior.n) if w.prior is not None else 0)


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
    args = ap.parse_args()
    assert args.warmup >= 3 or args.impl == "reference", "W >= 3 warm-up steps required"

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    with_marg = not args.no_marginalize

    if args.impl == "reference":
        if rank != 0:
            return 0
        scale = CONFIGS[args.config][0]
        wl = build_workload(args.config, 0, 32)
        steps = min(args.steps, 200)
        r = cpu_arm(wl, steps, min(args.warmup, 5), 0, with_marg)
        from oracle import backend as ob_
        from oracle import loop_ops

        o_ = ob_.default_opts()
        o_.row = 960
        _, ate_cpu = closed_loop_ate(loop_ops.oracle_ops(), o_)
        line = {
            "impl": "reference", "ate": {"cpu_ref": ate_cpu, "scenario": ATE_SCENARIO}, "metric": METRIC, "value": r["fps"], "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
            "warmup": min(args.warmup, 5), "ms_per_step": 1e3 / r["fps"], "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u8/f32 front end, f64 back end", "data": "synthetic",
            "config": {"workload": args.config, "scale": scale, "n_proj_factors": r["n_proj"], "n_landmarks": r["n_landmarks"], "prior_n": r["prior_n"],
                       "iterations": 8, "marginalize": with_marg},
            "cpu_baseline": {"value": r["fps"], "unit": UNIT, "cores": r["cores"], "kind": "port",
                             "sample": f"{steps} frames; front end = cv2 (OpenCV {r['cv2_threads']} threads, the library the "
                                       "reference calls) + restated control flow; back end = fp64 C++ oracle, 1 thread solve "
                                       "+ 4 thread marginalisation (as the reference)",
                             "ms_front": r["ms_front"], "ms_back": r["ms_back"]},
            "e2e": {"value": r["fps"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        }
        print(json.dumps(line))
        return 0

    import torch

    import lfvio
    from lfvio import backend as B

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

    def step_dev(k):
        tracker.read_image_dev(frames_dev[k % pool].data_ptr(), cols, k / hz, bool(PUB_PATTERN[k % len(PUB_PATTERN)]))
        window.optimize_dev(with_marg, B.MARGIN_OLD)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    # ---- value: inputs resident in HBM, no host synchronisation inside the timed region -------------
    clocks = ClockSampler(local_rank)
    clocks.start()
    clocks.wait_ready()
    for k in range(args.warmup):
        step_dev(k)
    tracker.sync()
    barrier()
    clocks.mark_begin()
    launches0 = lfvio.lib.lfvio_launch_count()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    ev[0].record(t_stream)
    ev[1].record(w_stream)
    # Two host threads, as in the reference (feature_tracker_node / estimator_node are two processes): each enqueues its
    # own CUDA stream; nothing inside the loops synchronises with the device.
    import threading

    enq = [0.0, 0.0]

    def enqueue_tracker():
        t0 = time.perf_counter()
        for k in range(args.warmup, args.warmup + args.steps):
            tracker.read_image_dev(frames_dev[k % pool].data_ptr(), cols, k / hz, bool(PUB_PATTERN[k % len(PUB_PATTERN)]))
        enq[0] = time.perf_counter() - t0

    def enqueue_window():
        t0 = time.perf_counter()
        for k in range(args.warmup, args.warmup + args.steps):
            window.optimize_dev(with_marg, B.MARGIN_OLD)
        enq[1] = time.perf_counter() - t0

    t_wall0 = time.perf_counter()
    th = [threading.Thread(target=enqueue_tracker), threading.Thread(target=enqueue_window)]
    for t in th:
        t.start()
    for t in th:
        t.join()
    ev[2].record(t_stream)
    ev[3].record(w_stream)
    torch.cuda.synchronize()
    t_wall = time.perf_counter() - t_wall0
    clocks.mark_end()
    launches = lfvio.lib.lfvio_launch_count() - launches0
    clk = clocks.stop()
    # both streams start together after the barrier; the step is done when both are
    ms_total = max(ev[0].elapsed_time(ev[2]), ev[1].elapsed_time(ev[3]), ev[0].elapsed_time(ev[3]), ev[1].elapsed_time(ev[2]))
    ms_total = max(ms_total, 0.0)
    ms_tracker, ms_window = ev[0].elapsed_time(ev[2]), ev[1].elapsed_time(ev[3])
    tracker.sync()
    n_tracks = len(tracker.ids)
    summ = window.download(wprob.clone())
    from lfvio import multi

    ms_total, frames_total = multi.reduce_counters(ms_total, args.steps, dev)  # MAX of times, SUM of frames (NCCL)
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
    import threading

    node_s = {"tracker": 0.0, "estimator": 0.0}

    def tracker_node(k0, n):
        t_node = time.perf_counter()
        for k in range(k0, k0 + n):
            tracker.read_image(src_frames[k % pool], (args.warmup + args.steps + k) / hz, bool(PUB_PATTERN[k % len(PUB_PATTERN)]))
            tracker.update_ids()
        node_s["tracker"] = time.perf_counter() - t_node

    def estimator_node(k0, n):
        t_node = time.perf_counter()
        for k in range(k0, k0 + n):
            for dst, s in zip((work.pose, work.sb, work.ex, work.feat, work.td), init):
                np.copyto(dst, s)
            if with_marg:  # Estimator::optimization() in one call: solve, double2vector gauge fix, next prior (all on device)
                window.optimize_full(work, B.MARGIN_OLD, prior_out)
            else:
                window.optimize(work)
        node_s["estimator"] = time.perf_counter() - t_node

    def run_nodes(k0, n):
        th = [threading.Thread(target=tracker_node, args=(k0, n)), threading.Thread(target=estimator_node, args=(k0, n))]
        t_start = time.perf_counter()
        for t in th:
            t.start()
        for t in th:
            t.join()
        torch.cuda.synchronize()
        return time.perf_counter() - t_start

    run_nodes(0, 3)
    barrier()
    e2e_s = run_nodes(3, e2e_steps)
    e2e_node_ms = {k: 1e3 * v / e2e_steps for k, v in node_s.items()}
    # the same calls made back to back by ONE thread (latency of a frame through both halves)
    t0 = time.perf_counter()
    n_seq = max(5, e2e_steps // 4)
    for k in range(n_seq):
        tracker_node(3 + e2e_steps + k, 1)
        estimator_node(0, 1)
    torch.cuda.synchronize()
    e2e_seq_ms = 1e3 * (time.perf_counter() - t0) / n_seq
    e2e_ms, e2e_frames = multi.reduce_counters(1e3 * e2e_s, e2e_steps, dev)
    e2e_s = e2e_ms * 1e-3
    blob_bytes = wprob.n_proj * C.sizeof(B.ProjFactor) + wprob.n_imu * C.sizeof(B.ImuFactor) + 8 * (16 * (wsz + 1) + 8 + wprob.L)
    if wprob.prior is not None:  # MarginalizationInfo travels with the problem: J0 (n x n), r0, x0, block tables
        blob_bytes += 8 * (int(wprob.prior.n) ** 2 + 3 * int(wprob.prior.n) + 64)
    h2d = frame_bytes + blob_bytes
    d2h = max_cnt * (8 + 4 + 4 + 12 + 12) + 8 + 8 * (16 * (wsz + 1) + 8 + wprob.L) + 2048

    # ---- per-kernel durations, live (CUDA events on the launching streams), same workload ---------------
    prof_steps = min(args.steps, 200)
    nt, nw = lfvio.lib.lfvio_tracker_stage_count(), lfvio.lib.lfvio_window_stage_count()
    lfvio.check(lfvio.lib.lfvio_tracker_profile_begin(tracker.handle, prof_steps), "profile_begin")
    lfvio.check(lfvio.lib.lfvio_window_profile_begin(window._h, prof_steps), "profile_begin")
    window.upload(wprob)
    for k in range(prof_steps):
        step_dev(args.warmup + args.steps + k)
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
    # algorithmic bytes per launch group (SURVEY.md §8d / DESIGN.md)
    px = rows * cols
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
    # of this command at the default workload (profiles/r1_traffic.json); null for other workloads / kernels
    traffic = None
    try:
        if args.config == "palvio_id01_1280x960_150feat_10kf":
            traffic = json.load(open(os.path.join(ROOT, "profiles", "r1_traffic.json")))["dram_bytes_per_launch"].get(dom_name)
    except Exception:
        traffic = None
    roofline = {
        "bound": "hbm", "kernel": dom_name, "achieved": achieved, "peak": peaks["hbm_gbs"], "unit": "GB/s",
        "frac": (achieved / peaks["hbm_gbs"]) if achieved else None, "traffic": traffic, "peak_kind": peak_kind,
        "algorithmic_bytes_per_launch": ab, "avg_launch_ms": dom_ms_launch,
        "stage_ms_per_step": {k: round(v[0], 5) for k, v in sorted(stages.items(), key=lambda kv: -kv[1][0])},
        "note": "configs 1-3 are launch/latency bound (a whole 1280x960 frame is ~3 us of HBM time, SURVEY.md §7)",
        # every stage that has an algorithmic byte count: [avg launch-group us, achieved GB/s, fraction of the measured HBM peak]
        "per_stage": {k: [round(1e3 * v[1], 2), round(algo_bytes[k] / (v[1] * 1e-3) / 1e9, 1),
                          round(algo_bytes[k] / (v[1] * 1e-3) / 1e9 / peaks["hbm_gbs"], 4)]
                      for k, v in stages.items() if k in algo_bytes and v[1] > 0},
    }

    # ---- CPU baseline on this box's host cores (rank 0, N = 1 only) -----------------------------------
    cpu = None
    if rank == 0 and world == 1:
        r = cpu_arm(wl, args.cpu_sample, 3, 0, with_marg)
        cpu = {"value": r["fps"], "unit": UNIT, "cores": r["cores"], "kind": "port",
               "sample": f"{args.cpu_sample} frames of the same workload; cv2 front end ({r['cv2_threads']} threads) + fp64 C++ "
                         "oracle back end (1 thread solve, 4 thread marginalisation)",
               "ms_front": r["ms_front"], "ms_back": r["ms_back"]}

    # ---- ATE of the closed loop, device arm next to the CPU reference arm (rank 0, N = 1 only) ------------------
    ate = None
    if rank == 0 and world == 1:
        from lfvio import violoop
        from oracle import backend as ob_
        from oracle import loop_ops

        d_opts = lfvio.solver_default_opts()
        d_opts.row = 960
        o_opts = ob_.default_opts()
        o_opts.row = 960
        win2 = lfvio.Window(10, 1000, 16384, device=local_rank)
        dev_loop, ate_dev = closed_loop_ate(violoop.device_ops(lfvio, win2, local_rank), d_opts)
        cpu_loop, ate_cpu = closed_loop_ate(loop_ops.oracle_ops(), o_opts)
        win2.close()
        diff = max(float(np.abs(dev_loop.estimates[k][0] - cpu_loop.estimates[k][0]).max()) for k in cpu_loop.estimates)
        ate = {"device": ate_dev, "cpu_ref": ate_cpu, "max_position_difference_m": diff, "scenario": ATE_SCENARIO}

    # ---- estimator initialisation: the numeric steps of initialStructure through the C-ABI with
    # host buffers next to the numpy port, same inputs (rank 0, N = 1 only; one-shot work, reported in ms, not in the metric)
    init_leg = None
    if rank == 0 and world == 1:
        init_leg = initialisation_leg(lfvio, local_rank)

    if rank == 0:
        fps = frames_total / (ms_total * 1e-3)
        line = {
            "metric": METRIC, "value": fps, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u8/f32 front end, f64 back end", "data": "synthetic",
            "config": {"workload": args.config, "image": f"{cols}x{rows}", "features": max_cnt, "tracked": n_tracks,
                       "window_frames": wsz + 1, "n_proj_factors": wprob.n_proj, "n_landmarks": wprob.L,
                       "prior_n": int(wprob.prior.n) if wprob.prior is not None else 0, "iterations": int(summ.iterations), "marginalize": with_marg, "pub_pattern": list(PUB_PATTERN), "reject_with_f": "device arm: on, every published frame; cpu arms: off (no native CPU port; the CPU arms do less work)",
                       "l2_policy": f"inputs larger than L2: {pool}-frame pool ({pool * px / 2**20:.0f} MiB) cycled",
                       "streams": "one independent stream per GPU; front end and back end on separate CUDA streams"},
            "ms_tracker_stream": ms_tracker / args.steps, "ms_window_stream": ms_window / args.steps,
            "wall_ms_per_step": 1e3 * t_wall / args.steps,
            "host_enqueue_ms_per_step": {"tracker_thread": 1e3 * enq[0] / args.steps, "window_thread": 1e3 * enq[1] / args.steps},
            "e2e": {"value": e2e_frames / e2e_s, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                    "steps": e2e_steps, "host_threads": 2, "single_thread_ms_per_frame": e2e_seq_ms, "node_ms_per_frame": e2e_node_ms,
                    "note": "two host threads = the reference's two ROS nodes; each makes blocking C-ABI calls with host buffers"},
            "gpu_launches": int(launches),
            "roofline": roofline,
            "cpu_baseline": cpu,
            "ate": ate,
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
