"""Times the tracker stream and the window stream alone and together (device-resident inputs), to separate device-side
bubbles from host enqueue limits. Usage: python tools/probe_streams.py [steps]"""
import os, sys, time, threading
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/lf-vio_b200')
import numpy as np, torch
import lfvio
from lfvio import backend as B
import bench

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 300
cfg = sys.argv[2] if len(sys.argv) > 2 else "palvio_id01_1280x960_150feat_10kf"
scale, max_cnt, wsz, n_lm, hz = bench.CONFIGS[cfg]
rows, cols = int(round(960 * scale)), int(round(1280 * scale))
pool = max(16, int(np.ceil(1.25 * 126 * 2 ** 20 / (rows * cols))))
wl = bench.build_workload(cfg, 0, pool)
dev = torch.device("cuda", 0)
frames_dev = torch.from_numpy(wl["frames"]).to(dev)
tracker = lfvio.FeatureTracker(wl["cam"], wl["tcfg"], device=0)
opts = lfvio.solver_default_opts(); opts.fixed_iterations = 1; opts.row = rows
window = lfvio.Window(wsz, max(1000, 2 * n_lm), max(16384, 24 * n_lm), device=0)
wprob = bench.make_window_problem(wl, opts, 0, window.optimize, window.marginalize)
window.upload(wprob)
ts = torch.cuda.ExternalStream(tracker.stream, device=dev); ws = torch.cuda.ExternalStream(window.stream, device=dev)
P = bench.PUB_PATTERN

def run(do_t, do_w, threaded):
    def et():
        for k in range(20, 20 + steps):
            tracker.read_image_dev(frames_dev[k % pool].data_ptr(), cols, k / hz, bool(P[k % 3]))
        e[2].record(ts)
    def ew():
        for k in range(steps):
            window.optimize_dev(True, B.MARGIN_OLD)
        e[3].record(ws)
    for k in range(20):
        if do_t: tracker.read_image_dev(frames_dev[k % pool].data_ptr(), cols, k / hz, bool(P[k % 3]))
        if do_w: window.optimize_dev(True, B.MARGIN_OLD)
    torch.cuda.synchronize()
    e = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    e[0].record(ts); e[1].record(ws)
    t0 = time.perf_counter()
    if threaded:
        th = ([threading.Thread(target=et)] if do_t else []) + ([threading.Thread(target=ew)] if do_w else [])
        [t.start() for t in th]; [t.join() for t in th]
    else:
        if do_t: et()
        if do_w: ew()
    host = time.perf_counter() - t0
    if not do_t: e[2].record(ts)
    if not do_w: e[3].record(ws)
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    tracker.sync()
    return dict(tracker_ms=e[0].elapsed_time(e[2]) / steps, window_ms=e[1].elapsed_time(e[3]) / steps, host_ms=1e3 * host / steps, wall_ms=1e3 * wall / steps)

print("tracker alone     ", run(True, False, False))
print("window alone      ", run(False, True, False))
print("both, two threads ", run(True, True, True))
print("both, one thread  ", run(True, True, False))
print("launch count", lfvio.lib.lfvio_launch_count())
