"""Prints the closed-loop-through-pixels ATEs (device arm / CPU arm, 5 RANSAC seeds each). Usage: python tools/pixel_loop.py [n_frames]"""
import sys, time, json
sys.path[:0] = ['/root/repo', '/root/repo/lf-vio_b200', '/root/repo/tests']
import numpy as np
import lfvio
from lfvio import panoworld, violoop
from oracle import backend as ob, frontend as fe, loop_ops
from test_pixel_loop_gpu import cpu_ransac_hook, SEEDS
n = int(sys.argv[1]) if len(sys.argv) > 1 else 40
cam, tcfg = lfvio.PAL_UP, dict(lfvio.tracker_defaults(), max_cnt=150)
t0 = time.time(); w = panoworld.PixelWorld(cam, tcfg, seed=1, n_frames=n); print("render %.1f s" % (time.time() - t0))
o = ob.default_opts(); o.row = 960
cpu, dev = [], []
for s in SEEDS:
    w.track(fe.FeatureTrackerOracle(cam, tcfg, "cv2", ransac=cpu_ransac_hook(s)))
    cpu.append(violoop.trajectory_error(violoop.VioLoop(w, loop_ops.oracle_ops(), o).run())[0])
od = lfvio.solver_default_opts(); od.row = 960
win = lfvio.Window(10, 1000, 16384)
for s in SEEDS:
    tr = lfvio.FeatureTracker(cam, dict(tcfg, ransac=1, ransac_seed=s)); w.track(tr); tr.close()
    dev.append(violoop.trajectory_error(violoop.VioLoop(w, violoop.device_ops(lfvio, win, 0), od).run())[0])
print(json.dumps({"cpu_ate_m": cpu, "device_ate_m": dev, "tracked_min": int(min(w.n_tracked[2:]))}))
