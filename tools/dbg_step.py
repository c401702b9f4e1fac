"""Prints the phase times of the fused step kernel (LFVIO_DEBUG_STEP=1) on the bench-shaped window with a prior."""
import os, sys
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/lf-vio_b200')
os.environ["LFVIO_DEBUG_STEP"] = "1"
import lfvio
from lfvio import synthwin
o = lfvio.solver_default_opts(); o.fixed_iterations = 1
win = lfvio.Window(10, 1000, 16384)
w = synthwin.make_window_with_prior(win.optimize, win.marginalize, seed=1, opts=o, perturb=0.3)
import io, contextlib
win.upload(w)
for _ in range(400):  # bring the SM clock up before the measured solves
    win.optimize_dev(True, 0)
import time, sys
t0 = time.perf_counter()
for _ in range(200):
    win.optimize_dev(False, 0)
s = win.download(w.clone())
print("optimize_dev: %.3f ms per solve (no marginalisation)" % ((time.perf_counter() - t0) / 200 * 1e3), file=sys.stderr)
t0 = time.perf_counter()
for _ in range(50):
    s = win.optimize(w.clone())
print("optimize: %.3f ms per call" % ((time.perf_counter() - t0) / 50 * 1e3), file=sys.stderr)
from lfvio import backend as B
pr = B.PriorData()
for _ in range(3):
    win.optimize_full(w.clone(), B.MARGIN_OLD, pr)
