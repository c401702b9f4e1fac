"""lfvio_window_optimize_full with host buffers next to the device-resident solve of the same window (lfvio_window_optimize_dev):
wall time per call, and with LFVIO_DEBUG_HOST=1 the host-side and device-side timeline of the call."""
import os, sys, time
os.environ["LFVIO_DEBUG_HOST"] = "1"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "lf-vio_b200"))
import lfvio
from lfvio import synthwin, backend as B
o = lfvio.solver_default_opts(); o.fixed_iterations = 1
win = lfvio.Window(10, 1000, 16384)
w = synthwin.make_window_with_prior(win.optimize, win.marginalize, seed=1, opts=o, perturb=0.3)
win.upload(w)
for rep in range(2):
    for _ in range(100): win.optimize_dev(True, 0)
    win.download(w.clone())
    t0 = time.perf_counter()
    for _ in range(400): win.optimize_dev(True, 0)
    win.download(w.clone())
    print("device-resident solve + marginalisation: %.1f us" % ((time.perf_counter() - t0) / 400 * 1e6))
work = w.clone()
call = win.prepared_optimize_full(work, B.MARGIN_OLD, B.PriorData())
init = (w.pose.copy(), w.sb.copy(), w.ex.copy(), w.feat.copy(), w.td.copy())
import numpy as np
for rep in range(3):
    t0 = time.perf_counter()
    for _ in range(200):
        for dst, s_ in zip((work.pose, work.sb, work.ex, work.feat, work.td), init): np.copyto(dst, s_)
        call()
    print("optimize_full with host buffers: %.1f us per call" % ((time.perf_counter() - t0) / 200 * 1e6))
