"""Prints the phase times of the fused step kernel (LFVIO_DEBUG_STEP=1) on the bench-shaped window with a prior."""
import os, sys
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/lf-vio_b200')
os.environ["LFVIO_DEBUG_STEP"] = "1"
import lfvio
from lfvio import synthwin
o = lfvio.solver_default_opts(); o.fixed_iterations = 1
win = lfvio.Window(10, 1000, 16384)
w = synthwin.make_window_with_prior(win.optimize, win.marginalize, seed=1, opts=o, perturb=0.3)
for _ in range(5):
    s = win.optimize(w.clone())
from lfvio import backend as B
pr = B.PriorData()
for _ in range(3):
    win.optimize_full(w.clone(), B.MARGIN_OLD, pr)
