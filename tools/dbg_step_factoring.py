"""Phase times of ONE factoring launch of the fused step kernel (LFVIO_DEBUG_STEP=1): max_num_iterations = 1, so the only
step of the solve factors the reduced system (a rejected step re-uses the factor and is ~10x shorter)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "lf-vio_b200"))
import lfvio
from lfvio import synthwin
o = lfvio.solver_default_opts(); o.fixed_iterations = 1
win = lfvio.Window(10, 1000, 16384)
w = synthwin.make_window_with_prior(win.optimize, win.marginalize, seed=1, opts=o, perturb=0.3)
o.max_num_iterations = 1
w.opts = o
win.upload(w)
for _ in range(600):
    win.optimize_dev(False, 0)
os.environ["LFVIO_DEBUG_STEP"] = "1"
win.download(w.clone())
