"""Device time per solve of lfvio_window_optimize_dev with and without torch in the process (CUDA events via the C-ABI stream are
not exposed without torch, so wall time over a long asynchronous run + one synchronising download is used)."""
import os, sys, time
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/lf-vio_b200')
if len(sys.argv) > 1 and sys.argv[1] == "torch":
    import torch
    torch.cuda.init()
import lfvio
from lfvio import synthwin, backend as B
o = lfvio.solver_default_opts(); o.fixed_iterations = 1
win = lfvio.Window(10, 1000, 16384)
w = synthwin.make_window_with_prior(win.optimize, win.marginalize, seed=1, opts=o, perturb=0.3)
win.upload(w)
for rep in range(3):
    for _ in range(50): win.optimize_dev(True, 0)
    win.download(w.clone())
    t0 = time.perf_counter()
    for _ in range(400): win.optimize_dev(True, 0)
    win.download(w.clone())
    print(sys.argv[1:], "ms per solve+marg: %.3f" % ((time.perf_counter() - t0) / 400 * 1e3))
