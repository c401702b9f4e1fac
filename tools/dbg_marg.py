import sys; sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/lf-vio_b200')
import lfvio
from lfvio import synthwin
o=lfvio.solver_default_opts(); o.fixed_iterations=1
w=synthwin.make_window(seed=1,opts=o,perturb=0.3)
win=lfvio.Window(10,1000,16384)
win.optimize(w)
for _ in range(5): win.marginalize(w, 0)
