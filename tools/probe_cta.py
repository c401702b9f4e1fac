"""Per-CTA timelines of evaluate_kernel / accumulate_kernel (library built with make EXTRA=-DLFVIO_CTA_PROFILE)."""
import os, sys, ctypes as C
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/lf-vio_b200')
import numpy as np
import lfvio
from lfvio import synthwin, backend as B
o = lfvio.solver_default_opts(); o.fixed_iterations = 1
o.max_num_iterations = int(sys.argv[1]) if len(sys.argv) > 1 else 8
win = lfvio.Window(10, 1000, 16384)
w = synthwin.make_window_with_prior(win.optimize, win.marginalize, seed=1, opts=o, perturb=0.3)
win.upload(w)
for _ in range(5):
    win.optimize_dev(True, B.MARGIN_OLD)
s = win.download(w.clone())
L = lfvio.lib
nbp = 2 * (((w.n_proj + 127) // 128 + 7) // 8 * 8)  # two threads per factor, CTA count bucketed to 8 (window.cu)
for name, fn, ncta, groups in (("evaluate", L.lfvio_debug_eval_prof, nbp + w.n_imu + 1, [("proj", 0, nbp), ("imu", nbp, nbp + w.n_imu), ("prior", nbp + w.n_imu, nbp + w.n_imu + 1)]),
                               ("accumulate", L.lfvio_debug_acc_prof, w.L + w.n_imu + 8, [("landmark", 0, w.L), ("imu", w.L, w.L + w.n_imu), ("prior", w.L + w.n_imu, w.L + w.n_imu + 8)])):
    buf = np.zeros((ncta, 4), np.int64)
    fn.argtypes = [C.c_void_p, C.c_int]
    assert fn(buf.ctypes.data, ncta) == 0
    t0 = buf[:, 0].min()
    print(f"== {name}: {ncta} CTAs, kernel span {(buf[:, 1:].max() - t0) / 1e3:.2f} us (first CTA start -> last stamp)")
    for g, a, b in groups:
        x = buf[a:b]
        if len(x) == 0: continue
        last = np.where(x[:, 3] > 0, x[:, 3], x[:, 2])
        print(f"  {g:9s} n={b - a:4d} start {((x[:, 0] - t0) / 1e3).min():6.2f}..{((x[:, 0] - t0) / 1e3).max():6.2f}  wait-end {((x[:, 1] - t0) / 1e3).min():6.2f}..{((x[:, 1] - t0) / 1e3).max():6.2f}"
              f"  work {((x[:, 2] - x[:, 1]) / 1e3).mean():6.2f} avg {((x[:, 2] - x[:, 1]) / 1e3).max():6.2f} max  end {((last - t0) / 1e3).max():6.2f}")
