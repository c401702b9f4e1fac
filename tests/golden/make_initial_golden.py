"""Generates tests/golden/initial_seed1.npz: the numpy oracle's outputs (oracle/initial.py) for the estimator initialisation
on the synthetic start-up window tests/initial_data.InitScenario(seed=1, F=11) — inputs included, so the GPU tests can check
the device against stored vectors without re-running the oracle, and the CPU tests can detect drift of the oracle itself.

The reference cannot be run (Eigen / Ceres / OpenCV C++ / ROS absent) and has no vectors for these functions: these are
OUR golden vectors ("parity unpinned" against the reference, as stated in oracle/initial.py and DESIGN.md §5).
Run from the repo root in the build container:  python tests/golden/make_initial_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for p in (ROOT, os.path.join(ROOT, "lf-vio_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)

from initial_data import InitScenario  # noqa: E402
from oracle import initial as OI  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
SEED_RANSAC = 7


def main():
    sc = InitScenario(seed=1, F=11)
    l = 0
    corres = sc.corres(l, 10)
    ok, Rr, Tr, cnt, E = OI.solve_relative_rt(corres, SEED_RANSAC, l)
    sfm = OI.global_sfm(11, l, Rr, Tr, sc.obs_first, sc.obs_frame, sc.obs_pts)
    pre_ba = OI.global_sfm(11, l, Rr, Tr, sc.obs_first, sc.obs_frame, sc.obs_pts, ba=False)
    # alignment inputs: the frames expressed in camera l, scaled by 1 / 2.5
    Rs = np.array([sc.Rc[l].T @ sc.Rb[k] for k in range(11)])
    Ts = np.array([sc.Rc[l].T @ (sc.Pc[k] - sc.Pc[l]) for k in range(11)]) / 2.5
    pre0 = sc.preintegrations()
    dq = np.array([p.delta_q for p in pre0])
    Jrg = np.array([p.jacobian[3:6, 12:15] for p in pre0])
    dbg = OI.solve_gyroscope_bias(Rs, dq, Jrg)
    pre = sc.preintegrations(bg=dbg)
    dp, dv, dts = np.array([p.delta_p for p in pre]), np.array([p.delta_v for p in pre]), np.array([p.sum_dt for p in pre])
    ok_a, g, x = OI.linear_alignment(Rs, Ts, dp, dv, dts, sc.tic, sc.g_norm)
    # one EPnP problem with bearings in both half-spaces
    rng = np.random.default_rng(11)
    pw = rng.uniform(-4, 4, (50, 3))
    pc = pw @ sc.Rc[4] + np.array([0.3, -0.2, 0.1])
    us = pc / np.linalg.norm(pc, axis=1, keepdims=True)
    us = us + rng.normal(0, 5e-4, us.shape)
    us /= np.linalg.norm(us, axis=1, keepdims=True)
    R_pnp, T_pnp, e_pnp = OI.epnp_compute_pose(pw, us)
    np.savez_compressed(
        os.path.join(OUT, "initial_seed1.npz"),
        obs_first=sc.obs_first, obs_frame=sc.obs_frame, obs_pts=sc.obs_pts, corres=corres, ransac_seed=SEED_RANSAC, l=l,
        rel_ok=ok, rel_R=Rr, rel_T=Tr, rel_cnt=cnt, rel_E=E,
        sfm_q=sfm["q"], sfm_T=sfm["T"], sfm_points=sfm["points"], sfm_state=sfm["state"], sfm_ba_costs=np.array(sfm["ba"]["costs"]),
        sfm_ba_iterations=sfm["ba"]["iterations"], pre_ba_q=pre_ba["q"], pre_ba_T=pre_ba["T"], pre_ba_points=pre_ba["points"],
        al_Rs=Rs, al_Ts=Ts, al_dq=dq, al_Jrg=Jrg, al_dbg=dbg, al_dp=dp, al_dv=dv, al_dts=dts, tic=sc.tic, g_norm=sc.g_norm,
        al_ok=ok_a, al_g=g, al_x=x, pnp_pw=pw, pnp_us=us, pnp_R=R_pnp, pnp_T=T_pnp, pnp_err=e_pnp)
    print("wrote initial_seed1.npz:", os.path.getsize(os.path.join(OUT, "initial_seed1.npz")), "bytes")


if __name__ == "__main__":
    main()
