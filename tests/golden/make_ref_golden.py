"""Generates tests/golden/ref_backend.npz: outputs of the REFERENCE'S OWN factor / marginalisation / pre-integration /
alignment code (oracle/_ref, compiled unmodified from /root/reference by `make -C oracle ref`) on small seeded inputs,
stored together with those inputs so that the fixture needs neither the reference tree nor the compiled library.
Run here (the container with /root/reference mounted):  python tests/golden/make_ref_golden.py
"""
import ctypes as C
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path[:0] = [ROOT, os.path.join(ROOT, "lf-vio_b200"), os.path.join(ROOT, "tests")]

from lfvio import backend as B  # noqa: E402
from lfvio import prestep as PS  # noqa: E402
from lfvio import synthwin  # noqa: E402
from oracle import backend as ob  # noqa: E402
from oracle import ref  # noqa: E402


def pack_window(w, prefix, out):
    out[prefix + "pose"], out[prefix + "sb"], out[prefix + "ex"] = w.pose, w.sb, w.ex
    out[prefix + "feat"], out[prefix + "td"] = w.feat, w.td
    out[prefix + "proj"] = np.frombuffer(bytes(w.proj), np.uint8)[: w.n_proj * C.sizeof(B.ProjFactor)].copy()
    out[prefix + "imu"] = np.frombuffer(bytes(w.imu), np.uint8)[: w.n_imu * C.sizeof(B.ImuFactor)].copy()
    p = w.prior
    out[prefix + "prior_ids"], out[prefix + "prior_idx"] = p.block_id[: p.n_blocks].copy(), p.block_idx[: p.n_blocks].copy()
    out[prefix + "prior_J0"], out[prefix + "prior_r0"] = p.jacobian().copy(), p.residual().copy()
    gs = sum(1 if i == 2 * w.F + 1 else (9 if i % 2 and i < 2 * w.F else 7) for i in p.block_id[: p.n_blocks])
    out[prefix + "prior_x0"] = p.x0[:gs].copy()


def main():
    assert ref.available(), "needs /root/reference (or a prebuilt oracle/_ref/libref_factors.so)"
    out = {}
    kw = dict(n_landmarks=100, feats_per_frame=50)
    w = synthwin.make_window_with_prior(ob.optimize, ob.marginalize, seed=7, opts=ob.default_opts(), perturb=0.3, **kw)
    w.opts.row, w.opts.tr = 961, 0.02  # odd ROW and a non-zero read-out time: exercises ROW / 2 as a floating division
    pack_window(w, "w_", out)
    out["w_row_tr"] = np.array([w.opts.row, w.opts.tr])
    for robust in (0, 1):
        for td in (1, 0):
            w.opts.estimate_td = td
            e = ref.evaluate(w, bool(robust))
            k = f"eval_r{robust}_td{td}_"
            out[k + "res_proj"], out[k + "jac_proj"] = e.res_proj, e.jac_proj
            out[k + "res_imu"], out[k + "jac_imu"], out[k + "res_prior"] = e.res_imu, e.jac_imu, e.res_prior
            out[k + "cost"] = np.array([e.cost.value])
    w.opts.estimate_td = 1
    for flag in (B.MARGIN_OLD, B.MARGIN_SECOND_NEW):
        pr, m = ref.marginalize(w, flag)
        A, g, x0 = ref.canonical_information(pr, w.F)
        out[f"marg{flag}_A"], out[f"marg{flag}_g"], out[f"marg{flag}_m_n"] = A, g, np.array([m, pr.n])
        out[f"marg{flag}_ids"] = np.array(sorted(x0), np.int32)
        out[f"marg{flag}_x0"] = np.concatenate([x0[k] for k in sorted(x0)])
        out[f"marg{flag}_r0sq"] = np.array([pr.residual() @ pr.residual()])
    # pre-integration of three intervals
    from prestep_data import G, NOISE, imu_stream

    s = imu_stream(9, n_intervals=3)
    ivs = [PS.make_interval(f, c, i, a0, g0, s["ba"], s["bg"]) for i, (f, c, a0, g0) in enumerate(s["intervals"])]
    arr = (PS.ImuInterval * len(ivs))(*ivs)
    res = (B.ImuFactor * len(ivs))()
    nz = PS.ImuNoise(*NOISE)
    dp = C.POINTER(C.c_double)
    L = ref.lib()
    L.ref_preintegrate.argtypes = [C.POINTER(PS.ImuInterval), C.c_int, dp, dp, dp, C.c_int, C.POINTER(PS.ImuNoise), dp, C.POINTER(B.ImuFactor)]
    assert L.ref_preintegrate(arr, len(ivs), s["dt"].ctypes.data_as(dp), s["acc"].ctypes.data_as(dp), s["gyr"].ctypes.data_as(dp),
                              len(s["dt"]), C.byref(nz), G.ctypes.data_as(dp), res) == 0
    out["pre_dt"], out["pre_acc"], out["pre_gyr"] = s["dt"], s["acc"], s["gyr"]
    out["pre_intervals"] = np.frombuffer(bytes(arr), np.uint8).copy()
    out["pre_noise_g"] = np.concatenate([NOISE, G])
    out["pre_out"] = np.frombuffer(bytes(res), np.uint8).copy()
    # visual-inertial alignment (solveGyroscopeBias, LinearAlignment + RefineGravity)
    from initial_data import InitScenario

    sc = InitScenario(seed=3, F=11, pixel_noise=0.3)
    imu = [sc.world.imu[k] for k in range(1, sc.F)]
    Rs, Ts = sc.Rb, sc.Pc * 0.41
    r0 = ref.visual_imu_alignment(Rs, Ts, imu, sc.world.noise, sc.tic, sc.g_norm, mode=0)
    r1 = ref.visual_imu_alignment(Rs, Ts, imu, sc.world.noise, sc.tic, sc.g_norm, mode=1)
    pre = r1["pre"]
    out["al_R"], out["al_T"], out["al_tic_gnorm"] = Rs, Ts, np.concatenate([sc.tic, [sc.g_norm]])
    out["al_dp"] = np.array([list(p.delta_p) for p in pre])
    out["al_dv"] = np.array([list(p.delta_v) for p in pre])
    out["al_dq"] = np.array([list(p.delta_q) for p in pre])
    out["al_dt"] = np.array([p.sum_dt for p in pre])
    out["al_Jrg"] = np.array([np.array(p.jacobian).reshape(15, 15)[3:6, 12:15] for p in pre])
    out["al_bg"], out["al_g"], out["al_x"], out["al_ok"] = r0["bg"], r1["g"], r1["x"], np.array([int(r1["ok"])])
    # rejectWithF's RANSAC (the reference's compute_E_21 / check_inliers / myfindFundamentalMat on the injected random source)
    import lfvio
    from oracle import ransac as R
    from ransac_data import two_view

    cur, forw, _ = two_view(lfvio.PAL_UP, lfvio.tracker_defaults(), 150, 12)
    b1, b2 = R.lift_pair(lfvio.PAL_UP, cur, forw)
    E, st = ref.find_fundamental(b1, b2, 5, 3)
    out["rs_cur"], out["rs_forw"], out["rs_seed_frame"] = cur, forw, np.array([5, 3])
    out["rs_status"], out["rs_E"] = st, E
    np.savez_compressed(os.path.join(HERE, "ref_backend.npz"), **out)
    print("wrote ref_backend.npz:", sum(v.nbytes for v in out.values()) // 1024, "KiB raw,", len(out), "arrays")


if __name__ == "__main__":
    main()
