"""ctypes view of the estimator-initialisation entries of include/lfvio.h (lfvio_solve_gyro_bias, lfvio_linear_alignment,
lfvio_solve_relative_rt, lfvio_epnp, lfvio_global_sfm, lfvio_sfm_bundle_adjust) — tests and bench only."""
from __future__ import annotations

import ctypes as C

import numpy as np

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)
_u8p = C.POINTER(C.c_ubyte)
MAX_BA_ITER = 50


class SfmSummary(C.Structure):
    _fields_ = [("ok", C.c_int), ("fail_stage", C.c_int), ("fail_frame", C.c_int), ("ba_iterations", C.c_int),
                ("ba_successful", C.c_int), ("ba_converged", C.c_int), ("ba_final_cost", C.c_double),
                ("ba_costs", C.c_double * (MAX_BA_ITER + 1))]

    def costs(self):
        return [self.ba_costs[k] for k in range(self.ba_successful + 1)]


def bind(lib):
    lib.lfvio_solve_gyro_bias.argtypes = [_dp, _dp, _dp, C.c_int, _dp, C.c_int]
    lib.lfvio_linear_alignment.argtypes = [_dp, _dp, _dp, _dp, _dp, C.c_int, _dp, C.c_double, _dp, _dp, _ip, C.c_int]
    lib.lfvio_solve_relative_rt.argtypes = [_dp, C.c_int, C.c_uint64, C.c_uint, _dp, _dp, _ip, _dp, C.c_int]
    lib.lfvio_epnp.argtypes = [_dp, _dp, _ip, C.c_int, _dp, _dp, _dp, C.c_int]
    lib.lfvio_solve_relative_r.argtypes = [_dp, C.c_int, C.c_uint64, C.c_uint, _dp, C.c_int]
    lib.lfvio_calibrate_ex_rotation.argtypes = [_dp, _dp, _dp, C.c_int, _dp, _dp, C.c_int]
    lib.lfvio_global_sfm.argtypes = [C.c_int, C.c_int, _dp, _dp, _ip, _ip, _dp, C.c_int, C.c_int, _dp, _dp, _dp, _u8p,
                                     C.POINTER(SfmSummary), C.c_int]
    lib.lfvio_sfm_bundle_adjust.argtypes = [C.c_int, C.c_int, _ip, _ip, _dp, C.c_int, _u8p, _dp, _dp, _dp, C.POINTER(SfmSummary),
                                            C.c_int]


def _d(a, shape=None):
    a = np.ascontiguousarray(a, np.float64)
    return a if shape is None else a.reshape(shape)


def _p(a, t=C.c_double):
    return a.ctypes.data_as(C.POINTER(t))


def solve_gyro_bias(lib, check, Rs, delta_q, J_rg, device=0):
    Rs, dq, J = _d(Rs, (-1, 9)), _d(delta_q, (-1, 4)), _d(J_rg, (-1, 9))
    out = np.zeros(3)
    check(lib.lfvio_solve_gyro_bias(_p(Rs), _p(dq), _p(J), len(Rs), _p(out), device), "lfvio_solve_gyro_bias")
    return out


def linear_alignment(lib, check, Rs, Ts, delta_p, delta_v, sum_dt, tic, g_norm, device=0):
    """-> (ok, g, x) like oracle.initial.linear_alignment"""
    Rs, Ts, dp, dv, dt, tic = _d(Rs, (-1, 9)), _d(Ts, (-1, 3)), _d(delta_p, (-1, 3)), _d(delta_v, (-1, 3)), _d(sum_dt), _d(tic)
    n = len(Rs)
    g, x, ok = np.zeros(3), np.zeros(3 * n + 4), C.c_int(0)
    check(lib.lfvio_linear_alignment(_p(Rs), _p(Ts), _p(dp), _p(dv), _p(dt), n, _p(tic), float(g_norm), _p(g), _p(x), C.byref(ok),
                                     device), "lfvio_linear_alignment")
    return bool(ok.value), g, (x[:3 * n + 3] if ok.value else x)


def solve_relative_rt(lib, check, corres, seed, frame, device=0):
    """-> (ok, Rotation, Translation, inlier_cnt, E)"""
    c = _d(corres, (-1, 6))
    Rm, T, E, cnt = np.zeros((3, 3)), np.zeros(3), np.zeros((3, 3)), C.c_int(0)
    check(lib.lfvio_solve_relative_rt(_p(c), len(c), int(seed), int(frame), _p(Rm), _p(T), C.byref(cnt), _p(E), device),
          "lfvio_solve_relative_rt")
    return cnt.value > 12, Rm, T, cnt.value, E


def epnp(lib, check, problems, device=0):
    """problems: [(pws (n,3), us (n,3)), ...] -> (R (P,3,3), T (P,3), err (P,))"""
    first = np.zeros(len(problems) + 1, np.int32)
    for i, (pw, _) in enumerate(problems):
        first[i + 1] = first[i] + len(pw)
    pws = _d(np.concatenate([np.asarray(p[0], float).reshape(-1, 3) for p in problems]))
    us = _d(np.concatenate([np.asarray(p[1], float).reshape(-1, 3) for p in problems]))
    n = len(problems)
    Rm, T, err = np.zeros((n, 3, 3)), np.zeros((n, 3)), np.zeros(n)
    check(lib.lfvio_epnp(_p(pws), _p(us), _p(first, C.c_int), n, _p(Rm), _p(T), _p(err), device), "lfvio_epnp")
    return Rm, T, err


def global_sfm(lib, check, frame_num, l, relative_R, relative_T, obs_first, obs_frame, obs_pts, ba=True, device=0):
    of, ofr, pts = np.ascontiguousarray(obs_first, np.int32), np.ascontiguousarray(obs_frame, np.int32), _d(obs_pts, (-1, 3))
    L = len(of) - 1
    q, T, points, state = np.zeros((frame_num, 4)), np.zeros((frame_num, 3)), np.zeros((L, 3)), np.zeros(L, np.uint8)
    s = SfmSummary()
    rR, rT = _d(relative_R, (9,)), _d(relative_T, (3,))
    check(lib.lfvio_global_sfm(frame_num, l, _p(rR), _p(rT), _p(of, C.c_int), _p(ofr, C.c_int), _p(pts), L, int(ba), _p(q), _p(T),
                               _p(points), _p(state, C.c_ubyte), C.byref(s), device), "lfvio_global_sfm")
    return dict(ok=bool(s.ok), q=q, T=T, points=points, state=state.astype(bool), summary=s)


def bundle_adjust(lib, check, c_rot, c_trans, points, state, obs_first, obs_frame, obs_pts, l, device=0):
    """In place on c_rot (F,4) / c_trans (F,3) / points (L,3) (float64, C-contiguous).  -> SfmSummary"""
    of, ofr, pts = np.ascontiguousarray(obs_first, np.int32), np.ascontiguousarray(obs_frame, np.int32), _d(obs_pts, (-1, 3))
    st = np.ascontiguousarray(state, np.uint8)
    s = SfmSummary()
    for a in (c_rot, c_trans, points):
        assert a.dtype == np.float64 and a.flags.c_contiguous
    check(lib.lfvio_sfm_bundle_adjust(len(c_rot), l, _p(of, C.c_int), _p(ofr, C.c_int), _p(pts), len(st), _p(st, C.c_ubyte), _p(c_rot),
                                      _p(c_trans), _p(points), C.byref(s), device), "lfvio_sfm_bundle_adjust")
    return s


def solve_relative_r(lib, check, corres, seed, frame, device=0):
    c = _d(corres, (-1, 6))
    Rm = np.zeros((3, 3))
    check(lib.lfvio_solve_relative_r(_p(c), len(c), int(seed), int(frame), _p(Rm), device), "lfvio_solve_relative_r")
    return Rm


def calibrate_ex_rotation(lib, check, Rc, Rimu, Rc_g, device=0):
    """-> (ric, singular values (4,))"""
    Rc, Ri, Rg = _d(Rc, (-1, 9)), _d(Rimu, (-1, 9)), _d(Rc_g, (-1, 9))
    ric, sv = np.zeros((3, 3)), np.zeros(4)
    check(lib.lfvio_calibrate_ex_rotation(_p(Rc), _p(Ri), _p(Rg), len(Rc), _p(ric), _p(sv), device), "lfvio_calibrate_ex_rotation")
    return ric, sv
