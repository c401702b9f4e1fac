"""CPU arm of the closed-loop VIO (lfvio.violoop): the oracle's numeric steps behind the same Ops interface —
TEST INFRASTRUCTURE ONLY (tests/ and bench.py's CPU leg).

optimize_full restates what lfvio_window_optimize_full fuses: ceres::Solve, Estimator::double2vector's yaw / position
re-anchoring (vins_estimator/src/estimator.cpp:532-583) followed by vector2double (:488-530), then the MARGIN_OLD prior.
"""
from __future__ import annotations

import numpy as np

from lfvio import backend as B
from lfvio import synthwin, violoop
from oracle import backend as ob
from oracle import prestep as O


def _R(q_xyzw, normalize):
    q = np.array([q_xyzw[3], q_xyzw[0], q_xyzw[1], q_xyzw[2]])
    if normalize:
        q = q / np.linalg.norm(q)
    return O._qrot(q)


def _yaw_pitch_deg(R):  # Utility::R2ypr (utility/utility.h:66-81)
    n = R[:, 0]
    y = np.arctan2(n[1], n[0])
    p = np.arctan2(-n[2], n[0] * np.cos(y) + n[1] * np.sin(y))
    return np.degrees(y), np.degrees(p)


def double2vector(before: B.WindowProblem, w: B.WindowProblem):
    R0_before, P0_before = _R(before.pose[0, 3:], False), before.pose[0, :3].copy()
    y0, p0 = _yaw_pitch_deg(R0_before)
    R00 = _R(w.pose[0, 3:], False)
    y00, p00 = _yaw_pitch_deg(R00)
    a = np.radians(y0 - y00)
    rot_diff = np.array([[np.cos(a), -np.sin(a), 0], [np.sin(a), np.cos(a), 0], [0, 0, 1]])
    if abs(abs(p0) - 90) < 1.0 or abs(abs(p00) - 90) < 1.0:
        rot_diff = R0_before @ R00.T
    P00 = w.pose[0, :3].copy()
    for i in range(w.F):
        Ri = rot_diff @ _R(w.pose[i, 3:], True)
        w.pose[i, :3] = rot_diff @ (w.pose[i, :3] - P00) + P0_before
        q = synthwin.rot_to_quat(Ri)
        w.pose[i, 3:] = [q[1], q[2], q[3], q[0]]
        w.sb[i, :3] = rot_diff @ w.sb[i, :3]
    q = synthwin.rot_to_quat(_R(w.ex[3:], False))
    w.ex[3:] = [q[1], q[2], q[3], q[0]]


def optimize_full(w: B.WindowProblem, flag: int):
    before = w.clone()
    s = ob.optimize(w)
    double2vector(before, w)
    return s, ob.marginalize(w, flag, 4)


def preintegrate(intervals, dt, acc, gyr, noise, g):
    out, state = [], np.zeros((len(intervals), 15))
    for i, iv in enumerate(intervals):
        ib = O.IntegrationOracle(list(iv.acc_0), list(iv.gyr_0), list(iv.ba), list(iv.bg), noise)
        samples = [(dt[k], acc[k], gyr[k]) for k in range(iv.first, iv.first + iv.count)]
        for s in samples:
            ib.push_back(*s)
        f = B.ImuFactor()
        f.frame_i, f.sum_dt = iv.frame_i, ib.sum_dt
        f.delta_p[:], f.delta_q[:], f.delta_v[:] = ib.delta_p.tolist(), ib.delta_q.tolist(), ib.delta_v.tolist()
        f.linearized_ba[:], f.linearized_bg[:] = list(iv.ba), list(iv.bg)
        f.jacobian[:], f.covariance[:] = ib.jacobian.ravel().tolist(), ib.covariance.ravel().tolist()
        out.append(f)
        if iv.propagate:
            P, V, R = O.process_imu(list(iv.P), list(iv.V), np.array(list(iv.R)).reshape(3, 3), np.array(list(iv.ba)),
                                    np.array(list(iv.bg)), np.asarray(g, float), list(iv.acc_0), list(iv.gyr_0), samples)
            state[i] = np.concatenate([P, V, R.ravel()])
    return out, state


def oracle_ops() -> violoop.Ops:
    return violoop.Ops("oracle", optimize_full, preintegrate,
                       lambda Ps, Rs, tic, ric, start, first, pts, depth, init: O.triangulate(Ps, Rs, tic, ric, start, first, pts, depth, init))
