"""Synthetic sliding windows for the back end (data generation for tests and bench.py — not product compute).

A rig moves on a smooth trajectory through a cloud of landmarks.  The generator produces what
Estimator::optimization() consumes: F window states, 200 Hz IMU pre-integrated per frame interval with the
reference's mid-point scheme (IntegrationBase::midPointIntegration, vins_estimator/src/factor/integration_base.h:54-158;
noise densities of config/mindvision/mindvision.yaml:138-142), unit-bearing observations with bearing velocities
and per-observation td, in the reference's factor order (estimator.cpp:717-772).
"""
from __future__ import annotations

import numpy as np

from . import backend as B
from .config import ESTIMATOR


# ---- small SO(3) helpers -----------------------------------------------------------------------
def quat_mul(a, b):  # (w, x, y, z)
    aw, ax, ay, az = a
    bw, bx, by, bz = b
    return np.array([aw * bw - ax * bx - ay * by - az * bz, aw * bx + ax * bw + ay * bz - az * by,
                     aw * by + ay * bw + az * bx - ax * bz, aw * bz + az * bw + ax * by - ay * bx])


def quat_to_rot(q):
    w, x, y, z = q
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y)],
                     [2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x)],
                     [2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)]])


def rot_to_quat(R):
    t = np.trace(R)
    if t > 0:
        s = np.sqrt(t + 1.0) * 2
        q = np.array([0.25 * s, (R[2, 1] - R[1, 2]) / s, (R[0, 2] - R[2, 0]) / s, (R[1, 0] - R[0, 1]) / s])
    else:
        i = int(np.argmax(np.diag(R)))
        j, k = (i + 1) % 3, (i + 2) % 3
        s = np.sqrt(1.0 + R[i, i] - R[j, j] - R[k, k]) * 2
        q = np.zeros(4)
        q[0] = (R[k, j] - R[j, k]) / s
        q[1 + i] = 0.25 * s
        q[1 + j] = (R[j, i] + R[i, j]) / s
        q[1 + k] = (R[k, i] + R[i, k]) / s
    if q[0] < 0:
        q = -q
    return q / np.linalg.norm(q)


def skew(v):
    return np.array([[0, -v[2], v[1]], [v[2], 0, -v[0]], [-v[1], v[0], 0]])


def rot_xyz(rz, ry, rx):
    cz, sz, cy, sy, cx, sx = np.cos(rz), np.sin(rz), np.cos(ry), np.sin(ry), np.cos(rx), np.sin(rx)
    Rz = np.array([[cz, -sz, 0], [sz, cz, 0], [0, 0, 1]])
    Ry = np.array([[cy, 0, sy], [0, 1, 0], [-sy, 0, cy]])
    Rx = np.array([[1, 0, 0], [0, cx, -sx], [0, sx, cx]])
    return Rz @ Ry @ Rx


# ---- IMU pre-integration (restated IntegrationBase) ---------------------------------------------
class PreIntegration:
    def __init__(self, acc0, gyr0, ba, bg, noise):
        self.acc0, self.gyr0 = np.array(acc0, float), np.array(gyr0, float)
        self.ba, self.bg = np.array(ba, float), np.array(bg, float)
        self.dp, self.dv, self.dq = np.zeros(3), np.zeros(3), np.array([1.0, 0, 0, 0])
        self.J, self.P, self.sum_dt = np.eye(15), np.zeros((15, 15)), 0.0
        an, gn, aw, gw = noise
        self.noise = np.diag([an * an] * 3 + [gn * gn] * 3 + [an * an] * 3 + [gn * gn] * 3 + [aw * aw] * 3 + [gw * gw] * 3)

    def push(self, dt, acc1, gyr1):
        acc1, gyr1 = np.array(acc1, float), np.array(gyr1, float)
        R0 = quat_to_rot(self.dq)
        un_acc_0 = R0 @ (self.acc0 - self.ba)
        un_gyr = 0.5 * (self.gyr0 + gyr1) - self.bg
        dq1 = quat_mul(self.dq, np.array([1.0, un_gyr[0] * dt / 2, un_gyr[1] * dt / 2, un_gyr[2] * dt / 2]))
        R1 = quat_to_rot(dq1)  # Eigen toRotationMatrix of the un-normalised product (normalised after the step)
        un_acc_1 = R1 @ (acc1 - self.ba)
        un_acc = 0.5 * (un_acc_0 + un_acc_1)
        dp1 = self.dp + self.dv * dt + 0.5 * un_acc * dt * dt
        dv1 = self.dv + un_acc * dt
        Rw, Ra0, Ra1 = skew(un_gyr), skew(self.acc0 - self.ba), skew(acc1 - self.ba)
        I3 = np.eye(3)
        F = np.zeros((15, 15))
        F[0:3, 0:3] = I3
        F[0:3, 3:6] = -0.25 * R0 @ Ra0 * dt * dt + -0.25 * R1 @ Ra1 @ (I3 - Rw * dt) * dt * dt
        F[0:3, 6:9] = I3 * dt
        F[0:3, 9:12] = -0.25 * (R0 + R1) * dt * dt
        F[0:3, 12:15] = -0.25 * R1 @ Ra1 * dt * dt * -dt
        F[3:6, 3:6] = I3 - Rw * dt
        F[3:6, 12:15] = -I3 * dt
        F[6:9, 3:6] = -0.5 * R0 @ Ra0 * dt + -0.5 * R1 @ Ra1 @ (I3 - Rw * dt) * dt
        F[6:9, 6:9] = I3
        F[6:9, 9:12] = -0.5 * (R0 + R1) * dt
        F[6:9, 12:15] = -0.5 * R1 @ Ra1 * dt * -dt
        F[9:12, 9:12] = I3
        F[12:15, 12:15] = I3
        V = np.zeros((15, 18))
        V[0:3, 0:3] = 0.25 * R0 * dt * dt
        V[0:3, 3:6] = 0.25 * -R1 @ Ra1 * dt * dt * 0.5 * dt
        V[0:3, 6:9] = 0.25 * R1 * dt * dt
        V[0:3, 9:12] = V[0:3, 3:6]
        V[3:6, 3:6] = 0.5 * I3 * dt
        V[3:6, 9:12] = 0.5 * I3 * dt
        V[6:9, 0:3] = 0.5 * R0 * dt
        V[6:9, 3:6] = 0.5 * -R1 @ Ra1 * dt * 0.5 * dt
        V[6:9, 6:9] = 0.5 * R1 * dt
        V[6:9, 9:12] = V[6:9, 3:6]
        V[9:12, 12:15] = I3 * dt
        V[12:15, 15:18] = I3 * dt
        self.J = F @ self.J
        self.P = F @ self.P @ F.T + V @ self.noise @ V.T
        self.dp, self.dv = dp1, dv1
        self.dq = dq1 / np.linalg.norm(dq1)
        self.sum_dt += dt
        self.acc0, self.gyr0 = acc1, gyr1

    def to_factor(self, frame_i: int) -> B.ImuFactor:
        f = B.ImuFactor()
        f.frame_i, f.sum_dt = frame_i, self.sum_dt
        f.delta_p[:] = self.dp.tolist()
        f.delta_q[:] = self.dq.tolist()
        f.delta_v[:] = self.dv.tolist()
        f.linearized_ba[:] = self.ba.tolist()
        f.linearized_bg[:] = self.bg.tolist()
        f.jacobian[:] = self.J.ravel().tolist()
        f.covariance[:] = self.P.ravel().tolist()
        return f


# ---- trajectory ----------------------------------------------------------------------------------
class Trajectory:
    def __init__(self, seed: int):
        rng = np.random.default_rng(seed)
        self.a = rng.uniform(0.8, 1.6, 3)
        self.w = rng.uniform(0.25, 0.6, 3)
        self.ph = rng.uniform(0, 2 * np.pi, 6)

    def pos(self, t):
        return np.array([self.a[0] * np.sin(self.w[0] * t + self.ph[0]), self.a[1] * np.cos(self.w[1] * t + self.ph[1]),
                         0.4 * self.a[2] * np.sin(self.w[2] * t + self.ph[2])])

    def rot(self, t):
        return rot_xyz(0.35 * t + 0.3 * np.sin(0.3 * t + self.ph[3]), 0.12 * np.sin(0.5 * t + self.ph[4]),
                       0.12 * np.cos(0.4 * t + self.ph[5]))

    def vel(self, t, h=1e-4):
        return (self.pos(t + h) - self.pos(t - h)) / (2 * h)

    def acc(self, t, h=1e-3):
        return (self.pos(t + h) - 2 * self.pos(t) + self.pos(t - h)) / (h * h)

    def omega_body(self, t, h=1e-4):
        R0, R1 = self.rot(t - h), self.rot(t + h)
        dR = self.rot(t).T @ (R1 - R0) / (2 * h)
        return np.array([dR[2, 1] - dR[1, 2], dR[0, 2] - dR[2, 0], dR[1, 0] - dR[0, 1]]) / 2


def make_window(seed: int = 1, n_frames: int = 11, n_landmarks: int = 260, feats_per_frame: int = 150, frame_dt: float = 0.1,
                imu_rate: float = 200.0, t0: float = 5.0, perturb: float = 1.0, pixel_noise: float = 0.3, opts=None,
                estimate_td: int = 1, outlier_frac: float = 0.02) -> B.WindowProblem:
    """One window (no prior).  perturb scales the offset of the initial state from ground truth."""
    rng = np.random.default_rng(1000 + seed)
    E = ESTIMATOR
    traj = Trajectory(seed)
    G = np.array([0, 0, E["g_norm"]])
    ric = np.array(E["extrinsicRotation"]).reshape(3, 3)
    tic = np.array(E["extrinsicTranslation"])
    ba_true, bg_true = rng.normal(0, 0.02, 3), rng.normal(0, 0.002, 3)
    td_true = E["td"]
    F = n_frames
    times = t0 + frame_dt * np.arange(F)
    Rs = [traj.rot(t) for t in times]
    Ps = [traj.pos(t) for t in times]
    Vs = [traj.vel(t) for t in times]
    # ---- IMU factors ----
    noise = (E["acc_n"], E["gyr_n"], E["acc_w"], E["gyr_w"])
    n_sub = int(round(frame_dt * imu_rate))
    dt = frame_dt / n_sub

    def imu_at(t):
        acc = traj.rot(t).T @ (traj.acc(t) + G) + ba_true + rng.normal(0, E["acc_n"], 3)
        gyr = traj.omega_body(t) + bg_true + rng.normal(0, E["gyr_n"], 3)
        return acc, gyr

    ba_lin, bg_lin = ba_true + rng.normal(0, 0.01, 3), bg_true + rng.normal(0, 0.001, 3)
    imu = []
    for i in range(F - 1):
        a0, g0 = imu_at(times[i])
        pre = PreIntegration(a0, g0, ba_lin, bg_lin, noise)
        for k in range(1, n_sub + 1):
            a1, g1 = imu_at(times[i] + k * dt)
            pre.push(dt, a1, g1)
        imu.append(pre.to_factor(i))
    # ---- landmarks + observations ----
    ang = rng.uniform(0, 2 * np.pi, n_landmarks)
    rad = rng.uniform(3.0, 8.0, n_landmarks)
    hgt = rng.uniform(-2.0, 3.0, n_landmarks)
    pw = np.stack([rad * np.cos(ang), rad * np.sin(ang), hgt], 1)
    focal = E["focal_length"]
    sigma = pixel_noise / focal

    def bearing(l, f, t_shift=0.0):
        t = times[f] + t_shift
        pc = ric.T @ (traj.rot(t).T @ (pw[l] - traj.pos(t)) - tic)
        return pc / np.linalg.norm(pc), np.linalg.norm(pc)

    # track start/length so that ~feats_per_frame are visible in every frame
    start = rng.integers(0, max(F - 3, 1), n_landmarks)
    start[: n_landmarks // 3] = 0
    length = rng.integers(2, F + 1, n_landmarks)
    proj, feat_depth, lm_index = [], [], {}
    per_frame = np.zeros(F, int)
    wsz = F - 1
    for l in np.argsort(start, kind="stable"):
        s = int(start[l])
        e = min(F, s + int(length[l]))
        if per_frame[s:e].max(initial=0) >= feats_per_frame:
            continue
        used = e - s
        if not (used >= 2 and s < wsz - 2):  # estimator.cpp:730
            continue
        per_frame[s:e] += 1
        idx = len(feat_depth)
        lm_index[l] = idx
        obs = []
        for f in range(s, e):
            b, dist = bearing(l, f, td_true)  # the image was really exposed td later than its stamp
            b = b + rng.normal(0, sigma, 3)
            if rng.uniform() < outlier_frac and f > s:
                b = b + rng.normal(0, 30 * sigma, 3)
            b = b / np.linalg.norm(b)
            obs.append(b.astype(np.float32).astype(np.float64))  # bearings travel as float32 in the PointCloud msg
        vel = [np.zeros(3)] + [((obs[k] - obs[k - 1]) / frame_dt).astype(np.float32).astype(np.float64) for k in range(1, len(obs))]
        _, d0 = bearing(l, s, td_true)
        feat_depth.append(1.0 / (d0 * (1.0 + perturb * rng.normal(0, 0.08))))
        for k in range(1, len(obs)):
            pf = B.ProjFactor()
            pf.frame_i, pf.frame_j, pf.landmark = s, s + k, idx
            pf.pts_i[:] = obs[0].tolist()
            pf.pts_j[:] = obs[k].tolist()
            pf.vel_i[:] = vel[0].tolist()
            pf.vel_j[:] = vel[k].tolist()
            pf.td_i = pf.td_j = td_true + 0.002
            pf.row_i, pf.row_j = 480 + 300 * obs[0][1], 480 + 300 * obs[k][1]
            proj.append(pf)
    # ---- initial state = truth + perturbation ----
    pose = np.zeros((F, 7))
    sb = np.zeros((F, 9))
    for f in range(F):
        dth = perturb * rng.normal(0, 0.01, 3) * (f > 0)
        R = Rs[f] @ rot_xyz(dth[2], dth[1], dth[0])
        q = rot_to_quat(R)
        pose[f, 0:3] = Ps[f] + perturb * rng.normal(0, 0.03, 3) * (f > 0)
        pose[f, 3:7] = [q[1], q[2], q[3], q[0]]
        sb[f, 0:3] = Vs[f] + perturb * rng.normal(0, 0.05, 3)
        sb[f, 3:6] = ba_lin
        sb[f, 6:9] = bg_lin
    qe = rot_to_quat(ric @ rot_xyz(*(perturb * rng.normal(0, 0.005, 3))))
    ex = np.concatenate([tic + perturb * rng.normal(0, 0.005, 3), [qe[1], qe[2], qe[3], qe[0]]])
    if opts is None:
        raise ValueError("pass solver opts (lfvio.backend.default_opts or oracle.backend.default_opts)")
    opts.estimate_td = estimate_td
    w = B.WindowProblem(F, pose, sb, ex, np.array(feat_depth), td_true + perturb * 0.003, proj, imu, None, opts)
    w.truth = dict(P=np.array(Ps), R=np.array(Rs), V=np.array(Vs), ba=ba_true, bg=bg_true, td=td_true)
    return w


def make_window_with_prior(optimize, marginalize, seed: int = 1, t0: float = 5.0, frame_dt: float = 0.1, depth: int = 3,
                           **kw) -> B.WindowProblem:
    """Steady-state window: the `depth` windows before t0 (same trajectory, one frame apart) are optimised in turn and
    their oldest frame marginalised (MARGIN_OLD, estimator.cpp:833-939), each one carrying the prior left by its
    predecessor; the last prior (n = 6 (F - 1) + 9 + 6 + 1) is attached to the window at t0, whose frames 0 .. F-2 are
    the frames 1 .. F-1 that prior was linearised on.  depth >= 3 reaches the regime a running estimator is in (the
    prior's information matrix numerically positive definite apart from the gauge directions).
    optimize(w) / marginalize(w, flag) are the device or the oracle implementations (same signatures)."""
    prior = None
    for k in range(depth, 0, -1):
        prev = make_window(seed=seed, t0=t0 - k * frame_dt, frame_dt=frame_dt, **kw)
        prev.prior = prior
        optimize(prev)
        prior = marginalize(prev, B.MARGIN_OLD)
    w = make_window(seed=seed, t0=t0, frame_dt=frame_dt, **kw)
    w.prior = prior
    return w
