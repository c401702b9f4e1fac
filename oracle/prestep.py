"""CPU oracle (TEST INFRASTRUCTURE ONLY) for the two steps before Estimator::optimization():

  IntegrationOracle   restates IntegrationBase (vins_estimator/src/factor/integration_base.h:13-157): constructor,
                      push_back -> propagate -> midPointIntegration with update_jacobian = 1, and the nominal-state
                      prediction of Estimator::processIMU (vins_estimator/src/estimator.cpp:86-120).
  triangulate         restates FeatureManager::triangulate (vins_estimator/src/feature_manager.cpp:199-253); the
                      Eigen::JacobiSVD of the reference is replaced by numpy's LAPACK SVD (same right singular vector
                      up to sign, which cancels in X / X_w).

Parity unpinned against a reference binary: the reference has no tests or vectors for these functions and cannot be built
here (Eigen absent).  Pinned instead by closed-form cases, finite differences of the bias Jacobians and noise-free
geometry in tests/test_oracle_prestep.py.  Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import this.
"""
from __future__ import annotations

import numpy as np


def _skew(v):
    return np.array([[0, -v[2], v[1]], [v[2], 0, -v[0]], [-v[1], v[0], 0]], float)


def _qmul(a, b):  # Hamilton, (w, x, y, z)
    return np.array([a[0] * b[0] - a[1] * b[1] - a[2] * b[2] - a[3] * b[3],
                     a[0] * b[1] + a[1] * b[0] + a[2] * b[3] - a[3] * b[2],
                     a[0] * b[2] + a[2] * b[0] + a[3] * b[1] - a[1] * b[3],
                     a[0] * b[3] + a[3] * b[0] + a[1] * b[2] - a[2] * b[1]])


def _qrot(q):  # Eigen::Quaternion::toRotationMatrix (no normalisation)
    w, x, y, z = q
    tx, ty, tz = 2 * x, 2 * y, 2 * z
    twx, twy, twz = tx * w, ty * w, tz * w
    txx, txy, txz = tx * x, ty * x, tz * x
    tyy, tyz, tzz = ty * y, tz * y, tz * z
    return np.array([[1 - (tyy + tzz), txy - twz, txz + twy], [txy + twz, 1 - (txx + tzz), tyz - twx],
                     [txz - twy, tyz + twx, 1 - (txx + tyy)]])


class IntegrationOracle:
    """integration_base.h:13-29 (ctor, noise), :31-37 (push_back), :55-127 (midPointIntegration), :129-157 (propagate)."""

    def __init__(self, acc_0, gyr_0, ba, bg, noise):
        self.acc_0, self.gyr_0 = np.array(acc_0, float), np.array(gyr_0, float)
        self.linearized_ba, self.linearized_bg = np.array(ba, float), np.array(bg, float)
        self.jacobian, self.covariance = np.eye(15), np.zeros((15, 15))
        self.sum_dt = 0.0
        self.delta_p, self.delta_v = np.zeros(3), np.zeros(3)
        self.delta_q = np.array([1.0, 0, 0, 0])
        acc_n, gyr_n, acc_w, gyr_w = noise
        d = [acc_n ** 2] * 3 + [gyr_n ** 2] * 3 + [acc_n ** 2] * 3 + [gyr_n ** 2] * 3 + [acc_w ** 2] * 3 + [gyr_w ** 2] * 3
        self.noise = np.diag(d)

    def push_back(self, dt, acc_1, gyr_1):
        acc_1, gyr_1 = np.array(acc_1, float), np.array(gyr_1, float)
        ba, bg, dq = self.linearized_ba, self.linearized_bg, self.delta_q
        R0 = _qrot(dq)
        un_acc_0 = R0 @ (self.acc_0 - ba)                                   # :65
        un_gyr = 0.5 * (self.gyr_0 + gyr_1) - bg                             # :66
        rq = _qmul(dq, np.array([1.0, un_gyr[0] * dt / 2, un_gyr[1] * dt / 2, un_gyr[2] * dt / 2]))  # :67
        R1 = _qrot(rq)
        un_acc_1 = R1 @ (acc_1 - ba)                                        # :68
        un_acc = 0.5 * (un_acc_0 + un_acc_1)
        rdp = self.delta_p + self.delta_v * dt + 0.5 * un_acc * dt * dt      # :70
        rdv = self.delta_v + un_acc * dt                                     # :71
        Rw, Ra0, Ra1 = _skew(un_gyr), _skew(self.acc_0 - ba), _skew(acc_1 - ba)
        I3 = np.eye(3)
        F = np.zeros((15, 15))                                               # :95-111
        F[0:3, 0:3] = I3
        F[0:3, 3:6] = -0.25 * R0 @ Ra0 * dt * dt + -0.25 * R1 @ Ra1 @ (I3 - Rw * dt) * dt * dt
        F[0:3, 6:9] = I3 * dt
        F[0:3, 9:12] = -0.25 * (R0 + R1) * dt * dt
        F[0:3, 12:15] = -0.25 * R1 @ Ra1 * dt * dt * -dt
        F[3:6, 3:6] = I3 - Rw * dt
        F[3:6, 12:15] = -1.0 * I3 * dt
        F[6:9, 3:6] = -0.5 * R0 @ Ra0 * dt + -0.5 * R1 @ Ra1 @ (I3 - Rw * dt) * dt
        F[6:9, 6:9] = I3
        F[6:9, 9:12] = -0.5 * (R0 + R1) * dt
        F[6:9, 12:15] = -0.5 * R1 @ Ra1 * dt * -dt
        F[9:12, 9:12] = I3
        F[12:15, 12:15] = I3
        V = np.zeros((15, 18))                                               # :114-126
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
        self.jacobian = F @ self.jacobian                                    # :130
        self.covariance = F @ self.covariance @ F.T + V @ self.noise @ V.T   # :131
        self.delta_p, self.delta_v = rdp, rdv                                # :146-151
        self.delta_q = rq / np.linalg.norm(rq)
        self.sum_dt += dt
        self.acc_0, self.gyr_0 = acc_1, gyr_1                                # :153-154


def process_imu(P, V, R, ba, bg, g, acc_0, gyr_0, samples):
    """Estimator::processIMU state prediction (estimator.cpp:107-116) over samples = [(dt, acc, gyr), ...]."""
    P, V, R = np.array(P, float), np.array(V, float), np.array(R, float).reshape(3, 3)
    acc_0, gyr_0 = np.array(acc_0, float), np.array(gyr_0, float)
    for dt, acc, gyr in samples:
        acc, gyr = np.array(acc, float), np.array(gyr, float)
        un_acc_0 = R @ (acc_0 - ba) - g
        un_gyr = 0.5 * (gyr_0 + gyr) - bg
        th = un_gyr * dt
        R = R @ _qrot(np.array([1.0, th[0] / 2, th[1] / 2, th[2] / 2]))      # Utility::deltaQ, utility.h:16-28
        un_acc_1 = R @ (acc - ba) - g
        un_acc = 0.5 * (un_acc_0 + un_acc_1)
        P = P + dt * V + 0.5 * dt * dt * un_acc
        V = V + dt * un_acc
        acc_0, gyr_0 = acc, gyr
    return P, V, R


def triangulate(Ps, Rs, tic, ric, start_frame, obs_first, obs_pts, depth, init_depth=5.0):
    """feature_manager.cpp:199-253.  Rs: (F, 3, 3); returns the updated depth array."""
    Ps, Rs = np.asarray(Ps, float), np.asarray(Rs, float).reshape(-1, 3, 3)
    tic, ric = np.asarray(tic, float), np.asarray(ric, float).reshape(3, 3)
    pts = np.asarray(obs_pts, float).reshape(-1, 3)
    out = np.array(depth, float, copy=True)
    window_size = len(Ps) - 1
    for l in range(len(start_frame)):
        K = obs_first[l + 1] - obs_first[l]
        if not (K >= 2 and start_frame[l] < window_size - 2):               # :203-205
            continue
        if out[l] > 0:                                                       # :207-208
            continue
        i = start_frame[l]
        t0 = Ps[i] + Rs[i] @ tic
        R0 = Rs[i] @ ric
        A = np.zeros((2 * K, 4))
        for k in range(K):
            j = i + k
            t1 = Ps[j] + Rs[j] @ tic
            R1 = Rs[j] @ ric
            t = R0.T @ (t1 - t0)
            R = R0.T @ R1
            Pm = np.zeros((3, 4))
            Pm[:, :3] = R.T
            Pm[:, 3] = -R.T @ t
            f = pts[obs_first[l] + k] / np.linalg.norm(pts[obs_first[l] + k])
            A[2 * k] = f[0] * Pm[2] - f[2] * Pm[0]                          # :234
            A[2 * k + 1] = f[1] * Pm[2] - f[2] * Pm[1]                      # :235
        Vh = np.linalg.svd(A)[2]
        X = Vh[-1]                                                           # matrixV().rightCols<1>()  (:241)
        d = (X[:3] / X[3]) @ pts[obs_first[l]]                               # :243-245
        out[l] = d if d >= 0 else init_depth                                 # :246-249
    return out
