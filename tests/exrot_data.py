"""Synthetic rotation-rich two-view stream for the extrinsic-rotation calibration tests (InitialEXRotation)."""
import numpy as np

from lfvio import synthwin
from oracle.initial import rot_to_quat


def _rodrigues(axis, ang):
    a = np.asarray(axis, float) / np.linalg.norm(axis)
    K = np.array([[0, -a[2], a[1]], [a[2], 0, -a[0]], [-a[1], a[0], 0]])
    return np.eye(3) + np.sin(ang) * K + (1 - np.cos(ang)) * K @ K


def make_stream(seed=0, n_frames=13, n_pts=120, angle=0.3, gyro_noise=1e-3):
    """-> (ric_true, [(corres (n,6), delta_q_imu (w,x,y,z))] for frames 1 .. n_frames-1).  corres = [bearing in frame k-1 |
    bearing in frame k]; X_{k-1} = R X_k + t with R = R_{c_{k-1} <- c_k}; the IMU rotation is ric R ric^T plus a small error."""
    rng = np.random.default_rng(seed)
    ric = synthwin.rot_xyz(0.7, -0.4, 1.1)
    out = []
    for _ in range(1, n_frames):
        Rk = _rodrigues(rng.normal(size=3), angle * rng.uniform(0.7, 1.3))
        t = rng.normal(0, 0.25, 3)
        X1 = rng.normal(size=(n_pts, 3))
        X1 = X1 / np.linalg.norm(X1, axis=1, keepdims=True) * rng.uniform(3, 8, (n_pts, 1))      # points in frame k, all around (PAL)
        X0 = X1 @ Rk.T + t
        b0, b1 = X0 / np.linalg.norm(X0, axis=1, keepdims=True), X1 / np.linalg.norm(X1, axis=1, keepdims=True)
        Rimu = ric @ Rk @ ric.T @ _rodrigues(rng.normal(size=3), gyro_noise)
        out.append((np.hstack([b0, b1]), rot_to_quat(Rimu)))
    return ric, out
