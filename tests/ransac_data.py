"""Synthetic two-view data for the rejectWithF tests: annulus pixels -> bearings -> rigid motion -> pixels."""
import numpy as np

from oracle import camera as ocam


def rot(rx, ry, rz):
    cx, sx, cy, sy, cz, sz = np.cos(rx), np.sin(rx), np.cos(ry), np.sin(ry), np.cos(rz), np.sin(rz)
    Rx = np.array([[1, 0, 0], [0, cx, -sx], [0, sx, cx]])
    Ry = np.array([[cy, 0, sy], [0, 1, 0], [-sy, 0, cy]])
    Rz = np.array([[cz, -sz, 0], [sz, cz, 0], [0, 0, 1]])
    return Rz @ Ry @ Rx


def two_view(cam, cfg, n, seed, outlier_frac=0.25, noise_px=0.15, t=(0.08, -0.03, 0.02), rpy=(0.01, -0.015, 0.03)):
    """Returns (cur_pts f32 n x 2, forw_pts f32 n x 2, is_outlier bool n)."""
    rng = np.random.default_rng(seed)
    r = rng.uniform(cfg["min_r"] + 15, cfg["max_r"] - 15, n)
    a = rng.uniform(0, 2 * np.pi, n)
    cur = np.stack([cfg["center_x"] + r * np.cos(a), cfg["center_y"] + r * np.sin(a)], 1).astype(np.float32)
    b1 = ocam.lift_sphere(cam, cur.astype(np.float64))
    depth = rng.uniform(1.5, 12.0, n)
    P2 = (rot(*rpy) @ (b1 * depth[:, None]).T).T + np.asarray(t)
    forw = ocam.space_to_plane(cam, P2) + rng.normal(0, noise_px, (n, 2))
    out = rng.random(n) < outlier_frac
    forw[out] += rng.uniform(-25, 25, (int(out.sum()), 2)) + np.sign(rng.normal(size=(int(out.sum()), 2))) * 6
    return cur, forw.astype(np.float32), out
