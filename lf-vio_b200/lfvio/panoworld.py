"""Synthetic panoramic-annular image stream of a textured 3-D room, rendered through the Scaramuzza model — harness only.

SURVEY.md §8(d) config 2: "camera on a smooth trajectory inside a textured box rendered through the Scaramuzza model".  The
image formation is the inverse of the tracker's lift: every pixel of the annulus is lifted to its bearing with the camera's
polynomial (OCAMCamera::liftProjective, ScaramuzzaCamera.cc:622-645), rotated into the world with the true camera pose at the
exposure time, intersected with the axis-aligned room and shaded with the wall texture at the hit point (bilinear).  A pixel
therefore shows exactly the world point its bearing points at, for rays above and below the camera's x-y plane alike.

PixelWorld is a violoop.World whose per-frame observations are not simulated bearings but what a FeatureTracker (the device
one through the C-ABI, or the cv2-backed CPU restatement) makes of these images: id, unit bearing, bearing velocity and
pixel row of every track with track_cnt > 1, as feature_tracker_node.cpp:113-172 publishes them.
"""
from __future__ import annotations

import numpy as np

from . import violoop


def lift_sphere(cam: dict, u: np.ndarray, v: np.ndarray) -> np.ndarray:
    """Unit bearings of pixel coordinates (same arithmetic as lfvio_lift_projective with normalize = 1), shape (..., 3)."""
    xc, yc = u - cam["cx"], v - cam["cy"]
    s = 1.0 / (cam["C"] - cam["D"] * cam["E"])
    xa = s * (xc - cam["D"] * yc)
    ya = s * (-cam["E"] * xc + cam["C"] * yc)
    phi = np.sqrt(xa * xa + ya * ya)
    z, p = np.zeros_like(phi), np.ones_like(phi)
    for c in cam["poly"]:
        z = z + p * c
        p = p * phi
    b = np.stack([xa, ya, -z], axis=-1)
    return b / np.linalg.norm(b, axis=-1, keepdims=True)


def _wall_texture(rng, h_px: int, w_px: int) -> np.ndarray:
    """Corner-rich, band-limited wall paper: overlapping rectangles of random grey + two octaves of smooth noise."""
    from scipy import ndimage

    img = np.full((h_px, w_px), 110.0, np.float32)
    for _ in range(int(h_px * w_px / 900)):
        rh, rw = int(rng.integers(6, min(70, h_px // 2))), int(rng.integers(6, min(70, w_px // 2)))
        y, x = int(rng.integers(0, h_px - rh)), int(rng.integers(0, w_px - rw))
        img[y:y + rh, x:x + rw] = rng.uniform(20, 235)
    for sigma, amp in ((9.0, 28.0), (2.5, 14.0)):
        n = ndimage.gaussian_filter(rng.standard_normal((h_px, w_px)).astype(np.float32), sigma)
        img += amp * n / max(float(n.std()), 1e-6)
    return ndimage.gaussian_filter(img, 1.2).clip(0, 255).astype(np.float32)


class PanoRenderer:
    """Room [-hx, hx] x [-hy, hy] x [z0, z1] (metres) seen by a PAL camera; textures at `res` metres per texel."""

    def __init__(self, cam: dict, tcfg: dict, seed: int = 1, half=(6.0, 6.0), z=(-2.0, 3.0), res=0.012, margin_px=32, n_movers=16,
                 noise_sigma=1.5):
        from scipy import ndimage  # noqa: F401  (fail early if scipy is missing)

        self.cam, self.H, self.W = cam, cam["image_height"], cam["image_width"]
        self.lo = np.array([-half[0], -half[1], z[0]])
        self.hi = np.array([half[0], half[1], z[1]])
        self.res = res
        vv, uu = np.mgrid[0:self.H, 0:self.W].astype(np.float64)
        r = np.hypot(uu - float(int(tcfg["center_x"])), vv - float(int(tcfg["center_y"])))
        self.sel = (r <= tcfg["max_r"] + margin_px) & (r >= max(tcfg["min_r"] - margin_px, 4))
        self.bear = lift_sphere(cam, uu[self.sel], vv[self.sel])  # (n, 3) camera-frame unit bearings
        rng = np.random.default_rng(4242 + seed)
        ext = self.hi - self.lo
        # face f: fixed axis a (0 x, 1 y, 2 z), side (0 low, 1 high); texture axes are the other two
        self.tex = {}
        for a in range(3):
            o = [k for k in range(3) if k != a]
            for side in range(2):
                self.tex[(a, side)] = _wall_texture(rng, int(ext[o[0]] / res) + 2, int(ext[o[1]] / res) + 2)
        # independently moving posters on the four walls: what their tracks do is not explained by the camera's motion, so
        # rejectWithF has outliers to find (on a static, noise-free scene every RANSAC seed keeps every track)
        self.movers = []
        for _ in range(n_movers):
            a, side = int(rng.integers(0, 2)), int(rng.integers(0, 2))
            o = [k for k in range(3) if k != a]
            size = rng.uniform(0.8, 1.5)
            c0 = np.array([rng.uniform(self.lo[o[0]] + 1.0, self.hi[o[0]] - 1.0), rng.uniform(-0.8, 1.6)])
            vel = np.array([rng.choice([-1.0, 1.0]) * rng.uniform(0.8, 2.0), rng.uniform(-0.6, 0.6)])  # 0.5 deg per frame and more at 6 m
            self.movers.append((a, side, c0, vel, size, _wall_texture(rng, int(size / res) + 2, int(size / res) + 2)))
        self.noise_sigma = noise_sigma
        self.noise_rng = np.random.default_rng(777 + seed)

    def render(self, R_wc: np.ndarray, p_wc: np.ndarray, t: float = 0.0) -> np.ndarray:
        from scipy import ndimage

        d = self.bear @ R_wc.T  # world-frame ray directions
        with np.errstate(divide="ignore", invalid="ignore"):
            t_axis = np.where(d > 0, (self.hi - p_wc) / d, (self.lo - p_wc) / d)  # distance to the wall ahead, per axis
        t_axis = np.where(np.isfinite(t_axis) & (t_axis > 0), t_axis, np.inf)
        a_hit = np.argmin(t_axis, axis=1)
        t_hit = t_axis[np.arange(len(d)), a_hit]
        hit = p_wc + d * t_hit[:, None]
        val = np.zeros(len(d), np.float32)
        for a in range(3):
            o = [k for k in range(3) if k != a]
            for side in range(2):
                m = (a_hit == a) & ((d[:, a] > 0) == bool(side))
                if not m.any():
                    continue
                c0 = (hit[m, o[0]] - self.lo[o[0]]) / self.res
                c1 = (hit[m, o[1]] - self.lo[o[1]]) / self.res
                v = ndimage.map_coordinates(self.tex[(a, side)], [c0, c1], order=1, mode="nearest")
                for (ma, ms, mc0, mvel, msize, mtex) in self.movers:
                    if ma != a or ms != side:
                        continue
                    ctr = mc0 + mvel * t
                    q0 = (hit[m, o[0]] - (ctr[0] - msize / 2)) / self.res
                    q1 = (hit[m, o[1]] - (ctr[1] - msize / 2)) / self.res
                    inside = (q0 >= 0) & (q1 >= 0) & (q0 <= msize / self.res) & (q1 <= msize / self.res)
                    if inside.any():
                        v[inside] = ndimage.map_coordinates(mtex, [q0[inside], q1[inside]], order=1, mode="nearest")
                val[m] = v
        if self.noise_sigma > 0:
            val = val + self.noise_rng.normal(0.0, self.noise_sigma, len(val)).astype(np.float32)
        img = np.zeros((self.H, self.W), np.uint8)
        img[self.sel] = np.clip(np.rint(val), 0, 255).astype(np.uint8)
        return img


class PixelWorld(violoop.World):
    """violoop.World (trajectory, 200 Hz IMU, ground truth) + rendered frames; `track()` fills `frames` from a tracker."""

    def __init__(self, cam: dict, tcfg: dict, seed=1, n_frames=45, frame_dt=0.1, render_seed=1):
        super().__init__(seed=seed, n_frames=n_frames, max_feat=0, frame_dt=frame_dt, pixel_noise=0.0)
        self.cam, self.tcfg = cam, tcfg
        self.renderer = PanoRenderer(cam, tcfg, seed=render_seed)
        self.images = []
        for t in self.times:
            ts = t + self.td_true  # the image is exposed td later than its stamp
            R_wb, p_wb = self.traj.rot(ts), self.traj.pos(ts)
            self.images.append(self.renderer.render(R_wb @ self.ric, p_wb + R_wb @ self.tic, ts - self.times[0]))

    def track(self, tracker) -> "PixelWorld":
        """tracker: object with read_image(img, t, pub) / update_ids() and the FeatureTracker members ids, track_cnt, cur_pts,
        cur_un_pts_3d (or bearing) and pts_velocity_3d.  Every frame is published (10 Hz camera = the reference's publish rate)."""
        self.frames, self.n_tracked = [], []
        for k, t in enumerate(self.times):
            tracker.read_image(self.images[k], float(t), True)
            tracker.update_ids()
            ids, cnt = np.asarray(tracker.ids), np.asarray(tracker.track_cnt)
            pts = np.asarray(tracker.cur_pts, np.float64).reshape(-1, 2)
            b = np.asarray(tracker.cur_un_pts_3d, np.float32).reshape(-1, 3).astype(np.float64)
            vel = np.asarray(tracker.pts_velocity_3d, np.float32).reshape(-1, 3).astype(np.float64)
            obs = {}
            for i in range(len(ids)):
                if cnt[i] > 1:  # feature_tracker_node.cpp:142
                    obs[int(ids[i])] = (b[i], vel[i], float(pts[i, 1]))
            self.frames.append(obs)
            self.n_tracked.append(len(obs))
        return self
