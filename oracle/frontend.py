"""CPU oracle of the panoramic front end — TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Two layers:

1. numpy restatements of the third-party image operators the reference calls.  Their source is
   OpenCV (NOT in /root/reference; effective pin ros:kinetic-perception => OpenCV 3.3.x,
   docker/Dockerfile:1).  The oracle of record is the in-container Python `cv2` 4.13.0 called with the
   reference's exact arguments; the restatements below state the published algorithms
    and are pinned against cv2 by tests/test_oracle_frontend.py:
     clahe()            <- cv::createCLAHE(3.0, Size(8,8))->apply       call site feature_tracker.cpp:103-105
     pyr_down()/scharr  <- inside cv::calcOpticalFlowPyrLK              call site feature_tracker.cpp:127
     lk_track()         <- cv::calcOpticalFlowPyrLK(.., Size(41,41), 3) call site feature_tracker.cpp:127
     min_eig_map()/good_features_to_track() <- cv::goodFeaturesToTrack  call site feature_tracker.cpp:166
                           (in-tree spec: pointcloud_image_fusion/src/pointcloud_image_fusion.cpp:210-396)
     circle_spans()/fill_circle() <- cv::circle(.., -1)                 call sites feature_tracker.cpp:55-56,80
2. FeatureTrackerOracle: line-by-line restatement of the reference's own control flow
   (feature_tracker/src/feature_tracker.cpp:46-184,388-398,441-504 and the updateID loop of
   feature_tracker_node.cpp:103-111) on top of either layer ("cv2" or "restated" backend).

Parity status: the reference has no tests for this path => pinned by golden vectors generated from
cv2 on seeded synthetic frames (tests/golden/, generator tests/golden/make_frontend_golden.py).
"""
from __future__ import annotations

import ctypes
import os

import numpy as np

f32 = np.float32
f64 = np.float64

WIN = 41  # LK window (feature_tracker.cpp:127)
LK_MAX_LEVEL = 3
W_BITS = 14


def _fma32(a, b, c):
    """f32 fused multiply-add emulated through f64 (product of two f32 is exact in f64)."""
    return (np.asarray(a, f64) * np.asarray(b, f64) + np.asarray(c, f64)).astype(f32)


def cv_round(x):
    """cvRound == lrint == round-half-to-even."""
    return np.rint(x).astype(np.int64)


def reflect101(idx, n):
    idx = np.asarray(idx)
    if n == 1:
        return np.zeros_like(idx)
    period = 2 * (n - 1)
    m = np.mod(idx, period)
    return np.where(m >= n, period - m, m)


# ------------------------------------------------------------------------------------------------
# CLAHE
# ------------------------------------------------------------------------------------------------
def clahe(img: np.ndarray, clip_limit: float = 3.0, tiles: int = 8) -> np.ndarray:
    H, W = img.shape
    if W % tiles == 0 and H % tiles == 0:
        ext = img
    else:
        pad_b, pad_r = tiles - H % tiles, tiles - W % tiles
        yy = reflect101(np.arange(H + pad_b), H)
        xx = reflect101(np.arange(W + pad_r), W)
        ext = img[np.ix_(yy, xx)]
    th, tw = ext.shape[0] // tiles, ext.shape[1] // tiles
    area = th * tw
    lut_scale = f32(255.0) / f32(area)
    clip = max(int(clip_limit * area / 256), 1)
    luts = np.zeros((tiles, tiles, 256), np.uint8)
    for ty in range(tiles):
        for tx in range(tiles):
            hist = np.bincount(ext[ty * th:(ty + 1) * th, tx * tw:(tx + 1) * tw].ravel(), minlength=256).astype(np.int64)
            over = np.maximum(hist - clip, 0)
            clipped = int(over.sum())
            hist = np.minimum(hist, clip)
            batch = clipped // 256
            residual = clipped - batch * 256
            hist += batch
            if residual != 0:
                step = max(256 // residual, 1)
                i = 0
                while i < 256 and residual > 0:
                    hist[i] += 1
                    i += step
                    residual -= 1
            cs = np.cumsum(hist)
            luts[ty, tx] = np.clip(cv_round(cs.astype(f32) * lut_scale), 0, 255).astype(np.uint8)
    inv_tw, inv_th = f32(1.0) / f32(tw), f32(1.0) / f32(th)
    xs = np.arange(W).astype(f32) * inv_tw - f32(0.5)
    tx1 = np.floor(xs).astype(np.int64)
    xa = (xs - tx1.astype(f32)).astype(f32)
    xa1 = (f32(1.0) - xa).astype(f32)
    tx2 = np.minimum(tx1 + 1, tiles - 1)
    tx1 = np.maximum(tx1, 0)
    ys = np.arange(H).astype(f32) * inv_th - f32(0.5)
    ty1 = np.floor(ys).astype(np.int64)
    ya = (ys - ty1.astype(f32)).astype(f32)
    ya1 = (f32(1.0) - ya).astype(f32)
    ty2 = np.minimum(ty1 + 1, tiles - 1)
    ty1 = np.maximum(ty1, 0)
    v = img.astype(np.int64)
    l11 = luts[ty1[:, None], tx1[None, :], v].astype(f32)
    l12 = luts[ty1[:, None], tx2[None, :], v].astype(f32)
    l21 = luts[ty2[:, None], tx1[None, :], v].astype(f32)
    l22 = luts[ty2[:, None], tx2[None, :], v].astype(f32)
    top = ((l11 * xa1[None, :]).astype(f32) + (l12 * xa[None, :]).astype(f32)).astype(f32)
    bot = ((l21 * xa1[None, :]).astype(f32) + (l22 * xa[None, :]).astype(f32)).astype(f32)
    res = ((top * ya1[:, None]).astype(f32) + (bot * ya[:, None]).astype(f32)).astype(f32)
    return np.clip(cv_round(res), 0, 255).astype(np.uint8)


# ------------------------------------------------------------------------------------------------
# Pyramid + Scharr
# ------------------------------------------------------------------------------------------------
def pyr_down(img: np.ndarray) -> np.ndarray:
    H, W = img.shape
    Ho, Wo = (H + 1) // 2, (W + 1) // 2
    k = np.array([1, 4, 6, 4, 1], np.int64)
    src = img.astype(np.int64)
    xs = reflect101(2 * np.arange(Wo)[:, None] + np.arange(-2, 3)[None, :], W)  # Wo x 5
    rows = (src[:, xs] * k[None, None, :]).sum(axis=2)  # H x Wo
    ys = reflect101(2 * np.arange(Ho)[:, None] + np.arange(-2, 3)[None, :], H)  # Ho x 5
    out = (rows[ys, :] * k[None, :, None]).sum(axis=1)  # Ho x Wo
    return ((out + 128) >> 8).astype(np.uint8)


def build_pyramid(img: np.ndarray, max_level: int = LK_MAX_LEVEL, win: int = WIN):
    pyr = [img]
    for _ in range(max_level):
        nxt = pyr_down(pyr[-1])
        if nxt.shape[1] <= win or nxt.shape[0] <= win:
            break
        pyr.append(nxt)
    return pyr


def scharr(img: np.ndarray):
    """int16 (Ix, Iy) as calcScharrDeriv: REFLECT_101 inside the image."""
    H, W = img.shape
    s = img.astype(np.int32)
    ym = reflect101(np.arange(H) - 1, H)
    yp = reflect101(np.arange(H) + 1, H)
    xm = reflect101(np.arange(W) - 1, W)
    xp = reflect101(np.arange(W) + 1, W)
    t0 = (s[ym] + s[yp]) * 3 + s * 10
    t1 = s[yp] - s[ym]
    ix = t0[:, xp] - t0[:, xm]
    iy = (t1[:, xp] + t1[:, xm]) * 3 + t1 * 10
    return ix.astype(np.int16), iy.astype(np.int16)


# ------------------------------------------------------------------------------------------------
# Pyramidal LK
# ------------------------------------------------------------------------------------------------
def _descale(x, n):
    return (x + (1 << (n - 1))) >> n


def _weights(a, b):
    one = f32(1.0)
    sc = f32(1 << W_BITS)
    iw00 = int(cv_round(((one - a) * (one - b)).astype(f32) * sc))
    iw01 = int(cv_round((a * (one - b)).astype(f32) * sc))
    iw10 = int(cv_round(((one - a) * b).astype(f32) * sc))
    iw11 = (1 << W_BITS) - iw00 - iw01 - iw10
    return iw00, iw01, iw10, iw11


def _window(img_pad, ix, iy, pad, n):
    """n x n block whose top-left is pixel (ix, iy) of the un-padded image."""
    return img_pad[iy + pad: iy + pad + n, ix + pad: ix + pad + n]


def lk_track(prev: np.ndarray, nxt: np.ndarray, pts: np.ndarray, win: int = WIN, max_level: int = LK_MAX_LEVEL,
             max_iter: int = 30, eps: float = 0.01, min_eig_thr: float = 1e-4):
    """Returns (next_pts f32 Nx2, status u8 N).  f32 accumulation in row-major pixel order (the scalar
    path of OpenCV's LKTrackerInvoker; its SIMD path sums in a different order => last-bit noise)."""
    pts = np.asarray(pts, f32).reshape(-1, 2)
    N = pts.shape[0]
    pyr_i = build_pyramid(prev, max_level, win)
    pyr_j = build_pyramid(nxt, max_level, win)
    levels = min(len(pyr_i), len(pyr_j)) - 1
    pad = win
    status = np.ones(N, np.uint8)
    next_pts = np.zeros((N, 2), f32)
    half = f32((win - 1) * 0.5)
    FLT_SCALE = f32(1.0 / (1 << 20))
    eps2 = eps * eps
    for level in range(levels, -1, -1):
        I, J = pyr_i[level], pyr_j[level]
        rows, cols = I.shape
        dxi, dyi = scharr(I)
        yy = reflect101(np.arange(-pad, rows + pad), rows)
        xx = reflect101(np.arange(-pad, cols + pad), cols)
        Ip = I[np.ix_(yy, xx)].astype(np.int64)
        Jp = J[np.ix_(yy, xx)].astype(np.int64)
        dxp = np.zeros((rows + 2 * pad, cols + 2 * pad), np.int64)
        dyp = np.zeros_like(dxp)
        dxp[pad:pad + rows, pad:pad + cols] = dxi
        dyp[pad:pad + rows, pad:pad + cols] = dyi
        scale = f32(1.0 / (1 << level))
        for p in range(N):
            prev_pt = (pts[p] * scale).astype(f32)
            if level == levels:
                next_pt = prev_pt.copy()
            else:
                next_pt = (next_pts[p] * f32(2.0)).astype(f32)
            next_pts[p] = next_pt
            pp = (prev_pt - half).astype(f32)
            ipx, ipy = int(np.floor(pp[0])), int(np.floor(pp[1]))
            if ipx < -win or ipx >= cols or ipy < -win or ipy >= rows:
                if level == 0:
                    status[p] = 0
                continue
            a = f32(pp[0] - f32(ipx))
            b = f32(pp[1] - f32(ipy))
            w00, w01, w10, w11 = _weights(a, b)
            Iw = _window(Ip, ipx, ipy, pad, win + 1)
            Dx = _window(dxp, ipx, ipy, pad, win + 1)
            Dy = _window(dyp, ipx, ipy, pad, win + 1)

            def bil(M, sh):
                return _descale(M[:-1, :-1] * w00 + M[:-1, 1:] * w01 + M[1:, :-1] * w10 + M[1:, 1:] * w11, sh)

            Iwin = bil(Iw, W_BITS - 5)
            Ix = bil(Dx, W_BITS)
            Iy = bil(Dy, W_BITS)
            A11 = f32(np.cumsum((Ix * Ix).astype(f32).ravel(), dtype=f32)[-1]) * FLT_SCALE
            A12 = f32(np.cumsum((Ix * Iy).astype(f32).ravel(), dtype=f32)[-1]) * FLT_SCALE
            A22 = f32(np.cumsum((Iy * Iy).astype(f32).ravel(), dtype=f32)[-1]) * FLT_SCALE
            D = f32(f32(A11 * A22) - f32(A12 * A12))
            dif = f32(A11 - A22)
            min_eig = f32(f32(f32(A22 + A11) - f32(np.sqrt(f32(f32(dif * dif) + f32(f32(f32(4.0) * A12) * A12))))) / f32(2 * win * win))
            if min_eig < min_eig_thr or D < np.finfo(f32).eps:
                if level == 0:
                    status[p] = 0
                continue
            D = f32(f32(1.0) / D)
            next_pt = (next_pt - half).astype(f32)
            prev_delta = np.zeros(2, f32)
            for j in range(max_iter):
                inx, iny = int(np.floor(next_pt[0])), int(np.floor(next_pt[1]))
                if inx < -win or inx >= cols or iny < -win or iny >= rows:
                    if level == 0:
                        status[p] = 0
                    break
                a = f32(next_pt[0] - f32(inx))
                b = f32(next_pt[1] - f32(iny))
                w00, w01, w10, w11 = _weights(a, b)
                Jw = _window(Jp, inx, iny, pad, win + 1)
                diff = _descale(Jw[:-1, :-1] * w00 + Jw[:-1, 1:] * w01 + Jw[1:, :-1] * w10 + Jw[1:, 1:] * w11, W_BITS - 5) - Iwin
                b1 = f32(np.cumsum((diff * Ix).astype(f32).ravel(), dtype=f32)[-1]) * FLT_SCALE
                b2 = f32(np.cumsum((diff * Iy).astype(f32).ravel(), dtype=f32)[-1]) * FLT_SCALE
                delta = np.array([f32(f32(f32(A12 * b2) - f32(A22 * b1)) * D), f32(f32(f32(A12 * b1) - f32(A11 * b2)) * D)], f32)
                next_pt = (next_pt + delta).astype(f32)
                next_pts[p] = (next_pt + half).astype(f32)
                if float(delta[0]) * float(delta[0]) + float(delta[1]) * float(delta[1]) <= eps2:
                    break
                if j > 0 and abs(f32(delta[0] + prev_delta[0])) < 0.01 and abs(f32(delta[1] + prev_delta[1])) < 0.01:
                    next_pts[p] = (next_pts[p] - (delta * f32(0.5)).astype(f32)).astype(f32)
                    break
                prev_delta = delta
            if status[p] and level == 0:
                # the caller requested `err`: final window must be in bounds or status is cleared
                q = (next_pts[p] - half).astype(f32)
                qx, qy = int(np.floor(q[0])), int(np.floor(q[1]))
                if qx < -win or qx >= cols or qy < -win or qy >= rows:
                    status[p] = 0
    return next_pts, status


# ------------------------------------------------------------------------------------------------
# Shi-Tomasi.  Rounding sequence matches the cv2 4.13 AVX2 build bit-for-bit.
# ------------------------------------------------------------------------------------------------
def min_eig_map(img: np.ndarray) -> np.ndarray:
    H, W = img.shape
    yy = reflect101(np.arange(-1, H + 1), H)
    xx = reflect101(np.arange(-1, W + 1), W)
    I = img[np.ix_(yy, xx)].astype(f32)
    s = f32(1.0 / (4 * 3 * 255.0))
    s2 = f32(2.0) * s
    # dx: rows [-1 0 1] exact, columns symmetric [s 2s s]: fma(S0+S2, s, S1*2s)
    rdx = (I[:, 2:] - I[:, :-2]).astype(f32)
    dx = _fma32((rdx[:-2] + rdx[2:]).astype(f32), s, (rdx[1:-1] * s2).astype(f32))
    # dy: rows [s 2s s] as fma(C, s, fma(B, 2s, A*s)), columns [-1 0 1]
    A, B, C = I[:, :-2], I[:, 1:-1], I[:, 2:]
    rdy = _fma32(C, s, _fma32(B, s2, (A * s).astype(f32)))
    dy = (rdy[2:] - rdy[:-2]).astype(f32)

    def box(a):
        p = a[np.ix_(reflect101(np.arange(-1, H + 1), H), reflect101(np.arange(-1, W + 1), W))].astype(f64)
        r = p[:, :-2] + p[:, 1:-1] + p[:, 2:]
        return (r[:-2] + r[1:-1] + r[2:]).astype(f32)

    a = (box((dx * dx).astype(f32)) * f32(0.5)).astype(f32)
    b = box((dx * dy).astype(f32))
    c = (box((dy * dy).astype(f32)) * f32(0.5)).astype(f32)
    t = (a - c).astype(f32)
    return ((a + c).astype(f32) - np.sqrt(((t * t).astype(f32) + (b * b).astype(f32)).astype(f32))).astype(f32)


def corner_candidates(eig: np.ndarray, mask: np.ndarray | None, quality: float):
    """(ys, xs) of the NMS survivors in descending (value, linear index) order — cv2 4.x total order."""
    H, W = eig.shape
    valid = np.ones((H, W), bool) if mask is None else (mask != 0)
    if not valid.any():
        return np.zeros(0, np.int64), np.zeros(0, np.int64)
    max_val = eig[valid].max()  # minMaxLoc is done in f64 on f32 data => exact
    thr = f32(f64(max_val) * quality)  # cv::threshold(eig, eig, maxVal*qualityLevel, 0, THRESH_TOZERO) (f64 thresh cast to f32)
    e = np.where(eig > thr, eig, f32(0))
    p = np.full((H + 2, W + 2), -np.inf, f32)
    p[1:-1, 1:-1] = e
    dil = np.max(np.stack([p[dy:dy + H, dx:dx + W] for dy in range(3) for dx in range(3)]), axis=0)
    keep = (e != 0) & (e == dil) & valid
    keep[0, :] = keep[-1, :] = False
    keep[:, 0] = keep[:, -1] = False
    ys, xs = np.nonzero(keep)
    vals = e[ys, xs]
    lin = ys * W + xs
    order = np.lexsort((-lin, -vals.astype(f64)))
    return ys[order], xs[order]


def select_min_distance(ys, xs, shape, max_corners: int, min_distance: float):
    """Greedy acceptance in the given order; reject if an accepted corner lies strictly within min_distance."""
    H, W = shape
    out = []
    if min_distance >= 1:
        cell = int(cv_round(min_distance))
        gw, gh = (W + cell - 1) // cell, (H + cell - 1) // cell
        grid = [[] for _ in range(gw * gh)]
        md2 = min_distance * min_distance
        for y, x in zip(ys.tolist(), xs.tolist()):
            xc, yc = x // cell, y // cell
            good = True
            for yy in range(max(0, yc - 1), min(gh - 1, yc + 1) + 1):
                for xx in range(max(0, xc - 1), min(gw - 1, xc + 1) + 1):
                    for (px, py) in grid[yy * gw + xx]:
                        dx, dy = x - px, y - py
                        if dx * dx + dy * dy < md2:
                            good = False
                            break
                    if not good:
                        break
                if not good:
                    break
            if good:
                grid[yc * gw + xc].append((x, y))
                out.append((x, y))
                if 0 < max_corners <= len(out):
                    break
    else:
        for y, x in zip(ys.tolist(), xs.tolist()):
            out.append((x, y))
            if 0 < max_corners <= len(out):
                break
    return np.array(out, f32).reshape(-1, 2)


def good_features_to_track(img, max_corners, quality, min_distance, mask=None):
    eig = min_eig_map(img)
    ys, xs = corner_candidates(eig, mask, quality)
    return select_min_distance(ys, xs, img.shape, max_corners, min_distance)


# ------------------------------------------------------------------------------------------------
# cv::circle(img, center, radius, color, -1)  (filled, no shift): midpoint spans
# ------------------------------------------------------------------------------------------------
def circle_spans(radius: int) -> np.ndarray:
    """half_width[|dy|] for dy in [0, radius]; row cy+dy is filled on [cx-hw, cx+hw]."""
    hw = np.full(radius + 1, -1, np.int64)
    err, dx, dy, plus, minus = 0, radius, 0, 1, (radius << 1) - 1
    while dx >= dy:
        hw[dy] = max(hw[dy], dx)
        hw[dx] = max(hw[dx], dy)
        dy += 1
        err += plus
        plus += 2
        m = (1 if err <= 0 else 0) - 1  # 0 or -1
        err -= minus & m
        dx += m
        minus -= m & 2
    return hw


def fill_circle(img: np.ndarray, cx: int, cy: int, radius: int, value: int) -> None:
    H, W = img.shape
    hw = circle_spans(radius)
    for d in range(-radius, radius + 1):
        y = cy + d
        if 0 <= y < H:
            h = hw[abs(d)]
            x0, x1 = max(cx - h, 0), min(cx + h, W - 1)
            if x0 <= x1:
                img[y, x0:x1 + 1] = value


# ------------------------------------------------------------------------------------------------
# libstdc++ std::sort permutation (unstable; feature_tracker.cpp:66-67)
# ------------------------------------------------------------------------------------------------
_STDSORT = None


def std_sort_desc_perm(keys: np.ndarray) -> np.ndarray:
    """Permutation produced by libstdc++ std::sort with comparator a.first > b.first on keys
    (compiled from oracle/stdsort.cpp into oracle/_build/libstdsort.so)."""
    global _STDSORT
    if _STDSORT is None:
        path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_build", "libstdsort.so")
        _STDSORT = ctypes.CDLL(path)
        _STDSORT.sort_desc_perm.argtypes = [ctypes.POINTER(ctypes.c_int), ctypes.POINTER(ctypes.c_int), ctypes.c_int]
    k = np.ascontiguousarray(keys, np.int32)
    perm = np.empty(k.shape[0], np.int32)
    _STDSORT.sort_desc_perm(k.ctypes.data_as(ctypes.POINTER(ctypes.c_int)), perm.ctypes.data_as(ctypes.POINTER(ctypes.c_int)), k.shape[0])
    return perm.astype(np.int64)


# ------------------------------------------------------------------------------------------------
# FeatureTracker control flow
# ------------------------------------------------------------------------------------------------
class FeatureTrackerOracle:
    """Restatement of class FeatureTracker (feature_tracker/src/feature_tracker.h:34-80).

    backend "cv2": image operators are the cv2 calls the reference makes (the oracle of record and the
    CPU baseline that is timed).  backend "restated": the numpy restatements above.
    RANSAC (rejectWithF, feature_tracker.cpp:329-386) is seeded from std::random_device in the reference; it runs
    when a `ransac` callable is injected (oracle.ransac.make_injected(seed) is the restatement with the injected
    random source the CUDA path shares).
    """

    def __init__(self, cam: dict, cfg: dict, backend: str = "cv2", ransac=None):
        from . import camera as ocam  # local import keeps module import light

        self._ocam = ocam
        self.cam, self.cfg, self.backend, self.ransac = cam, cfg, backend, ransac
        self.ROW, self.COL = cam["image_height"], cam["image_width"]
        self.n_id = 0
        self.frames_seen = 0
        self.prev_img = self.cur_img = self.forw_img = None
        self.cur_pts = np.zeros((0, 2), f32)
        self.forw_pts = np.zeros((0, 2), f32)
        self.ids = np.zeros(0, np.int64)
        self.track_cnt = np.zeros(0, np.int64)
        self.cur_un_pts_3d = np.zeros((0, 3), f32)
        self.pts_velocity_3d = np.zeros((0, 3), f32)
        self.prev_un_pts_map_3d: dict = {}
        self.mask = None
        self.cur_time = self.prev_time = 0.0
        self.n_pts = np.zeros((0, 2), f32)
        if backend == "cv2":
            import cv2

            self._cv2 = cv2
            self._clahe = cv2.createCLAHE(3.0, (8, 8))

    # -- image operators --------------------------------------------------------------------
    def _equalize(self, img):
        return self._clahe.apply(img) if self.backend == "cv2" else clahe(img)

    def _lk(self, a, b, pts):
        if self.backend == "cv2":
            nxt, st, _ = self._cv2.calcOpticalFlowPyrLK(a, b, pts.reshape(-1, 1, 2).astype(f32), None, winSize=(WIN, WIN), maxLevel=LK_MAX_LEVEL)
            return nxt.reshape(-1, 2), st.reshape(-1).astype(np.uint8)
        return lk_track(a, b, pts)

    def _gftt(self, img, n, mask):
        if self.backend == "cv2":
            c = self._cv2.goodFeaturesToTrack(img, n, 0.01, float(self.cfg["min_dist"]), mask=mask)
            return np.zeros((0, 2), f32) if c is None else c.reshape(-1, 2).astype(f32)
        return good_features_to_track(img, n, 0.01, float(self.cfg["min_dist"]), mask)

    def _circle(self, img, cx, cy, r, val):
        if self.backend == "cv2":
            self._cv2.circle(img, (int(cx), int(cy)), int(r), int(val), -1)
        else:
            fill_circle(img, int(cx), int(cy), int(r), int(val))

    # -- feature_tracker.cpp:6-12 -----------------------------------------------------------------
    def _in_border(self, pts):
        x = cv_round(pts[:, 0])
        y = cv_round(pts[:, 1])
        return (1 <= x) & (x < self.COL - 1) & (1 <= y) & (y < self.ROW - 1)

    # -- feature_tracker.cpp:46-83 ----------------------------------------------------------------
    def set_mask(self):
        c = self.cfg
        mask = np.zeros((self.ROW, self.COL), np.uint8)
        self._circle(mask, int(c["center_x"]), int(c["center_y"]), c["max_r"], 255)  # cv::Point(double,double) truncates
        self._circle(mask, int(c["center_x"]), int(c["center_y"]), c["min_r"], 0)
        perm = std_sort_desc_perm(self.track_cnt)
        pts, ids, cnt = [], [], []
        for k in perm:
            p = self.forw_pts[k]
            px, py = int(cv_round(p[0])), int(cv_round(p[1]))  # Mat::at(Point2f -> Point) rounds
            if mask[py, px] == 255:
                pts.append(p)
                ids.append(self.ids[k])
                cnt.append(self.track_cnt[k])
                self._circle(mask, px, py, c["min_dist"], 0)
        self.forw_pts = np.array(pts, f32).reshape(-1, 2)
        self.ids = np.array(ids, np.int64)
        self.track_cnt = np.array(cnt, np.int64)
        self.mask = mask

    # -- feature_tracker.cpp:95-184 ---------------------------------------------------------------
    def read_image(self, img: np.ndarray, cur_time: float, pub_this_frame: bool = True):
        self.cur_time = cur_time
        if self.cfg.get("equalize", 1):
            img = self._equalize(img)
        if self.forw_img is None:
            self.prev_img = self.cur_img = self.forw_img = img
        else:
            self.forw_img = img
        self.forw_pts = np.zeros((0, 2), f32)
        if len(self.cur_pts) > 0:
            forw, status = self._lk(self.cur_img, self.forw_img, self.cur_pts)
            status = status.astype(bool) & self._in_border(forw)
            self.cur_pts = self.cur_pts[status]
            self.forw_pts = forw[status]
            self.ids = self.ids[status]
            self.cur_un_pts_3d = self.cur_un_pts_3d[status]
            self.track_cnt = self.track_cnt[status]
        self.track_cnt = self.track_cnt + 1
        if pub_this_frame:
            if self.ransac is not None and len(self.forw_pts) >= 8:
                keep = self.ransac(self, self.cur_pts, self.forw_pts).astype(bool)
                self.cur_pts = self.cur_pts[keep]
                self.forw_pts = self.forw_pts[keep]
                self.ids = self.ids[keep]
                self.cur_un_pts_3d = self.cur_un_pts_3d[keep]
                self.track_cnt = self.track_cnt[keep]
            self.set_mask()
            n_max = self.cfg["max_cnt"] - len(self.forw_pts)
            if n_max > 0:
                self.n_pts = self._gftt(self.forw_img, n_max, self.mask)
            else:
                self.n_pts = np.zeros((0, 2), f32)
            # addPoints (:85-93)
            self.forw_pts = np.concatenate([self.forw_pts, self.n_pts.astype(f32)]).astype(f32)
            self.ids = np.concatenate([self.ids, -np.ones(len(self.n_pts), np.int64)])
            self.track_cnt = np.concatenate([self.track_cnt, np.ones(len(self.n_pts), np.int64)])
        self.prev_img = self.cur_img
        self.cur_img = self.forw_img
        self.cur_pts = self.forw_pts
        self.undistorted_points()
        self.prev_time = self.cur_time
        self.frames_seen += 1

    # -- feature_tracker.cpp:441-504 --------------------------------------------------------------
    def undistorted_points(self):
        b = self._ocam.lift_sphere(self.cam, self.cur_pts.astype(f64))  # normalise in f64 ...
        self.cur_un_pts_3d = b.astype(f32)  # ... then cast (cv::Point3f)
        cur_map = {}
        for i, idv in enumerate(self.ids.tolist()):
            cur_map.setdefault(idv, self.cur_un_pts_3d[i])  # std::map::insert keeps the first entry per key
        n = len(self.cur_pts)
        vel = np.zeros((n, 3), f32)
        if self.prev_un_pts_map_3d:
            dt = self.cur_time - self.prev_time
            for i, idv in enumerate(self.ids.tolist()):
                if idv != -1 and idv in self.prev_un_pts_map_3d:
                    d = (self.cur_un_pts_3d[i] - self.prev_un_pts_map_3d[idv]).astype(f32)  # float - float
                    vel[i] = (d.astype(f64) / dt).astype(f32)
        self.pts_velocity_3d = vel
        self.prev_un_pts_map_3d = cur_map

    # -- feature_tracker.cpp:388-398 + feature_tracker_node.cpp:103-111 ---------------------------------
    def update_ids(self):
        for i in range(len(self.ids)):
            if self.ids[i] == -1:
                self.ids[i] = self.n_id
                self.n_id += 1

    def snapshot(self) -> dict:
        return {
            "cur_pts": self.cur_pts.copy(), "ids": self.ids.copy(), "track_cnt": self.track_cnt.copy(),
            "cur_un_pts_3d": self.cur_un_pts_3d.copy(), "pts_velocity_3d": self.pts_velocity_3d.copy(),
        }
