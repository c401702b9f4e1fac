"""Deterministic synthetic panoramic frames (numpy integer arithmetic only => bit-reproducible on any box).

A fixed two-octave bilinear-noise texture is viewed through a slowly rotating + translating window, i.e.
frame k = texture sampled (fixed-point bilinear, 1/32 px) at R_k (x - c) + c + t_k.  This is the input for the
tracker parity tests, the golden fixtures and bench.py ("synthetic 1280x960 panoramic" of BASELINE.json).
No trigonometry: the rotation uses the rational parametrisation so that coordinates are plain IEEE +,*,/.
"""
from __future__ import annotations

import zlib

import numpy as np


def _bilinear_upsample_u16(low: np.ndarray, f: int, h: int, w: int) -> np.ndarray:
    """low: (h//f + 2, w//f + 2) u8 grid -> (h, w) values scaled by f*f (uint32), integer bilinear."""
    ys, xs = np.arange(h), np.arange(w)
    y0, fy = ys // f, (ys % f).astype(np.uint32)
    x0, fx = xs // f, (xs % f).astype(np.uint32)
    L = low.astype(np.uint32)
    a = L[y0][:, x0]
    b = L[y0][:, x0 + 1]
    c = L[y0 + 1][:, x0]
    d = L[y0 + 1][:, x0 + 1]
    FX, FY = fx[None, :], fy[:, None]
    return a * (f - FX) * (f - FY) + b * FX * (f - FY) + c * (f - FX) * FY + d * FX * FY


def make_texture(height: int, width: int, seed: int = 0) -> np.ndarray:
    """u8 texture: octave 1 (cell 16 px) * 3/4 + octave 2 (cell 6 px) * 1/4."""
    rng = np.random.default_rng(seed)
    l1 = rng.integers(0, 256, (height // 16 + 2, width // 16 + 2), dtype=np.uint8)
    l2 = rng.integers(0, 256, (height // 6 + 2, width // 6 + 2), dtype=np.uint8)
    o1 = _bilinear_upsample_u16(l1, 16, height, width)  # scaled by 256
    o2 = _bilinear_upsample_u16(l2, 6, height, width)   # scaled by 36
    v = (o1 * 3 * 9 + o2 * 64) // (256 * 9 * 4)         # (o1/256*3 + o2/36)/4
    return np.clip(v, 0, 255).astype(np.uint8)


class FrameSource:
    """frame(k): rows x cols u8 view of the texture after k steps of motion."""

    def __init__(self, rows: int, cols: int, seed: int = 0, margin: int = 160, rot_per_frame: float = 0.0012,
                 shift_per_frame=(1.9, -1.3)):
        self.rows, self.cols, self.margin = rows, cols, margin
        self.tex = make_texture(rows + 2 * margin, cols + 2 * margin, seed)
        self.rot, self.shift = rot_per_frame, shift_per_frame
        ys, xs = np.mgrid[0:rows, 0:cols]
        self._x = xs.astype(np.float64) - cols / 2.0
        self._y = ys.astype(np.float64) - rows / 2.0

    def frame(self, k: int) -> np.ndarray:
        t = self.rot * k / 2.0  # tan(theta/2)
        c, s = (1.0 - t * t) / (1.0 + t * t), 2.0 * t / (1.0 + t * t)
        # bounded, smooth, non-periodic-looking translation
        kk = k % 64
        tri = kk if kk < 32 else 64 - kk
        tx, ty = self.shift[0] * tri, self.shift[1] * tri
        X = c * self._x - s * self._y + self.cols / 2.0 + tx + self.margin
        Y = s * self._x + c * self._y + self.rows / 2.0 + ty + self.margin
        Xq = np.floor(X * 32.0 + 0.5).astype(np.int64)
        Yq = np.floor(Y * 32.0 + 0.5).astype(np.int64)
        x0, fx = Xq >> 5, (Xq & 31).astype(np.uint32)
        y0, fy = Yq >> 5, (Yq & 31).astype(np.uint32)
        H, W = self.tex.shape
        x0 = np.clip(x0, 0, W - 2)
        y0 = np.clip(y0, 0, H - 2)
        T = self.tex.astype(np.uint32)
        a, b = T[y0, x0], T[y0, x0 + 1]
        cc, d = T[y0 + 1, x0], T[y0 + 1, x0 + 1]
        v = a * (32 - fx) * (32 - fy) + b * fx * (32 - fy) + cc * (32 - fx) * fy + d * fx * fy
        return ((v + 512) >> 10).astype(np.uint8)

    def frames(self, n: int, start: int = 0):
        return [self.frame(start + k) for k in range(n)]


def crc(a: np.ndarray) -> int:
    return zlib.crc32(np.ascontiguousarray(a).tobytes()) & 0xFFFFFFFF
