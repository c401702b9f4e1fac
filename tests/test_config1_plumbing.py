"""BASELINE.json configs[0] / SURVEY.md §8(d) "Config 1" — the reference's own CPU-runnable case, no GPU: ONE 1280x960 PAL
frame pair (band-limited noise: rng(0) 160x120 uniform u8 -> bicubic x8 -> Gaussian 5x5 sigma 1.5; second frame = first
warped by (+2.3, -1.7) px), annulus (640, 490) r in [160, 500], 150 points from goodFeaturesToTrack(.., 150, 0.01, 20, mask),
LK(41x41, 3), lift of all points with the PAL-up calibration.  cv2 (the oracle of record for the OpenCV stages) against the
numpy restatements the kernels are built to, and both against the known shift; the C-ABI library is only loaded (plumbing)."""
import numpy as np
import pytest

cv2 = pytest.importorskip("cv2")

from lfvio.config import PAL_UP
from oracle import camera as ocam
from oracle import frontend as fe

SHIFT = (2.3, -1.7)


@pytest.fixture(scope="module")
def pair():
    rng = np.random.default_rng(0)
    small = rng.integers(0, 256, (120, 160), dtype=np.uint8)
    big = cv2.resize(small, (1280, 960), interpolation=cv2.INTER_CUBIC)
    f0 = cv2.GaussianBlur(big, (5, 5), 1.5)
    M = np.float32([[1, 0, SHIFT[0]], [0, 1, SHIFT[1]]])
    f1 = cv2.warpAffine(f0, M, (1280, 960), flags=cv2.INTER_LINEAR, borderMode=cv2.BORDER_REFLECT_101)
    mask = np.zeros((960, 1280), np.uint8)
    cv2.circle(mask, (640, 490), 500, 255, -1)
    cv2.circle(mask, (640, 490), 160, 0, -1)
    return f0, f1, mask


def test_config1_detect_track_lift(pair, lfvio_mod):
    f0, f1, mask = pair
    assert lfvio_mod.lib.lfvio_version()                              # the drop-in library loads on a CPU-only box
    pts_cv = cv2.goodFeaturesToTrack(f0, 150, 0.01, 20, mask=mask).reshape(-1, 2)
    pts = fe.good_features_to_track(f0, 150, 0.01, 20, mask)
    assert len(pts_cv) == 150 and np.array_equal(np.asarray(pts, np.float32).reshape(-1, 2), pts_cv)   # same corners, same order
    r = np.hypot(pts_cv[:, 0] - 640, pts_cv[:, 1] - 490)
    assert (r <= 500.5).all() and (r >= 159.5).all()                  # inside the annulus
    nxt_cv, st_cv, _ = cv2.calcOpticalFlowPyrLK(f0, f1, pts_cv.reshape(-1, 1, 2), None, winSize=(41, 41), maxLevel=3)
    nxt_cv, st_cv = nxt_cv.reshape(-1, 2), st_cv.ravel()
    nxt, st = fe.lk_track(f0, f1, pts_cv)[:2]
    assert np.array_equal(np.asarray(st).ravel().astype(np.uint8), st_cv)
    ok = st_cv == 1
    assert ok.sum() >= 145
    assert np.abs(np.asarray(nxt)[ok] - nxt_cv[ok]).max() < 2e-3     # restatement vs cv2 (SIMD summation order)
    d = nxt_cv[ok] - pts_cv[ok]
    assert np.abs(np.median(d, axis=0) - SHIFT).max() < 0.02          # the known shift
    assert np.abs(d - SHIFT).max() < 0.15
    rays = ocam.lift_projective(PAL_UP, nxt_cv[ok].astype(np.float64))
    b = ocam.lift_sphere(PAL_UP, nxt_cv[ok].astype(np.float64))
    assert np.allclose(np.linalg.norm(b, axis=1), 1.0, atol=1e-12)
    assert np.allclose(b, rays / np.linalg.norm(rays, axis=1, keepdims=True), atol=1e-15)
    rr = np.hypot(nxt_cv[ok][:, 0] - PAL_UP["cx"], nxt_cv[ok][:, 1] - PAL_UP["cy"])
    assert ((rays[:, 2] < 0) == (rr > PAL_UP["inv_poly"][0])).all()   # the negative half-space starts at phi = 376.85 px
    assert (rays[:, 2] < 0).any() and (rays[:, 2] > 0).any()
