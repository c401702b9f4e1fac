"""TEST INFRASTRUCTURE — CPU restatement of FeatureTracker::rejectWithF.

Follows feature_tracker/src/feature_tracker.cpp of the reference:
  compute_E_21            :186-218   8-point (or N-point) null vector of the bearing constraint matrix + rank-2 projection
  check_inliers           :219-264   symmetric epipolar-plane angular residual, float score, asymmetric quirks kept
  myfindFundamentalMat    :266-327   100 hypotheses, best = first strictly larger score, refit on its inliers
  rejectWithF             :329-386   lift cur/forw to the unit sphere, run the RANSAC, reduce by status
  create_random_array     random_array.cc:22-49  draw 9, sort, unique, keep the 8 smallest, repeat until 8 (then shuffle)

PARITY PINNING: the reference seeds a fresh std::mt19937 from std::random_device on EVERY call
(random_array.cc:12-20), so its output is not reproducible even against itself; there is no golden vector to pin
this restatement to.  The random source is therefore INJECTED: a counter-based generator (splitmix64 of
seed/frame/hypothesis/draw) that the CUDA path implements identically, feeding the reference's own sampling
procedure.  The final std::shuffle only permutes the rows of the 8x9 constraint matrix, which does not change its
null space, and is omitted.  Everything after the sampling follows the reference operation by operation in IEEE
double without fused multiply-add (the reference builds with plain -O3 on x86-64, feature_tracker/CMakeLists.txt:6-7);
the singular value decompositions are LAPACK's (numpy) where the reference uses Eigen::JacobiSVD — both return the
same subspaces to rounding, so residuals agree to ~1e-15 and inlier decisions agree except on a knife edge.
"""
from __future__ import annotations

import numpy as np

from . import camera as ocam

f32, f64 = np.float32, np.float64
MASK64 = (1 << 64) - 1
NUM_ITER = 100                         # max_num_iter            feature_tracker.cpp:278
MIN_SET = 8                            # min_set_size            :271
THR = float(f32(0.00872653549837))     # constexpr float residual_cos_thr  :227 (sin 0.5 deg, rounded to float)


def splitmix64(x: int) -> int:
    z = (x + 0x9E3779B97F4A7C15) & MASK64
    z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & MASK64
    z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & MASK64
    return z ^ (z >> 31)


def rng_word(seed: int, frame: int, hyp: int, k: int) -> int:
    """32-bit word number k of hypothesis `hyp` of frame `frame` (the injected random source)."""
    key = (seed + frame * 0xD1B54A32D192ED03 + hyp * 0x8CB92BA72F3D8DD7 + k * 0xA24BAED4963EE407) & MASK64
    return splitmix64(key) >> 32


def sample_indices(seed: int, frame: int, hyp: int, n: int) -> list[int]:
    """create_random_array(8, 0, n-1) (random_array.cc:22-49) on the injected source; result left sorted."""
    v: list[int] = []
    k = 0
    make_size = int(MIN_SET * 1.2)  # 9
    while len(v) != MIN_SET:
        while len(v) < make_size:
            v.append((rng_word(seed, frame, hyp, k) * n) >> 32)  # uniform in [0, n-1]
            k += 1
        v = sorted(set(v))
        if len(v) > MIN_SET:
            v = v[:MIN_SET]
    return v


def compute_E_21(b1: np.ndarray, b2: np.ndarray) -> np.ndarray:
    """feature_tracker.cpp:186-218."""
    A = np.concatenate([b2[:, 0:1] * b1, b2[:, 1:2] * b1, b2[:, 2:3] * b1], axis=1)  # rows: b2(r) * b1(c) at 3r+c
    if len(A) < 9:
        A = np.concatenate([A, np.zeros((9 - len(A), 9))])  # full V needs >= 9 rows of work space; null space unchanged
    _, _, Vt = np.linalg.svd(A, full_matrices=True)
    E = Vt[8].reshape(3, 3)  # Matrix3d(v.data()).transpose(): row-major read
    U, lam, Vt3 = np.linalg.svd(E)
    lam[2] = 0.0
    return (U * lam) @ Vt3


def _residual(E: np.ndarray, a: np.ndarray, b: np.ndarray) -> np.ndarray:
    """|(E a) . b| / |E a| per row, in the reference's operation order, no FMA."""
    e0 = E[0, 0] * a[:, 0] + E[0, 1] * a[:, 1] + E[0, 2] * a[:, 2]
    e1 = E[1, 0] * a[:, 0] + E[1, 1] * a[:, 1] + E[1, 2] * a[:, 2]
    e2 = E[2, 0] * a[:, 0] + E[2, 1] * a[:, 1] + E[2, 2] * a[:, 2]
    dot = e0 * b[:, 0] + e1 * b[:, 1] + e2 * b[:, 2]
    with np.errstate(invalid="ignore", divide="ignore"):
        return np.abs(dot / np.sqrt(e0 * e0 + e1 * e1 + e2 * e2))


def check_inliers(E21: np.ndarray, b1: np.ndarray, b2: np.ndarray):
    """feature_tracker.cpp:219-264.  Returns (score as the float the reference accumulates, inlier flags)."""
    r2 = _residual(E21, b1, b2)
    r1 = _residual(E21.T.copy(), b2, b1)
    pass2 = ~(THR < r2)  # NaN compares false, i.e. passes — as in the reference
    pass1 = ~(THR < r1)
    inl = pass2 & pass1
    score = f32(0.0)
    t2 = (THR - r2) * (THR - r2)  # pow(x, 2.0) == x * x
    t1 = (THR - r1) * (THR - r1)
    for i in range(len(b1)):  # sequential float accumulation (float score; `score += double`)
        if pass2[i]:
            score = f32(f64(score) + t2[i])
            if pass1[i]:
                score = f32(f64(score) + t1[i])
    return float(score), inl.astype(np.uint8)


def find_fundamental(b1: np.ndarray, b2: np.ndarray, seed: int, frame: int):
    """myfindFundamentalMat, feature_tracker.cpp:266-327.  Returns (E, status, scores[100], rows of the refit)."""
    n = len(b1)
    best_score, best_inl = 0.0, None
    scores = np.zeros(NUM_ITER, f32)
    for it in range(NUM_ITER):
        idx = sample_indices(seed, frame, it, n)
        E = compute_E_21(b1[idx], b2[idx])
        score, inl = check_inliers(E, b1, b2)
        scores[it] = score
        if best_score < score:
            best_score, best_inl = score, inl
    if best_inl is None:  # the reference would index an empty vector here; the CUDA path falls back to hypothesis 0
        _, best_inl = check_inliers(compute_E_21(b1[sample_indices(seed, frame, 0, n)], b2[sample_indices(seed, frame, 0, n)]), b1, b2)
    sel = best_inl.astype(bool)
    E = compute_E_21(b1[sel], b2[sel])
    _, status = check_inliers(E, b1, b2)
    return E, status, scores, int(sel.sum())


def lift_pair(cam: dict, cur_pts: np.ndarray, forw_pts: np.ndarray):
    """rejectWithF's un_cur_pts / un_forw_pts (:336-354): liftProjective then tmp_p / tmp_p.norm(), in double."""
    return ocam.lift_sphere(cam, cur_pts.astype(f64)), ocam.lift_sphere(cam, forw_pts.astype(f64))


def reject_with_f(cam: dict, cur_pts: np.ndarray, forw_pts: np.ndarray, seed: int, frame: int) -> np.ndarray:
    """Status vector of rejectWithF (all ones when fewer than 8 points, :331)."""
    if len(forw_pts) < 8:
        return np.ones(len(forw_pts), np.uint8)
    b1, b2 = lift_pair(cam, cur_pts, forw_pts)
    return find_fundamental(b1, b2, seed, frame)[1]


def make_injected(seed: int):
    """`ransac` callable for FeatureTrackerOracle; frame = number of frames the tracker has seen before this one."""

    def hook(tracker, cur_pts, forw_pts):
        return reject_with_f(tracker.cam, cur_pts, forw_pts, seed, tracker.frames_seen)

    return hook
