"""CPU tests of the rejectWithF restatement (oracle/ransac.py) — reference feature_tracker.cpp:186-386."""
import numpy as np

from lfvio import config
from oracle import ransac as R
from ransac_data import rot, two_view

CAM, CFG = config.PAL_UP, config.tracker_defaults()


def test_sample_indices_follow_create_random_array():
    for n in (8, 9, 30, 150, 2000):
        for h in range(20):
            idx = R.sample_indices(1234, 7, h, n)
            assert len(idx) == 8 and idx == sorted(set(idx)) and 0 <= idx[0] and idx[-1] < n
    assert R.sample_indices(1, 2, 3, 150) == R.sample_indices(1, 2, 3, 150)
    assert R.sample_indices(1, 2, 3, 150) != R.sample_indices(1, 2, 4, 150)
    assert R.sample_indices(5, 0, 0, 8) == list(range(8))


def test_compute_E_recovers_the_essential_matrix():
    rng = np.random.default_rng(0)
    b1 = rng.normal(size=(40, 3))
    b1 /= np.linalg.norm(b1, axis=1, keepdims=True)
    Rm, t = rot(0.02, -0.01, 0.05), np.array([0.3, -0.1, 0.2])
    P2 = (Rm @ (b1 * rng.uniform(2, 9, 40)[:, None]).T).T + t
    b2 = P2 / np.linalg.norm(P2, axis=1, keepdims=True)
    tx = np.array([[0, -t[2], t[1]], [t[2], 0, -t[0]], [-t[1], t[0], 0]])
    E_true = tx @ Rm
    for sel in (slice(0, 8), slice(0, 40)):
        E = R.compute_E_21(b1[sel], b2[sel])
        assert abs(np.linalg.svd(E)[1][2]) < 1e-12
        c = np.sum(E * E_true) / np.sum(E_true * E_true)
        assert np.allclose(E, c * E_true, atol=1e-9 * abs(c) + 1e-12)
        assert np.max(np.abs(np.einsum("ni,ij,nj->n", b2, E, b1))) < 1e-12


def test_check_inliers_keeps_the_reference_quirks():
    # one point passes in image 2 but fails in image 1: it adds its first margin to the score but is no inlier
    E = np.array([[0, -0.2, 0.1], [0.2, 0, -0.3], [-0.1, 0.3, 0.0]])
    rng = np.random.default_rng(3)
    b1 = rng.normal(size=(200, 3))
    b1 /= np.linalg.norm(b1, axis=1, keepdims=True)
    b2 = b1 + rng.normal(scale=0.004, size=b1.shape)
    b2 /= np.linalg.norm(b2, axis=1, keepdims=True)
    score, inl = R.check_inliers(E, b1, b2)
    r2, r1 = R._residual(E, b1, b2), R._residual(E.T.copy(), b2, b1)
    assert np.array_equal(inl.astype(bool), (r2 <= R.THR) & (r1 <= R.THR))
    want = np.float32(0)
    for i in range(len(b1)):
        if r2[i] <= R.THR:
            want = np.float32(np.float64(want) + (R.THR - r2[i]) ** 2)
            if r1[i] <= R.THR:
                want = np.float32(np.float64(want) + (R.THR - r1[i]) ** 2)
    assert score == float(want) and 0 < inl.sum() < len(inl)
    assert R.THR == float(np.float32(0.00872653549837)) and R.THR != 0.00872653549837


def test_find_fundamental_rejects_the_outliers():
    cur, forw, out = two_view(CAM, CFG, 150, seed=11)
    status = R.reject_with_f(CAM, cur, forw, seed=42, frame=3)
    assert status[~out].mean() > 0.97   # inliers kept (noise 0.15 px against a 0.5 degree gate)
    assert status[out].mean() < 0.35    # gross outliers rejected (those displaced along their epipolar plane pass)
    assert np.array_equal(status, R.reject_with_f(CAM, cur, forw, seed=42, frame=3))
    assert np.array_equal(R.reject_with_f(CAM, cur[:7], forw[:7], 1, 0), np.ones(7, np.uint8))
