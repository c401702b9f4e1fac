"""CPU tests: pin the front-end oracle (numpy restatements + control flow) against cv2 and the golden fixtures."""
import os

import numpy as np
import pytest

cv2 = pytest.importorskip("cv2")

from lfvio import synth
from lfvio.config import PAL_UP, tracker_defaults
from oracle import frontend as fe

ROWS, COLS = 960, 1280


@pytest.fixture(scope="module")
def stages(golden_dir):
    return np.load(os.path.join(golden_dir, "frontend_stages.npz"))


@pytest.fixture(scope="module")
def frames(stages):
    src = synth.FrameSource(ROWS, COLS, seed=int(stages["seed"]))
    f = [src.frame(int(k)) for k in stages["frame_idx"]]
    assert [synth.crc(x) for x in f] == [int(c) for c in stages["frame_crc"]], "synthetic frames are not reproducible"
    return f


def test_clahe_restatement_bit_exact(frames, stages):
    for f, want in zip(frames, stages["clahe_crc"]):
        got = fe.clahe(f)
        assert synth.crc(got) == int(want)
        assert np.array_equal(got, cv2.createCLAHE(3.0, (8, 8)).apply(f))


def test_clahe_ragged_size_bit_exact(frames):
    f = np.ascontiguousarray(frames[0][:477, :635])
    assert np.array_equal(fe.clahe(f), cv2.createCLAHE(3.0, (8, 8)).apply(f))


def test_pyramid_and_scharr_bit_exact(frames, stages):
    img = fe.clahe(frames[0])
    pyr = fe.build_pyramid(img)
    assert len(pyr) == 4
    assert [synth.crc(p) for p in pyr] == [int(c) for c in stages["pyr_crc"]]
    small = np.ascontiguousarray(img[:301, :403])
    assert np.array_equal(fe.pyr_down(small), cv2.pyrDown(small))
    ix, iy = fe.scharr(small)
    assert np.array_equal(ix, cv2.Scharr(small, cv2.CV_16S, 1, 0))
    assert np.array_equal(iy, cv2.Scharr(small, cv2.CV_16S, 0, 1))


def test_min_eig_map_bit_exact(frames, stages):
    img = fe.clahe(frames[0])
    eig = fe.min_eig_map(img)
    assert np.array_equal(eig.ravel()[stages["eig_samples_idx"]], stages["eig_samples"])
    assert synth.crc(eig) == int(stages["eig_crc"][0])


def test_good_features_matches_cv2_order(frames, stages):
    img = fe.clahe(frames[0])
    cfg = tracker_defaults()
    mask = np.zeros((ROWS, COLS), np.uint8)
    fe.fill_circle(mask, int(cfg["center_x"]), int(cfg["center_y"]), cfg["max_r"], 255)
    fe.fill_circle(mask, int(cfg["center_x"]), int(cfg["center_y"]), cfg["min_r"], 0)
    assert np.array_equal(fe.good_features_to_track(img, 150, 0.01, 20.0, mask), stages["gftt_masked"])
    assert np.array_equal(fe.good_features_to_track(img, 500, 0.01, 10.0, None), stages["gftt_nomask"])


@pytest.mark.parametrize("r", [1, 2, 7, 20, 33, 160])
def test_circle_raster_matches_cv2(r):
    a = np.zeros((400, 420), np.uint8)
    b = a.copy()
    cv2.circle(a, (200, 190), r, 255, -1)
    fe.fill_circle(b, 200, 190, r, 255)
    assert np.array_equal(a, b)
    a2 = np.zeros((60, 60), np.uint8)
    b2 = a2.copy()
    cv2.circle(a2, (3, 57), r, 9, -1)
    fe.fill_circle(b2, 3, 57, r, 9)
    assert np.array_equal(a2, b2)


def test_lk_restatement_matches_cv2(frames, stages):
    c0, c1 = fe.clahe(frames[0]), fe.clahe(frames[1])
    sel = np.r_[0:40, len(stages["lk_in"]) - 10:len(stages["lk_in"])]  # 40 interior + the 10 border cases
    nxt, st = fe.lk_track(c0, c1, stages["lk_in"][sel])
    assert np.array_equal(st, stages["lk_status"][sel])
    ok = st.astype(bool)
    assert np.abs(nxt - stages["lk_out"][sel])[ok].max() < 1e-3  # px; f32 accumulation order differs from SIMD


def test_std_sort_perm_is_a_valid_descending_permutation():
    rng = np.random.default_rng(0)
    for n in (0, 1, 5, 16, 17, 150, 2000):
        k = rng.integers(1, 8, n).astype(np.int32)
        p = fe.std_sort_desc_perm(k)
        assert sorted(p.tolist()) == list(range(n))
        assert np.all(np.diff(k[p]) <= 0)


def test_control_flow_restated_backend_equals_cv2_backend():
    """Three frames at reduced size: the numpy-restated operators drive the tracker to the same ids/cnt as cv2."""
    cam = dict(PAL_UP, image_width=640, image_height=480, cx=322.5, cy=243.0)
    cfg = dict(tracker_defaults(), center_x=320.5, center_y=245.5, max_r=230, min_r=70, max_cnt=40)
    src = synth.FrameSource(480, 640, seed=3, rot_per_frame=0.003, shift_per_frame=(3.1, -2.2))
    a = fe.FeatureTrackerOracle(cam, cfg, "cv2")
    b = fe.FeatureTrackerOracle(cam, cfg, "restated")
    for k in range(3):
        img = src.frame(k)
        for t in (a, b):
            t.read_image(img, 5.0 + k / 15, True)
            t.update_ids()
        assert np.array_equal(a.ids, b.ids) and np.array_equal(a.track_cnt, b.track_cnt)
        assert np.abs(a.cur_pts - b.cur_pts).max() < 2e-3
        assert np.array_equal(a.mask, b.mask)
        assert np.abs(a.pts_velocity_3d - b.pts_velocity_3d).max() < 1e-3


def test_golden_sequences_reproduce_from_cv2(golden_dir):
    """The committed fixtures are what cv2 + the restated control flow produce here (guards against drift)."""
    g = np.load(os.path.join(golden_dir, "frontend_seq_fast.npz"))
    src = synth.FrameSource(ROWS, COLS, seed=int(g["seed"]), rot_per_frame=0.004, shift_per_frame=(6.5, -4.2))
    tr = fe.FeatureTrackerOracle(PAL_UP, tracker_defaults(), "cv2")
    for k in range(4):
        img = src.frame(k)
        assert synth.crc(img) == int(g["frame_crc"][k])
        tr.read_image(img, 100.0 + k * float(g["dt"]), bool(g["pub"][k]))
        tr.update_ids()
        assert np.array_equal(tr.ids, g[f"f{k}_ids"]) and np.array_equal(tr.track_cnt, g[f"f{k}_track_cnt"])
        assert np.abs(tr.cur_pts - g[f"f{k}_cur_pts"]).max() < 1e-3
