"""GPU parity tests of the panoramic front end, called through the C-ABI (include/lfvio.h).

Bars: bit-exact for byte / integer / index work (CLAHE, pyramid, mask, corner coordinates and order, ids,
track_cnt) and for the f32 lambda_min map; LK positions within 2e-3 px of cv2 (f32 accumulation order), with
identical status; unit bearings within 2e-6, bearing velocities within 1e-4 (f32 of a difference / dt).
"""
import os

import numpy as np
import pytest

from lfvio import synth
from lfvio.config import PAL_UP, tracker_defaults
from oracle import frontend as fe

pytestmark = pytest.mark.gpu
ROWS, COLS = 960, 1280
LK_TOL_PX = 2e-3


@pytest.fixture(scope="module")
def stages(golden_dir):
    return np.load(os.path.join(golden_dir, "frontend_stages.npz"))


@pytest.fixture(scope="module")
def frames(stages):
    src = synth.FrameSource(ROWS, COLS, seed=int(stages["seed"]))
    f = [src.frame(int(k)) for k in stages["frame_idx"]]
    assert [synth.crc(x) for x in f] == [int(c) for c in stages["frame_crc"]], "synthetic frames are not reproducible"
    return f


@pytest.fixture(scope="module")
def eq_frames(frames):
    return [fe.clahe(f) for f in frames]  # oracle CLAHE (bit-exact with cv2, see test_oracle_frontend.py)


def annulus(cfg, rows=ROWS, cols=COLS):
    m = np.zeros((rows, cols), np.uint8)
    fe.fill_circle(m, int(cfg["center_x"]), int(cfg["center_y"]), cfg["max_r"], 255)
    fe.fill_circle(m, int(cfg["center_x"]), int(cfg["center_y"]), cfg["min_r"], 0)
    return m


# ---- K1 ---------------------------------------------------------------------------------------------
def test_clahe_bit_exact(lfvio_mod, gpu, frames, stages):
    for f, want in zip(frames, stages["clahe_crc"]):
        got = lfvio_mod.clahe(f, device=gpu)
        assert synth.crc(got) == int(want)


@pytest.mark.parametrize("shape", [(477, 635), (480, 640), (123, 1001), (64, 64)])
def test_clahe_ragged_and_small(lfvio_mod, gpu, frames, shape):
    f = np.ascontiguousarray(frames[0][:shape[0], :shape[1]])
    assert np.array_equal(lfvio_mod.clahe(f, device=gpu), fe.clahe(f))


def test_clahe_flat_and_saturated(lfvio_mod, gpu):
    for v in (0, 255, 17):
        f = np.full((240, 320), v, np.uint8)
        assert np.array_equal(lfvio_mod.clahe(f, device=gpu), fe.clahe(f))


# ---- K2 ---------------------------------------------------------------------------------------------
def test_pyramid_bit_exact(lfvio_mod, gpu, eq_frames, stages):
    img = eq_frames[0]
    for lvl in range(1, 4):
        img = lfvio_mod.pyr_down(img, device=gpu)
        assert synth.crc(img) == int(stages["pyr_crc"][lvl])


@pytest.mark.parametrize("shape", [(301, 403), (43, 45), (1, 7), (5, 1), (100, 2)])
def test_pyr_down_ragged(lfvio_mod, gpu, eq_frames, shape):
    f = np.ascontiguousarray(eq_frames[0][:shape[0], :shape[1]])
    assert np.array_equal(lfvio_mod.pyr_down(f, device=gpu), fe.pyr_down(f))


# ---- K3 ---------------------------------------------------------------------------------------------
def test_lk_matches_cv2_golden(lfvio_mod, gpu, eq_frames, stages):
    nxt, st = lfvio_mod.calc_optical_flow_pyr_lk(eq_frames[0], eq_frames[1], stages["lk_in"], device=gpu)
    assert np.array_equal(st, stages["lk_status"])
    ok = st.astype(bool)
    err = np.abs(nxt - stages["lk_out"])[ok]
    assert err.max() < LK_TOL_PX, err.max()


def test_lk_border_and_degenerate_points(lfvio_mod, gpu, eq_frames):
    """Points whose window leaves the image at some level, flat patches (minEig reject), against the restated oracle."""
    a, b = eq_frames[0].copy(), eq_frames[1].copy()
    a[400:520, 600:760] = 128  # flat patch -> minEig < 1e-4 => status 0 at level 0
    b[400:520, 600:760] = 128
    pts = np.array([[680, 460], [1.0, 1.0], [1278.6, 958.7], [20.4, 940.1], [1260.2, 19.9], [640, 0.2], [-3.0, 480.0],
                    [1300.0, 480.0], [640.0, 1000.0], [333.3, 222.2]], np.float32)
    nxt, st = lfvio_mod.calc_optical_flow_pyr_lk(a, b, pts, device=gpu)
    want, want_st = fe.lk_track(a, b, pts)
    assert np.array_equal(st, want_st)
    assert st[0] == 0
    ok = st.astype(bool)
    assert np.abs(nxt - want)[ok].max() < LK_TOL_PX


def test_lk_small_image_fewer_levels(lfvio_mod, gpu, eq_frames):
    a = np.ascontiguousarray(eq_frames[0][200:400, 300:520])  # 200x220: level 3 would be 25x28 <= 41 -> 3 levels
    b = np.ascontiguousarray(eq_frames[0][202:402, 303:523])
    pts = np.array([[100, 100], [60.5, 140.25], [180, 30]], np.float32)
    nxt, st = lfvio_mod.calc_optical_flow_pyr_lk(a, b, pts, device=gpu)
    want, want_st = fe.lk_track(a, b, pts)
    assert np.array_equal(st, want_st)
    assert np.abs(nxt - want)[st.astype(bool)].max() < LK_TOL_PX


def test_lk_zero_points(lfvio_mod, gpu, eq_frames):
    nxt, st = lfvio_mod.calc_optical_flow_pyr_lk(eq_frames[0], eq_frames[1], np.zeros((0, 2), np.float32), device=gpu)
    assert nxt.shape == (0, 2) and st.shape == (0,)


# ---- K4 ---------------------------------------------------------------------------------------------
def test_min_eig_map_bit_exact(lfvio_mod, gpu, eq_frames, stages):
    eig = lfvio_mod.corner_min_eigen_val(eq_frames[0], device=gpu)
    assert np.array_equal(eig.ravel()[stages["eig_samples_idx"]], stages["eig_samples"])
    assert synth.crc(eig) == int(stages["eig_crc"][0])
    assert eig.max() == stages["eig_max"]


@pytest.mark.parametrize("shape", [(100, 131), (33, 17), (16, 32), (3, 3)])
def test_min_eig_ragged(lfvio_mod, gpu, eq_frames, shape):
    f = np.ascontiguousarray(eq_frames[0][100:100 + shape[0], 200:200 + shape[1]])
    assert np.array_equal(lfvio_mod.corner_min_eigen_val(f, device=gpu), fe.min_eig_map(f))


def test_good_features_identical_to_cv2(lfvio_mod, gpu, eq_frames, stages):
    cfg = tracker_defaults()
    got = lfvio_mod.good_features_to_track(eq_frames[0], 150, 0.01, 20.0, annulus(cfg), device=gpu)
    assert np.array_equal(got, stages["gftt_masked"])
    got = lfvio_mod.good_features_to_track(eq_frames[0], 500, 0.01, 10.0, None, device=gpu)
    assert np.array_equal(got, stages["gftt_nomask"])


@pytest.mark.parametrize("k,md", [(1, 20.0), (7, 3.0), (2000, 20.0), (5000, 7.0), (300, 1.0), (50, 64.0)])
def test_good_features_vs_oracle_sweep(lfvio_mod, gpu, eq_frames, k, md):
    img = np.ascontiguousarray(eq_frames[1][:600, :800])
    mask = annulus(dict(center_x=400, center_y=300, max_r=290, min_r=60), 600, 800)
    got = lfvio_mod.good_features_to_track(img, k, 0.01, md, mask, device=gpu)
    want = fe.good_features_to_track(img, k, 0.01, md, mask)
    assert np.array_equal(got, want)


def test_good_features_empty_mask_and_flat_image(lfvio_mod, gpu, eq_frames):
    img = eq_frames[0]
    assert len(lfvio_mod.good_features_to_track(img, 10, 0.01, 20.0, np.zeros_like(img), device=gpu)) == 0
    flat = np.full((200, 200), 77, np.uint8)
    assert len(lfvio_mod.good_features_to_track(flat, 10, 0.01, 20.0, None, device=gpu)) == 0


def test_good_features_ties(lfvio_mod, gpu):
    """Periodic pattern => many exactly tied lambda values; order must follow (value desc, index desc)."""
    tile = np.zeros((16, 16), np.uint8)
    tile[4:12, 4:12] = 200
    img = np.tile(tile, (12, 15))
    got = lfvio_mod.good_features_to_track(img, 400, 0.01, 5.0, None, device=gpu)
    want = fe.good_features_to_track(img, 400, 0.01, 5.0, None)
    assert len(want) > 100
    assert np.array_equal(got, want)


# ---- whole tracker ------------------------------------------------------------------------------------
SEQ_KW = {"seq_slow": {}, "seq_fast": dict(rot_per_frame=0.004, shift_per_frame=(6.5, -4.2))}


@pytest.mark.parametrize("name", ["seq_slow", "seq_fast"])
def test_tracker_sequence_matches_golden(lfvio_mod, gpu, golden_dir, name):
    g = np.load(os.path.join(golden_dir, f"frontend_{name}.npz"))
    cfg = tracker_defaults()
    src = synth.FrameSource(ROWS, COLS, seed=int(g["seed"]), **SEQ_KW[name])
    tr = lfvio_mod.FeatureTracker(PAL_UP, cfg, device=gpu)
    for k in range(int(g["n_frames"])):
        img = src.frame(k)
        assert synth.crc(img) == int(g["frame_crc"][k])
        tr.read_image(img, 100.0 + k * float(g["dt"]), bool(g["pub"][k]))
        assert np.array_equal(tr.ids, g[f"f{k}_pre_ids"]), f"frame {k}: ids before updateID"
        tr.update_ids()
        assert np.array_equal(tr.ids, g[f"f{k}_ids"]), f"frame {k}: ids"
        assert np.array_equal(tr.track_cnt, g[f"f{k}_track_cnt"]), f"frame {k}: track_cnt"
        assert np.abs(tr.cur_pts - g[f"f{k}_cur_pts"]).max() < 5e-3, f"frame {k}: cur_pts"
        assert np.abs(tr.cur_un_pts_3d - g[f"f{k}_cur_un_pts_3d"]).max() < 2e-5, f"frame {k}: bearings"
        assert np.abs(tr.pts_velocity_3d - g[f"f{k}_pts_velocity_3d"]).max() < 5e-4, f"frame {k}: velocity"
        if g["pub"][k]:
            assert synth.crc(tr.mask) == int(g["mask_crc"][k]), f"frame {k}: mask bytes"
    tr.close()


def test_tracker_cur_img_is_clahe_pyramid(lfvio_mod, gpu, frames, stages):
    tr = lfvio_mod.FeatureTracker(PAL_UP, tracker_defaults(), device=gpu)
    tr.read_image(frames[0], 1.0, True)
    for lvl in range(4):
        assert synth.crc(tr.cur_img(lvl)) == int(stages["pyr_crc"][lvl])
    tr.close()


def test_set_mask_unstable_sort_order(lfvio_mod, gpu, frames):
    """Tie-heavy track_cnt: kept set and order must follow libstdc++'s std::sort permutation (feature_tracker.cpp:66-67)."""
    cfg = tracker_defaults(max_cnt=400)
    rng = np.random.default_rng(11)
    for n in (5, 16, 17, 150, 400):
        ang = rng.uniform(0, 2 * np.pi, n)
        r = rng.uniform(100, 560, n)  # some fall outside the annulus
        pts = np.stack([640 + r * np.cos(ang), 490 + r * np.sin(ang)], 1).astype(np.float32)
        pts[:, 0] = np.clip(pts[:, 0], 2, COLS - 3)
        pts[:, 1] = np.clip(pts[:, 1], 2, ROWS - 3)
        cnt = rng.integers(1, 5, n).astype(np.int32)
        ids = np.arange(n, dtype=np.int32)
        # survivors of the LK pass on identical images (same kernel, stage-level entry point) + inBorder
        flat = frames[0]
        nxt, st = lfvio_mod.calc_optical_flow_pyr_lk(flat, flat, pts, device=gpu)
        assert np.abs(nxt - pts)[st.astype(bool)].max() < 1e-4  # b == 0 => delta == 0; (x - 20) + 20 may move 1 ulp
        x, y = np.rint(nxt[:, 0]), np.rint(nxt[:, 1])
        ok = st.astype(bool) & (1 <= x) & (x < COLS - 1) & (1 <= y) & (y < ROWS - 1)
        # oracle: setMask on the surviving state
        o = fe.FeatureTrackerOracle(PAL_UP, cfg, "restated")
        o.forw_pts, o.ids, o.track_cnt = nxt[ok].copy(), ids[ok].astype(np.int64), cnt[ok].astype(np.int64)
        o.set_mask()
        tr = lfvio_mod.FeatureTracker(PAL_UP, dict(cfg, equalize=0), device=gpu)
        tr.read_image(flat, 0.0, False)          # first frame, no detection
        tr.set_points(pts, ids, cnt - 1)         # LK on identical images keeps positions, cnt++ restores cnt
        tr.read_image(flat, 0.1, True)
        old = tr.ids >= 0                        # tracked (not newly detected) entries come first, in setMask order
        assert np.array_equal(tr.ids[old], o.ids), f"n={n}"
        assert np.array_equal(tr.track_cnt[old], o.track_cnt)
        assert np.array_equal(tr.cur_pts[old], o.forw_pts)
        assert len(o.ids) < ok.sum() or n <= 5   # the greedy loop actually rejected something
        tr.close()


def test_device_async_path_equals_sync_path(lfvio_mod, gpu, golden_dir):
    """read_image_dev (image resident in HBM, ids assigned on device, no per-frame download) == read_image."""
    import torch

    g = np.load(os.path.join(golden_dir, "frontend_seq_fast.npz"))
    src = synth.FrameSource(ROWS, COLS, seed=int(g["seed"]), **SEQ_KW["seq_fast"])
    tr = lfvio_mod.FeatureTracker(PAL_UP, tracker_defaults(), device=gpu)
    n = int(g["n_frames"])
    dev = [torch.from_numpy(src.frame(k)).to(f"cuda:{gpu}") for k in range(n)]
    torch.cuda.synchronize()
    for k in range(n):
        tr.read_image_dev(dev[k].data_ptr(), COLS, 100.0 + k * float(g["dt"]), bool(g["pub"][k]))
    tr.sync()
    k = n - 1
    assert np.array_equal(tr.ids, g[f"f{k}_ids"]) and np.array_equal(tr.track_cnt, g[f"f{k}_track_cnt"])
    assert np.abs(tr.cur_pts - g[f"f{k}_cur_pts"]).max() < 5e-3
    tr.close()


# ---- BASELINE.json configs 3 and 4 shapes (larger frames, more features) against the live cv2 oracle -------------
@pytest.mark.parametrize("scale,max_cnt,n_frames", [(1.5, 500, 3), (3.0, 2000, 2)])
def test_tracker_large_configs_match_cv2_oracle(lfvio_mod, gpu, scale, max_cnt, n_frames):
    pytest.importorskip("cv2")
    from lfvio.config import scaled_camera

    cam, cfg = scaled_camera(scale), tracker_defaults(scale, max_cnt)
    rows, cols = cam["image_height"], cam["image_width"]
    src = synth.FrameSource(rows, cols, seed=5, margin=int(160 * scale))
    dev_tr = lfvio_mod.FeatureTracker(cam, cfg, device=gpu)
    ref_tr = fe.FeatureTrackerOracle(cam, cfg, "cv2")
    for k in range(n_frames):
        img = src.frame(2 * k)
        for t in (dev_tr, ref_tr):
            t.read_image(img, 3.0 + k / 30.0, True)
            t.update_ids()
        assert len(dev_tr.ids) == len(ref_tr.ids), f"frame {k}: {len(dev_tr.ids)} vs {len(ref_tr.ids)} tracks"
        assert np.array_equal(dev_tr.ids, ref_tr.ids), f"frame {k}: ids"
        assert np.array_equal(dev_tr.track_cnt, ref_tr.track_cnt), f"frame {k}: track_cnt"
        assert np.abs(dev_tr.cur_pts - ref_tr.cur_pts).max() < 5e-3, f"frame {k}: cur_pts"
        assert np.array_equal(dev_tr.mask, ref_tr.mask), f"frame {k}: mask"
    assert len(dev_tr.ids) > 0.9 * max_cnt
    dev_tr.close()
