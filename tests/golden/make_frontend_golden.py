"""Generates tests/golden/frontend_*.npz from the oracle of record: in-container Python cv2 (4.13.0) called with
the reference's arguments, driven by the restated FeatureTracker control flow (oracle/frontend.py, backend "cv2").

Run from the repo root in the build container:  python tests/golden/make_frontend_golden.py
The input frames are NOT stored: they are regenerated from a seed with integer-only numpy arithmetic
(lf-vio_b200/lfvio/synth.py); their CRC32s are stored and asserted by the tests before anything is compared.
"""
import os
import sys

import cv2
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "lf-vio_b200"))

from lfvio import synth  # noqa: E402
from lfvio.config import PAL_UP, tracker_defaults  # noqa: E402
from oracle import frontend as fe  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))

SEQUENCES = {
    # name: (seed, n_frames, pub pattern, FrameSource kwargs, dt)
    "seq_slow": (1, 8, [1] * 8, dict(), 1.0 / 15),
    "seq_fast": (2, 10, [1, 0, 1, 1, 0, 1, 0, 1, 1, 1], dict(rot_per_frame=0.004, shift_per_frame=(6.5, -4.2)), 1.0 / 15),
}


def annulus_mask(rows, cols, cfg):
    m = np.zeros((rows, cols), np.uint8)
    cv2.circle(m, (int(cfg["center_x"]), int(cfg["center_y"])), cfg["max_r"], 255, -1)
    cv2.circle(m, (int(cfg["center_x"]), int(cfg["center_y"])), cfg["min_r"], 0, -1)
    return m


def main():
    cam, cfg = PAL_UP, tracker_defaults()
    rows, cols = cam["image_height"], cam["image_width"]
    print("cv2", cv2.__version__)
    # ---- stage-level vectors on one frame pair --------------------------------------------------
    src = synth.FrameSource(rows, cols, seed=7)
    f0, f1 = src.frame(0), src.frame(3)
    c0 = cv2.createCLAHE(3.0, (8, 8)).apply(f0)
    c1 = cv2.createCLAHE(3.0, (8, 8)).apply(f1)
    mask = annulus_mask(rows, cols, cfg)
    eig = cv2.cornerMinEigenVal(c0, 3, ksize=3)
    pts = cv2.goodFeaturesToTrack(c0, 150, 0.01, 20.0, mask=mask).reshape(-1, 2)
    pts_nomask = cv2.goodFeaturesToTrack(c0, 500, 0.01, 10.0).reshape(-1, 2)
    extra = np.array([[3.2, 4.1], [1276.5, 955.2], [10, 950], [1270, 12], [640.5, 3.3], [0.4, 480.2], [25.7, 30.2],
                      [1279.0, 959.0], [0.0, 0.0], [639.5, 479.5]], np.float32)
    lk_in = np.concatenate([pts, extra]).astype(np.float32)
    lk_out, lk_st, _ = cv2.calcOpticalFlowPyrLK(c0, c1, lk_in.reshape(-1, 1, 2), None, winSize=(41, 41), maxLevel=3)
    pyr = [c0]
    for _ in range(3):
        pyr.append(cv2.pyrDown(pyr[-1]))
    np.savez_compressed(
        os.path.join(OUT, "frontend_stages.npz"),
        seed=7, frame_idx=np.array([0, 3]), frame_crc=np.array([synth.crc(f0), synth.crc(f1)], np.uint64),
        clahe_crc=np.array([synth.crc(c0), synth.crc(c1)], np.uint64),
        pyr_crc=np.array([synth.crc(p) for p in pyr], np.uint64),
        eig_crc=np.array([synth.crc(eig)], np.uint64), eig_max=np.float32(eig.max()),
        eig_samples_idx=np.arange(0, rows * cols, 9973), eig_samples=eig.ravel()[::9973].copy(),
        gftt_masked=pts, gftt_nomask=pts_nomask, lk_in=lk_in, lk_out=lk_out.reshape(-1, 2), lk_status=lk_st.reshape(-1),
        cv2_version=cv2.__version__,
    )
    # ---- whole-tracker sequences ----------------------------------------------------------------
    for name, (seed, n, pubs, kw, dt) in SEQUENCES.items():
        src = synth.FrameSource(rows, cols, seed=seed, **kw)
        tr = fe.FeatureTrackerOracle(cam, cfg, backend="cv2")
        out = {"seed": seed, "n_frames": n, "pub": np.array(pubs), "dt": dt, "cv2_version": cv2.__version__}
        crcs, mask_crcs = [], []
        for k in range(n):
            img = src.frame(k)
            crcs.append(synth.crc(img))
            tr.read_image(img, 100.0 + k * dt, bool(pubs[k]))
            pre_ids = tr.ids.copy()
            tr.update_ids()
            s = tr.snapshot()
            out[f"f{k}_pre_ids"] = pre_ids.astype(np.int32)
            out[f"f{k}_cur_pts"] = s["cur_pts"]
            out[f"f{k}_ids"] = s["ids"].astype(np.int32)
            out[f"f{k}_track_cnt"] = s["track_cnt"].astype(np.int32)
            out[f"f{k}_cur_un_pts_3d"] = s["cur_un_pts_3d"]
            out[f"f{k}_pts_velocity_3d"] = s["pts_velocity_3d"]
            mask_crcs.append(synth.crc(tr.mask) if tr.mask is not None else 0)
            print(name, k, "n =", len(s["ids"]), "new =", int((s["track_cnt"] == 1).sum()), "n_id =", tr.n_id)
        out["frame_crc"] = np.array(crcs, np.uint64)
        out["mask_crc"] = np.array(mask_crcs, np.uint64)
        np.savez_compressed(os.path.join(OUT, f"frontend_{name}.npz"), **out)


if __name__ == "__main__":
    main()
