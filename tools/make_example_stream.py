"""Writes the input files of examples/vio_from_files.cpp from the synthetic world the tests use (lfvio.violoop.World):
imu.txt, features.txt, calib.txt and truth_tum.txt (ground truth in the same TUM layout, for an ATE evaluation).

  python tools/make_example_stream.py OUT_DIR [--frames 60] [--seed 2]
  ./vio_from_files OUT_DIR/imu.txt OUT_DIR/features.txt OUT_DIR/result_tum.txt OUT_DIR/calib.txt 0.0
"""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "lf-vio_b200"))

from lfvio import synthwin, violoop  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("out")
    ap.add_argument("--frames", type=int, default=60)
    ap.add_argument("--seed", type=int, default=2)
    a = ap.parse_args()
    os.makedirs(a.out, exist_ok=True)
    E = synthwin.ESTIMATOR
    old_td, E["td"] = E["td"], 0.0           # the stream is cut exactly at the image stamps: td = 0
    try:
        w = violoop.World(seed=a.seed, n_frames=a.frames, pixel_noise=0.1)
    finally:
        E["td"] = old_td
    with open(os.path.join(a.out, "imu.txt"), "w") as f:
        a0, g0 = w.imu[1][0], w.imu[1][1]
        f.write("%.9f %s %s\n" % (w.times[0], " ".join("%.12g" % v for v in a0), " ".join("%.12g" % v for v in g0)))
        for k in range(1, w.n_frames):
            _, _, dt, acc, gyr = w.imu[k]
            n = len(dt)
            for j in range(n):
                t = w.times[k] if j == n - 1 else w.times[k - 1] + w.frame_dt * (j + 1) / n
                f.write("%.9f %s %s\n" % (t, " ".join("%.12g" % v for v in acc[j]), " ".join("%.12g" % v for v in gyr[j])))
    with open(os.path.join(a.out, "features.txt"), "w") as f:
        for k in range(w.n_frames):
            obs = w.frames[k]
            f.write("F %.9f %d\n" % (w.times[k], len(obs)))
            for fid, (b, vel, row) in obs.items():
                f.write("%d %.9g %.9g %.9g 0 %.6g %.9g %.9g %.9g\n" % (fid, b[0], b[1], b[2], row, vel[0], vel[1], vel[2]))
    with open(os.path.join(a.out, "calib.txt"), "w") as f:
        f.write(" ".join("%.12g" % v for v in np.concatenate([w.ric.ravel(), w.tic, w.noise, [np.linalg.norm(w.G)]])) + "\n")
    with open(os.path.join(a.out, "truth_tum.txt"), "w") as f:
        for k in range(w.n_frames):
            P, R, _ = w.truth(k)
            q = synthwin.rot_to_quat(R)
            f.write("%.12f %.12f %.12f %.12f %.12f %.12f %.12f %.12f\n" % (w.times[k], P[0], P[1], P[2], q[1], q[2], q[3], q[0]))
    print("wrote imu.txt, features.txt, calib.txt, truth_tum.txt to", a.out)


if __name__ == "__main__":
    main()
