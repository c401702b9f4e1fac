"""Synthetic initialisation scenario shared by the oracle and GPU tests of the estimator initialisation (SURVEY.md §8(f)
rank 4): the first F frames of a lfvio.violoop.World (hand-held trajectory, feature tracks as unit bearings, 200 Hz IMU)."""
from __future__ import annotations

import numpy as np

from lfvio import synthwin, violoop
from oracle import prestep as OP


class InitScenario:
    def __init__(self, seed=1, F=11, pixel_noise=0.0, td=0.0, frame_dt=0.1, max_feat=150):
        E = synthwin.ESTIMATOR
        old_td = E["td"]
        E["td"] = td
        try:
            self.world = violoop.World(seed=seed, n_frames=F, max_feat=max_feat, pixel_noise=pixel_noise, frame_dt=frame_dt)
        finally:
            E["td"] = old_td
        w = self.world
        self.F = F
        self.ric, self.tic, self.g_norm = w.ric, w.tic, float(np.linalg.norm(w.G))
        # true IMU and camera poses
        self.P = np.array([w.truth(k)[0] for k in range(F)])
        self.Rb = np.array([w.truth(k)[1] for k in range(F)])
        self.V = np.array([w.truth(k)[2] for k in range(F)])
        self.Rc = np.array([self.Rb[k] @ self.ric for k in range(F)])
        self.Pc = np.array([self.P[k] + self.Rb[k] @ self.tic for k in range(F)])
        # feature tracks in SFMFeature layout: observations per feature in ascending frame order
        ids = sorted({fid for fr in w.frames for fid in fr})
        self.ids = ids
        first, frames, pts = [0], [], []
        for fid in ids:
            for k in range(F):
                if fid in w.frames[k]:
                    frames.append(k)
                    pts.append(w.frames[k][fid][0])
            first.append(len(frames))
        self.obs_first, self.obs_frame = np.array(first, np.int32), np.array(frames, np.int32)
        self.obs_pts = np.array(pts, float)

    def corres(self, i, j):
        """FeatureManager::getCorresponding(i, j) (feature_manager.cpp:118-143): bearings of the features seen in both."""
        out = []
        for f in range(len(self.ids)):
            o = range(self.obs_first[f], self.obs_first[f + 1])
            fr = [self.obs_frame[k] for k in o]
            if fr and fr[0] <= i and fr[-1] >= j:
                out.append(np.concatenate([self.obs_pts[o[0] + i - fr[0]], self.obs_pts[o[0] + j - fr[0]]]))
        return np.array(out)

    def preintegrations(self, bg=None, ba=None):
        """IntegrationOracle per interval (k-1, k], linearised at the given biases (default: zero, as before initialisation)."""
        w = self.world
        out = []
        for k in range(1, self.F):
            a0, g0, dt, acc, gyr = w.imu[k]
            ib = OP.IntegrationOracle(a0, g0, np.zeros(3) if ba is None else ba, np.zeros(3) if bg is None else bg, w.noise)
            for s in range(len(dt)):
                ib.push_back(dt[s], acc[s], gyr[s])
            out.append(ib)
        return out
