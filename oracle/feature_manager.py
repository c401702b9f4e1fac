"""TEST INFRASTRUCTURE ONLY — pure-Python restatement of the reference's FeatureManager bookkeeping
(vins_estimator/src/feature_manager.cpp), used to check the C++ host mirror (lf-vio_b200/host/estimator.cpp) on a machine
without a GPU:

  add_feature_check_parallax   :45-93   (key-frame decision; bearing parallax acos(p_i . p_j) * 10, :353-369)
  get_corresponding            :118-143
  remove_back / remove_front   :312-345
  remove_back_shift_depth      :271-310 (depth = range along the bearing, ||.|| — "changed by wz")
  get_feature_count            :36-43

Parity unpinned against a reference binary (no tests / vectors in the reference); pinned by hand-checked small cases in
tests/test_feature_manager_host.py.
"""
from __future__ import annotations

import numpy as np


class Feature:
    def __init__(self, fid, start):
        self.id, self.start, self.pts, self.depth = fid, start, [], -1.0
        self.aux = []  # per observation: (velocity, uv, cur_td) or None — carried along, never read by the bookkeeping

    @property
    def end(self):
        return self.start + len(self.pts) - 1


class FeatureManagerOracle:
    def __init__(self, window_size=10, min_parallax=10.0 / 160.0, init_depth=5.0):
        self.W, self.min_parallax, self.init_depth = window_size, min_parallax, init_depth
        self.feature: list[Feature] = []
        self.last_track_num = 0

    def add_feature_check_parallax(self, frame_count, image: dict) -> bool:
        """image: {feature_id: bearing (3,)} or {feature_id: (bearing, (velocity, uv, cur_td))}, iterated in ascending id
        order (std::map)."""
        self.last_track_num = 0
        for fid in sorted(image):
            f = next((x for x in self.feature if x.id == fid), None)
            if f is None:
                f = Feature(fid, frame_count)
                self.feature.append(f)
            else:
                self.last_track_num += 1
            val = image[fid]
            b, aux = val if isinstance(val, tuple) else (val, None)
            f.pts.append(np.asarray(b, float))
            f.aux.append(aux)
        if frame_count < 2 or self.last_track_num < 20:
            return True
        s, n = 0.0, 0
        for f in self.feature:
            if f.start <= frame_count - 2 and f.end >= frame_count - 1:
                p_i, p_j = f.pts[frame_count - 2 - f.start], f.pts[frame_count - 1 - f.start]
                s += np.arccos(p_i @ p_j) * 10
                n += 1
        if n == 0:
            return True
        return s / n >= self.min_parallax

    def get_corresponding(self, l, r):
        return np.array([np.concatenate([f.pts[l - f.start], f.pts[r - f.start]]) for f in self.feature if f.start <= l and f.end >= r]).reshape(-1, 6)

    def get_feature_count(self):
        return sum(1 for f in self.feature if len(f.pts) >= 2 and f.start < self.W - 2)

    def remove_back(self):
        keep = []
        for f in self.feature:
            if f.start != 0:
                f.start -= 1
            else:
                f.pts.pop(0)
                f.aux.pop(0)
                if not f.pts:
                    continue
            keep.append(f)
        self.feature = keep

    def remove_front(self, frame_count):
        keep = []
        for f in self.feature:
            if f.start == frame_count:
                f.start -= 1
            else:
                j = self.W - 1 - f.start
                if f.end >= frame_count - 1:
                    f.pts.pop(j)
                    f.aux.pop(j)
                    if not f.pts:
                        continue
            keep.append(f)
        self.feature = keep

    def remove_back_shift_depth(self, marg_R, marg_P, new_R, new_P):
        keep = []
        for f in self.feature:
            if f.start != 0:
                f.start -= 1
            else:
                uv_i = f.pts.pop(0)
                f.aux.pop(0)
                if len(f.pts) < 2:
                    continue
                w = marg_R @ (uv_i * f.depth) + marg_P
                d = np.linalg.norm(new_R.T @ (w - new_P))
                f.depth = d if d > 0 else self.init_depth
            keep.append(f)
        self.feature = keep

    def _selected(self):
        return [f for f in self.feature if len(f.pts) >= 2 and f.start < self.W - 2]

    def set_depth(self, x):
        """setDepth / clearDepth (feature_manager.cpp:138-174): estimated_depth = 1 / x over the selected features."""
        for f, v in zip(self._selected(), x):
            f.depth = 1.0 / float(v)

    def depth_vector(self):
        """getDepthVector (feature_manager.cpp:176-190)."""
        return np.array([1.0 / f.depth for f in self._selected()])

    def factor_list(self):
        """The residual-block loop of Estimator::optimization (estimator.cpp:725-772) and para_Feature (vector2double,
        :1044-1046): [(imu_i, imu_j, feature_index, pts_i, pts_j, aux_i, aux_j)] in insertion order, landmark ranges,
        inverse depths."""
        out, first, lam = [], [0], []
        for l, f in enumerate(self._selected()):
            for j in range(1, len(f.pts)):
                out.append((f.start, f.start + j, l, f.pts[0], f.pts[j], f.aux[0], f.aux[j]))
            first.append(len(out))
            lam.append(1.0 / f.depth)
        return out, np.array(first, np.int32), np.array(lam)

    def snapshot(self):
        """[(id, start_frame, n_obs, depth)] in list order."""
        return [(f.id, f.start, len(f.pts), f.depth) for f in self.feature]
