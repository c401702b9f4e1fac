"""TEST INFRASTRUCTURE ONLY — Python restatement of the measurement synchronisation of the reference's estimator node
(vins_estimator/src/estimator_node.cpp): imu_callback (:136-162, out-of-order samples dropped), feature_callback (:164-176,
first message skipped), getMeasurements (:100-134) and the IMU hand-over of process() (:219-262: samples up to the image time
stamp + td are passed on with their own dt, the first one after it is linearly interpolated at the image time)."""
from __future__ import annotations

import numpy as np


class EstimatorNodeOracle:
    def __init__(self, td=0.0):
        self.td = td
        self.imu_buf, self.feature_buf = [], []
        self.last_imu_t, self.current_time = 0.0, -1.0
        self.init_feature = False
        self.sum_of_wait = 0
        self.frames = []          # per processed image: (stamp, [(dt, acc, gyr), ...]) = the processIMU calls before processImage

    def imu_callback(self, t, acc, gyr):
        if t <= self.last_imu_t:
            return
        self.last_imu_t = t
        self.imu_buf.append((t, np.asarray(acc, float), np.asarray(gyr, float)))

    def feature_callback(self, t):
        if not self.init_feature:
            self.init_feature = True
            return
        self.feature_buf.append(t)

    def process(self):
        n = 0
        while True:
            if not self.imu_buf or not self.feature_buf:
                break
            if not (self.imu_buf[-1][0] > self.feature_buf[0] + self.td):
                self.sum_of_wait += 1
                break
            if not (self.imu_buf[0][0] < self.feature_buf[0] + self.td):
                self.feature_buf.pop(0)
                continue
            img_t0 = self.feature_buf.pop(0)
            imus = []
            while self.imu_buf[0][0] < img_t0 + self.td:
                imus.append(self.imu_buf.pop(0))
            imus.append(self.imu_buf[0])
            calls = []
            a, g = np.zeros(3), np.zeros(3)
            for (t, acc, gyr) in imus:
                img_t = img_t0 + self.td
                if t <= img_t:
                    if self.current_time < 0:
                        self.current_time = t
                    dt = t - self.current_time
                    self.current_time = t
                    a, g = acc.copy(), gyr.copy()
                    calls.append((dt, a.copy(), g.copy()))
                else:
                    dt_1, dt_2 = img_t - self.current_time, t - img_t
                    self.current_time = img_t
                    w1, w2 = dt_2 / (dt_1 + dt_2), dt_1 / (dt_1 + dt_2)
                    a, g = w1 * a + w2 * acc, w1 * g + w2 * gyr
                    calls.append((dt_1, a.copy(), g.copy()))
            self.frames.append((img_t0, calls))
            n += 1
        return n
