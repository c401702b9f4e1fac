"""Closed-loop sliding-window VIO over a synthetic feature stream (tests / bench only).

A Python restatement of the NON_LINEAR branch of Estimator::processImage (vins_estimator/src/estimator.cpp:122-219):
processIMU -> f_manager.addFeature -> triangulate -> optimization() (solve + gauge fix + MARGIN_OLD prior) ->
slideWindow() (estimator.cpp:1011-1075) with FeatureManager::removeBackShiftDepth (feature_manager.cpp:271-310) and
removeFailures (:133-143).  Every frame is treated as a keyframe (MARGIN_OLD), the image front end is bypassed (the world
hands out noisy unit bearings, as feature_tracker_node publishes them) and the initialisation (initialStructure) is
replaced by a perturbed ground-truth bootstrap of the first window.  The numeric steps are injected (`Ops`): the device
library or the CPU oracle — the orchestration is the same code for both, so their trajectories can be compared (ATE).
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

from . import backend as B
from . import prestep as PS
from . import synthwin


# ------------------------------------------------------------------------------------------------------------------
# synthetic world: trajectory, IMU stream, feature tracks
# ------------------------------------------------------------------------------------------------------------------
class HandheldTrajectory(synthwin.Trajectory):
    """Same parametrisation as synthwin.Trajectory with hand-held-like excitation (accelerations of 1-3 m/s^2, body
    rates up to ~0.5 rad/s): scale, accelerometer bias and tilt are observable inside a one-second window."""

    def __init__(self, seed: int):
        super().__init__(seed)
        rng = np.random.default_rng(77 + seed)
        self.a = rng.uniform(0.45, 0.9, 3)
        self.w = rng.uniform(1.3, 2.3, 3)

    def rot(self, t):
        return synthwin.rot_xyz(0.35 * t + 0.3 * np.sin(0.9 * t + self.ph[3]), 0.2 * np.sin(1.3 * t + self.ph[4]),
                                0.2 * np.cos(1.1 * t + self.ph[5]))


class World:
    def __init__(self, seed=1, n_frames=60, max_feat=150, frame_dt=0.1, imu_rate=200.0, pixel_noise=0.3, t0=5.0,
                 handheld=True):
        E = synthwin.ESTIMATOR
        self.E = E
        rng = np.random.default_rng(5000 + seed)
        self.traj = HandheldTrajectory(seed) if handheld else synthwin.Trajectory(seed)
        self.G = np.array([0.0, 0.0, E["g_norm"]])
        self.ric = np.array(E["extrinsicRotation"], float).reshape(3, 3)
        self.tic = np.array(E["extrinsicTranslation"], float)
        self.td_true = E["td"]
        self.noise = (E["acc_n"], E["gyr_n"], E["acc_w"], E["gyr_w"])
        self.ba, self.bg = rng.normal(0, 0.02, 3), rng.normal(0, 0.002, 3)
        self.frame_dt, self.n_frames = frame_dt, n_frames
        self.times = t0 + frame_dt * np.arange(n_frames)
        n_sub = int(round(frame_dt * imu_rate))
        h = frame_dt / n_sub

        def meas(t):
            acc = self.traj.rot(t).T @ (self.traj.acc(t) + self.G) + self.ba + rng.normal(0, E["acc_n"], 3)
            gyr = self.traj.omega_body(t) + self.bg + rng.normal(0, E["gyr_n"], 3)
            return acc, gyr

        # imu[k] = samples of the interval (frame k-1, frame k]: (acc0, gyr0, dt[n], acc[n,3], gyr[n,3])
        self.imu = [None]
        for k in range(1, n_frames):
            a0, g0 = meas(self.times[k - 1])
            acc, gyr = zip(*[meas(self.times[k - 1] + (j + 1) * h) for j in range(n_sub)])
            self.imu.append((a0, g0, np.full(n_sub, h), np.array(acc), np.array(gyr)))
        # feature tracks
        sigma = pixel_noise / E["focal_length"]
        self.frames = []
        active = {}  # id -> [world point, frames left, previous bearing]
        next_id = 0
        for k, t in enumerate(self.times):
            for fid in [f for f, v in active.items() if v[1] <= 0]:
                del active[fid]
            while len(active) < max_feat:
                ang, rad, hgt = rng.uniform(0, 2 * np.pi), rng.uniform(3.0, 8.0), rng.uniform(-2.0, 3.0)
                pw = self.traj.pos(t) + np.array([rad * np.cos(ang), rad * np.sin(ang), hgt])
                active[next_id] = [pw, int(rng.integers(3, 26)), None]
                next_id += 1
            obs = {}
            ts = t + self.td_true  # the image was really exposed td later than its stamp
            Rw, Pw = self.traj.rot(ts), self.traj.pos(ts)
            Rp, Pp = self.traj.rot(ts - frame_dt), self.traj.pos(ts - frame_dt)

            def see(pw, R_, P_):
                pc = self.ric.T @ (R_.T @ (pw - P_) - self.tic)
                b = pc / np.linalg.norm(pc) + rng.normal(0, sigma, 3)
                return (b / np.linalg.norm(b)).astype(np.float32).astype(np.float64)  # bearings travel as float32

            for fid, v in active.items():
                if v[2] is None:
                    # the tracker publishes a feature from its second frame on (track_cnt > 1,
                    # feature_tracker_node.cpp:142), so the first published observation already has a velocity
                    v[2] = see(v[0], Rp, Pp)
                b = see(v[0], Rw, Pw)
                vel = ((b - v[2]) / frame_dt).astype(np.float32).astype(np.float64)
                obs[fid] = (b, vel, 480.0 + 300.0 * b[1])
                v[2] = b
                v[1] -= 1
            self.frames.append(obs)

    def truth(self, k):
        t = self.times[k]
        return self.traj.pos(t), self.traj.rot(t), self.traj.vel(t)


# ------------------------------------------------------------------------------------------------------------------
# numeric steps
# ------------------------------------------------------------------------------------------------------------------
@dataclass
class Ops:
    name: str
    optimize_full: callable   # (WindowProblem, flag) -> (SolveSummary, PriorData); updates the problem's state in place
    preintegrate: callable    # (intervals, dt, acc, gyr, noise, g) -> ([ImuFactor], state n x 15)
    triangulate: callable     # (Ps, Rs, tic, ric, start, first, pts, depth, init_depth) -> depth


def device_ops(lfvio_mod, window, device=0) -> Ops:
    return Ops("device", lambda w, flag: window.optimize_full(w, flag),
               lambda iv, dt, acc, gyr, nz, g: lfvio_mod.preintegrate(iv, dt, acc, gyr, nz, g, device=device),
               lambda *a: lfvio_mod.triangulate(*a, device=device))


def _quat_to_rot(q_xyzw):
    return synthwin.quat_to_rot(np.array([q_xyzw[3], q_xyzw[0], q_xyzw[1], q_xyzw[2]]))


def _rot_to_qxyzw(R):
    q = synthwin.rot_to_quat(R)
    return np.array([q[1], q[2], q[3], q[0]])


@dataclass
class Feature:
    fid: int
    start: int
    obs: list = field(default_factory=list)  # (point, velocity, td, row)
    depth: float = -1.0


class VioLoop:
    """WINDOW_SIZE = F - 1 keyframes + the newest frame, all state in the flat layout of include/lfvio.h."""

    def __init__(self, world: World, ops: Ops, opts, F=11, perturb=0.3, seed=0):
        self.w, self.ops, self.opts, self.F = world, ops, opts, F
        rng = np.random.default_rng(9000 + seed)
        self.pose, self.sb = np.zeros((F, 7)), np.zeros((F, 9))
        ba0, bg0 = world.ba + rng.normal(0, 0.01, 3), world.bg + rng.normal(0, 0.001, 3)
        for f in range(F):  # bootstrap: perturbed ground truth (replaces initialStructure)
            P, R, V = world.truth(f)
            dth = perturb * rng.normal(0, 0.01, 3) * (f > 0)
            self.pose[f, :3] = P + perturb * rng.normal(0, 0.03, 3) * (f > 0)
            self.pose[f, 3:] = _rot_to_qxyzw(R @ synthwin.rot_xyz(dth[2], dth[1], dth[0]))
            self.sb[f, :3] = V + perturb * rng.normal(0, 0.05, 3)
            self.sb[f, 3:6], self.sb[f, 6:9] = ba0, bg0
        self.ex = np.concatenate([world.tic, _rot_to_qxyzw(world.ric)])
        self.td = world.td_true + 0.003 * perturb
        self.features: list[Feature] = []
        self.imu = []  # ImuFactor for the pairs (i, i+1), i = 0 .. F-2
        for f in range(1, F):
            self.imu.append(self._integrate(f, f - 1, propagate=False)[0])
        for f in range(F):
            self._add_frame_obs(f, f)
        self.prior = None
        self.k_next = F  # next world frame to process
        self.estimates = {}  # world frame -> position estimate when it was the newest frame
        self.iterations = []

    # -- pieces of processImage ------------------------------------------------------------------------------
    def _integrate(self, k_world, slot_i, propagate):
        a0, g0, dt, acc, gyr = self.w.imu[k_world]
        j = min(slot_i + 1, self.F - 1)
        iv = PS.make_interval(0, len(dt), slot_i, a0, g0, self.sb[j, 3:6], self.sb[j, 6:9],
                              self.pose[j, :3] if propagate else None, self.sb[j, :3] if propagate else None,
                              _quat_to_rot(self.pose[j, 3:]) if propagate else None)
        out, state = self.ops.preintegrate([iv], dt, acc, gyr, self.w.noise, self.w.G)
        return out[0], state[0]

    def _add_frame_obs(self, k_world, slot):
        by_id = {f.fid: f for f in self.features}
        for fid, (b, vel, row) in self.w.frames[k_world].items():
            ft = by_id.get(fid)
            if ft is None:
                ft = Feature(fid, slot)
                self.features.append(ft)
            elif ft.start + len(ft.obs) != slot:
                continue  # (cannot happen: tracks are contiguous)
            # cur_td = the td used when the image was paired with the IMU stream (feature_manager.cpp:54).  The world cuts
            # the IMU intervals at the image STAMPS, i.e. it pairs with td = 0; the estimated td then converges to td_true.
            ft.obs.append((b, vel, 0.0, row))

    def _triangulate(self):
        F = self.F
        Ps = self.pose[:, :3]
        Rs = np.array([_quat_to_rot(self.pose[f, 3:]) for f in range(F)])
        start = np.array([f.start for f in self.features], np.int32)
        first = np.concatenate([[0], np.cumsum([len(f.obs) for f in self.features])]).astype(np.int32)
        pts = np.array([o[0] for f in self.features for o in f.obs])
        depth = np.array([f.depth for f in self.features])
        out = self.ops.triangulate(Ps, Rs, self.ex[:3], _quat_to_rot(self.ex[3:]), start, first, pts, depth, 5.0)
        for f, d in zip(self.features, out):
            f.depth = float(d)

    def _problem(self):
        ws = self.F - 1
        proj, depths, used = [], [], []
        for ft in self.features:
            if not (len(ft.obs) >= 2 and ft.start < ws - 2):  # estimator.cpp:730
                continue
            idx = len(depths)
            depths.append(1.0 / ft.depth)
            used.append(ft)
            p0 = ft.obs[0]
            for k in range(1, len(ft.obs)):
                pk = ft.obs[k]
                pf = B.ProjFactor()
                pf.frame_i, pf.frame_j, pf.landmark = ft.start, ft.start + k, idx
                pf.pts_i[:], pf.pts_j[:] = p0[0].tolist(), pk[0].tolist()
                pf.vel_i[:], pf.vel_j[:] = p0[1].tolist(), pk[1].tolist()
                pf.td_i, pf.td_j = p0[2], pk[2]
                pf.row_i, pf.row_j = p0[3], pk[3]
                proj.append(pf)
        imu = []
        for i, f in enumerate(self.imu):
            f.frame_i = i
            imu.append(f)
        return B.WindowProblem(self.F, self.pose, self.sb, self.ex, np.array(depths), self.td, proj, imu, self.prior, self.opts), used

    def _optimize(self):
        self._triangulate()
        prob, used = self._problem()
        summ, prior = self.ops.optimize_full(prob, B.MARGIN_OLD)
        self.pose[:], self.sb[:], self.ex[:], self.td = prob.pose, prob.sb, prob.ex, float(prob.td[0])
        # FeatureManager::setDepth (feature_manager.cpp:138-156).  LF-VIO sets solve_flag = 1 in BOTH branches, so
        # removeFailures (:158-167) never erases anything: a negative depth simply fails the `estimated_depth > 0` test
        # of the next triangulate() and is re-triangulated.
        for ft, lam in zip(used, prob.feat):
            ft.depth = 1.0 / lam
        self.prior = prior
        self.iterations.append(int(summ.iterations))

    def _slide_window(self):
        F = self.F
        R0, P0 = _quat_to_rot(self.pose[0, 3:]), self.pose[0, :3].copy()
        ric, tic = _quat_to_rot(self.ex[3:]), self.ex[:3]
        self.pose[:-1], self.sb[:-1] = self.pose[1:].copy(), self.sb[1:].copy()  # newest slot keeps a copy (:1040-1045)
        self.imu = self.imu[1:]
        R1, P1 = _quat_to_rot(self.pose[0, 3:]), self.pose[0, :3]
        marg_R, marg_P, new_R, new_P = R0 @ ric, P0 + R0 @ tic, R1 @ ric, P1 + R1 @ tic
        keep = []
        for ft in self.features:  # removeBackShiftDepth
            if ft.start != 0:
                ft.start -= 1
                keep.append(ft)
                continue
            uv_i = ft.obs.pop(0)[0]
            if len(ft.obs) < 2:
                continue
            pts_j = new_R.T @ (marg_R @ (uv_i * ft.depth) + marg_P - new_P)
            dep = float(np.linalg.norm(pts_j))
            ft.depth = dep if dep > 0 else 5.0
            keep.append(ft)
        self.features = keep  # (removeFailures is a no-op in LF-VIO, see _optimize)

    # -- one camera frame ------------------------------------------------------------------------------------
    def step(self):
        """Optimise the current window, slide, then take in the next world frame.  Returns False when the world is out."""
        self._optimize()
        newest_world = self.k_next - 1
        self.estimates[newest_world] = (self.pose[-1, :3].copy(), _quat_to_rot(self.pose[-1, 3:]))
        if self.k_next >= self.w.n_frames:
            return False
        self._slide_window()
        f, state = self._integrate(self.k_next, self.F - 2, propagate=True)
        self.imu.append(f)
        self.pose[-1, :3], self.sb[-1, :3] = state[0:3], state[3:6]
        self.pose[-1, 3:] = _rot_to_qxyzw(state[6:15].reshape(3, 3))
        self._add_frame_obs(self.k_next, self.F - 1)
        self.k_next += 1
        return True

    def run(self):
        while self.step():
            pass
        return self


class VioLoopStore(VioLoop):
    """The same loop with the FeatureManager on the device (lfvio.FeatureStore): observations go in once per frame, the
    triangulation, the factor list, setDepth and removeBackShiftDepth run where the tracks live, and the solve consumes the
    list without it ever existing on the host (lfvio_window_optimize_full_fm)."""

    def __init__(self, world: World, window, store, opts, F=11, perturb=0.3, seed=0):
        from . import _api  # (not importable in the library-free mode; this class needs the device)
        self.window, self.store = window, store
        store.clear()
        ops = Ops("device-store", None, lambda iv, dt, acc, gyr, nz, g: _api.preintegrate(iv, dt, acc, gyr, nz, g), None)
        super().__init__(world, ops, opts, F, perturb, seed)

    def _add_frame_obs(self, k_world, slot):
        obs = self.w.frames[k_world]
        ids = np.array(sorted(obs), np.int32)
        f8 = np.zeros((len(ids), 8))
        for r, i in enumerate(ids):
            b, vel, row = obs[int(i)]
            f8[r] = [b[0], b[1], b[2], 0.0, row, vel[0], vel[1], vel[2]]
        self.store.add_frame(slot, ids, f8, 0.0)

    def _triangulate(self):
        Rs = np.array([_quat_to_rot(self.pose[f, 3:]) for f in range(self.F)])
        self.store.triangulate(self.pose[:, :3], Rs, self.ex[:3], _quat_to_rot(self.ex[3:]))

    def _optimize(self):
        self._triangulate()
        imu = []
        for i, f in enumerate(self.imu):
            f.frame_i = i
            imu.append(f)
        prob = B.WindowProblem(self.F, self.pose, self.sb, self.ex, np.zeros(0), self.td, [], imu, self.prior, self.opts)
        summ, prior = self.window.optimize_full_fm(self.store, prob, B.MARGIN_OLD)
        self.pose[:], self.sb[:], self.ex[:], self.td = prob.pose, prob.sb, prob.ex, float(prob.td[0])
        self.prior = prior
        self.iterations.append(int(summ.iterations))

    def _slide_window(self):
        R0, P0 = _quat_to_rot(self.pose[0, 3:]), self.pose[0, :3].copy()
        ric, tic = _quat_to_rot(self.ex[3:]), self.ex[:3]
        self.pose[:-1], self.sb[:-1] = self.pose[1:].copy(), self.sb[1:].copy()
        self.imu = self.imu[1:]
        R1, P1 = _quat_to_rot(self.pose[0, 3:]), self.pose[0, :3]
        self.store.remove_back_shift_depth(R0 @ ric, P0 + R0 @ tic, R1 @ ric, P1 + R1 @ tic)


def ate_rmse(est_pos: np.ndarray, true_pos: np.ndarray) -> float:
    """Absolute trajectory error: RMSE after the least-squares SE(3) alignment (Umeyama without scale)."""
    mu_e, mu_t = est_pos.mean(0), true_pos.mean(0)
    H = (est_pos - mu_e).T @ (true_pos - mu_t)
    U, _, Vt = np.linalg.svd(H)
    S = np.diag([1, 1, np.sign(np.linalg.det(Vt.T @ U.T))])
    R = Vt.T @ S @ U.T
    err = (R @ (est_pos - mu_e).T).T + mu_t - true_pos
    return float(np.sqrt((err ** 2).sum(1).mean()))


def trajectory_error(loop: VioLoop):
    ks = sorted(loop.estimates)
    est = np.array([loop.estimates[k][0] for k in ks])
    tru = np.array([loop.w.truth(k)[0] for k in ks])
    return ate_rmse(est, tru), ks
