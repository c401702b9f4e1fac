"""Synthetic IMU streams and multi-view bearing tracks for the pre-integration / triangulation tests."""
import numpy as np

from lfvio import synthwin

NOISE = (0.02, 0.01, 0.04, 0.001)  # acc_n gyr_n acc_w gyr_w (config/mindvision/mindvision.yaml:138-141)
G = np.array([0.0, 0.0, 9.81007])


def imu_stream(seed, n_intervals=10, frame_dt=0.1, rate=200.0, t0=5.0, noisy=True):
    """Returns dict(dt, acc, gyr (flat sample arrays), intervals=[(first, count, acc0, gyr0)], ba, bg, traj, times)."""
    rng = np.random.default_rng(seed)
    traj = synthwin.Trajectory(seed)
    ba, bg = rng.normal(0, 0.02, 3), rng.normal(0, 0.002, 3)
    n_sub = int(round(frame_dt * rate))
    h = frame_dt / n_sub

    def meas(t):
        acc = traj.rot(t).T @ (traj.acc(t) + G) + ba
        gyr = traj.omega_body(t) + bg
        if noisy:
            acc = acc + rng.normal(0, NOISE[0], 3)
            gyr = gyr + rng.normal(0, NOISE[1], 3)
        return acc, gyr

    dt, acc, gyr, intervals = [], [], [], []
    for i in range(n_intervals):
        ts = t0 + i * frame_dt
        a0, g0 = meas(ts)
        first = len(dt)
        for k in range(1, n_sub + 1):
            a, g = meas(ts + k * h)
            dt.append(h * (1.0 + (0.05 * rng.uniform(-1, 1) if noisy else 0.0)))  # jittered stamps
            acc.append(a)
            gyr.append(g)
        intervals.append((first, n_sub, a0, g0))
    return dict(dt=np.array(dt), acc=np.array(acc), gyr=np.array(gyr), intervals=intervals, ba=ba, bg=bg, traj=traj,
                times=t0 + frame_dt * np.arange(n_intervals + 1))


def tracks(seed, n_frames=11, n_feat=200, noise=0.0, behind_frac=0.0):
    """Camera poses on a smooth trajectory and bearing tracks of random landmarks.
    Returns dict(Ps, Rs, tic, ric, start_frame, obs_first, obs_pts, true_depth)."""
    rng = np.random.default_rng(100 + seed)
    traj = synthwin.Trajectory(seed)
    E = synthwin.ESTIMATOR
    ric = np.array(E["extrinsicRotation"]).reshape(3, 3)
    tic = np.array(E["extrinsicTranslation"])
    times = 5.0 + 0.1 * np.arange(n_frames)
    Rs = np.array([traj.rot(t) for t in times])
    Ps = np.array([traj.pos(t) for t in times])
    start, first, pts, depth = [], [0], [], []
    for l in range(n_feat):
        s = int(rng.integers(0, n_frames - 1))
        K = int(rng.integers(1, n_frames - s + 1))
        ang, rad, hgt = rng.uniform(0, 2 * np.pi), rng.uniform(2.0, 9.0), rng.uniform(-2.0, 3.0)
        pw = np.array([rad * np.cos(ang), rad * np.sin(ang), hgt])
        if rng.uniform() < behind_frac:  # inconsistent track -> negative depth -> INIT_DEPTH
            pw = Ps[s] - (pw - Ps[s])
        for k in range(K):
            f = s + k
            pc = ric.T @ (Rs[f].T @ (pw - Ps[f]) - tic)
            if k == 0:
                d0 = np.linalg.norm(pc)
            b = pc / np.linalg.norm(pc) + rng.normal(0, noise, 3)
            if behind_frac and k > 0 and np.array_equal(pw, Ps[s] - (pw - Ps[s])):
                b = -b
            b = b / np.linalg.norm(b)
            pts.append(b.astype(np.float32).astype(np.float64))  # bearings travel as float32
        start.append(s)
        first.append(len(pts))
        depth.append(d0)
    return dict(Ps=Ps, Rs=Rs, tic=tic, ric=ric, start_frame=np.array(start, np.int32), obs_first=np.array(first, np.int32),
                obs_pts=np.array(pts), true_depth=np.array(depth))
