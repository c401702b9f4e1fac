"""Estimator::double2vector / vector2double of the C++ host mirror (pure host math, no device) against the numpy statement of
the reference's gauge re-anchoring (vins_estimator/src/estimator.cpp:532-583, oracle/loop_ops.double2vector): the optimised
window is rotated about gravity so that frame 0 keeps the yaw it had before the solve and shifted so that it keeps its
position; near the Euler singularity (pitch within 1 degree of +-90) the full rotation difference is used instead."""
import ctypes as C
import os

import numpy as np
import pytest

from conftest import ROOT
from lfvio import synthwin


@pytest.fixture(scope="module")
def host():
    lib = C.CDLL(os.path.join(ROOT, "lf-vio_b200", "liblfvio_host.so"))
    lib.lfvio_host_double2vector.argtypes = [C.c_int] + [C.c_void_p] * 10
    return lib


def _yaw_pitch(R):
    y = np.arctan2(R[1, 0], R[0, 0])
    p = np.arctan2(-R[2, 0], R[0, 0] * np.cos(y) + R[1, 0] * np.sin(y))
    return np.degrees(y), np.degrees(p)


def _reference(P0, R0, pose7, sb9):
    """numpy statement of estimator.cpp:532-583."""
    F = len(pose7)
    q = lambda p: synthwin.quat_to_rot(np.array([p[6], p[3], p[4], p[5]]) / np.linalg.norm(p[3:7]))
    y0, p0 = _yaw_pitch(R0)
    R00 = synthwin.quat_to_rot(np.array([pose7[0][6], pose7[0][3], pose7[0][4], pose7[0][5]]))
    y00, p00 = _yaw_pitch(R00)
    a = np.radians(y0 - y00)
    rot = np.array([[np.cos(a), -np.sin(a), 0], [np.sin(a), np.cos(a), 0], [0, 0, 1]])
    if abs(abs(p0) - 90) < 1.0 or abs(abs(p00) - 90) < 1.0:
        rot = R0 @ R00.T
    Rs = np.array([rot @ q(pose7[i]) for i in range(F)])
    Ps = np.array([rot @ (pose7[i][:3] - pose7[0][:3]) + P0 for i in range(F)])
    Vs = np.array([rot @ sb9[i][:3] for i in range(F)])
    return Ps, Rs, Vs


@pytest.mark.parametrize("case", ["generic", "pitch_singular"])
def test_double2vector_gauge_fix(host, case):
    rng = np.random.default_rng(7)
    W, F = 10, 11
    pitch = 0.3 if case == "generic" else np.radians(89.6)
    R_before = np.array([synthwin.rot_xyz(0.8 + 0.05 * i, pitch, 0.1 * i) for i in range(F)])
    P_before = rng.normal(size=(F, 3))
    # the "optimised" window: the true one moved by an arbitrary yaw + translation (the gauge freedom) plus small changes
    yaw = 0.4
    G = synthwin.rot_xyz(yaw, 0, 0)
    pose7, sb9 = np.zeros((F, 7)), rng.normal(size=(F, 9))
    for i in range(F):
        Ri = G @ R_before[i] @ synthwin.rot_xyz(*(1e-3 * rng.normal(size=3)))
        qw = synthwin.rot_to_quat(Ri)
        pose7[i, :3] = G @ P_before[i] + np.array([3.0, -2.0, 0.5]) + 1e-3 * rng.normal(size=3)
        pose7[i, 3:] = [qw[1], qw[2], qw[3], qw[0]]
    ex7 = np.array([0.01, 0.02, 0.03, 0, 0, 0, 1.0])
    P, R, V = P_before.copy(), np.ascontiguousarray(R_before.reshape(F, 9)), np.zeros((F, 3))
    Ba, Bg = np.zeros((F, 3)), np.zeros((F, 3))
    pose_rt, sb_rt = np.zeros((F, 7)), np.zeros((F, 9))
    p = lambda a: a.ctypes.data
    assert host.lfvio_host_double2vector(W, p(P), p(R), p(V), p(Ba), p(Bg), p(pose7), p(sb9), p(ex7), p(pose_rt), p(sb_rt)) == 0
    Ps, Rs, Vs = _reference(P_before[0], R_before[0], pose7, sb9)
    assert np.allclose(P, Ps, atol=1e-12) and np.allclose(R.reshape(F, 3, 3), Rs, atol=1e-12) and np.allclose(V, Vs, atol=1e-12)
    assert np.allclose(Ba, sb9[:, 3:6]) and np.allclose(Bg, sb9[:, 6:9])
    # frame 0 keeps its position exactly and its yaw (generic case) / its whole rotation (singular case)
    assert np.allclose(P[0], P_before[0], atol=1e-12)
    if case == "generic":
        assert abs(_yaw_pitch(R[0].reshape(3, 3))[0] - _yaw_pitch(R_before[0])[0]) < 1e-9
    else:
        assert np.allclose(R[0].reshape(3, 3), R_before[0], atol=1e-12)
    # the injected gauge motion is undone up to the small changes
    assert np.abs(P - P_before).max() < 0.02
    # vector2double packs the same states again: rotations as (qx qy qz qw) of Rs
    for i in range(F):
        q = pose_rt[i, 3:]
        assert np.allclose(synthwin.quat_to_rot(np.array([q[3], q[0], q[1], q[2]])), R[i].reshape(3, 3), atol=1e-12)
    assert np.allclose(pose_rt[:, :3], P, atol=1e-15) and np.allclose(sb_rt[:, :3], V, atol=1e-15)
