"""ctypes view of lfvio_preintegrate / lfvio_triangulate (include/lfvio.h): the two steps before optimization()."""
from __future__ import annotations

import ctypes as C

import numpy as np

from .backend import ImuFactor

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)


class ImuNoise(C.Structure):
    _fields_ = [("acc_n", C.c_double), ("gyr_n", C.c_double), ("acc_w", C.c_double), ("gyr_w", C.c_double)]


class ImuInterval(C.Structure):
    _fields_ = [("first", C.c_int), ("count", C.c_int), ("frame_i", C.c_int), ("propagate", C.c_int),
                ("acc_0", C.c_double * 3), ("gyr_0", C.c_double * 3), ("ba", C.c_double * 3), ("bg", C.c_double * 3),
                ("P", C.c_double * 3), ("V", C.c_double * 3), ("R", C.c_double * 9)]


def bind(lib):
    lib.lfvio_preintegrate.argtypes = [C.POINTER(ImuInterval), C.c_int, _dp, _dp, _dp, C.c_int, C.POINTER(ImuNoise), _dp,
                                       C.POINTER(ImuFactor), _dp, C.c_int]
    lib.lfvio_triangulate.argtypes = [_dp, _dp, C.c_int, _dp, _dp, _ip, _ip, _dp, C.c_int, C.c_double, _dp, C.c_int]


def _p(a, t=C.c_double):
    return a.ctypes.data_as(C.POINTER(t))


def make_interval(first, count, frame_i, acc0, gyr0, ba, bg, P=None, V=None, R=None) -> ImuInterval:
    iv = ImuInterval()
    iv.first, iv.count, iv.frame_i, iv.propagate = int(first), int(count), int(frame_i), int(P is not None)
    iv.acc_0[:] = list(map(float, acc0))
    iv.gyr_0[:] = list(map(float, gyr0))
    iv.ba[:] = list(map(float, ba))
    iv.bg[:] = list(map(float, bg))
    iv.P[:] = list(map(float, P if P is not None else np.zeros(3)))
    iv.V[:] = list(map(float, V if V is not None else np.zeros(3)))
    iv.R[:] = list(map(float, np.asarray(R if R is not None else np.eye(3)).ravel()))
    return iv


def preintegrate(lib, check, intervals, dt, acc, gyr, noise, g, device=0):
    """-> (list[ImuFactor], state (n x 15): P V R)"""
    n = len(intervals)
    arr = (ImuInterval * max(n, 1))(*intervals)
    dt = np.ascontiguousarray(dt, np.float64)
    acc = np.ascontiguousarray(acc, np.float64).reshape(-1, 3)
    gyr = np.ascontiguousarray(gyr, np.float64).reshape(-1, 3)
    nz = ImuNoise(*map(float, noise))
    gg = np.ascontiguousarray(g, np.float64)
    out = (ImuFactor * max(n, 1))()
    state = np.zeros((max(n, 1), 15))
    check(lib.lfvio_preintegrate(arr, n, _p(dt), _p(acc), _p(gyr), len(dt), C.byref(nz), _p(gg), out, _p(state), device),
          "lfvio_preintegrate")
    return [out[i] for i in range(n)], state[:n]


def triangulate(lib, check, Ps, Rs, tic, ric, start_frame, obs_first, obs_pts, depth, init_depth=5.0, device=0):
    Ps = np.ascontiguousarray(Ps, np.float64).reshape(-1, 3)
    Rs = np.ascontiguousarray(Rs, np.float64).reshape(-1, 9)
    tic = np.ascontiguousarray(tic, np.float64)
    ric = np.ascontiguousarray(ric, np.float64).reshape(9)
    sf = np.ascontiguousarray(start_frame, np.int32)
    of = np.ascontiguousarray(obs_first, np.int32)
    pts = np.ascontiguousarray(obs_pts, np.float64).reshape(-1, 3)
    d = np.array(depth, np.float64, copy=True)
    check(lib.lfvio_triangulate(_p(Ps), _p(Rs), len(Ps), _p(tic), _p(ric), _p(sf, C.c_int), _p(of, C.c_int), _p(pts), len(sf),
                                float(init_depth), _p(d), device), "lfvio_triangulate")
    return d
