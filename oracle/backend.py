"""ctypes wrapper of the C++ fp64 back-end oracle (oracle/backend/backend.cpp) — TEST INFRASTRUCTURE ONLY.

Consumes the interface structs of include/lfvio.h (declared once in lf-vio_b200/lfvio/backend.py).
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from lfvio import backend as B

_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_build", "liboracle_backend.so")
_lib = None


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(_PATH)
        dp = C.POINTER(C.c_double)
        L.oracle_default_opts.argtypes = [C.POINTER(B.SolverOpts)]
        L.oracle_default_opts.restype = None
        L.oracle_evaluate.argtypes = [C.POINTER(B.WindowState), C.POINTER(B.FactorList), C.POINTER(B.SolverOpts), C.c_int,
                                      dp, dp, dp, dp, dp, dp]
        L.oracle_imu_sqrt_info.argtypes = [C.POINTER(B.ImuFactor), dp]
        L.oracle_pose_plus.argtypes = [dp, dp, dp]
        L.oracle_pose_plus.restype = None
        L.oracle_optimize.argtypes = [C.POINTER(B.WindowState), C.POINTER(B.FactorList), C.POINTER(B.SolverOpts),
                                      C.POINTER(B.SolveSummary)]
        L.oracle_marginalize.argtypes = [C.POINTER(B.WindowState), C.POINTER(B.FactorList), C.POINTER(B.SolverOpts), C.c_int,
                                         C.c_int, C.POINTER(B.Prior)]
        L.oracle_time_backend.argtypes = [C.POINTER(B.WindowState), C.POINTER(B.FactorList), C.POINTER(B.SolverOpts), C.c_int,
                                          C.c_int, C.c_int, C.c_int]
        L.oracle_time_backend.restype = C.c_double
        _lib = L
    return _lib


def default_opts() -> B.SolverOpts:
    o = B.SolverOpts()
    lib().oracle_default_opts(C.byref(o))
    return o


def evaluate(w: B.WindowProblem, robust: bool = True) -> B.EvalResult:
    e = B.EvalResult(w)
    st, fl = w.state(), w.factors()
    rc = lib().oracle_evaluate(C.byref(st), C.byref(fl), C.byref(w.opts), int(robust), *e.args())
    assert rc == 0, rc
    return e.trim()


def optimize(w: B.WindowProblem) -> B.SolveSummary:
    s = B.SolveSummary()
    st, fl = w.state(), w.factors()
    rc = lib().oracle_optimize(C.byref(st), C.byref(fl), C.byref(w.opts), C.byref(s))
    assert rc == 0, rc
    return s


def marginalize(w: B.WindowProblem, flag: int, n_threads: int = 1) -> B.PriorData:
    out = B.PriorData()
    st, fl = w.state(), w.factors()
    rc = lib().oracle_marginalize(C.byref(st), C.byref(fl), C.byref(w.opts), flag, n_threads, C.byref(out.c))
    assert rc == 0, rc
    return out


def pose_plus(x, d):
    x = np.ascontiguousarray(x, np.float64)
    d = np.ascontiguousarray(d, np.float64)
    out = np.zeros(7)
    dp = C.POINTER(C.c_double)
    lib().oracle_pose_plus(x.ctypes.data_as(dp), d.ctypes.data_as(dp), out.ctypes.data_as(dp))
    return out


def imu_sqrt_info(f: B.ImuFactor) -> np.ndarray:
    out = np.zeros(225)
    rc = lib().oracle_imu_sqrt_info(C.byref(f), out.ctypes.data_as(C.POINTER(C.c_double)))
    assert rc == 0
    return out.reshape(15, 15)


def time_backend(w: B.WindowProblem, with_marginalize: bool, flag: int, marg_threads: int, reps: int) -> float:
    st, fl = w.state(), w.factors()
    return lib().oracle_time_backend(C.byref(st), C.byref(fl), C.byref(w.opts), int(with_marginalize), flag, marg_threads, reps)
