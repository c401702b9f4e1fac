"""ctypes wrapper of oracle/_ref/libref_factors.so — TEST INFRASTRUCTURE ONLY.

That library is the REFERENCE's own factor / pre-integration / marginalisation / camera code, compiled unmodified from
/root/reference by `make -C oracle ref` (oracle/Makefile) against the stand-in Eigen/Ceres/ROS headers of
oracle/ref_stubs. It is what pins the self-written oracle (oracle/backend/backend.cpp, oracle/prestep.py,
oracle/camera.py) to the reference: tests/test_oracle_ref.py compares them, and tests/golden/make_ref_golden.py stores its
outputs as fixtures for boxes where neither /root/reference nor the prebuilt library exist.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from lfvio import backend as B

_DIR = os.path.dirname(os.path.abspath(__file__))
_PATH = os.path.join(_DIR, "_ref", "libref_factors.so")
_lib = None
_dp = C.POINTER(C.c_double)


def available(build: bool = True) -> bool:
    """True if the library exists (building it first when the reference tree is mounted)."""
    if build and os.path.isdir("/root/reference/vins_estimator/src/factor"):
        subprocess.run(["make", "-s", "-C", _DIR, "ref"], check=True, stdout=subprocess.DEVNULL)
    return os.path.exists(_PATH)


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(_PATH)
        L.ref_evaluate.argtypes = [C.POINTER(B.WindowState), C.POINTER(B.FactorList), C.POINTER(B.SolverOpts), C.c_int,
                                   _dp, _dp, _dp, _dp, _dp, _dp, _dp]
        L.ref_marginalize.argtypes = [C.POINTER(B.WindowState), C.POINTER(B.FactorList), C.POINTER(B.SolverOpts), C.c_int,
                                      C.POINTER(B.Prior), C.POINTER(C.c_int)]
        L.ref_pose_plus.argtypes = [_dp, _dp, _dp]
        L.ref_pose_plus.restype = None
        L.ref_pose_plus_jacobian.argtypes = [_dp, _dp]
        L.ref_pose_plus_jacobian.restype = None
        L.ref_lift_projective.argtypes = [_dp, _dp, _dp, C.c_int, C.c_int]
        L.ref_lift_projective.restype = None
        L.ref_space_to_plane.argtypes = [_dp, _dp, _dp, C.c_int]
        L.ref_space_to_plane.restype = None
        for f in (L.ref_R2ypr, L.ref_ypr2R, L.ref_g2R):
            f.argtypes = [_dp, _dp]
            f.restype = None
        _lib = L
    return _lib


def _p(a):
    return a.ctypes.data_as(_dp)


class RefEval(B.EvalResult):
    def __init__(self, w):
        super().__init__(w)
        self.jac_prior = np.zeros(max(self.n_prior, 1) ** 2)


def evaluate(w: B.WindowProblem, robust: bool = True) -> RefEval:
    e = RefEval(w)
    st, fl = w.state(), w.factors()
    rc = lib().ref_evaluate(C.byref(st), C.byref(fl), C.byref(w.opts), int(robust), _p(e.res_proj), _p(e.jac_proj), _p(e.res_imu),
                            _p(e.jac_imu), _p(e.res_prior), _p(e.jac_prior), C.byref(e.cost))
    assert rc == 0, rc
    e.trim()
    e.jac_prior = e.jac_prior[: e.n_prior ** 2].reshape(e.n_prior, e.n_prior)
    return e


def marginalize(w: B.WindowProblem, flag: int):
    """(prior, m): MarginalizationInfo after marginalize() + getParameterBlocks(); blocks in the reference's map order."""
    out = B.PriorData()
    m = C.c_int(0)
    st, fl = w.state(), w.factors()
    rc = lib().ref_marginalize(C.byref(st), C.byref(fl), C.byref(w.opts), flag, C.byref(out.c), C.byref(m))
    assert rc == 0, rc
    return out, m.value


def canonical_information(p: B.PriorData, n_frames: int):
    """(J0^T J0, J0^T r0, x0 by block id) with the columns reordered by ascending block id (the order the oracle and the
    product emit), so that priors whose blocks come in a different order can be compared."""
    nb = p.n_blocks
    ids = p.block_id[:nb]
    idx = p.block_idx[:nb]
    gsz = [1 if i == 2 * n_frames + 1 else (7 if (i == 2 * n_frames or i % 2 == 0) else 9) for i in ids]
    lsz = [6 if g == 7 else g for g in gsz]
    off = np.concatenate([[0], np.cumsum(gsz)])
    cols, x0 = [], {}
    for b in np.argsort(ids, kind="stable"):
        cols += list(range(idx[b], idx[b] + lsz[b]))
        x0[int(ids[b])] = p.x0[off[b]:off[b] + gsz[b]].copy()
    A, g = p.information()
    cols = np.array(cols, int)
    return A[np.ix_(cols, cols)], g[cols], x0


def pose_plus(x, d):
    x = np.ascontiguousarray(x, np.float64)
    d = np.ascontiguousarray(d, np.float64)
    out = np.zeros(7)
    lib().ref_pose_plus(_p(x), _p(d), _p(out))
    return out


def cam_vector(cam: dict) -> np.ndarray:
    return np.array(list(cam["poly"]) + list(cam["inv_poly"]) + [cam["C"], cam["D"], cam["E"], cam["cx"], cam["cy"]], np.float64)


def lift(cam: dict, uv, normalize: bool) -> np.ndarray:
    uv = np.ascontiguousarray(uv, np.float64).reshape(-1, 2)
    out = np.zeros((len(uv), 3))
    c = cam_vector(cam)
    lib().ref_lift_projective(_p(c), _p(uv), _p(out), len(uv), int(normalize))
    return out


def space_to_plane(cam: dict, P) -> np.ndarray:
    P = np.ascontiguousarray(P, np.float64).reshape(-1, 3)
    out = np.zeros((len(P), 2))
    c = cam_vector(cam)
    lib().ref_space_to_plane(_p(c), _p(P), _p(out), len(P))
    return out


def _mat3(fn, v, n_out):
    v = np.ascontiguousarray(v, np.float64)
    out = np.zeros(n_out)
    fn(_p(v), _p(out))
    return out


def R2ypr(R):
    return _mat3(lib().ref_R2ypr, np.asarray(R, np.float64).reshape(9), 3)


def ypr2R(ypr):
    return _mat3(lib().ref_ypr2R, ypr, 9).reshape(3, 3)


def g2R(g):
    return _mat3(lib().ref_g2R, g, 9).reshape(3, 3)


# ---------------------------------------------------------------------------------------------------------------------
# initial/initial_aligment.cpp and feature_manager.cpp of the reference (oracle/ref_shim/estimator_parts_shim.cpp)
# ---------------------------------------------------------------------------------------------------------------------
def visual_imu_alignment(Rs, Ts, imu, noise, tic, g_norm, mode: int = 2, bg=(0.0, 0.0, 0.0), ba=(0.0, 0.0, 0.0)):
    """imu[k] = (acc_0, gyr_0, dt[], acc[], gyr[]) of the interval ending at frame k+1 (violoop.World.imu[k+1] layout).
    mode 0 solveGyroscopeBias, 1 LinearAlignment, 2 VisualIMUAlignment. -> dict(ok, bg, g, x, pre)."""
    from lfvio import prestep as PS

    L = lib()
    n = len(Rs)
    ivs, dts, accs, gyrs, first = [], [], [], [], 0
    for (a0, g0, dt, acc, gyr) in imu:
        ivs.append(PS.make_interval(first, len(dt), 0, a0, g0, ba, bg))
        first += len(dt)
        dts.append(np.asarray(dt, float))
        accs.append(np.asarray(acc, float).reshape(-1, 3))
        gyrs.append(np.asarray(gyr, float).reshape(-1, 3))
    arr = (PS.ImuInterval * max(len(ivs), 1))(*ivs)
    dt, acc, gyr = np.concatenate(dts), np.concatenate(accs), np.concatenate(gyrs)
    R = np.ascontiguousarray(Rs, np.float64).reshape(n, 9)
    T = np.ascontiguousarray(Ts, np.float64).reshape(n, 3)
    nz = PS.ImuNoise(*map(float, noise))
    tic = np.ascontiguousarray(tic, np.float64)
    bgv, g, x = np.array(bg, np.float64), np.zeros(3), np.zeros(3 * n + 4)
    xs, ok = C.c_int(0), C.c_int(0)
    pre = (B.ImuFactor * max(n - 1, 1))()
    L.ref_visual_imu_alignment.argtypes = [C.c_int, _dp, _dp, C.POINTER(PS.ImuInterval), _dp, _dp, _dp, C.POINTER(PS.ImuNoise), _dp,
                                           C.c_double, C.c_int, _dp, _dp, _dp, C.POINTER(C.c_int), C.POINTER(C.c_int),
                                           C.POINTER(B.ImuFactor)]
    rc = L.ref_visual_imu_alignment(n, _p(R), _p(T), arr, _p(dt), _p(acc), _p(gyr), C.byref(nz), _p(tic), float(g_norm), mode,
                                    _p(bgv), _p(g), _p(x), C.byref(xs), C.byref(ok), pre)
    assert rc == 0
    return dict(ok=bool(ok.value), bg=bgv, g=g, x=x[: xs.value].copy(), pre=[pre[i] for i in range(n - 1)])


class RefFeatureManager:
    """The reference's FeatureManager (WINDOW_SIZE = 10), driven like the host mirror's test hooks."""

    def __init__(self, min_parallax: float):
        L = lib()
        L.ref_fm_create.restype = C.c_void_p
        L.ref_fm_create.argtypes = [C.c_double]
        L.ref_fm_destroy.argtypes = [C.c_void_p]
        L.ref_fm_add.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_double]
        L.ref_fm_corresponding.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int]
        L.ref_fm_slide.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p]
        L.ref_fm_snapshot.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        L.ref_fm_set_rs.argtypes = [C.c_void_p, C.c_void_p]
        L.ref_fm_triangulate.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.ref_fm_get_depth_vector.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
        L.ref_fm_set_depth.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
        L.ref_fm_last_track_num.argtypes = [C.c_void_p]
        self.L, self.h = L, L.ref_fm_create(float(min_parallax))

    def __del__(self):
        if getattr(self, "h", None):
            self.L.ref_fm_destroy(self.h)
            self.h = None

    def add(self, frame_count, ids, points8, td=0.0) -> bool:
        ids = np.ascontiguousarray(ids, np.int32)
        p = np.ascontiguousarray(points8, np.float64).reshape(len(ids), 8)
        return bool(self.L.ref_fm_add(self.h, frame_count, ids.ctypes.data, p.ctypes.data, len(ids), float(td)))

    def corresponding(self, l, r):
        out = np.zeros((4096, 6))
        n = self.L.ref_fm_corresponding(self.h, l, r, out.ctypes.data, 4096)
        return out[:n]

    def slide(self, op, arg=0, pose=None):
        p = None if pose is None else np.ascontiguousarray(pose, np.float64)
        self.L.ref_fm_slide(self.h, op, arg, None if p is None else p.ctypes.data)

    def snapshot(self, set_depth=None):
        cap = 8192
        meta, depth, cnt = np.zeros((cap, 4), np.int32), np.zeros(cap), C.c_int(0)
        sd = None if set_depth is None else np.ascontiguousarray(set_depth, np.float64)
        n = self.L.ref_fm_snapshot(self.h, meta.ctypes.data, depth.ctypes.data, cap, None if sd is None else sd.ctypes.data, C.byref(cnt))
        return meta[:n].copy(), depth[:n].copy(), cnt.value

    def triangulate(self, Ps, Rs, tic, ric):
        Rs = np.ascontiguousarray(Rs, np.float64).reshape(11, 9)
        Ps = np.ascontiguousarray(Ps, np.float64).reshape(11, 3)
        tic, ric = np.ascontiguousarray(tic, np.float64), np.ascontiguousarray(ric, np.float64).reshape(9)
        self.L.ref_fm_set_rs(self.h, Rs.ctypes.data)
        self.L.ref_fm_triangulate(self.h, Ps.ctypes.data, tic.ctypes.data, ric.ctypes.data)

    def depth_vector(self):
        out = np.zeros(8192)
        n = self.L.ref_fm_get_depth_vector(self.h, out.ctypes.data, 8192)
        return out[:n]

    def set_depth(self, x):
        x = np.ascontiguousarray(x, np.float64)
        self.L.ref_fm_set_depth(self.h, x.ctypes.data, len(x))


def find_fundamental(b1, b2, seed: int, frame: int):
    """myfindFundamentalMat (feature_tracker.cpp:266-327) with the injected random source -> (E 3x3, status)."""
    b1 = np.ascontiguousarray(b1, np.float64).reshape(-1, 3)
    b2 = np.ascontiguousarray(b2, np.float64).reshape(-1, 3)
    n = len(b1)
    status, E = np.zeros(n, np.uint8), np.zeros(9)
    L = lib()
    L.ref_find_fundamental.argtypes = [_dp, _dp, C.c_int, C.c_ulonglong, C.c_uint, C.c_void_p, _dp]
    assert L.ref_find_fundamental(_p(b1), _p(b2), n, seed, frame, status.ctypes.data, _p(E)) == 0
    return E.reshape(3, 3), status
