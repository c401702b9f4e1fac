"""ctypes view of the back-end half of include/lfvio.h (structs + lfvio_window_* prototypes).

The struct classes are also used by the oracle's ctypes wrapper (oracle/backend.py): the oracle consumes
the same interface structs as the product so that one problem description feeds both.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)
MAX_ITER_LOG = 64
MARGIN_OLD, MARGIN_SECOND_NEW = 0, 1


class ProjFactor(C.Structure):
    _fields_ = [("frame_i", C.c_int), ("frame_j", C.c_int), ("landmark", C.c_int), ("reserved", C.c_int),
                ("pts_i", C.c_double * 3), ("pts_j", C.c_double * 3), ("vel_i", C.c_double * 3), ("vel_j", C.c_double * 3),
                ("td_i", C.c_double), ("td_j", C.c_double), ("row_i", C.c_double), ("row_j", C.c_double)]


class ImuFactor(C.Structure):
    _fields_ = [("frame_i", C.c_int), ("reserved", C.c_int), ("sum_dt", C.c_double), ("delta_p", C.c_double * 3),
                ("delta_q", C.c_double * 4), ("delta_v", C.c_double * 3), ("linearized_ba", C.c_double * 3),
                ("linearized_bg", C.c_double * 3), ("jacobian", C.c_double * 225), ("covariance", C.c_double * 225)]


class Prior(C.Structure):
    _fields_ = [("n", C.c_int), ("n_blocks", C.c_int), ("block_id", _ip), ("block_idx", _ip), ("x0", _dp), ("J0", _dp),
                ("r0", _dp), ("cap_n", C.c_int), ("cap_blocks", C.c_int)]


class WindowState(C.Structure):
    _fields_ = [("n_frames", C.c_int), ("n_landmarks", C.c_int), ("para_Pose", _dp), ("para_SpeedBias", _dp),
                ("para_Ex_Pose", _dp), ("para_Feature", _dp), ("para_Td", _dp)]


class FactorList(C.Structure):
    _fields_ = [("n_proj", C.c_int), ("proj", C.POINTER(ProjFactor)), ("n_imu", C.c_int), ("imu", C.POINTER(ImuFactor)),
                ("prior", C.POINTER(Prior))]


class SolverOpts(C.Structure):
    _fields_ = [("max_num_iterations", C.c_int), ("fixed_iterations", C.c_int), ("estimate_td", C.c_int),
                ("estimate_extrinsic", C.c_int), ("row", C.c_int), ("reserved", C.c_int), ("focal_length", C.c_double),
                ("tr", C.c_double), ("g", C.c_double * 3), ("cauchy_a", C.c_double),
                ("initial_trust_region_radius", C.c_double), ("max_trust_region_radius", C.c_double),
                ("min_trust_region_radius", C.c_double), ("min_relative_decrease", C.c_double),
                ("function_tolerance", C.c_double), ("gradient_tolerance", C.c_double), ("parameter_tolerance", C.c_double)]


class SolveSummary(C.Structure):
    _fields_ = [("iterations", C.c_int), ("num_successful_steps", C.c_int), ("termination", C.c_int), ("reserved", C.c_int),
                ("initial_cost", C.c_double), ("final_cost", C.c_double), ("cost", C.c_double * MAX_ITER_LOG),
                ("radius", C.c_double * MAX_ITER_LOG), ("accepted", C.c_int * MAX_ITER_LOG)]


def _p(a: np.ndarray, t=C.c_double):
    return a.ctypes.data_as(C.POINTER(t))


class PriorData:
    """numpy-owned storage behind a lfvio_prior."""

    def __init__(self, cap_n: int = 512, cap_blocks: int = 64):
        self.block_id = np.zeros(cap_blocks, np.int32)
        self.block_idx = np.zeros(cap_blocks, np.int32)
        self.x0 = np.zeros(cap_blocks * 9, np.float64)
        self.J0 = np.zeros(cap_n * cap_n, np.float64)
        self.r0 = np.zeros(cap_n, np.float64)
        self.c = Prior(0, 0, _p(self.block_id, C.c_int), _p(self.block_idx, C.c_int), _p(self.x0), _p(self.J0), _p(self.r0),
                       cap_n, cap_blocks)

    @property
    def n(self):
        return self.c.n

    @property
    def n_blocks(self):
        return self.c.n_blocks

    def jacobian(self) -> np.ndarray:
        return self.J0[: self.n * self.n].reshape(self.n, self.n)

    def residual(self) -> np.ndarray:
        return self.r0[: self.n]

    def information(self):
        """(J0^T J0, J0^T r0): the invariants of the prior (marginalization_factor.cpp:295-296)."""
        J, r = self.jacobian(), self.residual()
        return J.T @ J, J.T @ r


class WindowProblem:
    """One Estimator::optimization() problem in the flat layout of include/lfvio.h (numpy-owned)."""

    def __init__(self, n_frames: int, pose, speed_bias, ex_pose, feature, td, proj: list, imu: list, prior: PriorData | None,
                 opts: SolverOpts):
        self.F = n_frames
        self.pose = np.ascontiguousarray(pose, np.float64).reshape(n_frames, 7).copy()
        self.sb = np.ascontiguousarray(speed_bias, np.float64).reshape(n_frames, 9).copy()
        self.ex = np.ascontiguousarray(ex_pose, np.float64).reshape(7).copy()
        self.feat = np.ascontiguousarray(feature, np.float64).reshape(-1).copy()
        self.td = np.array([td], np.float64)
        self.L = self.feat.shape[0]
        self.proj = (ProjFactor * max(len(proj), 1))(*proj)
        self.n_proj = len(proj)
        self.imu = (ImuFactor * max(len(imu), 1))(*imu)
        self.n_imu = len(imu)
        self.prior = prior
        self.opts = opts
        self._feat_buf = self.feat if self.L > 0 else np.zeros(1)

    def state(self) -> WindowState:
        return WindowState(self.F, self.L, _p(self.pose), _p(self.sb), _p(self.ex), _p(self._feat_buf), _p(self.td))

    def factors(self) -> FactorList:
        pr = C.pointer(self.prior.c) if (self.prior is not None and self.prior.n > 0) else None
        return FactorList(self.n_proj, self.proj, self.n_imu, self.imu, pr)

    def clone(self) -> "WindowProblem":
        w = WindowProblem.__new__(WindowProblem)
        w.__dict__.update(self.__dict__)
        w.pose, w.sb, w.ex, w.feat, w.td = self.pose.copy(), self.sb.copy(), self.ex.copy(), self.feat.copy(), self.td.copy()
        w._feat_buf = w.feat if w.L > 0 else np.zeros(1)
        return w

    def state_vector(self) -> np.ndarray:
        return np.concatenate([self.pose.ravel(), self.sb.ravel(), self.ex, self.feat, self.td])


def bind(lib):
    """Declare the lfvio_window_* prototypes on a loaded liblfvio.so."""
    lib.lfvio_solver_default_opts.argtypes = [C.POINTER(SolverOpts)]
    lib.lfvio_solver_default_opts.restype = None
    lib.lfvio_window_create.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_void_p)]
    lib.lfvio_window_destroy.argtypes = [C.c_void_p]
    lib.lfvio_window_destroy.restype = None
    lib.lfvio_window_optimize.argtypes = [C.c_void_p, C.POINTER(WindowState), C.POINTER(FactorList), C.POINTER(SolverOpts),
                                          C.POINTER(SolveSummary)]
    lib.lfvio_window_marginalize.argtypes = [C.c_void_p, C.POINTER(WindowState), C.POINTER(FactorList), C.POINTER(SolverOpts),
                                             C.c_int, C.POINTER(Prior)]
    lib.lfvio_window_optimize_full.argtypes = [C.c_void_p, C.POINTER(WindowState), C.POINTER(FactorList), C.POINTER(SolverOpts),
                                               C.c_int, C.POINTER(SolveSummary), C.POINTER(Prior)]
    lib.lfvio_window_evaluate.argtypes = [C.c_void_p, C.POINTER(WindowState), C.POINTER(FactorList), C.POINTER(SolverOpts),
                                          C.c_int, _dp, _dp, _dp, _dp, _dp, _dp]
    lib.lfvio_window_stream.argtypes = [C.c_void_p]
    lib.lfvio_window_stream.restype = C.c_void_p
    lib.lfvio_window_upload.argtypes = [C.c_void_p, C.POINTER(WindowState), C.POINTER(FactorList), C.POINTER(SolverOpts)]
    lib.lfvio_window_optimize_dev.argtypes = [C.c_void_p, C.c_int, C.c_int]
    lib.lfvio_window_download.argtypes = [C.c_void_p, C.POINTER(WindowState), C.POINTER(SolveSummary)]
    # FeatureManager on the device
    lib.lfvio_fm_create.argtypes = [C.c_int, C.c_int, C.c_int, C.c_double, C.c_double, C.c_int, C.POINTER(C.c_void_p)]
    lib.lfvio_fm_destroy.argtypes = [C.c_void_p]
    lib.lfvio_fm_destroy.restype = None
    lib.lfvio_fm_clear.argtypes = [C.c_void_p]
    lib.lfvio_fm_add_frame.argtypes = [C.c_void_p, C.c_int, _ip, _dp, C.c_int, C.c_double, _ip, _ip]
    lib.lfvio_fm_remove_back.argtypes = [C.c_void_p]
    lib.lfvio_fm_remove_back_shift_depth.argtypes = [C.c_void_p, _dp, _dp, _dp, _dp]
    lib.lfvio_fm_remove_front.argtypes = [C.c_void_p, C.c_int]
    lib.lfvio_fm_remove_failures.argtypes = [C.c_void_p]
    lib.lfvio_fm_depth.argtypes = [C.c_void_p, C.c_int, _dp, C.c_int, _ip]
    lib.lfvio_fm_triangulate.argtypes = [C.c_void_p, _dp, _dp, _dp, _dp]
    lib.lfvio_fm_build_factors.argtypes = [C.c_void_p, _ip, _ip]
    lib.lfvio_fm_download_factors.argtypes = [C.c_void_p, C.POINTER(ProjFactor), C.c_int, _ip, _dp, C.c_int]
    lib.lfvio_fm_factors_dev.argtypes = [C.c_void_p, C.POINTER(C.POINTER(ProjFactor)), C.POINTER(_ip), C.POINTER(_dp), C.POINTER(C.c_void_p)]
    lib.lfvio_fm_snapshot.argtypes = [C.c_void_p, _ip, _dp, C.c_int]
    lib.lfvio_window_optimize_full_fm.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(WindowState), C.POINTER(FactorList),
                                                  C.POINTER(SolverOpts), C.c_int, C.POINTER(SolveSummary), C.POINTER(Prior)]


def default_opts(lib) -> SolverOpts:
    o = SolverOpts()
    lib.lfvio_solver_default_opts(C.byref(o))
    return o


class EvalResult:
    def __init__(self, w: WindowProblem):
        n = w.prior.n if (w.prior is not None) else 0
        self.res_proj = np.zeros(2 * max(w.n_proj, 1))
        self.jac_proj = np.zeros(40 * max(w.n_proj, 1))
        self.res_imu = np.zeros(15 * max(w.n_imu, 1))
        self.jac_imu = np.zeros(450 * max(w.n_imu, 1))
        self.res_prior = np.zeros(max(n, 1))
        self.cost = C.c_double(0)
        self.n_proj, self.n_imu, self.n_prior = w.n_proj, w.n_imu, n

    def args(self):
        return (_p(self.res_proj), _p(self.jac_proj), _p(self.res_imu), _p(self.jac_imu), _p(self.res_prior), C.byref(self.cost))

    def trim(self):
        self.res_proj = self.res_proj[: 2 * self.n_proj]
        self.jac_proj = self.jac_proj[: 40 * self.n_proj].reshape(-1, 2, 20)
        self.res_imu = self.res_imu[: 15 * self.n_imu]
        self.jac_imu = self.jac_imu[: 450 * self.n_imu].reshape(-1, 15, 30)
        self.res_prior = self.res_prior[: self.n_prior]
        return self


class FeatureStore:
    """lfvio_fm handle: the reference's FeatureManager as a structure of arrays on the device."""

    def __init__(self, lib, check, window_size: int = 10, capacity: int = 2048, max_proj: int = 16384, min_parallax: float = 10.0 / 160.0,
                 init_depth: float = 5.0, device: int = 0):
        self._lib, self._check = lib, check
        self.W, self.capacity, self.max_proj = window_size, capacity, max_proj
        self._h = C.c_void_p()
        check(lib.lfvio_fm_create(window_size, capacity, max_proj, min_parallax, init_depth, device, C.byref(self._h)), "lfvio_fm_create")

    def close(self):
        if self._h:
            self._lib.lfvio_fm_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def clear(self):
        self._check(self._lib.lfvio_fm_clear(self._h), "lfvio_fm_clear")

    def add_frame(self, frame_count: int, ids, feat8, td: float = 0.0):
        """addFeatureCheckParallax: ids ascending, feat8 n x [x y z u v vx vy vz]; returns (is_keyframe, last_track_num)."""
        ids = np.ascontiguousarray(ids, np.int32)
        f8 = np.ascontiguousarray(feat8, np.float64).reshape(-1, 8)
        key, trk = C.c_int(0), C.c_int(0)
        self._check(self._lib.lfvio_fm_add_frame(self._h, frame_count, _p(ids, C.c_int), _p(f8), len(ids), td, C.byref(key), C.byref(trk)),
                    "lfvio_fm_add_frame")
        return bool(key.value), trk.value

    def remove_back(self):
        self._check(self._lib.lfvio_fm_remove_back(self._h), "lfvio_fm_remove_back")

    def remove_back_shift_depth(self, marg_R, marg_P, new_R, new_P):
        a = [np.ascontiguousarray(x, np.float64) for x in (marg_R, marg_P, new_R, new_P)]
        self._check(self._lib.lfvio_fm_remove_back_shift_depth(self._h, *[_p(x) for x in a]), "lfvio_fm_remove_back_shift_depth")

    def remove_front(self, frame_count: int):
        self._check(self._lib.lfvio_fm_remove_front(self._h, frame_count), "lfvio_fm_remove_front")

    def remove_failures(self):
        self._check(self._lib.lfvio_fm_remove_failures(self._h), "lfvio_fm_remove_failures")

    def set_depth(self, x, clear: bool = False):
        x = np.ascontiguousarray(x, np.float64)
        n = C.c_int(0)
        self._check(self._lib.lfvio_fm_depth(self._h, 1 if clear else 0, _p(x), len(x), C.byref(n)), "lfvio_fm_depth")
        return n.value

    def depth_vector(self) -> np.ndarray:
        x, n = np.zeros(self.capacity), C.c_int(0)
        self._check(self._lib.lfvio_fm_depth(self._h, 2, _p(x), self.capacity, C.byref(n)), "lfvio_fm_depth")
        return x[: n.value].copy()

    def triangulate(self, Ps, Rs, tic, ric):
        a = [np.ascontiguousarray(x, np.float64) for x in (Ps, Rs, tic, ric)]
        assert a[0].size == 3 * (self.W + 1) and a[1].size == 9 * (self.W + 1)
        self._check(self._lib.lfvio_fm_triangulate(self._h, *[_p(x) for x in a]), "lfvio_fm_triangulate")

    def build_factors(self):
        """Returns (proj records, lm_first, para_Feature) of the list built on the device (copied out for inspection)."""
        R, L = C.c_int(0), C.c_int(0)
        self._check(self._lib.lfvio_fm_build_factors(self._h, C.byref(R), C.byref(L)), "lfvio_fm_build_factors")
        proj = (ProjFactor * max(R.value, 1))()
        lm = np.zeros(L.value + 1, np.int32)
        feat = np.zeros(max(L.value, 1))
        self._check(self._lib.lfvio_fm_download_factors(self._h, proj, max(R.value, 1), _p(lm, C.c_int), _p(feat), L.value),
                    "lfvio_fm_download_factors")
        return list(proj)[: R.value], lm, feat[: L.value]

    def snapshot(self):
        """[(feature_id, start_frame, n_obs, estimated_depth)] in list order."""
        meta, depth = np.zeros((self.capacity, 3), np.int32), np.zeros(self.capacity)
        n = self._lib.lfvio_fm_snapshot(self._h, _p(meta, C.c_int), _p(depth), self.capacity)
        if n < 0:
            self._check(n, "lfvio_fm_snapshot")
        return [(int(meta[i, 0]), int(meta[i, 1]), int(meta[i, 2]), float(depth[i])) for i in range(n)]


class Window:
    """lfvio_window handle: Estimator::optimization() on device."""

    def __init__(self, lib, check, window_size: int = 10, max_landmarks: int = 1000, max_proj: int = 16384, device: int = 0):
        self._lib, self._check = lib, check
        self._h = C.c_void_p()
        check(lib.lfvio_window_create(window_size, max_landmarks, max_proj, device, C.byref(self._h)), "lfvio_window_create")

    def close(self):
        if self._h:
            self._lib.lfvio_window_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def optimize(self, w: WindowProblem) -> SolveSummary:
        s = SolveSummary()
        st, fl = w.state(), w.factors()
        self._check(self._lib.lfvio_window_optimize(self._h, C.byref(st), C.byref(fl), C.byref(w.opts), C.byref(s)),
                    "lfvio_window_optimize")
        return s

    def optimize_full(self, w: WindowProblem, flag: int = MARGIN_OLD, out: PriorData | None = None):
        """solve + gauge re-anchoring + next prior in one call; returns (summary, prior)."""
        out = out or PriorData()
        s = SolveSummary()
        st, fl = w.state(), w.factors()
        self._check(self._lib.lfvio_window_optimize_full(self._h, C.byref(st), C.byref(fl), C.byref(w.opts), flag, C.byref(s),
                                                         C.byref(out.c)), "lfvio_window_optimize_full")
        return s, out

    def prepared_optimize_full(self, w: WindowProblem, flag: int = MARGIN_OLD, out: PriorData | None = None):
        """The same C-ABI call with its argument structs built once: returns call() -> status for loops that make the call many
        times on the same numpy-owned buffers (the structs only hold pointers to them), e.g. bench.py's estimator node."""
        out = out or PriorData()
        keep = (w.state(), w.factors(), SolveSummary(), out, w)
        fn, h, check = self._lib.lfvio_window_optimize_full, self._h, self._check
        a = (C.byref(keep[0]), C.byref(keep[1]), C.byref(w.opts), flag, C.byref(keep[2]), C.byref(out.c))

        def call():
            check(fn(h, *a), "lfvio_window_optimize_full")
            return keep[2]

        call.keep = keep
        return call

    def optimize_full_fm(self, fm: "FeatureStore", w: WindowProblem, flag: int = MARGIN_OLD, out: PriorData | None = None, max_landmarks: int = 4096):
        """optimize_full with the projection factors and para_Feature taken from the device FeatureStore (w.proj / w.feat are
        ignored on input; w.feat returns the optimised inverse depths); returns (summary, prior)."""
        out = out or PriorData()
        s = SolveSummary()
        feat = np.zeros(max_landmarks)
        st = WindowState(w.F, 0, _p(w.pose), _p(w.sb), _p(w.ex), _p(feat), _p(w.td))
        pr = C.pointer(w.prior.c) if (w.prior is not None and w.prior.n > 0) else None
        fl = FactorList(0, None, w.n_imu, w.imu, pr)
        self._check(self._lib.lfvio_window_optimize_full_fm(self._h, fm._h, C.byref(st), C.byref(fl), C.byref(w.opts), flag, C.byref(s),
                                                            C.byref(out.c)), "lfvio_window_optimize_full_fm")
        w.feat = feat[: st.n_landmarks].copy()
        w.L = st.n_landmarks
        w._feat_buf = w.feat if w.L > 0 else np.zeros(1)
        return s, out

    def marginalize(self, w: WindowProblem, flag: int, out: PriorData | None = None) -> PriorData:
        out = out or PriorData()
        st, fl = w.state(), w.factors()
        self._check(self._lib.lfvio_window_marginalize(self._h, C.byref(st), C.byref(fl), C.byref(w.opts), flag, C.byref(out.c)),
                    "lfvio_window_marginalize")
        return out

    def evaluate(self, w: WindowProblem, robust: bool = True) -> EvalResult:
        e = EvalResult(w)
        st, fl = w.state(), w.factors()
        self._check(self._lib.lfvio_window_evaluate(self._h, C.byref(st), C.byref(fl), C.byref(w.opts), int(robust), *e.args()),
                    "lfvio_window_evaluate")
        return e.trim()

    def upload(self, w: WindowProblem):
        st, fl = w.state(), w.factors()
        self._check(self._lib.lfvio_window_upload(self._h, C.byref(st), C.byref(fl), C.byref(w.opts)), "lfvio_window_upload")

    def optimize_dev(self, with_marginalize: bool = True, flag: int = MARGIN_OLD):
        self._check(self._lib.lfvio_window_optimize_dev(self._h, int(with_marginalize), flag), "lfvio_window_optimize_dev")

    def download(self, w: WindowProblem) -> SolveSummary:
        s = SolveSummary()
        st = w.state()
        self._check(self._lib.lfvio_window_download(self._h, C.byref(st), C.byref(s)), "lfvio_window_download")
        return s

    @property
    def stream(self) -> int:
        return self._lib.lfvio_window_stream(self._h) or 0
