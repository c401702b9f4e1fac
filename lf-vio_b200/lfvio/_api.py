"""ctypes view of the lfvio C-ABI (include/lfvio.h) — test/bench harness only.

The product is liblfvio.so (hand-written sm_100a kernels behind a C-ABI) plus the C++ host mirrors
of the reference classes in lf-vio_b200/host/.  This module only loads the library and declares
the prototypes so pytest and bench.py can call exactly what a reference-side binding would call.
It contains no compute and no fallback: if the library is missing, import fails loudly.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(os.path.dirname(_HERE), "liblfvio.so")

if not os.path.exists(LIB_PATH):
    raise ImportError(
        f"{LIB_PATH} not built — run `make -C lf-vio_b200` (or __graft_entry__.build()). "
        "There is no CPU fallback."
    )

lib = C.CDLL(LIB_PATH)


class LfvioError(RuntimeError):
    def __init__(self, status: int, where: str):
        msg = lib.lfvio_last_error().decode("utf-8", "replace")
        super().__init__(f"{where} failed with status {status}: {msg}")
        self.status = status


def check(status: int, where: str) -> None:
    if status != 0:
        raise LfvioError(status, where)


class Ocam(C.Structure):
    """lfvio_ocam (include/lfvio.h) == camodocal::OCAMCamera::Parameters."""

    _fields_ = [
        ("poly", C.c_double * 5),
        ("inv_poly", C.c_double * 20),
        ("C", C.c_double),
        ("D", C.c_double),
        ("E", C.c_double),
        ("cx", C.c_double),
        ("cy", C.c_double),
        ("image_width", C.c_int),
        ("image_height", C.c_int),
    ]

    @classmethod
    def from_dict(cls, d: dict) -> "Ocam":
        o = cls()
        for i, v in enumerate(d["poly"]):
            o.poly[i] = v
        for i, v in enumerate(d["inv_poly"]):
            o.inv_poly[i] = v
        o.C, o.D, o.E = d["C"], d["D"], d["E"]
        o.cx, o.cy = d["cx"], d["cy"]
        o.image_width, o.image_height = d["image_width"], d["image_height"]
        return o


lib.lfvio_last_error.restype = C.c_char_p
lib.lfvio_version.restype = C.c_char_p
lib.lfvio_device_count.restype = C.c_int
lib.lfvio_launch_count.restype = C.c_uint64

_dp = C.POINTER(C.c_double)
lib.lfvio_lift_projective.argtypes = [C.POINTER(Ocam), _dp, _dp, C.c_int, C.c_int, C.c_int]
lib.lfvio_lift_projective.restype = C.c_int
lib.lfvio_lift_projective_dev.argtypes = [C.POINTER(Ocam), C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p]
lib.lfvio_lift_projective_dev.restype = C.c_int


def _as(arr: np.ndarray, ctype):
    return arr.ctypes.data_as(C.POINTER(ctype))


def lift_projective(cam: Ocam, uv: np.ndarray, normalize: bool = False, device: int = 0) -> np.ndarray:
    """Batched OCAMCamera::liftProjective through the C-ABI (host buffers)."""
    uv = np.ascontiguousarray(uv, dtype=np.float64).reshape(-1, 2)
    out = np.empty((uv.shape[0], 3), dtype=np.float64)
    check(
        lib.lfvio_lift_projective(C.byref(cam), _as(uv, C.c_double), _as(out, C.c_double), uv.shape[0],
                                  int(normalize), device),
        "lfvio_lift_projective",
    )
    return out


from .config import PAL_UP, scaled_camera, tracker_defaults  # noqa: E402,F401

# ------------------------------------------------------------------------------------------------
# stage-level front-end operators
# ------------------------------------------------------------------------------------------------
_u8p = C.POINTER(C.c_uint8)
_fp = C.POINTER(C.c_float)
_ip = C.POINTER(C.c_int)

lib.lfvio_clahe.argtypes = [_u8p, C.c_int, C.c_int, C.c_int, C.c_double, _u8p, C.c_int, C.c_int]
lib.lfvio_pyr_down.argtypes = [_u8p, C.c_int, C.c_int, C.c_int, _u8p, C.c_int, C.c_int]
lib.lfvio_calc_optical_flow_pyr_lk.argtypes = [_u8p, _u8p, C.c_int, C.c_int, C.c_int, _fp, C.c_int, _fp, _u8p, C.c_int, C.c_int]
lib.lfvio_corner_min_eigen_val.argtypes = [_u8p, C.c_int, C.c_int, C.c_int, _fp, C.c_int]
lib.lfvio_good_features_to_track.argtypes = [_u8p, C.c_int, C.c_int, C.c_int, _u8p, C.c_int, C.c_int, C.c_double,
                                             C.c_double, _fp, _ip, C.c_int]


def _img(a: np.ndarray) -> np.ndarray:
    a = np.ascontiguousarray(a, dtype=np.uint8)
    assert a.ndim == 2
    return a


def clahe(img: np.ndarray, clip_limit: float = 3.0, device: int = 0) -> np.ndarray:
    img = _img(img)
    out = np.empty_like(img)
    check(lib.lfvio_clahe(_as(img, C.c_uint8), img.shape[0], img.shape[1], img.strides[0], clip_limit,
                          _as(out, C.c_uint8), out.strides[0], device), "lfvio_clahe")
    return out


def pyr_down(img: np.ndarray, device: int = 0) -> np.ndarray:
    img = _img(img)
    out = np.empty(((img.shape[0] + 1) // 2, (img.shape[1] + 1) // 2), np.uint8)
    check(lib.lfvio_pyr_down(_as(img, C.c_uint8), img.shape[0], img.shape[1], img.strides[0], _as(out, C.c_uint8),
                             out.strides[0], device), "lfvio_pyr_down")
    return out


def calc_optical_flow_pyr_lk(prev: np.ndarray, nxt: np.ndarray, pts: np.ndarray, max_level: int = 3, device: int = 0):
    prev, nxt = _img(prev), _img(nxt)
    assert prev.shape == nxt.shape
    pts = np.ascontiguousarray(pts, np.float32).reshape(-1, 2)
    n = pts.shape[0]
    out = np.zeros((n, 2), np.float32)
    st = np.zeros(n, np.uint8)
    check(lib.lfvio_calc_optical_flow_pyr_lk(_as(prev, C.c_uint8), _as(nxt, C.c_uint8), prev.shape[0], prev.shape[1],
                                             prev.strides[0], _as(pts, C.c_float), n, _as(out, C.c_float),
                                             _as(st, C.c_uint8), max_level, device), "lfvio_calc_optical_flow_pyr_lk")
    return out, st


def corner_min_eigen_val(img: np.ndarray, device: int = 0) -> np.ndarray:
    img = _img(img)
    out = np.empty(img.shape, np.float32)
    check(lib.lfvio_corner_min_eigen_val(_as(img, C.c_uint8), img.shape[0], img.shape[1], img.strides[0],
                                         _as(out, C.c_float), device), "lfvio_corner_min_eigen_val")
    return out


def good_features_to_track(img: np.ndarray, max_corners: int, quality: float, min_distance: float, mask=None,
                           device: int = 0) -> np.ndarray:
    img = _img(img)
    out = np.zeros((max_corners, 2), np.float32)
    n = C.c_int(0)
    if mask is not None:
        mask = _img(mask)
        mp, ms = _as(mask, C.c_uint8), mask.strides[0]
    else:
        mp, ms = None, 0
    check(lib.lfvio_good_features_to_track(_as(img, C.c_uint8), img.shape[0], img.shape[1], img.strides[0], mp, ms,
                                           max_corners, quality, min_distance, _as(out, C.c_float), C.byref(n), device),
          "lfvio_good_features_to_track")
    return out[: n.value].copy()


lib.lfvio_reject_with_f.argtypes = [C.POINTER(Ocam), _fp, _fp, C.c_int, C.c_uint64, C.c_uint, _u8p, C.POINTER(C.c_double),
                                    _fp, _ip]


def reject_with_f(cam: dict, cur_pts: np.ndarray, forw_pts: np.ndarray, seed: int, frame: int):
    """FeatureTracker::rejectWithF (feature_tracker.cpp:329-386).  Returns (status u8[n], E 3x3, scores f32[100], best)."""
    cur = np.ascontiguousarray(cur_pts, np.float32).reshape(-1, 2)
    forw = np.ascontiguousarray(forw_pts, np.float32).reshape(-1, 2)
    n = len(cur)
    status = np.zeros(n, np.uint8)
    E = np.zeros(9, np.float64)
    scores = np.zeros(100, np.float32)
    best = np.zeros(1, np.int32)
    oc = Ocam.from_dict(cam)
    check(lib.lfvio_reject_with_f(C.byref(oc), _as(cur, C.c_float), _as(forw, C.c_float), n, seed, frame, _as(status, C.c_uint8),
                                  E.ctypes.data_as(C.POINTER(C.c_double)), _as(scores, C.c_float), _as(best, C.c_int)),
          "lfvio_reject_with_f")
    return status, E.reshape(3, 3), scores, int(best[0])


# ------------------------------------------------------------------------------------------------
# FeatureTracker handle
# ------------------------------------------------------------------------------------------------
class TrackerCfg(C.Structure):
    _fields_ = [
        ("rows", C.c_int), ("cols", C.c_int), ("max_cnt", C.c_int), ("min_dist", C.c_int), ("equalize", C.c_int),
        ("center_x", C.c_double), ("center_y", C.c_double), ("max_r", C.c_int), ("min_r", C.c_int),
        ("capacity", C.c_int), ("ransac", C.c_int), ("cam", Ocam), ("ransac_seed", C.c_uint64),
    ]


class Tracks(C.Structure):
    _fields_ = [
        ("capacity", C.c_int), ("n", C.c_int), ("cur_pts", _fp), ("ids", _ip), ("track_cnt", _ip),
        ("cur_un_pts_3d", _fp), ("pts_velocity_3d", _fp),
    ]


lib.lfvio_tracker_create.argtypes = [C.POINTER(TrackerCfg), C.c_int, C.POINTER(C.c_void_p)]
lib.lfvio_tracker_destroy.argtypes = [C.c_void_p]
lib.lfvio_tracker_destroy.restype = None
lib.lfvio_tracker_read_image.argtypes = [C.c_void_p, _u8p, C.c_int, C.c_double, C.c_int]
lib.lfvio_tracker_read_image_dev.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_double, C.c_int]
lib.lfvio_tracker_sync.argtypes = [C.c_void_p]
lib.lfvio_tracker_count.argtypes = [C.c_void_p]
lib.lfvio_tracker_get.argtypes = [C.c_void_p, C.POINTER(Tracks)]
lib.lfvio_tracker_update_ids.argtypes = [C.c_void_p, _ip]
lib.lfvio_tracker_get_mask.argtypes = [C.c_void_p, _u8p, C.c_int]
lib.lfvio_tracker_get_image.argtypes = [C.c_void_p, C.c_int, _u8p, C.c_int]
lib.lfvio_tracker_set_points.argtypes = [C.c_void_p, _fp, _ip, _ip, C.c_int]
lib.lfvio_tracker_stream.argtypes = [C.c_void_p]
lib.lfvio_tracker_stream.restype = C.c_void_p


class FeatureTracker:
    """Python view of the C-ABI tracker handle; mirrors class FeatureTracker of the reference
    (feature_tracker/src/feature_tracker.h:34-80): readImage / updateID and the public members."""

    def __init__(self, cam: dict, cfg: dict, device: int = 0, capacity: int = 0):
        c = TrackerCfg()
        c.rows, c.cols = cam["image_height"], cam["image_width"]
        c.max_cnt, c.min_dist, c.equalize = cfg["max_cnt"], cfg["min_dist"], cfg.get("equalize", 1)
        c.center_x, c.center_y = cfg["center_x"], cfg["center_y"]
        c.max_r, c.min_r = cfg["max_r"], cfg["min_r"]
        c.capacity, c.ransac = capacity, int(bool(cfg.get("ransac", 0)))
        c.ransac_seed = int(cfg.get("ransac_seed", 0))
        c.cam = Ocam.from_dict(cam)
        self.rows, self.cols = c.rows, c.cols
        self._cap = max(capacity, c.max_cnt)
        self._h = C.c_void_p()
        check(lib.lfvio_tracker_create(C.byref(c), device, C.byref(self._h)), "lfvio_tracker_create")
        self.n_id = 0
        self.cur_pts = np.zeros((0, 2), np.float32)
        self.ids = np.zeros(0, np.int32)
        self.track_cnt = np.zeros(0, np.int32)
        self.cur_un_pts_3d = np.zeros((0, 3), np.float32)
        self.pts_velocity_3d = np.zeros((0, 3), np.float32)

    def close(self):
        if self._h:
            lib.lfvio_tracker_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def handle(self):
        return self._h

    def _refresh(self):
        cap = self._cap
        pts = np.zeros((cap, 2), np.float32)
        ids = np.zeros(cap, np.int32)
        cnt = np.zeros(cap, np.int32)
        b = np.zeros((cap, 3), np.float32)
        v = np.zeros((cap, 3), np.float32)
        t = Tracks(cap, 0, _as(pts, C.c_float), _as(ids, C.c_int), _as(cnt, C.c_int), _as(b, C.c_float), _as(v, C.c_float))
        check(lib.lfvio_tracker_get(self._h, C.byref(t)), "lfvio_tracker_get")
        n = t.n
        self.cur_pts, self.ids, self.track_cnt = pts[:n], ids[:n], cnt[:n]
        self.cur_un_pts_3d, self.pts_velocity_3d = b[:n], v[:n]

    def read_image(self, img: np.ndarray, cur_time: float, pub_this_frame: bool = True):
        img = _img(img)
        assert img.shape == (self.rows, self.cols)
        check(lib.lfvio_tracker_read_image(self._h, _as(img, C.c_uint8), img.strides[0], cur_time, int(pub_this_frame)),
              "lfvio_tracker_read_image")
        self._refresh()

    def read_image_dev(self, dev_ptr: int, stride: int, cur_time: float, pub_this_frame: bool = True):
        check(lib.lfvio_tracker_read_image_dev(self._h, C.c_void_p(dev_ptr), stride, cur_time, int(pub_this_frame)),
              "lfvio_tracker_read_image_dev")

    def sync(self):
        check(lib.lfvio_tracker_sync(self._h), "lfvio_tracker_sync")
        self._refresh()

    def update_ids(self):
        n = C.c_int(self.n_id)
        check(lib.lfvio_tracker_update_ids(self._h, C.byref(n)), "lfvio_tracker_update_ids")
        self.n_id = n.value
        self._refresh()

    def set_points(self, pts, ids, cnt):
        pts = np.ascontiguousarray(pts, np.float32).reshape(-1, 2)
        ids = np.ascontiguousarray(ids, np.int32)
        cnt = np.ascontiguousarray(cnt, np.int32)
        check(lib.lfvio_tracker_set_points(self._h, _as(pts, C.c_float), _as(ids, C.c_int), _as(cnt, C.c_int), len(ids)),
              "lfvio_tracker_set_points")

    @property
    def mask(self) -> np.ndarray:
        out = np.empty((self.rows, self.cols), np.uint8)
        check(lib.lfvio_tracker_get_mask(self._h, _as(out, C.c_uint8), out.strides[0]), "lfvio_tracker_get_mask")
        return out

    def cur_img(self, level: int = 0) -> np.ndarray:
        r, c = self.rows, self.cols
        for _ in range(level):
            r, c = (r + 1) // 2, (c + 1) // 2
        out = np.empty((r, c), np.uint8)
        check(lib.lfvio_tracker_get_image(self._h, level, _as(out, C.c_uint8), out.strides[0]), "lfvio_tracker_get_image")
        return out

    @property
    def stream(self) -> int:
        return lib.lfvio_tracker_stream(self._h) or 0

    def snapshot(self) -> dict:
        return {"cur_pts": self.cur_pts.copy(), "ids": self.ids.copy(), "track_cnt": self.track_cnt.copy(),
                "cur_un_pts_3d": self.cur_un_pts_3d.copy(), "pts_velocity_3d": self.pts_velocity_3d.copy()}


# ------------------------------------------------------------------------------------------------
# sliding-window back end (structs + handle live in backend.py)
# ------------------------------------------------------------------------------------------------
from . import backend  # noqa: E402

backend.bind(lib)


def solver_default_opts() -> "backend.SolverOpts":
    return backend.default_opts(lib)


def FeatureStore(window_size: int = 10, capacity: int = 2048, max_proj: int = 16384, min_parallax: float = 10.0 / 160.0,
                 init_depth: float = 5.0, device: int = 0) -> "backend.FeatureStore":
    return backend.FeatureStore(lib, check, window_size, capacity, max_proj, min_parallax, init_depth, device)


def Window(window_size: int = 10, max_landmarks: int = 1000, max_proj: int = 16384, device: int = 0) -> "backend.Window":
    return backend.Window(lib, check, window_size, max_landmarks, max_proj, device)

# live stage timing (bench.py)
lib.lfvio_tracker_stage_name.restype = C.c_char_p
lib.lfvio_tracker_stage_name.argtypes = [C.c_int]
lib.lfvio_window_stage_name.restype = C.c_char_p
lib.lfvio_window_stage_name.argtypes = [C.c_int]
lib.lfvio_tracker_profile_begin.argtypes = [C.c_void_p, C.c_int]
lib.lfvio_tracker_profile_end.argtypes = [C.c_void_p, C.POINTER(C.c_float), C.POINTER(C.c_int)]
lib.lfvio_window_profile_begin.argtypes = [C.c_void_p, C.c_int]
lib.lfvio_window_profile_end.argtypes = [C.c_void_p, C.POINTER(C.c_float), C.POINTER(C.c_int)]

# ------------------------------------------------------------------------------------------------
# the two steps before optimization(): IMU pre-integration and triangulation (prestep.py)
# ------------------------------------------------------------------------------------------------
from . import prestep  # noqa: E402

prestep.bind(lib)


def preintegrate(intervals, dt, acc, gyr, noise, g, device: int = 0):
    return prestep.preintegrate(lib, check, intervals, dt, acc, gyr, noise, g, device)


def triangulate(Ps, Rs, tic, ric, start_frame, obs_first, obs_pts, depth, init_depth: float = 5.0, device: int = 0):
    return prestep.triangulate(lib, check, Ps, Rs, tic, ric, start_frame, obs_first, obs_pts, depth, init_depth, device)


# ------------------------------------------------------------------------------------------------
# estimator initialisation (initial.py): gyro bias, gravity/scale alignment, relative pose, EPnP, global SfM
# ------------------------------------------------------------------------------------------------
from . import initial  # noqa: E402

initial.bind(lib)


def solve_gyro_bias(Rs, delta_q, J_rg, device: int = 0):
    return initial.solve_gyro_bias(lib, check, Rs, delta_q, J_rg, device)


def linear_alignment(Rs, Ts, delta_p, delta_v, sum_dt, tic, g_norm, device: int = 0):
    return initial.linear_alignment(lib, check, Rs, Ts, delta_p, delta_v, sum_dt, tic, g_norm, device)


def solve_relative_rt(corres, seed, frame, device: int = 0):
    return initial.solve_relative_rt(lib, check, corres, seed, frame, device)


def epnp(problems, device: int = 0):
    return initial.epnp(lib, check, problems, device)


def global_sfm(frame_num, l, relative_R, relative_T, obs_first, obs_frame, obs_pts, ba=True, device: int = 0):
    return initial.global_sfm(lib, check, frame_num, l, relative_R, relative_T, obs_first, obs_frame, obs_pts, ba, device)


def sfm_bundle_adjust(c_rot, c_trans, points, state, obs_first, obs_frame, obs_pts, l, device: int = 0):
    return initial.bundle_adjust(lib, check, c_rot, c_trans, points, state, obs_first, obs_frame, obs_pts, l, device)


def solve_relative_r(corres, seed, frame, device: int = 0):
    return initial.solve_relative_r(lib, check, corres, seed, frame, device)


def calibrate_ex_rotation(Rc, Rimu, Rc_g, device: int = 0):
    return initial.calibrate_ex_rotation(lib, check, Rc, Rimu, Rc_g, device)
