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
