"""The C-ABI library loads on a CPU-only box and exports every symbol include/lfvio.h declares."""
import ctypes
import os
import re

from conftest import ROOT


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "lfvio.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(lfvio_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_entry_points():
    syms = declared_symbols()
    assert "lfvio_lift_projective" in syms and "lfvio_last_error" in syms


def test_library_exports_every_declared_symbol(lfvio_mod):
    lib = ctypes.CDLL(lfvio_mod.LIB_PATH)
    missing = [s for s in declared_symbols() if not hasattr(lib, s)]
    assert not missing, f"declared in include/lfvio.h but not exported: {missing}"


def test_version_and_error_strings(lfvio_mod):
    assert b"sm_100a" in lfvio_mod.lib.lfvio_version()
    assert lfvio_mod.lib.lfvio_device_count() >= 0


def test_bad_arguments_fail_loudly(lfvio_mod):
    st = lfvio_mod.lib.lfvio_lift_projective(None, None, None, 4, 0, 0)
    assert st == -2
    assert b"bad arguments" in lfvio_mod.lib.lfvio_last_error()


def test_feature_store_arguments_and_missing_device_fail_loudly(lfvio_mod):
    """lfvio_fm_*: argument errors are reported before any device work, and without a CUDA device the handle is not created
    (there is no host-side fallback store)."""
    import ctypes as C

    L = lfvio_mod.lib
    h = C.c_void_p()
    assert L.lfvio_fm_create(1, 16, 16, 0.1, 5.0, 0, C.byref(h)) == -2 and not h.value            # window_size < 2
    assert L.lfvio_fm_create(40, 16, 16, 0.1, 5.0, 0, C.byref(h)) == -2 and not h.value           # more than 32 frames
    assert L.lfvio_fm_create(10, 0, 16, 0.1, 5.0, 0, C.byref(h)) == -2 and b"bad arguments" in L.lfvio_last_error()
    assert L.lfvio_fm_add_frame(None, 0, None, None, 0, 0.0, None, None) == -2
    assert L.lfvio_fm_build_factors(None, None, None) == -2 and b"null handle" in L.lfvio_last_error()
    assert L.lfvio_window_optimize_full_fm(None, None, None, None, None, 0, None, None) == -2
    if L.lfvio_device_count() == 0:
        assert L.lfvio_fm_create(10, 16, 16, 0.1, 5.0, 0, C.byref(h)) == -1 and not h.value       # LFVIO_ERR_CUDA, no handle


def test_product_never_imports_oracle():
    """The product tree must not reference the oracle (no CPU fallback)."""
    bad = []
    for dp, _, fns in os.walk(os.path.join(ROOT, "lf-vio_b200")):
        if "build" in dp:
            continue
        for fn in fns:
            if fn.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                src = open(os.path.join(dp, fn), errors="replace").read()
                if re.search(r"^\s*(from|import)\s+oracle\b|#include\s+\".*oracle/", src, flags=re.M):
                    bad.append(os.path.join(dp, fn))
    assert not bad, bad


def test_tum_writer_format():
    """The trajectory line of pubOdometry (vins_estimator/src/utility/visualization.cpp:173-179): fixed notation, 12
    decimals, "stamp px py pz qx qy qz qw" — pure host code, no GPU."""
    import numpy as np

    host = ctypes.CDLL(os.path.join(ROOT, "lf-vio_b200", "liblfvio_host.so"))
    host.lfvio_host_tum_line.argtypes = [ctypes.c_double, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_char_p, ctypes.c_int]
    buf = ctypes.create_string_buffer(512)
    p, q = np.array([1.5, -2.25, 3.0]), np.array([0.1, 0.2, 0.3, 0.9])
    n = host.lfvio_host_tum_line(1403636579.5, p.ctypes.data, q.ctypes.data, buf, 512)
    line = buf.value.decode()
    assert n == len(line) and line.endswith("\n")
    assert line == "1403636579.500000000000 1.500000000000 -2.250000000000 3.000000000000 0.100000000000 0.200000000000 0.300000000000 0.900000000000\n"
