"""FeatureManager of the C++ host mirror (pure host bookkeeping — no CUDA device needed) against the Python restatement of
the reference's logic (oracle/feature_manager.py): key-frame decision, correspondences, and the three ways a track list
slides (removeBack, removeFront, removeBackShiftDepth), over a synthetic stream with both kinds of frames."""
import ctypes as C
import os

import numpy as np
import pytest

from conftest import ROOT
from lfvio import violoop
from oracle.feature_manager import FeatureManagerOracle

W = 10


@pytest.fixture(scope="module")
def host():
    lib = C.CDLL(os.path.join(ROOT, "lf-vio_b200", "liblfvio_host.so"))
    lib.lfvio_host_fm_create.restype = C.c_void_p
    lib.lfvio_host_fm_create.argtypes = [C.c_int, C.c_double]
    lib.lfvio_host_fm_destroy.argtypes = [C.c_void_p]
    lib.lfvio_host_fm_add.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_double]
    lib.lfvio_host_fm_corresponding.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int]
    lib.lfvio_host_fm_slide.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p]
    lib.lfvio_host_fm_snapshot.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
    return lib


def _snapshot(host, h, set_depth=None):
    cap = 4096
    meta, depth, cnt = np.zeros((cap, 3), np.int32), np.zeros(cap), C.c_int(0)
    sd = None if set_depth is None else np.ascontiguousarray(set_depth, np.float64).ctypes.data
    n = host.lfvio_host_fm_snapshot(h, meta.ctypes.data, depth.ctypes.data, cap, sd, C.byref(cnt))
    return [(int(meta[i, 0]), int(meta[i, 1]), int(meta[i, 2]), float(depth[i])) for i in range(n)], cnt.value


@pytest.mark.parametrize("min_parallax", [0.0, 0.45, 1e9])
def test_feature_manager_stream(host, min_parallax):
    """min_parallax 0: every frame is a key frame (removeBack / removeBackShiftDepth); 1e9: only the frames the reference
    forces to be key frames (fewer than 20 tracked features or no feature old enough, :74-88) are; 0.45 (inside the range of
    this world's per-frame parallax, 0.25 .. 0.62 in the reference's acos * 10 units): a mixture."""
    world = violoop.World(seed=6, n_frames=34, max_feat=60)
    h = host.lfvio_host_fm_create(W, min_parallax)
    ref = FeatureManagerOracle(W, min_parallax)
    rng = np.random.default_rng(0)
    frame_count, kinds = 0, []
    for k in range(world.n_frames):
        obs = world.frames[k]
        ids = np.array(sorted(obs), np.int32)
        b = np.ascontiguousarray([obs[i][0] for i in ids])
        key = host.lfvio_host_fm_add(h, frame_count, ids.ctypes.data, b.ctypes.data, len(ids), 0.0)
        key_o = ref.add_feature_check_parallax(frame_count, {int(i): b[j] for j, i in enumerate(ids)})
        assert bool(key) == key_o, k
        snap, cnt = _snapshot(host, h)
        assert [s[:3] for s in snap] == [s[:3] for s in ref.snapshot()] and cnt == ref.get_feature_count(), k
        if frame_count >= 1:
            out = np.zeros((512, 6))
            n = host.lfvio_host_fm_corresponding(h, frame_count - 1, frame_count, out.ctypes.data, 512)
            assert np.array_equal(out[:n], ref.get_corresponding(frame_count - 1, frame_count))
        if frame_count < W:
            frame_count += 1
            continue
        kinds.append(bool(key))
        if key and k % 2 == 0:                        # MARGIN_OLD while initialising: removeBack
            host.lfvio_host_fm_slide(h, 0, 0, None)
            ref.remove_back()
        elif key:                                     # MARGIN_OLD when running: removeBackShiftDepth with some poses / depths
            depth = rng.uniform(2.0, 9.0, len(ref.feature))
            _snapshot(host, h, set_depth=depth)
            for f, d in zip(ref.feature, depth):
                f.depth = float(d)
            mR, nR = world.truth(k - W)[1], world.truth(k - W + 1)[1]
            mP, nP = world.truth(k - W)[0], world.truth(k - W + 1)[0]
            pose = np.concatenate([mR.ravel(), mP, nR.ravel(), nP])
            host.lfvio_host_fm_slide(h, 2, 0, pose.ctypes.data)
            ref.remove_back_shift_depth(mR, mP, nR, nP)
        else:                                         # MARGIN_SECOND_NEW: removeFront
            host.lfvio_host_fm_slide(h, 1, frame_count, None)
            ref.remove_front(frame_count)
        snap, cnt = _snapshot(host, h)
        rs = ref.snapshot()
        assert [s[:3] for s in snap] == [s[:3] for s in rs], k
        assert np.allclose([s[3] for s in snap], [s[3] for s in rs], rtol=1e-13, atol=0), k
    host.lfvio_host_fm_destroy(h)
    if min_parallax == 0.0:
        assert all(kinds)
    else:
        assert any(kinds) and not all(kinds), kinds
