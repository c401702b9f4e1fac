"""OCAMCamera::liftProjective — oracle vs known answers (CPU) and CUDA kernel vs oracle (GPU, bit-exact f64)."""
import json
import os

import numpy as np
import pytest

from oracle import camera as ocam_oracle


def _kat(golden_dir):
    return json.load(open(os.path.join(golden_dir, "lift_kat.json")))


def test_oracle_matches_known_answers(golden_dir):
    from lfvio.config import PAL_UP

    k = _kat(golden_dir)
    got = ocam_oracle.lift_projective(PAL_UP, np.array(k["uv"]))
    want = np.array(k["xyz"], dtype=np.float64)
    np.testing.assert_allclose(got, want, rtol=k["rel_tol"], atol=1e-6)
    # outer annulus maps to the negative half-space ("negative plane")
    assert got[3, 2] < 0 and got[4, 2] < 0 and got[1, 2] > 0


def test_oracle_z_sign_change_equals_inv_poly_p0():
    from lfvio.config import PAL_UP

    phi = np.linspace(376.0, 378.0, 20001)
    uv = np.stack([PAL_UP["cx"] + phi, np.full_like(phi, PAL_UP["cy"])], axis=1)
    z = ocam_oracle.lift_projective(PAL_UP, uv)[:, 2]
    cross = phi[np.argmax(z < 0)]
    assert abs(cross - PAL_UP["inv_poly"][0]) < 2e-3  # ScaramuzzaCamera.cc:55 consistency check


def test_oracle_lift_then_project_roundtrip():
    from lfvio.config import PAL_UP

    rng = np.random.default_rng(3)
    ang = rng.uniform(0, 2 * np.pi, 500)
    r = rng.uniform(160, 500, 500)
    uv = np.stack([640 + r * np.cos(ang), 490 + r * np.sin(ang)], axis=1)
    back = ocam_oracle.space_to_plane(PAL_UP, ocam_oracle.lift_sphere(PAL_UP, uv))
    assert np.abs(back - uv).max() < 0.1  # the 14-term inverse polynomial is a fit, not an inverse


@pytest.mark.gpu
@pytest.mark.parametrize("n", [0, 1, 5, 150, 2000, 100003])
@pytest.mark.parametrize("normalize", [False, True])
def test_cuda_lift_bit_exact(lfvio_mod, gpu, n, normalize):
    cam = lfvio_mod.PAL_UP
    rng = np.random.default_rng(n + 17)
    uv = np.stack([rng.uniform(0, 1280, n), rng.uniform(0, 960, n)], axis=1).astype(np.float32).astype(np.float64)
    got = lfvio_mod.lift_projective(lfvio_mod.Ocam.from_dict(cam), uv, normalize=normalize, device=gpu)
    want = ocam_oracle.lift_sphere(cam, uv) if normalize else ocam_oracle.lift_projective(cam, uv)
    assert got.shape == want.shape
    assert np.array_equal(got.view(np.uint64), want.view(np.uint64)), (
        f"max |d| = {np.abs(got - want).max() if n else 0}")


@pytest.mark.gpu
def test_cuda_lift_known_answers_and_scaled_camera(lfvio_mod, gpu, golden_dir):
    k = _kat(golden_dir)
    got = lfvio_mod.lift_projective(lfvio_mod.Ocam.from_dict(lfvio_mod.PAL_UP), np.array(k["uv"]), device=gpu)
    np.testing.assert_allclose(got, np.array(k["xyz"]), rtol=k["rel_tol"], atol=1e-6)
    # affine-skewed + scaled calibration (configs 3/4) against the oracle, bit-exact
    cam = lfvio_mod.scaled_camera(3.0)
    cam.update(C=1.000045, D=0.000016, E=-0.000022)
    rng = np.random.default_rng(5)
    uv = np.stack([rng.uniform(0, 3840, 4096), rng.uniform(0, 2880, 4096)], axis=1)
    got = lfvio_mod.lift_projective(lfvio_mod.Ocam.from_dict(cam), uv, normalize=True, device=gpu)
    want = ocam_oracle.lift_sphere(cam, uv)
    assert np.array_equal(got.view(np.uint64), want.view(np.uint64))
