"""numpy f64 restatement of the Scaramuzza/OCamCalib camera — ORACLE (test infrastructure only).

Follows camera_model/src/camera_models/ScaramuzzaCamera.cc of the reference:
  lift_projective  <- OCAMCamera::liftProjective  :622-645 (m_inv_scale :197)
  lift_sphere      <- OCAMCamera::liftSphere      :609-614
  space_to_plane   <- OCAMCamera::spaceToPlane    :653-674
Arithmetic is element-wise IEEE f64 in source order (numpy never fuses multiply-add), polynomial
accumulated low-order-first with a running power exactly as the reference loop does.

Pinned by the known-answer vectors of SURVEY.md §8c (tests/golden/lift_kat.json) — the reference has
no tests of its own for this function.
"""
from __future__ import annotations

import numpy as np


def inv_scale(cam: dict) -> float:
    return 1.0 / (cam["C"] - cam["D"] * cam["E"])


def lift_projective(cam: dict, uv: np.ndarray) -> np.ndarray:
    uv = np.asarray(uv, dtype=np.float64).reshape(-1, 2)
    xc0 = uv[:, 0] - cam["cx"]
    xc1 = uv[:, 1] - cam["cy"]
    s = inv_scale(cam)
    xa = s * (xc0 - cam["D"] * xc1)
    ya = s * (-cam["E"] * xc0 + cam["C"] * xc1)
    phi = np.sqrt(xa * xa + ya * ya)
    phi_i = np.ones_like(phi)
    z = np.zeros_like(phi)
    for i in range(5):
        z = z + phi_i * cam["poly"][i]
        phi_i = phi_i * phi
    return np.stack([xa, ya, -z], axis=1)


def lift_sphere(cam: dict, uv: np.ndarray) -> np.ndarray:
    P = lift_projective(cam, uv)
    nrm = np.sqrt((P[:, 0] * P[:, 0] + P[:, 1] * P[:, 1]) + P[:, 2] * P[:, 2])
    return P / nrm[:, None]


def space_to_plane(cam: dict, P: np.ndarray) -> np.ndarray:
    P = np.asarray(P, dtype=np.float64).reshape(-1, 3)
    norm = np.sqrt(P[:, 0] * P[:, 0] + P[:, 1] * P[:, 1])
    theta = np.arctan2(-P[:, 2], norm)
    rho = np.zeros_like(theta)
    theta_i = np.ones_like(theta)
    for i in range(20):
        rho = rho + theta_i * cam["inv_poly"][i]
        theta_i = theta_i * theta
    inv = 1.0 / norm
    xn0 = P[:, 0] * inv * rho
    xn1 = P[:, 1] * inv * rho
    return np.stack([xn0 * cam["C"] + xn1 * cam["D"] + cam["cx"], xn0 * cam["E"] + xn1 + cam["cy"]], axis=1)
