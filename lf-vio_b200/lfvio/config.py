"""Calibration / configuration fixtures.

Numbers are the PAL-up calibration and tracker/IMU settings of the reference's only panoramic rig
(config/mindvision/mindvision.yaml:48-80 camera, :110-122 tracker + annulus, :133-152 solver/IMU/td).
YAML key names are kept verbatim where they exist.
"""
from __future__ import annotations

import copy

PAL_UP = {
    "model_type": "scaramuzza",
    "camera_name": "pal",
    "image_width": 1280,
    "image_height": 960,
    # poly_parameters p0..p4
    "poly": [-2.445239e02, 0.0, 1.748610e-03, -1.757770e-06, 4.475965e-09],
    # inv_poly_parameters p0..p19
    "inv_poly": [376.845565, 246.746504, 19.035187, 23.840497, 18.991943, 6.066253, 1.560387, 5.854280,
                 3.458320, -1.995166, -1.509264, 1.089614, 1.340245, 0.255323, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0],
    # affine_parameters
    "C": 1.0, "D": 0.0, "E": 0.0, "cx": 645.107791, "cy": 486.025172,
}

TRACKER = {
    "max_cnt": 150,  # BASELINE.json configs use 150; mindvision.yaml:110 ships 200
    "min_dist": 20,
    "freq": 10,
    "equalize": 1,
    "fisheye": 0,
    "center_x": 640.991301,
    "center_y": 490.937512,
    "max_r": 500,
    "min_r": 160,
}

ESTIMATOR = {
    "max_solver_time": 0.04,
    "max_num_iterations": 8,
    "keyframe_parallax": 10.0,
    "acc_n": 0.02, "gyr_n": 0.01, "acc_w": 0.04, "gyr_w": 0.001, "g_norm": 9.81007,
    "estimate_td": 1, "td": -0.008,
    "estimate_extrinsic": 1,
    "rolling_shutter": 0, "rolling_shutter_tr": 0.0,
    "extrinsicRotation": [-1.0, 0.0, 0.0, 0.0, -1.0, 0.0, 0.0, 0.0, 1.0],
    "extrinsicTranslation": [0.0, 0.0, 0.03],
    "focal_length": 160.0,  # vins_estimator/src/parameters.h:11
}


def scaled_camera(scale: float) -> dict:
    """Calibration for a sensor with `scale`x the pixel pitch count:
    centre and image size scale by s, poly p_i' = p_i / s^(i-1) so that rays are unchanged up to
    the common factor s (z' = s z at phi' = s phi)."""
    c = copy.deepcopy(PAL_UP)
    s = float(scale)
    c["image_width"] = int(round(c["image_width"] * s))
    c["image_height"] = int(round(c["image_height"] * s))
    c["cx"] *= s
    c["cy"] *= s
    c["poly"] = [p / s ** (i - 1) for i, p in enumerate(c["poly"])]
    c["inv_poly"] = [p * s for p in c["inv_poly"]]
    return c


def tracker_defaults(scale: float = 1.0, max_cnt: int | None = None) -> dict:
    t = copy.deepcopy(TRACKER)
    s = float(scale)
    t["center_x"] *= s
    t["center_y"] *= s
    t["max_r"] = int(round(t["max_r"] * s))
    t["min_r"] = int(round(t["min_r"] * s))
    # min_dist and the 41x41 LK window are NOT scaled: more pixels => room for more features
    if max_cnt is not None:
        t["max_cnt"] = max_cnt
    return t
