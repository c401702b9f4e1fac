"""Golden vectors of the estimator initialisation (tests/golden/initial_seed1.npz, made by tests/golden/make_initial_golden.py):
CPU — the oracle still reproduces them; B200 — the device reproduces them through the C-ABI with the stored inputs."""
import os

import numpy as np
import pytest

from oracle import initial as OI


@pytest.fixture(scope="module")
def g(golden_dir):
    return np.load(os.path.join(golden_dir, "initial_seed1.npz"))


def test_oracle_reproduces_golden(g):
    l, seed = int(g["l"]), int(g["ransac_seed"])
    ok, Rr, Tr, cnt, E = OI.solve_relative_rt(g["corres"], seed, l)
    assert ok == bool(g["rel_ok"]) and cnt == int(g["rel_cnt"])
    assert np.allclose(Rr, g["rel_R"], atol=1e-12) and np.allclose(Tr, g["rel_T"], atol=1e-12)
    sfm = OI.global_sfm(11, l, g["rel_R"], g["rel_T"], g["obs_first"], g["obs_frame"], g["obs_pts"])
    assert sfm["ok"] and np.array_equal(sfm["state"], g["sfm_state"])
    assert np.allclose(sfm["q"], g["sfm_q"], atol=1e-10) and np.allclose(sfm["T"], g["sfm_T"], atol=1e-10)
    assert sfm["ba"]["iterations"] == int(g["sfm_ba_iterations"])
    assert np.allclose(OI.solve_gyroscope_bias(g["al_Rs"], g["al_dq"], g["al_Jrg"]), g["al_dbg"], atol=1e-14)
    ok_a, gv, x = OI.linear_alignment(g["al_Rs"], g["al_Ts"], g["al_dp"], g["al_dv"], g["al_dts"], g["tic"], float(g["g_norm"]))
    assert ok_a == bool(g["al_ok"]) and np.allclose(gv, g["al_g"], atol=1e-10) and np.allclose(x, g["al_x"], atol=1e-9)
    Rp, Tp, ep = OI.epnp_compute_pose(g["pnp_pw"], g["pnp_us"])
    assert np.allclose(Rp, g["pnp_R"], atol=1e-11) and np.allclose(Tp, g["pnp_T"], atol=1e-11)


@pytest.mark.gpu
def test_device_reproduces_golden(lfvio_mod, gpu, g):
    l, seed = int(g["l"]), int(g["ransac_seed"])
    ok, Rr, Tr, cnt, E = lfvio_mod.solve_relative_rt(g["corres"], seed, l)
    assert ok == bool(g["rel_ok"]) and cnt == int(g["rel_cnt"])
    assert np.allclose(Rr, g["rel_R"], atol=1e-8) and np.allclose(Tr, g["rel_T"], atol=1e-8)
    pre = lfvio_mod.global_sfm(11, l, g["rel_R"], g["rel_T"], g["obs_first"], g["obs_frame"], g["obs_pts"], ba=False)
    assert pre["ok"] and np.array_equal(pre["state"], g["sfm_state"])
    assert np.allclose(pre["q"], g["pre_ba_q"], atol=1e-7) and np.allclose(pre["T"], g["pre_ba_T"], atol=1e-7)
    sfm = lfvio_mod.global_sfm(11, l, g["rel_R"], g["rel_T"], g["obs_first"], g["obs_frame"], g["obs_pts"])
    assert sfm["ok"] and sfm["summary"].ba_iterations == int(g["sfm_ba_iterations"])
    assert np.allclose(sfm["summary"].costs(), g["sfm_ba_costs"], rtol=1e-5, atol=1e-15)
    assert np.allclose(sfm["q"], g["sfm_q"], atol=1e-7) and np.allclose(sfm["T"], g["sfm_T"], atol=1e-7)
    st = g["sfm_state"]
    assert np.allclose(sfm["points"][st], g["sfm_points"][st], rtol=1e-4, atol=1e-4)
    assert np.allclose(lfvio_mod.solve_gyro_bias(g["al_Rs"], g["al_dq"], g["al_Jrg"]), g["al_dbg"], rtol=1e-9, atol=1e-13)
    ok_a, gv, x = lfvio_mod.linear_alignment(g["al_Rs"], g["al_Ts"], g["al_dp"], g["al_dv"], g["al_dts"], g["tic"], float(g["g_norm"]))
    assert ok_a == bool(g["al_ok"]) and np.allclose(gv, g["al_g"], rtol=1e-8, atol=1e-8) and np.allclose(x, g["al_x"], rtol=1e-7, atol=1e-8)
    Rp, Tp, ep = lfvio_mod.epnp([(g["pnp_pw"], g["pnp_us"])])
    assert np.allclose(Rp[0], g["pnp_R"], atol=1e-8) and np.allclose(Tp[0], g["pnp_T"], atol=1e-8)
