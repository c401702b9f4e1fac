"""The committed outputs of the REFERENCE'S OWN code (tests/golden/ref_backend.npz, written by make_ref_golden.py from
oracle/_ref = /root/reference's factor / marginalisation / pre-integration / alignment sources compiled unmodified) against
(a) the self-written oracle — CPU, runs everywhere — and (b) the CUDA path through the C-ABI — `-m gpu`.
Needs neither /root/reference nor oracle/_ref at run time.

Stated fp64 tolerances (both sides are fp64; differences are FMA contraction, summation order and, for the IMU factor, the
inversion of a covariance with entries 1e-9 .. 1e-4): projection residual / Jacobian 1e-9, IMU 1e-7, prior information
(J0^T J0, J0^T r0) 1e-7 relative to the largest entry, alignment 1e-7.
"""
import numpy as np
import pytest

import ref_golden as RG
from lfvio import backend as B
from oracle import backend as ob
from oracle import initial as OI
from oracle import prestep as OP
from oracle.ref import canonical_information


def rel(a, b):
    return float(np.abs(np.asarray(a) - np.asarray(b)).max() / max(np.abs(np.asarray(b)).max(), 1e-300))


@pytest.fixture(scope="module")
def g():
    return RG.load()


def _check_eval(e, g, key, tp, ti):
    assert rel(e.res_proj, g[key + "res_proj"]) < tp and rel(e.jac_proj, g[key + "jac_proj"]) < tp
    assert rel(e.res_imu, g[key + "res_imu"]) < ti and rel(e.jac_imu, g[key + "jac_imu"]) < ti
    assert rel(e.res_prior, g[key + "res_prior"]) < ti
    assert abs(e.cost.value - g[key + "cost"][0]) < ti * g[key + "cost"][0]


def _check_prior(p, g, flag, F, tol):
    A, gg, x0 = canonical_information(p, F)
    assert p.n == g[f"marg{flag}_m_n"][1]
    assert list(sorted(x0)) == list(g[f"marg{flag}_ids"])
    assert np.abs(np.concatenate([x0[k] for k in sorted(x0)]) - g[f"marg{flag}_x0"]).max() < 1e-12
    assert rel(A, g[f"marg{flag}_A"]) < tol and rel(gg, g[f"marg{flag}_g"]) < tol


# ------------------------------------------------------------------------------------------------ (a) the oracle, on the CPU
@pytest.mark.parametrize("robust", [0, 1])
@pytest.mark.parametrize("td", [1, 0])
def test_oracle_factors_equal_reference(g, robust, td):
    w = RG.window(g, ob.default_opts())
    w.opts.estimate_td = td
    _check_eval(ob.evaluate(w, bool(robust)), g, f"eval_r{robust}_td{td}_", 1e-11, 1e-12)


@pytest.mark.parametrize("flag", [B.MARGIN_OLD, B.MARGIN_SECOND_NEW])
def test_oracle_prior_equals_reference(g, flag):
    w = RG.window(g, ob.default_opts())
    _check_prior(ob.marginalize(w, flag), g, flag, w.F, 2e-9)


def test_oracle_preintegration_equals_reference(g):
    ivs, dt, acc, gyr, noise, G, want = RG.preintegration_inputs(g)
    for iv, f in zip(ivs, want):
        ib = OP.IntegrationOracle(list(iv.acc_0), list(iv.gyr_0), list(iv.ba), list(iv.bg), noise)
        for k in range(iv.first, iv.first + iv.count):
            ib.push_back(dt[k], acc[k], gyr[k])
        assert np.abs(np.array(f.delta_p) - ib.delta_p).max() < 1e-15 and np.abs(np.array(f.delta_q) - ib.delta_q).max() < 1e-15
        assert rel(np.array(f.jacobian).reshape(15, 15), ib.jacobian) < 1e-13
        assert rel(np.array(f.covariance).reshape(15, 15), ib.covariance) < 1e-13


def test_oracle_alignment_equals_reference(g):
    tic, gn = g["al_tic_gnorm"][:3], g["al_tic_gnorm"][3]
    assert np.abs(OI.solve_gyroscope_bias(g["al_R"], g["al_dq"], g["al_Jrg"]) - g["al_bg"]).max() < 1e-13
    ok, gv, x = OI.linear_alignment(g["al_R"], g["al_T"], g["al_dp"], g["al_dv"], g["al_dt"], tic, gn)
    assert ok == bool(g["al_ok"][0]) and np.abs(gv - g["al_g"]).max() < 1e-9 and np.abs(x - g["al_x"]).max() < 1e-8


def test_oracle_ransac_equals_reference(g):
    import lfvio
    from oracle import ransac as R

    b1, b2 = R.lift_pair(lfvio.PAL_UP, g["rs_cur"], g["rs_forw"])
    E, st, _, _ = R.find_fundamental(b1, b2, int(g["rs_seed_frame"][0]), int(g["rs_seed_frame"][1]))
    assert np.array_equal(st, g["rs_status"])
    assert np.abs(E - np.sign((E * g["rs_E"]).sum()) * g["rs_E"]).max() < 1e-12


# ------------------------------------------------------------------------------------------------ (b) the CUDA path
@pytest.fixture(scope="module")
def win(lfvio_mod, gpu):
    return lfvio_mod.Window(10, 1000, 16384, device=gpu)


@pytest.mark.gpu
@pytest.mark.parametrize("robust", [0, 1])
@pytest.mark.parametrize("td", [1, 0])
def test_device_factors_equal_reference(lfvio_mod, win, g, robust, td):
    w = RG.window(g, lfvio_mod.solver_default_opts())
    w.opts.estimate_td = td
    _check_eval(win.evaluate(w, bool(robust)), g, f"eval_r{robust}_td{td}_", 1e-9, 1e-7)


@pytest.mark.gpu
@pytest.mark.parametrize("flag", [B.MARGIN_OLD, B.MARGIN_SECOND_NEW])
def test_device_prior_equals_reference(lfvio_mod, win, g, flag):
    w = RG.window(g, lfvio_mod.solver_default_opts())
    _check_prior(win.marginalize(w, flag), g, flag, w.F, 1e-7)


@pytest.mark.gpu
def test_device_preintegration_equals_reference(lfvio_mod, gpu, g):
    ivs, dt, acc, gyr, noise, G, want = RG.preintegration_inputs(g)
    got, _ = lfvio_mod.preintegrate(ivs, dt, acc, gyr, noise, G, device=gpu)
    for a, b in zip(got, want):
        assert abs(a.sum_dt - b.sum_dt) < 1e-15
        assert np.abs(np.array(a.delta_p) - np.array(b.delta_p)).max() < 1e-13
        assert np.abs(np.array(a.delta_q) - np.array(b.delta_q)).max() < 1e-13
        assert np.abs(np.array(a.delta_v) - np.array(b.delta_v)).max() < 1e-13
        assert rel(np.array(a.jacobian), np.array(b.jacobian)) < 1e-11
        assert rel(np.array(a.covariance), np.array(b.covariance)) < 1e-11


@pytest.mark.gpu
def test_device_ransac_equals_reference(lfvio_mod, gpu, g):
    """lfvio_reject_with_f (the tracker's rejectWithF kernels) on the stored point pairs: same status as the reference's code."""
    st = lfvio_mod.reject_with_f(lfvio_mod.PAL_UP, g["rs_cur"], g["rs_forw"], int(g["rs_seed_frame"][0]), int(g["rs_seed_frame"][1]))
    status = st[0] if isinstance(st, tuple) else st
    assert np.array_equal(np.asarray(status).astype(np.uint8), g["rs_status"])


@pytest.mark.gpu
def test_device_alignment_equals_reference(lfvio_mod, gpu, g):
    tic, gn = g["al_tic_gnorm"][:3], g["al_tic_gnorm"][3]
    assert np.abs(lfvio_mod.solve_gyro_bias(g["al_R"], g["al_dq"], g["al_Jrg"], device=gpu) - g["al_bg"]).max() < 1e-12
    ok, gv, x = lfvio_mod.linear_alignment(g["al_R"], g["al_T"], g["al_dp"], g["al_dv"], g["al_dt"], tic, gn, device=gpu)
    assert ok == bool(g["al_ok"][0]) and np.abs(gv - g["al_g"]).max() < 1e-7 and np.abs(x - g["al_x"]).max() < 1e-7
