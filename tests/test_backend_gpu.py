"""GPU parity tests of the sliding-window back end through the C-ABI (include/lfvio.h) against the fp64 oracle.

Everything is fp64 on both sides; differences come from FMA contraction and summation order (atomics), amplified
by the conditioning of the IMU information (covariance entries 1e-9..1e-4 are inverted, imu_factor.h:64).
Stated tolerances: residuals/Jacobians 1e-7 relative to the block's largest entry; per-iteration cost 1e-6 relative;
converged states 1e-6 absolute (metres / quaternion components / inverse depth).
"""
import numpy as np
import pytest

from lfvio import backend as B
from lfvio import synthwin
from oracle import backend as ob

pytestmark = pytest.mark.gpu


def rel_err(a, b):
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-300)


@pytest.fixture(scope="module")
def win(lfvio_mod, gpu):
    return lfvio_mod.Window(10, 1000, 16384, device=gpu)


@pytest.fixture(scope="module")
def problem(lfvio_mod):
    return synthwin.make_window(seed=1, opts=lfvio_mod.solver_default_opts(), perturb=0.3)


def test_default_opts_match_oracle(lfvio_mod):
    a, b = lfvio_mod.solver_default_opts(), ob.default_opts()
    assert bytes(a) == bytes(b)


@pytest.mark.parametrize("robust", [False, True])
def test_evaluate_matches_oracle(win, problem, robust):
    g, o = win.evaluate(problem, robust), ob.evaluate(problem, robust)
    assert rel_err(g.res_proj, o.res_proj) < 1e-9
    assert rel_err(g.jac_proj, o.jac_proj) < 1e-9
    assert rel_err(g.res_imu, o.res_imu) < 1e-7
    assert rel_err(g.jac_imu, o.jac_imu) < 1e-7
    assert abs(g.cost.value - o.cost.value) < 1e-7 * o.cost.value


def test_evaluate_variants(lfvio_mod, win):
    """estimate_td = 0 (ProjectionFactor), fixed extrinsics, rolling shutter term, no IMU factor for long gaps."""
    for kw in (dict(estimate_td=0), dict(estimate_extrinsic=0), dict(tr=0.03), dict(long_gap=True)):
        opts = lfvio_mod.solver_default_opts()
        w = synthwin.make_window(seed=5, opts=opts, perturb=0.5, estimate_td=kw.get("estimate_td", 1))
        if "estimate_extrinsic" in kw:
            w.opts.estimate_extrinsic = 0
        if "tr" in kw:
            w.opts.tr = kw["tr"]
        if kw.get("long_gap"):
            w.imu[4].sum_dt = 11.0  # estimator.cpp:720-721: pre-integrations longer than 10 s are skipped
        g, o = win.evaluate(w, True), ob.evaluate(w, True)
        assert rel_err(g.res_proj, o.res_proj) < 1e-9, kw
        assert rel_err(g.jac_proj, o.jac_proj) < 1e-9, kw
        assert rel_err(g.res_imu, o.res_imu) < 1e-7, kw
        assert rel_err(g.jac_imu, o.jac_imu) < 1e-7, kw
        assert abs(g.cost.value - o.cost.value) < 1e-7 * o.cost.value, kw


def _solve_both(win, w):
    a, b = w.clone(), w.clone()
    sg, so = win.optimize(a), ob.optimize(b)
    return a, sg, b, so


@pytest.mark.parametrize("seed,perturb", [(1, 0.3), (2, 1.0), (3, 0.1)])
def test_optimize_matches_oracle(lfvio_mod, win, seed, perturb):
    w = synthwin.make_window(seed=seed, opts=lfvio_mod.solver_default_opts(), perturb=perturb)
    a, sg, b, so = _solve_both(win, w)
    assert sg.iterations == so.iterations and sg.num_successful_steps == so.num_successful_steps
    assert sg.termination == so.termination
    for i in range(so.iterations + 1):
        assert abs(sg.cost[i] - so.cost[i]) <= 1e-6 * abs(so.cost[i]) + 1e-9, (i, sg.cost[i], so.cost[i])
        assert abs(sg.radius[i] - so.radius[i]) <= 1e-6 * so.radius[i], (i, sg.radius[i], so.radius[i])
        assert sg.accepted[i] == so.accepted[i]
    assert np.abs(a.pose - b.pose).max() < 1e-6
    assert np.abs(a.sb - b.sb).max() < 1e-5
    assert np.abs(a.ex - b.ex).max() < 1e-6
    assert np.abs(a.feat - b.feat).max() < 1e-6
    assert abs(a.td[0] - b.td[0]) < 1e-7


def test_optimize_fixed_iterations_and_truth(lfvio_mod, win):
    opts = lfvio_mod.solver_default_opts()
    opts.fixed_iterations = 1
    w = synthwin.make_window(seed=4, opts=opts, perturb=0.3)
    a, sg, b, so = _solve_both(win, w)
    assert sg.iterations == so.iterations == 8
    assert np.abs(a.pose - b.pose).max() < 1e-6
    assert np.abs(a.pose[:, :3] - w.truth["P"]).max() < 0.05


def test_optimize_without_td_and_fixed_extrinsic(lfvio_mod, win):
    opts = lfvio_mod.solver_default_opts()
    w = synthwin.make_window(seed=6, opts=opts, perturb=0.3, estimate_td=0)
    w.opts.estimate_extrinsic = 0
    a, sg, b, so = _solve_both(win, w)
    assert sg.iterations == so.iterations
    assert np.array_equal(a.ex, w.ex) and a.td[0] == w.td[0]  # constant blocks are untouched
    assert np.abs(a.pose - b.pose).max() < 1e-6


def test_device_resident_solve_repeats(lfvio_mod, win, problem):
    """upload once, solve twice from the uploaded state: identical results, equal to the host-buffer call."""
    ref = problem.clone()
    s_ref = win.optimize(ref)
    win.upload(problem)
    outs = []
    for _ in range(2):
        win.optimize_dev(False)
        o = problem.clone()
        s = win.download(o)
        outs.append((o, s))
    for o, s in outs:
        assert s.iterations == s_ref.iterations
        assert np.abs(o.pose - ref.pose).max() < 1e-7 and np.abs(o.feat - ref.feat).max() < 1e-7  # fp64 atomics => last-digit noise


def test_bad_problem_is_rejected(lfvio_mod, win, problem):
    w = problem.clone()
    a, b = w.proj[3].landmark, w.proj[400].landmark
    w.proj[3].landmark, w.proj[400].landmark = b, a  # breaks the grouping (the factor array is shared with `problem`)
    try:
        with pytest.raises(lfvio_mod.LfvioError):
            win.optimize(w)
    finally:
        w.proj[3].landmark, w.proj[400].landmark = a, b


# ---- K14: marginalisation ------------------------------------------------------------------------------
def _info_close(pg, po, tol=1e-7):
    assert pg.n == po.n and pg.n_blocks == po.n_blocks
    assert list(pg.block_id[: pg.n_blocks]) == list(po.block_id[: po.n_blocks])
    assert list(pg.block_idx[: pg.n_blocks]) == list(po.block_idx[: po.n_blocks])
    Ag, bg = pg.information()
    Ao, bo = po.information()
    assert np.abs(Ag - Ao).max() < tol * np.abs(Ao).max(), np.abs(Ag - Ao).max() / np.abs(Ao).max()
    assert np.abs(bg - bo).max() < 1e-6 * max(np.abs(bo).max(), 1.0)
    nx = sum(7 if (i < 2 * 11 and i % 2 == 0) or i == 22 else (9 if i < 22 else 1) for i in pg.block_id[: pg.n_blocks])
    assert np.array_equal(pg.x0[:nx], po.x0[:nx])  # keep_block_data are plain copies


def test_marginalize_old_matches_oracle(lfvio_mod, win):
    w = synthwin.make_window(seed=1, opts=lfvio_mod.solver_default_opts(), perturb=0.3)
    ob.optimize(w)
    pg, po = win.marginalize(w, B.MARGIN_OLD), ob.marginalize(w, B.MARGIN_OLD)
    assert pg.n == 6 * 10 + 9 + 6 + 1
    _info_close(pg, po)


def test_marginalize_chain_with_prior_and_second_new(lfvio_mod, win):
    """Window k -> prior -> window k+1 (prior factor inside the solve) -> MARGIN_OLD and MARGIN_SECOND_NEW again."""
    w = synthwin.make_window(seed=2, opts=lfvio_mod.solver_default_opts(), perturb=0.3)
    win.optimize(w)
    prior = win.marginalize(w, B.MARGIN_OLD)
    nw = synthwin.make_window(seed=2, opts=lfvio_mod.solver_default_opts(), perturb=0.3)
    nw.pose[:-1], nw.sb[:-1] = w.pose[1:], w.sb[1:]
    nw.pose[-1], nw.sb[-1] = w.pose[-1], w.sb[-1]
    nw.ex[:], nw.td[:] = w.ex, w.td
    nw.prior = prior
    # evaluation of the prior factor + full solve with the prior, device vs oracle
    g, o = win.evaluate(nw), ob.evaluate(nw)
    assert np.abs(g.res_prior - o.res_prior).max() < 1e-9 * max(1.0, np.abs(o.res_prior).max())
    a, b = nw.clone(), nw.clone()
    sg, so = win.optimize(a), ob.optimize(b)
    assert sg.iterations == so.iterations
    for i in range(so.iterations + 1):
        assert abs(sg.cost[i] - so.cost[i]) <= 1e-6 * abs(so.cost[i]) + 1e-9
    assert np.abs(a.pose - b.pose).max() < 1e-6
    for flag in (B.MARGIN_OLD, B.MARGIN_SECOND_NEW):
        pg, po = win.marginalize(b, flag), ob.marginalize(b, flag)
        _info_close(pg, po, 1e-6)
    assert win.marginalize(w, B.MARGIN_SECOND_NEW).n == 0  # no prior => nothing to do (estimator.cpp:942-943)


def test_device_resident_solve_then_marginalize(lfvio_mod, win, problem):
    win.upload(problem)
    win.optimize_dev(True, B.MARGIN_OLD)
    o = problem.clone()
    s = win.download(o)
    assert s.iterations > 0
    po = ob.marginalize(o, B.MARGIN_OLD)
    pg = win.marginalize(o, B.MARGIN_OLD)
    _info_close(pg, po)


# ---- larger windows (BASELINE.json configs 3/4 shapes): global-memory Cholesky path, more landmarks ------------
# (21, 3600, 2000) is the shape bench.py runs at BASELINE config 4 (R ~ 31 k projection factors, D = 322, L > 1024: pre_kernel / gnlm_kernel)
@pytest.mark.parametrize("n_frames,n_lm,feats", [(11, 900, 500), (21, 1500, 800), (21, 3600, 2000)])
def test_large_window_matches_oracle(lfvio_mod, gpu, n_frames, n_lm, feats):
    opts = lfvio_mod.solver_default_opts()
    w = synthwin.make_window(seed=7, n_frames=n_frames, n_landmarks=n_lm, feats_per_frame=feats, opts=opts, perturb=0.3)
    win = lfvio_mod.Window(n_frames - 1, 4096, 65536, device=gpu)
    g, o = win.evaluate(w), ob.evaluate(w)
    assert rel_err(g.res_proj, o.res_proj) < 1e-9 and rel_err(g.jac_imu, o.jac_imu) < 1e-7
    a, b = w.clone(), w.clone()
    sg, so = win.optimize(a), ob.optimize(b)
    assert sg.iterations == so.iterations and sg.num_successful_steps == so.num_successful_steps
    for i in range(so.iterations + 1):
        assert abs(sg.cost[i] - so.cost[i]) <= 1e-6 * abs(so.cost[i]) + 1e-9, (i, sg.cost[i], so.cost[i])
    assert np.abs(a.pose - b.pose).max() < 1e-6 and np.abs(a.feat - b.feat).max() < 1e-6
    pg, po = win.marginalize(b, B.MARGIN_OLD), ob.marginalize(b, B.MARGIN_OLD)
    assert pg.n == po.n
    Ag, bg = pg.information()
    Ao, bo = po.information()
    assert np.abs(Ag - Ao).max() < 1e-6 * np.abs(Ao).max()
    win.close()
