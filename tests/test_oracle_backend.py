"""CPU tests pinning the fp64 back-end oracle (oracle/backend/backend.cpp).

The reference has no tests for this path; its own debugging aids define the checks used here:
  * finite-difference Jacobians with the perturbation convention of ProjectionTdFactor::check
    (factor/projection_td_factor.cpp:215-268: P += d, Q = Q * deltaQ(d), inv_dep += d, eps 1e-6),
  * the marginalisation identities J0^T J0 = A', J0^T r0 = b' (factor/marginalization_factor.cpp:295-296).
"""
import numpy as np
import pytest

from lfvio import backend as B
from lfvio import synthwin
from oracle import backend as ob


@pytest.fixture(scope="module")
def window():
    return synthwin.make_window(seed=1, opts=ob.default_opts(), perturb=0.3)


def _perturbed(w, block, k, eps):
    """Clone of w with local coordinate k of a parameter block moved by eps (reference check() convention)."""
    c = w.clone()
    kind, idx = block
    d = np.zeros(9)
    d[k] = eps
    if kind == "pose":
        c.pose[idx] = ob.pose_plus(w.pose[idx], d[:6])
    elif kind == "sb":
        c.sb[idx] += d
    elif kind == "ex":
        c.ex[:] = ob.pose_plus(w.ex, d[:6])
    elif kind == "feat":
        c.feat[idx] += eps
    elif kind == "td":
        c.td[0] += eps
    return c


def test_pose_plus_keeps_unit_quaternion():
    rng = np.random.default_rng(0)
    for _ in range(20):
        q = rng.normal(size=4)
        q /= np.linalg.norm(q)
        x = np.concatenate([rng.normal(size=3), q])
        d = rng.normal(scale=0.3, size=6)
        y = ob.pose_plus(x, d)
        assert abs(np.linalg.norm(y[3:]) - 1) < 1e-14
        assert np.allclose(y[:3], x[:3] + d[:3])
        # first-order deltaQ: q (+) d = normalize(q * (1, d/2))
        qq = synthwin.quat_mul(np.array([q[3], q[0], q[1], q[2]]), np.array([1, d[3] / 2, d[4] / 2, d[5] / 2]))
        qq /= np.linalg.norm(qq)
        assert np.allclose(y[3:], [qq[1], qq[2], qq[3], qq[0]])


def test_projection_td_jacobian_finite_difference(window):
    w = window
    e0 = ob.evaluate(w, robust=False)
    eps = 1e-6
    rng = np.random.default_rng(1)
    for k in rng.choice(w.n_proj, 12, replace=False):
        f = w.proj[k]
        blocks = [("pose", f.frame_i)] * 6 + [("pose", f.frame_j)] * 6 + [("ex", 0)] * 6 + [("feat", f.landmark), ("td", 0)]
        loc = list(range(6)) * 3 + [0, 0]
        for col, (blk, kk) in enumerate(zip(blocks, loc)):
            e1 = ob.evaluate(_perturbed(w, blk, kk, eps), robust=False)
            num = (e1.res_proj[2 * k:2 * k + 2] - e0.res_proj[2 * k:2 * k + 2]) / eps
            ana = e0.jac_proj[k][:, col].copy()
            if col == 19:
                # reference quirk kept as is (projection_td_factor.cpp:140-146): the vel_j term is
                # sqrt_info * velocity_j.head(2), NOT the tangent-plane projection of the true derivative
                continue
            assert np.allclose(num, ana, rtol=2e-4, atol=2e-3 * max(1.0, np.abs(ana).max())), (k, col, num, ana)


def test_imu_jacobian_finite_difference(window):
    w = window
    e0 = ob.evaluate(w, robust=False)
    eps = 1e-7
    for k in (0, 4, 9):
        i = w.imu[k].frame_i
        blocks = [("pose", i)] * 6 + [("sb", i)] * 9 + [("pose", i + 1)] * 6 + [("sb", i + 1)] * 9
        loc = list(range(6)) + list(range(9)) + list(range(6)) + list(range(9))
        for col, (blk, kk) in enumerate(zip(blocks, loc)):
            e1 = ob.evaluate(_perturbed(w, blk, kk, eps), robust=False)
            num = (e1.res_imu[15 * k:15 * k + 15] - e0.res_imu[15 * k:15 * k + 15]) / eps
            ana = e0.jac_imu[k][:, col]
            scale = max(np.abs(ana).max(), 1.0)
            # the reference's rotation blocks are first-order approximations (Qleft/Qright of the current error
            # quaternion) => agreement to ~1e-3 relative where the rotation residual is not tiny
            assert np.abs(num - ana).max() < 5e-3 * scale, (k, col, np.abs(num - ana).max(), scale)


def test_imu_sqrt_information_factorises_inverse_covariance(window):
    f = window.imu[3]
    S = ob.imu_sqrt_info(f)
    cov = np.array(f.covariance[:]).reshape(15, 15)
    assert np.allclose(S.T @ S @ cov, np.eye(15), atol=1e-6)
    assert np.allclose(np.triu(S), S)  # matrixL().transpose() is upper triangular


def test_cauchy_correction_scales_residual_and_jacobian(window):
    a, b = ob.evaluate(window, robust=False), ob.evaluate(window, robust=True)
    s = (a.res_proj.reshape(-1, 2) ** 2).sum(1)
    sr = 1.0 / np.sqrt(1.0 + s)
    assert np.allclose(b.res_proj.reshape(-1, 2), a.res_proj.reshape(-1, 2) * sr[:, None])
    assert np.allclose(b.jac_proj, a.jac_proj * sr[:, None, None])
    want = 0.5 * np.log1p(s).sum() + 0.5 * (a.res_imu ** 2).sum()
    assert np.isclose(b.cost.value, want)


def test_solver_decreases_cost_and_recovers_truth(window):
    w = window.clone()
    s = ob.optimize(w)
    costs = [s.cost[i] for i in range(s.iterations + 1)]
    assert all(c1 <= c0 * (1 + 1e-12) for c0, c1 in zip(costs, costs[1:]))
    assert costs[-1] < 1e-3 * costs[0]
    # gauge: position/yaw are only observable relative to frame 0, which the generator keeps at truth
    assert np.abs(w.pose[:, :3] - w.truth["P"]).max() < 0.05
    assert abs(w.td[0] - w.truth["td"]) < 5e-3


def test_solver_fixed_iterations_runs_all(window):
    w = window.clone()
    w.opts = ob.default_opts()
    w.opts.fixed_iterations = 1
    s = ob.optimize(w)
    assert s.iterations == 8


def _numpy_marginalize(w, flag):
    """Independent numpy statement of MarginalizationInfo::marginalize for MARGIN_OLD without a prior."""
    e = ob.evaluate(w, robust=True)
    F = w.F
    lm = sorted({w.proj[k].landmark for k in range(w.n_proj) if w.proj[k].frame_i == 0})
    ids_m = [0, 1] + [2 * F + 2 + l for l in lm]
    frames_j = sorted({w.proj[k].frame_j for k in range(w.n_proj) if w.proj[k].frame_i == 0} | {1})
    ids_r = sorted({2 * f for f in frames_j} | {3, 2 * F, 2 * F + 1})
    size = lambda i: 1 if i >= 2 * F + 1 else (9 if (i < 2 * F and i & 1) else 6)
    off, pos = {}, 0
    for i in ids_m + ids_r:
        off[i] = pos
        pos += size(i)
    m = sum(size(i) for i in ids_m)
    A, b = np.zeros((pos, pos)), np.zeros(pos)

    def add(J, r, cols):
        nonlocal A, b
        A[np.ix_(cols, cols)] += J.T @ J
        b[cols] += J.T @ r

    k0 = [k for k in range(w.n_imu) if w.imu[k].frame_i == 0][0]
    cols = sum([list(range(off[i], off[i] + size(i))) for i in (0, 1, 2, 3)], [])
    add(e.jac_imu[k0], e.res_imu[15 * k0:15 * k0 + 15], cols)
    for k in range(w.n_proj):
        f = w.proj[k]
        if f.frame_i != 0:
            continue
        cols = (list(range(off[0], off[0] + 6)) + list(range(off[2 * f.frame_j], off[2 * f.frame_j] + 6)) +
                list(range(off[2 * F], off[2 * F] + 6)) + [off[2 * F + 2 + f.landmark], off[2 * F + 1]])
        add(e.jac_proj[k], e.res_proj[2 * k:2 * k + 2], cols)
    Amm = 0.5 * (A[:m, :m] + A[:m, :m].T)
    wv, V = np.linalg.eigh(Amm)
    inv = V @ np.diag(np.where(wv > 1e-8, 1.0 / np.where(wv > 1e-8, wv, 1), 0)) @ V.T
    Ar = A[m:, m:] - A[m:, :m] @ inv @ A[:m, m:]
    br = b[m:] - A[m:, :m] @ inv @ b[:m]
    return ids_r, Ar, br


def test_marginalize_old_identities(window):
    w = window.clone()
    ob.optimize(w)
    pr = ob.marginalize(w, B.MARGIN_OLD)
    ids_r, Ar, br = _numpy_marginalize(w, B.MARGIN_OLD)
    # kept blocks, shifted by addr_shift: pose/sb i -> i-1 (ids - 2); ex, td keep their ids
    F = w.F
    want_ids = [i - 2 if i < 2 * F else i for i in ids_r]
    assert list(pr.block_id[: pr.n_blocks]) == want_ids
    assert pr.n == Ar.shape[0] == 6 * (len(ids_r) - 3) + 9 + 6 + 1
    JtJ, Jtr = pr.information()
    scale = np.abs(Ar).max()
    assert np.abs(JtJ - Ar).max() < 1e-7 * scale
    # J0^T r0 = b' holds on the range of A' (eigenvalues below 1e-8 are dropped by the reference)
    assert np.abs(Jtr - br).max() < 1e-6 * max(np.abs(br).max(), 1.0)
    # the x0 copies are the CURRENT values of the kept blocks before the shift
    assert np.allclose(pr.x0[:7], w.pose[1]) and np.allclose(pr.x0[7:16], w.sb[1])


def test_marginalize_threads_agree(window):
    w = window.clone()
    a, b = ob.marginalize(w, B.MARGIN_OLD, 1), ob.marginalize(w, B.MARGIN_OLD, 4)
    Aa, ba = a.information()
    Ab, bb = b.information()
    assert np.allclose(Aa, Ab, rtol=1e-9, atol=1e-9 * np.abs(Aa).max()) and np.allclose(ba, bb, rtol=1e-8, atol=1e-8 * np.abs(ba).max())


def test_prior_factor_enters_next_solve_and_second_new_marginalisation(window):
    w = window.clone()
    ob.optimize(w)
    prior = ob.marginalize(w, B.MARGIN_OLD)
    # re-use the prior on a window whose states are the shifted, optimised ones (+ small perturbation)
    nw = window.clone()
    nw.pose[:-1], nw.sb[:-1] = w.pose[1:], w.sb[1:]
    nw.pose[-1], nw.sb[-1] = w.pose[-1], w.sb[-1]
    nw.ex[:], nw.td[:] = w.ex, w.td
    nw.prior = prior
    e = ob.evaluate(nw)
    # at the linearisation point (x == x0 for every kept block) the prior residual is exactly r0
    assert np.allclose(e.res_prior, prior.residual(), atol=1e-12)
    nw.pose[3, 0] += 0.01
    e2 = ob.evaluate(nw)
    col = prior.block_idx[list(prior.block_id[: prior.n_blocks]).index(2 * 3)]
    assert np.allclose(e2.res_prior - e.res_prior, prior.jacobian()[:, col] * 0.01, atol=1e-9)
    # MARGIN_SECOND_NEW drops pose F-2 from the prior only
    p2 = ob.marginalize(nw, B.MARGIN_SECOND_NEW)
    assert p2.n == prior.n - 6
    assert 2 * (nw.F - 2) not in list(p2.block_id[: p2.n_blocks])  # the second-newest pose is gone
    A1, _ = prior.information()
    A2, _ = p2.information()
    assert np.linalg.eigvalsh(A2).min() > -1e-6 * np.abs(A1).max()
