"""CPU oracle (TEST INFRASTRUCTURE ONLY) for the estimator initialisation, a numpy restatement of

  solve_gyroscope_bias / linear_alignment / refine_gravity / visual_imu_alignment
                        vins_estimator/src/initial/initial_aligment.cpp:3-38, 132-205, 56-130, 207-215
  compute_E_21 / check_inliers / find_fundamental        vins_estimator/src/initial/solve_5pts.cpp:201-335 (the same code
                        as feature_tracker.cpp:186-327, restated once in oracle/ransac.py and reused here)
  decompose_essential / recover_pose / solve_relative_rt solve_5pts.cpp:337-361, 392-536, 538-575
  triangulate_point     vins_estimator/src/initial/initial_sfm.cpp:6-23 (== solve_5pts.cpp:363-378)
  epnp_compute_pose     vins_estimator/src/pnp_solver.cpp:46-440 (EPnP on bearings, quirks kept: see the function)
  global_sfm            initial_sfm.cpp:127-316 (GlobalSFM::construct) with the Ceres 1.12 Levenberg-Marquardt loop of its
                        "full BA" restated (third party, not vendored: docker/Dockerfile:3 pins ceres-solver 1.12.0)

Stated deviations (also in DESIGN.md §4):
 * solve_5pts.cpp:556-569 passes `rot`, `trans` to recoverPose BY VALUE and then reads two uninitialised locals R, T
   (`Rotation = R.transpose()`), so the reference's own relative pose is undefined behaviour.  Restated as evidently
   intended (and as VINS-Mono does): Rotation = R^T, Translation = -R^T t of the pose recoverPose selects.
 * the random source of the RANSAC is injected (see oracle/ransac.py).
 * Ceres' max_solver_time_in_seconds = 0.2 (initial_sfm.cpp:279) is a wall-clock limit and cannot be restated; the loop
   runs to convergence or 50 iterations (the Ceres default max_num_iterations).
 * Eigen::JacobiSVD / LDLT / colPivHouseholderQr are replaced by LAPACK (numpy): same subspaces / solutions to rounding.

Parity unpinned against a reference binary: the reference has no tests or vectors for these functions and cannot be built
here (Eigen, Ceres, OpenCV and ROS are absent).  Pinned instead by noise-free synthetic geometry with known answers
(tests/test_oracle_initial.py).  Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import this module.
"""
from __future__ import annotations

import numpy as np

from . import ransac as R8
from .prestep import _qmul, _qrot

O_R, O_BG = 3, 12  # integration_base.h / parameters.h StateOrder


def rot_to_quat(Rm):
    """Eigen::Quaternion(Matrix3) (Shepperd), (w, x, y, z)."""
    Rm = np.asarray(Rm, float)
    t = np.trace(Rm)
    if t > 0:
        s = np.sqrt(t + 1.0) * 2
        q = [0.25 * s, (Rm[2, 1] - Rm[1, 2]) / s, (Rm[0, 2] - Rm[2, 0]) / s, (Rm[1, 0] - Rm[0, 1]) / s]
    else:
        i = int(np.argmax(np.diag(Rm)))
        j, k = (i + 1) % 3, (i + 2) % 3
        s = np.sqrt(Rm[i, i] - Rm[j, j] - Rm[k, k] + 1.0) * 2
        q = [0.0] * 4
        q[0] = (Rm[k, j] - Rm[j, k]) / s
        q[1 + i] = 0.25 * s
        q[1 + j] = (Rm[j, i] + Rm[i, j]) / s
        q[1 + k] = (Rm[k, i] + Rm[i, k]) / s
    return np.array(q)


# ------------------------------------------------------------------------------------------------------------------
# visual-inertial alignment
# ------------------------------------------------------------------------------------------------------------------
def solve_gyroscope_bias(Rs, delta_q, J_rg):
    """initial_aligment.cpp:3-29.  Rs: (n,3,3) ImageFrame::R; delta_q (n-1,4 wxyz) and J_rg (n-1,3,3) =
    jacobian.block<3,3>(O_R,O_BG) of the pre-integration ending at frame k+1.  Returns delta_bg."""
    A, b = np.zeros((3, 3)), np.zeros(3)
    for k in range(len(Rs) - 1):
        q_ij = rot_to_quat(Rs[k].T @ Rs[k + 1])
        dq = np.asarray(delta_q[k], float)
        dq_inv = np.array([dq[0], -dq[1], -dq[2], -dq[3]]) / (dq @ dq)  # Quaternion::inverse
        tmp_b = 2 * _qmul(dq_inv, q_ij)[1:]
        tmp_A = np.asarray(J_rg[k], float)
        A += tmp_A.T @ tmp_A
        b += tmp_A.T @ tmp_b
    return np.linalg.solve(A, b)  # A.ldlt().solve(b)


def tangent_basis(g0):
    """initial_aligment.cpp:40-54."""
    a = g0 / np.linalg.norm(g0)
    tmp = np.array([0.0, 0, 1])
    if np.array_equal(a, tmp):
        tmp = np.array([1.0, 0, 0])
    b = tmp - a * (a @ tmp)
    b = b / np.linalg.norm(b)
    return np.stack([b, np.cross(a, b)], axis=1)


def _alignment_system(Rs, Ts, dp, dv, dts, tic, g_cols, g_fixed, A=None, b=None):
    """The normal equations shared by LinearAlignment (:132-180, g_cols = I3, g_fixed = 0) and one RefineGravity pass
    (:73-123, g_cols = lxly, g_fixed = g0).  A / b: accumulate on top of these (RefineGravity zeroes them ONCE, before its
    four passes, :62-65) and scale the sum by 1000 (:119-120)."""
    n, k = len(Rs), g_cols.shape[1]
    ns = 3 * n + k + 1
    if A is None:
        A, b = np.zeros((ns, ns)), np.zeros(ns)
    for i in range(n - 1):
        dt = dts[i]
        Ri_T = Rs[i].T
        tA = np.zeros((6, 6 + k + 1))
        tb = np.zeros(6)
        tA[0:3, 0:3] = -dt * np.eye(3)
        tA[0:3, 6:6 + k] = Ri_T * dt * dt / 2 @ g_cols
        tA[0:3, 6 + k] = Ri_T @ (Ts[i + 1] - Ts[i]) / 100.0
        tb[0:3] = dp[i] + Ri_T @ Rs[i + 1] @ tic - tic - Ri_T * dt * dt / 2 @ g_fixed
        tA[3:6, 0:3] = -np.eye(3)
        tA[3:6, 3:6] = Ri_T @ Rs[i + 1]
        tA[3:6, 6:6 + k] = Ri_T * dt @ g_cols
        tb[3:6] = dv[i] - Ri_T * dt @ g_fixed
        rA, rb = tA.T @ tA, tA.T @ tb
        A[3 * i:3 * i + 6, 3 * i:3 * i + 6] += rA[:6, :6]
        b[3 * i:3 * i + 6] += rb[:6]
        A[-(k + 1):, -(k + 1):] += rA[6:, 6:]
        b[-(k + 1):] += rb[6:]
        A[3 * i:3 * i + 6, -(k + 1):] += rA[:6, 6:]
        A[-(k + 1):, 3 * i:3 * i + 6] += rA[6:, :6]
    return A * 1000.0, b * 1000.0


def refine_gravity(Rs, Ts, dp, dv, dts, tic, g, g_norm):
    """initial_aligment.cpp:56-130.  Returns (g, x) with x of size 3n + 2 + 1."""
    g0 = g / np.linalg.norm(g) * g_norm
    x = None
    A = b = None
    for _ in range(4):
        # The reference does NOT reset A and b between the passes: pass k solves 1000 (A_{k-1} + new_k), so the first
        # linearisation dominates every later solve by three orders of magnitude per pass (kept as is; pinned against the
        # reference's own code by tests/test_oracle_ref.py).
        lxly = tangent_basis(g0)
        A, b = _alignment_system(Rs, Ts, dp, dv, dts, tic, lxly, g0, A, b)
        x = np.linalg.solve(A, b)
        dg = x[-3:-1]
        g0 = g0 + lxly @ dg
        g0 = g0 / np.linalg.norm(g0) * g_norm
    return g0, x


def linear_alignment(Rs, Ts, dp, dv, dts, tic, g_norm):
    """initial_aligment.cpp:132-205.  Returns (ok, g, x): x = [v_0 .. v_{n-1} (body frame), dg (2), s] with s already / 100."""
    Rs, Ts = np.asarray(Rs, float), np.asarray(Ts, float)
    A, b = _alignment_system(Rs, Ts, dp, dv, dts, tic, np.eye(3), np.zeros(3))
    x = np.linalg.solve(A, b)
    s = x[-1] / 100.0
    g = x[-4:-1].copy()
    if abs(np.linalg.norm(g) - g_norm) > 1.0 or s < 0:
        return False, g, x
    g, x = refine_gravity(Rs, Ts, dp, dv, dts, tic, g, g_norm)
    x = x.copy()
    x[-1] = x[-1] / 100.0
    return bool(x[-1] >= 0.0), g, x


# ------------------------------------------------------------------------------------------------------------------
# relative pose
# ------------------------------------------------------------------------------------------------------------------
def triangulate_point(P0, P1, p0, p1):
    """initial_sfm.cpp:6-23 / solve_5pts.cpp:363-378."""
    D = np.stack([p0[0] * P0[2] - p0[2] * P0[0], p0[1] * P0[2] - p0[2] * P0[1],
                  p1[0] * P1[2] - p1[2] * P1[0], p1[1] * P1[2] - p1[2] * P1[1]])
    X = np.linalg.svd(D)[2][-1]
    return X[:3] / X[3]


def decompose_essential(E):
    """solve_5pts.cpp:337-361."""
    U, _, Vt = np.linalg.svd(E)
    V = Vt.T
    if np.linalg.det(U) < 0:
        U = -U
    if np.linalg.det(V) < 0:
        V = -V
    W = np.array([[0.0, 1, 0], [-1, 0, 0], [0, 0, 1]])
    return U @ W @ V.T, U @ W.T @ V.T, U[:, 2].copy()


def cheirality_count(Rm, t, b1, b2):
    """One of the four blocks of recoverPose (solve_5pts.cpp:430-448): both bearings must point towards the triangulated point."""
    P0 = np.hstack([np.eye(3), np.zeros((3, 1))])
    P1 = np.hstack([Rm, t.reshape(3, 1)])
    good = 0
    for i in range(len(b1)):
        Q = triangulate_point(P0, P1, b1[i], b2[i])
        l_dot = b1[i] @ Q / np.linalg.norm(Q)
        tq = Rm @ Q + t
        r_dot = b2[i] @ tq / np.linalg.norm(tq)
        good += bool(l_dot > 0 and r_dot > 0)
    return good


def recover_pose(E, b1, b2):
    """solve_5pts.cpp:392-536.  Returns (R, t, good, counts[4])."""
    R1, R2, t = decompose_essential(E)
    cands = [(R1, t), (R2, t), (R1, -t), (R2, -t)]
    counts = [cheirality_count(Rc, tc, b1, b2) for Rc, tc in cands]
    g1, g2, g3, g4 = counts
    if g1 >= g2 and g1 >= g3 and g1 >= g4:
        k = 0
    elif g2 >= g1 and g2 >= g3 and g2 >= g4:
        k = 1
    elif g3 >= g1 and g3 >= g2 and g3 >= g4:
        k = 2
    else:
        k = 3
    return cands[k][0], cands[k][1], counts[k], counts


def solve_relative_rt(corres, seed, frame):
    """solve_5pts.cpp:538-575 with the by-value bug repaired (module docstring).  corres: (n, 6) = [first xyz | second xyz].
    Returns (ok, Rotation, Translation, inlier_cnt, E)."""
    corres = np.asarray(corres, float).reshape(-1, 6)
    if len(corres) < 15:
        return False, np.eye(3), np.zeros(3), 0, np.zeros((3, 3))
    b1, b2 = corres[:, :3], corres[:, 3:]
    E = R8.find_fundamental(b1, b2, seed, frame)[0]
    Rm, t, cnt, _ = recover_pose(E, b1, b2)
    return cnt > 12, Rm.T, -Rm.T @ t, cnt, E


# ------------------------------------------------------------------------------------------------------------------
# EPnP on bearings
# ------------------------------------------------------------------------------------------------------------------
def _lstsq(A, b):  # colPivHouseholderQr().solve: least squares for the full-rank case
    return np.linalg.lstsq(A, b, rcond=None)[0]


def _epnp_L_6x10(Ut):
    """pnp_solver.cpp:98-136."""
    dv = np.zeros((4, 6, 3))
    for i in range(4):
        a, b = 0, 1
        for j in range(6):
            dv[i, j] = Ut[11 - i, 3 * a:3 * a + 3] - Ut[11 - i, 3 * b:3 * b + 3]
            b += 1
            if b > 3:
                a += 1
                b = a + 1
    L = np.zeros((6, 10))
    for i in range(6):
        d = dv[:, i]
        L[i] = [d[0] @ d[0], 2 * d[0] @ d[1], d[1] @ d[1], 2 * d[0] @ d[2], 2 * d[1] @ d[2], d[2] @ d[2],
                2 * d[0] @ d[3], 2 * d[1] @ d[3], 2 * d[2] @ d[3], d[3] @ d[3]]
    return L


def _epnp_gauss_newton(L, rho, betas):
    """pnp_solver.cpp:384-434 (15 iterations)."""
    b = betas.copy()
    for _ in range(15):
        A = np.zeros((6, 4))
        r = np.zeros(6)
        for i in range(6):
            l = L[i]
            A[i] = [2 * l[0] * b[0] + l[1] * b[1] + l[3] * b[2] + l[6] * b[3],
                    l[1] * b[0] + 2 * l[2] * b[1] + l[4] * b[2] + l[7] * b[3],
                    l[3] * b[0] + l[4] * b[1] + 2 * l[5] * b[2] + l[8] * b[3],
                    l[6] * b[0] + l[7] * b[1] + l[8] * b[2] + 2 * l[9] * b[3]]
            r[i] = rho[i] - (l[0] * b[0] * b[0] + l[1] * b[0] * b[1] + l[2] * b[1] * b[1] + l[3] * b[0] * b[2] +
                             l[4] * b[1] * b[2] + l[5] * b[2] * b[2] + l[6] * b[0] * b[3] + l[7] * b[1] * b[3] +
                             l[8] * b[2] * b[3] + l[9] * b[3] * b[3])
        b = b + _lstsq(A, r)
    return b


def epnp_compute_pose(pws, us):
    """PnpSolver::compute_pose with set_internal_parameters(0, 0, 1, 1) (pnp_solver.cpp:275-338; call sites
    initial_sfm.cpp:55-69, estimator.cpp:343-357).  pws: world points, us: bearings (unit, z may be negative).
    Returns (R, T, error) with p_cam = R p_w + T.  Quirks kept: perspective division by the bearing's z (:286-290), the
    sign fix looks at correspondence 0 only (:244-252), no det(R) check in estimate_R_and_t (:253-280), the
    "reprojection error" compares the unit bearing with the un-normalised camera point (:281-292), and the sign test of
    find_betas_2 writes to B5 instead of Betas (:217)."""
    pws, us = np.asarray(pws, float), np.asarray(us, float)
    n = len(pws)
    signs = np.where(us[:, 2] > 0, 1.0, -1.0)
    # choose_control_points :46-78
    cws = np.zeros((4, 3))
    cws[0] = pws.mean(axis=0)
    PW0 = pws - cws[0]
    U, DC, _ = np.linalg.svd(PW0.T @ PW0)
    for i in range(1, 4):
        u = U[:, i - 1]
        # The sign of a singular vector is the SVD implementation's choice (Eigen's is not reproducible here), and with
        # noisy data the EPnP result depends on it at noise level.  Convention shared with the CUDA path: the component of
        # largest magnitude is positive.
        if u[int(np.argmax(np.abs(u)))] < 0:
            u = -u
        cws[i] = cws[0] + np.sqrt(DC[i - 1] / n) * u
    # compute_barycentric_coordinates :79-101
    CC = (cws[1:] - cws[0]).T
    al = np.zeros((n, 4))
    al[:, 1:] = (np.linalg.inv(CC) @ (pws - cws[0]).T).T
    al[:, 0] = 1.0 - al[:, 1] - al[:, 2] - al[:, 3]
    # M :280-293
    M = np.zeros((2 * n, 12))
    for j in range(4):
        M[0::2, 3 * j] = al[:, j]
        M[0::2, 3 * j + 2] = al[:, j] * (0.0 - us[:, 0]) / us[:, 2]
        M[1::2, 3 * j + 1] = al[:, j]
        M[1::2, 3 * j + 2] = al[:, j] * (0.0 - us[:, 1]) / us[:, 2]
    Uu, _, _ = np.linalg.svd(M.T @ M)
    Ut = Uu.T
    L = _epnp_L_6x10(Ut)
    rho = np.array([np.sum((cws[a] - cws[b]) ** 2) for a, b in ((0, 1), (0, 2), (0, 3), (1, 2), (1, 3), (2, 3))])

    def R_and_t(betas):  # compute_ccs / compute_pcs / solve_for_sign / estimate_R_and_t / reprojection_error
        ccs = np.zeros((4, 3))
        for i in range(4):
            ccs += betas[i] * Ut[11 - i].reshape(4, 3)
        pcs = al @ ccs
        if (pcs[0, 2] < 0 and signs[0] > 0) or (pcs[0, 2] > 0 and signs[0] < 0):
            pcs = -pcs
        pc0, pw0 = pcs.mean(axis=0), pws.mean(axis=0)
        W = (pcs - pc0).T @ (pws - pw0)
        Uw, _, Vtw = np.linalg.svd(W)
        Rm = Uw @ Vtw
        T = pc0 - Rm @ pw0
        err = np.sum((us - (pws @ Rm.T + T)) ** 2) / n
        return Rm, T, err

    b4 = _lstsq(L[:, [0, 1, 3, 6]], rho)                                     # find_betas_0 :147-175
    if b4[0] < 0:
        b0 = np.array([np.sqrt(-b4[0]), 0, 0, 0])
        b0[1:] = -b4[1:] / b0[0]
    else:
        b0 = np.array([np.sqrt(b4[0]), 0, 0, 0])
        b0[1:] = b4[1:] / b0[0]
    b3 = _lstsq(L[:, [0, 1, 2]], rho)                                        # find_betas_1 :177-203
    if b3[0] < 0:
        b1 = np.array([np.sqrt(-b3[0]), np.sqrt(-b3[2]) if b3[2] < 0 else 0.0, 0, 0])
    else:
        b1 = np.array([np.sqrt(b3[0]), np.sqrt(b3[2]) if b3[2] > 0 else 0.0, 0, 0])
    if b3[1] < 0:
        b1[0] = -b1[0]
    b5 = _lstsq(L[:, [0, 1, 2, 3, 4]], rho)                                  # find_betas_2 :204-231
    if b5[0] < 0:
        b2 = np.array([np.sqrt(-b5[0]), np.sqrt(-b5[2]) if b5[2] < 0 else 0.0, 0, 0])
    else:
        b2 = np.array([np.sqrt(b5[0]), np.sqrt(b5[2]) if b5[2] > 0 else 0.0, 0, 0])
    b2[2] = b5[3] / b2[0]
    res = [R_and_t(_epnp_gauss_newton(L, rho, b)) for b in (b0, b1, b2)]
    N = 0
    if res[1][2] < res[0][2]:
        N = 1
    if res[2][2] < res[N][2]:
        N = 2
    return res[N]


# ------------------------------------------------------------------------------------------------------------------
# GlobalSFM::construct
# ------------------------------------------------------------------------------------------------------------------
def _ba_residual_jac(q, t, p, obs, want_jac=True):
    """ReprojectionError3D (initial_sfm.h:26-66) with ceres::QuaternionParameterization (x_plus = delta (x) x): residual
    n(R(q) p + t) - obs and its Jacobians in local coordinates: rotation (3x3, left-multiplied increment of angle 2|d|),
    translation (3x3), point (3x3)."""
    qn = q / np.linalg.norm(q)                       # QuaternionRotatePoint normalises
    Rm = _qrot(qn)
    rp = Rm @ p
    pc = rp + t
    nrm = np.linalg.norm(pc)
    r = pc / nrm - obs
    if not want_jac:
        return r, None, None, None
    dn = (np.eye(3) - np.outer(pc, pc) / (nrm * nrm)) / nrm
    skew = np.array([[0, -rp[2], rp[1]], [rp[2], 0, -rp[0]], [-rp[1], rp[0], 0]])
    return r, dn @ (-2.0 * skew), dn, dn @ Rm


def _quat_plus(q, d):
    """ceres::QuaternionParameterization::Plus (local_parameterization.cc, Ceres 1.12)."""
    nd = np.linalg.norm(d)
    if nd == 0.0:
        return q.copy()
    s = np.sin(nd) / nd
    return _qmul(np.array([np.cos(nd), s * d[0], s * d[1], s * d[2]]), q)


def bundle_adjust(c_rot, c_trans, points, state, obs_first, obs_frame, obs_pts, l, max_iter=50):
    """The "full BA" of initial_sfm.cpp:230-297: ceres::Solve with default options (trust region, Levenberg-Marquardt,
    DENSE_SCHUR, Jacobi scaling, no loss).  c_rot (F,4 wxyz) / c_trans (F,3) / points (L,3) are updated in place.
    Rotation l and translations l, F-1 are constant (:243-251).  Returns dict(converged, final_cost, iterations, costs)."""
    F = len(c_rot)
    rot_free = [i != l for i in range(F)]
    tr_free = [i != l and i != F - 1 for i in range(F)]
    pts = [j for j in range(len(points)) if state[j]]
    n_obs = [(j, o) for j in pts for o in range(obs_first[j], obs_first[j + 1])]
    # columns: rotation 3 per frame, translation 3 per frame, 3 per point
    col_r = lambda i: 3 * i
    col_t = lambda i: 3 * F + 3 * i
    pcol = {j: 6 * F + 3 * k for k, j in enumerate(pts)}
    ncol = 6 * F + 3 * len(pts)
    free = np.zeros(ncol, bool)
    for i in range(F):
        free[col_r(i):col_r(i) + 3] = rot_free[i]
        free[col_t(i):col_t(i) + 3] = tr_free[i]
    free[6 * F:] = True
    # a camera that sees no point has no residual block; Ceres would not even know its parameter blocks' gradients
    used = np.zeros(ncol, bool)

    def evaluate(cr, ct, pp, want_jac):
        r = np.zeros(3 * len(n_obs))
        J = np.zeros((3 * len(n_obs), ncol)) if want_jac else None
        for k, (j, o) in enumerate(n_obs):
            f = obs_frame[o]
            rr, Jr, Jt, Jp = _ba_residual_jac(cr[f], ct[f], pp[j], obs_pts[o], want_jac)
            r[3 * k:3 * k + 3] = rr
            if want_jac:
                J[3 * k:3 * k + 3, col_r(f):col_r(f) + 3] = Jr
                J[3 * k:3 * k + 3, col_t(f):col_t(f) + 3] = Jt
                J[3 * k:3 * k + 3, pcol[j]:pcol[j] + 3] = Jp
                used[col_r(f):col_r(f) + 3] = used[col_t(f):col_t(f) + 3] = True
                used[pcol[j]:pcol[j] + 3] = True
        return r, J

    r, J = evaluate(c_rot, c_trans, points, True)
    act = free & used
    cost = 0.5 * r @ r
    costs = [cost]
    Ja = J[:, act]
    scale = 1.0 / (1.0 + np.sqrt(np.sum(Ja * Ja, axis=0)))          # trust_region_minimizer.cc: EstimateScale, once
    radius, decrease = 1e4, 2.0
    converged, it = False, 0
    x_norm = lambda: np.sqrt(sum(np.sum(c_rot[i] ** 2) for i in range(F) if rot_free[i]) +
                             sum(np.sum(c_trans[i] ** 2) for i in range(F) if tr_free[i]) + sum(np.sum(points[j] ** 2) for j in pts))
    g = Ja.T @ r
    if np.max(np.abs(g)) <= 1e-10:
        return dict(converged=True, final_cost=cost, iterations=0, costs=costs)
    while it < max_iter:
        it += 1
        Js = Ja * scale
        gs = Js.T @ r
        H = Js.T @ Js
        D2 = np.clip(np.diag(H), 1e-6, 1e32) / radius             # levenberg_marquardt_strategy.cc
        try:
            step_s = np.linalg.solve(H + np.diag(D2), -gs)
        except np.linalg.LinAlgError:
            radius /= decrease
            decrease *= 2
            continue
        mr = Js @ step_s
        model_change = -(mr @ (r + mr / 2.0))
        if not model_change > 0:
            radius /= decrease
            decrease *= 2
            continue
        full = np.zeros(ncol)
        full[act] = step_s * scale
        cr = np.array([_quat_plus(c_rot[i], full[col_r(i):col_r(i) + 3]) if rot_free[i] else c_rot[i] for i in range(F)])
        ct = np.array([c_trans[i] + full[col_t(i):col_t(i) + 3] for i in range(F)])
        pp = points.copy()
        for j in pts:
            pp[j] = points[j] + full[pcol[j]:pcol[j] + 3]
        rn, _ = evaluate(cr, ct, pp, False)
        new_cost = 0.5 * rn @ rn
        step_norm = np.sqrt(sum(np.sum((cr[i] - c_rot[i]) ** 2) for i in range(F)) + np.sum((ct - c_trans) ** 2) + np.sum((pp - points) ** 2))
        if step_norm <= 1e-8 * (x_norm() + 1e-8):                  # parameter_tolerance (returns without taking the step)
            converged = True
            break
        change = cost - new_cost
        if abs(change) <= 1e-6 * cost:                             # function_tolerance (likewise)
            converged = True
            break
        rho = change / model_change
        if rho > 1e-3:                                             # min_relative_decrease
            c_rot[:], c_trans[:], points[:] = cr, ct, pp
            cost = new_cost
            costs.append(cost)
            r, J = evaluate(c_rot, c_trans, points, True)
            Ja = J[:, act]
            if np.max(np.abs(Ja.T @ r)) <= 1e-10:                  # gradient_tolerance
                converged = True
                break
            radius = min(1e16, radius / max(1.0 / 3.0, 1.0 - (2.0 * rho - 1.0) ** 3))
            decrease = 2.0
        else:
            radius /= decrease
            decrease *= 2
            if radius < 1e-32:
                break
    return dict(converged=converged, final_cost=cost, iterations=it, costs=costs)


def global_sfm(frame_num, l, relative_R, relative_T, obs_first, obs_frame, obs_pts, ba=True):
    """GlobalSFM::construct (initial_sfm.cpp:127-316).  Feature j has observations obs_first[j]..obs_first[j+1] with
    frames obs_frame[o] (ascending) and bearings obs_pts[o].  Returns dict(ok, q (F,4 wxyz: w_R_cam), T (F,3), points (L,3),
    state (L,), ba summary, c_rot/c_trans before BA)."""
    obs_first, obs_frame = np.asarray(obs_first, int), np.asarray(obs_frame, int)
    obs_pts = np.asarray(obs_pts, float).reshape(-1, 3)
    L = len(obs_first) - 1
    state = np.zeros(L, bool)
    points = np.zeros((L, 3))
    last = frame_num - 1
    c_R = [np.eye(3) for _ in range(frame_num)]
    c_T = [np.zeros(3) for _ in range(frame_num)]
    q_last = rot_to_quat(relative_R)                                # q[l] = identity, q[last] = q[l] * Quaterniond(relative_R)
    c_R[l] = np.eye(3)
    c_T[l] = np.zeros(3)
    c_R[last] = _qrot(np.array([q_last[0], -q_last[1], -q_last[2], -q_last[3]]) / (q_last @ q_last))
    c_T[last] = -(c_R[last] @ np.asarray(relative_T, float))
    pose = lambda i: np.hstack([c_R[i], c_T[i].reshape(3, 1)])

    def find(j, f):
        for o in range(obs_first[j], obs_first[j + 1]):
            if obs_frame[o] == f:
                return o
        return -1

    def tri_two(f0, f1):                                            # triangulateTwoFrames :74-114
        for j in range(L):
            if state[j]:
                continue
            o0, o1 = find(j, f0), find(j, f1)
            if o0 >= 0 and o1 >= 0:
                points[j] = triangulate_point(pose(f0), pose(f1), obs_pts[o0], obs_pts[o1])
                state[j] = True

    def pnp(i):                                                     # solveFrameByPnP :25-72
        p2, p3 = [], []
        for j in range(L):
            if not state[j]:
                continue
            o = find(j, i)
            if o >= 0:
                p2.append(obs_pts[o])
                p3.append(points[j].copy())
        if len(p2) < 10:
            return False
        Rm, T, _ = epnp_compute_pose(np.array(p3), np.array(p2))
        c_R[i], c_T[i] = Rm, T
        return True

    for i in range(l, last):                                        # :167-185
        if i > l and not pnp(i):
            return dict(ok=False, stage="pnp", frame=i)
        tri_two(i, last)
    for i in range(l + 1, last):                                    # :186-188
        tri_two(l, i)
    for i in range(l - 1, -1, -1):                                  # :191-206
        if not pnp(i):
            return dict(ok=False, stage="pnp", frame=i)
        tri_two(i, l)
    for j in range(L):                                              # :207-228
        if state[j] or obs_first[j + 1] - obs_first[j] < 2:
            continue
        o0, o1 = obs_first[j], obs_first[j + 1] - 1
        points[j] = triangulate_point(pose(obs_frame[o0]), pose(obs_frame[o1]), obs_pts[o0], obs_pts[o1])
        state[j] = True
    c_rot = np.array([rot_to_quat(c_R[i]) for i in range(frame_num)])
    c_rot[l] = [1.0, 0, 0, 0]                                       # c_Quat[l] = q[l].inverse()
    c_rot[last] = np.array([q_last[0], -q_last[1], -q_last[2], -q_last[3]]) / (q_last @ q_last)
    c_trans = np.array(c_T)
    pre = (c_rot.copy(), c_trans.copy(), points.copy())
    summ = dict(converged=True, final_cost=0.0, iterations=0, costs=[])
    if ba:
        summ = bundle_adjust(c_rot, c_trans, points, state, obs_first, obs_frame, obs_pts, l)
        if not (summ["converged"] or summ["final_cost"] < 5e-3):    # :283-294
            return dict(ok=False, stage="ba", ba=summ)
    q = np.array([np.array([c[0], -c[1], -c[2], -c[3]]) / (c @ c) for c in c_rot])   # :295-303
    T = np.array([-(_qrot(q[i]) @ c_trans[i]) for i in range(frame_num)])             # :304-309
    return dict(ok=True, q=q, T=T, points=points, state=state, ba=summ, pre_ba=pre)


# ------------------------------------------------------------------------------------------------------------------
# Estimator::initialStructure / visualInitialAlign (all frames key frames)
# ------------------------------------------------------------------------------------------------------------------
def _R2ypr_yaw_deg(Rm):  # Utility::R2ypr(R).x() (utility/utility.h:66-81)
    return np.degrees(np.arctan2(Rm[1, 0], Rm[0, 0]))


def _yaw_rot(deg):  # Utility::ypr2R({yaw, 0, 0})
    a = np.radians(deg)
    return np.array([[np.cos(a), -np.sin(a), 0], [np.sin(a), np.cos(a), 0], [0, 0, 1.0]])


def g2R(g):
    """utility/utility.cpp:3-12 (Quaterniond::FromTwoVectors(g.normalized(), e_z))."""
    v0 = g / np.linalg.norm(g)
    v1 = np.array([0.0, 0, 1])
    c = v1 @ v0
    axis = np.cross(v0, v1)
    s = np.sqrt((1 + c) * 2)
    R0 = _qrot(np.array([s * 0.5, *(axis / s)]))
    return _yaw_rot(-_R2ypr_yaw_deg(R0)) @ R0


def relative_pose(corres_fn, window_size, seed, attempt):
    """Estimator::relativePose (estimator.cpp:457-484); corres_fn(i, j) -> (n, 6)."""
    for i in range(window_size):
        c = corres_fn(i, window_size)
        if len(c) <= 20:
            continue
        par = np.mean(np.linalg.norm(c[:, :2] / c[:, 2:3] - c[:, 3:5] / c[:, 5:6], axis=1))
        if not par * 160 > 30:
            continue
        ok, Rm, T, cnt, _ = solve_relative_rt(c, seed, attempt * 64 + i)
        if ok:
            return True, Rm, T, i
    return False, None, None, -1


def initial_structure(F, obs_first, obs_frame, obs_pts, corres_fn, reintegrate, ric, tic, g_norm, bgs0, seed, attempt=1,
                      init_depth=5.0):
    """Estimator::initialStructure + visualInitialAlign (estimator.cpp:221-455) for a window in which every frame of
    all_image_frame is a key frame.  reintegrate(bg) -> list of IntegrationOracle of the F-1 intervals linearised at
    (ba = 0, bg).  Returns None on failure, else dict(Ps, Rs, Vs, Bgs, g, depth (per feature, -1 where untouched), l, s)."""
    from . import prestep as OP

    ok, rel_R, rel_T, l = relative_pose(corres_fn, F - 1, seed, attempt)
    if not ok:
        return None
    sfm = global_sfm(F, l, rel_R, rel_T, obs_first, obs_frame, obs_pts)
    if not sfm["ok"]:
        return None
    fR = np.array([_qrot(sfm["q"][i]) @ ric.T for i in range(F)])      # frame.R = Q[i] * RIC^T  (:301)
    fT = sfm["T"].copy()
    pre = reintegrate(np.asarray(bgs0, float))
    dbg = solve_gyroscope_bias(fR, [p.delta_q for p in pre], [p.jacobian[3:6, 12:15] for p in pre])
    Bgs = np.asarray(bgs0, float) + dbg
    pre = reintegrate(Bgs)
    ok, g, x = linear_alignment(fR, fT, [p.delta_p for p in pre], [p.delta_v for p in pre], [p.sum_dt for p in pre], tic, g_norm)
    if not ok:
        return None
    Ps, Rs = fT.copy(), fR.copy()
    L = len(obs_first) - 1
    start = np.array([obs_frame[obs_first[j]] for j in range(L)])
    depth = np.full(L, -1.0)
    depth = OP.triangulate(Ps, Rs, np.zeros(3), ric, start, obs_first, obs_pts, depth, init_depth)   # TIC_TMP = 0 (:405-410)
    s = x[-1]
    base = s * Ps[0] - Rs[0] @ tic
    Ps = np.array([s * Ps[i] - Rs[i] @ tic - base for i in range(F)])
    Vs = np.array([Rs[k] @ x[3 * k:3 * k + 3] for k in range(F)])
    elig = np.array([(obs_first[j + 1] - obs_first[j] >= 2) and (start[j] < F - 1 - 2) for j in range(L)])
    depth = np.where(elig, depth * s, depth)
    R0 = g2R(g)
    R0 = _yaw_rot(-_R2ypr_yaw_deg(R0 @ Rs[0])) @ R0
    g = R0 @ g
    Ps = Ps @ R0.T
    Rs = np.array([R0 @ Rs[i] for i in range(F)])
    Vs = Vs @ R0.T
    return dict(Ps=Ps, Rs=Rs, Vs=Vs, Bgs=Bgs, g=g, depth=depth, l=l, s=s)


# ------------------------------------------------------------------------------------------------------------------
# InitialEXRotation (ESTIMATE_EXTRINSIC == 2): extrinsic rotation from camera / IMU rotation pairs
# ------------------------------------------------------------------------------------------------------------------
def solve_relative_r(corres, seed, frame):
    """InitialEXRotation::solveRelativeR (initial/initial_ex_rotation.cpp:220-289).  The reference decomposes E with cv::SVD
    and, when the first candidate is a reflection, decomposes -E instead (:268-272); both candidates are then proper
    rotations.  Restated with the candidates made proper directly (decompose_essential); the reference's "R1" is built with
    the transposed W (:312-317), i.e. it is decompose_essential's R2 — which only matters for the tie of `ratio1 > ratio2`."""
    corres = np.asarray(corres, float).reshape(-1, 6)
    if len(corres) < 9:
        return np.eye(3)
    b1, b2 = corres[:, :3], corres[:, 3:]
    E = R8.find_fundamental(b1, b2, seed, frame)[0]
    Ra, Rb, t = decompose_essential(E)              # reference order: R1 = Rb, R2 = Ra
    n = len(corres)
    ratio1 = max(cheirality_count(Rb, t, b1, b2), cheirality_count(Rb, -t, b1, b2)) / n    # testTriangulation :291-318
    ratio2 = max(cheirality_count(Ra, t, b1, b2), cheirality_count(Ra, -t, b1, b2)) / n
    ans = Rb if ratio1 > ratio2 else Ra
    return ans.T                                    # ans_R_eigen(j, i) = ans_R_cv(i, j)  (:283-287)


def _angular_distance(qa, qb):
    """Eigen::Quaternion::angularDistance, (w, x, y, z)."""
    d = _qmul(qa, np.array([qb[0], -qb[1], -qb[2], -qb[3]]))
    return 2.0 * np.arctan2(np.linalg.norm(d[1:]), abs(d[0]))


def ex_rotation_solve(Rc, Rimu, Rc_g):
    """The linear system of CalibrationExRotation (initial_ex_rotation.cpp:20-60) for the pairs i = 1 .. N (arrays without the
    identity the constructor pushes).  Returns (ric, singular values (4,))."""
    N = len(Rc)
    A = np.zeros((4 * N, 4))
    for i in range(N):
        r1, r2 = rot_to_quat(Rc[i]), rot_to_quat(Rc_g[i])
        ang = 180.0 / np.pi * _angular_distance(r1, r2)
        huber = 5.0 / ang if ang > 5.0 else 1.0
        w, q = r1[0], r1[1:]
        Lm = np.zeros((4, 4))
        Lm[:3, :3] = w * np.eye(3) + np.array([[0, -q[2], q[1]], [q[2], 0, -q[0]], [-q[1], q[0], 0]])
        Lm[:3, 3], Lm[3, :3], Lm[3, 3] = q, -q, w
        rq = rot_to_quat(Rimu[i])
        w, q = rq[0], rq[1:]
        Rm = np.zeros((4, 4))
        Rm[:3, :3] = w * np.eye(3) - np.array([[0, -q[2], q[1]], [q[2], 0, -q[0]], [-q[1], q[0], 0]])
        Rm[:3, 3], Rm[3, :3], Rm[3, 3] = q, -q, w
        A[4 * i:4 * i + 4] = huber * (Lm - Rm)
    _, sv, Vt = np.linalg.svd(A)
    x = Vt[3]                                       # matrixV().col(3); Quaterniond(Vector4d) reads (x, y, z, w)
    ric = _qrot(np.array([x[3], x[0], x[1], x[2]]) / np.linalg.norm(x)).T   # toRotationMatrix().inverse()
    sv4 = np.zeros(4)
    sv4[:len(sv)] = sv
    return ric, sv4


class InitialEXRotationOracle:
    """InitialEXRotation (initial_ex_rotation.h:16-45, .cpp:4-68)."""

    def __init__(self, window_size=10):
        self.frame_count, self.window_size = 0, window_size
        self.Rc, self.Rimu, self.Rc_g = [], [], []
        self.ric = np.eye(3)

    def calibrate(self, corres, delta_q_imu, seed):
        """delta_q_imu (w, x, y, z).  Returns (ok, ric)."""
        self.frame_count += 1
        self.Rc.append(solve_relative_r(corres, seed, self.frame_count))
        Ri = _qrot(np.asarray(delta_q_imu, float))
        self.Rimu.append(Ri)
        self.Rc_g.append(self.ric.T @ Ri @ self.ric)            # ric.inverse() * delta_q_imu * ric
        self.ric, sv = ex_rotation_solve(self.Rc, self.Rimu, self.Rc_g)
        return bool(self.frame_count >= self.window_size and sv[2] > 0.25), self.ric
