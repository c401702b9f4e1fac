// The "full BA" of GlobalSFM::construct (vins_estimator/src/initial/initial_sfm.cpp:230-297) as the Ceres 1.12
// trust-region Levenberg-Marquardt loop with DENSE_SCHUR (third party, pinned by docker/Dockerfile:3, restated from its
// published algorithm: trust_region_minimizer.cc, levenberg_marquardt_strategy.cc, schur_eliminator).
//
// Residual: ReprojectionError3D (initial_sfm.h:26-66).  Parameter blocks: per frame a quaternion (ceres::Quaternion-
// Parameterization, 3 local dof) and a translation, per triangulated feature a 3-D point.  Constant: rotation l and
// translations l, F-1 (initial_sfm.cpp:243-251).  The points are the Schur-eliminated blocks; the reduced camera system
// is D = 6 F wide (constant / unobserved columns get an identity row).
//
// This header holds the per-observation arithmetic (ba_obs, ba_w, the LM diagonal, Jacobi scaling), the camera / point
// updates and the trust-region decision as __host__ __device__ functions, plus a scatter formulation of the Schur
// elimination per point (ba_point_linearize / _schur / _backsub, accumulation primitive passed in) that the host test shim
// tests/native/init_math_shim.cpp runs serially, so the arithmetic and the control flow are checked on the CPU against
// oracle/initial.py.  sfm_ba_kernel uses the same helpers and decisions but GATHERS the reduced system per entry (no
// atomics, every Jacobian evaluated once per iteration); the two formulations differ in summation order only.
#pragma once
#include "init_math.cuh"

namespace lfvio {
namespace im {

constexpr int kBaMaxIter = 50;  // Ceres default max_num_iterations (the reference's 0.2 s wall-clock cap is not restated)

struct BaProblem {
    int F, L, l;
    const int *obs_first, *obs_frame;
    const double *obs_pts;
    const unsigned char *state;
};

struct BaCtl {
    double radius, decrease, cost, cand_cost, model_change, step2, xnorm2, gmax;
    int iter, n_success, converged, done, chol_ok;
    double costs[kBaMaxIter + 1];
};

LFVIO_HD bool ba_rot_free(const BaProblem &P, int f) { return f != P.l; }
LFVIO_HD bool ba_tr_free(const BaProblem &P, int f) { return f != P.l && f != P.F - 1; }

// residual and the masked 3 x 6 camera Jacobian [rotation | translation] + 3 x 3 point Jacobian of observation o
LFVIO_HD void ba_obs(const BaProblem &P, const double *cq, const double *ct, const double *pt, int o, double *r, double *Jc,
                     double *Jp) {
    const int f = P.obs_frame[o];
    double Jr[9], Jt[9];
    ba_residual_jac(cq + 4 * f, ct + 3 * f, pt, P.obs_pts + 3 * o, true, r, Jr, Jt, Jp);
    const double mr = ba_rot_free(P, f) ? 1.0 : 0.0, mt = ba_tr_free(P, f) ? 1.0 : 0.0;
    for (int i = 0; i < 3; i++)
        for (int c = 0; c < 3; c++) {
            Jc[6 * i + c] = mr * Jr[3 * i + c];
            Jc[6 * i + 3 + c] = mt * Jt[3 * i + c];
        }
}

LFVIO_HD void ba_init(BaCtl *c) {
    c->radius = 1e4;  // initial_trust_region_radius
    c->decrease = 2.0;
    c->cost = c->cand_cost = c->model_change = c->step2 = c->xnorm2 = c->gmax = 0.0;
    c->iter = c->n_success = c->converged = c->done = 0;
    c->chol_ok = 1;
}

// Phase 1 (per point): accumulate U (block diagonal of S, full 6 x 6 blocks), g_c, the point's V (3 x 3) and g_p; returns
// the point's share of the cost and tracks max |g_p|.
template <class Add>
LFVIO_HD double ba_point_linearize(const BaProblem &P, const double *cq, const double *ct, const double *pts, int j, double *S,
                                   double *gc, double *V, double *gp, double *gmax, Add add) {
    const int D = 6 * P.F;
    double cost = 0;
    for (int k = 0; k < 9; k++) V[k] = 0;
    gp[0] = gp[1] = gp[2] = 0;
    for (int o = P.obs_first[j]; o < P.obs_first[j + 1]; o++) {
        const int f = P.obs_frame[o];
        double r[3], Jc[18], Jp[9];
        ba_obs(P, cq, ct, pts + 3 * j, o, r, Jc, Jp);
        cost += 0.5 * (r[0] * r[0] + r[1] * r[1] + r[2] * r[2]);
        for (int a = 0; a < 6; a++) {
            for (int b = 0; b < 6; b++) add(&S[(6 * f + a) * D + 6 * f + b], Jc[a] * Jc[b] + Jc[6 + a] * Jc[6 + b] + Jc[12 + a] * Jc[12 + b]);
            add(&gc[6 * f + a], Jc[a] * r[0] + Jc[6 + a] * r[1] + Jc[12 + a] * r[2]);
        }
        for (int a = 0; a < 3; a++) {
            for (int b = 0; b < 3; b++) V[3 * a + b] += Jp[a] * Jp[b] + Jp[3 + a] * Jp[3 + b] + Jp[6 + a] * Jp[6 + b];
            gp[a] += Jp[a] * r[0] + Jp[3 + a] * r[1] + Jp[6 + a] * r[2];
        }
    }
    for (int a = 0; a < 3; a++) *gmax = fmax(*gmax, fabs(gp[a]));
    return cost;
}

// Jacobi scaling of a column from its squared norm (trust_region_minimizer.cc EstimateScale): 1 / (1 + sqrt(sum J^2))
LFVIO_HD double ba_jacobi_scale(double col_sq_norm) { return 1.0 / (1.0 + sqrt(col_sq_norm)); }
// Levenberg-Marquardt diagonal (levenberg_marquardt_strategy.cc): clamp(diag(J^T J), 1e-6, 1e32) / radius
LFVIO_HD double ba_lm_diag(double d, double radius) { return fmin(fmax(d, 1e-6), 1e32) / radius; }

// scaled W_o = Jc_s^T Jp_s (6 x 3) of observation o
LFVIO_HD void ba_w(const double *Jc, const double *Jp, const double *sc /* 6 */, const double *sp /* 3 */, double *W) {
    for (int a = 0; a < 6; a++)
        for (int b = 0; b < 3; b++) W[3 * a + b] = sc[a] * sp[b] * (Jc[a] * Jp[b] + Jc[6 + a] * Jp[3 + b] + Jc[12 + a] * Jp[6 + b]);
}

// Phase 3 (per point): scaled, damped V -> Vinv (stored), Schur complement into S (lower block triangle + full diagonal
// blocks) and g_c.  V, gp: this point's UNSCALED blocks from phase 1; sp: the point's three Jacobi scales.
template <class Add>
LFVIO_HD void ba_point_schur(const BaProblem &P, const double *cq, const double *ct, const double *pts, int j, const double *V,
                             const double *gp, const double *scale_c, const double *sp, double radius, double *S, double *gc,
                             double *Vinv, Add add) {
    const int D = 6 * P.F;
    double Vs[9], gps[3];
    for (int a = 0; a < 3; a++) {
        for (int b = 0; b < 3; b++) Vs[3 * a + b] = sp[a] * sp[b] * V[3 * a + b];
        gps[a] = sp[a] * gp[a];
    }
    for (int a = 0; a < 3; a++) Vs[4 * a] += ba_lm_diag(Vs[4 * a], radius);
    inv3(Vs, Vinv);
    for (int o1 = P.obs_first[j]; o1 < P.obs_first[j + 1]; o1++) {
        const int f1 = P.obs_frame[o1];
        double r[3], Jc[18], Jp[9], W1[18], Y[18];
        ba_obs(P, cq, ct, pts + 3 * j, o1, r, Jc, Jp);
        ba_w(Jc, Jp, scale_c + 6 * f1, sp, W1);
        for (int a = 0; a < 6; a++)
            for (int b = 0; b < 3; b++) Y[3 * a + b] = W1[3 * a] * Vinv[b] + W1[3 * a + 1] * Vinv[3 + b] + W1[3 * a + 2] * Vinv[6 + b];
        for (int a = 0; a < 6; a++) add(&gc[6 * f1 + a], -(Y[3 * a] * gps[0] + Y[3 * a + 1] * gps[1] + Y[3 * a + 2] * gps[2]));
        for (int o2 = o1; o2 < P.obs_first[j + 1]; o2++) {
            const int f2 = P.obs_frame[o2];
            double W2[18];
            if (o2 == o1) {
                for (int k = 0; k < 18; k++) W2[k] = W1[k];
            } else {
                double r2[3], Jc2[18], Jp2[9];
                ba_obs(P, cq, ct, pts + 3 * j, o2, r2, Jc2, Jp2);
                ba_w(Jc2, Jp2, scale_c + 6 * f2, sp, W2);
            }
            for (int a = 0; a < 6; a++)
                for (int b = 0; b < 6; b++) {
                    const double v = Y[3 * a] * W2[3 * b] + Y[3 * a + 1] * W2[3 * b + 1] + Y[3 * a + 2] * W2[3 * b + 2];
                    add(&S[(6 * f2 + b) * D + 6 * f1 + a], -v);  // block (f2, f1), f2 >= f1: lower triangle
                    if (f2 == f1 && o2 != o1) add(&S[(6 * f1 + a) * D + 6 * f2 + b], -v);  // two observations in one frame
                }
        }
    }
}

// Phase 5 (per point): back-substitution, candidate point, this point's share of the model cost change, |step|^2, |x|^2.
// dxc: the SCALED camera step (solution of the reduced system, already negated).
LFVIO_HD void ba_point_backsub(const BaProblem &P, const double *cq, const double *ct, const double *pts, int j, const double *gp,
                               const double *scale_c, const double *sp, const double *Vinv, const double *dxc, double *cand,
                               double *model_change, double *step2, double *xnorm2) {
    double wdx[3] = {0, 0, 0};
    for (int o = P.obs_first[j]; o < P.obs_first[j + 1]; o++) {
        const int f = P.obs_frame[o];
        double r[3], Jc[18], Jp[9], W[18];
        ba_obs(P, cq, ct, pts + 3 * j, o, r, Jc, Jp);
        ba_w(Jc, Jp, scale_c + 6 * f, sp, W);
        for (int b = 0; b < 3; b++)
            for (int a = 0; a < 6; a++) wdx[b] += W[3 * a + b] * dxc[6 * f + a];
    }
    double rhs[3], dxp[3];
    for (int a = 0; a < 3; a++) rhs[a] = sp[a] * gp[a] + wdx[a];
    for (int a = 0; a < 3; a++) dxp[a] = -(Vinv[3 * a] * rhs[0] + Vinv[3 * a + 1] * rhs[1] + Vinv[3 * a + 2] * rhs[2]);
    for (int o = P.obs_first[j]; o < P.obs_first[j + 1]; o++) {
        const int f = P.obs_frame[o];
        double r[3], Jc[18], Jp[9];
        ba_obs(P, cq, ct, pts + 3 * j, o, r, Jc, Jp);
        for (int i = 0; i < 3; i++) {
            double m = 0;
            for (int a = 0; a < 6; a++) m += Jc[6 * i + a] * scale_c[6 * f + a] * dxc[6 * f + a];
            for (int a = 0; a < 3; a++) m += Jp[3 * i + a] * sp[a] * dxp[a];
            *model_change += -(m * (r[i] + m / 2.0));
        }
    }
    for (int a = 0; a < 3; a++) {
        const double d = sp[a] * dxp[a];
        cand[3 * j + a] = pts[3 * j + a] + d;
        *step2 += d * d;
        *xnorm2 += pts[3 * j + a] * pts[3 * j + a];
    }
}

// candidate camera f from the scaled step; adds its |x_plus - x|^2 and |x|^2 (free blocks only, ambient coordinates)
LFVIO_HD void ba_cam_candidate(const BaProblem &P, const double *cq, const double *ct, int f, const double *scale_c, const double *dxc,
                               double *cand_q, double *cand_t, double *step2, double *xnorm2) {
    double d[3];
    for (int a = 0; a < 3; a++) d[a] = scale_c[6 * f + a] * dxc[6 * f + a];
    if (ba_rot_free(P, f)) {
        quat_plus(cq + 4 * f, d, cand_q + 4 * f);
        for (int k = 0; k < 4; k++) {
            const double e = cand_q[4 * f + k] - cq[4 * f + k];
            *step2 += e * e;
            *xnorm2 += cq[4 * f + k] * cq[4 * f + k];
        }
    } else {
        for (int k = 0; k < 4; k++) cand_q[4 * f + k] = cq[4 * f + k];
    }
    for (int a = 0; a < 3; a++) {
        const double e = ba_tr_free(P, f) ? scale_c[6 * f + 3 + a] * dxc[6 * f + 3 + a] : 0.0;
        cand_t[3 * f + a] = ct[3 * f + a] + e;
        if (ba_tr_free(P, f)) {
            *step2 += e * e;
            *xnorm2 += ct[3 * f + a] * ct[3 * f + a];
        }
    }
}

LFVIO_HD double ba_point_cost(const BaProblem &P, const double *cq, const double *ct, const double *pts, int j) {
    double cost = 0;
    for (int o = P.obs_first[j]; o < P.obs_first[j + 1]; o++) {
        const int f = P.obs_frame[o];
        double r[3];
        ba_residual_jac(cq + 4 * f, ct + 3 * f, pts + 3 * j, P.obs_pts + 3 * o, false, r, nullptr, nullptr, nullptr);
        cost += 0.5 * (r[0] * r[0] + r[1] * r[1] + r[2] * r[2]);
    }
    return cost;
}

// Called once per iteration right after phase 1 (c->cost, c->gmax valid): iteration bookkeeping + gradient tolerance.
// Returns false when the loop is over.
LFVIO_HD bool ba_begin_iteration(BaCtl *c) {
    if (c->iter == 0) c->costs[0] = c->cost;
    if (c->gmax <= 1e-10) {  // gradient_tolerance
        c->converged = 1;
        c->done = 1;
        return false;
    }
    if (c->iter >= kBaMaxIter) {
        c->done = 1;
        return false;
    }
    c->iter++;
    return true;
}

// Trust-region decision after the candidate cost is known (trust_region_minimizer.cc of Ceres 1.12: parameter tolerance,
// function tolerance, relative decrease, LevenbergMarquardtStrategy::StepAccepted / StepRejected).  Returns 1 when the
// candidate is to be accepted.
LFVIO_HD int ba_decide(BaCtl *c) {
    if (!c->chol_ok || !(c->model_change > 0.0)) {  // StepIsInvalid
        c->radius /= c->decrease;
        c->decrease *= 2.0;
        if (c->radius < 1e-32) c->done = 1;
        return 0;
    }
    if (sqrt(c->step2) <= 1e-8 * (sqrt(c->xnorm2) + 1e-8)) {  // parameter_tolerance
        c->converged = 1;
        c->done = 1;
        return 0;
    }
    const double change = c->cost - c->cand_cost;
    if (fabs(change) <= 1e-6 * c->cost) {  // function_tolerance
        c->converged = 1;
        c->done = 1;
        return 0;
    }
    const double rho = change / c->model_change;
    if (rho > 1e-3) {  // min_relative_decrease
        const double t = 2.0 * rho - 1.0;
        c->radius = fmin(1e16, c->radius / fmax(1.0 / 3.0, 1.0 - t * t * t));
        c->decrease = 2.0;
        c->cost = c->cand_cost;
        c->n_success++;
        c->costs[c->n_success] = c->cand_cost;
        return 1;
    }
    c->radius /= c->decrease;
    c->decrease *= 2.0;
    if (c->radius < 1e-32) c->done = 1;
    return 0;
}

}  // namespace im
}  // namespace lfvio
