// K12/K13 — one trust-region step: Cauchy point, dense Cholesky solve of the reduced system, landmark back-substitution,
// traditional dogleg, model cost change and candidate state — one launch (three for very large windows), device-side control flow.
//
// Reference: ceres::Solve at vins_estimator/src/estimator.cpp:810-825 (DENSE_SCHUR + DOGLEG); Ceres 1.12.0
// DoglegStrategy::ComputeStep / ComputeCauchyPoint / ComputeGaussNewtonStep / ComputeTraditionalDoglegStep and the
// TrustRegionMinimizer step evaluation are restated in SURVEY.md Appendix B.4 (Ceres is third-party, not vendored).
//
//   chol_kernel  (one CTA)    the whole step for windows with up to kFuseMaxL landmarks:
//                             - diagonal_, gradient_, the Cauchy point's |J_s (gradient_/diagonal_)|^2 evaluated on the
//                               reduced system itself (S quadratic form on the way into shared memory + one dot product
//                               per landmark) instead of a pass over the residual rows;
//                             - reduced system as a packed upper triangle in shared memory (+ LM term); blocked (16)
//                               right-looking Cholesky: the 16x16 diagonal blocks are factored by ONE WARP IN REGISTERS
//                               (lane = column, pivots/rows exchanged with warp shuffles) one block ahead of the trailing
//                               update; the right-hand side rides along as a panel column; blocked back substitution;
//                             - landmark back-substitution y_l = (g_l - w_l^T y_p) / c_l (warp per landmark);
//                             - cand_tail: traditional dogleg, model cost change from the strategy's own scalars (see
//                               cand_tail), x_cand = Plus(x, step * scale), step published to the control block.
//   pre_kernel / gnlm_kernel  (multi-CTA) only for larger windows (config 4: L ~ 3.6 k): row-wise Cauchy point before,
//                             landmark back-substitution after, cand_tail in gnlm_kernel's last CTA.
#include "backend.cuh"
#include "solver_common.cuh"
#include "chol16.cuh"
#include "chol_blocked.cuh"

namespace lfvio {

namespace {

constexpr int CHOL_THREADS = 512;
constexpr int PRE_THREADS = 256;
constexpr int kFuseMaxL = 1024;  // up to this many landmarks the whole step runs in chol_kernel
constexpr int GNLM_WARPS = 8;

__device__ __forceinline__ long long gtimer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return (long long)t;
}
#define PHASE_MARK(i) do { if (threadIdx.x == 0) ctl->phase_clk[i] = gtimer_ns(); } while (0)

__device__ __forceinline__ double block_sum(double v, double *s_red) {  // any blockDim (multiple of 32, <= 1024)
    v = warp_sum(v);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = v;
    __syncthreads();
    double t = 0;
    if (threadIdx.x < 32) {
        t = (threadIdx.x < (blockDim.x >> 5)) ? s_red[threadIdx.x] : 0.0;
        t = warp_sum(t);
        if (threadIdx.x == 0) s_red[32] = t;
    }
    __syncthreads();
    return s_red[32];
}
__device__ __forceinline__ double block_max(double v, double *s_red) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    __syncthreads();
    if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = v;
    __syncthreads();
    double t = 0;
    if (threadIdx.x < 32) {
        t = (threadIdx.x < (blockDim.x >> 5)) ? s_red[threadIdx.x] : 0.0;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) t = fmax(t, __shfl_xor_sync(0xffffffffu, t, o));
        if (threadIdx.x == 0) s_red[32] = t;
    }
    __syncthreads();
    return s_red[32];
}

__device__ __forceinline__ void build_prior_map(const DevProblem &P, int *pmap) {
    const DevPrior &p = P.prior;
    for (int b = threadIdx.x; b < p.n_blocks; b += blockDim.x) {
        const int ls = P.lay.block_lsize(p.block_id[b]), c0 = P.lay.block_col(p.block_id[b]);
        for (int q = 0; q < ls; q++) pmap[p.block_idx[b] + q] = c0 + q;
    }
}

// Grid-strided sum over all residual rows of f(J_s v, r): vfun(col) returns v[col] (already including nothing of scale).
// mode 0: sum (Jv)^2        mode 1: sum -(Jv) (r + Jv / 2)
template <class VF>
__device__ double rows_partial(const DevProblem &P, const EvalBuf &e, const double *scale, VF vfun, int mode, const int *pmap) {
    const StateLayout lay = P.lay;
    const int gtid = blockIdx.x * blockDim.x + threadIdx.x, nth = gridDim.x * blockDim.x;
    double acc = 0;
    for (int k = gtid; k < P.n_proj; k += nth) {
        const lfvio_proj_factor &f = P.proj[k];
        const double *J = e.Jp + 40 * (size_t)k;
        double s0 = 0, s1 = 0;
#pragma unroll 4
        for (int c = 0; c < 20; c++) {
            const int col = proj_col(lay, f.frame_i, f.frame_j, f.landmark, c);
            const double vv = scale[col] * vfun(col);
            s0 += J[c] * vv;
            s1 += J[20 + c] * vv;
        }
        if (mode == 0) acc += s0 * s0 + s1 * s1;
        else acc -= s0 * (e.rp[2 * (size_t)k] + s0 / 2.0) + s1 * (e.rp[2 * (size_t)k + 1] + s1 / 2.0);
    }
    for (int q = gtid; q < P.n_imu * 15; q += nth) {
        const int k = q / 15, r = q % 15;
        if (P.imu[k].sum_dt > 10.0) continue;
        const double *J = e.Ji + 450 * (size_t)k + 30 * r;
        double s = 0;
        for (int c = 0; c < 30; c++) {
            const int col = imu_col(lay, P.imu[k].frame_i, c);
            s += J[c] * scale[col] * vfun(col);
        }
        if (mode == 0) acc += s * s;
        else acc -= s * (e.ri[q] + s / 2.0);
    }
    const DevPrior &p = P.prior;
    for (int i = gtid; i < p.n; i += nth) {
        double s = 0;
        for (int a = 0; a < p.n; a++) s += p.J0[(size_t)i * p.n + a] * scale[pmap[a]] * vfun(pmap[a]);
        if (mode == 0) acc += s * s;
        else acc -= s * (e.rm[i] + s / 2.0);
    }
    return acc;
}

// =====================================================================================================
__global__ void __launch_bounds__(PRE_THREADS) pre_kernel(const DevProblem *__restrict__ Pd, EvalBuf e0, EvalBuf e1, SolverScratch w, SolverCtl *ctl) {
    const DevProblem &P = *Pd;
    pdl_prologue();
    if (ctl->done != 0 || ctl->reuse != 0) return;
    __shared__ double s_red[33];
    __shared__ int pmap[15 * kMaxFrames + 8];
    const EvalBuf &e = ctl->cur ? e1 : e0;
    const int D = P.lay.D(), N = D + P.lay.L;
    build_prior_map(P, pmap);
    __syncthreads();
    const int gtid = blockIdx.x * blockDim.x + threadIdx.x, nth = gridDim.x * blockDim.x;
    double gg = 0, gmax = 0;
    for (int c = gtid; c < N; c += nth) {
        const double d2 = c < D ? w.diag2[c] : w.lm_c[c - D];
        const double g = c < D ? w.grad[c] : w.lm_g[c - D];
        const double dg = sqrt(clamp_diag(d2));
        w.diagonal[c] = dg;
        const double gs = g / dg;
        w.gradient[c] = gs;
        gg += gs * gs;
        gmax = fmax(gmax, fabs(g / w.scale[c]));  // un-scaled gradient J^T r
    }
    // Cauchy point: J_s (gradient_ / diagonal_) with gradient_/diagonal_ = (J_s^T r) / clamp(diag2), taken on the fly
    auto vfun = [&](int col) {
        const double d2 = col < D ? w.diag2[col] : w.lm_c[col - D];
        const double g = col < D ? w.grad[col] : w.lm_g[col - D];
        return g / clamp_diag(d2);
    };
    double jj = rows_partial(P, e, w.scale, vfun, 0, pmap);
    gg = block_sum(gg, s_red);
    jj = block_sum(jj, s_red);
    gmax = block_max(gmax, s_red);
    if (threadIdx.x == 0) {
        atomicAdd(&ctl->acc[0], gg);
        atomicAdd(&ctl->acc[1], jj);
        atomicMax(reinterpret_cast<unsigned long long *>(&ctl->acc[2]), (unsigned long long)__double_as_longlong(gmax));
    }
}

// =====================================================================================================
// ---- shared tail: traditional dogleg, model cost change, candidate state -----------------------------------------
// Whole CTA.  gn2 = |gauss_newton_step_|^2 and gdot = gradient_ . gauss_newton_step_ of THIS solve (ignored when reuse_).
//
// Model cost change without a pass over the residual rows.  With H = J_s^T J_s, g = J_s^T r, a = g / clamp(diag H)
// (Cauchy direction), y the solution of (H + mu diag) y = g, the step in J_s coordinates is v = cg a - cn y and
//   v^T g   = cg |gradient_|^2 + cn gdot                      (a^T g = |gradient_|^2,  y^T g = -gdot)
//   v^T H v = cg^2 a^T H a - 2 cg cn a^T H y + cn^2 y^T H y   (H y = g - mu diag y)
//           = cg^2 jj      - 2 cg cn (gg + mu gdot) + cn^2 (-gdot - mu gn2)
// so   model_cost_change = -(J v)^T (r + J v / 2) = -v^T g - v^T H v / 2   (ceres TrustRegionMinimizer) needs only the
// scalars the strategy already has; jj = |J_s a|^2 is the Cauchy-point denominator.
__device__ void cand_tail(const DevProblem &P, const SolverScratch &w, double *x0, double *x1, SolverCtl *ctl, double gn2_in,
                          double gdot_in, double *s_red, double *s_delta /* D + L, shared */) {
    const StateLayout lay = P.lay;
    const lfvio_solver_opts &o = P.o;
    const int D = lay.D(), L = lay.L;
    const int tid = threadIdx.x, nt = blockDim.x;
    const int cur = ctl->cur, reuse = ctl->reuse;
    const double *x = cur ? x1 : x0;
    double *xc = cur ? x0 : x1;
    const int ok = reuse ? 1 : ctl->pad;
    const double gnorm = ctl->gnorm, alpha = ctl->alpha, radius = ctl->radius, jj = ctl->jj, mu = ctl->mu_solve;
    const double gn2 = reuse ? ctl->gn_norm * ctl->gn_norm : gn2_in;
    const double gdot = reuse ? ctl->gdot : gdot_in;
    const double gnn = sqrt(gn2);
    double cg = 0, cn = 0, dsn = 0;  // step = cg * gradient_ + cn * gauss_newton_step_
    if (gnn <= radius) {
        cn = 1.0;
        dsn = gnn;
    } else if (gnorm * alpha >= radius) {
        cg = -(radius / gnorm);
        dsn = radius;
    } else {
        const double b_dot_a = -alpha * gdot;
        const double a_sq = (alpha * gnorm) * (alpha * gnorm);  // (pow(x, 2.0) of the original: exact either way)
        const double bma_sq = a_sq - 2 * b_dot_a + gnn * gnn;
        const double cc = b_dot_a - a_sq;
        const double dd = sqrt(cc * cc + bma_sq * (radius * radius - a_sq));
        const double beta = (cc <= 0) ? (dd - cc) / bma_sq : (radius * radius - a_sq) / (dd + cc);
        cg = -alpha * (1.0 - beta);
        cn = beta;
        dsn = sqrt(fmax(cg * cg * gnorm * gnorm + 2.0 * cg * cn * gdot + cn * cn * gn2, 0.0));
    }
    const double gg = gnorm * gnorm;
    const double vg = cg * gg + cn * gdot;
    const double vHv = cg * cg * jj - 2.0 * cg * cn * (gg + mu * gdot) + cn * cn * (-gdot - mu * gn2);
    const double model_cost_change = ok ? (-vg - 0.5 * vHv) : 0.0;
    double dx2 = 0;
    // ---- candidate x_plus_delta = Plus(x, step * scale): the scaled step of every column first (all threads, one round trip to
    // the four column arrays), then poses (warp 0, one frame per lane), speed / biases (the other warps, one element per
    // thread), extrinsic + td (last thread) and landmarks side by side ----
    // (s_delta == nullptr: gnlm_kernel's last CTA, windows too large for the staging buffer — columns straight from global)
    auto delta_global = [&](int col) { return (cg * w.gradient[col] + cn * w.gn[col]) / w.diagonal[col] * w.scale[col]; };
    if (ok && s_delta)
        for (int col = tid; col < D + L; col += nt) s_delta[col] = delta_global(col);
    __syncthreads();
    auto delta = [&](int col) { return s_delta ? s_delta[col] : delta_global(col); };
    if (ok) {
        for (int f = tid; f < lay.F; f += nt) {
            double d[6];
#pragma unroll
            for (int a = 0; a < 6; a++) d[a] = delta(lay.col_pose(f) + a);
            double xi[7], xo[7];  // (the candidate is formed in registers: reading it back from global costs a round trip)
#pragma unroll
            for (int a = 0; a < 7; a++) xi[a] = x[lay.pose(f) + a];
            pose_plus(xi, d, xo);
#pragma unroll
            for (int a = 0; a < 7; a++) {
                xc[lay.pose(f) + a] = xo[a];
                const double df = xi[a] - xo[a];
                dx2 += df * df;
            }
        }
        for (int q = tid - 32; q < 9 * lay.F; q += nt) {
            if (q < 0) continue;
            const int f = q / 9, a = q - 9 * f;
            const double xv = x[lay.sb(f) + a];
            const double nv = xv + delta(lay.col_sb(f) + a);
            xc[lay.sb(f) + a] = nv;
            const double df = xv - nv;
            dx2 += df * df;
        }
        if (tid == nt - 1) {  // extrinsic + td
            if (o.estimate_extrinsic) {
                double d[6];
#pragma unroll
                for (int a = 0; a < 6; a++) d[a] = delta(lay.col_ex() + a);
                double xi[7], xo[7];
#pragma unroll
                for (int a = 0; a < 7; a++) xi[a] = x[lay.ex() + a];
                pose_plus(xi, d, xo);
#pragma unroll
                for (int a = 0; a < 7; a++) {
                    xc[lay.ex() + a] = xo[a];
                    const double df = xi[a] - xo[a];
                    dx2 += df * df;
                }
            } else {
                for (int a = 0; a < 7; a++) xc[lay.ex() + a] = x[lay.ex() + a];
            }
            if (o.estimate_td) {
                const double xv = x[lay.td()];
                const double nv = xv + delta(lay.col_td());
                xc[lay.td()] = nv;
                dx2 += (xv - nv) * (xv - nv);
            } else {
                xc[lay.td()] = x[lay.td()];
            }
        }
        for (int l = tid; l < L; l += nt) {
            const double xv = x[lay.feat(l)];
            const double nv = xv + delta(D + l);
            xc[lay.feat(l)] = nv;
            dx2 += (xv - nv) * (xv - nv);
        }
    }
    if (tid == 0) ctl->phase_clk[15] = gtimer_ns();
    dx2 = block_sum(dx2, s_red);
    if (tid == 0) {
        const double step_norm = sqrt(dx2);
        const int step_valid = ok && (model_cost_change > 0.0);
        if (!reuse) {
            ctl->gn_norm = gnn;
            ctl->gdot = gdot;
        }
        ctl->dogleg_step_norm = dsn;
        ctl->model_cost_change = model_cost_change;
        ctl->step_norm = step_norm;
        ctl->step_valid = step_valid;
        ctl->cost_cand = 0.0;
        ctl->eval_pending = step_valid;
        if (step_valid && !o.fixed_iterations && step_norm <= o.parameter_tolerance * (ctl->x_norm + o.parameter_tolerance)) {
            ctl->done = 1;  // "Parameter tolerance reached": terminates without taking the step
            ctl->eval_pending = 0;
        }
    }
}

// w_l^T v_p for one landmark by ONE thread: v holds the reduced columns (shared or global).  Index arrays come from
// accumulate_kernel, so the only dependent load is lm_first.
__device__ __forceinline__ double landmark_dot(const DevProblem &P, const SolverScratch &w, const double *v, int l, int col_ex,
                                               int col_td) {
    const int ci = w.lm_coli[l];
    if (ci < 0) return 0.0;
    const int k0 = P.lm_first[l], k1 = P.lm_first[l + 1];
    const double *ws = w.lm_ws + 13 * (size_t)l;
    double s = ws[12] * v[col_td];
#pragma unroll
    for (int a = 0; a < 6; a++) s = fma(ws[a], v[ci + a], fma(ws[6 + a], v[col_ex + a], s));
    // four factors per pass, their index and weight loads issued together (one factor at a time serialises one L2 round trip each)
    for (int k = k0; k < k1; k += 4) {
        int cj[4];
        double wv[4][6];
#pragma unroll
        for (int u = 0; u < 4; u++) {
            const int kk = min(k + u, k1 - 1);
            cj[u] = w.proj_colj[kk];
#pragma unroll
            for (int a = 0; a < 6; a++) wv[u][a] = w.proj_wj[6 * (size_t)kk + a];
        }
#pragma unroll
        for (int u = 0; u < 4; u++) {
            if (k + u >= k1) break;
#pragma unroll
            for (int a = 0; a < 6; a++) s = fma(wv[u][a], v[cj[u] + a], s);
        }
    }
    return s;
}

// The same in two parallel passes (used when the factors fit the staging buffer): t_k = w_jk^T v[c_jk ..] of EVERY factor, one
// thread each with coalesced index / weight loads, into shared memory; then one thread per landmark adds its range of t and the
// 13 shared columns.  All global loads of a pass are independent of each other: one round trip instead of one per factor.
__device__ __forceinline__ void factor_dots(const SolverScratch &w, const double *v, int n_proj, double *s_t) {
    for (int k = threadIdx.x; k < n_proj; k += blockDim.x) {
        const int cj = w.proj_colj[k];
        const double *wj = w.proj_wj + 6 * (size_t)k;
        double wv[6];
#pragma unroll
        for (int a = 0; a < 6; a++) wv[a] = wj[a];
        double s = 0;
#pragma unroll
        for (int a = 0; a < 6; a++) s = fma(wv[a], v[cj + a], s);
        s_t[k] = s;
    }
}
struct LandmarkHead {  // everything one landmark needs from global memory, loaded together
    int ci, k0, k1;
    double ws[13], c, g;
};
__device__ __forceinline__ LandmarkHead load_landmark(const DevProblem &P, const SolverScratch &w, int l) {
    LandmarkHead h;
    h.ci = w.lm_coli[l];
    h.k0 = P.lm_first[l];
    h.k1 = P.lm_first[l + 1];
#pragma unroll
    for (int a = 0; a < 13; a++) h.ws[a] = w.lm_ws[13 * (size_t)l + a];
    h.c = w.lm_c[l];
    h.g = w.lm_g[l];
    return h;
}
__device__ __forceinline__ double landmark_dot_staged(const LandmarkHead &h, const double *v, const double *s_t, int col_ex, int col_td) {
    double s = h.ws[12] * v[col_td];
#pragma unroll
    for (int a = 0; a < 6; a++) s = fma(h.ws[a], v[h.ci + a], fma(h.ws[6 + a], v[col_ex + a], s));
    double s1 = 0;
    for (int k = h.k0; k + 1 < h.k1; k += 2) {
        s += s_t[k];
        s1 += s_t[k + 1];
    }
    if ((h.k1 - h.k0) & 1) s += s_t[h.k1 - 1];
    return s + s1;
}

// y_l = (g_l - w_l^T y_p) / (c_l + clamp(c_l) mu) for the landmarks l = first, first + stride, ... (one thread each);
// adds this thread's part of |gn|^2 and gradient_.gn.  yp: reduced solution (shared or global).
__device__ __forceinline__ void landmark_backsub(const DevProblem &P, const SolverScratch &w, const double *yp, double mu, int first,
                                                 int stride, double &gn2, double &gdot, int &bad) {
    const StateLayout lay = P.lay;
    const int D = lay.D(), L = lay.L, col_ex = lay.col_ex(), col_td = lay.col_td();
    for (int l = first; l < L; l += stride) {
        double yl = 0;
        if (w.lm_coli[l] >= 0) {
            const double c_l = w.lm_c[l];
            yl = (w.lm_g[l] - landmark_dot(P, w, yp, l, col_ex, col_td)) / (c_l + clamp_diag(c_l) * mu);
        }
        if (!isfinite(yl)) bad = 1;
        const double g = -w.diagonal[D + l] * yl;
        w.gn[D + l] = g;
        gn2 += g * g;
        gdot += g * w.gradient[D + l];
    }
}

// =====================================================================================================
// fused != 0 (L <= kFuseMaxL): the whole step in this one CTA — Cauchy point from the reduced system itself
//   |J_s a|^2 = a_p^T S a_p + sum_l [ (w_l^T a_p)^2 / c_mu,l + 2 a_l (w_l^T a_p) + c_l a_l^2 ]      (H_pp = S + sum_l w_l w_l^T / c_mu,l)
// then factor/solve, landmark back-substitution and cand_tail.  fused == 0: pre_kernel supplied the Cauchy scalars and
// gnlm_kernel finishes (back-substitution over many CTAs, cand_tail in its last CTA).
__global__ void __launch_bounds__(CHOL_THREADS) chol_kernel(const DevProblem *__restrict__ Pd, SolverScratch w, double *x0, double *x1,
                                                           SolverCtl *ctl, int use_smem, int fused) {
    const DevProblem &P = *Pd;
    pdl_prologue();
    if (ctl->done != 0) return;
    extern __shared__ double smem[];
    __shared__ double s_red[33];
    __shared__ __align__(16) double s_U[2 * NB * NB];
    __shared__ __align__(16) double s_row[2 * NB];
    __shared__ int s_okv[kMaxFrames * 15 / NB + 4];
    const StateLayout lay = P.lay;
    const lfvio_solver_opts &o = P.o;
    const int D = lay.D(), L = lay.L, N = D + L;
    const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31, wid = tid >> 5, nwarp = nt >> 5;
    const int nblk = (D + NB - 1) / NB;
    const int PW = chol_panel_pitch(D);
    double *s_panel = smem;                        // NB * PW (16-byte aligned rows)
    double *s_y = s_panel + (size_t)NB * PW;       // D (padded to even)
    double *s_invd = s_y + ((D + 1) & ~1);         // nblk * NB reciprocal pivots
    double *s_a = s_invd + (size_t)nblk * NB;      // D (padded to even): Cauchy direction, reduced columns
    double *s_PL = s_a + ((D + 1) & ~1);           // nblk x NB x NB: U11^-T of every diagonal block
    double *s_dd = s_PL + (size_t)nblk * NB * NB;  // D (padded to even): diagonal_^2, the weights of the LM term
    double *s_delta = s_dd + ((D + 1) & ~1);       // D + kFuseMaxL: the scaled step of cand_tail
    double *s_S = s_delta + ((D + kFuseMaxL + 1) & ~1);  // packed upper triangle when use_smem
    if (ctl->reuse != 0) {
        if (fused) cand_tail(P, w, x0, x1, ctl, 0.0, 0.0, s_red, s_delta);
        return;
    }
    const MatPacked Sp{s_S, D};
    if (use_smem) {
        // the reduced system on its way into shared memory while the Cauchy point is being formed: one asynchronous 8-byte
        // copy per element of the upper triangle (rows are contiguous both in global and in the packed layout), a row per warp
        for (int r = wid; r < D; r += nwarp) {
            const double *src = w.S + (size_t)r * D;
            const uint32_t drow = smem_u32(Sp.row(r));
            for (int c = r + lane; c < D; c += 32)
                asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(drow + 8u * (unsigned)c), "l"(src + c) : "memory");
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    }
    PHASE_MARK(0);
    const bool staged = fused && P.n_proj <= NB * PW;  // per-factor dot products staged in the (then idle) panel buffer
    const double mu = ctl->mu;
    double gg, jj_part = 0, gmax;
    if (fused) {
        // ---- diagonal_, gradient_ (DoglegStrategy::ComputeStep) and the Cauchy direction a = gradient_ / diagonal_ ----
        double gga = 0, gm = 0;
        for (int c = tid; c < N; c += nt) {
            const double d2 = clamp_diag(c < D ? w.diag2[c] : w.lm_c[c - D]);
            const double g = c < D ? w.grad[c] : w.lm_g[c - D];
            const double dg = sqrt(d2);
            w.diagonal[c] = dg;
            const double gs = g / dg;
            w.gradient[c] = gs;
            gga += gs * gs;
            gm = fmax(gm, fabs(g / w.scale[c]));  // un-scaled gradient J^T r
            if (c < D) {
                s_a[c] = g / d2;
                s_dd[c] = dg * dg;
                if (use_smem) {  // consumed: clear for the next accumulation (see launch_accumulate)
                    w.diag2[c] = 0.0;
                    w.grad[c] = 0.0;
                }
            }
        }
        PHASE_MARK(7);
        gg = block_sum(gga, s_red);
        gmax = block_max(gm, s_red);
        PHASE_MARK(8);
        // landmark part of |J_s a|^2 (one thread per landmark)
        if (staged) {
            factor_dots(w, s_a, P.n_proj, s_panel);  // (the panel is not in use yet: it is cleared after this phase)
            __syncthreads();
            for (int l = tid; l < L; l += nt) {
                const LandmarkHead h = load_landmark(P, w, l);
                if (h.ci < 0) continue;
                const double t = landmark_dot_staged(h, s_a, s_panel, lay.col_ex(), lay.col_td());
                const double cl = clamp_diag(h.c);
                const double a_l = h.g / cl;
                jj_part += t * t / (h.c + cl * mu) + 2.0 * a_l * t + h.c * a_l * a_l;
            }
            __syncthreads();  // all reads of the staging buffer before the panel is cleared
        } else {
            for (int l = tid; l < L; l += nt) {
                if (w.lm_coli[l] < 0) continue;
                const double t = landmark_dot(P, w, s_a, l, lay.col_ex(), lay.col_td());
                const double c_l = w.lm_c[l], cl = clamp_diag(c_l);
                const double a_l = w.lm_g[l] / cl;
                jj_part += t * t / (c_l + cl * mu) + 2.0 * a_l * t + c_l * a_l * a_l;
            }
        }
    } else {
        gg = ctl->acc[0];
        gmax = __longlong_as_double((long long)*reinterpret_cast<unsigned long long *>(&ctl->acc[2]));
    }
    PHASE_MARK(9);
    // ---- x_norm over the non-constant parameter blocks (global sizes) ----
    {
        const double *x = ctl->cur ? x1 : x0;
        double a = 0;
        for (int q = tid; q < lay.size(); q += nt) {
            const bool is_ex = q >= lay.ex() && q < lay.ex() + 7, is_td = q == lay.td();
            if ((is_ex && !o.estimate_extrinsic) || (is_td && !o.estimate_td)) continue;
            a += x[q] * x[q];
        }
        const double t = block_sum(a, s_red);
        if (tid == 0) ctl->x_norm = sqrt(t);
    }
    bool pd;
    PHASE_MARK(1);
    for (int q = tid; q < NB * PW; q += nt) s_panel[q] = 0.0;
    const bool clear_sys = fused && use_smem;
    for (int c = tid; c < D; c += nt) {
        s_y[c] = w.rhs[c];
        if (clear_sys) w.rhs[c] = 0.0;
        if (!fused) s_dd[c] = w.diagonal[c] * w.diagonal[c];  // (pre_kernel formed diagonal_)
    }
    PHASE_MARK(10);
    if (use_smem) {
        // the copy has landed: a_p^T S a_p (when fused) and the LM term from shared memory; the global copy is cleared for the next
        // accumulation (only now: an asynchronous copy is not ordered against a later store of the same thread)
        asm volatile("cp.async.wait_all;" ::: "memory");
        __syncthreads();
        for (int r = wid; r < D; r += nwarp) {
            double *drow = Sp.row(r);
            double *src = w.S + (size_t)r * D;
            const double ar = fused ? s_a[r] : 0.0;
            for (int c = r + lane; c < D; c += 32) {
                const double v = drow[c];
                if (fused) jj_part += ar * v * s_a[c] * (c == r ? 1.0 : 2.0);
                if (c == r) drow[c] = v + s_dd[c] * mu;
                if (clear_sys) src[c] = 0.0;
            }
        }
    } else {
        // regularise in place and, when fused, a_p^T S a_p on the way: coalesced upper triangle, four rows per warp and pass
        // with all their loads issued before the first use
        constexpr int RG = 4;
        for (int rb = wid * RG; rb < D; rb += nwarp * RG) {
            for (int cb = lane; rb + cb < D; cb += 32 * 6) {
                double v[RG][6];
#pragma unroll
                for (int rr = 0; rr < RG; rr++) {
                    const int r = rb + rr;
#pragma unroll
                    for (int u = 0; u < 6; u++) v[rr][u] = (r < D && r + cb + 32 * u < D) ? w.S[(size_t)r * D + r + cb + 32 * u] : 0.0;
                }
#pragma unroll
                for (int rr = 0; rr < RG; rr++) {
                    const int r = rb + rr;
                    if (r >= D) break;
                    double *src = w.S + (size_t)r * D;
                    const double ar = fused ? s_a[r] : 0.0;
#pragma unroll
                    for (int u = 0; u < 6; u++) {
                        const int c = r + cb + 32 * u;
                        if (c >= D) continue;
                        if (fused) jj_part += ar * v[rr][u] * s_a[c] * (c == r ? 1.0 : 2.0);
                        if (r == c) src[c] = v[rr][u] + s_dd[c] * mu;
                    }
                }
            }
        }
    }
    PHASE_MARK(11);
    const double jj = fused ? block_sum(jj_part, s_red) : ctl->acc[1];  // (block_sum synchronises the copy)
    PHASE_MARK(12);
    if (tid == 0) {
        ctl->alpha = gg / jj;
        ctl->gnorm = sqrt(gg);
        ctl->jj = jj;
        ctl->mu_solve = mu;
        ctl->acc[0] = ctl->acc[1] = ctl->acc[2] = 0.0;
        ctl->acc[3] = ctl->acc[4] = 0.0;
        ctl->pad = 1;  // chol_ok
        if (!o.fixed_iterations && gmax <= o.gradient_tolerance) ctl->done = 1;
    }
    __syncthreads();
    if (ctl->done != 0) return;
    PHASE_MARK(2);
    if (use_smem) pd = cholesky_lookahead(Sp, D, s_panel, PW, s_U, s_PL, s_invd, s_y, s_okv, s_row, WindowProfile{lay.F}, false, true);
    else pd = cholesky_lookahead(MatFull{w.S, D}, D, s_panel, PW, s_U, s_PL, s_invd, s_y, s_okv, s_row, WindowProfile{lay.F}, false, true);
    if (!pd) {
        if (tid == 0) ctl->pad = 0;
        __syncthreads();
        if (fused) cand_tail(P, w, x0, x1, ctl, 0.0, 0.0, s_red, s_delta);
        return;
    }
    PHASE_MARK(3);
    if (use_smem) solve_upper_back_la(Sp, D, s_y, s_PL);
    else solve_upper_back_la(MatFull{w.S, D}, D, s_y, s_PL);
    PHASE_MARK(4);
    double gn2 = 0, gdot = 0;
    int bad = 0;
    for (int c = tid; c < D; c += nt) {
        const double yv = s_y[c];
        if (!isfinite(yv)) bad = 1;
        w.y[c] = yv;
        const double g = -w.diagonal[c] * yv;  // scaled GN step: D * (-(J^T J)^-1 g)
        w.gn[c] = g;
        gn2 += g * g;
        gdot += g * w.gradient[c];
    }
    PHASE_MARK(13);
    if (staged) {
        factor_dots(w, s_y, P.n_proj, s_panel);  // (the factorisation is done with the panel)
        __syncthreads();
        for (int l = tid; l < L; l += nt) {
            const LandmarkHead h = load_landmark(P, w, l);
            const double dl = w.diagonal[D + l], gl = w.gradient[D + l];
            double yl = 0;
            if (h.ci >= 0) yl = (h.g - landmark_dot_staged(h, s_y, s_panel, lay.col_ex(), lay.col_td())) / (h.c + clamp_diag(h.c) * mu);
            if (!isfinite(yl)) bad = 1;
            const double g = -dl * yl;
            w.gn[D + l] = g;
            gn2 += g * g;
            gdot += g * gl;
        }
    } else if (fused) {
        landmark_backsub(P, w, s_y, mu, tid, nt, gn2, gdot, bad);
    }
    PHASE_MARK(14);
    gn2 = block_sum(gn2, s_red);
    gdot = block_sum(gdot, s_red);
    bad = __syncthreads_or(bad);
    if (tid == 0) {
        ctl->acc[3] = fused ? 0.0 : gn2;
        ctl->acc[4] = fused ? 0.0 : gdot;
        if (bad) ctl->pad = 0;
    }
    __syncthreads();
    PHASE_MARK(5);
    if (fused) cand_tail(P, w, x0, x1, ctl, gn2, gdot, s_red, s_delta);
    PHASE_MARK(6);
}

// =====================================================================================================
// Large windows only (L > kFuseMaxL): landmark back-substitution over many CTAs; the last CTA to finish runs cand_tail.
__global__ void __launch_bounds__(GNLM_WARPS * 32) gnlm_kernel(const DevProblem *__restrict__ Pd, SolverScratch w, double *x0, double *x1, SolverCtl *ctl) {
    const DevProblem &P = *Pd;
    pdl_prologue();
    if (ctl->done != 0) return;
    __shared__ double s_red[33];
    __shared__ int s_last;
    if (ctl->reuse != 0) {
        if (blockIdx.x == 0) cand_tail(P, w, x0, x1, ctl, 0.0, 0.0, s_red, nullptr);
        return;
    }
    double gn2 = 0, gdot = 0;
    int bad = 0;
    if (ctl->pad != 0) landmark_backsub(P, w, w.y, ctl->mu_solve, blockIdx.x * blockDim.x + threadIdx.x, gridDim.x * blockDim.x, gn2, gdot, bad);
    gn2 = block_sum(gn2, s_red);
    gdot = block_sum(gdot, s_red);
    bad = __syncthreads_or(bad);
    if (threadIdx.x == 0) {
        if (bad) ctl->pad = 0;
        if (gn2 != 0.0) atomicAdd(&ctl->acc[3], gn2);
        if (gdot != 0.0) atomicAdd(&ctl->acc[4], gdot);
        __threadfence();
        const unsigned t = atomicAdd(&ctl->ticket2, 1u);
        s_last = (t == gridDim.x - 1);
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    const double gn2_all = atomicAdd(&ctl->acc[3], 0.0), gdot_all = atomicAdd(&ctl->acc[4], 0.0);
    __syncthreads();
    cand_tail(P, w, x0, x1, ctl, gn2_all, gdot_all, s_red, nullptr);
    if (threadIdx.x == 0) {
        ctl->acc[3] = ctl->acc[4] = 0.0;
        ctl->ticket2 = 0;
    }
}

}  // namespace

#ifdef LFVIO_CHOL_PROFILE
void chol_profile_read(long long *out) { cudaMemcpyFromSymbol(out, g_sub, sizeof(long long) * 16); }
#endif

static size_t chol_base_bytes(int D) {
    const int F = (D - 7) / 15;
    (void)F;
    return sizeof(double) * ((size_t)NB * chol_panel_pitch(D) + 3 * (size_t)((D + 1) & ~1) + (size_t)cdiv(D, NB) * (NB + NB * NB) +
                             (size_t)((D + kFuseMaxL + 1) & ~1));
}
static bool chol_in_smem(int D) { return chol_base_bytes(D) + sizeof(double) * ((size_t)D * (D + 1) / 2) <= 218 * 1024; }

// fused / self-clearing is decided on the landmark BUCKET (g_L), so that it does not flip from frame to frame
bool step_zeroes_system(const DevProblem &P) { return P.g_L <= kFuseMaxL && chol_in_smem(P.lay.D()); }

int launch_step(const DevProblem &P, const EvalBuf &e0, const EvalBuf &e1, const SolverScratch &w, double *x0, double *x1,
                SolverCtl *ctl, cudaStream_t s) {
    const int D = P.lay.D(), L = P.g_L;
    const int fused = (L <= kFuseMaxL) ? 1 : 0;
    if (!fused) {
        const int work = std::max(std::max(D + L, P.g_nbp * 128), 1);
        LFVIO_CUDA_TRY(launch_pdl(pre_kernel, dim3(std::max(1, std::min(cdiv(work, PRE_THREADS), 148))), dim3(PRE_THREADS), 0, s, P.self_dev, e0, e1, w, ctl));
        LFVIO_LAUNCH_CHECK();
    }
    LFVIO_REQUIRE(D + 1 <= CHOL_THREADS - 128, "window too large for the single-pass panel solve (15 F + 8 <= 384)");
    const int nblk = cdiv(D, NB);
    const size_t base = chol_base_bytes(D);
    (void)nblk;
    const size_t packed = sizeof(double) * ((size_t)D * (D + 1) / 2);
    const size_t limit = 218 * 1024;
    const int use_smem = (base + packed <= limit) ? 1 : 0;
    LFVIO_TRY(ensure_dyn_smem(reinterpret_cast<const void *>(chol_kernel), (int)limit));
    LFVIO_CUDA_TRY(launch_pdl(chol_kernel, dim3(1), dim3(CHOL_THREADS), base + (use_smem ? packed : 0), s, P.self_dev, w, x0, x1, ctl, use_smem, fused));
    LFVIO_LAUNCH_CHECK();
    if (!fused) {
        LFVIO_CUDA_TRY(launch_pdl(gnlm_kernel, dim3(std::max(1, std::min(cdiv(L, GNLM_WARPS * 32), 148))), dim3(GNLM_WARPS * 32), 0, s, P.self_dev, w, x0, x1, ctl));
        LFVIO_LAUNCH_CHECK();
    }
    return LFVIO_OK;
}

}  // namespace lfvio
