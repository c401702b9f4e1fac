// K12/K13 — one trust-region step: Cauchy point, dense Cholesky solve of the reduced system, landmark back-substitution,
// traditional dogleg, model cost change and candidate state, as four launches with device-side control flow.
//
// Reference: ceres::Solve at vins_estimator/src/estimator.cpp:810-825 (DENSE_SCHUR + DOGLEG); Ceres 1.12.0
// DoglegStrategy::ComputeStep / ComputeCauchyPoint / ComputeGaussNewtonStep / ComputeTraditionalDoglegStep and the
// TrustRegionMinimizer step evaluation are restated in SURVEY.md Appendix B.4 (Ceres is third-party, not vendored).
//
//   pre_kernel   (multi-CTA)  diagonal_, gradient_ per column; |gradient_|^2, max |J^T r|, |J_s (gradient_/diagonal_)|^2
//                             over all residual rows, reduced with a few atomics.
//   chol_kernel  (one CTA)    alpha; reduced system copied into shared memory as a packed upper triangle (+ LM term);
//                             blocked (16) right-looking Cholesky: the 16x16 diagonal blocks are factored by ONE WARP IN
//                             REGISTERS (lane = column, pivots/rows exchanged with warp shuffles) and inverted, so the row
//                             panel and the triangular solves are dependency-free small mat-muls; 4x4 register tiles in
//                             the trailing update; forward/backward substitution blocked the same way.
//   gnlm_kernel  (multi-CTA)  y_l = (g_l - w_l^T y_p) / c_l, one warp per landmark, + |gn|^2 and gradient_.gn partial sums.
//   cand_kernel  (multi-CTA)  dogleg coefficients from the reduced scalars, step on the fly per column, model cost
//                             change over all rows, x_cand = Plus(x, step * scale); the last CTA publishes the step to
//                             the control block (validity, parameter-tolerance test, eval_pending).
#include "backend.cuh"
#include "solver_common.cuh"
#include "chol16.cuh"

namespace lfvio {

namespace {

constexpr int CHOL_THREADS = 512;
constexpr int PRE_THREADS = 256;
constexpr int CAND_THREADS = 256;
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
__global__ void __launch_bounds__(PRE_THREADS) pre_kernel(DevProblem P, EvalBuf e0, EvalBuf e1, SolverScratch w, SolverCtl *ctl) {
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
struct MatFull {
    double *p;
    int D;
    __device__ __forceinline__ double &at(int r, int c) const { return p[(size_t)r * D + c]; }
};
struct MatPacked {  // packed upper triangle, row-major (c >= r)
    double *p;
    int D;
    __device__ __forceinline__ double &at(int r, int c) const { return p[r * D - ((r * (r - 1)) >> 1) + (c - r)]; }
};

template <class M>
__device__ bool cholesky_upper(M S, int D, double *s_panel /* NB x D */, double *s_U /* NB x NB */, double *s_Xall, int *s_ok) {
    const int tid = threadIdx.x, nt = blockDim.x;
    if (tid == 0) *s_ok = 1;
    __syncthreads();
    for (int j0 = 0, blk = 0; j0 < D; j0 += NB, blk++) {
        const int nb = min(NB, D - j0);
        double *s_X = s_Xall + blk * NB * NB;
        // diagonal block -> shared (all threads), factor + invert (warp 0, registers), write back (all threads)
        for (int q = tid; q < NB * NB; q += nt) {
            const int r = q / NB, c = q % NB;
            s_U[q] = (r < nb && c < nb && c >= r) ? S.at(j0 + r, j0 + c) : ((r == c) ? 1.0 : 0.0);
        }
        __syncthreads();
        if (tid < 32) {
            const bool ok = factor_diag_block(s_U, s_X);
            if (!ok && tid == 0) *s_ok = 0;
        }
        __syncthreads();
        if (*s_ok == 0) return false;
        for (int q = tid; q < nb * nb; q += nt) {
            const int r = q / nb, c = q % nb;
            if (c >= r) S.at(j0 + r, j0 + c) = s_U[r * NB + c];
        }
        const int rest0 = j0 + nb, nrest = D - rest0;
        if (nrest > 0) {
            // row panel U12 = U11^-T A12 = X^T A12: dependency-free, one thread per column
            for (int c = tid; c < nrest; c += nt) {
                double a[NB], u[NB];
#pragma unroll
                for (int r = 0; r < NB; r++) a[r] = (r < nb) ? S.at(j0 + r, rest0 + c) : 0.0;
#pragma unroll
                for (int r = 0; r < NB; r++) {
                    double s = 0;
#pragma unroll
                    for (int t = 0; t < NB; t++)
                        if (t <= r) s += s_X[t * NB + r] * a[t];
                    u[r] = s;
                }
#pragma unroll
                for (int r = 0; r < NB; r++) {
                    if (r < nb) S.at(j0 + r, rest0 + c) = u[r];
                    s_panel[r * D + c] = (r < nb) ? u[r] : 0.0;
                }
            }
            if (tid < NB * 4) {  // zero padding so that the 4x4 tiles need no bounds checks inside the rank-16 loop
                const int r = tid >> 2, c = nrest + (tid & 3);
                if (c < D) s_panel[r * D + c] = 0.0;
            }
            __syncthreads();
            // trailing update A22[i][k] -= sum_t U12[t][i] U12[t][k] (k >= i), 4x4 register tiles
            const int nt4 = (nrest + 3) >> 2;
            for (int q = tid; q < nt4 * nt4; q += nt) {
                const int ti = q / nt4, tk = q - ti * nt4;
                if (tk < ti) continue;
                const int i0 = ti << 2, k0 = tk << 2;
                double acc[4][4];
#pragma unroll
                for (int u = 0; u < 4; u++)
#pragma unroll
                    for (int v = 0; v < 4; v++) acc[u][v] = 0.0;
#pragma unroll 4
                for (int t = 0; t < NB; t++) {
                    double a[4], b[4];
#pragma unroll
                    for (int u = 0; u < 4; u++) a[u] = s_panel[t * D + i0 + u];
#pragma unroll
                    for (int v = 0; v < 4; v++) b[v] = s_panel[t * D + k0 + v];
#pragma unroll
                    for (int u = 0; u < 4; u++)
#pragma unroll
                        for (int v = 0; v < 4; v++) acc[u][v] += a[u] * b[v];
                }
#pragma unroll
                for (int u = 0; u < 4; u++)
#pragma unroll
                    for (int v = 0; v < 4; v++) {
                        const int i = i0 + u, k = k0 + v;
                        if (i < nrest && k < nrest && k >= i) S.at(rest0 + i, rest0 + k) -= acc[u][v];
                    }
            }
        }
        __syncthreads();
    }
    return true;
}

// U^T z = b, U y = z on the shared vector y (in/out) using the inverted diagonal blocks: per block one 16x16 mat-vec
// by 16 threads plus a parallel update — no dependent scalar chain.  Whole CTA calls.
template <class M>
__device__ void solve_upper_blocked(M S, int D, double *y, const double *s_Xall, double *s_tmp /* NB */) {
    const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31, wid = tid >> 5;
    const int nblk = (D + NB - 1) / NB;
    for (int blk = 0; blk < nblk; blk++) {
        const int j0 = blk * NB, nb = min(NB, D - j0);
        const double *X = s_Xall + blk * NB * NB;
        double z = 0;
        if (tid < nb) {  // z[r] = sum_{t <= r} X[t][r] b[t]
            for (int t = 0; t <= tid; t++) z += X[t * NB + tid] * y[j0 + t];
        }
        __syncthreads();
        if (tid < nb) y[j0 + tid] = z;
        __syncthreads();
        for (int k = j0 + nb + tid; k < D; k += nt) {
            double s = 0;
#pragma unroll 4
            for (int t = 0; t < nb; t++) s += S.at(j0 + t, k) * y[j0 + t];
            y[k] -= s;
        }
        __syncthreads();
    }
    for (int blk = nblk - 1; blk >= 0; blk--) {
        const int j0 = blk * NB, nb = min(NB, D - j0), k0 = j0 + nb;
        const double *X = s_Xall + blk * NB * NB;
        for (int r = wid; r < nb; r += (nt >> 5)) {  // s_tmp[r] = sum_{k > block} U[j0+r][k] y[k]
            double s = 0;
            for (int k = k0 + lane; k < D; k += 32) s += S.at(j0 + r, k) * y[k];
            s = warp_sum(s);
            if (lane == 0) s_tmp[r] = s;
        }
        __syncthreads();
        double v = 0;
        if (tid < nb) {  // y[r] = sum_{t >= r} X[r][t] (z[t] - s_tmp[t])
            for (int t = tid; t < nb; t++) v += X[tid * NB + t] * (y[j0 + t] - s_tmp[t]);
        }
        __syncthreads();
        if (tid < nb) y[j0 + tid] = v;
        __syncthreads();
    }
}

__global__ void __launch_bounds__(CHOL_THREADS) chol_kernel(DevProblem P, SolverScratch w, const double *x0, const double *x1,
                                                           SolverCtl *ctl, int use_smem) {
    if (ctl->done != 0 || ctl->reuse != 0) return;
    extern __shared__ double smem[];
    __shared__ double s_red[33];
    __shared__ double s_U[NB * NB];
    __shared__ double s_tmp[NB];
    __shared__ int s_ok;
    const StateLayout lay = P.lay;
    const lfvio_solver_opts &o = P.o;
    const int D = lay.D();
    const int tid = threadIdx.x, nt = blockDim.x;
    const int nblk = (D + NB - 1) / NB;
    double *s_panel = smem;                        // NB * D
    double *s_y = s_panel + (size_t)NB * D;        // D
    double *s_X = s_y + D;                         // nblk * NB * NB inverted diagonal blocks
    double *s_S = s_X + (size_t)nblk * NB * NB;    // packed upper triangle when use_smem
    PHASE_MARK(0);
    // ---- scalars of the Cauchy point; gradient tolerance ----
    if (tid == 0) {
        const double gg = ctl->acc[0], jj = ctl->acc[1];
        const double gmax = __longlong_as_double((long long)*reinterpret_cast<unsigned long long *>(&ctl->acc[2]));
        ctl->alpha = gg / jj;
        ctl->gnorm = sqrt(gg);
        ctl->acc[0] = ctl->acc[1] = ctl->acc[2] = 0.0;
        ctl->acc[3] = ctl->acc[4] = 0.0;
        ctl->pad = 1;  // chol_ok
        if (!o.fixed_iterations && gmax <= o.gradient_tolerance) ctl->done = 1;
    }
    __syncthreads();
    if (ctl->done != 0) return;
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
    const double mu = ctl->mu;
    bool pd;
    PHASE_MARK(1);
    if (use_smem) {
        const MatPacked Sp{s_S, D};
        for (int q = tid; q < D * D; q += nt) {
            const int r = q / D, c = q - r * D;
            if (c >= r) Sp.at(r, c) = w.S[q] + (r == c ? w.diagonal[c] * w.diagonal[c] * mu : 0.0);
        }
        __syncthreads();
        PHASE_MARK(2);
        pd = cholesky_upper(Sp, D, s_panel, s_U, s_X, &s_ok);
    } else {
        for (int c = tid; c < D; c += nt) w.S[(size_t)c * D + c] += w.diagonal[c] * w.diagonal[c] * mu;
        __syncthreads();
        pd = cholesky_upper(MatFull{w.S, D}, D, s_panel, s_U, s_X, &s_ok);
    }
    if (!pd) {
        if (tid == 0) ctl->pad = 0;
        return;
    }
    PHASE_MARK(3);
    for (int c = tid; c < D; c += nt) s_y[c] = w.rhs[c];
    __syncthreads();
    if (use_smem) solve_upper_blocked(MatPacked{s_S, D}, D, s_y, s_X, s_tmp);
    else solve_upper_blocked(MatFull{w.S, D}, D, s_y, s_X, s_tmp);
    __syncthreads();
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
    gn2 = block_sum(gn2, s_red);
    gdot = block_sum(gdot, s_red);
    bad = __syncthreads_or(bad);
    if (tid == 0) {
        ctl->acc[3] = gn2;
        ctl->acc[4] = gdot;
        if (bad) ctl->pad = 0;
    }
    PHASE_MARK(5);
}

// =====================================================================================================
__global__ void __launch_bounds__(GNLM_WARPS * 32) gnlm_kernel(DevProblem P, SolverScratch w, SolverCtl *ctl) {
    if (ctl->done != 0 || ctl->reuse != 0 || ctl->pad == 0) return;
    __shared__ double s_a[GNLM_WARPS], s_b[GNLM_WARPS];
    const StateLayout lay = P.lay;
    const int D = lay.D(), L = lay.L;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const double mu = ctl->mu;
    double gn2 = 0, gdot = 0;
    for (int l = blockIdx.x * GNLM_WARPS + wid; l < L; l += gridDim.x * GNLM_WARPS) {
        const int k0 = P.lm_first[l], K = P.lm_first[l + 1] - k0;
        double yl = 0;
        if (K > 0) {
            const int fi = P.proj[k0].frame_i;
            double s = 0;
            if (lane < 13) {
                const int col = lane < 6 ? lay.col_pose(fi) + lane : (lane < 12 ? lay.col_ex() + (lane - 6) : lay.col_td());
                s = w.lm_ws[13 * (size_t)l + lane] * w.y[col];
            }
            for (int q = lane; q < 6 * K; q += 32) {
                const int k = q / 6, a = q - 6 * k;
                s += w.proj_wj[6 * (size_t)k0 + q] * w.y[lay.col_pose(P.proj[k0 + k].frame_j) + a];
            }
            s = warp_sum(s);
            const double c_l = w.lm_c[l];
            yl = (w.lm_g[l] - s) / (c_l + clamp_diag(c_l) * mu);
        }
        if (lane == 0) {
            if (!isfinite(yl)) ctl->pad = 0;
            const double g = -w.diagonal[D + l] * yl;
            w.gn[D + l] = g;
            gn2 += g * g;
            gdot += g * w.gradient[D + l];
        }
    }
    if (lane == 0) {
        s_a[wid] = gn2;
        s_b[wid] = gdot;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        double a = 0, b = 0;
        for (int k = 0; k < GNLM_WARPS; k++) {
            a += s_a[k];
            b += s_b[k];
        }
        if (a != 0.0) atomicAdd(&ctl->acc[3], a);
        if (b != 0.0) atomicAdd(&ctl->acc[4], b);
    }
}

// =====================================================================================================
__global__ void __launch_bounds__(CAND_THREADS) cand_kernel(DevProblem P, EvalBuf e0, EvalBuf e1, SolverScratch w, double *x0,
                                                           double *x1, SolverCtl *ctl) {
    if (ctl->done != 0) return;
    __shared__ double s_red[33];
    __shared__ int pmap[15 * kMaxFrames + 8];
    __shared__ int s_last;
    const StateLayout lay = P.lay;
    const lfvio_solver_opts &o = P.o;
    const int D = lay.D(), L = lay.L;
    const int cur = ctl->cur, reuse = ctl->reuse;
    const EvalBuf &e = cur ? e1 : e0;
    const double *x = cur ? x1 : x0;
    double *xc = cur ? x0 : x1;
    const int ok = reuse ? 1 : ctl->pad;
    // ---- traditional dogleg from the reduced scalars (every thread, identical arithmetic) ----
    const double gnorm = ctl->gnorm, alpha = ctl->alpha, radius = ctl->radius;
    const double gn2 = reuse ? ctl->gn_norm * ctl->gn_norm : ctl->acc[3];
    const double gdot = reuse ? ctl->gdot : ctl->acc[4];
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
        const double a_sq = pow(alpha * gnorm, 2.0);
        const double bma_sq = a_sq - 2 * b_dot_a + pow(gnn, 2.0);
        const double cc = b_dot_a - a_sq;
        const double dd = sqrt(cc * cc + bma_sq * (pow(radius, 2.0) - a_sq));
        const double beta = (cc <= 0) ? (dd - cc) / bma_sq : (radius * radius - a_sq) / (dd + cc);
        cg = -alpha * (1.0 - beta);
        cn = beta;
        dsn = sqrt(fmax(cg * cg * gnorm * gnorm + 2.0 * cg * cn * gdot + cn * cn * gn2, 0.0));
    }
    double model = 0, dx2 = 0;
    if (ok) {
        build_prior_map(P, pmap);
        __syncthreads();
        auto stepv = [&](int col) { return (cg * w.gradient[col] + cn * w.gn[col]) / w.diagonal[col]; };
        model = rows_partial(P, e, w.scale, stepv, 1, pmap);
        // ---- candidate x_plus_delta = Plus(x, step * scale) ----
        const int gtid = blockIdx.x * blockDim.x + threadIdx.x, nth = gridDim.x * blockDim.x;
        auto delta = [&](int col) { return stepv(col) * w.scale[col]; };
        for (int f = gtid; f < lay.F; f += nth) {
            double d[6];
#pragma unroll
            for (int a = 0; a < 6; a++) d[a] = delta(lay.col_pose(f) + a);
            pose_plus(x + lay.pose(f), d, xc + lay.pose(f));
            for (int a = 0; a < 7; a++) {
                const double df = x[lay.pose(f) + a] - xc[lay.pose(f) + a];
                dx2 += df * df;
            }
            for (int a = 0; a < 9; a++) {
                const double dv = delta(lay.col_sb(f) + a);
                const double nv = x[lay.sb(f) + a] + dv;
                xc[lay.sb(f) + a] = nv;
                const double df = x[lay.sb(f) + a] - nv;
                dx2 += df * df;
            }
        }
        if (gtid == nth - 1) {  // extrinsic + td, by the last thread of the grid
            if (o.estimate_extrinsic) {
                double d[6];
#pragma unroll
                for (int a = 0; a < 6; a++) d[a] = delta(lay.col_ex() + a);
                pose_plus(x + lay.ex(), d, xc + lay.ex());
                for (int a = 0; a < 7; a++) {
                    const double df = x[lay.ex() + a] - xc[lay.ex() + a];
                    dx2 += df * df;
                }
            } else {
                for (int a = 0; a < 7; a++) xc[lay.ex() + a] = x[lay.ex() + a];
            }
            if (o.estimate_td) {
                const double nv = x[lay.td()] + delta(lay.col_td());
                xc[lay.td()] = nv;
                dx2 += (x[lay.td()] - nv) * (x[lay.td()] - nv);
            } else {
                xc[lay.td()] = x[lay.td()];
            }
        }
        for (int l = gtid; l < L; l += nth) {
            const double nv = x[lay.feat(l)] + delta(D + l);
            xc[lay.feat(l)] = nv;
            dx2 += (x[lay.feat(l)] - nv) * (x[lay.feat(l)] - nv);
        }
    }
    model = block_sum(model, s_red);
    dx2 = block_sum(dx2, s_red);
    if (threadIdx.x == 0) {
        if (model != 0.0) atomicAdd(&ctl->acc[5], model);
        if (dx2 != 0.0) atomicAdd(&ctl->acc[6], dx2);
        __threadfence();
        const unsigned t = atomicAdd(&ctl->ticket2, 1u);
        s_last = (t == gridDim.x - 1);
    }
    __syncthreads();
    if (s_last && threadIdx.x == 0) {
        __threadfence();
        const double model_cost_change = atomicAdd(&ctl->acc[5], 0.0);
        const double step_norm = sqrt(atomicAdd(&ctl->acc[6], 0.0));
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
        ctl->acc[3] = ctl->acc[4] = ctl->acc[5] = ctl->acc[6] = 0.0;
        ctl->ticket2 = 0;
    }
}

}  // namespace

int launch_step(const DevProblem &P, const EvalBuf &e0, const EvalBuf &e1, const SolverScratch &w, double *x0, double *x1,
                SolverCtl *ctl, cudaStream_t s) {
    const int D = P.lay.D(), L = P.lay.L;
    const int work = std::max(std::max(D + L, P.n_proj), 1);
    const int grid_rows = std::max(1, std::min(cdiv(work, PRE_THREADS), 148));
    pre_kernel<<<grid_rows, PRE_THREADS, 0, s>>>(P, e0, e1, w, ctl);
    LFVIO_LAUNCH_CHECK();
    const int nblk = cdiv(D, NB);
    const size_t base = sizeof(double) * ((size_t)NB * D + D + (size_t)nblk * NB * NB);
    const size_t packed = sizeof(double) * ((size_t)D * (D + 1) / 2);
    const size_t limit = 220 * 1024;
    const int use_smem = (base + packed <= limit) ? 1 : 0;
    static bool attr_set = false;
    if (!attr_set) {
        LFVIO_CUDA_TRY(cudaFuncSetAttribute(chol_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)limit));
        attr_set = true;
    }
    chol_kernel<<<1, CHOL_THREADS, base + (use_smem ? packed : 0), s>>>(P, w, x0, x1, ctl, use_smem);
    LFVIO_LAUNCH_CHECK();
    if (L > 0) {
        gnlm_kernel<<<std::max(1, std::min(cdiv(L, GNLM_WARPS), 148)), GNLM_WARPS * 32, 0, s>>>(P, w, ctl);
        LFVIO_LAUNCH_CHECK();
    }
    cand_kernel<<<std::max(1, std::min(cdiv(work, CAND_THREADS), 148)), CAND_THREADS, 0, s>>>(P, e0, e1, w, x0, x1, ctl);
    LFVIO_LAUNCH_CHECK();
    return LFVIO_OK;
}

}  // namespace lfvio
