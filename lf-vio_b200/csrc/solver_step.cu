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
#ifdef LFVIO_CHOL_PROFILE
#define SUB_T0() long long sub_t = clock64(); long long sub_acc[8] = {0, 0, 0, 0, 0, 0, 0, 0}
#define SUB_MARK(i) do { const long long now = clock64(); sub_acc[i] += now - sub_t; sub_t = now; } while (0)
#define SUB_FLUSH() do { if (threadIdx.x == 0 || threadIdx.x == 32) for (int i = 0; i < 8; i++) g_sub[(threadIdx.x >> 5) * 8 + i] = sub_acc[i]; } while (0)
__device__ long long g_sub[16];
#else
#define SUB_T0()
#define SUB_MARK(i)
#define SUB_FLUSH()
#endif

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
// row pitch of the shared panel: trailing columns rounded up to the 4-wide tiles, + one tile whose last column holds the rhs
__host__ __device__ inline int chol_panel_pitch(int D) { return ((D + 3) & ~3) + 4; }

struct MatFull {
    double *p;
    int D;
    __device__ __forceinline__ double &at(int r, int c) const { return p[(size_t)r * D + c]; }
    __device__ __forceinline__ double *row(int r) const { return p + (size_t)r * D; }  // row(r)[c] == at(r, c)
};
struct MatPacked {  // packed upper triangle, row-major (c >= r)
    double *p;
    int D;
    __device__ __forceinline__ double &at(int r, int c) const { return p[r * D - ((r * (r - 1)) >> 1) + (c - r)]; }
    __device__ __forceinline__ double *row(int r) const { return p + (r * D - ((r * (r - 1)) >> 1) - r); }  // valid for c >= r
};

// ---- look-ahead blocked Cholesky + fused forward substitution ---------------------------------------------------
// S = U^T U on the upper triangle, 16-column block steps, one CTA.  Warp 0 is the "diagonal warp": while the other
// warps (workers) apply block step k to the trailing matrix, warp 0 already solves the 16 panel columns of block k+1,
// applies step k to diagonal block k+1 and factors it in registers — the serial pivot chain of block k+1 overlaps the
// throughput-bound trailing update of block k.  The right-hand side rides along as one extra panel column, so the
// forward substitution U^T z = b costs nothing extra; solve_upper_back finishes with U y = z.
// Barriers: id 1 = "diagonal block ready / trailing update done" (all threads), id 2 = "panel complete"
// (workers wait, warp 0 only arrives).
__device__ __forceinline__ void bar_sync(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }
__device__ __forceinline__ void bar_arrive(int id, int n) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(n) : "memory"); }

// one panel column: a <- U11^-T a by substitution with the factored diagonal block (broadcast reads of U, 1/pivot)
__device__ __forceinline__ void panel_column_solve(double (&a)[NB], const double *U, const double *invd) {
#pragma unroll
    for (int j = 0; j < NB; j++) {
        const double u = a[j] * invd[j];
        a[j] = u;
#pragma unroll
        for (int r = j + 1; r < NB; r++) a[r] = fma(-U[j * NB + r], u, a[r]);
    }
}

template <class M>
__device__ bool cholesky_lookahead(M S, int D, double *s_panel /* NB x PW */, int PW, double *s_U /* 2 x NB x NB */,
                                   double *s_invd /* nblk x NB */, double *s_y /* D: b in, z out */, int *s_okv /* nblk */, double *s_row /* 2 x NB */) {
    const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31, wid = tid >> 5;
    const int nblk = (D + NB - 1) / NB;
    const int nworkers = nt - 32, wtid = tid - 32;
    const int nbar = nt;  // threads taking part in the named barriers
    const int zc = PW - 1;           // panel column of the right-hand side
    SUB_T0();
    if (wid == 0) {  // factor diagonal block 0
        const int nb = min(NB, D);
        for (int q = lane; q < NB * NB; q += 32) {
            const int r = q / NB, c = q % NB;
            s_U[q] = (r < nb && c < nb && c >= r) ? S.at(r, c) : ((r == c) ? 1.0 : 0.0);
        }
        __syncwarp();
        const bool ok = factor_diag_block_split(s_U, s_invd, s_row);
        if (lane == 0) s_okv[0] = ok ? 1 : 0;
        for (int q = lane; q < NB * NB; q += 32) {
            const int r = q / NB, c = q % NB;
            if (c >= r && c < nb) S.at(r, c) = s_U[q];
        }
    }
    bar_sync(1, nbar);
#ifdef LFVIO_CHOL_PROFILE
    sub_t = clock64();
#endif
    for (int blk = 0; blk < nblk; blk++) {
        if (s_okv[blk] == 0) return false;
        const int j0 = blk * NB, nb = min(NB, D - j0);
        const int rest0 = j0 + nb, nrest = D - rest0, nb1 = min(NB, nrest);
        const double *U = s_U + (blk & 1) * NB * NB, *invd = s_invd + blk * NB;
        if (wid == 0) {
            // ---- diagonal warp: panel columns of block blk+1, then look ahead ----
            double *Un = s_U + ((blk + 1) & 1) * NB * NB;
            if (lane < nb1) {
                double a[NB];
#pragma unroll
                for (int r = 0; r < NB; r++) a[r] = (r < nb) ? S.at(j0 + r, rest0 + lane) : 0.0;
                panel_column_solve(a, U, invd);
#pragma unroll
                for (int r = 0; r < NB; r++) {
                    if (r < nb) S.at(j0 + r, rest0 + lane) = a[r];
                    s_panel[r * PW + lane] = a[r];
                }
            }
            __threadfence_block();
            bar_arrive(2, nbar);
            SUB_MARK(0);
            if (nb1 > 0) {
                __syncwarp();
                // step blk applied to diagonal block blk+1: lane (h, c) owns rows 8h .. 8h+7 of column c
                const int c = lane & 15, i0 = (lane >> 4) * 8;
                double acc[8];
#pragma unroll
                for (int u = 0; u < 8; u++) acc[u] = 0.0;
                if (c < nb1) {
#pragma unroll 4
                    for (int t = 0; t < NB; t++) {
                        const double pc = s_panel[t * PW + c];
#pragma unroll
                        for (int u = 0; u < 8; u++) acc[u] = fma(s_panel[t * PW + i0 + u], pc, acc[u]);
                    }
                }
#pragma unroll
                for (int u = 0; u < 8; u++) {
                    const int i = i0 + u;
                    Un[i * NB + c] = (i <= c && c < nb1) ? S.at(rest0 + i, rest0 + c) - acc[u] : ((i == c) ? 1.0 : 0.0);
                }
                __syncwarp();
                SUB_MARK(1);
                const bool ok = factor_diag_block_split(Un, s_invd + (blk + 1) * NB, s_row);
                if (lane == 0) s_okv[blk + 1] = ok ? 1 : 0;
                for (int q = lane; q < NB * NB; q += 32) {
                    const int r = q / NB, cc = q % NB;
                    if (cc >= r && cc < nb1) S.at(rest0 + r, rest0 + cc) = Un[q];
                }
                SUB_MARK(2);
            }
        } else {
            // ---- workers: remaining panel columns + right-hand side ----
            const int ncol = nrest - nb1 + 1;  // last one is the right-hand side
            // (one pass: D < number of workers, checked by launch_step)
            if (const int q = wtid; q < ncol) {
                const bool is_rhs = q == ncol - 1;
                const int c = nb1 + q;  // column inside the trailing matrix
                double a[NB];
#pragma unroll
                for (int r = 0; r < NB; r++) a[r] = (r < nb) ? (is_rhs ? s_y[j0 + r] : S.at(j0 + r, rest0 + c)) : 0.0;
                SUB_MARK(4);
                panel_column_solve(a, U, invd);
                SUB_MARK(5);
                const int pc = is_rhs ? zc : c;
#pragma unroll
                for (int r = 0; r < NB; r++) {
                    if (r < nb) {
                        if (is_rhs) s_y[j0 + r] = a[r];
                        else S.at(j0 + r, rest0 + c) = a[r];
                    }
                    s_panel[r * PW + pc] = a[r];
                }
            }
            SUB_MARK(0);
            bar_sync(2, nbar);
            SUB_MARK(1);
            // ---- trailing update A22[i][k] -= sum_t U12[t][i] U12[t][k] (k >= i) in 16 x 16 blocks: 16 threads per
            //      block, thread y owns column y of it (16 accumulators).  Per rank-1 step the 16 row factors are one
            //      broadcast read and the column factor one conflict-free read; the read-modify-write of S touches 16
            //      consecutive doubles per row.  Block (0, 0) — diagonal block blk+1 — belongs to warp 0. ----
            const int nbk = (nrest + NB - 1) / NB, nunits = nbk * (nbk + 1) / 2;
            const int y = wtid & 15;
            for (int q = 1 + (wtid >> 4); q < nunits; q += nworkers >> 4) {
                int bi = 0, rem = q;
                while (rem >= nbk - bi) {
                    rem -= nbk - bi;
                    bi++;
                }
                const int bk = bi + rem, i0 = bi * NB, k = bk * NB + y;
                double acc[NB];
#pragma unroll
                for (int u = 0; u < NB; u++) acc[u] = 0.0;
#pragma unroll 2
                for (int t = 0; t < NB; t++) {
                    const double pk = s_panel[t * PW + k];
                    const double2 *ap = reinterpret_cast<const double2 *>(&s_panel[t * PW + i0]);
#pragma unroll
                    for (int u2 = 0; u2 < NB / 2; u2++) {
                        const double2 av = ap[u2];
                        acc[2 * u2] = fma(av.x, pk, acc[2 * u2]);
                        acc[2 * u2 + 1] = fma(av.y, pk, acc[2 * u2 + 1]);
                    }
                }
                if (k < nrest) {
#pragma unroll
                    for (int u = 0; u < NB; u++) {
                        const int i = i0 + u;
                        if (i <= k) S.row(rest0 + i)[rest0 + k] -= acc[u];  // (i <= k < nrest)
                    }
                }
            }
            // right-hand side: b_rest -= U12^T z (taken by the workers that got the fewest tiles)
            for (int i = nworkers - 1 - wtid; i < nrest; i += nworkers) {
                double sacc = 0;
#pragma unroll 4
                for (int t = 0; t < NB; t++) sacc = fma(s_panel[t * PW + i], s_panel[t * PW + zc], sacc);
                s_y[rest0 + i] -= sacc;
            }
            SUB_MARK(2);
        }
        bar_sync(1, nbar);
        SUB_MARK(3);
    }
    SUB_FLUSH();
    return true;
}

// U y = z on the shared vector y (in/out), last block first: warp 0 back-substitutes the 16 x 16 diagonal block with
// shuffles, then every thread folds the 16 new unknowns into the rows above.  Whole CTA calls.
template <class M>
__device__ void solve_upper_back(M S, int D, double *y, const double *s_invd) {
    const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31;
    const int nblk = (D + NB - 1) / NB;
    for (int blk = nblk - 1; blk >= 0; blk--) {
        const int j0 = blk * NB, nb = min(NB, D - j0);
        if (tid < 32) {
            const int r = lane & 15;
            double urow[NB];
#pragma unroll
            for (int t = 0; t < NB; t++) urow[t] = (r < nb && t > r && t < nb) ? S.at(j0 + r, j0 + t) : 0.0;
            double wv = (r < nb) ? y[j0 + r] : 0.0;
            const double inv = (r < nb) ? s_invd[blk * NB + r] : 1.0;
#pragma unroll
            for (int t = NB - 1; t >= 0; t--) {
                const double yt = __shfl_sync(0xffffffffu, wv * inv, t);
                wv = fma(-urow[t], yt, wv);  // urow[t] == 0 for t <= r
            }
            if (lane < nb) y[j0 + lane] = wv * inv;
        }
        __syncthreads();
        for (int i = tid; i < j0; i += nt) {
            double sacc = 0;
#pragma unroll 4
            for (int t = 0; t < nb; t++) sacc = fma(S.at(i, j0 + t), y[j0 + t], sacc);
            y[i] -= sacc;
        }
        __syncthreads();
    }
}

__global__ void __launch_bounds__(CHOL_THREADS) chol_kernel(DevProblem P, SolverScratch w, const double *x0, const double *x1,
                                                           SolverCtl *ctl, int use_smem) {
    if (ctl->done != 0 || ctl->reuse != 0) return;
    extern __shared__ double smem[];
    __shared__ double s_red[33];
    __shared__ __align__(16) double s_U[2 * NB * NB];
    __shared__ __align__(16) double s_row[2 * NB];
    __shared__ int s_okv[kMaxFrames * 15 / NB + 4];
    const StateLayout lay = P.lay;
    const lfvio_solver_opts &o = P.o;
    const int D = lay.D();
    const int tid = threadIdx.x, nt = blockDim.x;
    const int nblk = (D + NB - 1) / NB;
    const int PW = chol_panel_pitch(D);
    double *s_panel = smem;                        // NB * PW (16-byte aligned rows)
    double *s_y = s_panel + (size_t)NB * PW;       // D (padded to even)
    double *s_invd = s_y + ((D + 1) & ~1);         // nblk * NB reciprocal pivots
    double *s_S = s_invd + (size_t)nblk * NB;      // packed upper triangle when use_smem
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
    for (int q = tid; q < NB * PW; q += nt) s_panel[q] = 0.0;
    for (int c = tid; c < D; c += nt) s_y[c] = w.rhs[c];
    if (use_smem) {
        const MatPacked Sp{s_S, D};
        for (int r = tid >> 5; r < D; r += nt >> 5) {  // one warp per row: coalesced reads of the upper triangle
            const double *src = w.S + (size_t)r * D;
            double *dst = &Sp.at(r, r) - r;
            for (int c = r + (tid & 31); c < D; c += 32) dst[c] = src[c] + (r == c ? w.diagonal[c] * w.diagonal[c] * mu : 0.0);
        }
        __syncthreads();
        PHASE_MARK(2);
        pd = cholesky_lookahead(Sp, D, s_panel, PW, s_U, s_invd, s_y, s_okv, s_row);
    } else {
        for (int c = tid; c < D; c += nt) w.S[(size_t)c * D + c] += w.diagonal[c] * w.diagonal[c] * mu;
        __syncthreads();
        pd = cholesky_lookahead(MatFull{w.S, D}, D, s_panel, PW, s_U, s_invd, s_y, s_okv, s_row);
    }
    if (!pd) {
        if (tid == 0) ctl->pad = 0;
        return;
    }
    PHASE_MARK(3);
    if (use_smem) solve_upper_back(MatPacked{s_S, D}, D, s_y, s_invd);
    else solve_upper_back(MatFull{w.S, D}, D, s_y, s_invd);
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

#ifdef LFVIO_CHOL_PROFILE
void chol_profile_read(long long *out) { cudaMemcpyFromSymbol(out, g_sub, sizeof(long long) * 16); }
#endif

int launch_step(const DevProblem &P, const EvalBuf &e0, const EvalBuf &e1, const SolverScratch &w, double *x0, double *x1,
                SolverCtl *ctl, cudaStream_t s) {
    const int D = P.lay.D(), L = P.lay.L;
    const int work = std::max(std::max(D + L, P.n_proj), 1);
    const int grid_rows = std::max(1, std::min(cdiv(work, PRE_THREADS), 148));
    pre_kernel<<<grid_rows, PRE_THREADS, 0, s>>>(P, e0, e1, w, ctl);
    LFVIO_LAUNCH_CHECK();
    LFVIO_REQUIRE(D + 1 <= CHOL_THREADS - 32, "window too large for the single-pass panel solve (15 F + 8 <= 480)");
    const int nblk = cdiv(D, NB);
    const size_t base = sizeof(double) * ((size_t)NB * chol_panel_pitch(D) + ((D + 1) & ~1) + (size_t)nblk * NB);
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
