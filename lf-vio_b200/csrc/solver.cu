// K11-K13 — normal equations, Schur elimination of the inverse-depth landmarks, dense pose solve, dogleg.
//
// Reference: ceres::Solve(options, &problem, &summary) at vins_estimator/src/estimator.cpp:810-825 with
// linear_solver_type = DENSE_SCHUR, trust_region_strategy_type = DOGLEG, max_num_iterations = NUM_ITERATIONS.
// Ceres 1.12.0 is third-party (pinned docker/Dockerfile:3, not vendored); its published algorithm —
// TrustRegionMinimizer, DoglegStrategy (TRADITIONAL_DOGLEG: Gauss-Newton + Cauchy point, mu-regularised
// linear solve, radius update), Jacobi column scaling, SchurEliminator — is restated in SURVEY.md App. B.4
// and followed here step for step; the fp64 CPU oracle (oracle/backend/backend.cpp) states the same thing.
// One deviation, shared with the oracle: a failed factorisation costs one (invalid) iteration instead of
// being retried with 10x mu inside the same iteration.
//
// Per trust-region iteration, all on device (the control block SolverCtl never visits the host):
//   accumulate_kernel  one warp per landmark: loads its factors' Jacobians (Jacobi-scaled) into shared memory,
//                      forms the landmark's 1x1 pivot c_l, the coupling vector w_l and adds the landmark's
//                      complete contribution  sum_k J_k^T J_k - w_l w_l^T / c_l  to the reduced system with
//                      fp64 atomics (upper triangle) — the Schur complement is never materialised per landmark;
//                      one warp per IMU factor (15x30 block), one CTA for the prior (precomputed J0^T J0).
//   step_kernel        one CTA: Cauchy point, blocked Cholesky of the D x D reduced system (panel in shared
//                      memory), back-substitution of the landmarks, dogleg interpolation, model cost change,
//                      candidate state x (+) delta, parameter-tolerance test.
//   evaluate_kernel    (factors.cu) candidate evaluation + accept/reject bookkeeping in its last CTA.
#include "backend.cuh"

namespace lfvio {

namespace {

constexpr int ACC_WARPS = 4;
constexpr int STEP_THREADS = 512;  // 128 registers per thread: the fp64 panel solve keeps 16 doubles live
constexpr int NB = 16;  // Cholesky panel width

__device__ __forceinline__ long long gtimer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return (long long)t;
}
__device__ __forceinline__ double clamp_diag(double v) { return fmin(fmax(v, 1e-6), 1e32); }

// local column c (0..19) of a projection factor -> column of the full (reduced | landmark) system
__device__ __forceinline__ int proj_col(const StateLayout &lay, int frame_i, int frame_j, int landmark, int c) {
    if (c < 6) return lay.col_pose(frame_i) + c;
    if (c < 12) return lay.col_pose(frame_j) + (c - 6);
    if (c < 18) return lay.col_ex() + (c - 12);
    if (c == 18) return lay.D() + landmark;
    return lay.col_td();
}
__device__ __forceinline__ int imu_col(const StateLayout &lay, int i, int c) {
    if (c < 6) return lay.col_pose(i) + c;
    if (c < 15) return lay.col_sb(i) + (c - 6);
    if (c < 21) return lay.col_pose(i + 1) + (c - 15);
    return lay.col_sb(i + 1) + (c - 21);
}

// ---- Jacobi scaling: scale = 1 / (1 + sqrt(sum_i J_ic^2)) at the initial point ---------------------
__global__ void colnorm_kernel(DevProblem P, EvalBuf e, double *scale) {
    const StateLayout lay = P.lay;
    const int tid = blockIdx.x * blockDim.x + threadIdx.x, nth = gridDim.x * blockDim.x;
    for (int k = tid; k < P.n_proj; k += nth) {
        const lfvio_proj_factor &f = P.proj[k];
        const double *J = e.Jp + 40 * (size_t)k;
        for (int c = 0; c < 20; c++) {
            const double v = J[c] * J[c] + J[20 + c] * J[20 + c];
            if (v != 0.0) atomicAdd(&scale[proj_col(lay, f.frame_i, f.frame_j, f.landmark, c)], v);
        }
    }
    for (int q = tid; q < P.n_imu * 30; q += nth) {
        const int k = q / 30, c = q % 30;
        const double *J = e.Ji + 450 * (size_t)k;
        double v = 0;
        for (int r = 0; r < 15; r++) v += J[r * 30 + c] * J[r * 30 + c];
        if (v != 0.0) atomicAdd(&scale[imu_col(lay, P.imu[k].frame_i, c)], v);
    }
    const DevPrior &p = P.prior;
    for (int b = 0; b < p.n_blocks; b++) {
        const int ls = lay.block_lsize(p.block_id[b]), c0 = lay.block_col(p.block_id[b]);
        for (int q = tid; q < ls; q += nth) {
            double v = 0;
            for (int i = 0; i < p.n; i++) {
                const double j = p.J0[(size_t)i * p.n + p.block_idx[b] + q];
                v += j * j;
            }
            atomicAdd(&scale[c0 + q], v);
        }
    }
}
__global__ void scale_finalize_kernel(double *scale, int n) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c < n) scale[c] = 1.0 / (1.0 + sqrt(scale[c]));
}

// ---- accumulate -----------------------------------------------------------------------------------
__device__ __forceinline__ void add_upper(double *S, int D, int ga, int gb, double v) {
    if (v == 0.0) return;
    if (ga <= gb) atomicAdd(&S[(size_t)ga * D + gb], v);
    else atomicAdd(&S[(size_t)gb * D + ga], v);
}

__global__ void __launch_bounds__(ACC_WARPS * 32) accumulate_kernel(DevProblem P, EvalBuf e0, EvalBuf e1, SolverScratch w,
                                                                   const SolverCtl *ctl, int n_lm_cta, int max_obs) {
    if (ctl->done != 0 || ctl->reuse != 0) return;
    extern __shared__ double smem[];
    const EvalBuf &e = ctl->cur ? e1 : e0;
    const StateLayout lay = P.lay;
    const int D = lay.D();
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const double mu = ctl->mu;
    if ((int)blockIdx.x < n_lm_cta) {
        const int item = blockIdx.x * ACC_WARPS + wid;
        // per-warp shared memory: sJ[max_obs][40], sR[max_obs][2], sW[max_obs][6], sWs[13], sFj[max_obs] (as double slots)
        const int per_warp = max_obs * 40 + max_obs * 2 + max_obs * 6 + 16 + max_obs;
        double *sJ = smem + (size_t)wid * per_warp;
        double *sR = sJ + max_obs * 40;
        double *sW = sR + max_obs * 2;
        double *sWs = sW + max_obs * 6;
        int *sFj = reinterpret_cast<int *>(sWs + 16);
        if (item < lay.L) {
            // ===== one landmark =====
            const int l = item;
            const int k0 = P.lm_first[l], K = min(P.lm_first[l + 1] - k0, max_obs);
            if (K <= 0) {
                if (lane == 0) {
                    w.lm_c[l] = 0.0;
                    w.lm_g[l] = 0.0;
                }
                return;
            }
            const int fi = P.proj[k0].frame_i;
            const double sc_l = w.scale[D + l];
            for (int k = lane; k < K; k += 32) sFj[k] = P.proj[k0 + k].frame_j;
            __syncwarp();
            for (int q = lane; q < K * 40; q += 32) {
                const int k = q / 40, c = (q % 40) % 20;
                const int col = proj_col(lay, fi, sFj[k], l, c);
                sJ[q] = e.Jp[40 * (size_t)(k0 + k) + (q % 40)] * w.scale[col];
            }
            for (int q = lane; q < K * 2; q += 32) sR[q] = e.rp[2 * (size_t)k0 + q];
            __syncwarp();
            (void)sc_l;
            // pivot and landmark gradient
            double c_l = 0, g_l = 0;
            for (int q = lane; q < K * 2; q += 32) {
                const double jl = sJ[(q >> 1) * 40 + (q & 1) * 20 + 18];
                c_l += jl * jl;
                g_l += jl * sR[q];
            }
            c_l = warp_sum(c_l);
            g_l = warp_sum(g_l);
            const double c_reg = c_l + clamp_diag(c_l) * mu;
            const double inv_c = 1.0 / c_reg;
            // shared columns a: 0..5 pose_i, 6..11 ex, 12 td -> factor-local column
            auto shared_c = [](int a) { return a < 6 ? a : (a < 12 ? a + 6 : 19); };
            if (lane < 13) {
                const int c = shared_c(lane);
                double s = 0;
                for (int k = 0; k < K; k++) s += sJ[k * 40 + c] * sJ[k * 40 + 18] + sJ[k * 40 + 20 + c] * sJ[k * 40 + 38];
                sWs[lane] = s;
            }
            for (int q = lane; q < K * 6; q += 32) {
                const int k = q / 6, c = 6 + q % 6;
                sW[q] = sJ[k * 40 + c] * sJ[k * 40 + 18] + sJ[k * 40 + 20 + c] * sJ[k * 40 + 38];
            }
            __syncwarp();
            // persist what the back-substitution needs
            if (lane == 0) {
                w.lm_c[l] = c_l;
                w.lm_g[l] = g_l;
            }
            if (lane < 13) w.lm_ws[13 * (size_t)l + lane] = sWs[lane];
            for (int q = lane; q < K * 6; q += 32) w.proj_wj[6 * (size_t)k0 + q] = sW[q];
            // local index A in [0, 13 + 6K): global column and accessors
            const int n_l = 13 + 6 * K;
            auto gcol = [&](int A) {
                if (A < 6) return lay.col_pose(fi) + A;
                if (A < 12) return lay.col_ex() + (A - 6);
                if (A == 12) return lay.col_td();
                return lay.col_pose(sFj[(A - 13) / 6]) + (A - 13) % 6;
            };
            auto wval = [&](int A) { return A < 13 ? sWs[A] : sW[A - 13]; };
            // gradient, squared column norms, reduced right-hand side
            for (int A = lane; A < n_l; A += 32) {
                double g = 0, d2 = 0;
                if (A < 13) {
                    const int c = shared_c(A);
                    for (int k = 0; k < K; k++) {
                        const double j0 = sJ[k * 40 + c], j1 = sJ[k * 40 + 20 + c];
                        g += j0 * sR[2 * k] + j1 * sR[2 * k + 1];
                        d2 += j0 * j0 + j1 * j1;
                    }
                } else {
                    const int k = (A - 13) / 6, c = 6 + (A - 13) % 6;
                    const double j0 = sJ[k * 40 + c], j1 = sJ[k * 40 + 20 + c];
                    g = j0 * sR[2 * k] + j1 * sR[2 * k + 1];
                    d2 = j0 * j0 + j1 * j1;
                }
                const int gc = gcol(A);
                if (g != 0.0) atomicAdd(&w.grad[gc], g);
                if (d2 != 0.0) atomicAdd(&w.diag2[gc], d2);
                const double rr = g - wval(A) * g_l * inv_c;
                if (rr != 0.0) atomicAdd(&w.rhs[gc], rr);
            }
            // reduced-system contribution: every unordered local pair (A <= B) exactly once
            for (int A = 0; A < n_l; A++) {
                const int gA = gcol(A);
                const double wA = wval(A);
                for (int B = A + lane; B < n_l; B += 32) {
                    double v = 0;
                    if (A < 13) {
                        const int ca = shared_c(A);
                        if (B < 13) {
                            const int cb = shared_c(B);
                            for (int k = 0; k < K; k++) v += sJ[k * 40 + ca] * sJ[k * 40 + cb] + sJ[k * 40 + 20 + ca] * sJ[k * 40 + 20 + cb];
                        } else {
                            const int k = (B - 13) / 6, cb = 6 + (B - 13) % 6;
                            v = sJ[k * 40 + ca] * sJ[k * 40 + cb] + sJ[k * 40 + 20 + ca] * sJ[k * 40 + 20 + cb];
                        }
                    } else {
                        const int ka = (A - 13) / 6, kb = (B - 13) / 6;
                        if (ka == kb) {
                            const int ca = 6 + (A - 13) % 6, cb = 6 + (B - 13) % 6;
                            v = sJ[ka * 40 + ca] * sJ[ka * 40 + cb] + sJ[ka * 40 + 20 + ca] * sJ[ka * 40 + 20 + cb];
                        }
                    }
                    v -= wA * wval(B) * inv_c;
                    add_upper(w.S, D, gA, gcol(B), v);
                }
            }
        } else if (item < lay.L + P.n_imu) {
            // ===== one IMU factor: dense 15 x 30 block =====
            const int k = item - lay.L;
            if (P.imu[k].sum_dt > 10.0) return;
            const int fi = P.imu[k].frame_i;
            double *sI = sJ;  // 450 doubles fit (per_warp >= 49 * max_obs + 16 >= 450 is ensured by the launcher)
            for (int q = lane; q < 450; q += 32) sI[q] = e.Ji[450 * (size_t)k + q] * w.scale[imu_col(lay, fi, q % 30)];
            double *sRi = sI + 450;
            for (int q = lane; q < 15; q += 32) sRi[q] = e.ri[15 * (size_t)k + q];
            __syncwarp();
            if (lane < 30) {
                double g = 0, d2 = 0;
                for (int r = 0; r < 15; r++) {
                    const double j = sI[r * 30 + lane];
                    g += j * sRi[r];
                    d2 += j * j;
                }
                const int gc = imu_col(lay, fi, lane);
                atomicAdd(&w.grad[gc], g);
                atomicAdd(&w.rhs[gc], g);
                atomicAdd(&w.diag2[gc], d2);
            }
            for (int A = 0; A < 30; A++) {
                const int gA = imu_col(lay, fi, A);
                for (int B = A + lane; B < 30; B += 32) {
                    double v = 0;
                    for (int r = 0; r < 15; r++) v += sI[r * 30 + A] * sI[r * 30 + B];
                    add_upper(w.S, D, gA, imu_col(lay, fi, B), v);
                }
            }
        }
        return;
    }
    // ===== prior CTA: S += scale (J0^T J0) scale, gradient += scale J0^T r =====
    const DevPrior &p = P.prior;
    if (p.n <= 0) return;
    int *pmap = reinterpret_cast<int *>(smem);
    for (int b = threadIdx.x; b < p.n_blocks; b += blockDim.x) {
        const int ls = lay.block_lsize(p.block_id[b]), c0 = lay.block_col(p.block_id[b]);
        for (int q = 0; q < ls; q++) pmap[p.block_idx[b] + q] = c0 + q;
    }
    __syncthreads();
    const int n = p.n;
    for (int q = threadIdx.x; q < n * n; q += blockDim.x) {
        const int a = q / n, b = q % n;
        const int ga = pmap[a], gb = pmap[b];
        if (ga <= gb) {
            const double v = p.H0[q] * w.scale[ga] * w.scale[gb];
            if (v != 0.0) atomicAdd(&w.S[(size_t)ga * D + gb], v);
            if (a == b) atomicAdd(&w.diag2[ga], v);
        }
    }
    for (int a = threadIdx.x; a < n; a += blockDim.x) {
        double g = 0;
        for (int i = 0; i < n; i++) g += p.J0[(size_t)i * n + a] * e.rm[i];
        g *= w.scale[pmap[a]];
        atomicAdd(&w.grad[pmap[a]], g);
        atomicAdd(&w.rhs[pmap[a]], g);
    }
}

// ---- step -----------------------------------------------------------------------------------------
__device__ __forceinline__ double block_sum_1024(double v, double *s_red) {
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
__device__ __forceinline__ double block_max_1024(double v, double *s_red) {
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

// sum over all residual rows of fn(row value of J_s v, residual): v is indexed by system column.
// mode 0: sum (Jv)^2        mode 1: sum -(Jv) (r + Jv / 2)
__device__ double rows_reduce(const DevProblem &P, const EvalBuf &e, const double *scale, const double *v, int mode,
                              const int *pmap, double *s_red) {
    const StateLayout lay = P.lay;
    double acc = 0;
    for (int k = threadIdx.x; k < P.n_proj; k += blockDim.x) {
        const lfvio_proj_factor &f = P.proj[k];
        const double *J = e.Jp + 40 * (size_t)k;
        double s0 = 0, s1 = 0;
        for (int c = 0; c < 20; c++) {
            const int col = proj_col(lay, f.frame_i, f.frame_j, f.landmark, c);
            const double vv = scale[col] * v[col];
            s0 += J[c] * vv;
            s1 += J[20 + c] * vv;
        }
        if (mode == 0) acc += s0 * s0 + s1 * s1;
        else acc -= s0 * (e.rp[2 * (size_t)k] + s0 / 2.0) + s1 * (e.rp[2 * (size_t)k + 1] + s1 / 2.0);
    }
    for (int q = threadIdx.x; q < P.n_imu * 15; q += blockDim.x) {
        const int k = q / 15, r = q % 15;
        if (P.imu[k].sum_dt > 10.0) continue;
        const double *J = e.Ji + 450 * (size_t)k + 30 * r;
        double s = 0;
        for (int c = 0; c < 30; c++) {
            const int col = imu_col(lay, P.imu[k].frame_i, c);
            s += J[c] * scale[col] * v[col];
        }
        if (mode == 0) acc += s * s;
        else acc -= s * (e.ri[q] + s / 2.0);
    }
    const DevPrior &p = P.prior;
    for (int i = threadIdx.x; i < p.n; i += blockDim.x) {
        double s = 0;
        for (int a = 0; a < p.n; a++) s += p.J0[(size_t)i * p.n + a] * scale[pmap[a]] * v[pmap[a]];
        if (mode == 0) acc += s * s;
        else acc -= s * (e.rm[i] + s / 2.0);
    }
    return block_sum_1024(acc, s_red);
}

// Blocked right-looking Cholesky S = U^T U on the upper triangle of the row-major D x D matrix, one CTA.
// Returns (to all threads) whether the matrix was positive definite.
// The matrix is addressed through an accessor: MatFull = row-major D x D in global memory (any D), MatPacked =
// packed upper triangle resident in shared memory (D <= ~220: the whole factorisation and both triangular
// solves then run without a single global-memory round trip).
struct MatFull {
    double *p;
    int D;
    __device__ __forceinline__ double &at(int r, int c) const { return p[(size_t)r * D + c]; }
};
struct MatPacked {
    double *p;
    int D;
    __device__ __forceinline__ double &at(int r, int c) const { return p[r * D - ((r * (r - 1)) >> 1) + (c - r)]; }
};

template <class M>
__device__ bool cholesky_upper(M S, int D, double *s_panel /* NB x D */, double *s_diag /* NB x NB */, int *s_ok) {
    const int tid = threadIdx.x, nt = blockDim.x;
    if (tid == 0) *s_ok = 1;
    __syncthreads();
    for (int j0 = 0; j0 < D; j0 += NB) {
        const int nb = min(NB, D - j0);
        // diagonal block -> shared, factor with one warp
        for (int q = tid; q < nb * nb; q += nt) {
            const int r = q / nb, c = q % nb;
            s_diag[r * NB + c] = (c >= r) ? S.at(j0 + r, j0 + c) : 0.0;
        }
        __syncthreads();
        if (tid < 32) {
            for (int j = 0; j < nb; j++) {
                double d = s_diag[j * NB + j];
                if (!(d > 0.0) || !isfinite(d)) {
                    if (tid == 0) *s_ok = 0;
                    d = 1.0;
                }
                d = sqrt(d);
                __syncwarp();
                if (tid == 0) s_diag[j * NB + j] = d;
                for (int c = j + 1 + tid; c < nb; c += 32) s_diag[j * NB + c] /= d;
                __syncwarp();
                for (int q = tid; q < (nb - j - 1) * (nb - j - 1); q += 32) {
                    const int r = j + 1 + q / (nb - j - 1), c = j + 1 + q % (nb - j - 1);
                    if (c >= r) s_diag[r * NB + c] -= s_diag[j * NB + r] * s_diag[j * NB + c];
                }
                __syncwarp();
            }
        }
        __syncthreads();
        if (*s_ok == 0) return false;
        for (int q = tid; q < nb * nb; q += nt) {
            const int r = q / nb, c = q % nb;
            if (c >= r) S.at(j0 + r, j0 + c) = s_diag[r * NB + c];
        }
        const int rest0 = j0 + nb, nrest = D - rest0;
        if (nrest > 0) {
            // row panel: U12 = U11^-T A12, one thread per column
            for (int c = tid; c < nrest; c += nt) {
                double u[NB];
#pragma unroll
                for (int r = 0; r < NB; r++) u[r] = (r < nb) ? S.at(j0 + r, rest0 + c) : 0.0;  // independent loads first
#pragma unroll
                for (int r = 0; r < NB; r++) {
                    if (r < nb) {
                        double s = u[r];
#pragma unroll
                        for (int t = 0; t < NB; t++)
                            if (t < r) s -= s_diag[t * NB + r] * u[t];
                        u[r] = s / s_diag[r * NB + r];
                    }
                }
#pragma unroll
                for (int r = 0; r < NB; r++) {
                    if (r < nb) {
                        S.at(j0 + r, rest0 + c) = u[r];
                        s_panel[r * D + c] = u[r];
                    }
                }
            }
            // zero padding so the 4x4 register tiles below need no bounds checks inside the rank-nb loop
            if (tid < NB * 4) {
                const int r = tid >> 2, c = nrest + (tid & 3);
                if (c < D) s_panel[r * D + c] = 0.0;
            }
            if (nb < NB)
                for (int q = tid; q < (NB - nb) * D; q += nt) s_panel[nb * D + q] = 0.0;
            __syncthreads();
            // trailing update of the upper triangle: A22[i][k] -= sum_t U12[t][i] U12[t][k], k >= i.
            // 4x4 register tiles: 8 shared-memory loads per 16 DFMA (the one-element-per-thread form is bound by
            // shared-memory bandwidth at 2 loads per DFMA).
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

// U^T z = b (forward) then U y = z (backward) on a shared-memory vector y (in/out), blocked by NB: the 16-step
// dependent chain of each diagonal block is run by warp 0, everything off the diagonal block is a parallel update
// (forward: all threads own a column; backward: one warp per row does a dot product).  Whole CTA must call.
template <class M>
__device__ void solve_upper_blocked(M S, int D, double *y, double *s_tmp /* NB */) {
    const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31, wid = tid >> 5;
    for (int j0 = 0; j0 < D; j0 += NB) {
        const int nb = min(NB, D - j0);
        if (wid == 0) {
            for (int j = j0; j < j0 + nb; j++) {
                const double zj = y[j] / S.at(j, j);
                __syncwarp();
                if (lane == 0) y[j] = zj;
                const int k = j + 1 + lane;
                if (k < j0 + nb) y[k] -= S.at(j, k) * zj;
                __syncwarp();
            }
        }
        __syncthreads();
        for (int k = j0 + nb + tid; k < D; k += nt) {
            double s = 0;
#pragma unroll 4
            for (int t = 0; t < nb; t++) s += S.at(j0 + t, k) * y[j0 + t];
            y[k] -= s;
        }
        __syncthreads();
    }
    const int nblk = (D + NB - 1) / NB;
    for (int bi = nblk - 1; bi >= 0; bi--) {
        const int j0 = bi * NB, nb = min(NB, D - j0), k0 = j0 + nb;
        for (int r = wid; r < nb; r += (nt >> 5)) {
            double s = 0;
            for (int k = k0 + lane; k < D; k += 32) s += S.at(j0 + r, k) * y[k];
            s = warp_sum(s);
            if (lane == 0) s_tmp[r] = s;
        }
        __syncthreads();
        if (wid == 0) {
            for (int j = j0 + nb - 1; j >= j0; j--) {
                double s = 0;
                const int k = j + 1 + lane;
                if (k < k0) s = S.at(j, k) * y[k];
                s = warp_sum(s);
                if (lane == 0) y[j] = (y[j] - s_tmp[j - j0] - s) / S.at(j, j);
                __syncwarp();
            }
        }
        __syncthreads();
    }
}

__global__ void __launch_bounds__(STEP_THREADS) step_kernel(DevProblem P, EvalBuf e0, EvalBuf e1, SolverScratch w, double *x0,
                                                           double *x1, SolverCtl *ctl, int use_smem) {
    if (ctl->done != 0) return;
    extern __shared__ double smem[];
    __shared__ double s_red[33];
    __shared__ double s_diag[NB * NB];
    __shared__ int s_ok;
    __shared__ double s_sc[NB];
    const StateLayout lay = P.lay;
    const lfvio_solver_opts &o = P.o;
    const int D = lay.D(), L = lay.L, N = D + L;
    const int tid = threadIdx.x, nt = blockDim.x;
    const int cur = ctl->cur;
    const EvalBuf &e = cur ? e1 : e0;
    double *x = cur ? x1 : x0, *xc = cur ? x0 : x1;
    double *s_panel = smem;                                  // NB * D doubles
    int *pmap = reinterpret_cast<int *>(smem + (size_t)NB * D);  // prior column map (15 * kMaxFrames + 8 ints)
    double *s_y = smem + (size_t)NB * D + (15 * kMaxFrames + 8 + 1) / 2;  // D doubles
    double *s_S = s_y + D;                                   // packed upper triangle when use_smem
    {
        const DevPrior &p = P.prior;
        for (int b = tid; b < p.n_blocks; b += nt) {
            const int ls = lay.block_lsize(p.block_id[b]), c0 = lay.block_col(p.block_id[b]);
            for (int q = 0; q < ls; q++) pmap[p.block_idx[b] + q] = c0 + q;
        }
    }
    if (tid == 0) ctl->phase_clk[0] = gtimer_ns();
    // x_norm over the non-constant parameter blocks (global sizes)
    {
        double a = 0;
        for (int q = tid; q < lay.size(); q += nt) {
            const bool is_ex = q >= lay.ex() && q < lay.ex() + 7, is_td = q == lay.td();
            if ((is_ex && !o.estimate_extrinsic) || (is_td && !o.estimate_td)) continue;
            a += x[q] * x[q];
        }
        const double t = block_sum_1024(a, s_red);
        if (tid == 0) ctl->x_norm = sqrt(t);
    }
    __syncthreads();
    int step_valid = 1;
    if (!ctl->reuse) {
        const double mu = ctl->mu;
        // diagonal_ and gradient_ of DoglegStrategy (in the Jacobi-scaled space)
        double gmax = 0;
        for (int c = tid; c < N; c += nt) {
            const double d2 = c < D ? w.diag2[c] : w.lm_c[c - D];
            const double g = c < D ? w.grad[c] : w.lm_g[c - D];
            const double dg = sqrt(clamp_diag(d2));
            w.diagonal[c] = dg;
            w.gradient[c] = g / dg;
            gmax = fmax(gmax, fabs(g / w.scale[c]));  // un-scaled gradient J^T r
        }
        gmax = block_max_1024(gmax, s_red);
        if (!o.fixed_iterations && gmax <= o.gradient_tolerance) {
            if (tid == 0) ctl->done = 1;
            return;
        }
        if (use_smem) {
            const MatPacked Sp{s_S, D};
            for (int q = tid; q < D * D; q += nt) {
                const int r = q / D, c = q - r * D;
                if (c >= r) Sp.at(r, c) = w.S[q] + (r == c ? w.diagonal[c] * w.diagonal[c] * mu : 0.0);
            }
        } else {
            for (int c = tid; c < D; c += nt) w.S[(size_t)c * D + c] += w.diagonal[c] * w.diagonal[c] * mu;
        }
        __syncthreads();
    if (tid == 0) ctl->phase_clk[1] = gtimer_ns();
        // Cauchy point: alpha = |gradient|^2 / |J_s (gradient / diagonal)|^2
        for (int c = tid; c < N; c += nt) w.step[c] = w.gradient[c] / w.diagonal[c];  // step used as scratch
        double gg = 0;
        for (int c = tid; c < N; c += nt) gg += w.gradient[c] * w.gradient[c];
        gg = block_sum_1024(gg, s_red);
        __syncthreads();
        const double jj = rows_reduce(P, e, w.scale, w.step, 0, pmap, s_red);
        if (tid == 0) ctl->alpha = gg / jj;
    if (tid == 0) ctl->phase_clk[2] = gtimer_ns();
        // Gauss-Newton step through the reduced system
        const bool pd = use_smem ? cholesky_upper(MatPacked{s_S, D}, D, s_panel, s_diag, &s_ok)
                                 : cholesky_upper(MatFull{w.S, D}, D, s_panel, s_diag, &s_ok);
        if (!pd) {
            step_valid = 0;
        } else {
            // U^T z = rhs (forward), U y = z (backward): one warp, row accesses are coalesced
    if (tid == 0) ctl->phase_clk[3] = gtimer_ns();
            for (int c = tid; c < D; c += nt) s_y[c] = w.rhs[c];
            __syncthreads();
            if (use_smem) solve_upper_blocked(MatPacked{s_S, D}, D, s_y, s_sc);
            else solve_upper_blocked(MatFull{w.S, D}, D, s_y, s_sc);
            __syncthreads();
    if (tid == 0) ctl->phase_clk[4] = gtimer_ns();
            int bad = 0;
            for (int c = tid; c < D; c += nt) {
                const double yv = s_y[c];
                if (!isfinite(yv)) bad = 1;
                w.gn[c] = -w.diagonal[c] * yv;  // scaled GN step: D * (-(J^T J)^-1 g)
            }
            // landmarks: y_l = (g_l - w_l^T y_p) / c_l
            for (int l = tid; l < L; l += nt) {
                const int k0 = P.lm_first[l], K = P.lm_first[l + 1] - k0;
                double yl = 0;
                if (K > 0) {
                    const int fi = P.proj[k0].frame_i;
                    const double *ws = w.lm_ws + 13 * (size_t)l;
                    double s = w.lm_g[l];
                    for (int a = 0; a < 6; a++) s -= ws[a] * s_y[lay.col_pose(fi) + a];
                    for (int a = 0; a < 6; a++) s -= ws[6 + a] * s_y[lay.col_ex() + a];
                    s -= ws[12] * s_y[lay.col_td()];
                    for (int k = 0; k < K; k++) {
                        const double *wj = w.proj_wj + 6 * (size_t)(k0 + k);
                        const int cj = lay.col_pose(P.proj[k0 + k].frame_j);
                        for (int a = 0; a < 6; a++) s -= wj[a] * s_y[cj + a];
                    }
                    const double c_l = w.lm_c[l];
                    yl = s / (c_l + clamp_diag(c_l) * mu);
                }
                if (!isfinite(yl)) bad = 1;
                w.gn[D + l] = -w.diagonal[D + l] * yl;
            }
            if (__syncthreads_or(bad)) step_valid = 0;
        }
    }
    __syncthreads();
    double model_cost_change = 0, step_norm_x = 0;
    if (step_valid) {
    if (tid == 0) ctl->phase_clk[5] = gtimer_ns();
        // ---- traditional dogleg ----
        double a0 = 0, a1 = 0, a2 = 0;
        for (int c = tid; c < N; c += nt) {
            const double g = w.gradient[c], n = w.gn[c];
            a0 += g * g;
            a1 += n * n;
            a2 += g * n;
        }
        const double gnorm = sqrt(block_sum_1024(a0, s_red));
        const double gnn = sqrt(block_sum_1024(a1, s_red));
        const double gdot = block_sum_1024(a2, s_red);
        const double radius = ctl->radius, alpha = ctl->alpha;
        double cg = 0, cn = 0;  // step = cg * gradient + cn * gn
        double dsn = -1.0;
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
        }
        double sn = 0;
        for (int c = tid; c < N; c += nt) {
            const double s = cg * w.gradient[c] + cn * w.gn[c];
            sn += s * s;
            w.step[c] = s / w.diagonal[c];
        }
        sn = block_sum_1024(sn, s_red);
        if (dsn < 0) dsn = sqrt(sn);
        __syncthreads();
    if (tid == 0) ctl->phase_clk[6] = gtimer_ns();
        // model_cost_change = -(J step) . (r + J step / 2)
        model_cost_change = rows_reduce(P, e, w.scale, w.step, 1, pmap, s_red);
        if (!(model_cost_change > 0.0)) step_valid = 0;
        if (tid == 0) {
            ctl->dogleg_step_norm = dsn;
            ctl->model_cost_change = model_cost_change;
        }
    }
    if (step_valid) {
    if (tid == 0) ctl->phase_clk[7] = gtimer_ns();
        // ---- candidate: x_plus_delta = Plus(x, step * scale) ----
        for (int q = tid; q < lay.size(); q += nt) xc[q] = x[q];
        __syncthreads();
        for (int f = tid; f < lay.F; f += nt) {
            double d[6];
            for (int a = 0; a < 6; a++) d[a] = w.step[lay.col_pose(f) + a] * w.scale[lay.col_pose(f) + a];
            pose_plus(x + lay.pose(f), d, xc + lay.pose(f));
            for (int a = 0; a < 9; a++) xc[lay.sb(f) + a] = x[lay.sb(f) + a] + w.step[lay.col_sb(f) + a] * w.scale[lay.col_sb(f) + a];
        }
        if (tid == 0) {
            if (o.estimate_extrinsic) {
                double d[6];
                for (int a = 0; a < 6; a++) d[a] = w.step[lay.col_ex() + a] * w.scale[lay.col_ex() + a];
                pose_plus(x + lay.ex(), d, xc + lay.ex());
            }
            if (o.estimate_td) xc[lay.td()] = x[lay.td()] + w.step[lay.col_td()] * w.scale[lay.col_td()];
        }
        for (int l = tid; l < L; l += nt) xc[lay.feat(l)] = x[lay.feat(l)] + w.step[D + l] * w.scale[D + l];
        __syncthreads();
        double a = 0;
        for (int q = tid; q < lay.size(); q += nt) {
            const bool is_ex = q >= lay.ex() && q < lay.ex() + 7, is_td = q == lay.td();
            if ((is_ex && !o.estimate_extrinsic) || (is_td && !o.estimate_td)) continue;
            const double dlt = x[q] - xc[q];
            a += dlt * dlt;
        }
        step_norm_x = sqrt(block_sum_1024(a, s_red));
    }
    if (tid == 0) ctl->phase_clk[8] = gtimer_ns();
    if (tid == 0) {
        (void)s_sc;
        ctl->step_valid = step_valid;
        ctl->step_norm = step_norm_x;
        ctl->cost_cand = 0.0;
        ctl->eval_pending = step_valid;
        if (step_valid && !o.fixed_iterations &&
            step_norm_x <= o.parameter_tolerance * (ctl->x_norm + o.parameter_tolerance)) {
            ctl->done = 1;  // "Parameter tolerance reached": terminates without taking the step
            ctl->eval_pending = 0;
        }
    }
}

__global__ void solver_init_kernel(SolverCtl *ctl, lfvio_solver_opts o) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    ctl->cur = 0;
    ctl->reuse = 0;
    ctl->done = 0;
    ctl->iter = 0;
    ctl->eval_pending = 0;
    ctl->step_valid = 0;
    ctl->successful = 0;
    ctl->invalid_in_a_row = 0;
    ctl->ticket = 0;
    ctl->radius = o.initial_trust_region_radius;
    ctl->mu = 1e-8;
    ctl->cost = 0.0;
    ctl->cost_cand = 0.0;
    ctl->x_norm = 0.0;
    ctl->dogleg_step_norm = 0.0;
    ctl->alpha = 0.0;
    ctl->model_cost_change = 0.0;
    ctl->step_norm = 0.0;
}

}  // namespace

int launch_solver_init(SolverCtl *ctl, const lfvio_solver_opts &o, cudaStream_t s) {
    solver_init_kernel<<<1, 32, 0, s>>>(ctl, o);
    LFVIO_LAUNCH_CHECK();
    return LFVIO_OK;
}

int launch_jacobi_scale(const DevProblem &P, const EvalBuf &e0, const SolverScratch &w, cudaStream_t s) {
    const int N = P.lay.D() + P.lay.L;
    LFVIO_CUDA_TRY(cudaMemsetAsync(w.scale, 0, sizeof(double) * N, s));
    colnorm_kernel<<<std::max(1, std::min(cdiv(P.n_proj, 128), 296)), 128, 0, s>>>(P, e0, w.scale);
    LFVIO_LAUNCH_CHECK();
    scale_finalize_kernel<<<cdiv(N, 256), 256, 0, s>>>(w.scale, N);
    LFVIO_LAUNCH_CHECK();
    return LFVIO_OK;
}

static int acc_per_warp_doubles(int max_obs) { return max_obs * 40 + max_obs * 2 + max_obs * 6 + 16 + max_obs; }

int launch_accumulate(const DevProblem &P, const EvalBuf &e0, const EvalBuf &e1, const SolverScratch &w, SolverCtl *ctl,
                      cudaStream_t s) {
    LFVIO_CUDA_TRY(cudaMemsetAsync(w.S, 0, w.zero_bytes, s));
    int max_obs = std::max(P.lay.F - 1, 10);  // 49 * 10 + 16 = 506 doubles >= 465 needed by an IMU warp
    const int items = P.lay.L + P.n_imu;
    const int n_lm_cta = cdiv(items, ACC_WARPS);
    size_t smem = sizeof(double) * (size_t)acc_per_warp_doubles(max_obs) * ACC_WARPS;
    smem = std::max(smem, sizeof(int) * (size_t)(15 * kMaxFrames + 8));
    static bool attr_set = false;
    if (!attr_set) {
        LFVIO_CUDA_TRY(cudaFuncSetAttribute(accumulate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
        attr_set = true;
    }
    accumulate_kernel<<<n_lm_cta + 1, ACC_WARPS * 32, smem, s>>>(P, e0, e1, w, ctl, n_lm_cta, max_obs);
    LFVIO_LAUNCH_CHECK();
    return LFVIO_OK;
}

int launch_step(const DevProblem &P, const EvalBuf &e0, const EvalBuf &e1, const SolverScratch &w, double *x0, double *x1,
                SolverCtl *ctl, cudaStream_t s) {
    const int D = P.lay.D();
    const size_t base = sizeof(double) * ((size_t)NB * D + (15 * kMaxFrames + 8 + 1) / 2 + D);
    const size_t packed = sizeof(double) * ((size_t)D * (D + 1) / 2);
    const size_t limit = 220 * 1024;  // 227 KB per CTA minus the static arrays
    const int use_smem = (base + packed <= limit) ? 1 : 0;
    const size_t smem = base + (use_smem ? packed : 0);
    static bool attr_set = false;
    if (!attr_set) {
        LFVIO_CUDA_TRY(cudaFuncSetAttribute(step_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)limit));
        attr_set = true;
    }
    step_kernel<<<1, STEP_THREADS, smem, s>>>(P, e0, e1, w, x0, x1, ctl, use_smem);
    LFVIO_LAUNCH_CHECK();
    return LFVIO_OK;
}

}  // namespace lfvio
