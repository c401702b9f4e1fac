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
//   accumulate_kernel  one CTA (4 warps) per landmark: loads its factors' Jacobians (Jacobi-scaled) into shared memory,
//                      forms the landmark's 1x1 pivot c_l, the coupling vector w_l and adds the landmark's
//                      complete contribution  sum_k J_k^T J_k - w_l w_l^T / c_l  to the reduced system with
//                      fp64 atomics (upper triangle) — the Schur complement is never materialised per landmark;
//                      one CTA per IMU factor (15x30 block), PRIOR_CTAS CTAs for the prior (precomputed J0^T J0).
//   pre/chol/gnlm/cand (solver_step.cu) Cauchy point (multi-CTA), Cholesky + triangular solves of the D x D reduced
//                      system (one CTA, shared-memory resident), landmark back-substitution (warp per landmark),
//                      dogleg + model cost + candidate state x (+) delta (multi-CTA).
//   evaluate_kernel    (factors.cu) candidate evaluation + accept/reject bookkeeping in its last CTA.
#include "backend.cuh"
#include "solver_common.cuh"

namespace lfvio {

namespace {

constexpr int ACC_WARPS = 8;    // warps of one CTA; its threads share the pairs of an item's local block
constexpr int PRIOR_CTAS = 16;  // CTAs that share the prior's n x n block and its gradient

// ---- Jacobi scaling: scale = 1 / (1 + sqrt(sum_i J_ic^2)) at the initial point ---------------------
// One thread per Jacobian ENTRY (coalesced reads of the 40-double factor rows; a thread per factor walked its 40 entries with a
// stride of 320 bytes between lanes: 30 us per solve), reduced columns summed in shared memory and flushed once per CTA,
// landmark columns straight to global (two entries per factor, few factors per address).
constexpr int CN_THREADS = 256, CN_MAX_CTAS = 48;
__global__ void __launch_bounds__(CN_THREADS) colnorm_kernel(const DevProblem *__restrict__ Pd, EvalBuf e, double *scale) {
    const DevProblem &P = *Pd;
    const StateLayout lay = P.lay;
    __shared__ double s_col[15 * kMaxFrames + 8];
    const int D = lay.D();
    const int tid = blockIdx.x * blockDim.x + threadIdx.x, nth = gridDim.x * blockDim.x;
    for (int c = threadIdx.x; c < D; c += blockDim.x) s_col[c] = 0.0;
    __syncthreads();
    for (int q = tid; q < 40 * P.n_proj; q += nth) {
        const double v = e.Jp[q];
        if (v == 0.0) continue;
        const int k = q / 40, c = (q - 40 * k) % 20;
        const lfvio_proj_factor &f = P.proj[k];
        const int col = proj_col(lay, f.frame_i, f.frame_j, f.landmark, c);
        if (col < D) atomicAdd(&s_col[col], v * v);
        else atomicAdd(&scale[col], v * v);
    }
    for (int q = tid; q < 450 * P.n_imu; q += nth) {
        const double v = e.Ji[q];
        if (v == 0.0) continue;
        const int k = q / 450, c = (q - 450 * k) % 30;
        atomicAdd(&s_col[imu_col(lay, P.imu[k].frame_i, c)], v * v);
    }
    // prior: squared column norms of J0 = diagonal of H0 = J0^T J0 (prior_prepare_kernel, once per upload)
    const DevPrior &p = P.prior;
    for (int b = 0; b < p.n_blocks; b++) {
        const int ls = lay.block_lsize(p.block_id[b]), c0 = lay.block_col(p.block_id[b]);
        for (int q = tid; q < ls; q += nth) atomicAdd(&scale[c0 + q], p.H0[(size_t)(p.block_idx[b] + q) * (p.n + 1)]);
    }
    __syncthreads();
    for (int c = threadIdx.x; c < D; c += blockDim.x)
        if (s_col[c] != 0.0) atomicAdd(&scale[c], s_col[c]);
}
__global__ void scale_finalize_kernel(const DevProblem *__restrict__ Pd, double *scale) {
    const int n = Pd->lay.D() + Pd->lay.L;
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c < n) scale[c] = 1.0 / (1.0 + sqrt(scale[c]));
}

// ---- accumulate -----------------------------------------------------------------------------------
__device__ __forceinline__ void add_upper(double *S, int D, int ga, int gb, double v) {
    if (v == 0.0) return;
#ifdef LFVIO_ACC_NOATOMIC  // timing experiment only (wrong sums)
    if (ga <= gb) S[(size_t)ga * D + gb] = v;
    else S[(size_t)gb * D + ga] = v;
    return;
#endif
    if (ga <= gb) atomicAdd(&S[(size_t)ga * D + gb], v);
    else atomicAdd(&S[(size_t)gb * D + ga], v);
}

CTA_PROF_DECL(g_acc_prof);
__global__ void __launch_bounds__(ACC_WARPS * 32) accumulate_kernel(const DevProblem *__restrict__ Pd, EvalBuf e0, EvalBuf e1, SolverScratch w,
                                                                   const SolverCtl *ctl, int n_lm_cta, int g_L, int max_obs) {
    const DevProblem &P = *Pd;
    CTA_PROF(g_acc_prof, 0);
    pdl_prologue();
    CTA_PROF(g_acc_prof, 1);
    if (ctl->done != 0 || ctl->reuse != 0) return;
    extern __shared__ double smem[];
    const EvalBuf &e = ctl->cur ? e1 : e0;
    const StateLayout lay = P.lay;
    const int D = lay.D();
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const double mu = ctl->mu;
    if ((int)blockIdx.x < n_lm_cta) {
        const int item = blockIdx.x;
        const int tid = threadIdx.x;
        // shared memory: sJ[max_obs][40], sR[max_obs][2], sW[max_obs][6], sWs[13], sFj[max_obs] (as double slots)
        double *sJ = smem;
        double *sR = sJ + max_obs * 40;
        double *sW = sR + max_obs * 2;
        double *sWs = sW + max_obs * 6;
        int *sFj = reinterpret_cast<int *>(sWs + 16);
        if (item < lay.L) {
            // ===== one landmark =====
            const int l = item;
            const int k0 = P.lm_first[l], K = min(P.lm_first[l + 1] - k0, max_obs);
            if (K <= 0) {
                if (tid == 0) {
                    w.lm_c[l] = 0.0;
                    w.lm_g[l] = 0.0;
                    w.lm_coli[l] = -1;
                }
                return;
            }
            const int fi = P.proj[k0].frame_i;
            const double sc_l = w.scale[D + l];
            for (int k = tid; k < K; k += ACC_WARPS * 32) sFj[k] = P.proj[k0 + k].frame_j;
            __syncthreads();
            for (int q = tid; q < K * 40; q += ACC_WARPS * 32) {
                const int k = q / 40, c = (q % 40) % 20;
                const int col = proj_col(lay, fi, sFj[k], l, c);
                sJ[q] = e.Jp[40 * (size_t)(k0 + k) + (q % 40)] * w.scale[col];
            }
            for (int q = tid; q < K * 2; q += ACC_WARPS * 32) sR[q] = e.rp[2 * (size_t)k0 + q];
            __syncthreads();
            (void)sc_l;
            // pivot and landmark gradient (every warp computes them, identically)
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
            if (tid < 13) {
                const int c = shared_c(tid);
                double s = 0;
                for (int k = 0; k < K; k++) s += sJ[k * 40 + c] * sJ[k * 40 + 18] + sJ[k * 40 + 20 + c] * sJ[k * 40 + 38];
                sWs[tid] = s;
            }
            for (int q = tid; q < K * 6; q += ACC_WARPS * 32) {
                const int k = q / 6, c = 6 + q % 6;
                sW[q] = sJ[k * 40 + c] * sJ[k * 40 + 18] + sJ[k * 40 + 20 + c] * sJ[k * 40 + 38];
            }
            __syncthreads();
            // persist what the back-substitution needs
            if (tid == 0) {
                w.lm_c[l] = c_l;
                w.lm_g[l] = g_l;
                w.lm_coli[l] = lay.col_pose(fi);
            }
            for (int k = tid; k < K; k += ACC_WARPS * 32) w.proj_colj[k0 + k] = lay.col_pose(sFj[k]);
            if (tid < 13) w.lm_ws[13 * (size_t)l + tid] = sWs[tid];
            for (int q = tid; q < K * 6; q += ACC_WARPS * 32) w.proj_wj[6 * (size_t)k0 + q] = sW[q];
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
            for (int A = tid; A < n_l; A += ACC_WARPS * 32) {
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
            // reduced-system contribution: every unordered local pair (A <= B) exactly once.  The pairs are independent and
            // cheap (a table look-up each way + at most two products): all threads of the CTA share them through a flat index
            // — a landmark seen from 10 frames has 2701 pairs, and walking them row by row with four warps left the CTA
            // latency-bound (18 us, measured); the shared 13 x 13 block J_s^T J_s is summed over the observations first.
            double *sH = sWs + 16 + ((max_obs + 1) >> 1);           // 91: J_s^T J_s, packed upper
            double *sWl = sH + 92;                                   // n_l: w of every local column
            int *sG = reinterpret_cast<int *>(sWl + 13 + 6 * max_obs);  // n_l: global column of every local column
            for (int q = tid; q < 91; q += ACC_WARPS * 32) {
                int a = 0, rem = q;
                while (rem >= 13 - a) {
                    rem -= 13 - a;
                    a++;
                }
                const int ca = shared_c(a), cb = shared_c(a + rem);
                double v = 0;
                for (int k = 0; k < K; k++) v += sJ[k * 40 + ca] * sJ[k * 40 + cb] + sJ[k * 40 + 20 + ca] * sJ[k * 40 + 20 + cb];
                sH[q] = v;
            }
            for (int A = tid; A < n_l; A += ACC_WARPS * 32) {
                sWl[A] = wval(A);
                sG[A] = gcol(A);
            }
            __syncthreads();
            // rows r and n_l - 1 - r of the upper triangle together have n_l + 1 entries: a rectangle of ceil(n_l / 2) such
            // double rows covers the triangle without idle iterations
            const int wrow = n_l + 1, nrow = (n_l + 1) >> 1, npair = nrow * wrow;
            const float inv_w = 1.0f / (float)wrow;
            for (int q = tid; q < npair; q += ACC_WARPS * 32) {
                int r = (int)((float)q * inv_w);
                r += (q - r * wrow >= wrow) - (q - r * wrow < 0);  // exact quotient (the float estimate is off by at most one)
                const int c = q - r * wrow, r2 = n_l - 1 - r;
                int A, B;
                if (c < n_l - r) {
                    A = r;
                    B = r + c;
                } else {
                    if (r2 == r) continue;  // middle row of an odd n_l: already covered
                    A = r2;
                    B = r2 + (c - (n_l - r));
                }
                double v;
                if (B < 13) {
                    v = sH[A * 13 - ((A * (A - 1)) >> 1) + (B - A)];
                } else {
                    const int kb = (B - 13) / 6, cb = 6 + (B - 13) % 6;
                    if (A < 13) {
                        const int ca = shared_c(A);
                        v = sJ[kb * 40 + ca] * sJ[kb * 40 + cb] + sJ[kb * 40 + 20 + ca] * sJ[kb * 40 + 20 + cb];
                    } else {
                        const int ka = (A - 13) / 6, ca = 6 + (A - 13) % 6;
                        v = (ka == kb) ? sJ[ka * 40 + ca] * sJ[ka * 40 + cb] + sJ[ka * 40 + 20 + ca] * sJ[ka * 40 + 20 + cb] : 0.0;
                    }
                }
                v -= sWl[A] * sWl[B] * inv_c;
                add_upper(w.S, D, sG[A], sG[B], v);
            }
        } else if (item >= g_L && item - g_L < P.n_imu) {
            // ===== one IMU factor: dense 15 x 30 block =====  (items [lay.L, g_L) are launch padding)
            const int k = item - g_L;
            if (P.imu[k].sum_dt > 10.0) return;
            const int fi = P.imu[k].frame_i;
            double *sI = sJ;  // 450 doubles fit (per_warp >= 49 * max_obs + 16 >= 450 is ensured by the launcher)
            for (int q = tid; q < 450; q += ACC_WARPS * 32) sI[q] = e.Ji[450 * (size_t)k + q] * w.scale[imu_col(lay, fi, q % 30)];
            double *sRi = sI + 450;
            for (int q = tid; q < 15; q += ACC_WARPS * 32) sRi[q] = e.ri[15 * (size_t)k + q];
            __syncthreads();
            if (tid < 30) {
                double g = 0, d2 = 0;
                for (int r = 0; r < 15; r++) {
                    const double j = sI[r * 30 + tid];
                    g += j * sRi[r];
                    d2 += j * j;
                }
                const int gc = imu_col(lay, fi, tid);
                atomicAdd(&w.grad[gc], g);
                atomicAdd(&w.rhs[gc], g);
                atomicAdd(&w.diag2[gc], d2);
            }
            for (int A = wid; A < 30; A += ACC_WARPS) {
                const int gA = imu_col(lay, fi, A);
                for (int B = A + lane; B < 30; B += 32) {
                    double v = 0;
                    for (int r = 0; r < 15; r++) v += sI[r * 30 + A] * sI[r * 30 + B];
                    add_upper(w.S, D, gA, imu_col(lay, fi, B), v);
                }
            }
        }
        CTA_PROF(g_acc_prof, 2);
        return;
    }
    // ===== prior CTAs: S += scale (J0^T J0) scale, gradient += scale J0^T r =====
    const DevPrior &p = P.prior;
    if (p.n <= 0) return;
    const int pj = blockIdx.x - n_lm_cta, ptid = pj * blockDim.x + threadIdx.x, pnt = PRIOR_CTAS * blockDim.x;
    int *pmap = reinterpret_cast<int *>(smem);
    for (int b = threadIdx.x; b < p.n_blocks; b += blockDim.x) {
        const int ls = lay.block_lsize(p.block_id[b]), c0 = lay.block_col(p.block_id[b]);
        for (int q = 0; q < ls; q++) pmap[p.block_idx[b] + q] = c0 + q;
    }
    __syncthreads();
    const int n = p.n;
    for (int q = ptid; q < n * n; q += pnt) {
        const int a = q / n, b = q % n;
        const int ga = pmap[a], gb = pmap[b];
        if (ga <= gb) {
            const double v = p.H0[q] * w.scale[ga] * w.scale[gb];
            if (v != 0.0) atomicAdd(&w.S[(size_t)ga * D + gb], v);
            if (a == b) atomicAdd(&w.diag2[ga], v);
        }
    }
    for (int a = threadIdx.x; a < n; a += blockDim.x) {  // CTA pj sums the rows i = pj, pj + PRIOR_CTAS, ...
        double g = 0;
        for (int i = pj; i < n; i += PRIOR_CTAS) g += p.J0[(size_t)i * n + a] * e.rm[i];
        g *= w.scale[pmap[a]];
        atomicAdd(&w.grad[pmap[a]], g);
        atomicAdd(&w.rhs[pmap[a]], g);
    }
    CTA_PROF(g_acc_prof, 2);
}

__global__ void solver_init_kernel(SolverCtl *ctl, const DevProblem *__restrict__ Pd) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    const lfvio_solver_opts &o = Pd->o;
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
    ctl->gnorm = ctl->gn_norm = ctl->gdot = 0.0;
    ctl->jj = 0.0;
    ctl->mu_solve = 1e-8;
    ctl->ticket2 = 0;
    for (int i = 0; i < 8; i++) ctl->acc[i] = 0.0;
}

}  // namespace
CTA_PROF_READER(lfvio_debug_acc_prof, g_acc_prof)

int launch_solver_init(SolverCtl *ctl, const DevProblem &P, cudaStream_t s) {
    solver_init_kernel<<<1, 32, 0, s>>>(ctl, P.self_dev);
    LFVIO_LAUNCH_CHECK();
    return LFVIO_OK;
}

int launch_jacobi_scale(const DevProblem &P, const EvalBuf &e0, const SolverScratch &w, cudaStream_t s) {
    const int N = P.lay.D() + P.cap_L;  // capacity: the launch geometry does not follow the landmark count
    LFVIO_CUDA_TRY(cudaMemsetAsync(w.scale, 0, sizeof(double) * N, s));
    // (g_nbp CTAs of 64 factors each in the evaluation: 40 entries per factor here)
    colnorm_kernel<<<std::max(1, std::min(cdiv(P.g_nbp * 64 * 40, CN_THREADS * 4), CN_MAX_CTAS)), CN_THREADS, 0, s>>>(P.self_dev, e0, w.scale);
    LFVIO_LAUNCH_CHECK();
    scale_finalize_kernel<<<cdiv(N, 256), 256, 0, s>>>(P.self_dev, w.scale);
    LFVIO_LAUNCH_CHECK();
    return LFVIO_OK;
}

// sJ 40 K | sR 2 K | sW 6 K | sWs 16 | sFj K ints | sH 92 | sWl 13 + 6 K | sG 13 + 6 K ints
static int acc_per_warp_doubles(int max_obs) { return max_obs * 40 + max_obs * 2 + max_obs * 6 + 16 + (max_obs + 1) / 2 + 92 + (13 + 6 * max_obs) + (13 + 6 * max_obs + 1) / 2 + 2; }

int launch_accumulate(const DevProblem &P, const EvalBuf &e0, const EvalBuf &e1, const SolverScratch &w, SolverCtl *ctl,
                      cudaStream_t s, bool clear_done) {
    // [S | rhs | diag2 | grad] must be zero on entry.  When the whole step runs in chol_kernel with S in shared memory
    // (step_zeroes_system), that kernel clears them as it consumes them; otherwise clear them here.
    if (!clear_done) LFVIO_CUDA_TRY(cudaMemsetAsync(w.S, 0, w.zero_bytes, s));
    int max_obs = std::max(P.lay.F - 1, 10);  // 49 * 10 + 16 = 506 doubles >= 465 needed by an IMU warp
    const int items = P.g_L + P.g_imu;  // landmark items padded to the bucket, then the IMU items
    const int n_lm_cta = items;
    size_t smem = sizeof(double) * (size_t)acc_per_warp_doubles(max_obs);
    smem = std::max(smem, sizeof(int) * (size_t)(15 * kMaxFrames + 8));
    LFVIO_TRY(ensure_dyn_smem(reinterpret_cast<const void *>(accumulate_kernel), 96 * 1024));
    LFVIO_CUDA_TRY(launch_pdl(accumulate_kernel, dim3(n_lm_cta + PRIOR_CTAS), dim3(ACC_WARPS * 32), smem, s, P.self_dev, e0, e1, w,
                              (const SolverCtl *)ctl, n_lm_cta, P.g_L, max_obs));
    LFVIO_LAUNCH_CHECK();
    return LFVIO_OK;
}

}  // namespace lfvio
