// K14 — construction of the next marginalisation prior on device.
//
// Reference: Estimator::optimization(), vins_estimator/src/estimator.cpp:833-939 (MARGIN_OLD) and :940-1005
// (MARGIN_SECOND_NEW); MarginalizationInfo::preMarginalize / marginalize / getParameterBlocks and
// ResidualBlockInfo::Evaluate, vins_estimator/src/factor/marginalization_factor.cpp:3-69,110-129,174-319.
//
// The reference builds a dense A = sum J^T J over [marginalised m | kept n] columns with 4 pthreads, inverts
// A_mm through a dense eigen-decomposition (eigenvalues <= 1e-8 dropped), forms the Schur complement and
// factors it back into (J0, r0) with a second eigen-decomposition.
// Here the structure of A_mm is used: it is [pose_0/speed-bias_0 (<= 15) | diagonal inverse-depth block], so
//   1. marg_accumulate_kernel: one warp per landmark hosted in the dropped frame eliminates its 1x1 pivot exactly
//      (pivots <= 1e-8 are dropped, which is what the pseudo-inverse does to a decoupled direction) and adds
//      sum_k J_k^T J_k - w w^T / c to the (m_p + n)^2 system T with fp64 atomics; the IMU factor and the previous
//      prior (J0^T J0 precomputed) are added by one warp / one CTA;
//   2. marg_finish_kernel (one CTA, matrices in shared memory): thresholded eigen-decomposition of the remaining
//      m_p x m_p block (m_p <= 15) by parallel cyclic Jacobi, Schur complement A' = T_rr - T_rm T_mm^+ T_mr,
//      second Jacobi eigen-decomposition of A' (n x n), J0 = sqrt(S) V^T, r0 = S^-1/2 V^T b'.
// This equals the reference whenever A_mm has no eigenvalue below 1e-8 other than decoupled landmark pivots
// (the reference itself carries a disabled assert that the spectrum is non-negative); the fp64 oracle implements
// the reference's dense route and the parity tests compare the invariants J0^T J0 and J0^T r0.
// Robust-loss correction: for CauchyLoss rho'' < 0, so ResidualBlockInfo::Evaluate (:37-68) scales residual and
// Jacobian by sqrt(rho') — exactly what evaluate_kernel already stored.
#include "backend.cuh"
#include "marg.cuh"

namespace lfvio {

namespace {

constexpr int MA_WARPS = 4;
constexpr int MF_THREADS = 512;
constexpr double kEps = 1e-8;  // MarginalizationInfo::eps (marginalization_factor.h:70)

__device__ __forceinline__ void add_sym(double *T, int ld, int a, int b, double v) {
    if (v == 0.0) return;
    atomicAdd(&T[(size_t)a * ld + b], v);
    if (a != b) atomicAdd(&T[(size_t)b * ld + a], v);
}

__global__ void __launch_bounds__(MA_WARPS * 32) marg_accumulate_kernel(DevProblem P, EvalBuf e0, EvalBuf e1,
                                                                       const SolverCtl *ctl, MargPlan plan, double *T,
                                                                       double *bvec, int n_item_cta, int max_obs) {
    extern __shared__ double smem[];
    const EvalBuf &e = (ctl && ctl->cur) ? e1 : e0;
    const StateLayout lay = P.lay;
    const int ld = plan.dim;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if ((int)blockIdx.x < n_item_cta) {
        const int item = blockIdx.x * MA_WARPS + wid;
        const int per_warp = max_obs * 40 + max_obs * 2 + max_obs * 6 + 16 + max_obs;
        double *sJ = smem + (size_t)wid * per_warp;
        double *sR = sJ + max_obs * 40;
        double *sW = sR + max_obs * 2;
        double *sWs = sW + max_obs * 6;
        int *sFj = reinterpret_cast<int *>(sWs + 16);
        if (item < plan.n_lm) {
            // ===== one landmark hosted in the dropped frame: eliminate its inverse depth =====
            const int l = plan.lm_list[item];
            const int k0 = P.lm_first[l], K = min(P.lm_first[l + 1] - k0, max_obs);
            if (K <= 0) return;
            const int fi = P.proj[k0].frame_i;
            for (int k = lane; k < K; k += 32) sFj[k] = P.proj[k0 + k].frame_j;
            for (int q = lane; q < K * 40; q += 32) sJ[q] = e.Jp[40 * (size_t)k0 + q];
            for (int q = lane; q < K * 2; q += 32) sR[q] = e.rp[2 * (size_t)k0 + q];
            __syncwarp();
            double c_l = 0, g_l = 0;
            for (int q = lane; q < K * 2; q += 32) {
                const double jl = sJ[(q >> 1) * 40 + (q & 1) * 20 + 18];
                c_l += jl * jl;
                g_l += jl * sR[q];
            }
            c_l = warp_sum(c_l);
            g_l = warp_sum(g_l);
            const double inv_c = c_l > kEps ? 1.0 / c_l : 0.0;
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
            const int n_l = 13 + 6 * K;
            const bool has_td = P.o.estimate_td != 0;
            auto mcol = [&](int A) {  // column in the marginalisation ordering, -1 if the block is absent
                if (A < 6) return plan.col[2 * fi] + A;
                if (A < 12) return plan.col[2 * lay.F] + (A - 6);
                if (A == 12) return has_td ? plan.col[2 * lay.F + 1] : -1;
                return plan.col[2 * sFj[(A - 13) / 6]] + (A - 13) % 6;
            };
            auto wval = [&](int A) { return A < 13 ? sWs[A] : sW[A - 13]; };
            for (int A = lane; A < n_l; A += 32) {
                const int ca = mcol(A);
                if (ca < 0) continue;
                double g = 0;
                if (A < 13) {
                    const int c = shared_c(A);
                    for (int k = 0; k < K; k++) g += sJ[k * 40 + c] * sR[2 * k] + sJ[k * 40 + 20 + c] * sR[2 * k + 1];
                } else {
                    const int k = (A - 13) / 6, c = 6 + (A - 13) % 6;
                    g = sJ[k * 40 + c] * sR[2 * k] + sJ[k * 40 + 20 + c] * sR[2 * k + 1];
                }
                const double v = g - wval(A) * g_l * inv_c;
                if (v != 0.0) atomicAdd(&bvec[ca], v);
            }
            for (int A = 0; A < n_l; A++) {
                const int ca = mcol(A);
                if (ca < 0) continue;
                const double wA = wval(A);
                for (int B = A + lane; B < n_l; B += 32) {
                    const int cb = mcol(B);
                    if (cb < 0) continue;
                    double v = 0;
                    if (A < 13) {
                        const int la = shared_c(A);
                        if (B < 13) {
                            const int lb = shared_c(B);
                            for (int k = 0; k < K; k++) v += sJ[k * 40 + la] * sJ[k * 40 + lb] + sJ[k * 40 + 20 + la] * sJ[k * 40 + 20 + lb];
                        } else {
                            const int k = (B - 13) / 6, lb = 6 + (B - 13) % 6;
                            v = sJ[k * 40 + la] * sJ[k * 40 + lb] + sJ[k * 40 + 20 + la] * sJ[k * 40 + 20 + lb];
                        }
                    } else {
                        const int ka = (A - 13) / 6, kb = (B - 13) / 6;
                        if (ka == kb) {
                            const int la = 6 + (A - 13) % 6, lb = 6 + (B - 13) % 6;
                            v = sJ[ka * 40 + la] * sJ[ka * 40 + lb] + sJ[ka * 40 + 20 + la] * sJ[ka * 40 + 20 + lb];
                        }
                    }
                    v -= wA * wval(B) * inv_c;
                    add_sym(T, ld, ca, cb, v);
                }
            }
        } else if (item == plan.n_lm && plan.imu_index >= 0) {
            // ===== IMUFactor(pre_integrations[1]) between frames 0 and 1 =====
            const int k = plan.imu_index;
            double *sI = sJ;
            for (int q = lane; q < 450; q += 32) sI[q] = e.Ji[450 * (size_t)k + q];
            double *sRi = sI + 450;
            for (int q = lane; q < 15; q += 32) sRi[q] = e.ri[15 * (size_t)k + q];
            __syncwarp();
            auto icol = [&](int c) {
                if (c < 6) return plan.col[0] + c;
                if (c < 15) return plan.col[1] + (c - 6);
                if (c < 21) return plan.col[2] + (c - 15);
                return plan.col[3] + (c - 21);
            };
            if (lane < 30) {
                double g = 0;
                for (int r = 0; r < 15; r++) g += sI[r * 30 + lane] * sRi[r];
                atomicAdd(&bvec[icol(lane)], g);
            }
            for (int A = 0; A < 30; A++)
                for (int B = A + lane; B < 30; B += 32) {
                    double v = 0;
                    for (int r = 0; r < 15; r++) v += sI[r * 30 + A] * sI[r * 30 + B];
                    add_sym(T, ld, icol(A), icol(B), v);
                }
        }
        return;
    }
    // ===== previous prior: T += P (J0^T J0) P^T, b += P J0^T r(x) =====
    if (!plan.use_prior) return;
    const DevPrior &p = P.prior;
    int *pmap = reinterpret_cast<int *>(smem);
    for (int b = threadIdx.x; b < p.n_blocks; b += blockDim.x) {
        const int ls = lay.block_lsize(p.block_id[b]), c0 = plan.col[p.block_id[b]];
        for (int q = 0; q < ls; q++) pmap[p.block_idx[b] + q] = c0 + q;
    }
    __syncthreads();
    const int n = p.n;
    for (int q = threadIdx.x; q < n * n; q += blockDim.x) {
        const int a = q / n, b = q % n;
        const double v = p.H0[q];
        if (v != 0.0) atomicAdd(&T[(size_t)pmap[a] * ld + pmap[b]], v);
    }
    for (int a = threadIdx.x; a < n; a += blockDim.x) {
        double g = 0;
        for (int i = 0; i < n; i++) g += p.J0[(size_t)i * n + a] * e.rm[i];
        atomicAdd(&bvec[pmap[a]], g);
    }
}

// Parallel cyclic Jacobi (round-robin ordering) on a symmetric n x n matrix A (leading dimension ld) held in
// shared or global memory; V (n x n, ld) receives the eigenvectors as columns; w <- diagonal.  Whole CTA calls.
__device__ void jacobi_eig(double *A, double *V, int n, int ld, double *s_c, double *s_s, int *s_p, int *s_q, double *s_red) {
    const int tid = threadIdx.x, nt = blockDim.x;
    for (int q = tid; q < n * n; q += nt) V[(q / n) * ld + (q % n)] = (q / n == q % n) ? 1.0 : 0.0;
    __syncthreads();
    if (n < 2) return;
    const int m = (n + 1) & ~1;  // players (one dummy when n is odd)
    const int half = m / 2;
    for (int sweep = 0; sweep < 30; sweep++) {
        // convergence: off-diagonal energy relative to the diagonal
        double off = 0, dg = 0;
        for (int q = tid; q < n * n; q += nt) {
            const int r = q / n, c = q % n;
            const double v = A[r * ld + c];
            if (r == c) dg += v * v;
            else off += v * v;
        }
        off = warp_sum(off);
        dg = warp_sum(dg);
        __syncthreads();
        if ((tid & 31) == 0) {
            s_red[tid >> 5] = off;
            s_red[32 + (tid >> 5)] = dg;
        }
        __syncthreads();
        double toff = 0, tdg = 0;
        for (int k = 0; k < (nt >> 5); k++) {
            toff += s_red[k];
            tdg += s_red[32 + k];
        }
        __syncthreads();
        if (toff <= 1e-30 * (tdg + toff) || toff == 0.0) break;
        for (int round = 0; round < m - 1; round++) {
            // round-robin pairing: player 0 fixed, the others rotate
            if (tid < half) {
                int a, b;  // circle method: player m-1 is fixed, the others rotate
                if (tid == 0) {
                    a = m - 1;
                    b = round % (m - 1);
                } else {
                    a = (round + tid) % (m - 1);
                    b = (round - tid + (m - 1)) % (m - 1);
                }
                int p = min(a, b), q = max(a, b);
                double c = 1.0, s = 0.0;
                if (q < n) {
                    const double apq = A[p * ld + q];
                    if (apq != 0.0) {
                        const double theta = (A[q * ld + q] - A[p * ld + p]) / (2.0 * apq);
                        const double t = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
                        c = 1.0 / sqrt(t * t + 1.0);
                        s = t * c;
                    }
                } else {
                    q = -1;  // dummy player: no rotation
                }
                s_p[tid] = p;
                s_q[tid] = q;
                s_c[tid] = c;
                s_s[tid] = s;
            }
            __syncthreads();
            // columns: A <- A J, V <- V J
            for (int e = tid; e < half * n; e += nt) {
                const int pr = e / n, k = e % n;
                const int p = s_p[pr], q = s_q[pr];
                if (q < 0) continue;
                const double c = s_c[pr], s = s_s[pr];
                const double akp = A[k * ld + p], akq = A[k * ld + q];
                A[k * ld + p] = c * akp - s * akq;
                A[k * ld + q] = s * akp + c * akq;
                const double vkp = V[k * ld + p], vkq = V[k * ld + q];
                V[k * ld + p] = c * vkp - s * vkq;
                V[k * ld + q] = s * vkp + c * vkq;
            }
            __syncthreads();
            // rows: A <- J^T A
            for (int e = tid; e < half * n; e += nt) {
                const int pr = e / n, k = e % n;
                const int p = s_p[pr], q = s_q[pr];
                if (q < 0) continue;
                const double c = s_c[pr], s = s_s[pr];
                const double apk = A[p * ld + k], aqk = A[q * ld + k];
                A[p * ld + k] = c * apk - s * aqk;
                A[q * ld + k] = s * apk + c * aqk;
            }
            __syncthreads();
        }
    }
}

__global__ void __launch_bounds__(MF_THREADS) marg_finish_kernel(DevProblem P, const double *x0, const double *x1,
                                                                const SolverCtl *ctl, MargPlan plan, const double *T,
                                                                const double *bvec, double *work, double *J0_out,
                                                                double *r0_out, double *x0_out) {
    extern __shared__ double smem[];
    __shared__ double s_c[64], s_s[64], s_red[64];
    __shared__ int s_p[64], s_q[64];
    const int tid = threadIdx.x, nt = blockDim.x;
    const int mp = plan.m_p, n = plan.n, ld = plan.dim;
    const double *x = (ctl && ctl->cur) ? x1 : x0;
    // kept-block data copies (keep_block_data), global sizes
    for (int b = tid; b < plan.n_keep; b += nt) {
        const int id = plan.keep_id[b], gs = P.lay.block_gsize(id), so = P.lay.block_state(id);
        for (int k = 0; k < gs; k++) x0_out[plan.keep_x0[b] + k] = x[so + k];
    }
    // workspace: [Amm mp*mp | Vm mp*mp | wm mp | Ar n*n | Vr n*n | br n | tmp mp*n]
    double *base = plan.use_smem ? smem : work;
    double *Amm = base;
    double *Vm = Amm + mp * mp;
    double *Minv = Vm + mp * mp;
    double *Ar = Minv + mp * mp;
    double *Vr = Ar + n * n;
    double *br = Vr + n * n;
    double *Tm = br + n;  // n x mp : T_rm * Amm^+
    for (int q = tid; q < mp * mp; q += nt) {
        const int r = q / mp, c = q % mp;
        Amm[q] = 0.5 * (T[(size_t)r * ld + c] + T[(size_t)c * ld + r]);
    }
    __syncthreads();
    jacobi_eig(Amm, Vm, mp, mp, s_c, s_s, s_p, s_q, s_red);
    __syncthreads();
    // Amm^+ = V diag(1/w > eps) V^T
    for (int q = tid; q < mp * mp; q += nt) {
        const int r = q / mp, c = q % mp;
        double s = 0;
        for (int k = 0; k < mp; k++) {
            const double wk = Amm[k * mp + k];
            if (wk > kEps) s += Vm[r * mp + k] * Vm[c * mp + k] / wk;
        }
        Minv[q] = s;
    }
    __syncthreads();
    for (int q = tid; q < n * mp; q += nt) {
        const int i = q / mp, j = q % mp;
        double s = 0;
        for (int k = 0; k < mp; k++) s += T[(size_t)(mp + i) * ld + k] * Minv[k * mp + j];
        Tm[q] = s;
    }
    __syncthreads();
    for (int q = tid; q < n * n; q += nt) {
        const int i = q / n, j = q % n;
        double s = T[(size_t)(mp + i) * ld + mp + j];
        for (int k = 0; k < mp; k++) s -= Tm[i * mp + k] * T[(size_t)k * ld + mp + j];
        Ar[q] = s;
    }
    for (int i = tid; i < n; i += nt) {
        double s = bvec[mp + i];
        for (int k = 0; k < mp; k++) s -= Tm[i * mp + k] * bvec[k];
        br[i] = s;
    }
    __syncthreads();
    // symmetrise against atomics round-off before the rotation sweeps
    for (int q = tid; q < n * n; q += nt) {
        const int i = q / n, j = q % n;
        if (i < j) {
            const double v = 0.5 * (Ar[i * n + j] + Ar[j * n + i]);
            Ar[i * n + j] = v;
            Ar[j * n + i] = v;
        }
    }
    __syncthreads();
    jacobi_eig(Ar, Vr, n, n, s_c, s_s, s_p, s_q, s_red);
    __syncthreads();
    // J0 = sqrt(S) V^T, r0 = S^-1/2 V^T b'   (eigenvalues <= eps dropped)
    for (int q = tid; q < n * n; q += nt) {
        const int k = q / n, i = q % n;
        const double wk = Ar[k * n + k];
        J0_out[q] = wk > kEps ? sqrt(wk) * Vr[i * n + k] : 0.0;
    }
    for (int k = tid; k < n; k += nt) {
        const double wk = Ar[k * n + k];
        double vb = 0;
        for (int i = 0; i < n; i++) vb += Vr[i * n + k] * br[i];
        r0_out[k] = wk > kEps ? vb / sqrt(wk) : 0.0;
    }
}

}  // namespace

size_t marg_finish_workspace_doubles(int m_p, int n) { return (size_t)3 * m_p * m_p + 2 * (size_t)n * n + n + (size_t)n * m_p; }

int launch_marginalize(const DevProblem &P, const EvalBuf &e0, const EvalBuf &e1, const double *x0, const double *x1,
                       const SolverCtl *ctl, MargPlan &plan, const MargBuffers &mb, cudaStream_t s) {
    const int dim = plan.dim;
    LFVIO_CUDA_TRY(cudaMemsetAsync(mb.T, 0, sizeof(double) * ((size_t)dim * dim + dim), s));
    const int max_obs = std::max(P.lay.F - 1, 10);
    const int items = plan.n_lm + 1;
    const int n_item_cta = cdiv(items, MA_WARPS);
    const int per_warp = max_obs * 40 + max_obs * 2 + max_obs * 6 + 16 + max_obs;
    size_t smem = sizeof(double) * (size_t)per_warp * MA_WARPS;
    smem = std::max(smem, sizeof(int) * (size_t)(15 * kMaxFrames + 8));
    static bool a1 = false, a2 = false;
    if (!a1) {
        LFVIO_CUDA_TRY(cudaFuncSetAttribute(marg_accumulate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
        a1 = true;
    }
    marg_accumulate_kernel<<<n_item_cta + 1, MA_WARPS * 32, smem, s>>>(P, e0, e1, ctl, plan, mb.T, mb.T + (size_t)dim * dim,
                                                                      n_item_cta, max_obs);
    LFVIO_LAUNCH_CHECK();
    const size_t ws = sizeof(double) * marg_finish_workspace_doubles(plan.m_p, plan.n);
    plan.use_smem = ws <= 200 * 1024 ? 1 : 0;
    if (!a2) {
        LFVIO_CUDA_TRY(cudaFuncSetAttribute(marg_finish_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        a2 = true;
    }
    marg_finish_kernel<<<1, MF_THREADS, plan.use_smem ? ws : 0, s>>>(P, x0, x1, ctl, plan, mb.T, mb.T + (size_t)dim * dim, mb.work,
                                                                     mb.J0, mb.r0, mb.x0);
    LFVIO_LAUNCH_CHECK();
    return LFVIO_OK;
}

}  // namespace lfvio
