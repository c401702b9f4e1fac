// K17-K22 — the estimator initialisation on device: everything Estimator::initialStructure
// (vins_estimator/src/estimator.cpp:221-376) and visualInitialAlign (:378-455) compute, behind stage-level C-ABI entries.
//
//   gyro_bias_kernel        solveGyroscopeBias            initial/initial_aligment.cpp:3-29      one warp, lane per IMU interval
//   alignment_kernel        LinearAlignment+RefineGravity initial/initial_aligment.cpp:56-205    one CTA: pair blocks scattered with
//                                                          fp64 atomics, in-place Cholesky of the (3n+4)^2 normal equations, the four
//                                                          gravity refinements without leaving the kernel
//   recover_pose_kernel     recoverPose                   initial/solve_5pts.cpp:392-536         4 candidates x n cheirality tests
//   sfm_*_kernel            GlobalSFM::construct          initial/initial_sfm.cpp:127-316        the whole chain (triangulate / PnP per
//                                                          frame / full BA) is ENQUEUED on one stream, state and failure flag live on
//                                                          the device, one synchronisation at the end
//   epnp_kernel             PnpSolver::compute_pose       pnp_solver.cpp:275-338                 one CTA per frame (all non-keyframes of
//                                                          estimator.cpp:320-357 in one launch)
//   sfm_ba_kernel           ceres::Solve of :230-297      Ceres 1.12 LM + DENSE_SCHUR restated   one CTA, reduced camera system in
//                                                          shared memory (see ba_core.cuh)
// The small dense cores are the __host__ __device__ functions of init_math.cuh / ba_core.cuh (also compiled for the host by
// the test shim, CPU-checked against oracle/initial.py); this file adds the parallel reductions, the atomics and the control.
// These run once per start-up / restart, on tens of frames and a few hundred features: they are latency-sized, not
// bandwidth-sized — the point of having them on the device is that initialisation needs no CPU numerics.
#include <mutex>
#include <vector>

#include "ba_core.cuh"
#include "frontend.cuh"

namespace lfvio {

namespace {

constexpr int NT = 256;      // threads of the single-CTA kernels
constexpr int BA_NT = 512;

template <class T>
struct DBuf {
    T *p = nullptr;
    ~DBuf() {
        if (p) cudaFree(p);
    }
    int put(const T *h, size_t n) {
        LFVIO_CUDA_TRY(cudaMalloc(&p, sizeof(T) * (n ? n : 1)));
        if (h && n) LFVIO_CUDA_TRY(cudaMemcpy(p, h, sizeof(T) * n, cudaMemcpyHostToDevice));
        else LFVIO_CUDA_TRY(cudaMemset(p, 0, sizeof(T) * (n ? n : 1)));
        return LFVIO_OK;
    }
    int get(T *h, size_t n) const {
        LFVIO_CUDA_TRY(cudaMemcpy(h, p, sizeof(T) * n, cudaMemcpyDeviceToHost));
        return LFVIO_OK;
    }
};

// sum of K values over the CTA; every thread receives the totals.  red: K * (blockDim / 32) doubles of shared memory.
template <int K>
__device__ void block_sum(double *v, double *red) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
#pragma unroll
    for (int k = 0; k < K; k++) v[k] = warp_sum(v[k]);
    __syncthreads();  // red may still be read from a previous call
    if (lane == 0)
        for (int k = 0; k < K; k++) red[k * nw + w] = v[k];
    __syncthreads();
    for (int k = 0; k < K; k++) {
        double s = 0;
        for (int i = 0; i < nw; i++) s += red[k * nw + i];
        v[k] = s;
    }
}
__device__ double block_max(double v, double *red) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    __syncthreads();
    if (lane == 0) red[w] = v;
    __syncthreads();
    double s = red[0];
    for (int i = 1; i < nw; i++) s = fmax(s, red[i]);
    return s;
}

// In-place Cholesky (lower triangle, row-major, leading dimension ld) of an n x n matrix in shared or global memory by the
// whole CTA, right-looking; *ok (shared) is cleared on a non-positive pivot.  Then solve A x = b in place in b.
__device__ void cta_cholesky_solve(double *A, int ld, int n, double *b, int *ok) {
    const int tid = threadIdx.x, nt = blockDim.x;
    for (int j = 0; j < n; j++) {
        if (tid == 0) {
            const double d = A[j * ld + j];
            if (!(d > 0.0)) *ok = 0;
            else A[j * ld + j] = sqrt(d);
        }
        __syncthreads();
        if (!*ok) return;
        const double inv = 1.0 / A[j * ld + j];
        for (int i = j + 1 + tid; i < n; i += nt) A[i * ld + j] *= inv;
        __syncthreads();
        const int m = n - j - 1;
        for (int idx = tid; idx < m * m; idx += nt) {
            const int ii = idx / m, kk = idx - ii * m;
            if (kk <= ii) A[(j + 1 + ii) * ld + j + 1 + kk] -= A[(j + 1 + ii) * ld + j] * A[(j + 1 + kk) * ld + j];
        }
        __syncthreads();
    }
    for (int j = 0; j < n; j++) {  // L y = b
        if (tid == 0) b[j] /= A[j * ld + j];
        __syncthreads();
        const double yj = b[j];
        for (int i = j + 1 + tid; i < n; i += nt) b[i] -= A[i * ld + j] * yj;
        __syncthreads();
    }
    for (int j = n - 1; j >= 0; j--) {  // L^T x = y
        if (tid == 0) b[j] /= A[j * ld + j];
        __syncthreads();
        const double xj = b[j];
        for (int i = tid; i < j; i += nt) b[i] -= A[j * ld + i] * xj;
        __syncthreads();
    }
}

// im::jacobi_sym on one warp (all 32 lanes call; n <= 32; A, V in shared memory): the same rotation sequence and formulas,
// with the n independent element updates of each half-rotation spread over the lanes.
__device__ void jacobi_sym_warp(int n, double *A, double *V) {
    const int lane = threadIdx.x & 31;
    for (int i = lane; i < n * n; i += 32) V[i] = (i / n == i % n) ? 1.0 : 0.0;
    __syncwarp();
    for (int sweep = 0; sweep < 40; sweep++) {
        double off = 0, dia = 0;
        if (lane < n) {
            dia = fabs(A[lane * n + lane]);
            for (int j = lane + 1; j < n; j++) off += fabs(A[lane * n + j]);
        }
        off = warp_sum(off);
        dia = warp_sum(dia);
        if (off <= 1e-300 || off <= 1e-18 * dia) break;
        for (int p = 0; p < n - 1; p++)
            for (int q = p + 1; q < n; q++) {
                const double apq = A[p * n + q];
                if (apq == 0.0) continue;  // (uniform: every lane reads the same element)
                const double theta = (A[q * n + q] - A[p * n + p]) / (2.0 * apq);
                const double t = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
                const double c = 1.0 / sqrt(t * t + 1.0), s = t * c;
                __syncwarp();
                if (lane < n) {  // A <- A J and V <- V J: row `lane`
                    const int k = lane;
                    const double akp = A[k * n + p], akq = A[k * n + q];
                    A[k * n + p] = c * akp - s * akq;
                    A[k * n + q] = s * akp + c * akq;
                    const double vkp = V[k * n + p], vkq = V[k * n + q];
                    V[k * n + p] = c * vkp - s * vkq;
                    V[k * n + q] = s * vkp + c * vkq;
                }
                __syncwarp();
                if (lane < n) {  // A <- J^T A: column `lane`
                    const int k = lane;
                    const double apk = A[p * n + k], aqk = A[q * n + k];
                    A[p * n + k] = c * apk - s * aqk;
                    A[q * n + k] = s * apk + c * aqk;
                }
                __syncwarp();
            }
    }
    __syncwarp();
}

// ---------------------------------------------------------------------------------------------------------------
// solveGyroscopeBias
// ---------------------------------------------------------------------------------------------------------------
__global__ void gyro_bias_kernel(const double *Rs, const double *dq, const double *Jrg, int n, double *delta_bg) {
    const int lane = threadIdx.x;
    double acc[12];
    for (int k = 0; k < 12; k++) acc[k] = 0;
    for (int i = lane; i < n - 1; i += 32) {
        double a[9], b[3];
        im::gyro_bias_pair(Rs + 9 * i, Rs + 9 * (i + 1), dq + 4 * i, Jrg + 9 * i, a, b);
        for (int k = 0; k < 9; k++) acc[k] += a[k];
        for (int k = 0; k < 3; k++) acc[9 + k] += b[k];
    }
    for (int k = 0; k < 12; k++) acc[k] = warp_sum(acc[k]);
    if (lane == 0) {  // A.ldlt().solve(b) of a 3 x 3 system
        double inv[9];
        im::inv3(acc, inv);
        im::mat3_vec(inv, acc + 9, delta_bg);
    }
}

// ---------------------------------------------------------------------------------------------------------------
// LinearAlignment + RefineGravity
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(NT) alignment_kernel(const double *Rs, const double *Ts, const double *dp, const double *dv,
                                                        const double *dts, int n, double t0, double t1, double t2, double g_norm,
                                                        double *A, double *bvec, double *Aacc, double *bacc, double *g_out,
                                                        double *x_out, int *ok_out) {
    __shared__ double s_gcols[9], s_gfixed[3], s_tic[3], s_g0[3];
    __shared__ int s_ok, s_stop;
    const int tid = threadIdx.x;
    if (tid == 0) {
        s_tic[0] = t0, s_tic[1] = t1, s_tic[2] = t2;
        for (int i = 0; i < 9; i++) s_gcols[i] = (i % 4 == 0) ? 1.0 : 0.0;
        s_gfixed[0] = s_gfixed[1] = s_gfixed[2] = 0.0;
        s_ok = 1, s_stop = 0;
    }
    __syncthreads();
    // pass 0 = LinearAlignment (k = 3), passes 1..4 = RefineGravity (k = 2)
    for (int pass = 0; pass < 5; pass++) {
        const int k = pass == 0 ? 3 : 2, kk = k + 1, ns = 3 * n + kk;
        // Aacc / bacc are the reference's A and b: LinearAlignment zeroes its own (:137-140); RefineGravity zeroes them ONCE
        // before its four passes (:62-65), so every later pass adds its blocks to the already scaled sum and scales again
        // (A = A * 1000 inside the loop, :119-120) — the first linearisation dominates the later solves. Kept as is.
        if (pass <= 1) {
            for (int i = tid; i < ns * ns; i += NT) Aacc[i] = 0.0;
            for (int i = tid; i < ns; i += NT) bacc[i] = 0.0;
        }
        __syncthreads();
        for (int i = tid; i < n - 1; i += NT) {
            double tA[60], tb[6];
            im::alignment_pair(Rs + 9 * i, Rs + 9 * (i + 1), Ts + 3 * i, Ts + 3 * (i + 1), dp + 3 * i, dv + 3 * i, dts[i], s_tic, s_gcols, k,
                               s_gfixed, tA, tb);
            const int w = 6 + kk;
            for (int a = 0; a < w; a++) {
                const int ga = a < 6 ? 3 * i + a : ns - kk + (a - 6);
                double rb = 0;
                for (int r = 0; r < 6; r++) rb += tA[10 * r + a] * tb[r];
                atomicAdd(&bacc[ga], rb);
                for (int c = 0; c <= a; c++) {  // lower triangle only (the Cholesky reads nothing else)
                    const int gc = c < 6 ? 3 * i + c : ns - kk + (c - 6);
                    double s = 0;
                    for (int r = 0; r < 6; r++) s += tA[10 * r + a] * tA[10 * r + c];
                    atomicAdd(&Aacc[ga * ns + gc], s);
                }
            }
        }
        __syncthreads();
        for (int i = tid; i < ns * ns; i += NT) A[i] = Aacc[i] = Aacc[i] * 1000.0;  // the factorisation below works in place on A
        for (int i = tid; i < ns; i += NT) bvec[i] = bacc[i] = bacc[i] * 1000.0;
        __syncthreads();
        cta_cholesky_solve(A, ns, ns, bvec, &s_ok);
        __syncthreads();
        if (tid == 0) {
            if (!s_ok) s_stop = 1;
            else if (pass == 0) {
                const double s = bvec[ns - 1] / 100.0;
                const double g[3] = {bvec[ns - 4], bvec[ns - 3], bvec[ns - 2]};
                const double gn = sqrt(g[0] * g[0] + g[1] * g[1] + g[2] * g[2]);
                if (fabs(gn - g_norm) > 1.0 || s < 0) {  // initial_aligment.cpp:186-194
                    s_ok = 0, s_stop = 1;
                    for (int q = 0; q < 3; q++) g_out[q] = g[q];
                } else {
                    for (int q = 0; q < 3; q++) s_g0[q] = g[q] / gn * g_norm;  // RefineGravity :58
                }
            } else {
                const double dg[2] = {bvec[ns - 3], bvec[ns - 2]};  // x.segment<2>(n_state - 3)
                double g[3];
                for (int q = 0; q < 3; q++) g[q] = s_g0[q] + s_gcols[2 * q] * dg[0] + s_gcols[2 * q + 1] * dg[1];
                const double gn = sqrt(g[0] * g[0] + g[1] * g[1] + g[2] * g[2]);
                for (int q = 0; q < 3; q++) s_g0[q] = g[q] / gn * g_norm;
            }
            if (!s_stop && pass < 4) {  // next RefineGravity pass linearises at the current g0
                im::tangent_basis(s_g0, s_gcols);
                for (int q = 0; q < 3; q++) s_gfixed[q] = s_g0[q];
            }
        }
        __syncthreads();
        if (s_stop || pass == 4) {
            for (int i = tid; i < ns; i += NT) x_out[i] = bvec[i];
            break;
        }
    }
    __syncthreads();
    if (tid == 0) {
        if (s_ok) {
            const int ns = 3 * n + 3;
            const double s = x_out[ns - 1] / 100.0;  // :199-200
            x_out[ns - 1] = s;
            for (int q = 0; q < 3; q++) g_out[q] = s_g0[q];
            if (s < 0.0) s_ok = 0;
        }
        *ok_out = s_ok;
    }
}

// ---------------------------------------------------------------------------------------------------------------
// recoverPose + the (repaired) tail of solveRelativeRT
// ---------------------------------------------------------------------------------------------------------------
// mode 0: recoverPose + the tail of solveRelativeRT.  mode 1: InitialEXRotation::solveRelativeR (initial_ex_rotation.cpp:
// 220-289): the same four cheirality counts (testTriangulation :291-318), rotation chosen by `ratio1 > ratio2 ? R1 : R2`
// where the reference's R1 (built with the transposed W, :312-317) is decompose_essential's R2; output = its transpose.
__global__ void __launch_bounds__(NT) recover_pose_kernel(const double *E, const double3 *b1, const double3 *b2, int n, int mode,
                                                           double *rotation, double *translation, int *inlier_cnt) {
    __shared__ double s_R[2][9], s_t[2][3];
    __shared__ int s_cnt[4];
    if (threadIdx.x == 0) {
        im::decompose_essential(E, s_R[0], s_R[1], s_t[0]);
        for (int k = 0; k < 3; k++) s_t[1][k] = -s_t[0][k];
        for (int k = 0; k < 4; k++) s_cnt[k] = 0;
    }
    __syncthreads();
    for (int idx = threadIdx.x; idx < 4 * n; idx += NT) {
        const int c = idx / n, i = idx - c * n;
        const double p[3] = {b1[i].x, b1[i].y, b1[i].z}, q[3] = {b2[i].x, b2[i].y, b2[i].z};
        if (im::cheirality(s_R[c & 1], s_t[c >> 1], p, q)) atomicAdd(&s_cnt[c], 1);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        const int g1 = s_cnt[0], g2 = s_cnt[1], g3 = s_cnt[2], g4 = s_cnt[3];
        if (mode == 1) {
            const int ratio1 = max(g2, g4), ratio2 = max(g1, g3);  // counts of the reference's R1 (= s_R[1]) and R2 (= s_R[0])
            const double *R = ratio1 > ratio2 ? s_R[1] : s_R[0];
            for (int r = 0; r < 3; r++)
                for (int c = 0; c < 3; c++) rotation[3 * r + c] = R[3 * c + r];
            *inlier_cnt = max(ratio1, ratio2);
            return;
        }
        int k;
        if (g1 >= g2 && g1 >= g3 && g1 >= g4) k = 0;
        else if (g2 >= g1 && g2 >= g3 && g2 >= g4) k = 1;
        else if (g3 >= g1 && g3 >= g2 && g3 >= g4) k = 2;
        else k = 3;
        const double *R = s_R[k & 1], *t = s_t[k >> 1];
        double Rt_t[3];
        im::mat3_tvec(R, t, Rt_t);
        for (int r = 0; r < 3; r++) {
            for (int c = 0; c < 3; c++) rotation[3 * r + c] = R[3 * c + r];  // Rotation = R^T
            translation[r] = -Rt_t[r];                                        // Translation = -R^T t
        }
        *inlier_cnt = s_cnt[k];
    }
}

// ---------------------------------------------------------------------------------------------------------------
// EPnP by one CTA
// ---------------------------------------------------------------------------------------------------------------
struct EpnpShared {
    double mean[3], cws[12], CCi[9], MtM[144], V[144], ccs[12], pc0[3], R[9], T[3], bestR[9], bestT[3];
    double sgn, best;
    im::EpnpCore core;
    double red[16 * (NT / 32)];
};

// pws, us: n x 3 (global); alphas: n x 4 global scratch.  All threads of the CTA call this; thread 0 ends up with the
// result in sh.bestR / sh.bestT / sh.best (visible to all after the final barrier).
__device__ void epnp_block(int n, const double *pws, const double *us, double *alphas, EpnpShared &sh) {
    const int tid = threadIdx.x;
    double v[9];
    v[0] = v[1] = v[2] = 0;
    for (int i = tid; i < n; i += NT)
        for (int k = 0; k < 3; k++) v[k] += pws[3 * i + k];
    block_sum<3>(v, sh.red);
    if (tid == 0)
        for (int k = 0; k < 3; k++) sh.mean[k] = v[k] / n;
    __syncthreads();
    for (int k = 0; k < 9; k++) v[k] = 0;
    for (int i = tid; i < n; i += NT)
        for (int a = 0; a < 3; a++)
            for (int b = 0; b < 3; b++) v[3 * a + b] += (pws[3 * i + a] - sh.mean[a]) * (pws[3 * i + b] - sh.mean[b]);
    block_sum<9>(v, sh.red);
    if (tid == 0) {
        im::epnp_control_points(n, sh.mean, v, sh.cws);
        im::epnp_cc_inv(sh.cws, sh.CCi);
    }
    __syncthreads();
    for (int i = tid; i < n; i += NT) im::epnp_alphas(sh.cws, sh.CCi, pws + 3 * i, alphas + 4 * i);
    __syncthreads();
    if (tid < 144) {
        const int r = tid / 12, c = tid - 12 * r;
        double acc = 0;
        for (int i = 0; i < n; i++) acc += im::epnp_mtm_entry(alphas + 4 * i, us + 3 * i, r, c);
        sh.MtM[tid] = acc;
    }
    __syncthreads();
    if (tid < 32) {
        jacobi_sym_warp(12, sh.MtM, sh.V);
        if (tid == 0) im::epnp_core_after_eig(sh.MtM, sh.V, sh.cws, &sh.core);
    }
    __syncthreads();
    for (int which = 0; which < 3; which++) {
        if (tid == 0) {
            im::epnp_ccs(&sh.core, which, sh.ccs);
            double z = 0;  // solve_for_sign (:244-252) looks at correspondence 0 only
            for (int j = 0; j < 4; j++) z += alphas[j] * sh.ccs[3 * j + 2];
            const double s0 = us[2] > 0 ? 1.0 : -1.0;
            sh.sgn = ((z < 0.0 && s0 > 0) || (z > 0.0 && s0 < 0)) ? -1.0 : 1.0;
        }
        __syncthreads();
        v[0] = v[1] = v[2] = 0;
        for (int i = tid; i < n; i += NT)
            for (int k = 0; k < 3; k++) {
                double s = 0;
                for (int j = 0; j < 4; j++) s += alphas[4 * i + j] * sh.ccs[3 * j + k];
                v[k] += sh.sgn * s;
            }
        block_sum<3>(v, sh.red);
        if (tid == 0)
            for (int k = 0; k < 3; k++) sh.pc0[k] = v[k] / n;
        __syncthreads();
        for (int k = 0; k < 9; k++) v[k] = 0;
        for (int i = tid; i < n; i += NT) {
            double pc[3];
            for (int k = 0; k < 3; k++) {
                double s = 0;
                for (int j = 0; j < 4; j++) s += alphas[4 * i + j] * sh.ccs[3 * j + k];
                pc[k] = sh.sgn * s - sh.pc0[k];
            }
            for (int a = 0; a < 3; a++)
                for (int b = 0; b < 3; b++) v[3 * a + b] += pc[a] * (pws[3 * i + b] - sh.mean[b]);
        }
        block_sum<9>(v, sh.red);
        if (tid == 0) im::epnp_rt_from_w(v, sh.pc0, sh.mean, sh.R, sh.T);
        __syncthreads();
        double err = 0;
        for (int i = tid; i < n; i += NT) {
            double p[3];
            im::mat3_vec(sh.R, pws + 3 * i, p);
            for (int k = 0; k < 3; k++) {
                const double d = us[3 * i + k] - (p[k] + sh.T[k]);
                err += d * d;
            }
        }
        block_sum<1>(&err, sh.red);
        if (tid == 0) {
            err /= n;
            if (which == 0 || err < sh.best) {  // compute_pose :318-337
                sh.best = err;
                for (int k = 0; k < 9; k++) sh.bestR[k] = sh.R[k];
                for (int k = 0; k < 3; k++) sh.bestT[k] = sh.T[k];
            }
        }
        __syncthreads();
    }
}

// one CTA per problem: correspondences [first[p], first[p+1])
__global__ void __launch_bounds__(NT) epnp_kernel(const double *pws, const double *us, const int *first, double *alphas, double *R,
                                                   double *T, double *err) {
    __shared__ EpnpShared sh;
    const int p = blockIdx.x, o = first[p], n = first[p + 1] - o;
    if (n < 4) {
        if (threadIdx.x == 0) err[p] = -1.0;
        return;
    }
    epnp_block(n, pws + 3 * (size_t)o, us + 3 * (size_t)o, alphas + 4 * (size_t)o, sh);
    if (threadIdx.x == 0) {
        for (int k = 0; k < 9; k++) R[9 * p + k] = sh.bestR[k];
        for (int k = 0; k < 3; k++) T[3 * p + k] = sh.bestT[k];
        err[p] = sh.best;
    }
}

// ---------------------------------------------------------------------------------------------------------------
// GlobalSFM::construct
// ---------------------------------------------------------------------------------------------------------------
struct SfmDev {
    int F, L, l;
    const int *obs_first, *obs_frame;
    const double *obs_pts;
    unsigned char *state;
    double *points;      // L x 3
    double *cR, *cT;     // F x 9, F x 3: c_Rotation / c_Translation (camera <- world of frame l)
    double *cq;          // F x 4: c_Quat (w, x, y, z), the BA's rotation blocks
    int *fail;           // [0] stage (0 ok, 1 PnP, 2 BA), [1] frame
};

__global__ void sfm_init_kernel(SfmDev d, const double *relative_R, const double *relative_T) {
    if (threadIdx.x != 0) return;
    const int last = d.F - 1;
    for (int k = 0; k < 9; k++) d.cR[9 * d.l + k] = (k % 4 == 0) ? 1.0 : 0.0;   // q[l] = identity, T[l] = 0 (:136-140)
    for (int k = 0; k < 3; k++) d.cT[3 * d.l + k] = 0.0;
    d.cq[4 * d.l] = 1.0, d.cq[4 * d.l + 1] = d.cq[4 * d.l + 2] = d.cq[4 * d.l + 3] = 0.0;
    double q[4], qi[4], t[3];
    im::rot_to_quat(relative_R, q);                                              // q[last] = q[l] * Quaterniond(relative_R)
    const double n2 = q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3];
    qi[0] = q[0] / n2, qi[1] = -q[1] / n2, qi[2] = -q[2] / n2, qi[3] = -q[3] / n2;  // c_Quat[last] = q[last].inverse()
    im::quat_to_rot(qi, d.cR + 9 * last);
    im::mat3_vec(d.cR + 9 * last, relative_T, t);
    for (int k = 0; k < 3; k++) d.cT[3 * last + k] = -t[k], d.cq[4 * last + k] = qi[k];
    d.cq[4 * last + 3] = qi[3];
}

__device__ __forceinline__ void sfm_pose(const SfmDev &d, int f, double *P) {
    for (int r = 0; r < 3; r++) {
        for (int c = 0; c < 3; c++) P[4 * r + c] = d.cR[9 * f + 3 * r + c];
        P[4 * r + 3] = d.cT[3 * f + r];
    }
}

// triangulateTwoFrames (:74-114) for f0 >= 0; the "all other points" loop (:207-228) for f0 < 0
__global__ void sfm_triangulate_kernel(SfmDev d, int f0, int f1) {
    if (d.fail[0]) return;
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= d.L || d.state[j]) return;
    const int a = d.obs_first[j], b = d.obs_first[j + 1];
    int o0 = -1, o1 = -1;
    if (f0 >= 0) {
        for (int o = a; o < b; o++) {
            if (d.obs_frame[o] == f0) o0 = o;
            if (d.obs_frame[o] == f1) o1 = o;
        }
        if (o0 < 0 || o1 < 0) return;
    } else {
        if (b - a < 2) return;
        o0 = a, o1 = b - 1;
    }
    double P0[12], P1[12], X[3];
    sfm_pose(d, d.obs_frame[o0], P0);
    sfm_pose(d, d.obs_frame[o1], P1);
    im::triangulate_point(P0, P1, d.obs_pts + 3 * o0, d.obs_pts + 3 * o1, X);
    for (int k = 0; k < 3; k++) d.points[3 * j + k] = X[k];
    d.state[j] = 1;
}

// solveFrameByPnP (:25-72): gather the triangulated features seen in frame i (in feature order: the sign fix of the EPnP
// looks at the first one), EPnP, write the pose of frame i.
__global__ void __launch_bounds__(NT) sfm_pnp_kernel(SfmDev d, int i, double *pw_list, double *us_list, double *alphas) {
    __shared__ EpnpShared sh;
    __shared__ int s_cnt[NT + 1];
    if (d.fail[0]) return;
    const int tid = threadIdx.x;
    const int chunk = (d.L + NT - 1) / NT, j0 = tid * chunk, j1 = min(d.L, j0 + chunk);
    int cnt = 0;
    for (int j = j0; j < j1; j++) {
        if (!d.state[j]) continue;
        for (int o = d.obs_first[j]; o < d.obs_first[j + 1]; o++)
            if (d.obs_frame[o] == i) {
                cnt++;
                break;
            }
    }
    s_cnt[tid + 1] = cnt;
    __syncthreads();
    if (tid == 0) {
        s_cnt[0] = 0;
        for (int t = 0; t < NT; t++) s_cnt[t + 1] += s_cnt[t];
    }
    __syncthreads();
    int pos = s_cnt[tid];
    const int n = s_cnt[NT];
    for (int j = j0; j < j1; j++) {
        if (!d.state[j]) continue;
        for (int o = d.obs_first[j]; o < d.obs_first[j + 1]; o++)
            if (d.obs_frame[o] == i) {
                for (int k = 0; k < 3; k++) {
                    pw_list[3 * pos + k] = d.points[3 * j + k];
                    us_list[3 * pos + k] = d.obs_pts[3 * o + k];
                }
                pos++;
                break;
            }
    }
    __syncthreads();
    if (n < 10) {  // :50-55 (below 15 the reference only prints a warning)
        if (tid == 0) d.fail[0] = 1, d.fail[1] = i;
        return;
    }
    epnp_block(n, pw_list, us_list, alphas, sh);
    if (tid == 0) {
        for (int k = 0; k < 9; k++) d.cR[9 * i + k] = sh.bestR[k];
        for (int k = 0; k < 3; k++) d.cT[3 * i + k] = sh.bestT[k];
        im::rot_to_quat(sh.bestR, d.cq + 4 * i);  // c_Quat[i] = c_Rotation[i]
    }
}

// Per-observation record kept between the phases of one LM iteration (L2 resident): r (3) | Jc (18, masked) | Jp (9) |
// W = Jc_s^T Jp_s (18) | Y = W Vinv (18)
constexpr int kObsRec = 66;
constexpr int OB_R = 0, OB_JC = 3, OB_JP = 21, OB_W = 30, OB_Y = 48;

// The LM loop of ba_core.cuh by one CTA.  Same per-observation arithmetic and the same trust-region decisions as the
// CPU-checked work items there (ba_obs, ba_w, ba_lm_diag, ba_jacobi_scale, ba_cam_candidate, ba_point_cost, ba_begin_iteration,
// ba_decide); the reduced camera system is GATHERED, not scattered: a thread owns an entry of S and sums over the points seen
// in both of its frames through the table tab[j][f] (observation of point j in frame f, or -1), so there are no atomics and
// every Jacobian is evaluated once per iteration.
// Dynamic shared memory: S (D x D) | gc (D) | scale_c (D) | dxc (D) | cq (4F) | ct (3F) | cand_q (4F) | cand_t (3F) | red
__global__ void __launch_bounds__(BA_NT) sfm_ba_kernel(SfmDev d, double *ct_io /* F x 3 */, double *V, double *gp, double *sp, double *Vinv,
                                                       double *cand, double *obs_rec, int *tab, im::BaCtl *ctl_out) {
    extern __shared__ double smem[];
    __shared__ im::BaCtl c;
    __shared__ int s_go, s_accept;
    if (d.fail[0]) return;
    const int tid = threadIdx.x, F = d.F, L = d.L, D = 6 * F;
    double *S = smem, *gc = S + D * D, *scale_c = gc + D, *dxc = scale_c + D, *cq = dxc + D, *ct = cq + 4 * F, *cand_q = ct + 3 * F,
           *cand_t = cand_q + 4 * F, *red = cand_t + 3 * F;
    const im::BaProblem P{F, L, d.l, d.obs_first, d.obs_frame, d.obs_pts, d.state};
    for (int k = tid; k < 4 * F; k += BA_NT) cq[k] = d.cq[k];
    for (int k = tid; k < 3 * F; k += BA_NT) ct[k] = ct_io[k];
    for (int k = tid; k < D; k += BA_NT) scale_c[k] = 1.0;
    for (int k = tid; k < L * F; k += BA_NT) tab[k] = -1;
    if (tid == 0) im::ba_init(&c);
    __syncthreads();
    for (int j = tid; j < L; j += BA_NT)
        if (d.state[j])
            for (int o = d.obs_first[j]; o < d.obs_first[j + 1]; o++) tab[j * F + d.obs_frame[o]] = o;
    bool first = true;
    __syncthreads();
    while (true) {
        // ---- phase 1 (thread per point): residuals and Jacobians of its observations, V, g_p, cost ----
        double v[3] = {0, 0, 0}, gmax = 0;
        for (int j = tid; j < L; j += BA_NT) {
            if (!d.state[j]) continue;
            double Vj[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0}, gpj[3] = {0, 0, 0};
            for (int o = d.obs_first[j]; o < d.obs_first[j + 1]; o++) {
                double r[3], Jc[18], Jp[9];
                im::ba_obs(P, cq, ct, d.points + 3 * j, o, r, Jc, Jp);
                double *rec = obs_rec + (size_t)kObsRec * o;
                for (int k = 0; k < 3; k++) rec[OB_R + k] = r[k];
                for (int k = 0; k < 18; k++) rec[OB_JC + k] = Jc[k];
                for (int k = 0; k < 9; k++) rec[OB_JP + k] = Jp[k];
                v[0] += 0.5 * (r[0] * r[0] + r[1] * r[1] + r[2] * r[2]);
                for (int a = 0; a < 3; a++) {
                    for (int b = 0; b < 3; b++) Vj[3 * a + b] += Jp[a] * Jp[b] + Jp[3 + a] * Jp[3 + b] + Jp[6 + a] * Jp[6 + b];
                    gpj[a] += Jp[a] * r[0] + Jp[3 + a] * r[1] + Jp[6 + a] * r[2];
                }
            }
            for (int k = 0; k < 9; k++) V[9 * j + k] = Vj[k];
            for (int a = 0; a < 3; a++) {
                gp[3 * j + a] = gpj[a];
                gmax = fmax(gmax, fabs(gpj[a]));
            }
        }
        block_sum<1>(v, red);  // (its barriers publish obs_rec / V / gp)
        // ---- U (diagonal blocks of S) and g_c gathered per entry over the observations of the frame ----
        for (int k = tid; k < D * D; k += BA_NT) S[k] = 0.0;
        __syncthreads();
        for (int e = tid; e < 42 * F; e += BA_NT) {
            const int f = e / 42, q = e - 42 * f, a = q < 36 ? q / 6 : q - 36, b = q < 36 ? q % 6 : -1;
            double acc = 0;
            for (int j = 0; j < L; j++) {
                const int o = tab[j * F + f];
                if (o < 0) continue;
                const double *rec = obs_rec + (size_t)kObsRec * o;
                if (b >= 0) acc += rec[OB_JC + a] * rec[OB_JC + b] + rec[OB_JC + 6 + a] * rec[OB_JC + 6 + b] + rec[OB_JC + 12 + a] * rec[OB_JC + 12 + b];
                else acc += rec[OB_JC + a] * rec[OB_R] + rec[OB_JC + 6 + a] * rec[OB_R + 1] + rec[OB_JC + 12 + a] * rec[OB_R + 2];
            }
            if (b >= 0) S[(6 * f + a) * D + 6 * f + b] = acc;
            else gc[6 * f + a] = acc;
        }
        __syncthreads();
        for (int k = tid; k < D; k += BA_NT) gmax = fmax(gmax, fabs(gc[k]));
        gmax = block_max(gmax, red);
        if (tid == 0) {
            c.cost = v[0], c.gmax = gmax;
            s_go = im::ba_begin_iteration(&c) ? 1 : 0;
        }
        __syncthreads();
        if (!s_go) break;
        if (first) {
            for (int k = tid; k < D; k += BA_NT) scale_c[k] = im::ba_jacobi_scale(S[k * D + k]);
            for (int j = tid; j < L; j += BA_NT)
                if (d.state[j])
                    for (int a = 0; a < 3; a++) sp[3 * j + a] = im::ba_jacobi_scale(V[9 * j + 4 * a]);
            first = false;
            __syncthreads();
        }
        const double radius = c.radius;
        // ---- phase 2 (thread per point): scaled, damped V -> Vinv; W and Y of its observations ----
        for (int j = tid; j < L; j += BA_NT) {
            if (!d.state[j]) continue;
            const double *spj = sp + 3 * j;
            double Vs[9], Vi[9];
            for (int a = 0; a < 3; a++)
                for (int b = 0; b < 3; b++) Vs[3 * a + b] = spj[a] * spj[b] * V[9 * j + 3 * a + b];
            for (int a = 0; a < 3; a++) Vs[4 * a] += im::ba_lm_diag(Vs[4 * a], radius);
            im::inv3(Vs, Vi);
            for (int k = 0; k < 9; k++) Vinv[9 * j + k] = Vi[k];
            for (int o = d.obs_first[j]; o < d.obs_first[j + 1]; o++) {
                double *rec = obs_rec + (size_t)kObsRec * o;
                double W[18];
                im::ba_w(rec + OB_JC, rec + OB_JP, scale_c + 6 * d.obs_frame[o], spj, W);
                for (int a = 0; a < 6; a++)
                    for (int b = 0; b < 3; b++) {
                        rec[OB_W + 3 * a + b] = W[3 * a + b];
                        rec[OB_Y + 3 * a + b] = W[3 * a] * Vi[b] + W[3 * a + 1] * Vi[3 + b] + W[3 * a + 2] * Vi[6 + b];
                    }
            }
        }
        __syncthreads();
        // ---- phase 3 (thread per entry of the lower block triangle): S = U_s + D^2 - sum_j Y W^T, rhs = g_c,s - sum_j Y g_p,s ----
        const int n_blk = F * (F + 1) / 2;
        for (int e = tid; e < 36 * n_blk; e += BA_NT) {
            const int blk = e / 36, q = e - 36 * blk, b = q / 6, a = q - 6 * b;  // entry (6 f2 + b, 6 f1 + a), f2 >= f1
            int f2 = 0, rem = blk;
            while (rem > f2) rem -= f2 + 1, f2++;
            const int f1 = rem;
            double acc = 0;
            for (int j = 0; j < L; j++) {
                const int o1 = tab[j * F + f1];
                if (o1 < 0) continue;
                const int o2 = f1 == f2 ? o1 : tab[j * F + f2];
                if (o2 < 0) continue;
                const double *y = obs_rec + (size_t)kObsRec * o1 + OB_Y + 3 * a, *w = obs_rec + (size_t)kObsRec * o2 + OB_W + 3 * b;
                acc += y[0] * w[0] + y[1] * w[1] + y[2] * w[2];
            }
            const int row = 6 * f2 + b, col = 6 * f1 + a;
            double u = 0;
            if (f1 == f2) {
                u = scale_c[row] * scale_c[col] * S[row * D + col];
                if (row == col) u = (u == 0.0) ? 1.0 : u + im::ba_lm_diag(u, radius);  // constant / unobserved column: identity row
            }
            if (f1 == f2 && row < col) continue;  // upper part of a diagonal block: the Cholesky reads the lower triangle only
            S[row * D + col] = u - acc;           // (each thread reads and writes only its own entry of S)
        }
        for (int e = tid; e < D; e += BA_NT) {
            const int f = e / 6, a = e - 6 * f;
            double acc = 0;
            for (int j = 0; j < L; j++) {
                const int o = tab[j * F + f];
                if (o < 0) continue;
                const double *y = obs_rec + (size_t)kObsRec * o + OB_Y + 3 * a;
                acc += y[0] * sp[3 * j] * gp[3 * j] + y[1] * sp[3 * j + 1] * gp[3 * j + 1] + y[2] * sp[3 * j + 2] * gp[3 * j + 2];
            }
            dxc[e] = scale_c[e] * gc[e] - acc;
        }
        if (tid == 0) c.chol_ok = 1;
        __syncthreads();
        cta_cholesky_solve(S, D, D, dxc, &c.chol_ok);
        __syncthreads();
        v[0] = v[1] = v[2] = 0;
        double cand_cost = 0;
        if (c.chol_ok) {
            for (int k = tid; k < D; k += BA_NT) dxc[k] = -dxc[k];
            __syncthreads();
            // ---- phase 5 (thread per point): back-substitution, candidate, model cost change ----
            for (int j = tid; j < L; j += BA_NT) {
                if (!d.state[j]) continue;
                const double *spj = sp + 3 * j, *Vi = Vinv + 9 * j;
                double wdx[3] = {0, 0, 0};
                for (int o = d.obs_first[j]; o < d.obs_first[j + 1]; o++) {
                    const double *W = obs_rec + (size_t)kObsRec * o + OB_W, *dx = dxc + 6 * d.obs_frame[o];
                    for (int b = 0; b < 3; b++)
                        for (int a = 0; a < 6; a++) wdx[b] += W[3 * a + b] * dx[a];
                }
                double rhs[3], dxp[3];
                for (int a = 0; a < 3; a++) rhs[a] = spj[a] * gp[3 * j + a] + wdx[a];
                for (int a = 0; a < 3; a++) dxp[a] = -(Vi[3 * a] * rhs[0] + Vi[3 * a + 1] * rhs[1] + Vi[3 * a + 2] * rhs[2]);
                for (int o = d.obs_first[j]; o < d.obs_first[j + 1]; o++) {
                    const int f = d.obs_frame[o];
                    const double *rec = obs_rec + (size_t)kObsRec * o;
                    for (int i = 0; i < 3; i++) {
                        double m = 0;
                        for (int a = 0; a < 6; a++) m += rec[OB_JC + 6 * i + a] * scale_c[6 * f + a] * dxc[6 * f + a];
                        for (int a = 0; a < 3; a++) m += rec[OB_JP + 3 * i + a] * spj[a] * dxp[a];
                        v[0] += -(m * (rec[OB_R + i] + m / 2.0));
                    }
                }
                for (int a = 0; a < 3; a++) {
                    const double dd = spj[a] * dxp[a];
                    cand[3 * j + a] = d.points[3 * j + a] + dd;
                    v[1] += dd * dd;
                    v[2] += d.points[3 * j + a] * d.points[3 * j + a];
                }
            }
            if (tid < F) im::ba_cam_candidate(P, cq, ct, tid, scale_c, dxc, cand_q, cand_t, &v[1], &v[2]);
            block_sum<3>(v, red);  // (its barriers also publish cand / cand_q / cand_t)
            for (int j = tid; j < L; j += BA_NT)
                if (d.state[j]) cand_cost += im::ba_point_cost(P, cand_q, cand_t, cand, j);
            block_sum<1>(&cand_cost, red);
        }
        if (tid == 0) {
            c.model_change = v[0], c.step2 = v[1], c.xnorm2 = v[2], c.cand_cost = cand_cost;
            s_accept = im::ba_decide(&c);
        }
        __syncthreads();
        if (s_accept) {
            for (int j = tid; j < L; j += BA_NT)
                if (d.state[j])
                    for (int a = 0; a < 3; a++) d.points[3 * j + a] = cand[3 * j + a];
            for (int k = tid; k < 4 * F; k += BA_NT) cq[k] = cand_q[k];
            for (int k = tid; k < 3 * F; k += BA_NT) ct[k] = cand_t[k];
        }
        __syncthreads();
        if (c.done) break;
    }
    for (int k = tid; k < 4 * F; k += BA_NT) d.cq[k] = cq[k];
    for (int k = tid; k < 3 * F; k += BA_NT) ct_io[k] = ct[k];
    if (tid == 0) {
        *ctl_out = c;
        // initial_sfm.cpp:283-294: "CONVERGENCE || final_cost < 5e-03", otherwise construct() fails
        if (!(c.converged || c.cost < 5e-3) && d.fail[0] == 0) d.fail[0] = 2, d.fail[1] = -1;
    }
}

// q[i] = c_Quat[i].inverse(), T[i] = -(q[i] * c_translation[i])  (:295-309)
__global__ void sfm_finish_kernel(SfmDev d, const double *ct, double *q_out, double *T_out) {
    const int i = threadIdx.x;
    if (i >= d.F) return;
    const double *c = d.cq + 4 * i;
    const double n2 = c[0] * c[0] + c[1] * c[1] + c[2] * c[2] + c[3] * c[3];
    const double q[4] = {c[0] / n2, -c[1] / n2, -c[2] / n2, -c[3] / n2};
    double R[9], t[3];
    im::quat_to_rot(q, R);
    im::mat3_vec(R, ct + 3 * i, t);
    for (int k = 0; k < 4; k++) q_out[4 * i + k] = q[k];
    for (int k = 0; k < 3; k++) T_out[3 * i + k] = -t[k];
}

size_t ba_smem_bytes(int F) {
    const int D = 6 * F;
    return sizeof(double) * ((size_t)D * D + 3 * D + 14 * F + 16 * (BA_NT / 32));
}

void fill_summary(const im::BaCtl &c, lfvio_sfm_summary *s) {
    s->ba_iterations = c.iter;
    s->ba_successful = c.n_success;
    s->ba_converged = c.converged;
    s->ba_final_cost = c.cost;
    for (int k = 0; k <= LFVIO_SFM_MAX_BA_ITER; k++) s->ba_costs[k] = k <= c.n_success ? c.costs[k] : 0.0;
}

int check_tracks(const int *obs_first, const int *obs_frame, int n_features, int frame_num) {
    LFVIO_REQUIRE(obs_first[0] == 0, "sfm: obs_first[0] must be 0");
    for (int j = 0; j < n_features; j++) {
        LFVIO_REQUIRE(obs_first[j + 1] >= obs_first[j], "sfm: obs_first must be non-decreasing");
        for (int o = obs_first[j]; o < obs_first[j + 1]; o++) {
            LFVIO_REQUIRE(obs_frame[o] >= 0 && obs_frame[o] < frame_num, "sfm: observation frame outside the window");
            LFVIO_REQUIRE(o == obs_first[j] || obs_frame[o] > obs_frame[o - 1], "sfm: observations of a feature must be in ascending frame order");
        }
    }
    return LFVIO_OK;
}

}  // namespace

}  // namespace lfvio

using namespace lfvio;

extern "C" int lfvio_solve_gyro_bias(const double *R, const double *delta_q, const double *J_rg, int n_frames, double *delta_bg,
                                     int device) {
    clear_error();
    LFVIO_REQUIRE(R && delta_q && J_rg && delta_bg && n_frames >= 2, "lfvio_solve_gyro_bias: bad arguments");
    LFVIO_CUDA_TRY(cudaSetDevice(device));
    DBuf<double> dR, dq, dJ, out;
    LFVIO_TRY(dR.put(R, 9 * (size_t)n_frames));
    LFVIO_TRY(dq.put(delta_q, 4 * (size_t)(n_frames - 1)));
    LFVIO_TRY(dJ.put(J_rg, 9 * (size_t)(n_frames - 1)));
    LFVIO_TRY(out.put(nullptr, 3));
    gyro_bias_kernel<<<1, 32>>>(dR.p, dq.p, dJ.p, n_frames, out.p);
    LFVIO_LAUNCH_CHECK();
    return out.get(delta_bg, 3);
}

extern "C" int lfvio_linear_alignment(const double *R, const double *T, const double *delta_p, const double *delta_v,
                                      const double *sum_dt, int n_frames, const double *tic, double g_norm, double *g, double *x,
                                      int *ok, int device) {
    clear_error();
    LFVIO_REQUIRE(R && T && delta_p && delta_v && sum_dt && tic && g && x && ok, "lfvio_linear_alignment: null argument");
    LFVIO_REQUIRE(n_frames >= 2 && n_frames <= LFVIO_INIT_MAX_FRAMES, "lfvio_linear_alignment: 2 <= n_frames <= LFVIO_INIT_MAX_FRAMES");
    LFVIO_CUDA_TRY(cudaSetDevice(device));
    const int ns = 3 * n_frames + 4;
    DBuf<double> dR, dT, dp, dv, dt, A, b, Aacc, bacc, dg, dx;
    DBuf<int> dok;
    LFVIO_TRY(dR.put(R, 9 * (size_t)n_frames));
    LFVIO_TRY(dT.put(T, 3 * (size_t)n_frames));
    LFVIO_TRY(dp.put(delta_p, 3 * (size_t)(n_frames - 1)));
    LFVIO_TRY(dv.put(delta_v, 3 * (size_t)(n_frames - 1)));
    LFVIO_TRY(dt.put(sum_dt, (size_t)(n_frames - 1)));
    LFVIO_TRY(A.put(nullptr, (size_t)ns * ns));
    LFVIO_TRY(b.put(nullptr, ns));
    LFVIO_TRY(Aacc.put(nullptr, (size_t)ns * ns));
    LFVIO_TRY(bacc.put(nullptr, ns));
    LFVIO_TRY(dg.put(nullptr, 3));
    LFVIO_TRY(dx.put(nullptr, ns));
    LFVIO_TRY(dok.put(nullptr, 1));
    alignment_kernel<<<1, NT>>>(dR.p, dT.p, dp.p, dv.p, dt.p, n_frames, tic[0], tic[1], tic[2], g_norm, A.p, b.p, Aacc.p, bacc.p, dg.p, dx.p, dok.p);
    LFVIO_LAUNCH_CHECK();
    LFVIO_TRY(dg.get(g, 3));
    LFVIO_TRY(dx.get(x, ns));
    return dok.get(ok, 1);
}

static int relative_pose_impl(const double *corres, int n, unsigned long long seed, unsigned frame, int mode, double *rotation,
                              double *translation, int *inlier_cnt, double *E_out, int device) {
    LFVIO_CUDA_TRY(cudaSetDevice(device));
    std::vector<double> h1(3 * (size_t)n), h2(3 * (size_t)n);
    for (int i = 0; i < n; i++)
        for (int k = 0; k < 3; k++) h1[3 * i + k] = corres[6 * i + k], h2[3 * i + k] = corres[6 * i + 3 + k];
    DBuf<double> b1, b2, E, Eb, R, T;
    DBuf<float> score;
    DBuf<uint8_t> flags, st;
    DBuf<int> nd, best, cnt;
    LFVIO_TRY(b1.put(h1.data(), h1.size()));
    LFVIO_TRY(b2.put(h2.data(), h2.size()));
    LFVIO_TRY(E.put(nullptr, 900));
    LFVIO_TRY(Eb.put(nullptr, 9));
    LFVIO_TRY(R.put(nullptr, 9));
    LFVIO_TRY(T.put(nullptr, 3));
    LFVIO_TRY(score.put(nullptr, 100));
    LFVIO_TRY(flags.put(nullptr, 100 * (size_t)n));
    LFVIO_TRY(st.put(nullptr, n));
    LFVIO_TRY(nd.put(&n, 1));
    LFVIO_TRY(best.put(nullptr, 1));
    LFVIO_TRY(cnt.put(nullptr, 1));
    const RansacScratch rs{reinterpret_cast<double3 *>(b1.p), reinterpret_cast<double3 *>(b2.p), E.p, score.p, flags.p, Eb.p, best.p};
    LFVIO_TRY(launch_ransac_bearings(nd.p, n, n, seed, frame, rs, st.p, 0));
    recover_pose_kernel<<<1, NT>>>(Eb.p, rs.b1, rs.b2, n, mode, R.p, T.p, cnt.p);
    LFVIO_LAUNCH_CHECK();
    LFVIO_TRY(R.get(rotation, 9));
    if (translation) LFVIO_TRY(T.get(translation, 3));
    if (inlier_cnt) LFVIO_TRY(cnt.get(inlier_cnt, 1));
    if (E_out) LFVIO_TRY(Eb.get(E_out, 9));
    return LFVIO_OK;
}

extern "C" int lfvio_solve_relative_rt(const double *corres, int n, unsigned long long seed, unsigned frame, double *rotation,
                                       double *translation, int *inlier_cnt, double *E_out, int device) {
    clear_error();
    LFVIO_REQUIRE(n >= 0 && rotation && translation && inlier_cnt && (n == 0 || corres), "lfvio_solve_relative_rt: bad arguments");
    *inlier_cnt = 0;
    if (n < 15) return LFVIO_OK;  // solve_5pts.cpp:540: below 15 correspondences the reference returns false untouched
    return relative_pose_impl(corres, n, seed, frame, 0, rotation, translation, inlier_cnt, E_out, device);
}

extern "C" int lfvio_solve_relative_r(const double *corres, int n, unsigned long long seed, unsigned frame, double *rotation,
                                      int device) {
    clear_error();
    LFVIO_REQUIRE(n >= 0 && rotation && (n == 0 || corres), "lfvio_solve_relative_r: bad arguments");
    if (n < 9) {  // initial_ex_rotation.cpp:222, :288: identity below 9 correspondences
        for (int k = 0; k < 9; k++) rotation[k] = (k % 4 == 0) ? 1.0 : 0.0;
        return LFVIO_OK;
    }
    return relative_pose_impl(corres, n, seed, frame, 1, rotation, nullptr, nullptr, nullptr, device);
}

namespace lfvio {
namespace {
__global__ void ex_rotation_kernel(int N, const double *Rc, const double *Rimu, const double *Rc_g, double *A, double *ric, double *sv) {
    if (threadIdx.x == 0) im::ex_rotation_solve(N, Rc, Rimu, Rc_g, A, ric, sv);
}
}  // namespace
}  // namespace lfvio

extern "C" int lfvio_calibrate_ex_rotation(const double *Rc, const double *Rimu, const double *Rc_g, int n, double *ric, double *sv,
                                           int device) {
    clear_error();
    LFVIO_REQUIRE(Rc && Rimu && Rc_g && ric && sv && n >= 1 && n <= 4096, "lfvio_calibrate_ex_rotation: bad arguments");
    LFVIO_CUDA_TRY(cudaSetDevice(device));
    DBuf<double> dRc, dRi, dRg, dA, dric, dsv;
    LFVIO_TRY(dRc.put(Rc, 9 * (size_t)n));
    LFVIO_TRY(dRi.put(Rimu, 9 * (size_t)n));
    LFVIO_TRY(dRg.put(Rc_g, 9 * (size_t)n));
    LFVIO_TRY(dA.put(nullptr, 16 * (size_t)n));
    LFVIO_TRY(dric.put(nullptr, 9));
    LFVIO_TRY(dsv.put(nullptr, 4));
    ex_rotation_kernel<<<1, 32>>>(n, dRc.p, dRi.p, dRg.p, dA.p, dric.p, dsv.p);
    LFVIO_LAUNCH_CHECK();
    LFVIO_TRY(dric.get(ric, 9));
    return dsv.get(sv, 4);
}

extern "C" int lfvio_epnp(const double *pws, const double *us, const int *first, int n_problems, double *R, double *T, double *err,
                          int device) {
    clear_error();
    LFVIO_REQUIRE(n_problems >= 0 && (n_problems == 0 || (pws && us && first && R && T)), "lfvio_epnp: bad arguments");
    if (n_problems == 0) return LFVIO_OK;
    LFVIO_REQUIRE(first[0] == 0, "lfvio_epnp: first[0] must be 0");
    for (int p = 0; p < n_problems; p++) LFVIO_REQUIRE(first[p + 1] - first[p] >= 4, "lfvio_epnp: a problem needs at least 4 correspondences");
    const size_t n = first[n_problems];
    LFVIO_CUDA_TRY(cudaSetDevice(device));
    DBuf<double> dpw, dus, al, dR, dT, de;
    DBuf<int> df;
    LFVIO_TRY(dpw.put(pws, 3 * n));
    LFVIO_TRY(dus.put(us, 3 * n));
    LFVIO_TRY(al.put(nullptr, 4 * n));
    LFVIO_TRY(df.put(first, (size_t)n_problems + 1));
    LFVIO_TRY(dR.put(nullptr, 9 * (size_t)n_problems));
    LFVIO_TRY(dT.put(nullptr, 3 * (size_t)n_problems));
    LFVIO_TRY(de.put(nullptr, n_problems));
    epnp_kernel<<<n_problems, NT>>>(dpw.p, dus.p, df.p, al.p, dR.p, dT.p, de.p);
    LFVIO_LAUNCH_CHECK();
    LFVIO_TRY(dR.get(R, 9 * (size_t)n_problems));
    LFVIO_TRY(dT.get(T, 3 * (size_t)n_problems));
    if (err) LFVIO_TRY(de.get(err, n_problems));
    return LFVIO_OK;
}

namespace {

// Device scratch of the SfM entries: one grow-only allocation per device, kept for the life of the process and handed to one
// call at a time (a cudaMalloc / cudaFree pair costs more than most of the kernels here, and cudaFree synchronises the device).
struct SfmScratch {
    std::mutex mu;
    unsigned char *p[64] = {};
    size_t cap[64] = {};
};
SfmScratch &sfm_scratch() {
    static SfmScratch s;
    return s;
}

struct SfmBuffers {
    std::unique_lock<std::mutex> lock;
    struct {
        unsigned char *p = nullptr;
    } arena;
    int acquire(size_t bytes) {
        int dev = 0;
        LFVIO_CUDA_TRY(cudaGetDevice(&dev));
        SfmScratch &sc = sfm_scratch();
        lock = std::unique_lock<std::mutex>(sc.mu);
        dev &= 63;
        if (sc.cap[dev] < bytes) {
            if (sc.p[dev]) LFVIO_CUDA_TRY(cudaFree(sc.p[dev]));
            sc.p[dev] = nullptr, sc.cap[dev] = 0;
            const size_t want = (bytes + (bytes >> 2) + ((size_t)1 << 20)) & ~(((size_t)1 << 20) - 1);
            LFVIO_CUDA_TRY(cudaMalloc(&sc.p[dev], want));
            sc.cap[dev] = want;
        }
        arena.p = sc.p[dev];
        LFVIO_CUDA_TRY(cudaMemsetAsync(arena.p, 0, bytes, 0));
        return LFVIO_OK;
    }
    template <class T>
    struct Ref {
        T *p = nullptr;
        size_t n = 0;
        int get(T *h, size_t count) const {
            LFVIO_CUDA_TRY(cudaMemcpy(h, p, sizeof(T) * count, cudaMemcpyDeviceToHost));
            return LFVIO_OK;
        }
        int put(const T *h, size_t count) {
            LFVIO_CUDA_TRY(cudaMemcpy(p, h, sizeof(T) * count, cudaMemcpyHostToDevice));
            return LFVIO_OK;
        }
    };
    Ref<int> of, ofr, fail;
    Ref<double> pts_obs, points, cR, cT, cq, V, gp, sp, Vinv, cand, pw_list, us_list, alphas, relR, relT, q_out, T_out, obs_rec;
    Ref<int> tab;
    Ref<unsigned char> state;
    Ref<im::BaCtl> ctl;
    SfmDev dev{};
    int setup(int F, int l, const int *obs_first, const int *obs_frame, const double *obs_pts, int L) {
        const size_t n_obs = obs_first[L];
        size_t total = 0;
        auto plan = [&](auto &r, size_t n) {
            r.n = n;
            r.p = reinterpret_cast<decltype(r.p)>(total);  // offset for now
            total += (sizeof(*r.p) * (n ? n : 1) + 255) / 256 * 256;
        };
        plan(of, (size_t)L + 1), plan(ofr, n_obs), plan(pts_obs, 3 * n_obs), plan(fail, 2), plan(points, 3 * (size_t)L), plan(state, L);
        plan(cR, 9 * (size_t)F), plan(cT, 3 * (size_t)F), plan(cq, 4 * (size_t)F), plan(V, 9 * (size_t)L), plan(gp, 3 * (size_t)L);
        plan(sp, 3 * (size_t)L), plan(Vinv, 9 * (size_t)L), plan(cand, 3 * (size_t)L), plan(pw_list, 3 * (size_t)L), plan(us_list, 3 * (size_t)L);
        plan(alphas, 4 * (size_t)L), plan(relR, 9), plan(relT, 3), plan(q_out, 4 * (size_t)F), plan(T_out, 3 * (size_t)F), plan(ctl, 1);
        plan(obs_rec, (size_t)kObsRec * n_obs), plan(tab, (size_t)L * F);
        LFVIO_TRY(acquire(total));  // (zero-filled)
        auto fix = [&](auto &r) { r.p = reinterpret_cast<decltype(r.p)>(arena.p + reinterpret_cast<size_t>(r.p)); };
        fix(of), fix(ofr), fix(pts_obs), fix(fail), fix(points), fix(state), fix(cR), fix(cT), fix(cq), fix(V), fix(gp), fix(sp), fix(Vinv);
        fix(cand), fix(pw_list), fix(us_list), fix(alphas), fix(relR), fix(relT), fix(q_out), fix(T_out), fix(ctl), fix(obs_rec), fix(tab);
        LFVIO_TRY(of.put(obs_first, (size_t)L + 1));
        LFVIO_TRY(ofr.put(obs_frame, n_obs));
        LFVIO_TRY(pts_obs.put(obs_pts, 3 * n_obs));
        dev = SfmDev{F, L, l, of.p, ofr.p, pts_obs.p, state.p, points.p, cR.p, cT.p, cq.p, fail.p};
        return LFVIO_OK;
    }
    int run_ba(int F) {
        const size_t smem = ba_smem_bytes(F);
        LFVIO_TRY(ensure_dyn_smem((const void *)sfm_ba_kernel, (int)ba_smem_bytes(LFVIO_SFM_MAX_FRAMES)));
        sfm_ba_kernel<<<1, BA_NT, smem>>>(dev, cT.p, V.p, gp.p, sp.p, Vinv.p, cand.p, obs_rec.p, tab.p, ctl.p);
        LFVIO_LAUNCH_CHECK();
        return LFVIO_OK;
    }
};

}  // namespace

extern "C" int lfvio_sfm_bundle_adjust(int frame_num, int l, const int *obs_first, const int *obs_frame, const double *obs_pts,
                                       int n_features, const unsigned char *state, double *c_rot, double *c_trans, double *points,
                                       lfvio_sfm_summary *summary, int device) {
    clear_error();
    LFVIO_REQUIRE(obs_first && obs_frame && obs_pts && state && c_rot && c_trans && points && summary, "lfvio_sfm_bundle_adjust: null argument");
    LFVIO_REQUIRE(frame_num >= 2 && frame_num <= LFVIO_SFM_MAX_FRAMES && l >= 0 && l < frame_num && n_features > 0,
                  "lfvio_sfm_bundle_adjust: bad sizes");
    LFVIO_TRY(check_tracks(obs_first, obs_frame, n_features, frame_num));
    LFVIO_CUDA_TRY(cudaSetDevice(device));
    SfmBuffers B;
    LFVIO_TRY(B.setup(frame_num, l, obs_first, obs_frame, obs_pts, n_features));
    LFVIO_CUDA_TRY(cudaMemcpy(B.state.p, state, n_features, cudaMemcpyHostToDevice));
    LFVIO_CUDA_TRY(cudaMemcpy(B.points.p, points, sizeof(double) * 3 * n_features, cudaMemcpyHostToDevice));
    LFVIO_CUDA_TRY(cudaMemcpy(B.cq.p, c_rot, sizeof(double) * 4 * frame_num, cudaMemcpyHostToDevice));
    LFVIO_CUDA_TRY(cudaMemcpy(B.cT.p, c_trans, sizeof(double) * 3 * frame_num, cudaMemcpyHostToDevice));
    LFVIO_TRY(B.run_ba(frame_num));
    im::BaCtl c;
    LFVIO_TRY(B.ctl.get(&c, 1));
    LFVIO_TRY(B.cq.get(c_rot, 4 * (size_t)frame_num));
    LFVIO_TRY(B.cT.get(c_trans, 3 * (size_t)frame_num));
    LFVIO_TRY(B.points.get(points, 3 * (size_t)n_features));
    int fail[2];
    LFVIO_TRY(B.fail.get(fail, 2));
    summary->ok = fail[0] == 0, summary->fail_stage = fail[0], summary->fail_frame = fail[1];
    fill_summary(c, summary);
    return LFVIO_OK;
}

extern "C" int lfvio_global_sfm(int frame_num, int l, const double *relative_R, const double *relative_T, const int *obs_first,
                                const int *obs_frame, const double *obs_pts, int n_features, int run_ba, double *q, double *T,
                                double *points, unsigned char *state, lfvio_sfm_summary *summary, int device) {
    clear_error();
    LFVIO_REQUIRE(relative_R && relative_T && obs_first && obs_frame && obs_pts && q && T && summary, "lfvio_global_sfm: null argument");
    LFVIO_REQUIRE(frame_num >= 2 && frame_num <= LFVIO_SFM_MAX_FRAMES && l >= 0 && l < frame_num - 1 && n_features > 0,
                  "lfvio_global_sfm: bad sizes");
    LFVIO_TRY(check_tracks(obs_first, obs_frame, n_features, frame_num));
    LFVIO_CUDA_TRY(cudaSetDevice(device));
    SfmBuffers B;
    const int F = frame_num, L = n_features, last = F - 1;
    LFVIO_TRY(B.setup(F, l, obs_first, obs_frame, obs_pts, L));
    LFVIO_TRY(B.relR.put(relative_R, 9));
    LFVIO_TRY(B.relT.put(relative_T, 3));
    const SfmDev &d = B.dev;
    const int tg = cdiv(L, 128);
    sfm_init_kernel<<<1, 32>>>(d, B.relR.p, B.relT.p);
    LFVIO_LAUNCH_CHECK();
    auto tri = [&](int f0, int f1) -> int {
        sfm_triangulate_kernel<<<tg, 128>>>(d, f0, f1);
        LFVIO_LAUNCH_CHECK();
        return LFVIO_OK;
    };
    auto pnp = [&](int i) -> int {
        sfm_pnp_kernel<<<1, NT>>>(d, i, B.pw_list.p, B.us_list.p, B.alphas.p);
        LFVIO_LAUNCH_CHECK();
        return LFVIO_OK;
    };
    for (int i = l; i < last; i++) {  // initial_sfm.cpp:167-185
        if (i > l) LFVIO_TRY(pnp(i));
        LFVIO_TRY(tri(i, last));
    }
    for (int i = l + 1; i < last; i++) LFVIO_TRY(tri(l, i));  // :186-188
    for (int i = l - 1; i >= 0; i--) {                         // :191-206
        LFVIO_TRY(pnp(i));
        LFVIO_TRY(tri(i, l));
    }
    LFVIO_TRY(tri(-1, -1));                                    // :207-228
    if (run_ba) LFVIO_TRY(B.run_ba(F));
    sfm_finish_kernel<<<1, 64>>>(d, B.cT.p, B.q_out.p, B.T_out.p);
    LFVIO_LAUNCH_CHECK();
    int fail[2];
    LFVIO_TRY(B.fail.get(fail, 2));  // (synchronises)
    im::BaCtl c;
    LFVIO_TRY(B.ctl.get(&c, 1));
    summary->ok = fail[0] == 0, summary->fail_stage = fail[0], summary->fail_frame = fail[1];
    if (!run_ba || fail[0] == 1) {
        im::ba_init(&c);
        c.converged = run_ba ? 0 : 1;
    }
    fill_summary(c, summary);
    LFVIO_TRY(B.q_out.get(q, 4 * (size_t)F));
    LFVIO_TRY(B.T_out.get(T, 3 * (size_t)F));
    if (points) LFVIO_TRY(B.points.get(points, 3 * (size_t)L));
    if (state) LFVIO_TRY(B.state.get(state, L));
    return LFVIO_OK;
}
