// K7 — FeatureTracker::rejectWithF on device.
//
// Reference: feature_tracker/src/feature_tracker.cpp
//   rejectWithF            :329-386   lift cur/forw to the unit sphere, RANSAC, reduceVector by status
//   myfindFundamentalMat   :266-327   100 hypotheses, best = first strictly larger float score, refit on its inliers
//   compute_E_21           :186-218   null vector of the N x 9 bearing constraint matrix, rank-2 projection
//   check_inliers          :219-264   symmetric epipolar-plane residual |(E a).b| / |E a| against sin(0.5 deg) (a float)
//   create_random_array    random_array.cc:22-49
//
// The reference seeds a new mt19937 from std::random_device on every call, so its sample sets cannot be reproduced; the
// random source is injected instead: word k of hypothesis h of frame f is the top half of splitmix64(seed, f, h, k),
// consumed by the reference's own procedure (draw 9, sort, unique, keep the 8 smallest, repeat until 8).
//
// Kernels (all on the tracker stream, no host round trip; N = *n_dev is only known on the device):
//   ransac_lift_kernel        b1, b2 (f64 unit bearings), one thread per track
//   ransac_hypothesis_kernel  one warp per hypothesis: sample, 8 x 9 null vector by Gauss-Jordan elimination with
//                             complete pivoting in shared memory (Eigen's JacobiSVD null vector to rounding), rank-2
//                             projection through the smallest eigenvector of E^T E
//   ransac_score_kernel       one CTA per hypothesis: residuals in parallel (IEEE double, no FMA, reference operation
//                             order), then the reference's SEQUENTIAL float accumulation of the score by one thread, so
//                             that the best-hypothesis choice is the reference's
//   ransac_refit_kernel       one CTA: first-maximum hypothesis, A^T A over its inliers, smallest eigenvector by shifted
//                             inverse iteration, rank-2 projection, final inlier test -> status
// The tracker then compacts the track arrays with that status (compact_kernel, no track_cnt increment).
#include "frontend.cuh"

namespace lfvio {

namespace {

constexpr int kHyp = 100;   // max_num_iter
constexpr int kWarpsPerCta = 4;
constexpr int kScoreThreads = 256, kScoreChunk = 1024;
constexpr int kRefitThreads = 512;

__device__ __forceinline__ unsigned long long splitmix64(unsigned long long x) {
    unsigned long long z = x + 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
__device__ __forceinline__ unsigned rng_word(unsigned long long seed, unsigned frame, unsigned hyp, unsigned k) {
    const unsigned long long key =
        seed + frame * 0xD1B54A32D192ED03ull + hyp * 0x8CB92BA72F3D8DD7ull + k * 0xA24BAED4963EE407ull;
    return (unsigned)(splitmix64(key) >> 32);
}

// create_random_array(8, 0, n - 1): one thread
__device__ void sample_indices(unsigned long long seed, unsigned frame, unsigned hyp, int n, int *out8) {
    int v[9];
    int cnt = 0;
    unsigned k = 0;
    while (cnt != 8) {
        while (cnt < 9) v[cnt++] = (int)(((unsigned long long)rng_word(seed, frame, hyp, k++) * (unsigned long long)n) >> 32);
        for (int i = 1; i < cnt; i++) {  // insertion sort
            const int x = v[i];
            int j = i - 1;
            while (j >= 0 && v[j] > x) {
                v[j + 1] = v[j];
                j--;
            }
            v[j + 1] = x;
        }
        int m = 0;
        for (int i = 0; i < cnt; i++)
            if (i == 0 || v[i] != v[m - 1]) v[m++] = v[i];
        cnt = m > 8 ? 8 : m;
    }
    for (int i = 0; i < 8; i++) out8[i] = v[i];
}

// Smallest-eigenvalue eigenvector of the symmetric 3 x 3 matrix B by cyclic Jacobi (one thread).
__device__ void smallest_eigvec3(const double B[3][3], double v[3]) {
    double A[3][3], V[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) A[i][j] = B[i][j];
    for (int sweep = 0; sweep < 12; sweep++) {
        const double off = fabs(A[0][1]) + fabs(A[0][2]) + fabs(A[1][2]);
        if (off <= 1e-300 || off <= 1e-17 * (fabs(A[0][0]) + fabs(A[1][1]) + fabs(A[2][2]))) break;
#pragma unroll
        for (int pq = 0; pq < 3; pq++) {
            const int p = pq == 2 ? 1 : 0, q = pq == 0 ? 1 : 2;
            if (A[p][q] == 0.0) continue;
            const double theta = (A[q][q] - A[p][p]) / (2.0 * A[p][q]);
            const double t = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
            const double c = 1.0 / sqrt(t * t + 1.0), s = t * c;
            for (int k = 0; k < 3; k++) {  // A <- A J
                const double akp = A[k][p], akq = A[k][q];
                A[k][p] = c * akp - s * akq;
                A[k][q] = s * akp + c * akq;
            }
            for (int k = 0; k < 3; k++) {  // A <- J^T A
                const double apk = A[p][k], aqk = A[q][k];
                A[p][k] = c * apk - s * aqk;
                A[q][k] = s * apk + c * aqk;
            }
            for (int k = 0; k < 3; k++) {
                const double vkp = V[k][p], vkq = V[k][q];
                V[k][p] = c * vkp - s * vkq;
                V[k][q] = s * vkp + c * vkq;
            }
        }
    }
    int m = 0;
    if (A[1][1] < A[m][m]) m = 1;
    if (A[2][2] < A[m][m]) m = 2;
    for (int k = 0; k < 3; k++) v[k] = V[k][m];
}

// lambda(2) = 0 of the 3 x 3 SVD (compute_E_21 :204-214): E - (E v3) v3^T with v3 the right singular vector of the
// smallest singular value.  E is row-major 9 doubles, in place.  One thread.
__device__ void project_rank2(double *E) {
    double B[3][3], v[3];
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) B[i][j] = E[0 + i] * E[0 + j] + E[3 + i] * E[3 + j] + E[6 + i] * E[6 + j];
    smallest_eigvec3(B, v);
    for (int r = 0; r < 3; r++) {
        const double ev = E[3 * r] * v[0] + E[3 * r + 1] * v[1] + E[3 * r + 2] * v[2];
        for (int c = 0; c < 3; c++) E[3 * r + c] -= ev * v[c];
    }
}

// |(E a) . b| / |E a|, reference operation order, IEEE double without contraction (check_inliers :236-239)
__device__ __forceinline__ double residual(const double *E, double3 a, double3 b) {
    const double e0 = __dadd_rn(__dadd_rn(__dmul_rn(E[0], a.x), __dmul_rn(E[1], a.y)), __dmul_rn(E[2], a.z));
    const double e1 = __dadd_rn(__dadd_rn(__dmul_rn(E[3], a.x), __dmul_rn(E[4], a.y)), __dmul_rn(E[5], a.z));
    const double e2 = __dadd_rn(__dadd_rn(__dmul_rn(E[6], a.x), __dmul_rn(E[7], a.y)), __dmul_rn(E[8], a.z));
    const double dot = __dadd_rn(__dadd_rn(__dmul_rn(e0, b.x), __dmul_rn(e1, b.y)), __dmul_rn(e2, b.z));
    const double nrm = __dsqrt_rn(__dadd_rn(__dadd_rn(__dmul_rn(e0, e0), __dmul_rn(e1, e1)), __dmul_rn(e2, e2)));
    return fabs(__ddiv_rn(dot, nrm));
}

// One point against E_21 (E) and E_12 (Et): code 0 = fails in image 2, 1 = passes 2 / fails 1, 2 = inlier; t2, t1 are
// the squared margins the score accumulates (:241-262).
__device__ __forceinline__ int check_point(const double *E, const double *Et, double3 b1, double3 b2, double thr, double *t2,
                                           double *t1) {
    const double r2 = residual(E, b1, b2);
    if (thr < r2) return 0;
    const double d2 = __dadd_rn(thr, -r2);
    *t2 = __dmul_rn(d2, d2);
    const double r1 = residual(Et, b2, b1);
    if (thr < r1) return 1;
    const double d1 = __dadd_rn(thr, -r1);
    *t1 = __dmul_rn(d1, d1);
    return 2;
}

__device__ __forceinline__ double thr_value() { return (double)0.00872653549837f; }  // constexpr float residual_cos_thr

}  // namespace

// ---------------------------------------------------------------------------------------------------------------
__global__ void ransac_lift_kernel(OcamDev cam, const float2 *__restrict__ cur, const float2 *__restrict__ forw,
                                   const int *n_dev, int cap, double3 *b1, double3 *b2) {
    const int n = min(*n_dev, cap);
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float2 p = cur[i], q = forw[i];
    b1[i] = ocam_lift(cam, (double)p.x, (double)p.y, true);
    b2[i] = ocam_lift(cam, (double)q.x, (double)q.y, true);
}

// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kWarpsPerCta * 32) ransac_hypothesis_kernel(const double3 *__restrict__ b1,
                                                                              const double3 *__restrict__ b2, const int *n_dev,
                                                                              int cap, unsigned long long seed, unsigned frame,
                                                                              const unsigned *frame_dev,
                                                                              double *E_out /* kHyp x 9 */) {
    __shared__ double s_A[kWarpsPerCta][8][9];
    __shared__ int s_idx[kWarpsPerCta][8];
    if (frame_dev) frame = *frame_dev;  // per-frame parameter block (the launch itself can then be replayed from a CUDA graph)
    const int n = min(*n_dev, cap);
    if (n < 8) return;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int hyp = blockIdx.x * kWarpsPerCta + w;
    if (hyp >= kHyp) return;
    double(*A)[9] = s_A[w];
    if (lane == 0) sample_indices(seed, frame, hyp, n, s_idx[w]);
    __syncwarp();
    if (lane < 8) {  // row i: b2(r) * b1(c) at column 3r + c (compute_E_21 :194-199)
        const double3 p = b1[s_idx[w][lane]], q = b2[s_idx[w][lane]];
        A[lane][0] = q.x * p.x, A[lane][1] = q.x * p.y, A[lane][2] = q.x * p.z;
        A[lane][3] = q.y * p.x, A[lane][4] = q.y * p.y, A[lane][5] = q.y * p.z;
        A[lane][6] = q.z * p.x, A[lane][7] = q.z * p.y, A[lane][8] = q.z * p.z;
    }
    __syncwarp();
    // Gauss-Jordan with complete pivoting; lane l owns elements l, l + 32, l + 64 of the 72
    unsigned row_used = 0, col_used = 0;
    int prow[8], pcol[8];
#pragma unroll
    for (int k = 0; k < 8; k++) {
        double best = -1.0;
        int bi = 0;
#pragma unroll
        for (int q = 0; q < 3; q++) {
            const int e = lane + 32 * q;
            if (e < 72) {
                const int r = e / 9, c = e - 9 * r;
                const double v = fabs(A[r][c]);
                if (!((row_used >> r) & 1) && !((col_used >> c) & 1) && v > best) best = v, bi = e;
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const double ob = __shfl_xor_sync(0xffffffffu, best, o);
            const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
            if (ob > best || (ob == best && oi < bi)) best = ob, bi = oi;
        }
        const int pr = bi / 9, pc = bi - 9 * pr;
        prow[k] = pr, pcol[k] = pc;
        row_used |= 1u << pr, col_used |= 1u << pc;
        const double inv = 1.0 / A[pr][pc];
        double f[3], pv[3];
#pragma unroll
        for (int q = 0; q < 3; q++) {
            const int e = lane + 32 * q;
            if (e < 72) {
                const int r = e / 9, c = e - 9 * r;
                f[q] = A[r][pc] * inv;
                pv[q] = A[pr][c];
            }
        }
        __syncwarp();
#pragma unroll
        for (int q = 0; q < 3; q++) {
            const int e = lane + 32 * q;
            if (e < 72) {
                const int r = e / 9, c = e - 9 * r;
                if (r != pr) A[r][c] = (c == pc) ? 0.0 : A[r][c] - f[q] * pv[q];
            }
        }
        __syncwarp();
    }
    if (lane == 0) {
        int fc = 0;
        while ((col_used >> fc) & 1) fc++;
        double v[9];
        v[fc] = 1.0;
        double nn = 1.0;
#pragma unroll
        for (int k = 0; k < 8; k++) {
            const double x = -A[prow[k]][fc] / A[prow[k]][pcol[k]];
            v[pcol[k]] = x;
            nn += x * x;
        }
        const double s = 1.0 / sqrt(nn);
        double E[9];
        for (int i = 0; i < 9; i++) E[i] = v[i] * s;  // Matrix3d(v.data()).transpose(): row-major
        project_rank2(E);
        for (int i = 0; i < 9; i++) E_out[hyp * 9 + i] = E[i];
    }
}

// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kScoreThreads) ransac_score_kernel(const double3 *__restrict__ b1, const double3 *__restrict__ b2,
                                                                     const int *n_dev, int cap, const double *__restrict__ E_all,
                                                                     float *score_out, uint8_t *flags /* kHyp x cap */) {
    __shared__ double s_t2[kScoreChunk], s_t1[kScoreChunk];
    __shared__ uint8_t s_code[kScoreChunk];
    __shared__ double s_E[9], s_Et[9];
    const int n = min(*n_dev, cap);
    if (n < 8) return;
    const int hyp = blockIdx.x;
    if (threadIdx.x < 9) {
        const int r = threadIdx.x / 3, c = threadIdx.x % 3;
        const double v = E_all[hyp * 9 + threadIdx.x];
        s_E[threadIdx.x] = v;
        s_Et[3 * c + r] = v;
    }
    __syncthreads();
    const double thr = thr_value();
    float score = 0.f;
    for (int i0 = 0; i0 < n; i0 += kScoreChunk) {
        const int m = min(kScoreChunk, n - i0);
        for (int j = threadIdx.x; j < m; j += kScoreThreads) {
            double t2 = 0, t1 = 0;
            const int code = check_point(s_E, s_Et, b1[i0 + j], b2[i0 + j], thr, &t2, &t1);
            s_t2[j] = t2, s_t1[j] = t1, s_code[j] = (uint8_t)code;
            flags[(size_t)hyp * cap + i0 + j] = code == 2;
        }
        __syncthreads();
        if (threadIdx.x == 0) {  // float score; score += pow(thr - res, 2.0), in point order (:244, :259)
            for (int j = 0; j < m; j++) {
                const int code = s_code[j];
                if (code >= 1) score = (float)((double)score + s_t2[j]);
                if (code == 2) score = (float)((double)score + s_t1[j]);
            }
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) score_out[hyp] = score;
}

// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kRefitThreads) ransac_refit_kernel(const double3 *__restrict__ b1, const double3 *__restrict__ b2,
                                                                     const int *n_dev, int cap, const float *__restrict__ score,
                                                                     const uint8_t *__restrict__ flags, double *E_best /* 9 */,
                                                                     int *best_out, uint8_t *status) {
    constexpr int kChunks = kRefitThreads / 45;  // 11
    __shared__ double s_part[kChunks][45];
    __shared__ double s_B[9][9], s_E[9], s_Et[9];
    __shared__ int s_best;
    const int n = min(*n_dev, cap);
    if (n < 8) {  // rejectWithF does nothing below 8 points (:331)
        for (int i = threadIdx.x; i < n; i += kRefitThreads) status[i] = 1;
        return;
    }
    if (threadIdx.x < 32) {  // best_score_ < score_in_sac from best_score_ = 0: first strict maximum, NaN never wins
        float bs = 0.f;
        int bi = -1;
        for (int h = threadIdx.x; h < kHyp; h += 32) {
            const float s = score[h];
            if (s > bs) bs = s, bi = h;
        }
        for (int o = 16; o > 0; o >>= 1) {
            const float os = __shfl_xor_sync(0xffffffffu, bs, o);
            const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
            if (oi >= 0 && (os > bs || (os == bs && (bi < 0 || oi < bi)))) bs = os, bi = oi;
        }
        if (threadIdx.x == 0) {
            s_best = bi < 0 ? 0 : bi;  // (the reference indexes an empty vector in that case)
            *best_out = s_best;
        }
    }
    __syncthreads();
    const uint8_t *inl = flags + (size_t)s_best * cap;
    // A^T A of the inlier rows, upper triangle: entry e = (a, b), a <= b; thread (chunk, e)
    if (threadIdx.x < kChunks * 45) {
        const int e = threadIdx.x % 45, chunk = threadIdx.x / 45;
        int a = 0, rem = e;
        while (rem >= 9 - a) rem -= 9 - a, a++;
        const int b = a + rem;
        const int ar = a / 3, ac = a % 3, br = b / 3, bc = b % 3;
        double acc = 0;
        for (int i = chunk; i < n; i += kChunks) {
            if (!inl[i]) continue;
            const double3 p = b1[i], q = b2[i];
            const double qa = ar == 0 ? q.x : (ar == 1 ? q.y : q.z), pa = ac == 0 ? p.x : (ac == 1 ? p.y : p.z);
            const double qb = br == 0 ? q.x : (br == 1 ? q.y : q.z), pb = bc == 0 ? p.x : (bc == 1 ? p.y : p.z);
            acc += (qa * pa) * (qb * pb);
        }
        s_part[chunk][e] = acc;
    }
    __syncthreads();
    if (threadIdx.x < 45) {
        const int e = threadIdx.x;
        int a = 0, rem = e;
        while (rem >= 9 - a) rem -= 9 - a, a++;
        const int b = a + rem;
        double acc = 0;
        for (int c = 0; c < kChunks; c++) acc += s_part[c][e];
        s_B[a][b] = acc;
        s_B[b][a] = acc;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        // smallest eigenvector of B = A^T A (the right singular vector V.col(8) of compute_E_21): inverse iteration on
        // B + delta I (Cholesky), which keeps B's eigenvectors and is positive definite for an exactly singular B
        double tr = 0;
        for (int i = 0; i < 9; i++) tr += s_B[i][i];
        const double delta = 1e-15 * tr + 1e-300;
        double L[9][9];
        bool ok = true;
        for (int j = 0; j < 9; j++) {
            double d = s_B[j][j] + delta;
            for (int k = 0; k < j; k++) d -= L[j][k] * L[j][k];
            if (!(d > 0.0)) ok = false, d = delta;
            const double lj = sqrt(d);
            L[j][j] = lj;
            for (int i = j + 1; i < 9; i++) {
                double s = s_B[i][j];
                for (int k = 0; k < j; k++) s -= L[i][k] * L[j][k];
                L[i][j] = s / lj;
            }
        }
        double x[9];
        for (int i = 0; i < 9; i++) x[i] = 1.0 + 0.37 * i;
        for (int it = 0; it < (ok ? 16 : 0); it++) {
            for (int i = 0; i < 9; i++) {  // L y = x
                double s = x[i];
                for (int k = 0; k < i; k++) s -= L[i][k] * x[k];
                x[i] = s / L[i][i];
            }
            for (int i = 8; i >= 0; i--) {  // L^T z = y
                double s = x[i];
                for (int k = i + 1; k < 9; k++) s -= L[k][i] * x[k];
                x[i] = s / L[i][i];
            }
            double nn = 0;
            for (int i = 0; i < 9; i++) nn += x[i] * x[i];
            const double sc = 1.0 / sqrt(nn);
            for (int i = 0; i < 9; i++) x[i] *= sc;
        }
        double E[9];
        for (int i = 0; i < 9; i++) E[i] = x[i];
        project_rank2(E);
        for (int i = 0; i < 9; i++) {
            s_E[i] = E[i];
            s_Et[3 * (i % 3) + i / 3] = E[i];
            E_best[i] = E[i];
        }
    }
    __syncthreads();
    const double thr = thr_value();
    for (int i = threadIdx.x; i < n; i += kRefitThreads) {
        double t2, t1;
        status[i] = check_point(s_E, s_Et, b1[i], b2[i], thr, &t2, &t1) == 2;
    }
}

// ---------------------------------------------------------------------------------------------------------------
int launch_reject_with_f(const OcamDev &cam, const float2 *cur, const float2 *forw, const int *n_dev, int n_upper, int cap,
                         unsigned long long seed, unsigned frame, const RansacScratch &rs, uint8_t *status, cudaStream_t s,
                         const unsigned *frame_dev) {
    if (n_upper <= 0) return LFVIO_OK;
    ransac_lift_kernel<<<cdiv(n_upper, 128), 128, 0, s>>>(cam, cur, forw, n_dev, cap, rs.b1, rs.b2);
    LFVIO_LAUNCH_CHECK();
    ransac_hypothesis_kernel<<<cdiv(kHyp, kWarpsPerCta), kWarpsPerCta * 32, 0, s>>>(rs.b1, rs.b2, n_dev, cap, seed, frame, frame_dev, rs.E);
    LFVIO_LAUNCH_CHECK();
    ransac_score_kernel<<<kHyp, kScoreThreads, 0, s>>>(rs.b1, rs.b2, n_dev, cap, rs.E, rs.score, rs.flags);
    LFVIO_LAUNCH_CHECK();
    ransac_refit_kernel<<<1, kRefitThreads, 0, s>>>(rs.b1, rs.b2, n_dev, cap, rs.score, rs.flags, rs.E_best, rs.best, status);
    LFVIO_LAUNCH_CHECK();
    return LFVIO_OK;
}

// myfindFundamentalMat on bearings that are already on the device (MotionEstimator::myfindFundamentalMat,
// vins_estimator/src/initial/solve_5pts.cpp:266-335 — the same code as the tracker's copy): hypotheses, scores, refit.
int launch_ransac_bearings(const int *n_dev, int n_upper, int cap, unsigned long long seed, unsigned frame, const RansacScratch &rs,
                           uint8_t *status, cudaStream_t s) {
    if (n_upper <= 0) return LFVIO_OK;
    ransac_hypothesis_kernel<<<cdiv(kHyp, kWarpsPerCta), kWarpsPerCta * 32, 0, s>>>(rs.b1, rs.b2, n_dev, cap, seed, frame, nullptr, rs.E);
    LFVIO_LAUNCH_CHECK();
    ransac_score_kernel<<<kHyp, kScoreThreads, 0, s>>>(rs.b1, rs.b2, n_dev, cap, rs.E, rs.score, rs.flags);
    LFVIO_LAUNCH_CHECK();
    ransac_refit_kernel<<<1, kRefitThreads, 0, s>>>(rs.b1, rs.b2, n_dev, cap, rs.score, rs.flags, rs.E_best, rs.best, status);
    LFVIO_LAUNCH_CHECK();
    return LFVIO_OK;
}

}  // namespace lfvio
