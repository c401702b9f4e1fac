// K15/K16 — the two steps right before Estimator::optimization(), batched on device.
//
//   preintegrate_kernel   IntegrationBase::push_back / propagate / midPointIntegration
//                         (vins_estimator/src/factor/integration_base.h:31-157) for whole inter-frame intervals, plus the
//                         nominal-state prediction of Estimator::processIMU (vins_estimator/src/estimator.cpp:86-120).
//                         One CTA per interval: the samples are inherently sequential; per sample one warp forms the
//                         3x3 blocks of F (15x15) and V (15x18), then 225 threads take one entry each of
//                         jacobian <- F jacobian and covariance <- F covariance F^T + V noise V^T in shared memory.
//                         The output record is exactly lfvio_imu_factor, i.e. what IMUFactor (imu_factor.h:19-200) reads.
//   triangulate_kernel    FeatureManager::triangulate (vins_estimator/src/feature_manager.cpp:199-253): per landmark the
//                         (2K x 4) system of its K bearing observations; the right singular vector of the smallest
//                         singular value by a ONE-SIDED Jacobi SVD on the matrix itself (no A^T A squaring), one warp per
//                         landmark with the rows spread over the lanes and the column inner products reduced by
//                         shuffles.  depth = (X / X_w) . pts_0 (range along the bearing), INIT_DEPTH when negative.
#include "backend.cuh"

namespace lfvio {

namespace {

constexpr int PI_THREADS = 256;

struct M3 {
    double m[9];
};
__device__ __forceinline__ M3 m3_from(const DM3 &a) {
    M3 r;
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) r.m[3 * i + j] = a.m[i][j];
    return r;
}

// writes the 3x3 block b (row-major 9) scaled by s into M (row-major, ld columns) at (r0, c0)
__device__ __forceinline__ void put3(double *M, int ld, int r0, int c0, const DM3 &b, double s) {
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) M[(r0 + i) * ld + c0 + j] = s * b.m[i][j];
}

__global__ void __launch_bounds__(PI_THREADS) preintegrate_kernel(const lfvio_imu_interval *iv, const double *dt_all,
                                                                  const double *acc_all, const double *gyr_all,
                                                                  lfvio_imu_noise noise, double gx, double gy, double gz,
                                                                  lfvio_imu_factor *out, double *state_out) {
    __shared__ double sJ[225], sP[225], sF[225], sV[270], sT[225], sT2[225];
    __shared__ double s_nz[18];
    const int tid = threadIdx.x;
    const lfvio_imu_interval I = iv[blockIdx.x];
    for (int q = tid; q < 225; q += PI_THREADS) {
        sJ[q] = (q / 15 == q % 15) ? 1.0 : 0.0;
        sP[q] = 0.0;
    }
    if (tid < 18) {
        const int b = tid / 3;
        const double n = (b == 0 || b == 2) ? noise.acc_n : ((b == 1 || b == 3) ? noise.gyr_n : (b == 4 ? noise.acc_w : noise.gyr_w));
        s_nz[tid] = n * n;
    }
    // per-interval running values live in thread 0's registers
    D3 acc0 = d3(I.acc_0[0], I.acc_0[1], I.acc_0[2]), gyr0 = d3(I.gyr_0[0], I.gyr_0[1], I.gyr_0[2]);
    const D3 ba = d3(I.ba[0], I.ba[1], I.ba[2]), bg = d3(I.bg[0], I.bg[1], I.bg[2]);
    D3 dp = d3(0, 0, 0), dv = d3(0, 0, 0);
    DQ dqq = dq(1, 0, 0, 0);
    double sum_dt = 0;
    // Estimator::processIMU nominal state (estimator.cpp:107-116)
    D3 Ps = d3(I.P[0], I.P[1], I.P[2]), Vs = d3(I.V[0], I.V[1], I.V[2]);
    DM3 Rs;
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) Rs.m[i][j] = I.R[3 * i + j];
    const D3 g = d3(gx, gy, gz);
    __syncthreads();
    for (int sidx = 0; sidx < I.count; sidx++) {
        const int k = I.first + sidx;
        const double dt = dt_all[k];
        if (tid == 0) {
            const D3 acc1 = d3(acc_all[3 * k], acc_all[3 * k + 1], acc_all[3 * k + 2]);
            const D3 gyr1 = d3(gyr_all[3 * k], gyr_all[3 * k + 1], gyr_all[3 * k + 2]);
            // midPointIntegration (integration_base.h:65-76)
            const D3 un_acc_0 = qapply(dqq, acc0 - ba);
            const D3 un_gyr = 0.5 * (gyr0 + gyr1) - bg;
            const DQ rq = dqq * dq(1.0, un_gyr.x * dt / 2, un_gyr.y * dt / 2, un_gyr.z * dt / 2);
            const D3 un_acc_1 = qapply(rq, acc1 - ba);
            const D3 un_acc = 0.5 * (un_acc_0 + un_acc_1);
            const D3 rdp = dp + dv * dt + 0.5 * un_acc * dt * dt;
            const D3 rdv = dv + un_acc * dt;
            // Jacobian / noise blocks (:78-127)
            const DM3 R_w_x = skew3(un_gyr), R_a_0_x = skew3(acc0 - ba), R_a_1_x = skew3(acc1 - ba);
            const DM3 R0 = qrot(dqq), R1 = qrot(rq);  // toRotationMatrix() of the not yet normalised product, as in the reference
            const DM3 I3 = dm3_identity();
            const DM3 IwT = I3 - dt * R_w_x;
            const DM3 R0a0 = R0 * R_a_0_x, R1a1 = R1 * R_a_1_x, R1a1Iw = R1a1 * IwT;
            for (int q = 0; q < 225; q++) sF[q] = 0.0;
            for (int q = 0; q < 270; q++) sV[q] = 0.0;
            put3(sF, 15, 0, 0, I3, 1.0);
            put3(sF, 15, 0, 3, (-0.25 * dt * dt) * R0a0 + (-0.25 * dt * dt) * R1a1Iw, 1.0);
            put3(sF, 15, 0, 6, I3, dt);
            put3(sF, 15, 0, 9, R0 + R1, -0.25 * dt * dt);
            put3(sF, 15, 0, 12, R1a1, -0.25 * dt * dt * -dt);
            put3(sF, 15, 3, 3, IwT, 1.0);
            put3(sF, 15, 3, 12, I3, -1.0 * dt);
            put3(sF, 15, 6, 3, (-0.5 * dt) * R0a0 + (-0.5 * dt) * R1a1Iw, 1.0);
            put3(sF, 15, 6, 6, I3, 1.0);
            put3(sF, 15, 6, 9, R0 + R1, -0.5 * dt);
            put3(sF, 15, 6, 12, R1a1, -0.5 * dt * -dt);
            put3(sF, 15, 9, 9, I3, 1.0);
            put3(sF, 15, 12, 12, I3, 1.0);
            put3(sV, 18, 0, 0, R0, 0.25 * dt * dt);
            put3(sV, 18, 0, 3, R1a1, -0.25 * dt * dt * 0.5 * dt);
            put3(sV, 18, 0, 6, R1, 0.25 * dt * dt);
            put3(sV, 18, 0, 9, R1a1, -0.25 * dt * dt * 0.5 * dt);
            put3(sV, 18, 3, 3, I3, 0.5 * dt);
            put3(sV, 18, 3, 9, I3, 0.5 * dt);
            put3(sV, 18, 6, 0, R0, 0.5 * dt);
            put3(sV, 18, 6, 3, R1a1, -0.5 * dt * 0.5 * dt);
            put3(sV, 18, 6, 6, R1, 0.5 * dt);
            put3(sV, 18, 6, 9, R1a1, -0.5 * dt * 0.5 * dt);
            put3(sV, 18, 9, 12, I3, dt);
            put3(sV, 18, 12, 15, I3, dt);
            // propagate() tail (:146-156)
            dp = rdp;
            dv = rdv;
            dqq = qnormalized(rq);
            sum_dt += dt;
            if (I.propagate) {  // estimator.cpp:107-116 (uses the same acc_0 / gyr_0 pair and the frame's own biases)
                const D3 w_acc_0 = Rs * (acc0 - ba) - g;
                const DQ dR = delta_q(un_gyr * dt);
                Rs = Rs * qrot(dR);
                const D3 w_acc_1 = Rs * (acc1 - ba) - g;
                const D3 w_acc = 0.5 * (w_acc_0 + w_acc_1);
                Ps = Ps + dt * Vs + 0.5 * dt * dt * w_acc;
                Vs = Vs + dt * w_acc;
            }
            acc0 = acc1;
            gyr0 = gyr1;
        }
        __syncthreads();
        // jacobian <- F jacobian ; T <- F covariance
        if (tid < 225) {
            const int r = tid / 15, c = tid % 15;
            double a = 0, b = 0;
#pragma unroll
            for (int q = 0; q < 15; q++) {
                a = fma(sF[r * 15 + q], sJ[q * 15 + c], a);
                b = fma(sF[r * 15 + q], sP[q * 15 + c], b);
            }
            sT[tid] = a;
            sT2[tid] = b;
        }
        __syncthreads();
        // covariance <- T F^T + V noise V^T
        if (tid < 225) {
            const int r = tid / 15, c = tid % 15;
            double a = 0, b = 0;
#pragma unroll
            for (int q = 0; q < 15; q++) a = fma(sT2[r * 15 + q], sF[c * 15 + q], a);
#pragma unroll
            for (int q = 0; q < 18; q++) b = fma(sV[r * 18 + q] * s_nz[q], sV[c * 18 + q], b);
            sP[tid] = a + b;
            sJ[tid] = sT[tid];
        }
        __syncthreads();
    }
    lfvio_imu_factor &o = out[blockIdx.x];
    for (int q = tid; q < 225; q += PI_THREADS) {
        o.jacobian[q] = sJ[q];
        o.covariance[q] = sP[q];
    }
    if (tid == 0) {
        o.frame_i = I.frame_i;
        o.reserved = 0;
        o.sum_dt = sum_dt;
        o.delta_p[0] = dp.x; o.delta_p[1] = dp.y; o.delta_p[2] = dp.z;
        o.delta_q[0] = dqq.w; o.delta_q[1] = dqq.x; o.delta_q[2] = dqq.y; o.delta_q[3] = dqq.z;
        o.delta_v[0] = dv.x; o.delta_v[1] = dv.y; o.delta_v[2] = dv.z;
        for (int a = 0; a < 3; a++) {
            o.linearized_ba[a] = I.ba[a];
            o.linearized_bg[a] = I.bg[a];
        }
        if (state_out) {
            double *so = state_out + 15 * (size_t)blockIdx.x;
            so[0] = Ps.x; so[1] = Ps.y; so[2] = Ps.z;
            so[3] = Vs.x; so[4] = Vs.y; so[5] = Vs.z;
            for (int i = 0; i < 3; i++)
                for (int j = 0; j < 3; j++) so[6 + 3 * i + j] = Rs.m[i][j];
        }
    }
}

// ---- triangulation -----------------------------------------------------------------------------------------
constexpr int TRI_WARPS = 4;
constexpr int TRI_MAX_ROWS = 64;  // 2 K <= 64: up to 32 observations (kMaxFrames)

__global__ void __launch_bounds__(TRI_WARPS * 32) triangulate_kernel(const double *Ps, const double *Rs, const double *tic,
                                                                    const double *ric, const int *start_frame,
                                                                    const int *obs_first, const double *obs_pts, int n,
                                                                    int n_frames, double init_depth, double *depth,
                                                                    const int *n_obs, int rec) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int l = blockIdx.x * TRI_WARPS + wid;
    if (l >= n) return;
    if (depth[l] > 0) return;  // feature_manager.cpp:207-208: already has a depth
    // packed tracks (obs_first) or the device FeatureManager's slots (n_obs: n_frames records of `rec` doubles per slot)
    const int o0 = n_obs ? l * n_frames : obs_first[l], K = n_obs ? n_obs[l] : obs_first[l + 1] - o0;
    if (K < 2 || 2 * K > TRI_MAX_ROWS) return;  // used_num >= 2 (:204); longer tracks cannot exist (K <= F)
    const int f0 = start_frame[l];
    if (!(f0 < (n_frames - 1) - 2)) return;     // start_frame < WINDOW_SIZE - 2 (:204)
    DM3 Ric, R0w;
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) {
            Ric.m[i][j] = ric[3 * i + j];
            R0w.m[i][j] = Rs[9 * (size_t)f0 + 3 * i + j];
        }
    const D3 Tic = d3(tic[0], tic[1], tic[2]);
    const D3 t0 = d3(Ps[3 * f0], Ps[3 * f0 + 1], Ps[3 * f0 + 2]) + R0w * Tic;  // :215-216
    const DM3 R0 = R0w * Ric;
    // lane r owns rows r and r + 32 of svd_A (row 2k, 2k+1 belong to observation k)
    double a[2][4];
#pragma unroll
    for (int h = 0; h < 2; h++) {
        const int row = lane + 32 * h, k = row >> 1;
#pragma unroll
        for (int c = 0; c < 4; c++) a[h][c] = 0.0;
        if (row < 2 * K && f0 + k < n_frames) {
            const int fj = f0 + k;
            DM3 Rjw;
            for (int i = 0; i < 3; i++)
                for (int j = 0; j < 3; j++) Rjw.m[i][j] = Rs[9 * (size_t)fj + 3 * i + j];
            const D3 t1 = d3(Ps[3 * fj], Ps[3 * fj + 1], Ps[3 * fj + 2]) + Rjw * Tic;
            const DM3 R1 = Rjw * Ric;
            const D3 t = transpose3(R0) * (t1 - t0);
            const DM3 R = transpose3(R0) * R1;
            const DM3 Rt = transpose3(R);       // P.leftCols<3>()
            const D3 pt = -(Rt * t);            // P.rightCols<1>()
            const D3 f = unit3(d3(obs_pts[rec * (size_t)(o0 + k)], obs_pts[rec * (size_t)(o0 + k) + 1], obs_pts[rec * (size_t)(o0 + k) + 2]));
            const int sel = row & 1;  // row 2k: f0 P2 - f2 P0 ; row 2k+1: f1 P2 - f2 P1   (:234-235)
            const double fa = sel ? f.y : f.x;
#pragma unroll
            for (int c = 0; c < 3; c++) a[h][c] = fa * Rt.m[2][c] - f.z * Rt.m[sel][c];
            a[h][3] = fa * pt.z - f.z * (sel ? pt.y : pt.x);
        }
    }
    // one-sided Jacobi (Hestenes): rotate column pairs until mutually orthogonal; V accumulates the rotations
    double V[4][4];
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) V[i][j] = (i == j) ? 1.0 : 0.0;
    for (int sweep = 0; sweep < 30; sweep++) {
        double off = 0.0;
#pragma unroll
        for (int p = 0; p < 3; p++)
#pragma unroll
            for (int q = p + 1; q < 4; q++) {
                double al = a[0][p] * a[0][p] + a[1][p] * a[1][p];
                double be = a[0][q] * a[0][q] + a[1][q] * a[1][q];
                double ga = a[0][p] * a[0][q] + a[1][p] * a[1][q];
                al = warp_sum(al);
                be = warp_sum(be);
                ga = warp_sum(ga);
                if (ga == 0.0 || fabs(ga) <= 1e-17 * sqrt(al * be)) continue;
                off = fmax(off, fabs(ga) / sqrt(al * be));
                const double zeta = (be - al) / (2.0 * ga);
                const double t = (zeta >= 0 ? 1.0 : -1.0) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
                const double c = 1.0 / sqrt(1.0 + t * t), s = c * t;
#pragma unroll
                for (int h = 0; h < 2; h++) {
                    const double ap = a[h][p], aq = a[h][q];
                    a[h][p] = c * ap - s * aq;
                    a[h][q] = s * ap + c * aq;
                }
#pragma unroll
                for (int i = 0; i < 4; i++) {
                    const double vp = V[i][p], vq = V[i][q];
                    V[i][p] = c * vp - s * vq;
                    V[i][q] = s * vp + c * vq;
                }
            }
        if (off < 1e-15) break;  // (uniform across the warp: every lane sees the same reduced values)
    }
    // smallest singular value = smallest column norm
    double best = 1e300;
    int bi = 0;
#pragma unroll
    for (int c = 0; c < 4; c++) {
        double nn = a[0][c] * a[0][c] + a[1][c] * a[1][c];
        nn = warp_sum(nn);
        if (nn < best) {
            best = nn;
            bi = c;
        }
    }
    if (lane == 0) {
        double X[4];
#pragma unroll
        for (int i = 0; i < 4; i++) X[i] = (bi == 0) ? V[i][0] : (bi == 1 ? V[i][1] : (bi == 2 ? V[i][2] : V[i][3]));
        const double px = obs_pts[rec * (size_t)o0], py = obs_pts[rec * (size_t)o0 + 1], pz = obs_pts[rec * (size_t)o0 + 2];
        double d = (X[0] / X[3]) * px + (X[1] / X[3]) * py + (X[2] / X[3]) * pz;  // :243-245 (range along the bearing)
        if (d < 0) d = init_depth;                                                 // :246-249
        depth[l] = d;
    }
}

// Per-thread arena of the two stage-level entry points below (the host Estimator calls both once per image): one device
// buffer, one page-locked mirror of it and one stream, grown on demand and kept.  A call packs its inputs into the mirror, sends
// them in ONE asynchronous copy, launches on the arena's stream, brings the outputs back in one copy and synchronises that stream
// only — no cudaMalloc / cudaFree and no default-stream synchronisation per call (round 1: 6-8 of each, which serialised the
// estimator against the tracker and window streams).
struct StepArena {
    int device = -1;
    cudaStream_t s = nullptr;
    char *d = nullptr, *h = nullptr;
    size_t cap = 0;
    void release() {
        if (device >= 0) cudaSetDevice(device);
        if (d) cudaFree(d);
        if (h) cudaFreeHost(h);
        if (s) cudaStreamDestroy(s);
        d = h = nullptr;
        s = nullptr;
        cap = 0;
        device = -1;
    }
    ~StepArena() { release(); }
    int ensure(int dev, size_t bytes) {
        if (dev != device) release();
        if (!s) {
            device = dev;
            LFVIO_CUDA_TRY(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
        }
        if (bytes > cap) {
            if (d) cudaFree(d);
            if (h) cudaFreeHost(h);
            d = h = nullptr;
            cap = 0;
            const size_t want = std::max<size_t>(bytes + bytes / 2, 1 << 16);
            LFVIO_CUDA_TRY(cudaMalloc(&d, want));
            LFVIO_CUDA_TRY(cudaMallocHost(&h, want));
            LFVIO_CUDA_TRY(cudaMemset(d, 0, want));  // (the alignment gaps between the sub-buffers travel in the one D2H copy)
            memset(h, 0, want);
            cap = want;
        }
        return LFVIO_OK;
    }
};
static thread_local StepArena t_arena;
struct Packer {  // sub-buffers at 256-byte offsets, same offsets on both sides
    size_t off = 0;
    size_t take(size_t bytes) {
        const size_t o = off;
        off = (off + bytes + 255) & ~(size_t)255;
        return o;
    }
};

}  // namespace

int launch_preintegrate(const lfvio_imu_interval *iv_dev, int n_intervals, const double *dt_dev, const double *acc_dev,
                        const double *gyr_dev, const lfvio_imu_noise &noise, const double *g, lfvio_imu_factor *out_dev,
                        double *state_out_dev, cudaStream_t s) {
    if (n_intervals <= 0) return LFVIO_OK;
    preintegrate_kernel<<<n_intervals, PI_THREADS, 0, s>>>(iv_dev, dt_dev, acc_dev, gyr_dev, noise, g[0], g[1], g[2], out_dev,
                                                           state_out_dev);
    LFVIO_LAUNCH_CHECK();
    return LFVIO_OK;
}

int launch_triangulate(const double *Ps, const double *Rs, const double *tic, const double *ric, const int *start_frame,
                       const int *obs_first, const double *obs_pts, int n, int n_frames, double init_depth, double *depth,
                       cudaStream_t s) {
    if (n <= 0) return LFVIO_OK;
    triangulate_kernel<<<cdiv(n, TRI_WARPS), TRI_WARPS * 32, 0, s>>>(Ps, Rs, tic, ric, start_frame, obs_first, obs_pts, n, n_frames,
                                                                    init_depth, depth, nullptr, 3);
    LFVIO_LAUNCH_CHECK();
    return LFVIO_OK;
}

int launch_triangulate_slots(const double *Ps, const double *Rs, const double *tic, const double *ric, const int *start_frame,
                             const int *n_obs, const double *obs, int rec, int n_slots, int n_frames, double init_depth, double *depth,
                             cudaStream_t s) {
    if (n_slots <= 0) return LFVIO_OK;
    triangulate_kernel<<<cdiv(n_slots, TRI_WARPS), TRI_WARPS * 32, 0, s>>>(Ps, Rs, tic, ric, start_frame, nullptr, obs, n_slots, n_frames,
                                                                          init_depth, depth, n_obs, rec);
    LFVIO_LAUNCH_CHECK();
    return LFVIO_OK;
}

}  // namespace lfvio

using namespace lfvio;

extern "C" int lfvio_preintegrate(const lfvio_imu_interval *intervals, int n_intervals, const double *dt, const double *acc,
                                  const double *gyr, int n_samples, const lfvio_imu_noise *noise, const double *g,
                                  lfvio_imu_factor *out, double *state_out, int device) {
    clear_error();
    LFVIO_REQUIRE(n_intervals >= 0 && n_samples >= 0 && noise && g, "lfvio_preintegrate: bad arguments");
    if (n_intervals == 0) return LFVIO_OK;
    LFVIO_REQUIRE(intervals && out && (n_samples == 0 || (dt && acc && gyr)), "lfvio_preintegrate: null buffer");
    for (int i = 0; i < n_intervals; i++)
        LFVIO_REQUIRE(intervals[i].first >= 0 && intervals[i].count >= 0 && intervals[i].first + intervals[i].count <= n_samples,
                      "lfvio_preintegrate: interval outside the sample arrays");
    LFVIO_CUDA_TRY(cudaSetDevice(device));
    Packer pk;
    const size_t o_iv = pk.take(sizeof(lfvio_imu_interval) * n_intervals), o_dt = pk.take(sizeof(double) * n_samples);
    const size_t o_acc = pk.take(sizeof(double) * 3 * n_samples), o_gyr = pk.take(sizeof(double) * 3 * n_samples);
    const size_t in_bytes = pk.off;
    const size_t o_out = pk.take(sizeof(lfvio_imu_factor) * n_intervals), o_state = pk.take(sizeof(double) * 15 * n_intervals);
    StepArena &A = t_arena;
    LFVIO_TRY(A.ensure(device, pk.off));
    memcpy(A.h + o_iv, intervals, sizeof(lfvio_imu_interval) * n_intervals);
    if (n_samples) {
        memcpy(A.h + o_dt, dt, sizeof(double) * n_samples);
        memcpy(A.h + o_acc, acc, sizeof(double) * 3 * n_samples);
        memcpy(A.h + o_gyr, gyr, sizeof(double) * 3 * n_samples);
    }
    LFVIO_CUDA_TRY(cudaMemcpyAsync(A.d, A.h, in_bytes, cudaMemcpyHostToDevice, A.s));
    LFVIO_TRY(launch_preintegrate(reinterpret_cast<lfvio_imu_interval *>(A.d + o_iv), n_intervals, reinterpret_cast<double *>(A.d + o_dt),
                                  reinterpret_cast<double *>(A.d + o_acc), reinterpret_cast<double *>(A.d + o_gyr), *noise, g,
                                  reinterpret_cast<lfvio_imu_factor *>(A.d + o_out), reinterpret_cast<double *>(A.d + o_state), A.s));
    LFVIO_CUDA_TRY(cudaMemcpyAsync(A.h + o_out, A.d + o_out, pk.off - o_out, cudaMemcpyDeviceToHost, A.s));
    LFVIO_CUDA_TRY(cudaStreamSynchronize(A.s));
    memcpy(out, A.h + o_out, sizeof(lfvio_imu_factor) * n_intervals);
    if (state_out) memcpy(state_out, A.h + o_state, sizeof(double) * 15 * n_intervals);
    return LFVIO_OK;
}

extern "C" int lfvio_triangulate(const double *Ps, const double *Rs, int n_frames, const double *tic, const double *ric,
                                 const int *start_frame, const int *obs_first, const double *obs_pts, int n_features,
                                 double init_depth, double *depth, int device) {
    clear_error();
    LFVIO_REQUIRE(n_features >= 0 && n_frames > 0 && n_frames <= kMaxFrames, "lfvio_triangulate: bad arguments");
    if (n_features == 0) return LFVIO_OK;
    LFVIO_REQUIRE(Ps && Rs && tic && ric && start_frame && obs_first && obs_pts && depth, "lfvio_triangulate: null buffer");
    const int n_obs = obs_first[n_features];
    for (int l = 0; l < n_features; l++) {
        const int K = obs_first[l + 1] - obs_first[l];
        LFVIO_REQUIRE(K >= 0 && start_frame[l] >= 0 && start_frame[l] + K <= n_frames, "lfvio_triangulate: track leaves the window");
    }
    LFVIO_CUDA_TRY(cudaSetDevice(device));
    Packer pk;
    const size_t o_P = pk.take(sizeof(double) * 3 * n_frames), o_R = pk.take(sizeof(double) * 9 * n_frames);
    const size_t o_t = pk.take(sizeof(double) * 3), o_r = pk.take(sizeof(double) * 9);
    const size_t o_pts = pk.take(sizeof(double) * 3 * (size_t)n_obs), o_sf = pk.take(sizeof(int) * n_features);
    const size_t o_of = pk.take(sizeof(int) * ((size_t)n_features + 1)), o_depth = pk.take(sizeof(double) * n_features);
    StepArena &A = t_arena;
    LFVIO_TRY(A.ensure(device, pk.off));
    memcpy(A.h + o_P, Ps, sizeof(double) * 3 * n_frames);
    memcpy(A.h + o_R, Rs, sizeof(double) * 9 * n_frames);
    memcpy(A.h + o_t, tic, sizeof(double) * 3);
    memcpy(A.h + o_r, ric, sizeof(double) * 9);
    if (n_obs) memcpy(A.h + o_pts, obs_pts, sizeof(double) * 3 * (size_t)n_obs);
    memcpy(A.h + o_sf, start_frame, sizeof(int) * n_features);
    memcpy(A.h + o_of, obs_first, sizeof(int) * ((size_t)n_features + 1));
    memcpy(A.h + o_depth, depth, sizeof(double) * n_features);
    LFVIO_CUDA_TRY(cudaMemcpyAsync(A.d, A.h, pk.off, cudaMemcpyHostToDevice, A.s));
    LFVIO_TRY(launch_triangulate(reinterpret_cast<double *>(A.d + o_P), reinterpret_cast<double *>(A.d + o_R), reinterpret_cast<double *>(A.d + o_t),
                                 reinterpret_cast<double *>(A.d + o_r), reinterpret_cast<int *>(A.d + o_sf), reinterpret_cast<int *>(A.d + o_of),
                                 reinterpret_cast<double *>(A.d + o_pts), n_features, n_frames, init_depth, reinterpret_cast<double *>(A.d + o_depth), A.s));
    LFVIO_CUDA_TRY(cudaMemcpyAsync(A.h + o_depth, A.d + o_depth, sizeof(double) * n_features, cudaMemcpyDeviceToHost, A.s));
    LFVIO_CUDA_TRY(cudaStreamSynchronize(A.s));
    memcpy(depth, A.h + o_depth, sizeof(double) * n_features);
    return LFVIO_OK;
}
