// K3 — pyramidal Lucas-Kanade, 41x41 window, following cv::calcOpticalFlowPyrLK's fixed-point scheme.
//
// Reference call site: feature_tracker/src/feature_tracker.cpp:127
//   cv::calcOpticalFlowPyrLK(cur_img, forw_img, cur_pts, forw_pts, status, err, cv::Size(41, 41), 3);
// with the inBorder() filter of :129-131 (and :6-12) fused into `status`.
// Algorithm: OpenCV's LKTrackerInvoker (not in the reference tree; restated in SURVEY.md Appendix B.1):
//   W_BITS = 14 fixed-point bilinear weights, template = value*32 (s16), s16 Scharr gradients with
//   REFLECT_101 inside the image and ZERO outside, A/b scaled by 2^-20, <=30 iterations, eps^2 = 1e-4,
//   oscillation back-off, level-0 bounds => status 0 (incl. the extra bounds check of the `err` pass).
//
// One CTA (4 warps) per feature, all pyramid levels inside the CTA (coarse -> fine is a serial chain).
// Per level the 44x44 u8 tile of the previous image is staged in shared memory (TMA box load for interior tiles), the Scharr gradients and
// the 41x41 template are built there once, and every iteration samples the next image (L1/L2-resident)
// and reduces the 2x2 normal-equation right-hand side with warp shuffles.  The sums are accumulated in
// exact integer arithmetic (int32 per thread, int64 across the CTA) — order independent, deterministic —
// and converted to f32 once; OpenCV accumulates in f32 (SIMD order), so positions agree to ~1e-4 px, not
// bitwise.  The scalar f32 update math is written with un-fused intrinsics to follow OpenCV's rounding.
// Algorithmic HBM bytes: 2 * 43^2 B per feature per level.
#include "frontend.cuh"
#include "tma.cuh"

namespace lfvio {

namespace {

constexpr int WIN = 41;
constexpr int HALF = 20;
constexpr int TILE = WIN + 3;     // 44: I at ip-1 .. ip+42
constexpr int TILE_P = 64;        // row pitch == TMA box width: 44 columns + up to 15 of 16-byte alignment slack
constexpr int DW = WIN + 1;       // 42: derivative positions ip .. ip+41
constexpr int LK_THREADS = 128;
constexpr int SEGS = 3, SEG_LEN = 14;  // 41 rows x 3 segments = 123 working threads
constexpr int W_BITS = 14;
constexpr int JM = 8;                        // margin of the staged search tile of the next image around the first guess
constexpr int JT_H = WIN + 1 + 2 * JM;       // 58 rows
constexpr int JT_W = 80;                     // bytes per row: 42 + 2 * 8 columns + up to 15 of 16-byte alignment slack

__device__ __forceinline__ int descale(int x, int n) { return (x + (1 << (n - 1))) >> n; }

struct Weights {
    int w00, w01, w10, w11;
};
__device__ __forceinline__ Weights make_weights(float a, float b) {
    Weights w;
    const float one_a = __fsub_rn(1.f, a), one_b = __fsub_rn(1.f, b);
    const float sc = (float)(1 << W_BITS);
    w.w00 = cv_round(__fmul_rn(__fmul_rn(one_a, one_b), sc));
    w.w01 = cv_round(__fmul_rn(__fmul_rn(a, one_b), sc));
    w.w10 = cv_round(__fmul_rn(__fmul_rn(one_a, b), sc));
    w.w11 = (1 << W_BITS) - w.w00 - w.w01 - w.w10;
    return w;
}

// CTA-wide sum of K long long values; result broadcast to all threads. `buf` is [4][K] shared, caller
// alternates two buffers so one __syncthreads per reduction suffices.
template <int K>
__device__ __forceinline__ void block_sum(long long (&v)[K], long long *buf) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < K; k++) v[k] = warp_sum(v[k]);
    if (lane == 0) {
#pragma unroll
        for (int k = 0; k < K; k++) buf[w * K + k] = v[k];
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < K; k++) v[k] = buf[0 * K + k] + buf[1 * K + k] + buf[2 * K + k] + buf[3 * K + k];
}

__global__ void __launch_bounds__(LK_THREADS) lk_kernel(Pyramid prev, const __grid_constant__ PyrTmaps prev_tm, Pyramid next,
                                                        const float2 *__restrict__ prev_pts,
                                                        float2 *__restrict__ next_pts, uint8_t *__restrict__ status,
                                                        int n_max, const int *__restrict__ n_dev, LkParams prm) {
    const int p = blockIdx.x;
    const int n = n_dev ? min(*n_dev, n_max) : n_max;
    if (p >= n) return;

    __shared__ __align__(128) uint8_t s_I[TILE][TILE_P];  // 44 rows x 64 bytes == one TMA box
    __shared__ uint64_t s_mbar;
    __shared__ short s_dx[DW][DW + 2], s_dy[DW][DW + 2];
    __shared__ short s_Iw[WIN][WIN + 1], s_Ix[WIN][WIN + 1], s_Iy[WIN][WIN + 1];
    __shared__ long long s_red[2][4 * 3];
    __shared__ __align__(16) uint8_t s_J[JT_H][JT_W];  // search tile of the next image (round 2): the iterations read shared memory

    const int tid = threadIdx.x;
    const bool worker = tid < WIN * SEGS;
    const int wrow = tid / SEGS;
    const int wx0 = (tid % SEGS) * SEG_LEN;
    const int wx1 = min(WIN, wx0 + SEG_LEN);

    if (threadIdx.x == 0) mbar_init(&s_mbar, 1);
    uint32_t tma_phase = 0;
    __syncthreads();
    const float2 pt0 = prev_pts[p];
    const float FLT_SCALE = 1.f / (1 << 20);
    const float halfWin = (float)HALF;
    float2 nextFull = make_float2(0.f, 0.f);
    bool st = true;
    int red_par = 0;
    const int levels = min(prev.levels, next.levels);

    for (int level = levels - 1; level >= 0; level--) {
        const ImgLevel I = prev.lvl[level], J = next.lvl[level];
        const float scale = (float)(1.0 / (double)(1 << level));
        float2 prevPt = make_float2(__fmul_rn(pt0.x, scale), __fmul_rn(pt0.y, scale));
        if (level == levels - 1) nextFull = prevPt;
        else nextFull = make_float2(__fmul_rn(nextFull.x, 2.f), __fmul_rn(nextFull.y, 2.f));

        prevPt.x = __fsub_rn(prevPt.x, halfWin);
        prevPt.y = __fsub_rn(prevPt.y, halfWin);
        const int ipx = cv_floor(prevPt.x), ipy = cv_floor(prevPt.y);
        if (ipx < -WIN || ipx >= I.cols || ipy < -WIN || ipy >= I.rows) {
            if (level == 0) st = false;
            continue;
        }
        __syncthreads();  // previous level's shared tiles are dead
        // ---- stage the 44x44 tile of I (virtual coords ip-1 ..) ----
        // interior tiles: one TMA box load (cp.async.bulk.tensor.2d) issued by one thread, completion on an mbarrier;
        // tiles touching the image border need REFLECT_101 and are gathered with plain loads.
        const bool tile_inside = (ipx - 1 >= 0) && (ipy - 1 >= 0) && (ipx - 1 + TILE <= I.cols) && (ipy - 1 + TILE <= I.rows);
        int xo = 0;  // column of the tile's first pixel inside s_I rows
        if (prev_tm.valid && level < kTmaLevels && tile_inside) {
            // a TMA box row must start on a 16-byte boundary of the image row (u8: x % 16 == 0, measured: any other
            // start faults), so the box starts at the aligned column below ip-1 and the tile sits xo bytes into it
            xo = (ipx - 1) & 15;
            if (tid == 0) {
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy reads of s_I are done
                mbar_expect_tx(&s_mbar, TILE * TILE_P);
                tma_load_tile(&s_I[0][0], &prev_tm.m[level], (ipx - 1) - xo, ipy - 1, &s_mbar);
            }
            mbar_wait(&s_mbar, tma_phase);
            tma_phase ^= 1;
        } else {
            for (int k = tid; k < TILE * TILE; k += LK_THREADS) {
                const int ty = k / TILE, tx = k % TILE;
                const int y = reflect101(ipy - 1 + ty, I.rows), x = reflect101(ipx - 1 + tx, I.cols);
                s_I[ty][tx] = I.data[(size_t)y * I.pitch + x];
            }
            __syncthreads();
        }
        // ---- stage the search tile of J around this level's first guess: one L2 round trip per level instead of one per
        //      iteration (the per-pixel byte loads of the window were the iteration's latency).  Only when the tile lies inside the
        //      image and the rows are 16-byte aligned; an iteration whose window leaves the tile reads global memory as before. ----
        int jtx0 = 0, jty0 = 0;
        bool jt_ok = false;
        {
            const int gx = cv_floor(__fsub_rn(nextFull.x, halfWin)), gy = cv_floor(__fsub_rn(nextFull.y, halfWin));
            jtx0 = (gx - JM) & ~15;
            jty0 = gy - JM;
            jt_ok = gx - JM >= 0 && jty0 >= 0 && jty0 + JT_H <= J.rows && jtx0 + JT_W <= J.cols && ((J.pitch & 15) == 0) &&
                    ((reinterpret_cast<uintptr_t>(J.data) & 15) == 0);
            if (jt_ok) {
                for (int k = tid; k < JT_H * (JT_W / 16); k += LK_THREADS) {
                    const int ty = k / (JT_W / 16), q = k % (JT_W / 16);
                    *reinterpret_cast<uint4 *>(&s_J[ty][16 * q]) =
                        __ldg(reinterpret_cast<const uint4 *>(J.data + (size_t)(jty0 + ty) * J.pitch + jtx0) + q);
                }
            }
        }
        // ---- Scharr gradients at ip .. ip+41 (zero outside the image) ----
        for (int k = tid; k < DW * DW; k += LK_THREADS) {
            const int dy = k / DW, dx = k % DW;
            const int y = ipy + dy, x = ipx + dx;
            short gx = 0, gy = 0;
            if (x >= 0 && x < I.cols && y >= 0 && y < I.rows) {
                // tile index of (y, x) is (dy+1, dx+1)
                const uint8_t *r0 = &s_I[dy][dx + xo], *r1 = r0 + TILE_P, *r2 = r1 + TILE_P;
                const int a00 = r0[0], a01 = r0[1], a02 = r0[2];
                const int a10 = r1[0], a12 = r1[2];
                const int a20 = r2[0], a21 = r2[1], a22 = r2[2];
                // Ix = smooth_y[3 10 3](x+1) - smooth_y(x-1);  Iy = smooth_x[3 10 3] of (row+1 - row-1)
                gx = (short)(((a02 + a22) * 3 + a12 * 10) - ((a00 + a20) * 3 + a10 * 10));
                gy = (short)(((a20 - a00) + (a22 - a02)) * 3 + (a21 - a01) * 10);
            }
            s_dx[dy][dx] = gx;
            s_dy[dy][dx] = gy;
        }
        __syncthreads();
        // ---- template window + spatial gradient matrix ----
        const Weights w = make_weights(__fsub_rn(prevPt.x, (float)ipx), __fsub_rn(prevPt.y, (float)ipy));
        long long acc[3] = {0, 0, 0};
        if (worker) {
            int a11 = 0, a12 = 0, a22 = 0;
            const int y = wrow;
            for (int x = wx0; x < wx1; x++) {
                const uint8_t *q0 = &s_I[y + 1][x + 1 + xo], *q1 = q0 + TILE_P;
                const int iv = descale(q0[0] * w.w00 + q0[1] * w.w01 + q1[0] * w.w10 + q1[1] * w.w11, W_BITS - 5);
                const int ix = descale(s_dx[y][x] * w.w00 + s_dx[y][x + 1] * w.w01 + s_dx[y + 1][x] * w.w10 +
                                           s_dx[y + 1][x + 1] * w.w11, W_BITS);
                const int iy = descale(s_dy[y][x] * w.w00 + s_dy[y][x + 1] * w.w01 + s_dy[y + 1][x] * w.w10 +
                                           s_dy[y + 1][x + 1] * w.w11, W_BITS);
                s_Iw[y][x] = (short)iv;
                s_Ix[y][x] = (short)ix;
                s_Iy[y][x] = (short)iy;
                a11 += ix * ix;
                a12 += ix * iy;
                a22 += iy * iy;
            }
            acc[0] = a11;
            acc[1] = a12;
            acc[2] = a22;
        }
        block_sum<3>(acc, s_red[red_par]);
        red_par ^= 1;
        const float A11 = __fmul_rn((float)acc[0], FLT_SCALE);
        const float A12 = __fmul_rn((float)acc[1], FLT_SCALE);
        const float A22 = __fmul_rn((float)acc[2], FLT_SCALE);
        float D = __fsub_rn(__fmul_rn(A11, A22), __fmul_rn(A12, A12));
        const float dif = __fsub_rn(A11, A22);
        const float minEig = __fdiv_rn(
            __fsub_rn(__fadd_rn(A22, A11),
                      __fsqrt_rn(__fadd_rn(__fmul_rn(dif, dif), __fmul_rn(__fmul_rn(4.f, A12), A12)))),
            (float)(2 * WIN * WIN));
        if ((double)minEig < prm.min_eig_thr || D < 1.1920929e-07f /* FLT_EPSILON */) {
            if (level == 0) st = false;
            continue;
        }
        D = __fdiv_rn(1.f, D);
        float2 nextPt = make_float2(__fsub_rn(nextFull.x, halfWin), __fsub_rn(nextFull.y, halfWin));
        float2 prevDelta = make_float2(0.f, 0.f);

        for (int j = 0; j < prm.max_iter; j++) {
            const int inx = cv_floor(nextPt.x), iny = cv_floor(nextPt.y);
            if (inx < -WIN || inx >= J.cols || iny < -WIN || iny >= J.rows) {
                if (level == 0) st = false;
                break;
            }
            const Weights wj = make_weights(__fsub_rn(nextPt.x, (float)inx), __fsub_rn(nextPt.y, (float)iny));
            long long bb[2] = {0, 0};
            if (worker) {
                const int y = iny + wrow;
                int b1 = 0, b2 = 0;
                const bool interior = (inx >= 0) && (inx + WIN < J.cols) && (iny >= 0) && (iny + WIN < J.rows);
                const bool in_tile = jt_ok && interior && inx >= jtx0 && inx + WIN + 1 <= jtx0 + JT_W && iny >= jty0 && iny + WIN + 1 <= jty0 + JT_H;
                if (in_tile) {
                    const uint8_t *r0 = &s_J[y - jty0][inx - jtx0 + wx0];
                    const uint8_t *r1 = r0 + JT_W;
                    int p00 = r0[0], p10 = r1[0];
                    for (int x = wx0; x < wx1; x++) {
                        const int p01 = r0[(x - wx0) + 1], p11 = r1[(x - wx0) + 1];
                        const int diff = descale(p00 * wj.w00 + p01 * wj.w01 + p10 * wj.w10 + p11 * wj.w11, W_BITS - 5) -
                                         s_Iw[wrow][x];
                        b1 += diff * s_Ix[wrow][x];
                        b2 += diff * s_Iy[wrow][x];
                        p00 = p01;
                        p10 = p11;
                    }
                } else if (interior) {
                    const uint8_t *r0 = J.data + (size_t)y * J.pitch + inx + wx0;
                    const uint8_t *r1 = r0 + J.pitch;
                    int p00 = __ldg(r0), p10 = __ldg(r1);
                    for (int x = wx0; x < wx1; x++) {
                        const int p01 = __ldg(r0 + (x - wx0) + 1), p11 = __ldg(r1 + (x - wx0) + 1);
                        const int diff = descale(p00 * wj.w00 + p01 * wj.w01 + p10 * wj.w10 + p11 * wj.w11, W_BITS - 5) -
                                         s_Iw[wrow][x];
                        b1 += diff * s_Ix[wrow][x];
                        b2 += diff * s_Iy[wrow][x];
                        p00 = p01;
                        p10 = p11;
                    }
                } else {
                    const uint8_t *r0 = J.data + (size_t)reflect101(y, J.rows) * J.pitch;
                    const uint8_t *r1 = J.data + (size_t)reflect101(y + 1, J.rows) * J.pitch;
                    int xr = reflect101(inx + wx0, J.cols);
                    int p00 = __ldg(r0 + xr), p10 = __ldg(r1 + xr);
                    for (int x = wx0; x < wx1; x++) {
                        xr = reflect101(inx + x + 1, J.cols);
                        const int p01 = __ldg(r0 + xr), p11 = __ldg(r1 + xr);
                        const int diff = descale(p00 * wj.w00 + p01 * wj.w01 + p10 * wj.w10 + p11 * wj.w11, W_BITS - 5) -
                                         s_Iw[wrow][x];
                        b1 += diff * s_Ix[wrow][x];
                        b2 += diff * s_Iy[wrow][x];
                        p00 = p01;
                        p10 = p11;
                    }
                }
                bb[0] = b1;
                bb[1] = b2;
            }
            block_sum<2>(bb, s_red[red_par]);
            red_par ^= 1;
            const float b1f = __fmul_rn((float)bb[0], FLT_SCALE);
            const float b2f = __fmul_rn((float)bb[1], FLT_SCALE);
            const float2 delta = make_float2(__fmul_rn(__fsub_rn(__fmul_rn(A12, b2f), __fmul_rn(A22, b1f)), D),
                                             __fmul_rn(__fsub_rn(__fmul_rn(A12, b1f), __fmul_rn(A11, b2f)), D));
            nextPt.x = __fadd_rn(nextPt.x, delta.x);
            nextPt.y = __fadd_rn(nextPt.y, delta.y);
            nextFull = make_float2(__fadd_rn(nextPt.x, halfWin), __fadd_rn(nextPt.y, halfWin));
            if (__dadd_rn(__dmul_rn((double)delta.x, (double)delta.x), __dmul_rn((double)delta.y, (double)delta.y)) <= prm.eps2)
                break;
            if (j > 0 && (double)fabsf(__fadd_rn(delta.x, prevDelta.x)) < 0.01 &&
                (double)fabsf(__fadd_rn(delta.y, prevDelta.y)) < 0.01) {
                nextFull.x = __fsub_rn(nextFull.x, __fmul_rn(delta.x, 0.5f));
                nextFull.y = __fsub_rn(nextFull.y, __fmul_rn(delta.y, 0.5f));
                break;
            }
            prevDelta = delta;
        }
        if (st && level == 0) {
            // the caller asks for `err`: OpenCV re-checks the final window bounds and may clear status
            const float qx = __fsub_rn(nextFull.x, halfWin), qy = __fsub_rn(nextFull.y, halfWin);
            const int iqx = cv_floor(qx), iqy = cv_floor(qy);
            if (iqx < -WIN || iqx >= J.cols || iqy < -WIN || iqy >= J.rows) st = false;
        }
    }
    if (tid == 0) {
        next_pts[p] = nextFull;
        if (st && prm.border_check) {
            const int x = cv_round(nextFull.x), y = cv_round(nextFull.y);
            st = (1 <= x && x < prm.cols - 1 && 1 <= y && y < prm.rows - 1);
        }
        status[p] = st ? 1 : 0;
    }
}

}  // namespace

int launch_lk(const Pyramid &prev, const PyrTmaps &prev_tm, const Pyramid &next, const float2 *prev_pts, float2 *next_pts,
              uint8_t *status, int n_max, const int *n_dev, const LkParams &prm, cudaStream_t s) {
    if (n_max <= 0) return LFVIO_OK;
    LFVIO_REQUIRE(prm.win == WIN, "launch_lk: only the reference's 41x41 window is built");
    lk_kernel<<<n_max, LK_THREADS, 0, s>>>(prev, prev_tm, next, prev_pts, next_pts, status, n_max, n_dev, prm);
    LFVIO_LAUNCH_CHECK();
    return LFVIO_OK;
}

}  // namespace lfvio
