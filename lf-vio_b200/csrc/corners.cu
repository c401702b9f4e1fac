// K4 — Shi-Tomasi corner detection, matching cv::goodFeaturesToTrack(img, K, 0.01, minDist, mask).
//
// Reference call site: feature_tracker/src/feature_tracker.cpp:166; in-tree specification of the
// selection stage: pointcloud_image_fusion/src/pointcloud_image_fusion.cpp:210-396.
// Algorithm: cornerMinEigenVal(blockSize 3, Sobel 3) -> max under the mask ->
// threshold(0.01*max, TOZERO) -> 3x3 dilate NMS -> descending (value, index) order -> greedy
// min-distance acceptance until maxCorners.
//
// Three launches:
//  1. min_eig_kernel: fused Sobel -> products -> 3x3 box -> lambda_min on a shared-memory tile, with the
//     masked maximum reduced per CTA into one atomicMax.  The f32 rounding sequence reproduces the
//     cv2 4.13 (AVX2/FMA) build bit for bit: dx = fma(S0+S2, s, S1*2s), dy rows = fma(C, s, fma(B, 2s, A*s)),
//     un-fused products, box sums in f64, un-fused lambda (probed in-container, see oracle/frontend.py).
//  2. candidates_kernel: threshold + 3x3 NMS + mask -> 64-bit keys (ordered value bits << 32 | index),
//     warp-aggregated append, per-cell counts and a 1024-bin value histogram.
//  3. select_kernel (one CTA): the greedy acceptance is the lexicographically-first maximal independent
//     set w.r.t. the key order, so it is computed in parallel rounds over cell lists (a candidate is
//     accepted once every higher-priority neighbour within minDistance is rejected, rejected as soon as
//     one is accepted); only the top-valued prefix needed to produce K corners is activated (histogram
//     cut), and the K best accepted keys are emitted in descending order == OpenCV's acceptance order.
// Traffic: read 1 B/px + write 4 B/px (eig), then read 4 B/px + 1 B/px mask = 10 B/px.
#include <cstring>

#include "frontend.cuh"
#include "tma.cuh"

namespace lfvio {

namespace {

__host__ __device__ __forceinline__ uint32_t enc_ordered(uint32_t bits) {
    return bits ^ ((bits >> 31) ? 0xFFFFFFFFu : 0x80000000u);
}
__device__ __forceinline__ float dec_ordered(uint32_t e) {
    uint32_t bits = e ^ ((e >> 31) ? 0x80000000u : 0xFFFFFFFFu);
    return __uint_as_float(bits);
}

// ------------------------------------------------------------------------------------------------
// Round 2: 64 x 64 outputs per CTA (256 threads).  The image tile goes to shared memory once as bytes (aligned 32-bit loads away
// from the border); a thread then owns a 4-wide, 4-high strip: it forms the derivative rows it needs from three byte rows of the
// tile, keeps the f64 horizontal 3-sums of the products of three derivative rows in registers (the box filter is separable, as in
// OpenCV's boxFilter) and rolls them down the strip — 3 conversions + 12 f64 additions per pixel instead of 27 + 27, no shared
// memory for floats, one barrier.  Rounding sequence unchanged (see the header of this file): bit-exact against cv2.
constexpr int ET_W = 64;       // tile width; the tile is 16 strips of ET_R rows high
constexpr int ET_TW = ET_W + 32;  // byte tile: virtual columns x0-16 .. x0+79 (one TMA box: 16-byte aligned start), rows y0-2 .. y0+ET_H+1
constexpr int ET_XO = 16;        // tile column of virtual column x0

template <int ET_R>
__global__ void __launch_bounds__(256) min_eig_kernel(const uint8_t *__restrict__ img, int pitch,
                                                      const uint8_t *__restrict__ mask, int mask_pitch,
                                                      float *__restrict__ eig, int rows, int cols,
                                                      uint32_t *__restrict__ max_bits, const __grid_constant__ CUtensorMap tmap, int use_tma) {
    constexpr int ET_H = 16 * ET_R, ET_TH = ET_H + 4;
    __shared__ __align__(128) uint8_t s_I[ET_TH][ET_TW];
    __shared__ uint64_t s_mbar;
    __shared__ uint32_t s_max[8];
    const int x0 = blockIdx.x * ET_W, y0 = blockIdx.y * ET_H;
    const int tid = threadIdx.x;
    // ---- byte tile at virtual coordinates: one TMA box load (cp.async.bulk.tensor.2d, completion on an mbarrier) for tiles away
    //      from the image border, REFLECT_101 gathered with plain loads for the others ----
    const bool interior = x0 >= ET_XO && x0 + ET_TW - ET_XO <= cols && y0 >= 2 && y0 + ET_H + 2 <= rows;
    if (interior && use_tma) {
        if (tid == 0) {
            mbar_init(&s_mbar, 1);
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            mbar_expect_tx(&s_mbar, ET_TH * ET_TW);
            tma_load_tile(&s_I[0][0], &tmap, x0 - ET_XO, y0 - 2, &s_mbar);
        }
        __syncthreads();  // the barrier is initialised before anybody waits on it
        mbar_wait(&s_mbar, 0);
    } else if (interior && ((pitch & 3) == 0) && ((reinterpret_cast<uintptr_t>(img) & 3) == 0)) {
        const uint8_t *base = img + (size_t)(y0 - 2) * pitch + (x0 - ET_XO);
        for (int k = tid; k < ET_TH * (ET_TW / 4); k += 256) {
            const int ty = k / (ET_TW / 4), w = k % (ET_TW / 4);
            *reinterpret_cast<uint32_t *>(&s_I[ty][4 * w]) = __ldg(reinterpret_cast<const uint32_t *>(base + (size_t)ty * pitch) + w);
        }
        __syncthreads();
    } else {
        for (int k = tid; k < ET_TH * ET_TW; k += 256) {
            const int ty = k / ET_TW, tx = k % ET_TW;
            const int y = reflect101(y0 - 2 + ty, rows), x = reflect101(x0 - ET_XO + tx, cols);
            s_I[ty][tx] = img[(size_t)y * pitch + x];
        }
        __syncthreads();
    }
    const float s = (float)(1.0 / (4 * 3 * 255.0));  // Sobel scale for u8, ksize 3, blockSize 3
    const float s2 = 2.0f * s;
    const int X = x0 + 4 * (tid & 15), Y0 = y0 + ET_R * (tid >> 4);  // first output of this strip
    uint32_t local_max = 0;                                          // enc_ordered(-NaN..) minimum
    if (X < cols && Y0 < rows) {
        // derivative row at VIRTUAL row vy, virtual columns X-1 .. X+4: a position outside the image takes the value at its
        // REFLECT_101 position (the box filter's border rule on the covariance images)
        auto deriv_row = [&](int vy, float (&gx)[6], float (&gy)[6]) {
            const int t = reflect101(vy, rows) - (y0 - 2);  // tile row of the centre
            float a[3][8];                                  // tile columns 4 tx + 14 .. 4 tx + 21 = virtual X-2 .. X+5
#pragma unroll
            for (int r = 0; r < 3; r++) {
                const uint32_t *w = reinterpret_cast<const uint32_t *>(&s_I[t - 1 + r][4 * (tid & 15) + ET_XO - 4]);
                const uint32_t w0 = w[0], w1 = w[1], w2 = w[2];
                a[r][0] = (float)((w0 >> 16) & 255u); a[r][1] = (float)(w0 >> 24);
                a[r][2] = (float)(w1 & 255u); a[r][3] = (float)((w1 >> 8) & 255u); a[r][4] = (float)((w1 >> 16) & 255u); a[r][5] = (float)(w1 >> 24);
                a[r][6] = (float)(w2 & 255u); a[r][7] = (float)((w2 >> 8) & 255u);
            }
#pragma unroll
            for (int k = 0; k < 6; k++) {
                // dx: row filter [-1 0 1] (exact), column filter [s 2s s] as fma(S0+S2, s, S1*2s)
                const float S0 = a[0][k + 2] - a[0][k], S1 = a[1][k + 2] - a[1][k], S2 = a[2][k + 2] - a[2][k];
                gx[k] = __fmaf_rn(__fadd_rn(S0, S2), s, __fmul_rn(S1, s2));
                // dy: row filter [s 2s s] as fma(C, s, fma(B, 2s, A*s)), column filter [-1 0 1]
                const float R0 = __fmaf_rn(a[0][k + 2], s, __fmaf_rn(a[0][k + 1], s2, __fmul_rn(a[0][k], s)));
                const float R2 = __fmaf_rn(a[2][k + 2], s, __fmaf_rn(a[2][k + 1], s2, __fmul_rn(a[2][k], s)));
                gy[k] = __fsub_rn(R2, R0);
            }
            // virtual columns -1 and cols: the value at column 1 / cols - 2
#pragma unroll
            for (int k = 0; k < 6; k++) {
                const int vx = X - 1 + k;
                if (vx == -1 && k + 2 < 6) {
                    gx[k] = gx[k + 2];
                    gy[k] = gy[k + 2];
                }
            }
#pragma unroll
            for (int k = 2; k < 6; k++) {
                const int vx = X - 1 + k;
                if (vx == cols) {
                    gx[k] = gx[k - 2];
                    gy[k] = gy[k - 2];
                }
            }
        };
        // f64 horizontal 3-sums of the products of one derivative row, for the four output columns
        auto hsums = [&](int vy, double (&hxx)[4], double (&hxy)[4], double (&hyy)[4]) {
            float gx[6], gy[6];
            deriv_row(vy, gx, gy);
            double pxx[6], pxy[6], pyy[6];
#pragma unroll
            for (int k = 0; k < 6; k++) {
                pxx[k] = (double)__fmul_rn(gx[k], gx[k]);
                pxy[k] = (double)__fmul_rn(gx[k], gy[k]);
                pyy[k] = (double)__fmul_rn(gy[k], gy[k]);
            }
#pragma unroll
            for (int q = 0; q < 4; q++) {
                hxx[q] = (pxx[q] + pxx[q + 1]) + pxx[q + 2];
                hxy[q] = (pxy[q] + pxy[q + 1]) + pxy[q + 2];
                hyy[q] = (pyy[q] + pyy[q + 1]) + pyy[q + 2];
            }
        };
        double hxx[3][4], hxy[3][4], hyy[3][4];  // derivative rows oy-1, oy, oy+1 (rolling)
        hsums(Y0 - 1, hxx[0], hxy[0], hyy[0]);
        hsums(Y0, hxx[1], hxy[1], hyy[1]);
#pragma unroll
        for (int rr = 0; rr < ET_R; rr++) {
            const int y = Y0 + rr;
            if (y >= rows) break;
            hsums(y + 1, hxx[(rr + 2) % 3], hxy[(rr + 2) % 3], hyy[(rr + 2) % 3]);
            const int r0 = rr % 3, r1 = (rr + 1) % 3, r2 = (rr + 2) % 3;
            float lam[4];
#pragma unroll
            for (int q = 0; q < 4; q++) {
                const double sxx = (hxx[r0][q] + hxx[r1][q]) + hxx[r2][q];
                const double sxy = (hxy[r0][q] + hxy[r1][q]) + hxy[r2][q];
                const double syy = (hyy[r0][q] + hyy[r1][q]) + hyy[r2][q];
                const float a = __fmul_rn((float)sxx, 0.5f), b = (float)sxy, c = __fmul_rn((float)syy, 0.5f);
                const float t = __fsub_rn(a, c);
                lam[q] = __fsub_rn(__fadd_rn(a, c), __fsqrt_rn(__fadd_rn(__fmul_rn(t, t), __fmul_rn(b, b))));
            }
            float *out = eig + (size_t)y * cols + X;
            const bool full = X + 4 <= cols;
            if (full && ((cols & 3) == 0) && ((reinterpret_cast<uintptr_t>(eig) & 15) == 0)) {
                *reinterpret_cast<float4 *>(out) = make_float4(lam[0], lam[1], lam[2], lam[3]);
            } else {
#pragma unroll
                for (int q = 0; q < 4; q++)
                    if (X + q < cols) out[q] = lam[q];
            }
            unsigned mbits = 0x01010101u;
            if (mask != nullptr) {
                if (full && ((mask_pitch & 3) == 0) && ((reinterpret_cast<uintptr_t>(mask) & 3) == 0)) {
                    mbits = __ldg(reinterpret_cast<const unsigned *>(mask + (size_t)y * mask_pitch + X));
                } else {
                    mbits = 0;
#pragma unroll
                    for (int q = 0; q < 4; q++)
                        if (X + q < cols && mask[(size_t)y * mask_pitch + X + q] != 0) mbits |= 1u << (8 * q);
                }
            }
#pragma unroll
            for (int q = 0; q < 4; q++)
                if (X + q < cols && ((mbits >> (8 * q)) & 255u) != 0) local_max = max(local_max, enc_ordered(__float_as_uint(lam[q])));
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) local_max = max(local_max, __shfl_xor_sync(0xffffffffu, local_max, o));
    if ((tid & 31) == 0) s_max[tid >> 5] = local_max;
    __syncthreads();
    if (tid == 0) {
        uint32_t m = s_max[0];
#pragma unroll
        for (int k = 1; k < 8; k++) m = max(m, s_max[k]);
        if (m) atomicMax(max_bits, m);
    }
}

// ------------------------------------------------------------------------------------------------
struct CandParams {
    int rows, cols, mask_pitch;
    double quality;
    int cap;
    int gw, cell;
};

__device__ __forceinline__ int cand_hist_shift(uint32_t range) {
    const int bits = 32 - __clz(range | 1u);
    return max(0, bits - 10);
}

__global__ void __launch_bounds__(256) candidates_kernel(const float *__restrict__ eig, const uint8_t *__restrict__ mask,
                                                         const uint32_t *__restrict__ max_bits, CandParams p,
                                                         unsigned long long *__restrict__ cand,
                                                         int *__restrict__ cell_count, uint32_t *__restrict__ hist,
                                                         int *__restrict__ counters) {
    const uint32_t menc = *max_bits;
    if (menc == 0) return;  // empty mask: no maximum, no corners
    const float max_val = dec_ordered(menc);
    const float thr = (float)((double)max_val * p.quality);  // cv::threshold takes a double, compares in f32
    const uint32_t tenc = enc_ordered(__float_as_uint(thr));
    const int shift = cand_hist_shift(menc - tenc);
    if (blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0) {
        counters[3] = (int)tenc;
        counters[4] = shift;
    }
    // Grid-stride over chunks of four consecutive pixels of one row (one 128-bit load of the eigenvalue map and one 32-bit load
    // of the mask per chunk when the rows are 16-byte aligned); a pixel that fails the threshold costs nothing more, the eight
    // neighbours are read only for the few that pass.  One atomicAdd per warp and pass reserves the output slots.
    const int lane = threadIdx.x & 31;
    const int chunks_per_row = (p.cols + 3) >> 2;
    const bool vec_ok = ((p.cols & 3) == 0) && ((reinterpret_cast<uintptr_t>(eig) & 15) == 0);
    const bool mvec_ok = mask != nullptr && ((p.mask_pitch & 3) == 0) && ((reinterpret_cast<uintptr_t>(mask) & 3) == 0);
    // rows over blockIdx.y, chunks of a row over blockIdx.x (no division per chunk; whole warps stay in the loops together
    // because of the ballots below: the chunk loop runs to the warp's common bound)
    const int chunks_pad = (chunks_per_row + 31) & ~31;
    for (int y = 1 + blockIdx.y; y < p.rows - 1; y += gridDim.y)
    for (int ch = blockIdx.x * blockDim.x + threadIdx.x; ch < chunks_pad; ch += gridDim.x * blockDim.x) {
        const bool in_row = ch < chunks_per_row;
        float v[4] = {0.f, 0.f, 0.f, 0.f};
        unsigned pass = 0;
        const int x0 = 4 * ch;
        if (in_row) {
            const float *row = eig + (size_t)y * p.cols;
            if (vec_ok) {
                const float4 q = __ldg(reinterpret_cast<const float4 *>(row + x0));
                v[0] = q.x; v[1] = q.y; v[2] = q.z; v[3] = q.w;
            } else {
#pragma unroll
                for (int k = 0; k < 4; k++)
                    if (x0 + k < p.cols) v[k] = row[x0 + k];
            }
            unsigned mbits = 0xffffffffu;
            if (mvec_ok && x0 + 4 <= p.cols) mbits = __ldg(reinterpret_cast<const unsigned *>(mask + (size_t)y * p.mask_pitch + x0));
#pragma unroll
            for (int k = 0; k < 4; k++) {
                const int x = x0 + k;
                if (x < 1 || x >= p.cols - 1) continue;
                if (!(v[k] > thr) || v[k] == 0.f) continue;
                bool m_ok = true;
                if (mask != nullptr) m_ok = (mvec_ok && x0 + 4 <= p.cols) ? (((mbits >> (8 * k)) & 255u) != 0) : (mask[(size_t)y * p.mask_pitch + x] != 0);
                if (!m_ok) continue;
                const float *cc = row + x;
                bool is_max = true;
#pragma unroll
                for (int dy = -1; dy <= 1; dy++)
#pragma unroll
                    for (int dx = -1; dx <= 1; dx++) {
                        const float nb = cc[dy * p.cols + dx];
                        const float tn = nb > thr ? nb : 0.f;
                        if (tn > v[k]) is_max = false;
                    }
                if (is_max) pass |= 1u << k;
            }
        }
        if (!__any_sync(0xffffffffu, pass != 0)) continue;
        // warp-aggregated append: exclusive scan of the per-lane counts
        const int cnt = __popc(pass);
        int inc = cnt;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += t;
        }
        const int warp_total = __shfl_sync(0xffffffffu, inc, 31);
        int base = 0;
        if (lane == 0) base = atomicAdd(&counters[0], warp_total);
        base = __shfl_sync(0xffffffffu, base, 0);
        int idx = base + inc - cnt;
#pragma unroll
        for (int k = 0; k < 4; k++) {
            if (!(pass & (1u << k))) continue;
            const int x = x0 + k;
            const uint32_t venc = enc_ordered(__float_as_uint(v[k]));
            if (idx < p.cap) {
                cand[idx] = ((unsigned long long)venc << 32) | (uint32_t)(y * p.cols + x);
                atomicAdd(&cell_count[(y / p.cell) * p.gw + (x / p.cell)], 1);
                atomicAdd(&hist[min((venc - tenc) >> shift, (uint32_t)(kCandHistBins - 1))], 1u);
            } else {
                counters[1] = 1;  // overflow: reported as LFVIO_ERR_CAPACITY by the caller
            }
            idx++;
        }
    }
}

// ------------------------------------------------------------------------------------------------
constexpr int SEL_THREADS = 1024;

struct SelParams {
    int cols, gw, gh, cell, cap;
    double md2;
    int max_total, out_cap;
};

__device__ __forceinline__ int block_exclusive_scan_1024(int v, int *s_warp, int *total) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    int inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        int t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
    }
    __syncthreads();
    if (lane == 31) s_warp[w] = inc;
    __syncthreads();
    if (w == 0) {
        int t = s_warp[lane];
        int ti = t;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            int u = __shfl_up_sync(0xffffffffu, ti, o);
            if (lane >= o) ti += u;
        }
        s_warp[lane] = ti - t;
        if (lane == 31) s_warp[32] = ti;
    }
    __syncthreads();
    *total = s_warp[32];
    return inc - v + s_warp[w];
}

__global__ void __launch_bounds__(SEL_THREADS) select_kernel(SelParams p, unsigned long long *cand, unsigned long long *sorted,
                                                             int *cell_count, int *cell_start,
                                                             const uint32_t *__restrict__ hist, volatile uint8_t *state,
                                                             int *counters, int *n_inout, float2 *out_pts, int *out_ids,
                                                             int *out_cnt, uint8_t *out_prev_valid) {
    __shared__ int s_warp[33];
    __shared__ int s_misc[4];
    const int tid = threadIdx.x;
    const int M = min(counters[0], p.cap);
    const int n_cur = n_inout ? *n_inout : 0;
    const int K = min(p.max_total - n_cur, p.out_cap - n_cur);
    if (K <= 0 || M == 0) {
        if (tid == 0) counters[2] = 0;
        return;
    }
    // ---- cell_start = exclusive scan of cell_count; cell_count becomes the fill cursor ----
    const int ncell = p.gw * p.gh;
    int carry = 0;
    for (int c0 = 0; c0 < ncell; c0 += SEL_THREADS) {
        const int c = c0 + tid;
        const int v = c < ncell ? cell_count[c] : 0;
        int total;
        const int ex = block_exclusive_scan_1024(v, s_warp, &total);
        if (c < ncell) {
            cell_start[c] = carry + ex;
            cell_count[c] = 0;
        }
        carry += total;
        __syncthreads();
    }
    if (tid == 0) cell_start[ncell] = carry;
    __syncthreads();
    // ---- scatter candidates into cell order ----
    for (int i = tid; i < M; i += SEL_THREADS) {
        const unsigned long long key = cand[i];
        const int lin = (int)(key & 0xffffffffu);
        const int y = lin / p.cols, x = lin - y * p.cols;
        const int c = (y / p.cell) * p.gw + (x / p.cell);
        const int pos = cell_start[c] + atomicAdd(&cell_count[c], 1);
        sorted[pos] = key;
        state[pos] = 0;
    }
    __syncthreads();
    // ---- histogram cut: activate only the top-valued prefix (bins = (enc(v) - enc(thr)) >> shift) ----
    int want = min(M, 4 * K + 256);
    int accepted_total = 0;
    int bin_thr = kCandHistBins;  // nothing active yet
    const uint32_t tenc = (uint32_t)counters[3];  // written by candidates_kernel: enc(threshold), histogram shift
    const int shift = counters[4];

    for (;;) {
        // lower bin_thr until `want` candidates are active
        if (tid == 0) {
            int cum = 0, b = kCandHistBins;
            while (b > 0 && cum < want) {
                b--;
                cum += (int)hist[b];
            }
            s_misc[0] = b;
            s_misc[1] = cum;
        }
        __syncthreads();
        bin_thr = s_misc[0];
        const int n_active = s_misc[1];
        __syncthreads();
        // ---- MIS rounds over active, undecided candidates ----
        for (;;) {
            int undecided = 0;
            for (int i = tid; i < M; i += SEL_THREADS) {
                if (state[i] != 0) continue;
                const unsigned long long key = sorted[i];
                const uint32_t venc = (uint32_t)(key >> 32);
                const int bin = (int)min((venc - tenc) >> shift, (uint32_t)(kCandHistBins - 1));
                if (bin < bin_thr) continue;
                const int lin = (int)(key & 0xffffffffu);
                const int y = lin / p.cols, x = lin - y * p.cols;
                const int cx = x / p.cell, cy = y / p.cell;
                bool rejected = false, blocked = false;
                for (int yy = max(cy - 1, 0); yy <= min(cy + 1, p.gh - 1) && !rejected; yy++) {
                    const int c_lo = yy * p.gw + max(cx - 1, 0), c_hi = yy * p.gw + min(cx + 1, p.gw - 1);
                    const int j0 = cell_start[c_lo], j1 = cell_start[c_hi + 1];  // cells of one row are contiguous
                    for (int j = j0; j < j1; j++) {
                        const unsigned long long kj = sorted[j];
                        if (kj <= key) continue;  // only higher priority matters (j == i excluded too)
                        const int lj = (int)(kj & 0xffffffffu);
                        const int yj = lj / p.cols, xj = lj - yj * p.cols;
                        const int dx = x - xj, dy = y - yj;
                        if ((double)(dx * dx + dy * dy) < p.md2) {
                            const uint8_t sj = state[j];
                            if (sj == 1) {
                                rejected = true;
                                break;
                            }
                            if (sj == 0) blocked = true;
                        }
                    }
                }
                if (rejected) state[i] = 2;
                else if (!blocked) state[i] = 1;
                else undecided = 1;
            }
            if (__syncthreads_or(undecided) == 0) break;
        }
        // ---- count accepted ----
        int acc = 0;
        for (int i = tid; i < M; i += SEL_THREADS) acc += (state[i] == 1);
        int total;
        block_exclusive_scan_1024(acc, s_warp, &total);
        accepted_total = total;
        __syncthreads();
        if (accepted_total >= K || n_active >= M || bin_thr == 0) break;
        want = min(M, 2 * want);
    }
    // ---- gather accepted keys (into `cand`, no longer needed), rank them, emit the best K ----
    if (tid == 0) s_misc[2] = 0;
    __syncthreads();
    for (int i = tid; i < M; i += SEL_THREADS)
        if (state[i] == 1) cand[atomicAdd(&s_misc[2], 1)] = sorted[i];
    __syncthreads();
    const int A = s_misc[2];
    for (int i = tid; i < A; i += SEL_THREADS) {
        const unsigned long long key = cand[i];
        int rank = 0;
        for (int j = 0; j < A; j++) rank += (cand[j] > key);
        if (rank < K) {
            const int lin = (int)(key & 0xffffffffu);
            const int y = lin / p.cols, x = lin - y * p.cols;
            const int o = n_cur + rank;
            out_pts[o] = make_float2((float)x, (float)y);
            if (out_ids) out_ids[o] = -1;            // addPoints(): ids.push_back(-1)
            if (out_cnt) out_cnt[o] = 1;             //              track_cnt.push_back(1)
            if (out_prev_valid) out_prev_valid[o] = 0;
        }
    }
    __syncthreads();
    if (tid == 0) {
        const int n_sel = min(A, K);
        counters[2] = n_sel;
        if (n_inout) *n_inout = n_cur + n_sel;
    }
}

}  // namespace

int launch_min_eig(const uint8_t *img, int pitch, const uint8_t *mask, int mask_pitch, float *eig, int rows,
                   int cols, uint32_t *max_bits, cudaStream_t s, const CUtensorMap *tile_map) {
    LFVIO_CUDA_TRY(cudaMemsetAsync(max_bits, 0, sizeof(uint32_t), s));
    // (strips of 8 rows — 1.25 derivative rows per output row instead of 1.5 — measured slower at 3840 x 2880: 120 against 112 us)
    // tile_map: tensor map of `img` with a 96 x 68 box (encode_u8_tmap); null -> the tiles are staged with plain loads
    CUtensorMap tm;
    if (tile_map) tm = *tile_map;
    else memset(&tm, 0, sizeof(tm));
    min_eig_kernel<4><<<dim3(cdiv(cols, ET_W), cdiv(rows, 64)), 256, 0, s>>>(img, pitch, mask, mask_pitch, eig, rows, cols, max_bits, tm,
                                                                              tile_map ? 1 : 0);
    LFVIO_LAUNCH_CHECK();
    return LFVIO_OK;
}

int launch_corner_candidates(const float *eig, const uint8_t *mask, int mask_pitch, int rows, int cols,
                             const uint32_t *max_bits, double quality, const CornerScratch &cs, cudaStream_t s) {
    LFVIO_CUDA_TRY(cudaMemsetAsync(cs.counters, 0, sizeof(int) * 8, s));
    LFVIO_CUDA_TRY(cudaMemsetAsync(cs.cell_count, 0, sizeof(int) * (cs.gw * cs.gh + 1), s));
    LFVIO_CUDA_TRY(cudaMemsetAsync(cs.hist, 0, sizeof(uint32_t) * kCandHistBins, s));
    if (rows < 3 || cols < 3) return LFVIO_OK;
    CandParams p;
    p.rows = rows;
    p.cols = cols;
    p.mask_pitch = mask_pitch;
    p.quality = quality;
    p.cap = cs.cand_cap;
    p.gw = cs.gw;
    p.cell = cs.cell;
    const int gx = std::max(1, cdiv((cols + 3) / 4, 256));
    const int gy = std::max(1, std::min(rows - 2, cdiv(148 * 8, gx)));
    candidates_kernel<<<dim3(gx, gy), 256, 0, s>>>(eig, mask, max_bits, p, cs.cand, cs.cell_count, cs.hist, cs.counters);
    LFVIO_LAUNCH_CHECK();
    return LFVIO_OK;
}

int launch_corner_select(const CornerScratch &cs, int cols, float min_distance, int max_total, int *n_inout,
                         float2 *out_pts, int *out_ids, int *out_cnt, uint8_t *out_prev_valid, int out_cap,
                         cudaStream_t s) {
    SelParams p;
    p.cols = cols;
    p.gw = cs.gw;
    p.gh = cs.gh;
    p.cell = cs.cell;
    p.cap = cs.cand_cap;
    p.md2 = (double)min_distance * (double)min_distance;  // minDistance *= minDistance (double)
    p.max_total = max_total;
    p.out_cap = out_cap;
    select_kernel<<<1, SEL_THREADS, 0, s>>>(p, cs.cand, cs.cand_sorted, cs.cell_count, cs.cell_start, cs.hist, cs.state,
                                            cs.counters, n_inout, out_pts, out_ids, out_cnt, out_prev_valid);
    LFVIO_LAUNCH_CHECK();
    return LFVIO_OK;
}

}  // namespace lfvio
