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
#include "frontend.cuh"

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
constexpr int ET_W = 32, ET_H = 16;  // outputs per CTA; 256 threads, 2 rows each

__global__ void __launch_bounds__(256) min_eig_kernel(const uint8_t *__restrict__ img, int pitch,
                                                      const uint8_t *__restrict__ mask, int mask_pitch,
                                                      float *__restrict__ eig, int rows, int cols,
                                                      uint32_t *__restrict__ max_bits) {
    __shared__ float s_I[ET_H + 4][ET_W + 4 + 1];
    __shared__ float s_dx[ET_H + 2][ET_W + 2 + 1], s_dy[ET_H + 2][ET_W + 2 + 1];
    __shared__ uint32_t s_max[8];
    const int x0 = blockIdx.x * ET_W, y0 = blockIdx.y * ET_H;
    const int tid = threadIdx.x;
    // I tile at virtual coords [y0-2, y0+ET_H+2) x [x0-2, x0+ET_W+2), REFLECT_101
    for (int k = tid; k < (ET_H + 4) * (ET_W + 4); k += 256) {
        const int ty = k / (ET_W + 4), tx = k % (ET_W + 4);
        const int y = reflect101(y0 - 2 + ty, rows), x = reflect101(x0 - 2 + tx, cols);
        s_I[ty][tx] = (float)img[(size_t)y * pitch + x];
    }
    __syncthreads();
    const float s = (float)(1.0 / (4 * 3 * 255.0));  // Sobel scale for u8, ksize 3, blockSize 3
    const float s2 = 2.0f * s;
    // derivatives at virtual coords [y0-1, y0+ET_H+1) x [x0-1, x0+ET_W+1); positions outside the image take
    // the value at the REFLECT_101 position (the box filter's border rule on the covariance images)
    for (int k = tid; k < (ET_H + 2) * (ET_W + 2); k += 256) {
        const int ty = k / (ET_W + 2), tx = k % (ET_W + 2);
        int ry = reflect101(y0 - 1 + ty, rows) - (y0 - 2);  // tile-local index of the reflected centre
        int rx = reflect101(x0 - 1 + tx, cols) - (x0 - 2);
        ry = min(max(ry, 1), ET_H + 2);
        rx = min(max(rx, 1), ET_W + 2);
        const float a00 = s_I[ry - 1][rx - 1], a01 = s_I[ry - 1][rx], a02 = s_I[ry - 1][rx + 1];
        const float a10 = s_I[ry][rx - 1], a11 = s_I[ry][rx], a12 = s_I[ry][rx + 1];
        const float a20 = s_I[ry + 1][rx - 1], a21 = s_I[ry + 1][rx], a22 = s_I[ry + 1][rx + 1];
        // dx: row filter [-1 0 1] (exact), column filter [s 2s s] as fma(S0+S2, s, S1*2s)
        const float S0 = a02 - a00, S1 = a12 - a10, S2 = a22 - a20;
        s_dx[ty][tx] = __fmaf_rn(__fadd_rn(S0, S2), s, __fmul_rn(S1, s2));
        // dy: row filter [s 2s s] as fma(C, s, fma(B, 2s, A*s)), column filter [-1 0 1]
        const float R0 = __fmaf_rn(a02, s, __fmaf_rn(a01, s2, __fmul_rn(a00, s)));
        const float R2 = __fmaf_rn(a22, s, __fmaf_rn(a21, s2, __fmul_rn(a20, s)));
        s_dy[ty][tx] = __fsub_rn(R2, R0);
        (void)a11;
    }
    __syncthreads();
    uint32_t local_max = 0;  // enc_ordered(-NaN..) minimum
    const int lx = tid % ET_W;
#pragma unroll
    for (int rr = 0; rr < 2; rr++) {
        const int ly = tid / ET_W + rr * (ET_H / 2);
        const int x = x0 + lx, y = y0 + ly;
        if (x < cols && y < rows) {
            double sxx = 0, sxy = 0, syy = 0;
#pragma unroll
            for (int dy = 0; dy < 3; dy++) {
                double rxx = 0, rxy = 0, ryy = 0;
#pragma unroll
                for (int dx = 0; dx < 3; dx++) {
                    const float gx = s_dx[ly + dy][lx + dx], gy = s_dy[ly + dy][lx + dx];
                    rxx += (double)__fmul_rn(gx, gx);
                    rxy += (double)__fmul_rn(gx, gy);
                    ryy += (double)__fmul_rn(gy, gy);
                }
                sxx += rxx;
                sxy += rxy;
                syy += ryy;
            }
            const float a = __fmul_rn((float)sxx, 0.5f), b = (float)sxy, c = __fmul_rn((float)syy, 0.5f);
            const float t = __fsub_rn(a, c);
            const float lam = __fsub_rn(__fadd_rn(a, c), __fsqrt_rn(__fadd_rn(__fmul_rn(t, t), __fmul_rn(b, b))));
            eig[(size_t)y * cols + x] = lam;
            if (mask == nullptr || mask[(size_t)y * mask_pitch + x] != 0)
                local_max = max(local_max, enc_ordered(__float_as_uint(lam)));
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
    const int x = blockIdx.x * blockDim.x + threadIdx.x + 1;
    const int y = blockIdx.y + 1;
    bool is_cand = false;
    float v = 0.f;
    if (x < p.cols - 1 && y < p.rows - 1) {
        const float *c = eig + (size_t)y * p.cols + x;
        v = *c;
        if (v > thr && v != 0.f && (mask == nullptr || mask[(size_t)y * p.mask_pitch + x] != 0)) {
            is_cand = true;
#pragma unroll
            for (int dy = -1; dy <= 1; dy++)
#pragma unroll
                for (int dx = -1; dx <= 1; dx++) {
                    const float nb = c[dy * p.cols + dx];
                    const float tn = nb > thr ? nb : 0.f;
                    if (tn > v) is_cand = false;
                }
        }
    }
    const unsigned ballot = __ballot_sync(0xffffffffu, is_cand);
    if (ballot == 0) return;
    const int lane = threadIdx.x & 31;
    int base = 0;
    if (lane == __ffs(ballot) - 1) base = atomicAdd(&counters[0], __popc(ballot));
    base = __shfl_sync(0xffffffffu, base, __ffs(ballot) - 1);
    if (is_cand) {
        const int idx = base + __popc(ballot & ((1u << lane) - 1));
        const uint32_t venc = enc_ordered(__float_as_uint(v));
        if (idx < p.cap) {
            cand[idx] = ((unsigned long long)venc << 32) | (uint32_t)(y * p.cols + x);
            atomicAdd(&cell_count[(y / p.cell) * p.gw + (x / p.cell)], 1);
            atomicAdd(&hist[min((venc - tenc) >> shift, (uint32_t)(kCandHistBins - 1))], 1u);
        } else {
            counters[1] = 1;  // overflow: reported as LFVIO_ERR_CAPACITY by the caller
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
                   int cols, uint32_t *max_bits, cudaStream_t s) {
    LFVIO_CUDA_TRY(cudaMemsetAsync(max_bits, 0, sizeof(uint32_t), s));
    dim3 grid(cdiv(cols, ET_W), cdiv(rows, ET_H));
    min_eig_kernel<<<grid, 256, 0, s>>>(img, pitch, mask, mask_pitch, eig, rows, cols, max_bits);
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
    dim3 grid(cdiv(cols - 2, 256), rows - 2);
    candidates_kernel<<<grid, 256, 0, s>>>(eig, mask, max_bits, p, cs.cand, cs.cell_count, cs.hist, cs.counters);
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
