// K1 — CLAHE(clipLimit 3.0, 8x8 tiles), bit-exact with cv::CLAHE::apply on CV_8UC1.
//
// Reference call site: feature_tracker/src/feature_tracker.cpp:101-107 (cv::createCLAHE(3.0, Size(8,8))).
// Algorithm (OpenCV, restated in SURVEY.md Appendix B.3): per-tile 256-bin histogram -> clip ->
// redistribute (batch + strided residual) -> cumulative LUT (f32 scale, cvRound, saturate) ->
// per-pixel bilinear blend of the 4 neighbouring tile LUTs in UN-FUSED f32.
//
// Three launches: histogram (smem atomics, tiles x splits CTAs), LUT (64 CTAs), apply (persistent
// grid, all 64 LUTs = 16 KB staged in shared memory once per CTA, uchar4 vector I/O).
// HBM traffic: read W*H (hist) + read W*H + write W*H (apply) = 3 B/px.
#include "frontend.cuh"

namespace lfvio {

namespace {

struct ClaheGeom {
    int rows, cols;          // source
    int tile_w, tile_h;      // of the (possibly REFLECT_101-extended) image
    int splits;              // CTAs per tile in the histogram pass
};

__global__ void __launch_bounds__(256) clahe_hist_kernel(const uint8_t *__restrict__ src, int pitch, ClaheGeom g,
                                                         uint32_t *__restrict__ hist) {
    // one private histogram per warp (fewer atomics on one word), merged at the end; the slice of the tile is walked as ONE flat
    // range of 4-pixel chunks so that all 256 threads work whatever the tile width (a 160-pixel tile row has 40 chunks)
    __shared__ uint32_t sh[8][256];
    const int tile = blockIdx.x;
    const int ty = tile / kClaheTiles, tx = tile % kClaheTiles;
    for (int k = threadIdx.x; k < 8 * 256; k += blockDim.x) (&sh[0][0])[k] = 0;
    __syncthreads();
    uint32_t *my = sh[threadIdx.x >> 5];
    // this CTA's slice of tile rows
    const int r0 = (int)(((long long)g.tile_h * blockIdx.y) / g.splits);
    const int r1 = (int)(((long long)g.tile_h * (blockIdx.y + 1)) / g.splits);
    const int x_base = tx * g.tile_w;
    const bool interior_x = (x_base + g.tile_w <= g.cols);
    if (interior_x && ((x_base & 3) == 0) && ((g.tile_w & 3) == 0) && ((pitch & 3) == 0) && ((reinterpret_cast<uintptr_t>(src) & 3) == 0)) {
        const int w4 = g.tile_w >> 2, n = (r1 - r0) * w4;
        int r = r0 + (int)threadIdx.x / w4, x = (int)threadIdx.x % w4;
        const int dr = 256 / w4, dx = 256 % w4;
        for (int i = threadIdx.x; i < n; i += 256) {
            const int y = reflect101(ty * g.tile_h + r, g.rows);
            const uchar4 v = __ldg(reinterpret_cast<const uchar4 *>(src + (size_t)y * pitch + x_base) + x);
            atomicAdd(&my[v.x], 1u);
            atomicAdd(&my[v.y], 1u);
            atomicAdd(&my[v.z], 1u);
            atomicAdd(&my[v.w], 1u);
            r += dr;
            x += dx;
            if (x >= w4) {
                x -= w4;
                r++;
            }
        }
    } else {
        for (int r = r0; r < r1; r++) {
            const int y = reflect101(ty * g.tile_h + r, g.rows);
            const uint8_t *row = src + (size_t)y * pitch;
            for (int x = threadIdx.x; x < g.tile_w; x += blockDim.x) atomicAdd(&my[row[reflect101(x_base + x, g.cols)]], 1u);
        }
    }
    __syncthreads();
    uint32_t v = 0;
#pragma unroll
    for (int k = 0; k < 8; k++) v += sh[k][threadIdx.x];
    if (v) atomicAdd(&hist[tile * 256 + threadIdx.x], v);
}

__global__ void __launch_bounds__(256) clahe_lut_kernel(const uint32_t *__restrict__ hist, uint8_t *__restrict__ lut,
                                                        int clip_limit, float lut_scale) {
    __shared__ int s_scan[256];
    __shared__ int s_warp[8];
    const int tile = blockIdx.x, i = threadIdx.x, lane = i & 31, w = i >> 5;
    int h = (int)hist[tile * 256 + i];
    int over = 0;
    if (clip_limit > 0 && h > clip_limit) {
        over = h - clip_limit;
        h = clip_limit;
    }
    // clipped = sum(over)
    int c = warp_sum(over);
    if (lane == 0) s_warp[w] = c;
    __syncthreads();
    int clipped = 0;
#pragma unroll
    for (int k = 0; k < 8; k++) clipped += s_warp[k];
    if (clip_limit > 0) {
        int batch = clipped / 256;
        int residual = clipped - batch * 256;
        h += batch;
        if (residual != 0) {
            int step = 256 / residual;
            if (step < 1) step = 1;
            // for (i = 0; i < 256 && residual > 0; i += step, residual--) hist[i]++
            if (i % step == 0 && i / step < residual) h += 1;
        }
    }
    // inclusive scan over 256 bins
    int v = h;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        int t = __shfl_up_sync(0xffffffffu, v, o);
        if (lane >= o) v += t;
    }
    __syncthreads();
    if (lane == 31) s_warp[w] = v;
    __syncthreads();
    int base = 0;
    for (int k = 0; k < w; k++) base += s_warp[k];
    int sum = v + base;
    (void)s_scan;
    int r = cv_round(__fmul_rn((float)sum, lut_scale));
    lut[tile * 256 + i] = (uint8_t)min(max(r, 0), 255);
}

// Round 2: a CTA works inside ONE band of rows between two tile-centre lines (the rows that interpolate between tile rows ty1 and
// ty2: nine bands) and stages, for the nine column bands, the four LUT bytes a pixel of value v blends as ONE 32-bit word
// (l11 | l12 << 8 | l21 << 16 | l22 << 24, 9 KB): one shared load per pixel instead of four byte gathers out of the 16 KB of all
// 64 LUTs (139 k bank conflicts per launch at 3840 x 2880 in round 1).  Same un-fused f32 blend, bit-exact.
constexpr int kClaheBands = kClaheTiles + 1;

__global__ void __launch_bounds__(256) clahe_apply_kernel(const uint8_t *__restrict__ src, int src_pitch,
                                                          uint8_t *__restrict__ dst, int dst_pitch, ClaheGeom g,
                                                          const uint8_t *__restrict__ lut, int chunks_per_band) {
    __shared__ uint32_t s_tab[kClaheBands][256];
    const int band = blockIdx.y;  // rows [ceil((band - 0.5) tile_h), ceil((band + 0.5) tile_h)): ty1 = band - 1 (clamped), ty2 = band (clamped)
    const int ty1 = max(band - 1, 0), ty2 = min(band, kClaheTiles - 1);
    for (int k = threadIdx.x; k < kClaheBands * 256; k += blockDim.x) {
        const int xb = k >> 8, v = k & 255;
        const int tx1 = max(xb - 1, 0), tx2 = min(xb, kClaheTiles - 1);
        const uint32_t l11 = lut[(ty1 * kClaheTiles + tx1) * 256 + v], l12 = lut[(ty1 * kClaheTiles + tx2) * 256 + v];
        const uint32_t l21 = lut[(ty2 * kClaheTiles + tx1) * 256 + v], l22 = lut[(ty2 * kClaheTiles + tx2) * 256 + v];
        s_tab[xb][v] = l11 | (l12 << 8) | (l21 << 16) | (l22 << 24);
    }
    __syncthreads();
    const float inv_tw = __fdiv_rn(1.0f, (float)g.tile_w);
    const float inv_th = __fdiv_rn(1.0f, (float)g.tile_h);
    // first row y with floor(y / tile_h - 0.5) == band - 1, i.e. y >= (band - 0.5) tile_h: (2 band - 1) tile_h / 2 rounded up
    const int ya0 = max(0, ((2 * band - 1) * g.tile_h + 1) >> 1), ya1 = min(g.rows, ((2 * band + 1) * g.tile_h + 1) >> 1);
    const int chunks_per_row = (g.cols + 3) / 4;
    const bool vec_ok = ((src_pitch & 3) == 0) && ((dst_pitch & 3) == 0) && ((g.cols & 3) == 0);
    // rows of the band over blockIdx.x, chunks of a row over the threads: no division per chunk, the row weights once per row
    // a thread keeps its four columns (column weights and table rows once), the band's rows in turn
    for (int ch = threadIdx.x; ch < chunks_per_row; ch += blockDim.x) {
        const int x0 = 4 * ch;
        const int nvalid = min(4, g.cols - x0);
        float xa[4], xa1[4];
        int tx1f[4];
        const uint32_t *tab[4];
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const float txf = __fsub_rn(__fmul_rn((float)(x0 + k), inv_tw), 0.5f);
            tx1f[k] = cv_floor(txf);
            xa[k] = __fsub_rn(txf, (float)tx1f[k]);
            xa1[k] = __fsub_rn(1.0f, xa[k]);
            tab[k] = s_tab[min(max(tx1f[k] + 1, 0), kClaheBands - 1)];
        }
        for (int y = ya0 + blockIdx.x; y < ya1; y += gridDim.x) {
            const float tyf = __fsub_rn(__fmul_rn((float)y, inv_th), 0.5f);
            const int ty1f = cv_floor(tyf);
            const float ya = __fsub_rn(tyf, (float)ty1f);
            const float yb = __fsub_rn(1.0f, ya);
            // (the band limits above are the exact integer statement of ty1f == band - 1; should the f32 product ever land on the
            // other side of a boundary, the table of this band would be the wrong one, so such a row falls back to the byte LUTs)
            const bool band_ok = (ty1f == band - 1);
            uint8_t in[4], out[4];
            if (vec_ok) {
                const uchar4 v = *reinterpret_cast<const uchar4 *>(src + (size_t)y * src_pitch + x0);
                in[0] = v.x; in[1] = v.y; in[2] = v.z; in[3] = v.w;
            } else {
                for (int k = 0; k < nvalid; k++) in[k] = src[(size_t)y * src_pitch + x0 + k];
            }
#pragma unroll
            for (int k = 0; k < 4; k++) {
                if (k >= nvalid) break;
                const int sv = in[k];
                float l11, l12, l21, l22;
                if (band_ok) {
                    const uint32_t w = tab[k][sv];
                    l11 = (float)(w & 255u);
                    l12 = (float)((w >> 8) & 255u);
                    l21 = (float)((w >> 16) & 255u);
                    l22 = (float)(w >> 24);
                } else {
                    const int a1 = max(ty1f, 0), a2 = min(ty1f + 1, kClaheTiles - 1), b1 = max(tx1f[k], 0), b2 = min(tx1f[k] + 1, kClaheTiles - 1);
                    l11 = (float)lut[(a1 * kClaheTiles + b1) * 256 + sv];
                    l12 = (float)lut[(a1 * kClaheTiles + b2) * 256 + sv];
                    l21 = (float)lut[(a2 * kClaheTiles + b1) * 256 + sv];
                    l22 = (float)lut[(a2 * kClaheTiles + b2) * 256 + sv];
                }
                const float top = __fadd_rn(__fmul_rn(l11, xa1[k]), __fmul_rn(l12, xa[k]));
                const float bot = __fadd_rn(__fmul_rn(l21, xa1[k]), __fmul_rn(l22, xa[k]));
                const float res = __fadd_rn(__fmul_rn(top, yb), __fmul_rn(bot, ya));
                out[k] = (uint8_t)min(max(cv_round(res), 0), 255);
            }
            if (vec_ok) {
                *reinterpret_cast<uchar4 *>(dst + (size_t)y * dst_pitch + x0) = make_uchar4(out[0], out[1], out[2], out[3]);
            } else {
                for (int k = 0; k < nvalid; k++) dst[(size_t)y * dst_pitch + x0 + k] = out[k];
            }
        }
    }
    (void)chunks_per_band;
}

}  // namespace

int launch_clahe(const uint8_t *src, int src_pitch, uint8_t *dst, int dst_pitch, int rows, int cols,
                 float clip_limit, uint32_t *hist, uint8_t *lut, int n_sm, cudaStream_t s) {
    ClaheGeom g;
    g.rows = rows;
    g.cols = cols;
    int ext_rows = rows, ext_cols = cols;
    if (cols % kClaheTiles != 0 || rows % kClaheTiles != 0) {
        // OpenCV pads BOTH dimensions: bottom = tiles - rows%tiles, right = tiles - cols%tiles
        ext_rows = rows + (kClaheTiles - rows % kClaheTiles);
        ext_cols = cols + (kClaheTiles - cols % kClaheTiles);
    }
    g.tile_w = ext_cols / kClaheTiles;
    g.tile_h = ext_rows / kClaheTiles;
    const int area = g.tile_w * g.tile_h;
    // enough CTAs to fill the machine: 64 tiles x splits ~ 4 CTAs per SM
    g.splits = max(1, min(g.tile_h, cdiv(4 * n_sm, kClaheTiles * kClaheTiles)));
    int clip = 0;
    if (clip_limit > 0.0f) {
        clip = (int)((double)clip_limit * area / 256);  // static_cast<int>(clipLimit_ * tileSizeTotal / histSize)
        if (clip < 1) clip = 1;
    }
    const float lut_scale = (float)(255) / (float)area;  // static_cast<float>(histSize - 1) / tileSizeTotal
    LFVIO_CUDA_TRY(cudaMemsetAsync(hist, 0, sizeof(uint32_t) * kClaheTiles * kClaheTiles * 256, s));
    clahe_hist_kernel<<<dim3(kClaheTiles * kClaheTiles, g.splits), 256, 0, s>>>(src, src_pitch, g, hist);
    LFVIO_LAUNCH_CHECK();
    clahe_lut_kernel<<<kClaheTiles * kClaheTiles, 256, 0, s>>>(hist, lut, clip, lut_scale);
    LFVIO_LAUNCH_CHECK();
    // nine row bands x enough CTAs per band (each takes rows of the band in turn) to put about four CTAs on every SM
    const int per_band = std::max(1, std::min(g.tile_h + 1, cdiv(4 * n_sm, kClaheTiles + 1)));
    clahe_apply_kernel<<<dim3(per_band, kClaheTiles + 1), 256, 0, s>>>(src, src_pitch, dst, dst_pitch, g, lut, 0);
    LFVIO_LAUNCH_CHECK();
    return LFVIO_OK;
}

}  // namespace lfvio
