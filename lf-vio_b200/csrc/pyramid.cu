// K2 — Gaussian pyramid for pyramidal LK, bit-exact with cv::pyrDown on CV_8UC1.
//
// Reference call site: inside cv::calcOpticalFlowPyrLK (feature_tracker/src/feature_tracker.cpp:127), which
// rebuilds BOTH pyramids on every frame; here the previous frame's pyramid stays resident in HBM and only
// the new frame's is built.  Algorithm: separable [1 4 6 4 1], REFLECT_101,
// integer sum, (sum + 128) >> 8; level size (w+1)/2 x (h+1)/2; stop when a level would be <= winSize.
// Scharr derivatives are NOT materialised: the LK kernel recomputes them from its shared-memory tile.
//
// Register kernel (round 2): a thread produces a 4 x 2 block of output pixels from a 12 x 7 input window held in registers
// (seven rows of three aligned 32-bit loads; rows are shared between the two output rows, columns between the four outputs);
// no shared memory, no barrier, one 32-bit store per output row.  The window of a block away from the image border needs no
// reflection: only border blocks take the slow path.  Traffic: read 1 B/px (re-reads of the window overlap are L1 hits),
// write 1/4 B/px.
#include <cstring>

#include "frontend.cuh"
#include "tma.cuh"

namespace lfvio {

namespace {

constexpr int PD_THREADS = 128;

__device__ __forceinline__ uint32_t hsum5(uint32_t a, uint32_t b, uint32_t c, uint32_t d, uint32_t e) { return a + 4 * b + 6 * c + 4 * d + e; }

__global__ void __launch_bounds__(PD_THREADS) pyr_down_kernel(ImgLevel src, ImgLevel dst) {
    const int bx = blockIdx.x * PD_THREADS + threadIdx.x;  // block of 4 output columns
    const int by = blockIdx.y;                             // block of 2 output rows
    const int ox = 4 * bx, oy = 2 * by;
    if (ox >= dst.cols || oy >= dst.rows) return;
    const int ix = 2 * ox - 2, iy = 2 * oy - 2;  // window: input columns ix .. ix+10, rows iy .. iy+6
    uint32_t h[7][4];                            // horizontally filtered rows
    const bool interior = ix >= 2 && ix + 14 <= src.cols && iy >= 0 && iy + 7 <= src.rows && ((src.pitch & 3) == 0) &&
                          ((reinterpret_cast<uintptr_t>(src.data) & 3) == 0);
    if (interior) {
        // ix - 2 is a multiple of 4: three aligned words cover bytes ix-2 .. ix+9; byte ix+10 comes from the fourth
        const uint8_t *base = src.data + (size_t)iy * src.pitch + (ix - 2);
#pragma unroll
        for (int r = 0; r < 7; r++) {
            const uint32_t *w = reinterpret_cast<const uint32_t *>(base + (size_t)r * src.pitch);
            const uint32_t w0 = __ldg(w), w1 = __ldg(w + 1), w2 = __ldg(w + 2), w3 = __ldg(w + 3);
            uint32_t p[11];
            p[0] = (w0 >> 16) & 255u; p[1] = w0 >> 24;
            p[2] = w1 & 255u; p[3] = (w1 >> 8) & 255u; p[4] = (w1 >> 16) & 255u; p[5] = w1 >> 24;
            p[6] = w2 & 255u; p[7] = (w2 >> 8) & 255u; p[8] = (w2 >> 16) & 255u; p[9] = w2 >> 24;
            p[10] = w3 & 255u;
#pragma unroll
            for (int q = 0; q < 4; q++) h[r][q] = hsum5(p[2 * q], p[2 * q + 1], p[2 * q + 2], p[2 * q + 3], p[2 * q + 4]);
        }
    } else {
#pragma unroll
        for (int r = 0; r < 7; r++) {
            const uint8_t *row = src.data + (size_t)reflect101(iy + r, src.rows) * src.pitch;
            uint32_t p[11];
#pragma unroll
            for (int c = 0; c < 11; c++) p[c] = row[reflect101(ix + c, src.cols)];
#pragma unroll
            for (int q = 0; q < 4; q++) h[r][q] = hsum5(p[2 * q], p[2 * q + 1], p[2 * q + 2], p[2 * q + 3], p[2 * q + 4]);
        }
    }
#pragma unroll
    for (int rr = 0; rr < 2; rr++) {
        if (oy + rr >= dst.rows) break;
        uint8_t o[4];
#pragma unroll
        for (int q = 0; q < 4; q++) o[q] = (uint8_t)((hsum5(h[2 * rr][q], h[2 * rr + 1][q], h[2 * rr + 2][q], h[2 * rr + 3][q], h[2 * rr + 4][q]) + 128u) >> 8);
        uint8_t *out = dst.data + (size_t)(oy + rr) * dst.pitch + ox;
        if (ox + 4 <= dst.cols && ((dst.pitch & 3) == 0) && ((reinterpret_cast<uintptr_t>(dst.data) & 3) == 0)) {
            *reinterpret_cast<uint32_t *>(out) = o[0] | (o[1] << 8) | (o[2] << 16) | ((uint32_t)o[3] << 24);
        } else {
            for (int q = 0; q < 4 && ox + q < dst.cols; q++) out[q] = o[q];
        }
    }
}

}  // namespace

int launch_pyr_down(const ImgLevel &src, const ImgLevel &dst, cudaStream_t s) {
    dim3 grid(cdiv(cdiv(dst.cols, 4), PD_THREADS), cdiv(dst.rows, 2));
    pyr_down_kernel<<<grid, PD_THREADS, 0, s>>>(src, dst);
    LFVIO_LAUNCH_CHECK();
    return LFVIO_OK;
}

int launch_build_pyramid(const Pyramid &p, cudaStream_t s) {
    for (int l = 1; l < p.levels; l++) LFVIO_TRY(launch_pyr_down(p.lvl[l - 1], p.lvl[l], s));
    return LFVIO_OK;
}

bool encode_u8_tmap(CUtensorMap *out, const uint8_t *data, int rows, int cols, int pitch, int box_w, int box_h) {
    typedef CUresult (*EncodeTiled)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                    const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    static EncodeTiled encode = [] {
        void *fn = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) != cudaSuccess || !fn ||
            qres != cudaDriverEntryPointSuccess) {
            cudaGetLastError();
            fn = nullptr;
        }
        return reinterpret_cast<EncodeTiled>(fn);
    }();
    if (!encode) return false;
    if ((reinterpret_cast<uintptr_t>(data) & 15) || (pitch & 15)) return false;  // TMA needs 16-byte aligned base and pitch
    const cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    const cuuint64_t gstride[1] = {(cuuint64_t)pitch};
    const cuuint32_t box[2] = {(cuuint32_t)box_w, (cuuint32_t)box_h};
    const cuuint32_t estride[2] = {1, 1};
    return encode(out, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, const_cast<uint8_t *>(data), gdim, gstride, box, estride, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

void make_pyramid_tmaps(const Pyramid &p, PyrTmaps *out) {
    memset(out, 0, sizeof(*out));
    for (int l = 0; l < p.levels && l < kTmaLevels; l++) {
        const ImgLevel &L = p.lvl[l];
        // box: the 44 x 44 LK tile + 16-byte alignment slack of the row start
        if (!encode_u8_tmap(&out->m[l], L.data, L.rows, L.cols, L.pitch, 64, 44)) return;
    }
    out->valid = 1;
}

void make_pyramid_layout(Pyramid *p, int rows, int cols, int max_level, int win) {
    p->levels = 1;
    p->lvl[0].rows = rows;
    p->lvl[0].cols = cols;
    int r = rows, c = cols;
    for (int l = 1; l <= max_level && l < kMaxPyrLevels; l++) {
        r = (r + 1) / 2;
        c = (c + 1) / 2;
        if (c <= win || r <= win) break;  // buildOpticalFlowPyramid early stop
        p->lvl[l].rows = r;
        p->lvl[l].cols = c;
        p->levels = l + 1;
    }
}

}  // namespace lfvio
