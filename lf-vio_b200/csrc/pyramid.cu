// K2 — Gaussian pyramid for pyramidal LK, bit-exact with cv::pyrDown on CV_8UC1.
//
// Reference call site: inside cv::calcOpticalFlowPyrLK (feature_tracker/src/feature_tracker.cpp:127), which
// rebuilds BOTH pyramids on every frame; here the previous frame's pyramid stays resident in HBM and only
// the new frame's is built.  Algorithm: separable [1 4 6 4 1], REFLECT_101,
// integer sum, (sum + 128) >> 8; level size (w+1)/2 x (h+1)/2; stop when a level would be <= winSize.
// Scharr derivatives are NOT materialised: the LK kernel recomputes them from its shared-memory tile.
//
// Tile kernel: a CTA produces 32x8 output pixels from a 67x19 input tile staged in shared memory
// (separable: horizontal pass into smem, vertical pass out).  Traffic: read 1 B/px, write 1/4 B/px.
#include <cstring>

#include "frontend.cuh"

namespace lfvio {

namespace {

constexpr int TW = 32, TH = 8;
constexpr int IN_W = 2 * TW + 3, IN_H = 2 * TH + 3;

__global__ void __launch_bounds__(TW *TH) pyr_down_kernel(ImgLevel src, ImgLevel dst) {
    __shared__ uint8_t s_in[IN_H][IN_W + 1];
    __shared__ uint16_t s_row[IN_H][TW];
    const int ox0 = blockIdx.x * TW, oy0 = blockIdx.y * TH;
    const int ix0 = 2 * ox0 - 2, iy0 = 2 * oy0 - 2;
    const int tid = threadIdx.y * TW + threadIdx.x;
    for (int k = tid; k < IN_W * IN_H; k += TW * TH) {
        const int ty = k / IN_W, tx = k % IN_W;
        const int y = reflect101(iy0 + ty, src.rows), x = reflect101(ix0 + tx, src.cols);
        s_in[ty][tx] = src.data[(size_t)y * src.pitch + x];
    }
    __syncthreads();
    for (int k = tid; k < IN_H * TW; k += TW * TH) {
        const int ty = k / TW, ox = k % TW;
        const uint8_t *p = &s_in[ty][2 * ox];
        s_row[ty][ox] = (uint16_t)(p[0] + 4 * p[1] + 6 * p[2] + 4 * p[3] + p[4]);
    }
    __syncthreads();
    const int ox = ox0 + threadIdx.x, oy = oy0 + threadIdx.y;
    if (ox < dst.cols && oy < dst.rows) {
        const int ty = 2 * threadIdx.y;
        const int v = s_row[ty][threadIdx.x] + 4 * s_row[ty + 1][threadIdx.x] + 6 * s_row[ty + 2][threadIdx.x] +
                      4 * s_row[ty + 3][threadIdx.x] + s_row[ty + 4][threadIdx.x];
        dst.data[(size_t)oy * dst.pitch + ox] = (uint8_t)((v + 128) >> 8);
    }
}

}  // namespace

int launch_pyr_down(const ImgLevel &src, const ImgLevel &dst, cudaStream_t s) {
    dim3 grid(cdiv(dst.cols, TW), cdiv(dst.rows, TH));
    pyr_down_kernel<<<grid, dim3(TW, TH), 0, s>>>(src, dst);
    LFVIO_LAUNCH_CHECK();
    return LFVIO_OK;
}

int launch_build_pyramid(const Pyramid &p, cudaStream_t s) {
    for (int l = 1; l < p.levels; l++) LFVIO_TRY(launch_pyr_down(p.lvl[l - 1], p.lvl[l], s));
    return LFVIO_OK;
}

void make_pyramid_tmaps(const Pyramid &p, PyrTmaps *out) {
    typedef CUresult (*EncodeTiled)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                    const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    memset(out, 0, sizeof(*out));
    void *fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) != cudaSuccess || !fn ||
        qres != cudaDriverEntryPointSuccess) {
        cudaGetLastError();
        return;
    }
    EncodeTiled encode = reinterpret_cast<EncodeTiled>(fn);
    for (int l = 0; l < p.levels && l < kTmaLevels; l++) {
        const ImgLevel &L = p.lvl[l];
        if ((reinterpret_cast<uintptr_t>(L.data) & 15) || (L.pitch & 15)) return;  // TMA needs 16-byte aligned base and pitch
        const cuuint64_t gdim[2] = {(cuuint64_t)L.cols, (cuuint64_t)L.rows};
        const cuuint64_t gstride[1] = {(cuuint64_t)L.pitch};
        const cuuint32_t box[2] = {64, 44};  // the 44 x 44 LK tile + 16-byte alignment slack of the row start
        const cuuint32_t estride[2] = {1, 1};
        if (encode(&out->m[l], CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, L.data, gdim, gstride, box, estride, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
            return;
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
