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
