// Stand-alone probe: which cp.async.bulk.tensor.2d (u8) configurations run on this GPU.
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
struct alignas(64) Maps { CUtensorMap m[4]; int valid; };
struct Pad { char c[200]; };
__device__ __forceinline__ uint32_t s32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
template <int BW, int BH>
__device__ void body(const CUtensorMap *tm, int x, int y, uint8_t *out) {
    __shared__ __align__(128) uint8_t tile[BH * BW];
    __shared__ uint64_t bar;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(&bar)), "r"(BW * BH) : "memory");
        asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
                         s32(tile)),
                     "l"(reinterpret_cast<uint64_t>(tm)), "r"(x), "r"(y), "r"(s32(&bar))
                     : "memory");
    }
    uint32_t done = 0;
    while (!done)
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done)
                     : "r"(s32(&bar)), "r"(0)
                     : "memory");
    for (int k = threadIdx.x; k < BW * BH; k += blockDim.x) out[k] = tile[k];
}
template <int BW, int BH>
__global__ void k_single(const __grid_constant__ CUtensorMap tm, int x, int y, uint8_t *out) { body<BW, BH>(&tm, x, y, out); }
template <int BW, int BH>
__global__ void k_struct(Pad pad, const __grid_constant__ Maps tm, int lvl, int x, int y, uint8_t *out) {
    if (pad.c[0] == 77) return;
    body<BW, BH>(&tm.m[lvl], x, y, out);
}
template <int BW, int BH>
__global__ void k_global(const CUtensorMap *tm, int x, int y, uint8_t *out) { body<BW, BH>(tm, x, y, out); }

typedef CUresult (*EncodeTiled)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *, const cuuint32_t *,
                                const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
int main(int argc, char **argv) {
    const char *mode = argc > 1 ? argv[1] : "single48";
    const int cols = 640, rows = 480, pitch = 640;
    uint8_t *img, *out;
    cudaMalloc(&img, pitch * rows);
    cudaMalloc(&out, 64 * 64);
    uint8_t *h = (uint8_t *)malloc(pitch * rows);
    for (int i = 0; i < pitch * rows; i++) h[i] = (uint8_t)((i * 7 + (i / pitch) * 13) & 255);
    cudaMemcpy(img, h, pitch * rows, cudaMemcpyHostToDevice);
    void *fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
    if (!fn) { printf("%s: no entry point\n", mode); return 1; }
    const int bw = strstr(mode, "64") ? 64 : 48, bh = 44;
    Maps maps;
    memset(&maps, 0, sizeof(maps));
    for (int l = 0; l < 4; l++) {
        cuuint64_t gd[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
        cuuint64_t gs[1] = {(cuuint64_t)pitch};
        cuuint32_t box[2] = {(cuuint32_t)bw, (cuuint32_t)bh};
        cuuint32_t es[2] = {1, 1};
        CUresult r = ((EncodeTiled)fn)(&maps.m[l], CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, img, gd, gs, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                       CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { printf("%s: encode failed %d\n", mode, (int)r); return 1; }
    }
    const int x = argc > 2 ? atoi(argv[2]) : 101, y = argc > 3 ? atoi(argv[3]) : 57;
    Pad pad; memset(&pad, 0, sizeof(pad));
    if (!strcmp(mode, "single48")) k_single<48, 44><<<1, 128>>>(maps.m[0], x, y, out);
    else if (!strcmp(mode, "single64")) k_single<64, 44><<<1, 128>>>(maps.m[0], x, y, out);
    else if (!strcmp(mode, "struct48")) k_struct<48, 44><<<1, 128>>>(pad, maps, 2, x, y, out);
    else if (!strcmp(mode, "struct64")) k_struct<64, 44><<<1, 128>>>(pad, maps, 2, x, y, out);
    else {
        CUtensorMap *d; cudaMalloc(&d, sizeof(CUtensorMap));
        cudaMemcpy(d, &maps.m[0], sizeof(CUtensorMap), cudaMemcpyHostToDevice);
        if (bw == 48) k_global<48, 44><<<1, 128>>>(d, x, y, out); else k_global<64, 44><<<1, 128>>>(d, x, y, out);
    }
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("%s x=%d y=%d: %s\n", mode, x, y, cudaGetErrorString(e)); return 1; }
    uint8_t ho[64 * 64];
    cudaMemcpy(ho, out, bw * bh, cudaMemcpyDeviceToHost);
    int bad = 0;
    for (int r = 0; r < bh; r++) for (int c = 0; c < bw; c++) {
        const int gx = x + c, gy = y + r;
        const uint8_t want = (gx < cols && gy < rows && gx >= 0 && gy >= 0) ? h[gy * pitch + gx] : 0;
        bad += ho[r * bw + c] != want;
    }
    printf("%s x=%d y=%d: ok, mismatches=%d\n", mode, x, y, bad);
    return 0;
}
