// TMA (cp.async.bulk.tensor) + mbarrier helpers shared by the tile kernels of the front end (lk.cu, corners.cu).
#pragma once
#include <cuda.h>
#include <stdint.h>

namespace lfvio {

#ifdef __CUDACC__
__device__ __forceinline__ uint32_t tma_smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(tma_smem_u32(bar)), "r"(count));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(tma_smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    uint32_t done = 0;
    while (!done) {
        asm volatile(
            "{\n\t"
            ".reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t"
            "}"
            : "=r"(done)
            : "r"(tma_smem_u32(bar)), "r"(parity)
            : "memory");
    }
}
// one 48 x 44 u8 box whose top-left pixel is (x, y); out-of-range bytes (x >= cols) are zero-filled by the TMA unit
__device__ __forceinline__ void tma_load_tile(void *dst, const CUtensorMap *tmap, int x, int y, uint64_t *bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(tma_smem_u32(dst)),
        "l"(reinterpret_cast<uint64_t>(tmap)), "r"(x), "r"(y), "r"(tma_smem_u32(bar))
        : "memory");
}

#endif

// cuTensorMapEncodeTiled for a u8 image (row pitch in bytes, 16-byte aligned base and pitch) with a box of box_w x box_h bytes;
// false when the driver entry point is missing or the encode fails (the callers then stage their tiles with plain loads).
bool encode_u8_tmap(CUtensorMap *out, const uint8_t *data, int rows, int cols, int pitch, int box_w, int box_h);

}  // namespace lfvio
