// One-way producer -> consumer hand-over latency between two warps of one CTA (sm_100a): named barrier (arrive / sync),
// mbarrier (arrive / try_wait.parity), volatile shared flag.  Producer stamps clock64() just before signalling, consumer
// stamps it right after it has read the payload.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *b, int n) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(n)); }
__device__ __forceinline__ void mbar_arrive(uint64_t *b) { asm volatile("mbarrier.arrive.release.cta.shared::cta.b64 _, [%0];" ::"r"(smem_u32(b)) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t *b, int parity) {
    asm volatile("{\n.reg .pred p;\nW: mbarrier.try_wait.parity.acquire.cta.shared::cta.b64 p, [%0], %1;\n@p bra D;\nbra W;\nD:\n}" ::"r"(smem_u32(b)), "r"(parity) : "memory");
}
constexpr int N = 64;
__global__ void k(long long *out, int mode) {
    __shared__ uint64_t bars[16];
    __shared__ volatile int flag;
    __shared__ double box[32];
    __shared__ long long stamp[N];
    const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) { for (int i = 0; i < 16; i++) mbar_init(&bars[i], 1); flag = 0; }
    __syncthreads();
    long long acc = 0;
    for (int i = 0; i < N; i++) {
        if (wid == 0) {
            // some dependent work so that the consumer is already waiting
            double x = i + lane;
            for (int q = 0; q < 40; q++) x = fma(x, 1.0000001, 1e-9);
            box[lane] = x;
            if (mode == 0) {
                __syncwarp();
                if (lane == 0) stamp[i] = clock64();
                asm volatile("bar.arrive %0, 64;" ::"r"(1 + (i & 7)) : "memory");
            } else if (mode == 1) {
                __syncwarp();
                if (lane == 0) { stamp[i] = clock64(); mbar_arrive(&bars[i & 15]); }
            } else {
                __threadfence_block();
                __syncwarp();
                if (lane == 0) { stamp[i] = clock64(); flag = i + 1; }
            }
            asm volatile("bar.sync 9, 64;" ::: "memory");  // back-pressure once per round (not timed)
        } else {
            if (mode == 0) asm volatile("bar.sync %0, 64;" ::"r"(1 + (i & 7)) : "memory");
            else if (mode == 1) mbar_wait(&bars[i & 15], (i >> 4) & 1);
            else { while (flag != i + 1) {} }
            const double v = box[lane];
            const long long t = clock64();
            if (lane == 0 && v > -1.0) acc += t - stamp[i];
            asm volatile("bar.sync 9, 64;" ::: "memory");
        }
    }
    if (wid == 1 && lane == 0) out[mode] = acc;
}
int main() {
    long long *o, h[3];
    cudaMalloc(&o, 64);
    for (int mode = 0; mode < 3; mode++)
        for (int r = 0; r < 2; r++) { k<<<1, 64>>>(o, mode); cudaDeviceSynchronize(); }
    cudaMemcpy(h, o, 24, cudaMemcpyDeviceToHost);
    printf("one-way hand-over: bar.arrive/bar.sync %.1f cycles, mbarrier arrive/try_wait %.1f cycles, volatile flag %.1f cycles (%s)\n", h[0] / (double)N, h[1] / (double)N, h[2] / (double)N, cudaGetErrorString(cudaGetLastError()));
}
