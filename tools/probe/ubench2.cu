#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(long long *out) {
    long long t0 = clock64();
    unsigned long long g = 0, acc = 0;
#pragma unroll 1
    for (int i = 0; i < 64; i++) { asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g)); acc += g; }
    long long t1 = clock64();
    long long c = 0;
#pragma unroll 1
    for (int i = 0; i < 64; i++) { c += clock64(); }
    long long t2 = clock64();
    __shared__ double sh[32];
    sh[threadIdx.x & 31] = 1.0;
    __syncthreads();
    long long t3 = clock64();
#pragma unroll 1
    for (int i = 0; i < 64; i++) __syncthreads();
    long long t4 = clock64();
#pragma unroll 1
    for (int i = 0; i < 64; i++) atomicAdd(&sh[(threadIdx.x + i) & 31], 1.0);
    long long t5 = clock64();
    if (threadIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t1; out[2] = t4 - t3; out[3] = t5 - t4; out[4] = (long long)acc + c; }
}
int main() {
    long long *o, h[5];
    cudaMalloc(&o, 64);
    for (int r = 0; r < 2; r++) { k<<<1, 512>>>(o); cudaDeviceSynchronize(); }
    cudaMemcpy(h, o, 40, cudaMemcpyDeviceToHost);
    printf("globaltimer read %.1f cycles\nclock64 read %.1f cycles\n__syncthreads (512 thr) %.1f cycles\nshared atomicAdd(double), 32-way distinct per warp, 16 warps %.1f cycles\n", h[0] / 64.0, h[1] / 64.0, h[2] / 64.0, h[3] / 64.0);
}
