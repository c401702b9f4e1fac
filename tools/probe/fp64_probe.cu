// Micro-probe: fp64 latencies / throughput on one SM (cycles via clock64).
#include <cuda_runtime.h>
#include <cstdio>
__global__ void k_lat(double *out, long long *cyc, double seed) {
    double a = seed + threadIdx.x * 1e-9, b = 1.0000001, c = 1e-9;
    long long t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < 256; i++) {
#pragma unroll
        for (int k = 0; k < 16; k++) a = fma(a, b, c);
    }
    long long t1 = clock64();
    double r = a;
#pragma unroll 1
    for (int i = 0; i < 256; i++) {
#pragma unroll
        for (int k = 0; k < 4; k++) r = rsqrt(r + 1.5);
    }
    long long t2 = clock64();
    double s = r;
#pragma unroll 1
    for (int i = 0; i < 256; i++) {
#pragma unroll
        for (int k = 0; k < 16; k++) s = __shfl_sync(0xffffffffu, s, (k * 7 + 1) & 31) + 1.0;
    }
    long long t3 = clock64();
    float f = (float)s;
#pragma unroll 1
    for (int i = 0; i < 256; i++) {
#pragma unroll
        for (int k = 0; k < 16; k++) f = fmaf(f, 1.0001f, 0.5f);
    }
    long long t4 = clock64();
    double q = s;
#pragma unroll 1
    for (int i = 0; i < 256; i++) {
#pragma unroll
        for (int k = 0; k < 4; k++) q = 1.0 / (q + 1.5);
    }
    long long t5 = clock64();
    double w = q;
#pragma unroll 1
    for (int i = 0; i < 256; i++) {
#pragma unroll
        for (int k = 0; k < 4; k++) w = (double)rsqrtf((float)(w + 1.5));
    }
    long long t6 = clock64();
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        cyc[0] = (t1 - t0); cyc[1] = (t2 - t1); cyc[2] = (t3 - t2); cyc[3] = t4 - t3; cyc[4] = t5 - t4; cyc[5] = t6 - t5;
    }
    out[threadIdx.x] = a + r + s + f + q + w;
}
// throughput: each thread 8 independent chains
__global__ void k_thr(double *out, long long *cyc, double seed) {
    double a[8];
    for (int k = 0; k < 8; k++) a[k] = seed + k + threadIdx.x * 1e-9;
    const double b = 1.0000001, c = 1e-9;
    __syncthreads();
    long long t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < 512; i++) {
#pragma unroll
        for (int k = 0; k < 8; k++) a[k] = fma(a[k], b, c);
    }
    __syncthreads();
    long long t1 = clock64();
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
    double s = 0;
    for (int k = 0; k < 8; k++) s += a[k];
    out[threadIdx.x] = s;
}
int main() {
    double *out; long long *cyc, h[8];
    cudaMalloc(&out, 8192); cudaMalloc(&cyc, 64);
    k_lat<<<1, 32>>>(out, cyc, 1.0); k_lat<<<1, 32>>>(out, cyc, 1.0);
    cudaMemcpy(h, cyc, 48, cudaMemcpyDeviceToHost);
    printf("dependent latency (cycles/op): dfma %.1f  rsqrt(double) %.1f  shfl64+dadd %.1f  ffma %.1f  ddiv+dadd %.1f  rsqrtf-cvt+dadd %.1f\n", h[0] / 4096.0, h[1] / 1024.0,
           h[2] / 4096.0, h[3] / 4096.0, h[4] / 1024.0, h[5] / 1024.0);
    for (int nt = 32; nt <= 1024; nt *= 2) {
        k_thr<<<1, nt>>>(out, cyc, 1.0); k_thr<<<1, nt>>>(out, cyc, 1.0);
        cudaMemcpy(h, cyc, 8, cudaMemcpyDeviceToHost);
        printf("dfma throughput, %4d threads: %.2f dfma/clk/SM\n", nt, 512.0 * 8 * nt / h[0]);
    }
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
