// Dependent-chain latencies of the primitives on the Cholesky pivot chain (one warp), in SM clocks.
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(double *out, long long *clk, double x0) {
    const int lane = threadIdx.x;
    double x = x0 + lane * 1e-3, y = 1.000001;
    long long t[12];
    __shared__ double sh[64];
    sh[lane] = x;
    __syncwarp();
    t[0] = clock64();
#pragma unroll 1
    for (int i = 0; i < 256; i++) x = fma(x, y, 1e-9);
    t[1] = clock64();
#pragma unroll 1
    for (int i = 0; i < 256; i++) x = rsqrt(x + 2.0);
    t[2] = clock64();
#pragma unroll 1
    for (int i = 0; i < 256; i++) x = __shfl_sync(0xffffffffu, x, (lane + 1) & 31);
    t[3] = clock64();
#pragma unroll 1
    for (int i = 0; i < 256; i++) { sh[lane] = x; __syncwarp(); x = sh[(lane + 1) & 31]; __syncwarp(); }
    t[4] = clock64();
#pragma unroll 1
    for (int i = 0; i < 256; i++) x = 1.0 / (x + 2.0);
    t[5] = clock64();
#pragma unroll 1
    for (int i = 0; i < 256; i++) x = sqrt(x + 2.0);
    t[6] = clock64();
    // 8 independent DFMA chains (issue rate of one warp)
    double a0 = x, a1 = x + 1, a2 = x + 2, a3 = x + 3, a4 = x + 4, a5 = x + 5, a6 = x + 6, a7 = x + 7;
#pragma unroll 1
    for (int i = 0; i < 256; i++) {
        a0 = fma(a0, y, 1e-9); a1 = fma(a1, y, 1e-9); a2 = fma(a2, y, 1e-9); a3 = fma(a3, y, 1e-9);
        a4 = fma(a4, y, 1e-9); a5 = fma(a5, y, 1e-9); a6 = fma(a6, y, 1e-9); a7 = fma(a7, y, 1e-9);
    }
    t[7] = clock64();
    // fp32 rsqrt seed + 2 Newton steps in fp64
#pragma unroll 1
    for (int i = 0; i < 256; i++) {
        const double d = x + 2.0;
        double r = (double)rsqrtf((float)d);
        r = r * fma(-0.5 * d * r, r, 1.5);
        r = r * fma(-0.5 * d * r, r, 1.5);
        x = r;
    }
    t[8] = clock64();
    x += a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
    out[lane] = x;
    if (lane == 0) for (int i = 0; i < 9; i++) clk[i] = t[i];
}
int main() {
    double *o; long long *c, h[12];
    cudaMalloc(&o, 256 * 8); cudaMalloc(&c, 12 * 8);
    for (int rep = 0; rep < 2; rep++) { k<<<1, 32>>>(o, c, 1.5); cudaDeviceSynchronize(); }
    cudaMemcpy(h, c, 12 * 8, cudaMemcpyDeviceToHost);
    const char *n[8] = {"dfma dependent", "rsqrt(double)", "shfl", "sts+syncwarp+lds+syncwarp", "1.0/x", "sqrt(double)", "8 independent dfma (per group of 8)", "rsqrtf seed + 2 NR (fp64)"};
    for (int i = 0; i < 8; i++) printf("%-40s %7.1f cycles\n", n[i], (h[i + 1] - h[i]) / 256.0);
    // 16-lane variant of issue rate
    return 0;
}
