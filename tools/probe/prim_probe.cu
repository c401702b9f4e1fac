// Latency / throughput of the primitives the pivot chain of the reduced-system factorisation is made of (sm_100a).
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o prim_probe prim_probe.cu && ./prim_probe
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ double rcp_seed(double x) {  // MUFU.RCP64H on the high word
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    return r;
}
__device__ __forceinline__ double rcp_nr2(double x) {  // seed + 2 Newton steps (4 dependent DFMA)
    double r = rcp_seed(x);
    double e = fma(-x, r, 1.0);
    r = fma(r, e, r);
    e = fma(-x, r, 1.0);
    r = fma(r, e, r);
    return r;
}
__device__ __forceinline__ double rcp_nr2b(double x) {  // seed + one quadratic-in-e step (3 dependent DFMA): r (1 + e + e^2), then a last correction
    double r = rcp_seed(x);
    const double e = fma(-x, r, 1.0);
    const double p = fma(e, e, e);       // e + e^2
    r = fma(r, p, r);                    // ~ 2^-60
    return r;
}

__global__ void lat(double *out, long long *clk, double x0) {
    const int lane = threadIdx.x & 31;
    double x = x0 + lane * 1e-3;
    long long t[16];
    __shared__ double sh[16][64];
    double *my = sh[threadIdx.x >> 5];
    my[lane] = x;
    __syncwarp();
    int i = 0;
    t[0] = clock64();
#pragma unroll 1
    for (i = 0; i < 32; i++) { x = fma(x, 1.000001, 1e-9); x = fma(x, 1.000001, 1e-9); x = fma(x, 1.000001, 1e-9); x = fma(x, 1.000001, 1e-9); x = fma(x, 1.000001, 1e-9); x = fma(x, 1.000001, 1e-9); x = fma(x, 1.000001, 1e-9); x = fma(x, 1.000001, 1e-9); }
    t[1] = clock64();
#pragma unroll 1
    for (i = 0; i < 32; i++) { x = 1.0 / x; x = 1.0 / x; x = 1.0 / x; x = 1.0 / x; x = 1.0 / x; x = 1.0 / x; x = 1.0 / x; x = 1.0 / x; }
    t[2] = clock64();
#pragma unroll 1
    for (i = 0; i < 32; i++) { x = __drcp_rn(x); x = __drcp_rn(x); x = __drcp_rn(x); x = __drcp_rn(x); x = __drcp_rn(x); x = __drcp_rn(x); x = __drcp_rn(x); x = __drcp_rn(x); }
    t[3] = clock64();
#pragma unroll 1
    for (i = 0; i < 32; i++) { x = rcp_seed(x); x = rcp_seed(x); x = rcp_seed(x); x = rcp_seed(x); x = rcp_seed(x); x = rcp_seed(x); x = rcp_seed(x); x = rcp_seed(x); }
    t[4] = clock64();
#pragma unroll 1
    for (i = 0; i < 32; i++) { x = rcp_nr2(x); x = rcp_nr2(x); x = rcp_nr2(x); x = rcp_nr2(x); x = rcp_nr2(x); x = rcp_nr2(x); x = rcp_nr2(x); x = rcp_nr2(x); }
    t[5] = clock64();
#pragma unroll 1
    for (i = 0; i < 32; i++) { x = rcp_nr2b(x); x = rcp_nr2b(x); x = rcp_nr2b(x); x = rcp_nr2b(x); x = rcp_nr2b(x); x = rcp_nr2b(x); x = rcp_nr2b(x); x = rcp_nr2b(x); }
    t[6] = clock64();
#pragma unroll 1
    for (i = 0; i < 32; i++) { x = rsqrt(x); x = rsqrt(x); x = rsqrt(x); x = rsqrt(x); x = rsqrt(x); x = rsqrt(x); x = rsqrt(x); x = rsqrt(x); }
    t[7] = clock64();
#pragma unroll 1
    for (i = 0; i < 32; i++) { x = __shfl_sync(0xffffffffu, x, (lane + 1) & 31); x = __shfl_sync(0xffffffffu, x, (lane + 1) & 31); x = __shfl_sync(0xffffffffu, x, (lane + 1) & 31); x = __shfl_sync(0xffffffffu, x, (lane + 1) & 31); x = __shfl_sync(0xffffffffu, x, (lane + 1) & 31); x = __shfl_sync(0xffffffffu, x, (lane + 1) & 31); x = __shfl_sync(0xffffffffu, x, (lane + 1) & 31); x = __shfl_sync(0xffffffffu, x, (lane + 1) & 31); }
    t[8] = clock64();
#pragma unroll 1
    for (i = 0; i < 32; i++) {
#pragma unroll
        for (int q = 0; q < 8; q++) {
            my[lane] = x;
            __syncwarp();
            x = my[(lane + 1) & 31];
            __syncwarp();
        }
    }
    t[9] = clock64();
#pragma unroll 1
    for (i = 0; i < 32; i++) { x = sqrt(x); x = sqrt(x); x = sqrt(x); x = sqrt(x); x = sqrt(x); x = sqrt(x); x = sqrt(x); x = sqrt(x); }
    t[10] = clock64();
    // accuracy of the hand-rolled reciprocals against 1.0 / x
    double worst = 0, worst_b = 0;
    for (i = 0; i < 64; i++) {
        const double v = 1.0 + 0.37 * i + lane * 0.011;
        const double ex = 1.0 / v;
        worst = fmax(worst, fabs(rcp_nr2(v) - ex) / ex);
        worst_b = fmax(worst_b, fabs(rcp_nr2b(v) - ex) / ex);
    }
    out[lane] = x;
    out[32 + lane] = worst;
    out[64 + lane] = worst_b;
    if (threadIdx.x == 0)
        for (i = 0; i < 11; i++) clk[i] = t[i];
}

// DFMA issue rate: W warps of one CTA, each lane 8 independent chains
__global__ void thr(double *out, long long *clk) {
    double a0 = threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
    const double y = 1.000001;
    __syncthreads();
    const long long t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < 512; i++) {
        a0 = fma(a0, y, 1e-9); a1 = fma(a1, y, 1e-9); a2 = fma(a2, y, 1e-9); a3 = fma(a3, y, 1e-9);
        a4 = fma(a4, y, 1e-9); a5 = fma(a5, y, 1e-9); a6 = fma(a6, y, 1e-9); a7 = fma(a7, y, 1e-9);
    }
    __syncthreads();
    const long long t1 = clock64();
    out[threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
    if (threadIdx.x == 0) clk[0] = t1 - t0;
}

// named-barrier round trip between two warps, and a shared-memory flag hand-over (producer warp -> consumer warp)
__global__ void sync_probe(long long *clk) {
    __shared__ volatile int flag;
    __shared__ volatile double box[32];
    const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) flag = 0;
    __syncthreads();
    long long t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < 256; i++) asm volatile("bar.sync 1, 64;" ::: "memory");
    long long t1 = clock64();
    __syncthreads();
    // ping-pong: warp 0 publishes i, warp 1 answers -i
    long long t2 = clock64();
#pragma unroll 1
    for (int i = 1; i <= 256; i++) {
        if (wid == 0) {
            box[lane] = i;
            __threadfence_block();
            __syncwarp();
            if (lane == 0) flag = i;
            while (flag != -i) {}
        } else {
            while (flag != i) {}
            const double v = box[lane];
            __syncwarp();
            if (lane == 0 && v == i) flag = -i;
            __syncwarp();
        }
    }
    long long t3 = clock64();
    __syncthreads();
    long long t4 = clock64();
#pragma unroll 1
    for (int i = 0; i < 256; i++) __syncwarp();
    long long t5 = clock64();
    if (threadIdx.x == 0) {
        clk[0] = t1 - t0;
        clk[1] = t3 - t2;
        clk[2] = t5 - t4;
    }
}

int main() {
    double *o, ho[96];
    long long *c, h[16];
    cudaMalloc(&o, 2048 * 8);
    cudaMalloc(&c, 16 * 8);
    for (int rep = 0; rep < 3; rep++) {
        lat<<<1, 32>>>(o, c, 1.5);
        cudaDeviceSynchronize();
    }
    cudaMemcpy(h, c, 16 * 8, cudaMemcpyDeviceToHost);
    cudaMemcpy(ho, o, 96 * 8, cudaMemcpyDeviceToHost);
    const char *n[10] = {"dfma dependent", "1.0 / x", "__drcp_rn", "rcp.approx.ftz.f64 (seed only)", "seed + 2 Newton (4 dfma)", "seed + (e + e^2) (3 dfma)",
                         "rsqrt(double)", "shfl", "sts + syncwarp + lds + syncwarp", "sqrt(double)"};
    for (int i = 0; i < 10; i++) printf("%-40s %7.1f cycles\n", n[i], (h[i + 1] - h[i]) / 256.0);
    double w = 0, wb = 0;
    for (int i = 0; i < 32; i++) {
        w = w > ho[32 + i] ? w : ho[32 + i];
        wb = wb > ho[64 + i] ? wb : ho[64 + i];
    }
    printf("max rel. error: seed + 2 Newton %.3g, seed + (e + e^2) %.3g\n", w, wb);
    for (int warps : {1, 2, 4, 8, 16, 32}) {
        for (int rep = 0; rep < 2; rep++) {
            thr<<<1, 32 * warps>>>(o, c);
            cudaDeviceSynchronize();
        }
        cudaMemcpy(h, c, 8, cudaMemcpyDeviceToHost);
        printf("dfma throughput, %2d warps: %.2f warp-instructions / cycle / SM (%.1f cycles per warp instruction per warp)\n", warps,
               warps * 512.0 * 8 / h[0], h[0] / (512.0 * 8));
    }
    for (int rep = 0; rep < 2; rep++) {
        sync_probe<<<1, 64>>>(c);
        cudaDeviceSynchronize();
    }
    cudaMemcpy(h, c, 3 * 8, cudaMemcpyDeviceToHost);
    printf("bar.sync (2 warps) %.1f cycles; shared flag ping-pong round trip %.1f cycles; __syncwarp %.1f cycles\n", h[0] / 256.0, h[1] / 256.0, h[2] / 256.0);
    return 0;
}
