// Times the 16x16 warp-level Cholesky variants of csrc/chol16.cuh on one warp and checks them against each other.
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include "../../lf-vio_b200/csrc/chol16.cuh"
using namespace lfvio;
__global__ void k(const double *A, double *U0, double *U1, double *U2, long long *cyc) {
    __shared__ __align__(16) double s_U[256], s_X[256], s_invd[16], s_row[32];
    const int lane = threadIdx.x;
    for (int rep = 0; rep < 3; rep++) {
        for (int q = lane; q < 256; q += 32) s_U[q] = A[q];
        __syncwarp();
        long long t0 = clock64();
        bool ok = factor_diag_block(s_U, s_X);
        long long t1 = clock64();
        if (lane == 0) cyc[0] = t1 - t0 + (ok ? 0 : 1);
        for (int q = lane; q < 256; q += 32) U0[q] = s_U[q];
        __syncwarp();
        for (int q = lane; q < 256; q += 32) s_U[q] = A[q];
        __syncwarp();
        t0 = clock64();
        ok = factor_diag_block_noinv(s_U, s_invd);
        t1 = clock64();
        if (lane == 0) cyc[1] = t1 - t0 + (ok ? 0 : 1);
        for (int q = lane; q < 256; q += 32) U1[q] = s_U[q];
        __syncwarp();
        for (int q = lane; q < 256; q += 32) s_U[q] = A[q];
        __syncwarp();
        t0 = clock64();
        ok = factor_diag_block_split(s_U, s_invd, s_row);
        t1 = clock64();
        if (lane == 0) cyc[2] = t1 - t0 + (ok ? 0 : 1);
        for (int q = lane; q < 256; q += 32) U2[q] = s_U[q];
        __syncwarp();
    }
}
int main() {
    double hA[256], hB[16 * 16];
    srand(1);
    for (int i = 0; i < 256; i++) hB[i] = rand() / (double)RAND_MAX - 0.5;
    for (int r = 0; r < 16; r++) for (int c = 0; c < 16; c++) {
        double s = (r == c) ? 0.5 : 0.0;
        for (int t = 0; t < 16; t++) s += hB[t * 16 + r] * hB[t * 16 + c];
        hA[r * 16 + c] = (c >= r) ? s : 0.0;
    }
    double *A, *U0, *U1, *U2; long long *cyc, hc[3];
    cudaMalloc(&A, 2048); cudaMalloc(&U0, 2048); cudaMalloc(&U1, 2048); cudaMalloc(&U2, 2048); cudaMalloc(&cyc, 24);
    cudaMemcpy(A, hA, 2048, cudaMemcpyHostToDevice);
    k<<<1, 32>>>(A, U0, U1, U2, cyc);
    double h0[256], h1[256], h2[256];
    cudaMemcpy(h0, U0, 2048, cudaMemcpyDeviceToHost); cudaMemcpy(h1, U1, 2048, cudaMemcpyDeviceToHost); cudaMemcpy(h2, U2, 2048, cudaMemcpyDeviceToHost);
    cudaMemcpy(hc, cyc, 24, cudaMemcpyDeviceToHost);
    double d1 = 0, d2 = 0, rec = 0;
    for (int i = 0; i < 256; i++) { d1 = fmax(d1, fabs(h0[i] - h1[i])); d2 = fmax(d2, fabs(h0[i] - h2[i])); }
    for (int r = 0; r < 16; r++) for (int c = r; c < 16; c++) {
        double s = 0; for (int t = 0; t < 16; t++) s += h2[t * 16 + r] * h2[t * 16 + c];
        rec = fmax(rec, fabs(s - hA[r * 16 + c]));
    }
    printf("%s | cycles: with-inverse %lld  noinv %lld  split %lld | max|U0-U1| %.3g  max|U0-U2| %.3g  max|U2^T U2 - A| %.3g\n",
           cudaGetErrorString(cudaDeviceSynchronize()), hc[0], hc[1], hc[2], d1, d2, rec);
    return 0;
}
