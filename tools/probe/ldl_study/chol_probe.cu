// Stand-alone check + timing of the reduced-system factorisations (chol_blocked.cuh: blocked look-ahead Cholesky; chol_ldl.cuh:
// LDL^T with the inverse ride-along) on a matrix with the sliding window's structure.  One CTA, 512 threads, SM clocks.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a --expt-relaxed-constexpr -I../../../lf-vio_b200/csrc -I. -o chol_probe chol_probe.cu
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "chol_ldl.cuh"

using namespace lfvio;

__global__ void __launch_bounds__(512) k_old(const double *Sg, const double *bg, double *yg, int D, int F, long long *clk, int *okout) {
    extern __shared__ double smem[];
    __shared__ __align__(16) double s_U[2 * NB * NB];
    __shared__ __align__(16) double s_row[2 * NB];
    __shared__ int s_okv[64];
    const int PW = chol_panel_pitch(D), nblk = (D + NB - 1) / NB;
    double *s_panel = smem, *s_y = s_panel + (size_t)NB * PW, *s_invd = s_y + ((D + 1) & ~1), *s_S = s_invd + nblk * NB;
    const MatPacked Sp{s_S, D};
    for (int q = threadIdx.x; q < D * D; q += blockDim.x) {
        const int r = q / D, c = q % D;
        if (c >= r) Sp.at(r, c) = Sg[q];
    }
    for (int q = threadIdx.x; q < NB * PW; q += blockDim.x) s_panel[q] = 0.0;
    for (int c = threadIdx.x; c < D; c += blockDim.x) s_y[c] = bg[c];
    __syncthreads();
    const long long t0 = clock64();
    const bool pd = F > 0 ? cholesky_lookahead(Sp, D, s_panel, PW, s_U, s_invd, s_y, s_okv, s_row, WindowProfile{F})
                          : cholesky_lookahead(Sp, D, s_panel, PW, s_U, s_invd, s_y, s_okv, s_row);
    __syncthreads();
    const long long t1 = clock64();
    if (pd) solve_upper_back(Sp, D, s_y, s_invd);
    __syncthreads();
    const long long t2 = clock64();
    for (int c = threadIdx.x; c < D; c += blockDim.x) yg[c] = s_y[c];
    if (threadIdx.x == 0) {
        clk[0] = t1 - t0;
        clk[1] = t2 - t1;
        *okout = pd;
    }
}

__global__ void __launch_bounds__(512) k_new(const double *Sg, const double *bg, double *yg, int D, int F, long long *clk, int *okout) {
    extern __shared__ __align__(16) double smem[];
    const int PW = chol_panel_pitch(D);
    double *s_y = smem + ((ldl_smem_doubles(D) + 1) & ~(size_t)1), *s_S = s_y + ((D + 1) & ~1);
    const LdlSmem sm = ldl_carve(smem, D, s_y);
    const MatPacked Sp{s_S, D};
    const SmPacked Sq{smem_u32(s_S), D};
    for (int q = threadIdx.x; q < D * D; q += blockDim.x) {
        const int r = q / D, c = q % D;
        if (c >= r) Sp.at(r, c) = Sg[q];
    }
    for (int c = threadIdx.x; c < D; c += blockDim.x) s_y[c] = bg[c];
    __syncthreads();
    const long long t0 = clock64();
    const bool pd = F > 0 ? ldl_lookahead(Sq, D, sm, PW, WindowProfile{F}) : ldl_lookahead(Sq, D, sm, PW);
    const long long t1 = clock64();
    if (pd) ldl_solve_back(Sq, D, sm);
    __syncthreads();
    const long long t2 = clock64();
    for (int c = threadIdx.x; c < D; c += blockDim.x) yg[c] = s_y[c];
    if (threadIdx.x == 0) {
        clk[0] = t1 - t0;
        clk[1] = t2 - t1;
        *okout = pd;
    }
}

void run_isolated(const double *dS, int D, long long *dclk);
void run_contend(const double *dS, int D, long long *dclk);
static double urand() { return rand() / (double)RAND_MAX - 0.5; }

int main(int argc, char **argv) {
    const int F = argc > 1 ? atoi(argv[1]) : 11;
    const int structured = argc > 2 ? atoi(argv[2]) : 1;
    const int D = 15 * F + 7;
    srand(7);
    // column order [sb_{F-1} .. sb_1 | sb_0 | pose_{F-1} .. pose_0 | ex | td]
    auto col_sb = [&](int k) { return 9 * (F - 1 - k); };
    auto col_pose = [&](int k) { return 9 * F + 6 * (F - 1 - k); };
    std::vector<double> H((size_t)D * D, 0.0), b(D);
    auto add_factor = [&](const std::vector<int> &cols, int rows) {
        std::vector<double> J((size_t)rows * cols.size());
        for (auto &v : J) v = urand();
        for (size_t a = 0; a < cols.size(); a++)
            for (size_t c = 0; c < cols.size(); c++) {
                double s = 0;
                for (int r = 0; r < rows; r++) s += J[r * cols.size() + a] * J[r * cols.size() + c];
                H[(size_t)cols[a] * D + cols[c]] += s;
            }
    };
    if (structured) {
        for (int k = 0; k + 1 < F; k++) {  // IMU factor between frames k and k+1
            std::vector<int> cols;
            for (int a = 0; a < 6; a++) cols.push_back(col_pose(k) + a);
            for (int a = 0; a < 9; a++) cols.push_back(col_sb(k) + a);
            for (int a = 0; a < 6; a++) cols.push_back(col_pose(k + 1) + a);
            for (int a = 0; a < 9; a++) cols.push_back(col_sb(k + 1) + a);
            add_factor(cols, 15);
        }
        std::vector<int> dense;  // prior + landmarks' Schur complement: sb_0, all poses, ex, td
        for (int c = 9 * (F - 1); c < D; c++) dense.push_back(c);
        add_factor(dense, D);
    } else {
        std::vector<int> all;
        for (int c = 0; c < D; c++) all.push_back(c);
        add_factor(all, D + 8);
    }
    for (int c = 0; c < D; c++) {
        H[(size_t)c * D + c] += 0.5;
        b[c] = urand();
    }
    // host reference: long double Cholesky
    std::vector<long double> L((size_t)D * D, 0.0L), yref(D);
    for (int j = 0; j < D; j++) {
        long double s = H[(size_t)j * D + j];
        for (int k = 0; k < j; k++) s -= L[(size_t)j * D + k] * L[(size_t)j * D + k];
        L[(size_t)j * D + j] = sqrtl(s);
        for (int i = j + 1; i < D; i++) {
            long double t = H[(size_t)i * D + j];
            for (int k = 0; k < j; k++) t -= L[(size_t)i * D + k] * L[(size_t)j * D + k];
            L[(size_t)i * D + j] = t / L[(size_t)j * D + j];
        }
    }
    for (int i = 0; i < D; i++) {
        long double s = b[i];
        for (int k = 0; k < i; k++) s -= L[(size_t)i * D + k] * yref[k];
        yref[i] = s / L[(size_t)i * D + i];
    }
    for (int i = D - 1; i >= 0; i--) {
        long double s = yref[i];
        for (int k = i + 1; k < D; k++) s -= L[(size_t)k * D + i] * yref[k];
        yref[i] = s / L[(size_t)i * D + i];
    }
    double *dS, *db, *dy;
    long long *dclk, hclk[2];
    int *dok, hok;
    cudaMalloc(&dS, sizeof(double) * D * D);
    cudaMalloc(&db, sizeof(double) * D);
    cudaMalloc(&dy, sizeof(double) * D);
    cudaMalloc(&dclk, 16);
    cudaMalloc(&dok, 4);
    cudaMemcpy(dS, H.data(), sizeof(double) * D * D, cudaMemcpyHostToDevice);
    cudaMemcpy(db, b.data(), sizeof(double) * D, cudaMemcpyHostToDevice);
    const int PW = chol_panel_pitch(D), nblk = (D + NB - 1) / NB;
    const size_t packed = (size_t)D * (D + 1) / 2;
    const size_t smem_old = sizeof(double) * ((size_t)NB * PW + ((D + 1) & ~1) + nblk * NB + packed);
    const size_t smem_new = sizeof(double) * (((ldl_smem_doubles(D) + 1) & ~(size_t)1) + ((D + 1) & ~1) + packed);
    cudaFuncSetAttribute(k_old, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    cudaFuncSetAttribute(k_new, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (argc > 3) { run_isolated(dS, D, dclk); run_contend(dS, D, dclk); return 0; }
    std::vector<double> y(D);
    for (int which = 0; which < 2; which++) {
        const int Fp = structured ? F : 0;
        for (int rep = 0; rep < 3; rep++) {
            if (which == 0) k_old<<<1, 512, smem_old>>>(dS, db, dy, D, Fp, dclk, dok);
            else k_new<<<1, 512, smem_new>>>(dS, db, dy, D, Fp, dclk, dok);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) {
                printf("%s: %s\n", which ? "ldl" : "chol", cudaGetErrorString(e));
                return 1;
            }
        }
        cudaMemcpy(hclk, dclk, 16, cudaMemcpyDeviceToHost);
        cudaMemcpy(&hok, dok, 4, cudaMemcpyDeviceToHost);
        cudaMemcpy(y.data(), dy, sizeof(double) * D, cudaMemcpyDeviceToHost);
        double err = 0, nrm = 0;
        for (int i = 0; i < D; i++) {
            err = fmax(err, fabs(y[i] - (double)yref[i]));
            nrm = fmax(nrm, fabs((double)yref[i]));
        }
#ifdef LDL_PROFILE
        if (which == 1) {
            static long long ev[32][32][8];
            cudaMemcpyFromSymbol(ev, g_ldl_ev, sizeof(ev));
            const long long *e = ev[0][0], *f = ev[5][0];
            printf("  warp 0 totals: factor %lld | wait CTA %lld | P1 %lld | wait %lld | D1 %lld | wait %lld | loop top %lld\n", e[0], e[1], e[2], e[3], e[4], e[5], e[7]);
            printf("  warp 5 totals: wait CTA %lld | panel %lld | wait %lld | write-back + trailing %lld | loop top %lld\n", f[1] + f[0], f[2], f[3], f[4], f[7]);
        }
#endif
        printf("%-28s F=%d D=%d %s: ok=%d factor %lld cycles, back-substitution %lld cycles, max |y - y_ref| / max |y_ref| = %.3g (smem %zu B)\n",
               which ? "ldl_lookahead (chol_ldl.cuh)" : "cholesky_lookahead (old)", F, D, structured ? "window profile" : "dense", hok, hclk[0], hclk[1],
               err / nrm, which ? smem_new : smem_old);
    }
    return 0;
}

// ---- isolated pieces (run with a third argument) ------------------------------------------------------------------
__global__ void k_factor_only(const double *Sg, int D, long long *clk, int reps) {
    __shared__ __align__(16) double sD[256], sL[256], sRd[16], sRow[32];
    for (int q = threadIdx.x; q < 256; q += blockDim.x) {
        const int r = q / 16, c = q % 16;
        sD[q] = c >= r ? Sg[(size_t)r * D + c] : 0.0;
    }
    __syncthreads();
    if (threadIdx.x < 32) {
        const long long t0 = clock64();
        bool ok = true;
        for (int i = 0; i < reps; i++) ok = ldl_factor_inv(smem_u32(sD), smem_u32(sL), smem_u32(sRd), smem_u32(sRow)) && ok;
        const long long t1 = clock64();
        if (threadIdx.x == 0) clk[0] = (t1 - t0) / reps + (ok ? 0 : 1000000);
    }
}
void run_isolated(const double *dS, int D, long long *dclk) {
    long long h;
    for (int nt : {32, 512}) {
        for (int r = 0; r < 2; r++) { k_factor_only<<<1, nt>>>(dS, D, dclk, 16); cudaDeviceSynchronize(); }
        cudaMemcpy(&h, dclk, 8, cudaMemcpyDeviceToHost);
        printf("ldl_factor_inv alone (%d threads in the CTA, warp 0 calls): %lld cycles per 16 x 16 block (%s)\n", nt, h, cudaGetErrorString(cudaGetLastError()));
    }
}

// ---- what slows the chain warp down?  Warp 0 factors; warps 1.. (nw of them) run a given kind of load next to it. ---------
template <int N>
__device__ __forceinline__ void long_code(unsigned &x) {  // N * 8 distinct straight-line integer instructions
    if constexpr (N > 0) {
        x = x * 1664525u + (unsigned)N;
        x ^= x >> 7;
        x += 0x9e3779b9u * (unsigned)N;
        x ^= x << 3;
        x = x * 22695477u + 1u;
        x ^= x >> 11;
        x += (unsigned)N * 77u;
        x ^= x << 5;
        long_code<N - 1>(x);
    }
}
__global__ void __launch_bounds__(512) k_contend(const double *Sg, int D, long long *clk, int mode, int nw, volatile int *stop_g) {
    __shared__ __align__(16) double sD[256], sL[256], sRd[16], sRow[32], sBuf[2048];
    __shared__ volatile int stop;
    for (int q = threadIdx.x; q < 256; q += blockDim.x) {
        const int r = q / 16, c = q % 16;
        sD[q] = c >= r ? Sg[(size_t)r * D + c] : 0.0;
    }
    for (int q = threadIdx.x; q < 2048; q += blockDim.x) sBuf[q] = 1.0 + q * 1e-6;
    if (threadIdx.x == 0) stop = 0;
    __syncthreads();
    const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (wid == 0) {
        const long long t0 = clock64();
        bool ok = true;
        for (int i = 0; i < 16; i++) ok = ldl_factor_inv(smem_u32(sD), smem_u32(sL), smem_u32(sRd), smem_u32(sRow)) && ok;
        const long long t1 = clock64();
        if (lane == 0) {
            clk[0] = (t1 - t0) / 16 + (ok ? 0 : 1000000);
            stop = 1;
        }
    } else if (wid <= nw && (mode < 6 || (wid & 3) != 0)) {
        double a0 = lane, a1 = 1, a2 = 2, a3 = 3, a4 = 4, a5 = 5, a6 = 6, a7 = 7;
        unsigned x = threadIdx.x;
        const uint32_t b = smem_u32(sBuf);
        while (!stop) {
            if (mode == 1 || mode == 6) {
#pragma unroll 4
                for (int i = 0; i < 32; i++) {
                    a0 = fma(a0, 1.0000001, 1e-9); a1 = fma(a1, 1.0000001, 1e-9); a2 = fma(a2, 1.0000001, 1e-9); a3 = fma(a3, 1.0000001, 1e-9);
                    a4 = fma(a4, 1.0000001, 1e-9); a5 = fma(a5, 1.0000001, 1e-9); a6 = fma(a6, 1.0000001, 1e-9); a7 = fma(a7, 1.0000001, 1e-9);
                }
            } else if (mode == 2 || mode == 7) {
#pragma unroll 8
                for (int i = 0; i < 64; i++) { const double2 v = lds2(b + 16u * ((i + wid) & 127)); a0 += v.x; a1 += v.y; }
            } else if (mode == 3 || mode == 8) {
#pragma unroll 8
                for (int i = 0; i < 64; i++) a0 += lds(b + 8u * (((i * 32 + lane) & 2047)));
            } else if (mode == 4 || mode == 9) {
#pragma unroll 8
                for (int i = 0; i < 64; i++) sts(b + 8u * (((i * 32 + lane + 64 * wid) & 2047)), a0);
            } else if (mode == 5) {
                long_code<100>(x); long_code<99>(x); long_code<98>(x); long_code<97>(x);
            }
        }
        if (a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7 + x == 12345.678) clk[1] = 1;
    }
}
void run_contend(const double *dS, int D, long long *dclk) {
    long long h;
    const char *names[10] = {"nothing", "DFMA (8 chains / lane)", "LDS.128 broadcast", "LDS.64 distinct", "STS.64", "long straight-line integer code (3200 instr)", "DFMA, not on warp 0's scheduler (warps 4, 8, 12 idle)", "LDS.128 broadcast, warps 4, 8, 12 idle", "LDS.64 distinct, warps 4, 8, 12 idle", "STS.64, warps 4, 8, 12 idle"};
    for (int mode = 0; mode < 10; mode++)
        for (int nw : {3, 15}) {
            if (mode == 0 && nw == 15) continue;
            for (int r = 0; r < 2; r++) { k_contend<<<1, 512>>>(dS, D, dclk, mode, nw, nullptr); cudaDeviceSynchronize(); }
            cudaMemcpy(&h, dclk, 8, cudaMemcpyDeviceToHost);
            printf("factor next to %2d warps of %-55s: %6lld cycles per block (%s)\n", nw, names[mode], h, cudaGetErrorString(cudaGetLastError()));
        }
}
