// 16 x 16 warp-level Cholesky + inverse shared by the solver (diagonal blocks of the reduced system) and the
// marginalisation (inverse of the dropped pose / speed-bias block).
#pragma once
#include "common.cuh"

namespace lfvio {

constexpr int NB = 16;  // Cholesky panel width

#ifdef __CUDACC__
// Warp-level Cholesky of one 16 x 16 diagonal block held in shared memory (identity-padded), U^T U = A, in registers:
// lane c owns column c; pivots and row entries are exchanged with warp shuffles; no fp64 division or square root
// (rsqrt + multiplies).  Overwrites s_U with U (upper, zeros below) and writes U^-1 (upper) into s_X.  All 32 lanes call.
static __device__ __noinline__ bool factor_diag_block(double *s_U, double *s_X, double *pivot_min = nullptr, double *pivot_max = nullptr) {
    const int c = threadIdx.x & 15;  // lanes 16..31 mirror lanes 0..15 (keeps the shuffles convergent and cheap)
    double a[NB];
#pragma unroll
    for (int r = 0; r < NB; r++) a[r] = s_U[r * NB + c];
    bool ok = true;
    double pmin = 1e300, pmax = 0.0;
    double invd[NB];
#pragma unroll
    for (int j = 0; j < NB; j++) {
        double djj = __shfl_sync(0xffffffffu, a[j], j);
        if (!(djj > 0.0) || !(djj < 1e300)) {
            ok = false;
            djj = 1.0;
        }
        pmin = fmin(pmin, djj);
        pmax = fmax(pmax, djj);
        const double inv = rsqrt(djj);
        invd[j] = inv;
        // U[j][c] = a[j] / sqrt(djj) (lane j holds a[j] == djj).  Entries below the diagonal of a column (r > c) are
        // never read by another lane, so they are updated unconditionally (no selects) and zeroed at write-out.
        const double ujc = a[j] * inv;
        a[j] = ujc;
#pragma unroll
        for (int r = j + 1; r < NB; r++) {
            const double ujr = __shfl_sync(0xffffffffu, ujc, r);
            a[r] = fma(-ujr, ujc, a[r]);
        }
    }
    __syncwarp();
    if (threadIdx.x < NB) {
#pragma unroll
        for (int r = 0; r < NB; r++) s_U[r * NB + c] = (r <= c) ? a[r] : 0.0;
    }
    __syncwarp();
    // inverse, column c: x[c] = 1/U[c][c]; x[r] = -(sum_{k=r+1..c} U[r][k] x[k]) / U[r][r]
    if (threadIdx.x < NB) {
        double x[NB];
#pragma unroll
        for (int r = NB - 1; r >= 0; r--) {
            double s = 0;
#pragma unroll
            for (int k = r + 1; k < NB; k++) s = fma(s_U[r * NB + k], x[k], s);  // x[k] == 0 for k > c
            x[r] = (r > c) ? 0.0 : ((r == c) ? invd[r] : -s * invd[r]);
        }
#pragma unroll
        for (int r = 0; r < NB; r++) s_X[r * NB + c] = x[r];
    }
    __syncwarp();
    if (pivot_min) *pivot_min = pmin;
    if (pivot_max) *pivot_max = pmax;
    return ok;
}

#endif

}  // namespace lfvio
