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

// Same factorisation without the explicit inverse: overwrites s_U with U (upper, zeros below) and writes the reciprocal
// pivots 1/U[j][j] to s_invd[0..15].  Used by the solver's look-ahead Cholesky, whose panel and triangular solves
// substitute with U and the reciprocal pivots directly.  All 32 lanes call.
static __device__ __noinline__ bool factor_diag_block_noinv(double *s_U, double *s_invd) {
    const int c = threadIdx.x & 15;
    double a[NB];
#pragma unroll
    for (int r = 0; r < NB; r++) a[r] = s_U[r * NB + c];
    // No per-pivot test on the dependent chain: a non-positive or non-finite pivot turns its reciprocal root into
    // NaN / inf / 0, which is detected once at the end (the caller discards the factor in that case).
    double inv_min = 1e300, inv_max = 0.0;
#pragma unroll
    for (int j = 0; j < NB; j++) {
        const double djj = __shfl_sync(0xffffffffu, a[j], j);
        const double inv = rsqrt(djj);
        inv_min = fmin(inv_min, inv);  // fmin/fmax drop NaN operands, so NaN is tested separately below
        inv_max = fmax(inv_max, inv);
        if (threadIdx.x == j) s_invd[j] = inv;
        const double ujc = a[j] * inv;
        a[j] = ujc;
#pragma unroll
        for (int r = j + 1; r < NB; r++) {
            const double ujr = __shfl_sync(0xffffffffu, ujc, r);
            a[r] = fma(-ujr, ujc, a[r]);
        }
    }
    __syncwarp();
    bool finite = true;
    if (threadIdx.x < NB) {
#pragma unroll
        for (int r = 0; r < NB; r++) {
            const double v = (r <= c) ? a[r] : 0.0;
            finite = finite && (fabs(v) < 1e300);  // false for NaN and inf
            s_U[r * NB + c] = v;
        }
    }
    finite = __all_sync(0xffffffffu, finite);
    __syncwarp();
    return finite && inv_min > 0.0 && inv_max < 1e150;
}

// Variant used on the solver's critical path.  Lane (h, c) = (lane / 16, lane % 16) holds rows 8h .. 8h+7 of column
// c, so a pivot step costs every lane at most 8 fp64 FMAs (one warp issues an fp64 FMA only every ~4.3 cycles on
// sm_100, measured), and row j of U is exchanged through a 16-double shared row (one store + four broadcast 128-bit
// loads per lane) instead of 30 shuffles.  s_row: 2 x 16 doubles, 16-byte aligned.  All 32 lanes call.
static __device__ __noinline__ bool factor_diag_block_split(double *s_U, double *s_invd, double *s_row) {
    const int lane = threadIdx.x & 31, c = lane & 15, h = lane >> 4, r0 = 8 * h;
    double a[8];
#pragma unroll
    for (int u = 0; u < 8; u++) a[u] = s_U[(r0 + u) * NB + c];
    double inv_min = 1e300, inv_max = 0.0;
#pragma unroll
    for (int j = 0; j < NB; j++) {
        const int hj = j >> 3, uj = j & 7;
        double *row = s_row + (j & 1) * NB;
        const double djj = __shfl_sync(0xffffffffu, a[uj], hj * 16 + j);
        const double inv = rsqrt(djj);
        inv_min = fmin(inv_min, inv);
        inv_max = fmax(inv_max, inv);
        if (lane == j) s_invd[j] = inv;
        if (h == hj) {
            const double ujc = a[uj] * inv;
            a[uj] = ujc;
            row[c] = ujc;
        }
        __syncwarp();
        const double my = row[c];
        const double2 *rr = reinterpret_cast<const double2 *>(row + r0);
        const double2 q0 = rr[0], q1 = rr[1], q2 = rr[2], q3 = rr[3];
        const double ur[8] = {q0.x, q0.y, q1.x, q1.y, q2.x, q2.y, q3.x, q3.y};
#pragma unroll
        for (int u = 0; u < 8; u++)
            if (r0 + u > j) a[u] = fma(-ur[u], my, a[u]);
    }
    __syncwarp();
    bool finite = true;
#pragma unroll
    for (int u = 0; u < 8; u++) {
        const double v = (r0 + u <= c) ? a[u] : 0.0;
        finite = finite && (fabs(v) < 1e300);
        s_U[(r0 + u) * NB + c] = v;
    }
    finite = __all_sync(0xffffffffu, finite);
    __syncwarp();
    return finite && inv_min > 0.0 && inv_max < 1e150;
}

#endif

}  // namespace lfvio
