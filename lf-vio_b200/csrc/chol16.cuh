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

// ---- shared memory by 32-bit address ---------------------------------------------------------------------------
// A generic pointer into shared memory costs an S2R + LEA per access to rebuild the shared window base (nvcc 12.9, sm_100:
// 600 such pairs in this library); on the factorisation's dependent chain every instruction counts, so the chain code
// converts once and uses ld/st.shared with 32-bit addresses.
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ double lds(uint32_t a) {
    double v;
    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(a));
    return v;
}
__device__ __forceinline__ double2 lds2(uint32_t a) {  // 16-byte aligned
    double2 v;
    asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(a));
    return v;
}
__device__ __forceinline__ void sts(uint32_t a, double v) { asm volatile("st.shared.f64 [%0], %1;" ::"r"(a), "d"(v) : "memory"); }

// Variant on the solver's critical path since round 2, built on what tools/probe/prim_probe.cu measured on a B200 (dependent
// DFMA 9.4 cycles, MUFU.RCP64H + 3 DFMA 40, shfl of a double 26 and two SHFL, shared-memory exchange 35, rsqrt(double) 68,
// one warp issues an fp64 instruction every 4.5 cycles; fp64 issue is per scheduler).
// The block is factored as Uu^T diag(d) Uu — no square root, no normalised row on the dependent chain: lanes 0-15 hold the
// columns of the block, lanes 16-31 the columns of the identity, all 16 rows in registers, and ONE elimination loop factors
// the block and inverts its unit factor (the same FMAs run on both halves).  Row j is broadcast through shared memory while
// 1 / d_j is computed from a shuffled pivot (seed + one cubic step: three dependent DFMA), and row j + 1 — what the next
// pivot waits for — is updated first, as  a -= (a_j,j+1 a_j,c) (1 / d_j)  with the product formed ahead; the bulk of pivot j
// is issued inside the shuffle / reciprocal latency of pivot j + 1 (software-pipelined by hand: ptxas keeps volatile shared
// accesses in order).  Chain per pivot: DFMA + shfl + reciprocal, 113 cycles measured against 181 for the rsqrt form.
// Out: s_U = Cholesky factor U (upper, zeros below), s_invd[j] = 1 / U_jj, s_PL = U^-T (lower triangular) — with it the
// panel rows U^-T A12 are independent dot products instead of a 16-step substitution.  s_row: 2 x 16 doubles.  All 32 lanes call.
static __device__ __noinline__ bool factor_diag_block_ldl(double *g_U, double *g_invd, double *g_PL, double *g_row) {
    const uint32_t s_U = smem_u32(g_U), s_invd = smem_u32(g_invd), s_PL = smem_u32(g_PL), s_row = smem_u32(g_row);
    const int lane = threadIdx.x & 31, c = lane & 15;
    const bool is_a = lane < 16;
    double a[NB];
#pragma unroll
    for (int r = 0; r < NB; r++) {
        const double v = lds(s_U + 8u * (r * NB + c));  // (all lanes load; the identity half overrides)
        a[r] = is_a ? v : ((r == c) ? 1.0 : 0.0);
    }
    bool good = true;
    double xp[NB], tp = 0.0;  // raw row and scaled own entry of the previous pivot (bulk still to apply)
#pragma unroll
    for (int r = 0; r < NB; r++) xp[r] = 0.0;
#pragma unroll
    for (int j = 0; j < NB; j++) {
        const uint32_t row = s_row + 8u * NB * (j & 1);
        const double d = __shfl_sync(0xffffffffu, a[j], j);
        if (is_a) sts(row + 8u * c, a[j]);
        __syncwarp();
        double xr[NB];  // raw row j: xr[r] = A[j][r], r > j
        if (j + 1 < NB) {
#pragma unroll
            for (int r = (j + 1) & ~1; r < NB; r += 2) {
                const double2 v = lds2(row + 8u * r);
                xr[r] = v.x;
                xr[r + 1] = v.y;
            }
        }
        // bulk of pivot j - 1 in four parts around the reciprocal's dependent steps
        constexpr int K0 = 5;
#pragma unroll
        for (int r = j + 1; r < NB; r++)
            if (j >= 1 && r - (j + 1) < K0) a[r] = fma(-xp[r], tp, a[r]);
        double r0;
        asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r0) : "d"(d));
#pragma unroll
        for (int r = j + 1; r < NB; r++)
            if (j >= 1 && r - (j + 1) >= K0 && r - (j + 1) < K0 + 3) a[r] = fma(-xp[r], tp, a[r]);
        const double e = fma(-d, r0, 1.0);
#pragma unroll
        for (int r = j + 1; r < NB; r++)
            if (j >= 1 && r - (j + 1) >= K0 + 3 && r - (j + 1) < K0 + 5) a[r] = fma(-xp[r], tp, a[r]);
        const double pe = fma(e, e, e);
#pragma unroll
        for (int r = j + 1; r < NB; r++)
            if (j >= 1 && r - (j + 1) >= K0 + 5) a[r] = fma(-xp[r], tp, a[r]);
        const double rd = fma(r0, pe, r0);  // 1 / d: agrees with 1.0 / d to the last bit on every value of the probe
        good = good && (d > 0.0) && (d < 1e300);  // false for NaN too
        if (lane == j) sts(s_invd + 8u * j, rd);
        if (j + 1 < NB) {
            const double qn = xr[j + 1] * a[j];  // formed while the reciprocal is in flight
            a[j + 1] = fma(-qn, rd, a[j + 1]);   // row j + 1 first: the next pivot waits for it
            tp = a[j] * rd;                      // Uu[j][c] (lanes 0-15), Linv[j][c] / d_j (lanes 16-31)
#pragma unroll
            for (int r = j + 2; r < NB; r++) xp[r] = xr[r];
        }
    }
    __syncwarp();
    if (is_a) sts(s_invd + 8u * c, sqrt(lds(s_invd + 8u * c)));  // 1 / U_jj = sqrt(1 / d_j), one square root per lane, off the chain
    __syncwarp();
    // U = diag(sqrt(1/d)) (raw rows), U^-T = diag(sqrt(1/d)) Linv: row r of either is scaled by 1 / U_rr
    bool finite = true;
#pragma unroll
    for (int r = 0; r < NB; r += 2) {
        const double2 sq = lds2(s_invd + 8u * r);
        const double v0 = a[r] * sq.x, v1 = a[r + 1] * sq.y;
        finite = finite && (fabs(sq.x) < 1e150) && (fabs(sq.y) < 1e150);
        if (is_a) {
            sts(s_U + 8u * (r * NB + c), (r <= c) ? v0 : 0.0);
            sts(s_U + 8u * ((r + 1) * NB + c), (r + 1 <= c) ? v1 : 0.0);
        } else {
            sts(s_PL + 8u * (r * NB + c), (r >= c) ? v0 : 0.0);
            sts(s_PL + 8u * ((r + 1) * NB + c), (r + 1 >= c) ? v1 : 0.0);
        }
    }
    __syncwarp();
    return __all_sync(0xffffffffu, good && finite);
}

#endif

}  // namespace lfvio
