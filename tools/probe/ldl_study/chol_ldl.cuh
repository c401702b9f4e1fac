// Blocked LDL^T factorisation + solve of the reduced system, one CTA, built around what was measured on a B200
// (tools/probe/prim_probe.cu: dependent DFMA 9.4 cycles, MUFU.RCP64H + 3 DFMA 40, shfl 26 (two SHFL per double),
// shared-memory exchange 35, __syncwarp 23, rsqrt(double) 68; one warp issues an fp64 instruction every 4.5 cycles, the SM
// 1.5 per cycle; warp -> warp hand-over through a named barrier ~70) and on what the first versions of this file
// showed under ncu: the kernel is bound by INSTRUCTION ISSUE on the chain warp, not by arithmetic — a generic pointer
// into shared memory costs an S2R + LEA per access to rebuild the shared window base, the packed-triangle index five
// integer operations, a shuffled double two SHFL.  Hence: 32-bit shared addresses + ld/st.shared throughout, pointer
// walks instead of index arithmetic, and as few instructions per pivot as possible on warp 0.
//
// S = Uu^T diag(d) Uu (Uu unit upper).  Per 16-column block step:
//   warp 0   lanes 0-15 hold the columns of the 16 x 16 diagonal block, lanes 16-31 the columns of the identity, all 16
//            rows in registers.  One elimination loop factors the block AND produces Linv = (Uu^T)^-1 (the same FMAs
//            run on both halves): no square roots, no normalised rows; row j is broadcast through shared memory while
//            1 / d_j is computed from a shuffled pivot, and row j + 1 is updated first, as a -= (a_j,j+1 a_j,c) (1 / d_j)
//            with the product formed ahead: chain per pivot = DFMA + shfl + reciprocal (~82 cycles; the Cholesky form
//            with rsqrt and normalised rows measured 184).  Warps 4, 8 and 12 — the other warps of warp 0's scheduler —
//            stay out of the whole factorisation, so the chain never waits for an issue slot or the fp64 pipe.
//            With Linv, NOTHING downstream is a substitution: panel rows W = Linv A12 are independent dot products.
//   warps 0-3 then form the 16 look-ahead panel columns and the next diagonal block  A22 - W^T diag(1/d) W  (two short
//            128-thread phases) and warp 0 goes on factoring while
//   the workers (warps 1, 2, 3, 5, 6, 7, 9, ...) form the other panel columns (+ the right-hand side as one more column;
//            warp-uniform static row pairs (t, 15 - t): the triangular Linv costs 17 FMAs per pair) and apply the step to
//            the trailing matrix (skyline profile of chol_blocked.cuh), 16 x 16 tiles.
// The solve keeps Linv of every block: back substitution per block is Linv^T applied to the scaled right-hand side (one
// warp, two half dot products + one shuffle) followed by the update of the rows above.
#pragma once
#include "chol_blocked.cuh"

namespace lfvio {

#ifdef __CUDACC__

constexpr int LDL_BAR_CTA = 1;   // all taking part: block factored + previous trailing update done
constexpr int LDL_BAR_WORK = 2;  // workers: panel complete
constexpr int LDL_BAR_PRIO = 3;  // warps 0-3
constexpr int LDL_BAR_GO = 4;    // priority phase done: the workers start

#ifdef LDL_PROFILE
static __device__ long long g_ldl_ev[32][32][8];  // [warp][block][event] clock64 stamps
#endif

// ---- shared memory by 32-bit address ---------------------------------------------------------------------------
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
__device__ __forceinline__ int ldsi(uint32_t a) {
    int v;
    asm volatile("ld.shared.s32 %0, [%1];" : "=r"(v) : "r"(a));
    return v;
}
__device__ __forceinline__ void stsi(uint32_t a, int v) { asm volatile("st.shared.s32 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }

// ---- the two homes of S: packed upper triangle in shared memory, full square in global memory ---------------------
struct SmPacked {
    uint32_t base;
    int D;
    typedef uint32_t ptr;
    __device__ __forceinline__ ptr at(int r, int c) const { return base + 8u * (unsigned)(r * D - ((r * (r - 1)) >> 1) + (c - r)); }
    __device__ __forceinline__ ptr down(ptr p, int r) const { return p + 8u * (unsigned)(D - r - 1); }  // (r, c) -> (r + 1, c)
    __device__ __forceinline__ ptr row0(int r) const { return base + 8u * (unsigned)(r * D - ((r * (r - 1)) >> 1) - r); }  // right(row0(r), c) == at(r, c), c >= r
    static __device__ __forceinline__ ptr right(ptr p, int n) { return p + 8u * (unsigned)n; }
    static __device__ __forceinline__ double ld(ptr p) { return lds(p); }
    static __device__ __forceinline__ void st(ptr p, double v) { sts(p, v); }
};
struct GmFull {
    double *base;
    int D;
    typedef double *ptr;
    __device__ __forceinline__ ptr at(int r, int c) const { return base + (size_t)r * D + c; }
    __device__ __forceinline__ ptr down(ptr p, int r) const { return p + D; }
    __device__ __forceinline__ ptr row0(int r) const { return base + (size_t)r * D; }
    static __device__ __forceinline__ ptr right(ptr p, int n) { return p + n; }
    static __device__ __forceinline__ double ld(ptr p) { return *p; }
    static __device__ __forceinline__ void st(ptr p, double v) { *p = v; }
};

struct LdlSmem {      // 32-bit shared addresses
    uint32_t panel;   // NB x PW   W rows of the current step, compact trailing columns (+ rhs in the last column)
    uint32_t panelV;  // NB x PW   V = diag(1/d) W
    uint32_t Linv;    // nblk x 256
    uint32_t rd;      // nblk x 16  1 / d
    uint32_t Dg;      // 256        diagonal block to factor (upper, identity padded)
    uint32_t rowx;    // 2 x 16     raw pivot row (warp 0)
    uint32_t ok;      // nblk + 1 ints (last: the whole factorisation)
    uint32_t y;       // D          right-hand side / solution
};
__host__ __device__ inline size_t ldl_smem_doubles(int D) {  // without y
    const int nblk = (D + NB - 1) / NB;
    return 2 * (size_t)NB * chol_panel_pitch(D) + (size_t)nblk * (NB * NB + NB) + NB * NB + 2 * NB + (size_t)((nblk + 3) / 2);
}
__device__ inline LdlSmem ldl_carve(double *p, int D, double *s_y) {
    const int nblk = (D + NB - 1) / NB;
    LdlSmem s;
    uint32_t a = smem_u32(p);
    s.panel = a;
    a += 8u * NB * chol_panel_pitch(D);
    s.panelV = a;
    a += 8u * NB * chol_panel_pitch(D);
    s.Linv = a;
    a += 8u * nblk * NB * NB;
    s.rd = a;
    a += 8u * nblk * NB;
    s.Dg = a;
    a += 8u * NB * NB;
    s.rowx = a;
    a += 8u * 2 * NB;
    s.ok = a;
    s.y = smem_u32(s_y);
    return s;
}

// 1 / x: MUFU.RCP64H seed + one cubic step; agrees with 1.0 / x to the last bit on every value of the probe.
// Non-positive, non-finite or denormal pivots give inf / NaN, which the caller detects.
__device__ __forceinline__ double rcp_fast(double x) {
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    const double e = fma(-x, r, 1.0);
    const double p = fma(e, e, e);
    return fma(r, p, r);
}

// Warp 0: factor the block in s_D and invert its unit factor (see the header).  Writes Linv (row-major, lower) and 1 / d.
// Software-pipelined by hand: the bulk eliminations of pivot j - 1 (rows > j) are issued inside the shuffle / reciprocal
// latency of pivot j; only row j itself ("critical", updated first) is on the chain.
static __device__ __noinline__ bool ldl_factor_inv(uint32_t s_D, uint32_t s_Linv, uint32_t s_rd, uint32_t s_rowx) {
    const int lane = threadIdx.x & 31, c = lane & 15;
    const bool is_a = lane < 16;
    double a[NB];
#pragma unroll
    for (int r = 0; r < NB; r++) {
        const double v = lds(s_D + 8u * (r * NB + c));  // (all lanes load; the identity half overrides)
        a[r] = is_a ? v : ((r == c) ? 1.0 : 0.0);
    }
    bool good = true;
    double xp[NB], tp = 0.0;  // raw row and scaled own entry of the previous pivot (bulk still to apply)
#pragma unroll
    for (int r = 0; r < NB; r++) xp[r] = 0.0;
#pragma unroll
    for (int j = 0; j < NB; j++) {
        const uint32_t row = s_rowx + 8u * NB * (j & 1);
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
        // bulk of pivot j - 1, first part (hides the shuffle)
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
        const double rd = fma(r0, pe, r0);
        good = good && (d > 0.0) && (d < 1e300);  // false for NaN too
        if (lane == j) sts(s_rd + 8u * j, rd);
        if (j + 1 < NB) {
            const double qn = xr[j + 1] * a[j];  // formed while the reciprocal is in flight
            a[j + 1] = fma(-qn, rd, a[j + 1]);   // row j + 1 first: the next pivot waits for it
            tp = a[j] * rd;                      // Uu[j][c] (lanes 0-15), Linv[j][c] / d_j (lanes 16-31)
#pragma unroll
            for (int r = j + 2; r < NB; r++) xp[r] = xr[r];
        }
    }
    if (!is_a) {
#pragma unroll
        for (int r = 0; r < NB; r++) sts(s_Linv + 8u * (r * NB + c), a[r]);
    }
    return __all_sync(0xffffffffu, good);
}

// rows P and 15 - P of W = Linv a for one column held in registers; Linv is unit lower triangular: 17 FMAs for the pair
template <int P>
__device__ __forceinline__ void ldl_panel_pair(uint32_t Li, const double (&a)[NB], double &w1, double &w2) {
    constexpr int T1 = P, T2 = NB - 1 - P;
    double e1 = 0, o1 = 0, e2 = 0, o2 = 0;
#pragma unroll
    for (int r = 0; r <= T1; r += 2) {
        const double2 l = lds2(Li + 8u * (T1 * NB + r));
        e1 = fma(l.x, a[r], e1);
        if (r + 1 <= T1) o1 = fma(l.y, a[r + 1], o1);
    }
#pragma unroll
    for (int r = 0; r <= T2; r += 2) {
        const double2 l = lds2(Li + 8u * (T2 * NB + r));
        e2 = fma(l.x, a[r], e2);
        if (r + 1 <= T2) o2 = fma(l.y, a[r + 1], o2);
    }
    w1 = e1 + o1;
    w2 = e2 + o2;
}
// pairs P0 and P0 + 1 of one column into the W and V panels (all the arithmetic first, then the eight stores)
template <int P0>
__device__ __forceinline__ void ldl_panel_2pairs(uint32_t Li, uint32_t rd, const double (&a)[NB], uint32_t pcol, uint32_t vcol, int PW) {
    constexpr int TA = P0, TB = NB - 1 - P0, TC = P0 + 1, TD = NB - 2 - P0;
    const double ra = lds(rd + 8u * TA), rb = lds(rd + 8u * TB), rc = lds(rd + 8u * TC), rdd = lds(rd + 8u * TD);
    double wa, wb, wc, wd;
    ldl_panel_pair<P0>(Li, a, wa, wb);
    ldl_panel_pair<P0 + 1>(Li, a, wc, wd);
    sts(pcol + 8u * (unsigned)(TA * PW), wa);
    sts(pcol + 8u * (unsigned)(TB * PW), wb);
    sts(pcol + 8u * (unsigned)(TC * PW), wc);
    sts(pcol + 8u * (unsigned)(TD * PW), wd);
    sts(vcol + 8u * (unsigned)(TA * PW), wa * ra);
    sts(vcol + 8u * (unsigned)(TB * PW), wb * rb);
    sts(vcol + 8u * (unsigned)(TC * PW), wc * rc);
    sts(vcol + 8u * (unsigned)(TD * PW), wd * rdd);
}

// On return (true): the rows of S above the diagonal blocks hold W (un-normalised: U = diag(1/d) W), sm.Linv / sm.rd the
// block inverses and reciprocal pivots, sm.y the forward-substituted right-hand side in the same un-normalised form.
// Whole CTA calls with blockDim.x = 512; no thread may skip it.  Ends with a CTA barrier.
// Roles: warp 0 factors; warps 0-3 (one per scheduler) form the look-ahead columns and the next diagonal block while
// everybody else waits; warps 4, 8, 12 (warp 0's scheduler) never run; the other twelve warps are the workers.
template <class M, class Prof = DenseProfile>
__device__ bool ldl_lookahead(M S, int D, const LdlSmem &sm, int PW, Prof prof = Prof()) {
    typedef typename M::ptr mptr;
    const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31, wid = tid >> 5;
    const int nblk = (D + NB - 1) / NB;
    const int zc = PW - 1;
    const bool prio = wid < 4;                 // one warp per scheduler: fp64 issue is per scheduler
    const bool idle = wid >= 4 && (wid & 3) == 0;  // warps 4, 8, 12: nothing but warp 0 ever runs on scheduler 0
    const int ptid = tid;                      // 0..127 inside the priority group
    const int nworkers = nt - 128;             // 384: warps 1-3, 5-7, 9-11, 13-15
    const int wrank = wid - 1 - (wid / 4);     // dense rank of a worker warp (1, 2, 3 -> 0, 1, 2; 5 -> 3; ..)
    const int nww = nworkers >> 5;             // 12
    const int wtid = wrank * 32 + lane;
    const int n_part = nt - 96;
    if (tid < 128) {
        const int nb = min(NB, D);
        for (int q = tid; q < NB * NB; q += 128) {
            const int r = q / NB, c = q % NB;
            sts(sm.Dg + 8u * q, (r < nb && c < nb && c >= r) ? M::ld(S.at(r, c)) : ((r == c) ? 1.0 : 0.0));
        }
    }
    __syncthreads();
    bool all_ok = true;
#ifdef LDL_PROFILE
    // stamps of warps 0 and 5 only, kept in registers of lane 0 until the end (stores would perturb the run)
    long long ev_[8] = {0, 0, 0, 0, 0, 0, 0, 0}, evt_ = clock64();
#define LDL_MARK(i) do { if (wid == 0 || wid == 5) { const long long n_ = clock64(); ev_[i] += n_ - evt_; evt_ = n_; } } while (0)
#else
#define LDL_MARK(i)
#endif
    if (!idle)
    for (int blk = 0; blk < nblk; blk++) {
        const int j0 = blk * NB, nb = min(NB, D - j0);
        const int rest0 = j0 + nb, nrest = D - rest0, nb1 = min(NB, nrest);
        const uint32_t Li = sm.Linv + 8u * (blk * NB * NB), rd = sm.rd + 8u * (blk * NB);
        LDL_MARK(7);
        if (wid == 0) {
            const bool ok = ldl_factor_inv(sm.Dg, Li, rd, sm.rowx);
            if (lane == 0) stsi(sm.ok + 4u * blk, ok ? 1 : 0);
        }
        LDL_MARK(0);
        bar_sync(LDL_BAR_CTA, n_part);
        LDL_MARK(1);
        if (ldsi(sm.ok + 4u * blk) == 0) {
            all_ok = false;
            break;
        }
        int lenA, startB, lenB;
        prof.active(j0, nb, D, nb1, lenA, startB, lenB);
        const int nact = lenA + lenB;
        auto ccol = [&](int c) { return c < lenA ? rest0 + c : startB + (c - lenA); };
        if (prio) {
#if defined(LDL_SKIP) && (LDL_SKIP & 4)
            if (false) {
#else
            if (nb1 > 0) {
#endif
                // ---- the 16 look-ahead panel columns, then the next diagonal block ----
                const int c = ptid & 15, p = ptid >> 4, t1 = p, t2 = NB - 1 - p;
#if defined(LDL_SKIP) && (LDL_SKIP & 32)
                if (false) {
#else
                if (c < nb1) {
#endif
                    mptr sp = S.at(j0, rest0 + c);
                    const uint32_t l1 = Li + 8u * (t1 * NB), l2 = Li + 8u * (t2 * NB);
                    const double r1 = lds(rd + 8u * t1), r2 = lds(rd + 8u * t2);
                    double av[NB];
#pragma unroll
                    for (int r = 0; r < NB; r++) {
                        av[r] = (r < nb) ? M::ld(sp) : 0.0;
                        sp = S.down(sp, j0 + r);
                    }
                    double e1 = 0, o1 = 0, e2 = 0, o2 = 0;
#pragma unroll
                    for (int r = 0; r < NB; r += 2) {
                        const double2 x1 = lds2(l1 + 8u * r), x2 = lds2(l2 + 8u * r);
                        e1 = fma(x1.x, av[r], e1);
                        o1 = fma(x1.y, av[r + 1], o1);
                        e2 = fma(x2.x, av[r], e2);
                        o2 = fma(x2.y, av[r + 1], o2);
                    }
                    const double w1 = e1 + o1, w2 = e2 + o2;
                    sts(sm.panel + 8u * (unsigned)(t1 * PW + c), w1);
                    sts(sm.panel + 8u * (unsigned)(t2 * PW + c), w2);
                    sts(sm.panelV + 8u * (unsigned)(t1 * PW + c), w1 * r1);
                    sts(sm.panelV + 8u * (unsigned)(t2 * PW + c), w2 * r2);
                }
                LDL_MARK(2);
                bar_sync(LDL_BAR_PRIO, 128);
                LDL_MARK(3);
#if defined(LDL_SKIP) && (LDL_SKIP & 16)
                if (false)
#endif
                {
                    const bool use1 = t1 <= c && c < nb1, use2 = t2 <= c && c < nb1;
                    const double s1 = use1 ? M::ld(S.at(rest0 + t1, rest0 + c)) : ((t1 == c) ? 1.0 : 0.0);
                    const double s2 = use2 ? M::ld(S.at(rest0 + t2, rest0 + c)) : ((t2 == c) ? 1.0 : 0.0);
                    double a1e = 0, a1o = 0, a2e = 0, a2o = 0;
                    if (c < nb1) {
#pragma unroll
                        for (int t = 0; t < NB; t += 2) {
                            const uint32_t pr0 = sm.panel + 8u * (unsigned)(t * PW), pr1 = pr0 + 8u * (unsigned)PW;
                            const double v0 = lds(sm.panelV + 8u * (unsigned)(t * PW + c)), v1 = lds(sm.panelV + 8u * (unsigned)((t + 1) * PW + c));
                            a1e = fma(lds(pr0 + 8u * t1), v0, a1e);
                            a2e = fma(lds(pr0 + 8u * t2), v0, a2e);
                            a1o = fma(lds(pr1 + 8u * t1), v1, a1o);
                            a2o = fma(lds(pr1 + 8u * t2), v1, a2o);
                        }
                    }
                    sts(sm.Dg + 8u * (t1 * NB + c), use1 ? s1 - (a1e + a1o) : s1);
                    sts(sm.Dg + 8u * (t2 * NB + c), use2 ? s2 - (a2e + a2o) : s2);
                }
                LDL_MARK(4);
                bar_sync(LDL_BAR_PRIO, 128);
                LDL_MARK(5);
            }
            // the other warps were held back until here: the priority phase has the fp64 pipes and shared memory to itself
            if (wid == 0) {
                bar_arrive(LDL_BAR_GO, n_part);
                continue;  // warp 0 goes on with the next block
            }
            bar_sync(LDL_BAR_GO, n_part);
        } else {
            bar_sync(LDL_BAR_GO, n_part);
        }
        // ---- workers: the other panel columns + the right-hand side.  Unit = (chunk of 32 columns, group of 2 row pairs):
        //      the pair index is warp-uniform and static, the column is read once per unit. ----
        {
            const int ncol = nact - nb1 + 1, nchunk = (ncol + 31) >> 5;
#if defined(LDL_SKIP) && (LDL_SKIP & 2)
            if (false)
#endif
            for (int unit = wrank; unit < 4 * nchunk; unit += nww) {
                const int g = unit & 3, q = (unit >> 2) * 32 + lane;
                if (q < ncol) {
                    const bool is_rhs = q == ncol - 1;
                    const int c = nb1 + q;
                    double a[NB];
                    if (is_rhs) {
#pragma unroll
                        for (int r = 0; r < NB; r++) a[r] = (r < nb) ? lds(sm.y + 8u * (j0 + r)) : 0.0;
                    } else {
                        mptr sp = S.at(j0, ccol(c));
#pragma unroll
                        for (int r = 0; r < NB; r++) {
                            a[r] = (r < nb) ? M::ld(sp) : 0.0;
                            sp = S.down(sp, j0 + r);
                        }
                    }
                    const uint32_t pcol = sm.panel + 8u * (unsigned)(is_rhs ? zc : c), vcol = sm.panelV + 8u * (unsigned)(is_rhs ? zc : c);
                    if (g == 0) ldl_panel_2pairs<0>(Li, rd, a, pcol, vcol, PW);
                    else if (g == 1) ldl_panel_2pairs<2>(Li, rd, a, pcol, vcol, PW);
                    else if (g == 2) ldl_panel_2pairs<4>(Li, rd, a, pcol, vcol, PW);
                    else ldl_panel_2pairs<6>(Li, rd, a, pcol, vcol, PW);
                }
            }
        }
        LDL_MARK(2);
        bar_sync(LDL_BAR_WORK, nworkers);
        LDL_MARK(3);
        // ---- W back into S (the rows of this block; read again by the back substitution) and into the right-hand side ----
#if defined(LDL_SKIP) && (LDL_SKIP & 8)
        if (false)
#endif
        for (int c = wtid; c <= nact; c += nworkers) {
            if (c == nact) {
                for (int r = 0; r < nb; r++) sts(sm.y + 8u * (j0 + r), lds(sm.panel + 8u * (unsigned)(r * PW + zc)));
            } else {
                mptr sp = S.at(j0, ccol(c));
                for (int r = 0; r < nb; r++) {
                    M::st(sp, lds(sm.panel + 8u * (unsigned)(r * PW + c)));
                    sp = S.down(sp, j0 + r);
                }
            }
        }
        // ---- trailing update A22[i][k] -= sum_t W[t][i] V[t][k] (k >= i): one warp per 16 x 8 half tile, each thread a 2 x 2
        //      sub-tile in registers (two 128-bit shared loads per 4 FMAs).  Tile (0, 0) — the next diagonal block — was
        //      done by the priority group. ----
        const int nbk = (nact + NB - 1) / NB, nunits = nbk * (nbk + 1) / 2;
        const int ty = lane >> 2, tx = lane & 3;
#if defined(LDL_SKIP) && (LDL_SKIP & 1)
        if (false)
#endif
        for (int u = 2 + wrank; u < 2 * nunits; u += nww) {
            const int q = u >> 1;
            int bi = 0, rem = q;
            while (rem >= nbk - bi) {
                rem -= nbk - bi;
                bi++;
            }
            const int bk = bi + rem, i0 = bi * NB + 2 * ty, k0 = bk * NB + 8 * (u & 1) + 2 * tx;
            if (i0 > k0 + 1 || k0 >= nact) continue;  // below the diagonal / past the last column
            double c00 = 0, c01 = 0, c10 = 0, c11 = 0;
#pragma unroll 8
            for (int t = 0; t < NB; t++) {
                const double2 w = lds2(sm.panel + 8u * (unsigned)(t * PW + i0)), v = lds2(sm.panelV + 8u * (unsigned)(t * PW + k0));
                c00 = fma(w.x, v.x, c00);
                c01 = fma(w.x, v.y, c01);
                c10 = fma(w.y, v.x, c10);
                c11 = fma(w.y, v.y, c11);
            }
            const int rk0 = ccol(k0), rk1 = ccol(k0 + 1);
            const bool k1ok = k0 + 1 < nact;
            {
                const mptr rp = S.row0(ccol(i0));
                if (i0 <= k0) {
                    const mptr e = M::right(rp, rk0);
                    M::st(e, M::ld(e) - c00);
                }
                if (k1ok) {
                    const mptr e = M::right(rp, rk1);
                    M::st(e, M::ld(e) - c01);
                }
            }
            if (i0 + 1 < nact) {
                const mptr rp = S.row0(ccol(i0 + 1));
                if (i0 + 1 <= k0) {
                    const mptr e = M::right(rp, rk0);
                    M::st(e, M::ld(e) - c10);
                }
                if (k1ok && i0 + 1 <= k0 + 1) {
                    const mptr e = M::right(rp, rk1);
                    M::st(e, M::ld(e) - c11);
                }
            }
        }
        // right-hand side: b_rest -= W^T (z / d)
        for (int i = nworkers - 1 - wtid; i < nact; i += nworkers) {
            double sacc = 0;
#pragma unroll 4
            for (int t = 0; t < NB; t++) sacc = fma(lds(sm.panel + 8u * (unsigned)(t * PW + i)), lds(sm.panelV + 8u * (unsigned)(t * PW + zc)), sacc);
            const uint32_t ya = sm.y + 8u * ccol(i);
            sts(ya, lds(ya) - sacc);
        }
        LDL_MARK(4);
    }
#ifdef LDL_PROFILE
    if (lane == 0 && (wid == 0 || wid == 5)) for (int i = 0; i < 8; i++) g_ldl_ev[wid][0][i] = ev_[i];
#endif
    if (tid == 0) stsi(sm.ok + 4u * nblk, all_ok ? 1 : 0);
    __syncthreads();
    return ldsi(sm.ok + 4u * nblk) != 0;
}

// Uu y = diag(1/d) z on the shared vector (in: the un-normalised z of ldl_lookahead, out: y), last block first.
template <class M>
__device__ void ldl_solve_back(M S, int D, const LdlSmem &sm) {
    typedef typename M::ptr mptr;
    const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31;
    const int nblk = (D + NB - 1) / NB;
    for (int blk = nblk - 1; blk >= 0; blk--) {
        const int j0 = blk * NB, nb = min(NB, D - j0);
        const uint32_t yb = sm.y + 8u * j0;
        if (tid < 32) {
            const int c = lane & 15, r0 = (lane >> 4) * 8;
            const uint32_t Li = sm.Linv + 8u * (blk * NB * NB), rd = sm.rd + 8u * (blk * NB);
            double e = 0, o = 0;
#pragma unroll
            for (int u = 0; u < 8; u += 2) {
                const int r = r0 + u;
                const double z0 = (r < nb) ? lds(yb + 8u * r) * lds(rd + 8u * r) : 0.0;
                const double z1 = (r + 1 < nb) ? lds(yb + 8u * (r + 1)) * lds(rd + 8u * (r + 1)) : 0.0;
                e = fma(lds(Li + 8u * (r * NB + c)), z0, e);
                o = fma(lds(Li + 8u * ((r + 1) * NB + c)), z1, o);
            }
            double acc = e + o;
            acc += __shfl_xor_sync(0xffffffffu, acc, 16);
            if (lane < nb) sts(yb + 8u * lane, acc);
        }
        __syncthreads();
        for (int i = tid; i < j0; i += nt) {
            mptr sp = S.at(i, j0);
            double e = 0, o = 0;
#pragma unroll 4
            for (int t = 0; t < NB; t += 2) {
                if (t < nb) e = fma(M::ld(M::right(sp, t)), lds(yb + 8u * t), e);
                if (t + 1 < nb) o = fma(M::ld(M::right(sp, t + 1)), lds(yb + 8u * (t + 1)), o);
            }
            const uint32_t ya = sm.y + 8u * i;
            sts(ya, lds(ya) - (e + o));
        }
        __syncthreads();
    }
}

#endif  // __CUDACC__

}  // namespace lfvio
