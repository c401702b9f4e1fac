// Blocked look-ahead Cholesky (one CTA) shared by the solver's reduced system (solver_step.cu) and the square root of
// the marginalisation's Schur complement (marg.cu).  See cholesky_lookahead below for the scheme.
#pragma once
#include "chol16.cuh"

namespace lfvio {

#ifdef __CUDACC__
#ifdef LFVIO_CHOL_PROFILE
#define SUB_T0() long long sub_t = clock64(); long long sub_acc[8] = {0, 0, 0, 0, 0, 0, 0, 0}
#define SUB_MARK(i) do { const long long now = clock64(); sub_acc[i] += now - sub_t; sub_t = now; } while (0)
#define SUB_FLUSH() do { if (threadIdx.x == 0 || threadIdx.x == 32) for (int i = 0; i < 8; i++) g_sub[(threadIdx.x >> 5) * 8 + i] = sub_acc[i]; } while (0)
static __device__ long long g_sub[16];
#else
#define SUB_T0()
#define SUB_MARK(i)
#define SUB_FLUSH()
#endif
// row pitch of the shared panel: trailing columns rounded up to the 4-wide tiles, + one tile whose last column holds the rhs
__host__ __device__ inline int chol_panel_pitch(int D) { return ((D + 3) & ~3) + 4; }

// Element handles for the hot loops: a pointer walk down a column costs one add per row; for the shared-memory triangle the
// handle is a 32-bit shared address (a generic pointer into shared memory costs an S2R + LEA per access to rebuild the
// window base, the packed index five integer operations — measured: the workers spent half their time "loading 16 values").
struct MatFull {
    double *p;
    int D;
    __device__ __forceinline__ double &at(int r, int c) const { return p[(size_t)r * D + c]; }
    __device__ __forceinline__ double *row(int r) const { return p + (size_t)r * D; }  // row(r)[c] == at(r, c)
    typedef double *hnd;
    __device__ __forceinline__ hnd h(int r, int c) const { return p + (size_t)r * D + c; }
    __device__ __forceinline__ hnd down(hnd e, int r) const { return e + D; }  // (r, c) -> (r + 1, c)
    static __device__ __forceinline__ hnd right(hnd e, int n) { return e + n; }
    static __device__ __forceinline__ double ld(hnd e) { return *e; }
    static __device__ __forceinline__ void st(hnd e, double v) { *e = v; }
};
struct MatPacked {  // packed upper triangle, row-major (c >= r)
    double *p;
    int D;
    __device__ __forceinline__ double &at(int r, int c) const { return p[r * D - ((r * (r - 1)) >> 1) + (c - r)]; }
    __device__ __forceinline__ double *row(int r) const { return p + (r * D - ((r * (r - 1)) >> 1) - r); }  // valid for c >= r
    typedef uint32_t hnd;
    __device__ __forceinline__ hnd h(int r, int c) const { return smem_u32(p) + 8u * (unsigned)(r * D - ((r * (r - 1)) >> 1) + (c - r)); }
    __device__ __forceinline__ hnd down(hnd e, int r) const { return e + 8u * (unsigned)(D - r - 1); }
    static __device__ __forceinline__ hnd right(hnd e, int n) { return e + 8u * (unsigned)n; }
    static __device__ __forceinline__ double ld(hnd e) { return lds(e); }
    static __device__ __forceinline__ void st(hnd e, double v) { sts(e, v); }
};

// ---- look-ahead blocked Cholesky + fused forward substitution ---------------------------------------------------
// S = U^T U on the upper triangle, 16-column block steps, one CTA.  Warp 0 is the "diagonal warp": while the other
// warps (workers) apply block step k to the trailing matrix, warp 0 already forms the 16 panel columns of block k+1,
// applies step k to diagonal block k+1 and factors it in registers (factor_diag_block_ldl, which also hands out U11^-T so
// that no panel column is a substitution) — the serial pivot chain of block k+1 overlaps the trailing update of block k.  The right-hand side rides along as one extra panel column, so the
// forward substitution U^T z = b costs nothing extra; solve_upper_back finishes with U y = z.
// Barriers: id 1 = "diagonal block ready / trailing update done" (all threads), id 2 = "panel complete"
// (workers wait, warp 0 only arrives).
__device__ __forceinline__ void bar_sync(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }
__device__ __forceinline__ void bar_arrive(int id, int n) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(n) : "memory"); }

// one panel column through the explicit inverse of the diagonal block: a <- U11^-T a with U11^-T lower triangular in shared
// memory (factor_diag_block_ldl) — 136 independent-ish FMAs instead of a 16-step substitution; in place, last row first
__device__ __forceinline__ void panel_column_apply(double (&a)[NB], uint32_t pl) {
#pragma unroll
    for (int t = NB - 1; t >= 0; t--) {
        double e = 0, o = 0;
#pragma unroll
        for (int r = 0; r <= t; r += 2) {
            const double2 l = lds2(pl + 8u * (t * NB + r));
            e = fma(l.x, a[r], e);
            if (r + 1 <= t) o = fma(l.y, a[r + 1], o);
        }
        a[t] = e + o;
    }
}

// Sparsity profiles: which trailing columns a block step [j0, j0 + nb) can touch, as two contiguous ranges
// [rest0, rest0 + lenA) and [startB, startB + lenB) (lenA >= nb1 so that the look-ahead block is always inside).
struct DenseProfile {
    __device__ __forceinline__ void active(int j0, int nb, int D, int nb1, int &lenA, int &startB, int &lenB) const {
        lenA = D - (j0 + nb);
        startB = 0;
        lenB = 0;
    }
};
// Reduced system of the sliding window in the column order of StateLayout: [sb_{F-1} .. sb_1 | sb_0 | pose_{F-1} ..
// pose_0 | ex | td].  A speed-bias block sb_k (k >= 1) is coupled only to sb_{k+-1} and pose_{k-1..k+1} (IMU factors);
// eliminating sb_{F-1}, sb_{F-2}, ... in this order the fill of row sb_k stays inside sb_{k-1} and pose_{F-1} .. pose_{k-1},
// a PREFIX of the pose section.  Everything from sb_0 on (prior, landmarks' Schur complement) is dense.
struct WindowProfile {
    int F;
    __device__ __forceinline__ void active(int j0, int nb, int D, int nb1, int &lenA, int &startB, int &lenB) const {
        const int rest0 = j0 + nb, jl = rest0 - 1;
        lenA = D - rest0;
        startB = 0;
        lenB = 0;
        if (jl >= 9 * (F - 1)) return;               // the block reaches sb_0: dense tail
        const int k_lo = F - 1 - jl / 9;             // lowest speed-bias index inside the block (>= 1)
        const int a_end = 9 * (F - k_lo + 1);        // end of sb_{k_lo - 1}
        const int la = max(a_end - rest0, nb1), sb = 9 * F, lb = 6 * (F - k_lo + 1);
        if (rest0 + la >= sb) {
            lenA = sb + lb - rest0;                  // ranges touch: one contiguous range up to the end of the pose prefix
            return;
        }
        lenA = la;
        startB = sb;
        lenB = lb;
    }
};

template <class M, class Prof = DenseProfile>
__device__ bool cholesky_lookahead(M S, int D, double *s_panel /* NB x PW */, int PW, double *s_U /* 2 x NB x NB */, double *s_PL /* 2 (or, pl_all, nblk) x NB x NB */,
                                   double *s_invd /* nblk x NB */, double *s_y /* D: b in, z out */, int *s_okv /* nblk */, double *s_row /* 2 x NB */,
                                   Prof prof = Prof(), bool leave_last = false, bool pl_all = false) {
    // pl_all: keep U11^-T of every block (s_PL holds nblk of them) for solve_upper_back_inv; else two buffers alternate
    // leave_last: stop before the last diagonal block — on return it holds the fully updated (unfactored) Schur
    // complement of the leading blocks and s_y its updated right-hand side; the caller factors it its own way
    // (marg.cu: with diagonal pivoting, because that is where a gauge-deficient matrix shows its tiny pivots).
    const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31, wid = tid >> 5;
    const int nblk = (D + NB - 1) / NB;
    // Warps 4, 8, 12, .. share warp 0's scheduler (warp w runs on scheduler w % 4; fp64 issue is per scheduler: measured, the
    // diagonal-block factorisation takes 1.8 k cycles alone and 3.3 k next to three busy warps there): they sit the whole
    // factorisation out, so that the chain warp never waits for an issue slot or the fp64 pipe.
    const bool idle = wid >= 4 && (wid & 3) == 0;
    const int n_idle = ((nt >> 5) - 1) / 4;
    const int nworkers = nt - 32 - 32 * n_idle, wtid = (wid - 1 - wid / 4) * 32 + lane;
    const int nbar = nt - 32 * n_idle;  // threads taking part in the named barriers
    const int zc = PW - 1;           // panel column of the right-hand side
    SUB_T0();
    if (wid == 0) {  // factor diagonal block 0
        const int nb = min(NB, D);
        for (int q = lane; q < NB * NB; q += 32) {
            const int r = q / NB, c = q % NB;
            s_U[q] = (r < nb && c < nb && c >= r) ? S.at(r, c) : ((r == c) ? 1.0 : 0.0);
        }
        __syncwarp();
        const bool ok = factor_diag_block_ldl(s_U, s_invd, s_PL, s_row);  // (block 0 -> buffer 0 either way)
        if (lane == 0) s_okv[0] = ok ? 1 : 0;
        for (int q = lane; q < NB * NB; q += 32) {
            const int r = q / NB, c = q % NB;
            if (c >= r && c < nb) S.at(r, c) = s_U[q];
        }
    }
    if (!idle) bar_sync(1, nbar);
#ifdef LFVIO_CHOL_PROFILE
    sub_t = clock64();
#endif
    bool ok_all = true;
    for (int blk = 0; blk < nblk && !idle; blk++) {
        if (s_okv[blk] == 0) {
            ok_all = false;
            break;
        }
        if (leave_last && blk == nblk - 1) break;
        const int j0 = blk * NB, nb = min(NB, D - j0);
        const int rest0 = j0 + nb, nrest = D - rest0, nb1 = min(NB, nrest);
        const uint32_t pl = smem_u32(s_PL + (pl_all ? blk : (blk & 1)) * NB * NB);  // U11^-T of this block
        if (wid == 0) {
            // ---- diagonal warp: panel columns of block blk+1, then look ahead ----
            double *Un = s_U + ((blk + 1) & 1) * NB * NB;
            {
                // the 16 look-ahead columns through U11^-T: lane (h, c) forms rows h, h + 2, .. of column c (72 FMAs per lane)
                const int c = lane & 15, h = lane >> 4;
                if (c < nb1) {
                    double a[NB];
#pragma unroll
                    {
                        typename M::hnd e = S.h(j0, rest0 + c);
#pragma unroll
                        for (int r = 0; r < NB; r++) {
                            a[r] = (r < nb) ? M::ld(e) : 0.0;
                            e = S.down(e, j0 + r);
                        }
                    }
                    double out[NB / 2];
#pragma unroll
                    for (int u = 0; u < NB / 2; u++) {
                        const uint32_t prow = pl + 8u * ((2 * u + h) * NB);
                        double e = 0, o = 0;
#pragma unroll
                        for (int r = 0; r <= 2 * u + 1; r += 2) {  // (entry 2u + 1 of row 2u is zero)
                            const double2 l = lds2(prow + 8u * r);
                            e = fma(l.x, a[r], e);
                            o = fma(l.y, a[r + 1], o);
                        }
                        out[u] = e + o;
                    }
                    typename M::hnd e = S.h(j0 + h, rest0 + c);
#pragma unroll
                    for (int u = 0; u < NB / 2; u++) {
                        const int t = 2 * u + h;
                        if (t < nb) M::st(e, out[u]);
                        e = S.down(S.down(e, j0 + t), j0 + t + 1);
                        s_panel[t * PW + c] = out[u];
                    }
                }
            }
            __threadfence_block();
            bar_arrive(2, nbar);
            SUB_MARK(0);
            if (nb1 > 0) {
                __syncwarp();
                // step blk applied to diagonal block blk+1: lane (h, c) owns rows 8h .. 8h+7 of column c
                const int c = lane & 15, i0 = (lane >> 4) * 8;
                double acc[8];
#pragma unroll
                for (int u = 0; u < 8; u++) acc[u] = 0.0;
                if (c < nb1) {
#pragma unroll 4
                    for (int t = 0; t < NB; t++) {
                        const double pc = s_panel[t * PW + c];
#pragma unroll
                        for (int u = 0; u < 8; u++) acc[u] = fma(s_panel[t * PW + i0 + u], pc, acc[u]);
                    }
                }
#pragma unroll
                for (int u = 0; u < 8; u++) {
                    const int i = i0 + u;
                    Un[i * NB + c] = (i <= c && c < nb1) ? S.at(rest0 + i, rest0 + c) - acc[u] : ((i == c) ? 1.0 : 0.0);
                }
                __syncwarp();
                SUB_MARK(1);
                const bool ok = (leave_last && blk + 1 == nblk - 1) ? true : factor_diag_block_ldl(Un, s_invd + (blk + 1) * NB, s_PL + (pl_all ? blk + 1 : ((blk + 1) & 1)) * NB * NB, s_row);
                if (lane == 0) s_okv[blk + 1] = ok ? 1 : 0;
                for (int q = lane; q < NB * NB; q += 32) {
                    const int r = q / NB, cc = q % NB;
                    if (cc >= r && cc < nb1) S.at(rest0 + r, rest0 + cc) = Un[q];
                }
                SUB_MARK(2);
            }
        } else {
            // ---- workers: remaining panel columns + right-hand side ----
            // The trailing matrix is addressed through COMPACT indices 0 .. nact-1 over the profile's two ranges;
            // the panel is stored compactly, only the reads/writes of S translate to real columns.
            int lenA, startB, lenB;
            prof.active(j0, nb, D, nb1, lenA, startB, lenB);
            const int nact = lenA + lenB;
            auto ccol = [&](int c) { return c < lenA ? rest0 + c : startB + (c - lenA); };
            const int ncol = nact - nb1 + 1;  // last one is the right-hand side
            // (one pass: D < number of workers, checked by launch_step)
            if (const int q = wtid; q < ncol) {
                const bool is_rhs = q == ncol - 1;
                const int c = nb1 + q;  // compact column inside the trailing matrix
                const int rc = ccol(c);
                double a[NB];
#pragma unroll
                for (int r = 0; r < NB; r++) a[r] = 0.0;
                const typename M::hnd e0 = S.h(j0, is_rhs ? j0 : rc);
                if (is_rhs) {
#pragma unroll
                    for (int r = 0; r < NB; r++)
                        if (r < nb) a[r] = s_y[j0 + r];
                } else {
                    typename M::hnd e = e0;
#pragma unroll
                    for (int r = 0; r < NB; r++) {
                        if (r < nb) a[r] = M::ld(e);
                        e = S.down(e, j0 + r);
                    }
                }
                SUB_MARK(4);
                panel_column_apply(a, pl);
                SUB_MARK(5);
                const int pc = is_rhs ? zc : c;
                typename M::hnd e = e0;
#pragma unroll
                for (int r = 0; r < NB; r++) {
                    if (r < nb) {
                        if (is_rhs) s_y[j0 + r] = a[r];
                        else M::st(e, a[r]);
                    }
                    e = S.down(e, j0 + r);
                    s_panel[r * PW + pc] = a[r];
                }
            }
            SUB_MARK(0);
            bar_sync(2, nbar);
            SUB_MARK(1);
            // ---- trailing update A22[i][k] -= sum_t U12[t][i] U12[t][k] (k >= i) in 16 x 16 blocks: 16 threads per
            //      block, thread y owns column y of it (16 accumulators).  Per rank-1 step the 16 row factors are one
            //      broadcast read and the column factor one conflict-free read; the read-modify-write of S touches 16
            //      consecutive doubles per row.  Block (0, 0) — diagonal block blk+1 — belongs to warp 0. ----
            const int nbk = (nact + NB - 1) / NB, nunits = nbk * (nbk + 1) / 2;
            const int y = wtid & 15;
            for (int q = 1 + (wtid >> 4); q < nunits; q += nworkers >> 4) {
                int bi = 0, rem = q;
                while (rem >= nbk - bi) {
                    rem -= nbk - bi;
                    bi++;
                }
                const int bk = bi + rem, i0 = bi * NB, k = bk * NB + y;
                double acc[NB];
#pragma unroll
                for (int u = 0; u < NB; u++) acc[u] = 0.0;
#pragma unroll 2
                for (int t = 0; t < NB; t++) {
                    const double pk = s_panel[t * PW + k];
                    const double2 *ap = reinterpret_cast<const double2 *>(&s_panel[t * PW + i0]);
#pragma unroll
                    for (int u2 = 0; u2 < NB / 2; u2++) {
                        const double2 av = ap[u2];
                        acc[2 * u2] = fma(av.x, pk, acc[2 * u2]);
                        acc[2 * u2 + 1] = fma(av.y, pk, acc[2 * u2 + 1]);
                    }
                }
                if (k < nact) {
                    const int rk = ccol(k);
#pragma unroll
                    for (int u = 0; u < NB; u++) {
                        const int i = i0 + u;
                        if (i <= k) S.row(ccol(i))[rk] -= acc[u];  // (i <= k < nact; the mapping is monotone)
                    }
                }
            }
            // right-hand side: b_rest -= U12^T z (taken by the workers that got the fewest tiles)
            for (int i = nworkers - 1 - wtid; i < nact; i += nworkers) {
                double sacc = 0;
#pragma unroll 4
                for (int t = 0; t < NB; t++) sacc = fma(s_panel[t * PW + i], s_panel[t * PW + zc], sacc);
                s_y[ccol(i)] -= sacc;
            }
            SUB_MARK(2);
        }
        bar_sync(1, nbar);
        SUB_MARK(3);
    }
    SUB_FLUSH();
    if (tid == 0) s_okv[nblk] = ok_all ? 1 : 0;  // (the idle warps never looked)
    __syncthreads();
    return s_okv[nblk] != 0;
}

// U y = z on the shared vector y (in/out), last block first: warp 0 back-substitutes the 16 x 16 diagonal block with
// shuffles, then every thread folds the 16 new unknowns into the rows above.  Whole CTA calls.
template <class M>
__device__ void solve_upper_back(M S, int D, double *y, const double *s_invd) {
    const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31;
    const int nblk = (D + NB - 1) / NB;
    for (int blk = nblk - 1; blk >= 0; blk--) {
        const int j0 = blk * NB, nb = min(NB, D - j0);
        if (tid < 32) {
            const int r = lane & 15;
            double urow[NB];
#pragma unroll
            for (int t = 0; t < NB; t++) urow[t] = (r < nb && t > r && t < nb) ? S.at(j0 + r, j0 + t) : 0.0;
            double wv = (r < nb) ? y[j0 + r] : 0.0;
            const double inv = (r < nb) ? s_invd[blk * NB + r] : 1.0;
#pragma unroll
            for (int t = NB - 1; t >= 0; t--) {
                const double yt = __shfl_sync(0xffffffffu, wv * inv, t);
                wv = fma(-urow[t], yt, wv);  // urow[t] == 0 for t <= r
            }
            if (lane < nb) y[j0 + lane] = wv * inv;
        }
        __syncthreads();
        for (int i = tid; i < j0; i += nt) {
            double sacc = 0;
#pragma unroll 4
            for (int t = 0; t < nb; t++) sacc = fma(S.at(i, j0 + t), y[j0 + t], sacc);
            y[i] -= sacc;
        }
        __syncthreads();
    }
}

// The same with the explicit inverses U11^-T of all diagonal blocks (cholesky_lookahead with pl_all): the block solve is a
// 16-term dot product per unknown (two half sums + one shuffle) instead of 16 dependent shuffle + FMA steps.
template <class M>
__device__ void solve_upper_back_inv(M S, int D, double *y, const double *s_PL) {
    const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31;
    const int nblk = (D + NB - 1) / NB;
    const uint32_t ya = smem_u32(y), pla = smem_u32(s_PL);
    for (int blk = nblk - 1; blk >= 0; blk--) {
        const int j0 = blk * NB, nb = min(NB, D - j0);
        if (tid < 32) {
            const int c = lane & 15, r0 = (lane >> 4) * 8;
            const uint32_t pl = pla + 8u * (blk * NB * NB);
            double e = 0, o = 0;
#pragma unroll
            for (int u = 0; u < 8; u += 2) {
                const int r = r0 + u;
                const double z0 = (r < nb) ? lds(ya + 8u * (j0 + r)) : 0.0, z1 = (r + 1 < nb) ? lds(ya + 8u * (j0 + r + 1)) : 0.0;
                e = fma(lds(pl + 8u * (r * NB + c)), z0, e);  // U^-1 = (U^-T)^T: column c of the stored rows (zero above the diagonal)
                o = fma(lds(pl + 8u * ((r + 1) * NB + c)), z1, o);
            }
            double acc = e + o;
            acc += __shfl_xor_sync(0xffffffffu, acc, 16);
            __syncwarp();  // every lane has read its z before any lane overwrites one
            if (lane < nb) sts(ya + 8u * (j0 + lane), acc);
        }
        __syncthreads();
        for (int i = tid; i < j0; i += nt) {
            const typename M::hnd e = S.h(i, j0);
            double s0 = 0, s1 = 0;
#pragma unroll 4
            for (int t = 0; t < NB; t += 2) {
                if (t < nb) s0 = fma(M::ld(M::right(e, t)), lds(ya + 8u * (j0 + t)), s0);
                if (t + 1 < nb) s1 = fma(M::ld(M::right(e, t + 1)), lds(ya + 8u * (j0 + t + 1)), s1);
            }
            sts(ya + 8u * i, lds(ya + 8u * i) - (s0 + s1));
        }
        __syncthreads();
    }
}

// The same with the block chain taken off the barrier: warp 0 solves block b through U11^-1, publishes y_b and forms ITSELF the
// correction of the 16 rows of block b - 1 (kept in registers and applied when that block's turn comes), while the other warps
// subtract y_b from the rows of the blocks below b - 1 — one chain step behind.  Named barriers: 3/4 "y_b published" (workers
// wait, warp 0 arrives), 5/6 "rows updated with y_b" (warp 0 waits, workers arrive), alternating by block parity.  Warp 0 waits
// for the update with y_{b+2} before it touches block b, and a worker can only arrive for b - 2 after warp 0 has passed that
// wait, so no barrier is ever entered for two blocks at once.  Warps 4, 8, 12 (warp 0's scheduler) sit it out as in the
// factorisation.  11 blocks of the 10-frame window: 8.2 us -> see DESIGN.md.
template <class M>
__device__ void solve_upper_back_la(M S, int D, double *y, const double *s_PL) {
    const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31, wid = tid >> 5;
    const int nblk = (D + NB - 1) / NB;
    const bool idle = wid >= 4 && (wid & 3) == 0;
    const int n_idle = ((nt >> 5) - 1) / 4;
    const int nworkers = nt - 32 - 32 * n_idle, wtid = (wid - 1 - wid / 4) * 32 + lane;
    const int nbar = nt - 32 * n_idle;
    const uint32_t ya = smem_u32(y), pla = smem_u32(s_PL);
    __syncthreads();  // z complete
    if (wid == 0) {
        const int c = lane & 15, hh = lane >> 4;
        double corr = 0;  // lane c: what y_{b+1} takes off row c of block b
        for (int blk = nblk - 1; blk >= 0; blk--) {
            const int j0 = blk * NB, nb = min(NB, D - j0);
            if (blk + 2 < nblk) bar_sync(5 + ((blk + 2) & 1), nbar);  // rows of this block carry y_{b+2}, y_{b+3}, ..
            if (blk + 1 < nblk) {
                if (lane < NB) sts(ya + 8u * (j0 + lane), lds(ya + 8u * (j0 + lane)) - corr);
                __syncwarp();
            }
            {
                const int r0 = hh * 8;
                const uint32_t pl = pla + 8u * (blk * NB * NB);
                double e = 0, o = 0;
#pragma unroll
                for (int u = 0; u < 8; u += 2) {
                    const int r = r0 + u;
                    const double z0 = (r < nb) ? lds(ya + 8u * (j0 + r)) : 0.0, z1 = (r + 1 < nb) ? lds(ya + 8u * (j0 + r + 1)) : 0.0;
                    e = fma(lds(pl + 8u * (r * NB + c)), z0, e);
                    o = fma(lds(pl + 8u * ((r + 1) * NB + c)), z1, o);
                }
                double acc = e + o;
                acc += __shfl_xor_sync(0xffffffffu, acc, 16);
                __syncwarp();  // every lane has read its z before any lane overwrites one
                if (lane < nb) sts(ya + 8u * (j0 + lane), acc);
                __syncwarp();
            }
            if (blk >= 2) bar_arrive(3 + (blk & 1), nbar);  // y_b is out
            if (blk >= 1) {  // look ahead: rows of block b - 1 (a full block), eight columns of block b per half warp
                const typename M::hnd e = S.h(j0 - NB + c, j0 + hh * 8);
                double s0 = 0, s1 = 0;
#pragma unroll
                for (int u = 0; u < 8; u += 2) {
                    const int t = hh * 8 + u;
                    if (t < nb) s0 = fma(M::ld(M::right(e, u)), lds(ya + 8u * (j0 + t)), s0);
                    if (t + 1 < nb) s1 = fma(M::ld(M::right(e, u + 1)), lds(ya + 8u * (j0 + t + 1)), s1);
                }
                double sum = s0 + s1;
                sum += __shfl_xor_sync(0xffffffffu, sum, 16);
                corr = sum;
            }
        }
    } else if (!idle) {
        for (int blk = nblk - 1; blk >= 2; blk--) {
            const int j0 = blk * NB, nb = min(NB, D - j0);
            bar_sync(3 + (blk & 1), nbar);
            for (int i = wtid; i < j0 - NB; i += nworkers) {
                const typename M::hnd e = S.h(i, j0);
                double s0 = 0, s1 = 0;
#pragma unroll 4
                for (int t = 0; t < NB; t += 2) {
                    if (t < nb) s0 = fma(M::ld(M::right(e, t)), lds(ya + 8u * (j0 + t)), s0);
                    if (t + 1 < nb) s1 = fma(M::ld(M::right(e, t + 1)), lds(ya + 8u * (j0 + t + 1)), s1);
                }
                sts(ya + 8u * i, lds(ya + 8u * i) - (s0 + s1));
            }
            bar_arrive(5 + (blk & 1), nbar);
        }
    }
    __syncthreads();
}

#endif  // __CUDACC__

}  // namespace lfvio
