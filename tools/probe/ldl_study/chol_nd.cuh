// Nested dissection of the speed-bias chain in the reduced system (one CTA), in front of the dense blocked Cholesky.
//
// Column order of the reduced system (StateLayout): [sb_{F-1} .. sb_1 | sb_0 | pose_{F-1} .. pose_0 | ex | td].  A speed-bias block
// sb_k (k >= 1) is coupled only to sb_{k-1}, sb_{k+1} and pose_{k-1 .. k+1} (the two IMU factors that touch it); the prior and the
// landmarks' Schur complement make everything from sb_0 on dense.  Eliminating sb_{F-1}, sb_{F-2}, ... one after the other is a
// chain of 9 (F - 1) dependent pivots.  Here they go in LEVELS instead: level m eliminates every sb_k with k = (2 i + 1) 2^m — no
// two of them are coupled, their neighbours are sb_{k - 2^m}, sb_{k + 2^m} and pose_{k - 2^m .. k + 2^m} — so one warp per block
// factors them side by side and the chain shrinks to 9 ceil(log2 F) pivots (36 instead of 90 at F = 11, 45 instead of 180 at
// F = 21).  What remains, [sb_0 | poses | ex | td], is dense and goes through cholesky_lookahead (chol_blocked.cuh).
// The factor is the same Cholesky factor up to the elimination order, i.e. the same solution of the same linear system.
#pragma once
#include "chol_blocked.cuh"

namespace lfvio {

#ifdef __CUDACC__
constexpr int SBW = 9;               // columns of one speed-bias block
constexpr int ND_SLOT = SBW * SBW + SBW;  // per block: U (9 x 9, row-major, zeros below the diagonal) | reciprocal pivots

// trailing principal sub-matrix [off, D) of an upper-stored symmetric matrix
template <class M>
struct MatSub {
    M m;
    int off;
    __device__ __forceinline__ double &at(int r, int c) const { return m.at(r + off, c + off); }
    __device__ __forceinline__ double *row(int r) const { return m.row(r + off) + off; }
};

__host__ __device__ inline int nd_max_blocks(int F) { return F / 2; }  // blocks of level 0: k = 1, 3, ... <= F - 1
// staging doubles: panel columns (neighbours + right-hand side) of all blocks of one level, 9 each, + the factor slots
__host__ __device__ inline int nd_stage_doubles(int F) {
    const int a = nd_max_blocks(F) * (2 * SBW + 18 + 1), b = 2 * SBW + 6 * F + 1;
    return (a > b ? a : b) * SBW + nd_max_blocks(F) * ND_SLOT + ((a > b ? a : b) + 1) / 2;
}
__host__ __device__ inline int nd_items_max(int F) {
    const int a = nd_max_blocks(F) * (2 * SBW + 18 + 1), b = 2 * SBW + 6 * F + 1;
    return a > b ? a : b;
}

struct NdGeom {
    int F;
    __device__ __forceinline__ int nblocks(int h) const { return (h > F - 1) ? 0 : ((F - 1) / h + 1) / 2; }
    __device__ __forceinline__ int col_sb(int k) const { return SBW * (F - 1 - k); }
    __device__ __forceinline__ int col_pose(int f) const { return SBW * F + 6 * (F - 1 - f); }
};
struct NdBlock {  // sb_k eliminated at step h: neighbours sb_kr (kr < 0: none), sb_kl, pose_pl .. pose_pu
    int k, kl, kr, pl, pu, c0, n;
    __device__ __forceinline__ NdBlock(const NdGeom &g, int h, int slot) {
        k = (2 * slot + 1) * h;
        kl = k - h;
        kr = (k + h <= g.F - 1) ? k + h : -1;
        pl = kl;
        pu = kr >= 0 ? kr : g.F - 1;
        c0 = g.col_sb(k);
        n = (kr >= 0 ? SBW : 0) + SBW + 6 * (pu - pl + 1);
    }
    // neighbour columns in ascending order of their column index: [sb_kr][sb_kl][pose_pu .. pose_pl]
    __device__ __forceinline__ int col(const NdGeom &g, int j) const {
        if (kr >= 0) {
            if (j < SBW) return g.col_sb(kr) + j;
            j -= SBW;
        }
        if (j < SBW) return g.col_sb(kl) + j;
        return g.col_pose(pu) + (j - SBW);
    }
};
template <class M>
__device__ __forceinline__ double &nd_sym(const M &S, int r, int c) { return r <= c ? S.at(r, c) : S.at(c, r); }

__device__ __forceinline__ int nd_find_slot(const int *s_off, int nb, int q) {
    int slot = 0;
    while (slot + 1 < nb && q >= s_off[slot + 1]) slot++;
    return slot;
}

// Eliminates sb_{F-1} .. sb_1 level by level.  s_y: right-hand side in, z (forward-substituted) out for the eliminated columns,
// updated right-hand side for the rest.  Rows of U are left in S (the panel entries at their symmetric positions).
// s_stage: nd_stage_doubles(F).  s_off: 2 x 17 ints, s_flag: 1 int.  Whole CTA calls; returns false on a non-positive pivot.
template <class M>
__device__ bool nd_eliminate(M S, int F, double *s_y, double *s_stage, int *s_off, int *s_flag, long long *dbg_clk = nullptr) {
#define ND_MARK(i)                                                                   \
    do {                                                                             \
        if (dbg_clk && threadIdx.x == 0) {                                           \
            unsigned long long t_;                                                   \
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_));                    \
            dbg_clk[i] = (long long)t_;                                              \
        }                                                                            \
    } while (0)
    const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31, wid = tid >> 5;
    const NdGeom g{F};
    double *s_pan = s_stage, *s_u9 = s_stage + (size_t)nd_items_max(F) * SBW;
    int *s_colt = reinterpret_cast<int *>(s_u9 + nd_max_blocks(F) * ND_SLOT);  // global column of every panel column
    int *s_offc[2] = {s_off + 17, s_off + 34};  // stage-3 item offsets of the even / odd blocks
    if (tid == 0) *s_flag = 0;
    __syncthreads();
    ND_MARK(0);
    for (int h = 1; h <= F - 1; h <<= 1) {
        const int nb = g.nblocks(h);
        // ---- stage 1: one warp per block factors its 9 x 9 diagonal block in registers (lane = column) ----
        if (wid < nb) {
            const NdBlock B(g, h, wid);
            const int c = lane < SBW ? lane : SBW - 1;
            double a[SBW];
#pragma unroll
            for (int r = 0; r < SBW; r++) a[r] = (r <= c) ? S.at(B.c0 + r, B.c0 + c) : 0.0;
            double inv_mine = 0.0;
            bool ok = true;
#pragma unroll
            for (int j = 0; j < SBW; j++) {
                const double djj = __shfl_sync(0xffffffffu, a[j], j);
                ok = ok && (djj > 0.0) && (djj < 1e300);
                const double inv = rsqrt(djj);
                if (lane == j) inv_mine = inv;
                const double ujc = a[j] * inv;
                a[j] = ujc;
#pragma unroll
                for (int r = j + 1; r < SBW; r++) {
                    const double ujr = __shfl_sync(0xffffffffu, ujc, r);
                    a[r] = fma(-ujr, ujc, a[r]);
                }
            }
            double *U = s_u9 + wid * ND_SLOT;
            if (lane < SBW) {
#pragma unroll
                for (int r = 0; r < SBW; r++) {
                    const double v = (r <= c) ? a[r] : 0.0;
                    U[r * SBW + c] = v;
                    if (r <= c) S.at(B.c0 + r, B.c0 + c) = v;
                }
                U[SBW * SBW + lane] = inv_mine;
            }
            if (!ok && lane == 0) *s_flag = 1;
        }
        if (tid == 0) {  // item offsets of stages 2 and 3
            int o2 = 0, o3[2] = {0, 0};
            for (int b = 0; b < nb; b++) {
                const NdBlock B(g, h, b);
                s_off[b] = o2;
                s_offc[0][b] = o3[0];
                s_offc[1][b] = o3[1];
                o2 += B.n + 1;                                           // neighbour columns + right-hand side
                o3[b & 1] += ((B.n + 1) >> 1) * (B.n + 1);               // double rows x their n + 1 entries
            }
            s_off[nb] = o2;
            s_offc[0][nb] = o3[0];
            s_offc[1][nb] = o3[1];
        }
        __syncthreads();
        if (h == 1) ND_MARK(1);
        if (*s_flag) return false;
        // ---- stage 2: panel columns p_j = U^-T S[block, j] (and z = U^-T b), one thread each ----
        for (int q = tid; q < s_off[nb]; q += nt) {
            const int slot = nd_find_slot(s_off, nb, q), jl = q - s_off[slot];
            const NdBlock B(g, h, slot);
            const bool is_rhs = jl == B.n;
            const int j = is_rhs ? 0 : B.col(g, jl);
            const double *U = s_u9 + slot * ND_SLOT, *invd = U + SBW * SBW;
            double a[SBW];
#pragma unroll
            for (int r = 0; r < SBW; r++) a[r] = is_rhs ? s_y[B.c0 + r] : nd_sym(S, B.c0 + r, j);
#pragma unroll
            for (int jj = 0; jj < SBW; jj++) {
                const double u = a[jj] * invd[jj];
                a[jj] = u;
#pragma unroll
                for (int r = jj + 1; r < SBW; r++) a[r] = fma(-U[jj * SBW + r], u, a[r]);
            }
            double *pan = s_pan + (size_t)q * SBW;
            s_colt[q] = is_rhs ? -1 : j;
#pragma unroll
            for (int r = 0; r < SBW; r++) {
                pan[r] = a[r];
                if (is_rhs) s_y[B.c0 + r] = a[r];
                else nd_sym(S, B.c0 + r, j) = a[r];
            }
        }
        __syncthreads();
        if (h == 1) ND_MARK(2);
        // ---- stage 3: S[i, j] -= p_i . p_j over the neighbours of every block, b[j] -= p_j . z.  Neighbouring blocks of a level
        //      share a neighbour (sb_kr of one is sb_kl of the next): even blocks first, then odd blocks, plain read-modify-write.
        //      Rows r and n - 1 - r of a block's upper triangle form a double row of n + 1 entries; one item = 8 of them. ----
        for (int colour = 0; colour < 2; colour++) {
            const int *off3 = s_offc[colour];
            const int nitem = off3[nb];
            // One ENTRY per thread and pass, consecutive threads on consecutive columns j: the panel columns are 9 doubles
            // apart, an odd stride, so a warp's loads of p_j are bank-conflict free (8 entries per thread put the lanes 72
            // doubles apart: 8-way conflicts, measured 3x slower).  UNR passes are interleaved for instruction-level
            // parallelism (an fp64 FMA has a 23-cycle dependent latency on sm_100).
            constexpr int UNR = 4;
            for (int q0 = tid; q0 < nitem; q0 += UNR * nt) {
                int ii[UNR], jj[UNR], gi[UNR], gj[UNR];
                const double *pp[UNR];
                bool valid[UNR];
#pragma unroll
                for (int u = 0; u < UNR; u++) {
                    const int q = q0 + u * nt;
                    valid[u] = q < nitem;
                    int slot = colour;
                    while (slot + 2 < nb && q >= off3[slot + 2]) slot += 2;
                    const int n = s_off[slot + 1] - s_off[slot] - 1, wrow = n + 1;
                    const int ql = valid[u] ? q - off3[slot] : 0;
                    int r = (int)((float)ql * (1.0f / (float)wrow));
                    r += (ql - r * wrow >= wrow) - (ql - r * wrow < 0);
                    const int e = ql - r * wrow, r2 = n - 1 - r;
                    const bool first = e < n - r;
                    valid[u] = valid[u] && (first || r2 != r);
                    ii[u] = first ? r : r2;
                    jj[u] = first ? r + e : r2 + (e - (n - r));
                    if (!valid[u]) ii[u] = jj[u] = 0;
                    pp[u] = s_pan + (size_t)s_off[slot] * SBW;
                    const int *colt = s_colt + s_off[slot];
                    gi[u] = colt[ii[u]];
                    gj[u] = colt[jj[u]];
                }
                double d[UNR];
#pragma unroll
                for (int u = 0; u < UNR; u++) d[u] = 0.0;
#pragma unroll
                for (int t = 0; t < SBW; t++)
#pragma unroll
                    for (int u = 0; u < UNR; u++) d[u] = fma(pp[u][ii[u] * SBW + t], pp[u][jj[u] * SBW + t], d[u]);
#pragma unroll
                for (int u = 0; u < UNR; u++)
                    if (valid[u]) S.at(gi[u], gj[u]) -= d[u];
            }
            // right-hand side: b[j] -= p_j . z, by threads from the top of the CTA (they run next to the items above)
            for (int q = nt - 1 - tid; q < s_off[nb]; q += nt) {
                const int slot = nd_find_slot(s_off, nb, q);
                const int gj = s_colt[q];
                if ((slot & 1) == colour && gj >= 0) {
                    const double *pj = s_pan + (size_t)q * SBW, *z = s_pan + (size_t)(s_off[slot + 1] - 1) * SBW;
                    double d0 = 0, d1 = 0, d2 = 0;
#pragma unroll
                    for (int t = 0; t < SBW; t += 3) {
                        d0 = fma(pj[t], z[t], d0);
                        d1 = fma(pj[t + 1], z[t + 1], d1);
                        d2 = fma(pj[t + 2], z[t + 2], d2);
                    }
                    s_y[gj] -= (d0 + d1) + d2;
                }
            }
            __syncthreads();
        }
        if (h == 1) ND_MARK(3);
    }
    ND_MARK(4);
#undef ND_MARK
    return true;
}

// U y = z for the eliminated speed-bias blocks, levels in reverse; y of everything from sb_0 on is already in s_y.
template <class M>
__device__ void nd_backsolve(M S, int F, double *s_y) {
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const NdGeom g{F};
    int hmax = 1;
    while (hmax * 2 <= F - 1) hmax <<= 1;
    for (int h = hmax; h >= 1; h >>= 1) {
        const int nb = g.nblocks(h);
        if (wid < nb) {
            const NdBlock B(g, h, wid);
            double acc[SBW];
#pragma unroll
            for (int r = 0; r < SBW; r++) acc[r] = 0.0;
            for (int jl = lane; jl < B.n; jl += 32) {
                const int j = B.col(g, jl);
                const double yj = s_y[j];
#pragma unroll
                for (int r = 0; r < SBW; r++) acc[r] = fma(nd_sym(S, B.c0 + r, j), yj, acc[r]);
            }
            double mine = 0.0;
#pragma unroll
            for (int r = 0; r < SBW; r++) {
                const double t = warp_sum(acc[r]);
                if (lane == r) mine = t;
            }
            const int r = lane < SBW ? lane : SBW - 1;
            double urow[SBW];
#pragma unroll
            for (int c = 0; c < SBW; c++) urow[c] = (c > r) ? S.at(B.c0 + r, B.c0 + c) : 0.0;
            const double inv = 1.0 / S.at(B.c0 + r, B.c0 + r);
            double wv = s_y[B.c0 + r] - mine;
#pragma unroll
            for (int c = SBW - 1; c >= 0; c--) {
                const double yc = __shfl_sync(0xffffffffu, wv * inv, c);
                wv = fma(-urow[c], yc, wv);  // urow[c] == 0 for c <= r
            }
            if (lane < SBW) s_y[B.c0 + lane] = wv * inv;
        }
        __syncthreads();
    }
}
#endif

}  // namespace lfvio
