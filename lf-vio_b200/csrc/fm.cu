// FeatureManager on the device: the track list of the sliding window as a structure of arrays in HBM,
// its bookkeeping as kernels, and the projection-factor list of Estimator::optimization() built where it is consumed.
//
// Reference: vins_estimator/src/feature_manager.{h,cpp} — FeaturePerFrame (.h:17-38), FeaturePerId (.h:41-64),
// addFeatureCheckParallax (.cpp:45-95), getFeatureCount (:36-43), getDepthVector / setDepth / clearDepth (:145-197),
// removeFailures (:134-143), removeBackShiftDepth (:271-310), removeBack (:312-327), removeFront (:329-351),
// compensatedParallax2 (:353-384, LF-VIO's bearing form), and the factor insertion loop estimator.cpp:725-772.
//
// Layout.  A feature lives in a SLOT (0 .. capacity-1): id, start_frame, number of observations, estimated_depth, solve_flag,
// and a [window + 1] x 9 array of observations (point 3, velocity 3, uv 2, cur_td).  The reference's std::list order (insertion
// order; the tracker hands out ids in increasing order and std::map iterates them ascending, so the list is sorted by id) is
// the array `order` of live slots.  Every operation is ONE single-CTA kernel on the handle's stream:
//   add_frame       binary search of the incoming ids in the (id-sorted) list; new features take free slots in incoming order
//                   (block scans); observations appended; last_track_num and the parallax test reduced in the same launch.
//   remove_*        per list entry in place (a feature that loses its first observation shifts its rows), then a stable
//                   compaction of `order`.
//   set_depth       1 / x over the selected features in list order (block scan gives the index); remove_failures likewise.
//   build_factors   landmark index and factor offset of every selected feature by two block scans, then one thread per
//                   feature writes its lfvio_proj_factor records, its range and 1 / depth — the arrays lfvio_window consumes.
// Counts come back in one small pinned block (key-frame decision, list length, L, R, capacity error).
#include <cstring>
#include <vector>

#include "backend.cuh"

namespace lfvio {

namespace {

constexpr int FM_THREADS = 1024;
constexpr int OBS_W = 9;  // point 3, velocity 3, uv 2, cur_td

enum { M_NLIST = 0, M_TRACKED = 1, M_KEY = 2, M_ERR = 3, M_NLM = 4, M_NPROJ = 5, M_MASK = 6, M_NMARG = 7, M_COUNT = 8 };

struct FmDev {
    int W, NF, C;  // window_size, frames in the window, slots
    int *id, *start, *nobs, *flag, *order, *order2, *freel;
    double *depth, *obs;  // obs: C x NF x OBS_W
    int *meta;            // M_COUNT ints
};

// exclusive scan of one int per thread over the CTA; returns the exclusive prefix, *total the sum.  blockDim = FM_THREADS.
__device__ int block_excl_scan(int v, int *s_warp, int *total) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    int inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
    }
    __syncthreads();
    if (lane == 31) s_warp[w] = inc;
    __syncthreads();
    if (w == 0) {
        int t = s_warp[lane], ti = t;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int u = __shfl_up_sync(0xffffffffu, ti, o);
            if (lane >= o) ti += u;
        }
        s_warp[lane] = ti - t;
        if (lane == 31) s_warp[32] = ti;
    }
    __syncthreads();
    *total = s_warp[32];
    return s_warp[w] + inc - v;
}

__device__ __forceinline__ bool selected(const FmDev &f, int s) { return f.nobs[s] >= 2 && f.start[s] < f.W - 2; }
__device__ __forceinline__ double *obs_at(const FmDev &f, int s, int k) { return f.obs + ((size_t)s * f.NF + k) * OBS_W; }

// stable compaction of `order` keeping the entries whose slot is still live (id >= 0); single CTA
__device__ void compact_order(const FmDev &f, int n_list, int *s_warp) {
    int base = 0;
    for (int i0 = 0; i0 < n_list; i0 += FM_THREADS) {
        const int i = i0 + threadIdx.x;
        const int s = i < n_list ? f.order[i] : -1;
        const int keep = (s >= 0 && f.id[s] >= 0) ? 1 : 0;
        int total;
        const int pos = block_excl_scan(keep, s_warp, &total);
        if (keep) f.order2[base + pos] = s;
        base += total;
        __syncthreads();
    }
    for (int i = threadIdx.x; i < base; i += FM_THREADS) f.order[i] = f.order2[i];
    if (threadIdx.x == 0) f.meta[M_NLIST] = base;
    __syncthreads();
}

__global__ void __launch_bounds__(FM_THREADS) fm_clear_kernel(FmDev f) {
    for (int s = threadIdx.x; s < f.C; s += FM_THREADS) {
        f.id[s] = -1;
        f.nobs[s] = 0;
    }
    if (threadIdx.x < M_COUNT) f.meta[threadIdx.x] = 0;
}

// addFeatureCheckParallax.  in_ids ascending, in_feat n x 8 = [x y z u v vx vy vz].
__global__ void __launch_bounds__(FM_THREADS) fm_add_frame_kernel(FmDev f, int frame_count, const int *__restrict__ in_ids,
                                                                  const double *__restrict__ in_feat, int n, double td, double min_parallax) {
    __shared__ int s_warp[33];
    __shared__ double s_red[32];
    __shared__ int s_cnt[32];
    const int tid = threadIdx.x;
    const int n_list = f.meta[M_NLIST];
    // ---- free slots, in slot order ----
    int n_free = 0;
    for (int s0 = 0; s0 < f.C; s0 += FM_THREADS) {
        const int s = s0 + tid;
        const int fr = (s < f.C && f.id[s] < 0) ? 1 : 0;
        int total;
        const int pos = block_excl_scan(fr, s_warp, &total);
        if (fr) f.freel[n_free + pos] = s;
        n_free += total;
        __syncthreads();
    }
    // ---- match the incoming observations (the list is sorted by id: binary search), append, count ----
    int n_new = 0, tracked = 0;
    for (int k0 = 0; k0 < n; k0 += FM_THREADS) {
        const int k = k0 + tid;
        int slot = -1, is_new = 0;
        if (k < n) {
            const int want = in_ids[k];
            int lo = 0, hi = n_list - 1;
            while (lo <= hi) {
                const int mid = (lo + hi) >> 1, s = f.order[mid], v = f.id[s];
                if (v == want) {
                    slot = s;
                    break;
                }
                if (v < want) lo = mid + 1;
                else hi = mid - 1;
            }
            is_new = slot < 0;
            // the list stays sorted by id only if new ids exceed every id handed out before (the tracker's n_id++ does)
            if (is_new && n_list > 0 && want < f.id[f.order[n_list - 1]]) f.meta[M_ERR] = 4;
        }
        int total_new;
        const int r = block_excl_scan(is_new, s_warp, &total_new);
        if (k < n) {
            int j;
            if (is_new) {
                if (n_new + r < n_free) {
                    slot = f.freel[n_new + r];
                    f.id[slot] = in_ids[k];
                    f.start[slot] = frame_count;
                    f.nobs[slot] = 1;
                    f.depth[slot] = -1.0;  // FeaturePerId ctor (feature_manager.h:57-60)
                    f.flag[slot] = 0;
                    f.order[n_list + n_new + r] = slot;
                    j = 0;
                } else {
                    f.meta[M_ERR] = 1;  // more live features than slots
                    j = -1;
                }
            } else {
                j = f.nobs[slot];
                if (j < f.NF) f.nobs[slot] = j + 1;
                else {
                    f.meta[M_ERR] = 2;  // a track longer than the window: the caller did not slide
                    j = -1;
                }
            }
            if (j >= 0) {
                double *o = obs_at(f, slot, j);
                const double *p = in_feat + 8 * (size_t)k;
                o[0] = p[0]; o[1] = p[1]; o[2] = p[2];   // point
                o[3] = p[5]; o[4] = p[6]; o[5] = p[7];   // velocity
                o[6] = p[3]; o[7] = p[4];                // uv
                o[8] = td;                               // cur_td
            }
        }
        int t_total;
        block_excl_scan((k < n && !is_new) ? 1 : 0, s_warp, &t_total);
        tracked += t_total;
        n_new += min(total_new, max(n_free - n_new, 0));
        __syncthreads();
    }
    const int n_list2 = n_list + n_new;
    __syncthreads();
    // ---- parallax between the second last and the third last frame (feature_manager.cpp:74-94) ----
    double psum = 0;
    int pnum = 0;
    for (int i = tid; i < n_list2; i += FM_THREADS) {
        const int s = f.order[i];
        const int st = f.start[s], no = f.nobs[s];
        if (st <= frame_count - 2 && st + no - 1 >= frame_count - 1) {
            const double *pi = obs_at(f, s, frame_count - 2 - st), *pj = obs_at(f, s, frame_count - 1 - st);
            psum += acos(pi[0] * pj[0] + pi[1] * pj[1] + pi[2] * pj[2]) * 10;  // compensatedParallax2
            pnum++;
        }
    }
    psum = warp_sum(psum);
    pnum = warp_sum(pnum);
    if ((tid & 31) == 0) {
        s_red[tid >> 5] = psum;
        s_cnt[tid >> 5] = pnum;
    }
    __syncthreads();
    if (tid == 0) {
        double ps = 0;
        int pn = 0;
        for (int w = 0; w < FM_THREADS / 32; w++) {
            ps += s_red[w];
            pn += s_cnt[w];
        }
        int key;
        if (frame_count < 2 || tracked < 20) key = 1;
        else if (pn == 0) key = 1;
        else key = (ps / pn >= min_parallax) ? 1 : 0;
        f.meta[M_NLIST] = n_list2;
        f.meta[M_TRACKED] = tracked;
        f.meta[M_KEY] = key;
    }
}

// op 0: removeBack, 1: removeBackShiftDepth (pose = marg_R 9 | marg_P 3 | new_R 9 | new_P 3), 2: removeFront(frame_count),
// 3: removeFailures
struct SlidePose {
    double v[24];
};
__global__ void __launch_bounds__(FM_THREADS) fm_slide_kernel(FmDev f, int op, int frame_count, SlidePose ps, double init_depth) {
    __shared__ int s_warp[33];
    const int n_list = f.meta[M_NLIST];
    for (int i = threadIdx.x; i < n_list; i += FM_THREADS) {
        const int s = f.order[i];
        if (op == 3) {
            if (f.flag[s] == 2) f.id[s] = -1, f.nobs[s] = 0;
            continue;
        }
        if (op == 2) {
            if (f.start[s] == frame_count) {
                f.start[s]--;
                continue;
            }
            const int j = f.W - 1 - f.start[s];
            if (f.start[s] + f.nobs[s] - 1 < frame_count - 1) continue;
            for (int k = j; k + 1 < f.nobs[s]; k++) {
                const double *src = obs_at(f, s, k + 1);
                double *dst = obs_at(f, s, k);
                for (int q = 0; q < OBS_W; q++) dst[q] = src[q];
            }
            if (--f.nobs[s] == 0) f.id[s] = -1;
            continue;
        }
        if (f.start[s] != 0) {
            f.start[s]--;
            continue;
        }
        const double u0 = obs_at(f, s, 0)[0], u1 = obs_at(f, s, 0)[1], u2 = obs_at(f, s, 0)[2];
        for (int k = 0; k + 1 < f.nobs[s]; k++) {
            const double *src = obs_at(f, s, k + 1);
            double *dst = obs_at(f, s, k);
            for (int q = 0; q < OBS_W; q++) dst[q] = src[q];
        }
        const int no = --f.nobs[s];
        if (op == 0) {
            if (no == 0) f.id[s] = -1;
        } else {
            if (no < 2) {
                f.id[s] = -1;
                f.nobs[s] = 0;
            } else {
                const double *mR = ps.v, *mP = ps.v + 9, *nR = ps.v + 12, *nP = ps.v + 21;
                const double d = f.depth[s];
                const double p0 = u0 * d, p1 = u1 * d, p2 = u2 * d;
                const double w0 = mR[0] * p0 + mR[1] * p1 + mR[2] * p2, w1 = mR[3] * p0 + mR[4] * p1 + mR[5] * p2,
                             w2 = mR[6] * p0 + mR[7] * p1 + mR[8] * p2;
                const double d0 = w0 + mP[0] - nP[0], d1 = w1 + mP[1] - nP[1], d2 = w2 + mP[2] - nP[2];
                const double q0 = nR[0] * d0 + nR[3] * d1 + nR[6] * d2, q1 = nR[1] * d0 + nR[4] * d1 + nR[7] * d2,
                             q2 = nR[2] * d0 + nR[5] * d1 + nR[8] * d2;  // new_R^T d
                const double dep_j = sqrt(q0 * q0 + q1 * q1 + q2 * q2);
                f.depth[s] = dep_j > 0 ? dep_j : init_depth;
            }
        }
    }
    __syncthreads();
    compact_order(f, n_list, s_warp);
}

// setDepth (mode 0: depth = 1 / x, solve_flag = 1) / clearDepth (mode 1: depth = 1 / x) / getDepthVector (mode 2: x = 1 / depth)
__global__ void __launch_bounds__(FM_THREADS) fm_depth_kernel(FmDev f, double *x, int mode) {
    __shared__ int s_warp[33];
    const int n_list = f.meta[M_NLIST];
    int base = 0;
    for (int i0 = 0; i0 < n_list; i0 += FM_THREADS) {
        const int i = i0 + threadIdx.x;
        const int s = i < n_list ? f.order[i] : -1;
        const int sel = (s >= 0 && selected(f, s)) ? 1 : 0;
        int total;
        const int idx = base + block_excl_scan(sel, s_warp, &total);
        if (sel) {
            if (mode == 2) x[idx] = 1.0 / f.depth[s];
            else {
                f.depth[s] = 1.0 / x[idx];
                if (mode == 0) f.flag[s] = 1;  // both branches of the reference set 1 (feature_manager.cpp:151-157)
            }
        }
        base += total;
        __syncthreads();
    }
    if (threadIdx.x == 0) f.meta[M_NLM] = base;
}

// The residual blocks of estimator.cpp:725-772 in the reference's insertion order + para_Feature (vector2double, :1044-1046).
// Also what the marginalisation plan needs (estimator.cpp:862-899): the landmarks hosted in frame 0 and the frames they are seen from.
__global__ void __launch_bounds__(FM_THREADS) fm_build_factors_kernel(FmDev f, FmBuildDst dst) {
    __shared__ int s_warp[33];
    __shared__ unsigned s_mask;
    const int n_list = f.meta[M_NLIST];
    int lm_base = 0, k_base = 0, m_base = 0;
    if (threadIdx.x == 0) s_mask = 0;
    __syncthreads();
    for (int i0 = 0; i0 < n_list; i0 += FM_THREADS) {
        const int i = i0 + threadIdx.x;
        const int s = i < n_list ? f.order[i] : -1;
        const int sel = (s >= 0 && selected(f, s)) ? 1 : 0;
        const int nfac = sel ? f.nobs[s] - 1 : 0;
        const int host0 = (sel && f.start[s] == 0) ? 1 : 0;
        int tot_l, tot_k, tot_m;
        const int l = lm_base + block_excl_scan(sel, s_warp, &tot_l);
        const int k0 = k_base + block_excl_scan(nfac, s_warp, &tot_k);
        const int m = m_base + block_excl_scan(host0, s_warp, &tot_m);
        if (sel) {
            if (l < dst.lm_cap && k0 + nfac <= dst.proj_cap) {
                dst.lm_first[l] = k0;
                dst.para_feature[l] = 1.0 / f.depth[s];
                f.order2[l] = s;  // landmark -> slot, for the factor pass below (order2 is scratch outside the slides)
                if (host0) {
                    dst.marg_lm[m] = l;
                    atomicOr(&s_mask, ((2u << nfac) - 1u) & ~1u);  // frames 1 .. nfac
                }
            } else {
                f.meta[M_ERR] = 3;  // factor list larger than the handle's capacity
            }
        }
        lm_base += tot_l;
        k_base += tot_k;
        m_base += tot_m;
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        dst.lm_first[min(lm_base, dst.lm_cap)] = k_base;
        f.meta[M_NLM] = lm_base;
        f.meta[M_NPROJ] = k_base;
        f.meta[M_MASK] = (int)s_mask;
        f.meta[M_NMARG] = m_base;
    }
    __syncthreads();
    if (f.meta[M_ERR] != 0) return;
    // the records, one thread per FACTOR (consecutive threads write consecutive 144-byte records): its landmark is the last
    // range that starts at or before it
    for (int k = threadIdx.x; k < k_base; k += FM_THREADS) {
        int lo = 0, hi = lm_base - 1;
        while (lo < hi) {
            const int mid = (lo + hi + 1) >> 1;
            if (dst.lm_first[mid] <= k) lo = mid;
            else hi = mid - 1;
        }
        const int l = lo, s = f.order2[l], j = k - dst.lm_first[l] + 1, fi = f.start[s];
        const double *o0 = obs_at(f, s, 0), *oj = obs_at(f, s, j);
        lfvio_proj_factor r;
        r.frame_i = fi;
        r.frame_j = fi + j;
        r.landmark = l;
        r.reserved = 0;
        for (int q = 0; q < 3; q++) {
            r.pts_i[q] = o0[q];
            r.pts_j[q] = oj[q];
            r.vel_i[q] = o0[3 + q];
            r.vel_j[q] = oj[3 + q];
        }
        r.td_i = o0[8];
        r.td_j = oj[8];
        r.row_i = o0[7];
        r.row_j = oj[7];
        dst.proj[k] = r;
    }
}

// test hook: the list as (feature_id, start_frame, n_obs) + depth, in list order
__global__ void __launch_bounds__(FM_THREADS) fm_snapshot_kernel(FmDev f, int *meta3, double *depth, int cap) {
    const int n_list = f.meta[M_NLIST];
    for (int i = threadIdx.x; i < n_list && i < cap; i += FM_THREADS) {
        const int s = f.order[i];
        meta3[3 * i] = f.id[s];
        meta3[3 * i + 1] = f.start[s];
        meta3[3 * i + 2] = f.nobs[s];
        depth[i] = f.depth[s];
    }
}

template <class T>
struct Buf {
    T *p = nullptr;
    ~Buf() {
        if (p) cudaFree(p);
    }
    int alloc(size_t n) {
        LFVIO_CUDA_TRY(cudaMalloc(&p, sizeof(T) * (n ? n : 1)));
        LFVIO_CUDA_TRY(cudaMemset(p, 0, sizeof(T) * (n ? n : 1)));
        return LFVIO_OK;
    }
};

}  // namespace

}  // namespace lfvio

using namespace lfvio;

struct lfvio_fm {
    int device = 0, W = 0, C = 0, in_cap = 0;
    double min_parallax = 0, init_depth = 5.0;
    cudaStream_t stream = nullptr;
    FmDev d{};
    Buf<int> id, start, nobs, flag, order, order2, freel, meta, lm_first, marg_lm, in_ids, snap_meta;
    int proj_cap = 0;
    Buf<double> depth, obs, para_feature, in_feat, x, snap_depth, poses;
    Buf<lfvio_proj_factor> proj;
    int *h_meta = nullptr;      // pinned
    double *h_stage = nullptr;  // pinned staging of one incoming frame / depth vector
    size_t h_stage_bytes = 0;
    ~lfvio_fm() {
        if (h_meta) cudaFreeHost(h_meta);
        if (h_stage) cudaFreeHost(h_stage);
        if (stream) cudaStreamDestroy(stream);
    }
};

static int fm_meta(lfvio_fm *f) {  // counts back to the host: one small copy + one synchronisation of the handle's stream
    LFVIO_CUDA_TRY(cudaMemcpyAsync(f->h_meta, f->d.meta, sizeof(int) * M_COUNT, cudaMemcpyDeviceToHost, f->stream));
    LFVIO_CUDA_TRY(cudaStreamSynchronize(f->stream));
    if (f->h_meta[M_ERR] == 1) {
        set_error("lfvio_fm: more live features than slots");
        return LFVIO_ERR_CAPACITY;
    }
    if (f->h_meta[M_ERR] == 2) {
        set_error("lfvio_fm: a track is longer than the window (slide before adding the next frame)");
        return LFVIO_ERR_ARG;
    }
    if (f->h_meta[M_ERR] == 4) {
        set_error("lfvio_fm: a new feature id is smaller than one already in the list (ids must be handed out in increasing order)");
        return LFVIO_ERR_ARG;
    }
    if (f->h_meta[M_ERR] == 3) {
        set_error("lfvio_fm: factor list larger than the handle's capacity");
        return LFVIO_ERR_CAPACITY;
    }
    return LFVIO_OK;
}

extern "C" int lfvio_fm_create(int window_size, int capacity, int max_proj_factors, double min_parallax, double init_depth, int device,
                               lfvio_fm **out) {
    clear_error();
    if (out) *out = nullptr;
    LFVIO_REQUIRE(out && window_size >= 2 && window_size + 1 <= 32 && capacity > 0 && max_proj_factors > 0, "lfvio_fm_create: bad arguments");
    LFVIO_CUDA_TRY(cudaSetDevice(device));
    lfvio_fm *f = new lfvio_fm();
    f->device = device;
    f->W = window_size;
    f->C = capacity;
    f->min_parallax = min_parallax;
    f->init_depth = init_depth;
    f->in_cap = capacity;
    auto fail = [&](int st) {
        delete f;
        return st;
    };
    int st;
#define FM_ALLOC(b, n) if ((st = f->b.alloc(n)) != LFVIO_OK) return fail(st)
    FM_ALLOC(id, capacity);
    FM_ALLOC(start, capacity);
    FM_ALLOC(nobs, capacity);
    FM_ALLOC(flag, capacity);
    FM_ALLOC(order, capacity);
    FM_ALLOC(order2, capacity);
    FM_ALLOC(freel, capacity);
    FM_ALLOC(meta, M_COUNT);
    FM_ALLOC(depth, capacity);
    FM_ALLOC(obs, (size_t)capacity * (window_size + 1) * OBS_W);
    FM_ALLOC(proj, max_proj_factors);
    FM_ALLOC(lm_first, capacity + 1);
    FM_ALLOC(marg_lm, capacity);
    FM_ALLOC(para_feature, capacity);
    FM_ALLOC(in_ids, capacity);
    FM_ALLOC(in_feat, (size_t)capacity * 8);
    FM_ALLOC(x, capacity);
    FM_ALLOC(snap_meta, (size_t)capacity * 3);
    FM_ALLOC(snap_depth, capacity);
    FM_ALLOC(poses, 12 * (size_t)(window_size + 1) + 12);
#undef FM_ALLOC
    if (cudaStreamCreateWithFlags(&f->stream, cudaStreamNonBlocking) != cudaSuccess || cudaMallocHost(&f->h_meta, sizeof(int) * M_COUNT) != cudaSuccess) {
        set_error("lfvio_fm_create: stream / pinned allocation failed");
        return fail(LFVIO_ERR_CUDA);
    }
    f->h_stage_bytes = sizeof(double) * ((size_t)capacity * 9 + 64 + 12 * (size_t)(window_size + 2));
    if (cudaMallocHost(&f->h_stage, f->h_stage_bytes) != cudaSuccess) {
        set_error("lfvio_fm_create: pinned allocation failed");
        return fail(LFVIO_ERR_CUDA);
    }
    FmDev &d = f->d;
    d.W = window_size;
    d.NF = window_size + 1;
    d.C = capacity;
    d.id = f->id.p; d.start = f->start.p; d.nobs = f->nobs.p; d.flag = f->flag.p; d.order = f->order.p; d.order2 = f->order2.p;
    d.freel = f->freel.p; d.depth = f->depth.p; d.obs = f->obs.p; d.meta = f->meta.p;
    f->proj_cap = max_proj_factors;
    fm_clear_kernel<<<1, FM_THREADS, 0, f->stream>>>(d);
    count_launch();
    if (cudaPeekAtLastError() != cudaSuccess || cudaStreamSynchronize(f->stream) != cudaSuccess) {
        set_error("lfvio_fm_create: clearing the store failed: %s", cudaGetErrorString(cudaGetLastError()));
        return fail(LFVIO_ERR_CUDA);
    }
    *out = f;
    return LFVIO_OK;
}

extern "C" void lfvio_fm_destroy(lfvio_fm *f) {
    if (!f) return;
    cudaSetDevice(f->device);
    if (f->stream) cudaStreamSynchronize(f->stream);
    delete f;
}

extern "C" int lfvio_fm_clear(lfvio_fm *f) {  // FeatureManager::clearState (feature_manager.cpp:31-34)
    clear_error();
    LFVIO_REQUIRE(f, "lfvio_fm_clear: null handle");
    LFVIO_CUDA_TRY(cudaSetDevice(f->device));
    fm_clear_kernel<<<1, FM_THREADS, 0, f->stream>>>(f->d);
    LFVIO_LAUNCH_CHECK();
    return LFVIO_OK;
}

// Replace the content of the store (the host-side list at the moment the estimator leaves its initialisation): feature i of the
// list goes to slot i.
extern "C" int lfvio_fm_load(lfvio_fm *f, int n, const int *ids, const int *start_frame, const int *obs_first, const double *depth,
                             const int *solve_flag, const double *obs9) {
    clear_error();
    LFVIO_REQUIRE(f && n >= 0 && n <= f->C && (n == 0 || (ids && start_frame && obs_first && depth && obs9)), "lfvio_fm_load: bad arguments");
    const int NF = f->W + 1;
    std::vector<int> id(f->C, -1), st(f->C, 0), no(f->C, 0), fl(f->C, 0), ord(f->C, 0);
    std::vector<double> dep(f->C, -1.0), obs((size_t)(n ? n : 1) * NF * OBS_W, 0.0);
    for (int i = 0; i < n; i++) {
        const int k = obs_first[i + 1] - obs_first[i];
        LFVIO_REQUIRE(k >= 1 && k <= NF && start_frame[i] >= 0 && start_frame[i] + k <= NF, "lfvio_fm_load: a track leaves the window");
        LFVIO_REQUIRE(i == 0 || ids[i] > ids[i - 1], "lfvio_fm_load: ids must ascend");
        id[i] = ids[i];
        st[i] = start_frame[i];
        no[i] = k;
        fl[i] = solve_flag ? solve_flag[i] : 0;
        ord[i] = i;
        dep[i] = depth[i];
        memcpy(&obs[(size_t)i * NF * OBS_W], obs9 + (size_t)obs_first[i] * OBS_W, sizeof(double) * OBS_W * k);
    }
    int meta[M_COUNT] = {0};
    meta[M_NLIST] = n;
    LFVIO_CUDA_TRY(cudaSetDevice(f->device));
    cudaStream_t s = f->stream;
    LFVIO_CUDA_TRY(cudaMemcpyAsync(f->d.id, id.data(), sizeof(int) * f->C, cudaMemcpyHostToDevice, s));
    LFVIO_CUDA_TRY(cudaMemcpyAsync(f->d.start, st.data(), sizeof(int) * f->C, cudaMemcpyHostToDevice, s));
    LFVIO_CUDA_TRY(cudaMemcpyAsync(f->d.nobs, no.data(), sizeof(int) * f->C, cudaMemcpyHostToDevice, s));
    LFVIO_CUDA_TRY(cudaMemcpyAsync(f->d.flag, fl.data(), sizeof(int) * f->C, cudaMemcpyHostToDevice, s));
    LFVIO_CUDA_TRY(cudaMemcpyAsync(f->d.order, ord.data(), sizeof(int) * f->C, cudaMemcpyHostToDevice, s));
    LFVIO_CUDA_TRY(cudaMemcpyAsync(f->d.depth, dep.data(), sizeof(double) * f->C, cudaMemcpyHostToDevice, s));
    if (n) LFVIO_CUDA_TRY(cudaMemcpyAsync(f->d.obs, obs.data(), sizeof(double) * (size_t)n * NF * OBS_W, cudaMemcpyHostToDevice, s));
    LFVIO_CUDA_TRY(cudaMemcpyAsync(f->d.meta, meta, sizeof(meta), cudaMemcpyHostToDevice, s));
    LFVIO_CUDA_TRY(cudaStreamSynchronize(s));  // (the sources are stack / heap buffers of this call)
    return LFVIO_OK;
}

extern "C" int lfvio_fm_add_frame(lfvio_fm *f, int frame_count, const int *ids, const double *feat8, int n, double td, int *is_keyframe,
                                  int *last_track_num) {
    clear_error();
    LFVIO_REQUIRE(f && n >= 0 && n <= f->in_cap && (n == 0 || (ids && feat8)) && frame_count >= 0 && frame_count <= f->W, "lfvio_fm_add_frame: bad arguments");
    for (int k = 1; k < n; k++) LFVIO_REQUIRE(ids[k] > ids[k - 1], "lfvio_fm_add_frame: ids must ascend (the reference iterates a std::map)");
    LFVIO_CUDA_TRY(cudaSetDevice(f->device));
    // (the previous call synchronised the stream before it returned, so the staging block is free)
    int *hs_ids = reinterpret_cast<int *>(f->h_stage + (size_t)f->in_cap * 8);
    if (n) {
        memcpy(f->h_stage, feat8, sizeof(double) * 8 * (size_t)n);
        memcpy(hs_ids, ids, sizeof(int) * (size_t)n);
        LFVIO_CUDA_TRY(cudaMemcpyAsync(f->in_feat.p, f->h_stage, sizeof(double) * 8 * (size_t)n, cudaMemcpyHostToDevice, f->stream));
        LFVIO_CUDA_TRY(cudaMemcpyAsync(f->in_ids.p, hs_ids, sizeof(int) * (size_t)n, cudaMemcpyHostToDevice, f->stream));
    }
    fm_add_frame_kernel<<<1, FM_THREADS, 0, f->stream>>>(f->d, frame_count, f->in_ids.p, f->in_feat.p, n, td, f->min_parallax);
    LFVIO_LAUNCH_CHECK();
    LFVIO_TRY(fm_meta(f));
    if (is_keyframe) *is_keyframe = f->h_meta[M_KEY];
    if (last_track_num) *last_track_num = f->h_meta[M_TRACKED];
    return LFVIO_OK;
}

static int fm_slide(lfvio_fm *f, int op, int frame_count, const double *pose24) {
    LFVIO_CUDA_TRY(cudaSetDevice(f->device));
    SlidePose ps;
    memset(&ps, 0, sizeof(ps));
    if (pose24) memcpy(ps.v, pose24, sizeof(ps.v));
    fm_slide_kernel<<<1, FM_THREADS, 0, f->stream>>>(f->d, op, frame_count, ps, f->init_depth);
    LFVIO_LAUNCH_CHECK();
    return LFVIO_OK;
}
extern "C" int lfvio_fm_remove_back(lfvio_fm *f) {
    clear_error();
    LFVIO_REQUIRE(f, "lfvio_fm_remove_back: null handle");
    return fm_slide(f, 0, 0, nullptr);
}
extern "C" int lfvio_fm_remove_back_shift_depth(lfvio_fm *f, const double *marg_R, const double *marg_P, const double *new_R, const double *new_P) {
    clear_error();
    LFVIO_REQUIRE(f && marg_R && marg_P && new_R && new_P, "lfvio_fm_remove_back_shift_depth: null argument");
    double p[24];
    memcpy(p, marg_R, 72);
    memcpy(p + 9, marg_P, 24);
    memcpy(p + 12, new_R, 72);
    memcpy(p + 21, new_P, 24);
    return fm_slide(f, 1, 0, p);
}
extern "C" int lfvio_fm_remove_front(lfvio_fm *f, int frame_count) {
    clear_error();
    LFVIO_REQUIRE(f, "lfvio_fm_remove_front: null handle");
    return fm_slide(f, 2, frame_count, nullptr);
}
extern "C" int lfvio_fm_remove_failures(lfvio_fm *f) {
    clear_error();
    LFVIO_REQUIRE(f, "lfvio_fm_remove_failures: null handle");
    return fm_slide(f, 3, 0, nullptr);
}

// mode 0 setDepth, 1 clearDepth: x (n values, list order over the selected features) in; mode 2 getDepthVector: x out, returns n
extern "C" int lfvio_fm_depth(lfvio_fm *f, int mode, double *x, int n_cap, int *n_out) {
    clear_error();
    LFVIO_REQUIRE(f && x && mode >= 0 && mode <= 2 && n_cap >= 0 && n_cap <= f->C, "lfvio_fm_depth: bad arguments");
    LFVIO_CUDA_TRY(cudaSetDevice(f->device));
    if (mode != 2 && n_cap) {
        memcpy(f->h_stage, x, sizeof(double) * n_cap);
        LFVIO_CUDA_TRY(cudaMemcpyAsync(f->x.p, f->h_stage, sizeof(double) * n_cap, cudaMemcpyHostToDevice, f->stream));
    }
    fm_depth_kernel<<<1, FM_THREADS, 0, f->stream>>>(f->d, f->x.p, mode);
    LFVIO_LAUNCH_CHECK();
    if (mode == 2 && n_cap) LFVIO_CUDA_TRY(cudaMemcpyAsync(f->h_stage, f->x.p, sizeof(double) * n_cap, cudaMemcpyDeviceToHost, f->stream));
    LFVIO_TRY(fm_meta(f));
    const int n = f->h_meta[M_NLM];
    if (n_out) *n_out = n;
    if (mode == 2) memcpy(x, f->h_stage, sizeof(double) * (size_t)(n < n_cap ? n : n_cap));
    else LFVIO_REQUIRE(n <= n_cap, "lfvio_fm_depth: fewer values than selected features");
    return LFVIO_OK;
}

namespace lfvio {
int fm_build_factors_into(lfvio_fm *f, const FmBuildDst &dst, DevFactorSrc *src) {
    LFVIO_CUDA_TRY(cudaSetDevice(f->device));
    fm_build_factors_kernel<<<1, FM_THREADS, 0, f->stream>>>(f->d, dst);
    LFVIO_LAUNCH_CHECK();
    LFVIO_TRY(fm_meta(f));
    src->n_proj = f->h_meta[M_NPROJ];
    src->n_landmarks = f->h_meta[M_NLM];
    src->frame_j_mask = (unsigned)f->h_meta[M_MASK];
    src->n_marg_lm = f->h_meta[M_NMARG];
    return LFVIO_OK;
}
int fm_set_depth_from_dev(lfvio_fm *f, const double *x_dev) {  // the producer of x_dev has been synchronised by the caller
    LFVIO_CUDA_TRY(cudaSetDevice(f->device));
    fm_depth_kernel<<<1, FM_THREADS, 0, f->stream>>>(f->d, const_cast<double *>(x_dev), 0);
    LFVIO_LAUNCH_CHECK();
    return LFVIO_OK;
}
int fm_window_size(const lfvio_fm *f) { return f->W; }
int fm_device(const lfvio_fm *f) { return f->device; }
}  // namespace lfvio

extern "C" int lfvio_fm_build_factors(lfvio_fm *f, int *n_proj, int *n_landmarks) {
    clear_error();
    LFVIO_REQUIRE(f, "lfvio_fm_build_factors: null handle");
    FmBuildDst dst{f->proj.p, f->lm_first.p, f->para_feature.p, f->marg_lm.p, f->proj_cap, f->C};
    DevFactorSrc src;
    LFVIO_TRY(fm_build_factors_into(f, dst, &src));
    if (n_proj) *n_proj = src.n_proj;
    if (n_landmarks) *n_landmarks = src.n_landmarks;
    return LFVIO_OK;
}

// device views of the last build (consumed by lfvio_window_optimize_full_fm): valid until the next call on this handle
// FeatureManager::triangulate (feature_manager.cpp:199-253) over the slots of the store: Ps (F x 3), Rs (F x 9 row-major),
// tic (3), ric (9) from the host, depths in place on the device.
extern "C" int lfvio_fm_triangulate(lfvio_fm *f, const double *Ps, const double *Rs, const double *tic, const double *ric) {
    clear_error();
    LFVIO_REQUIRE(f && Ps && Rs && tic && ric, "lfvio_fm_triangulate: null argument");
    LFVIO_CUDA_TRY(cudaSetDevice(f->device));
    LFVIO_CUDA_TRY(cudaStreamSynchronize(f->stream));  // the staging block may still be in flight (slides do not synchronise)
    const int F = f->W + 1;
    double *h = f->h_stage;
    memcpy(h, Ps, sizeof(double) * 3 * F);
    memcpy(h + 3 * F, Rs, sizeof(double) * 9 * F);
    memcpy(h + 12 * F, tic, sizeof(double) * 3);
    memcpy(h + 12 * F + 3, ric, sizeof(double) * 9);
    LFVIO_CUDA_TRY(cudaMemcpyAsync(f->poses.p, h, sizeof(double) * (12 * (size_t)F + 12), cudaMemcpyHostToDevice, f->stream));
    const double *d = f->poses.p;
    return launch_triangulate_slots(d, d + 3 * F, d + 12 * F, d + 12 * F + 3, f->d.start, f->d.nobs, f->d.obs, OBS_W, f->C, F, f->init_depth,
                                    f->d.depth, f->stream);
}

extern "C" int lfvio_fm_factors_dev(lfvio_fm *f, const lfvio_proj_factor **proj, const int **lm_first, const double **para_feature, void **stream) {
    clear_error();
    LFVIO_REQUIRE(f, "lfvio_fm_factors_dev: null handle");
    if (proj) *proj = f->proj.p;
    if (lm_first) *lm_first = f->lm_first.p;
    if (para_feature) *para_feature = f->para_feature.p;
    if (stream) *stream = f->stream;
    return LFVIO_OK;
}

// test hooks: the list (feature_id, start_frame, n_obs, depth) in list order; the factor list of the last build
extern "C" int lfvio_fm_snapshot(lfvio_fm *f, int *meta3, double *depth, int cap) {
    clear_error();
    LFVIO_REQUIRE(f && meta3 && depth && cap >= 0, "lfvio_fm_snapshot: bad arguments");
    LFVIO_CUDA_TRY(cudaSetDevice(f->device));
    const int c = cap < f->C ? cap : f->C;
    fm_snapshot_kernel<<<1, FM_THREADS, 0, f->stream>>>(f->d, f->snap_meta.p, f->snap_depth.p, c);
    LFVIO_LAUNCH_CHECK();
    LFVIO_TRY(fm_meta(f));
    const int n = f->h_meta[M_NLIST] < c ? f->h_meta[M_NLIST] : c;
    LFVIO_CUDA_TRY(cudaMemcpy(meta3, f->snap_meta.p, sizeof(int) * 3 * (size_t)n, cudaMemcpyDeviceToHost));
    LFVIO_CUDA_TRY(cudaMemcpy(depth, f->snap_depth.p, sizeof(double) * (size_t)n, cudaMemcpyDeviceToHost));
    return f->h_meta[M_NLIST];
}
extern "C" int lfvio_fm_download_factors(lfvio_fm *f, lfvio_proj_factor *proj, int proj_cap, int *lm_first, double *para_feature, int lm_cap) {
    clear_error();
    LFVIO_REQUIRE(f && proj && lm_first && para_feature, "lfvio_fm_download_factors: null argument");
    LFVIO_CUDA_TRY(cudaSetDevice(f->device));
    LFVIO_CUDA_TRY(cudaStreamSynchronize(f->stream));
    const int R = f->h_meta[M_NPROJ], L = f->h_meta[M_NLM];
    LFVIO_REQUIRE(R <= proj_cap && L <= lm_cap, "lfvio_fm_download_factors: output arrays too small");
    LFVIO_CUDA_TRY(cudaMemcpy(proj, f->proj.p, sizeof(lfvio_proj_factor) * (size_t)R, cudaMemcpyDeviceToHost));
    LFVIO_CUDA_TRY(cudaMemcpy(lm_first, f->lm_first.p, sizeof(int) * ((size_t)L + 1), cudaMemcpyDeviceToHost));
    LFVIO_CUDA_TRY(cudaMemcpy(para_feature, f->para_feature.p, sizeof(double) * (size_t)L, cudaMemcpyDeviceToHost));
    return LFVIO_OK;
}
