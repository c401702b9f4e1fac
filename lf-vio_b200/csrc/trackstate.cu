// K6 + track bookkeeping of FeatureTracker::readImage, on device (no host round trip inside a frame).
//
// Reference: feature_tracker/src/feature_tracker.cpp
//   reduceVector / track_cnt++            :14-40, :132-145   -> compact_kernel
//   setMask()                             :46-83             -> set_mask_select_kernel + mask_raster_kernel
//   addPoints()                           :85-93             -> fused into corners.cu select_kernel
//   undistortedPoints()                   :441-504           -> undistort_kernel
//   updateID() + node loop                :388-398, feature_tracker_node.cpp:103-111 -> update_ids_kernel
//
// Order-dependent semantics are preserved exactly:
//  * setMask sorts by track_cnt with libstdc++'s UNSTABLE std::sort; the permutation (hence which of two
//    equally old tracks survives and the output order, hence feature IDs) is reproduced by a device
//    transliteration of libstdc++'s introsort (median-of-3 quicksort, depth limit 2*lg n with heapsort
//    fallback, final insertion sort with threshold 16), run by one thread on a few hundred keys.
//  * the keep/stamp loop is sequential in that order; "pixel still 255" is evaluated geometrically
//    (annulus base mask AND outside every previously kept disc, using cv::circle's midpoint half-widths),
//    so no raster is needed inside the loop; the mask itself is rastered afterwards in parallel.
#include "frontend.cuh"

namespace lfvio {

void circle_half_widths(int radius, int *hw) {
    // cv::circle(img, c, radius, color, -1): midpoint algorithm filling horizontal spans (OpenCV drawing.cpp Circle())
    for (int i = 0; i <= radius; i++) hw[i] = -1;
    int err = 0, dx = radius, dy = 0, plus = 1, minus = (radius << 1) - 1;
    while (dx >= dy) {
        if (hw[dy] < dx) hw[dy] = dx;
        if (hw[dx] < dy) hw[dx] = dy;
        dy++;
        err += plus;
        plus += 2;
        const int m = (err <= 0) - 1;
        err -= minus & m;
        dx += m;
        minus -= m & 2;
    }
}

namespace {

constexpr int BT = 1024;

__device__ __forceinline__ int block_excl_scan(int v, int *s_warp, int *total) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    int inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        int t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
    }
    __syncthreads();
    if (lane == 31) s_warp[w] = inc;
    __syncthreads();
    if (w == 0) {
        int t = s_warp[lane], ti = t;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            int u = __shfl_up_sync(0xffffffffu, ti, o);
            if (lane >= o) ti += u;
        }
        s_warp[lane] = ti - t;
        if (lane == 31) s_warp[32] = ti;
    }
    __syncthreads();
    *total = s_warp[32];
    return inc - v + s_warp[w];
}

__device__ __forceinline__ void copy_track(const TrackArrays &in, int i, const TrackArrays &out, int o, int cnt_add) {
    out.cur[o] = in.cur[i];
    out.forw[o] = in.forw[i];
    out.ids[o] = in.ids[i];
    out.cnt[o] = in.cnt[i] + cnt_add;
    out.bearing[3 * o + 0] = in.bearing[3 * i + 0];
    out.bearing[3 * o + 1] = in.bearing[3 * i + 1];
    out.bearing[3 * o + 2] = in.bearing[3 * i + 2];
    out.prev_valid[o] = in.prev_valid[i];
}

__global__ void __launch_bounds__(BT) compact_kernel(TrackArrays in, TrackArrays out, const uint8_t *__restrict__ status,
                                                     int *n_dev, int cap, int cnt_add) {
    __shared__ int s_warp[33];
    const int n = min(*n_dev, cap);
    int carry = 0;
    for (int i0 = 0; i0 < n; i0 += BT) {
        const int i = i0 + threadIdx.x;
        const int keep = (i < n && status[i]) ? 1 : 0;
        int total;
        const int ex = block_excl_scan(keep, s_warp, &total);
        if (keep) copy_track(in, i, out, carry + ex, cnt_add);  // 1 after LK: for (auto &n : track_cnt) n++
        carry += total;
        __syncthreads();
    }
    __syncthreads();
    if (threadIdx.x == 0) *n_dev = carry;
}

// ------------------------------------------------------------------------------------------------
// libstdc++ std::sort(first, last, [](a, b){ return a.first > b.first; }) on (key, payload) pairs packed as
// (key << 32 | payload); only the key takes part in comparisons.
struct DescByKey {
    __device__ __forceinline__ bool operator()(long long a, long long b) const { return (int)(a >> 32) > (int)(b >> 32); }
};

__device__ void adjust_heap(long long *first, int hole, int len, long long value, DescByKey comp) {
    const int top = hole;
    int child = hole;
    while (child < (len - 1) / 2) {
        child = 2 * (child + 1);
        if (comp(first[child], first[child - 1])) child--;
        first[hole] = first[child];
        hole = child;
    }
    if ((len & 1) == 0 && child == (len - 2) / 2) {
        child = 2 * (child + 1);
        first[hole] = first[child - 1];
        hole = child - 1;
    }
    int parent = (hole - 1) / 2;  // __push_heap
    while (hole > top && comp(first[parent], value)) {
        first[hole] = first[parent];
        hole = parent;
        parent = (hole - 1) / 2;
    }
    first[hole] = value;
}

__device__ void heap_sort(long long *first, int len, DescByKey comp) {  // __partial_sort(first, last, last)
    if (len >= 2) {
        int parent = (len - 2) / 2;
        while (true) {
            long long v = first[parent];
            adjust_heap(first, parent, len, v, comp);
            if (parent == 0) break;
            parent--;
        }
    }
    int last = len;
    while (last > 1) {
        --last;
        long long v = first[last];
        first[last] = first[0];
        adjust_heap(first, 0, last, v, comp);
    }
}

__device__ void unguarded_linear_insert(long long *a, int last, DescByKey comp) {
    long long val = a[last];
    int next = last - 1;
    while (comp(val, a[next])) {
        a[last] = a[next];
        last = next;
        --next;
    }
    a[last] = val;
}

__device__ void insertion_sort(long long *a, int first, int last, DescByKey comp) {
    if (first == last) return;
    for (int i = first + 1; i != last; ++i) {
        if (comp(a[i], a[first])) {
            long long val = a[i];
            for (int k = i; k > first; --k) a[k] = a[k - 1];  // move_backward
            a[first] = val;
        } else {
            unguarded_linear_insert(a, i, comp);
        }
    }
}

__device__ void std_sort_desc(long long *a, int n) {
    DescByKey comp;
    if (n <= 1) return;
    // __introsort_loop with an explicit stack (the sub-ranges are independent, so order does not matter)
    int stack_first[64], stack_last[64], stack_depth[64];
    int sp = 0;
    int lg = 31 - __clz(n);
    stack_first[sp] = 0;
    stack_last[sp] = n;
    stack_depth[sp] = 2 * lg;
    sp++;
    while (sp > 0) {
        sp--;
        int first = stack_first[sp], last = stack_last[sp], depth = stack_depth[sp];
        while (last - first > 16) {
            if (depth == 0) {
                heap_sort(a + first, last - first, comp);
                break;
            }
            --depth;
            // __unguarded_partition_pivot: median of (first+1, mid, last-1) moved to first
            const int mid = first + (last - first) / 2;
            const int ia = first + 1, ib = mid, ic = last - 1;
            int m;
            if (comp(a[ia], a[ib])) {
                if (comp(a[ib], a[ic])) m = ib;
                else if (comp(a[ia], a[ic])) m = ic;
                else m = ia;
            } else if (comp(a[ia], a[ic])) m = ia;
            else if (comp(a[ib], a[ic])) m = ic;
            else m = ib;
            {
                long long t = a[first];
                a[first] = a[m];
                a[m] = t;
            }
            // __unguarded_partition(first + 1, last, pivot = first)
            int lo = first + 1, hi = last;
            const long long pivot = a[first];
            while (true) {
                while (comp(a[lo], pivot)) ++lo;
                --hi;
                while (comp(pivot, a[hi])) --hi;
                if (!(lo < hi)) break;
                long long t = a[lo];
                a[lo] = a[hi];
                a[hi] = t;
                ++lo;
            }
            const int cut = lo;
            // recurse on [cut, last), loop on [first, cut)
            stack_first[sp] = cut;
            stack_last[sp] = last;
            stack_depth[sp] = depth;
            sp++;
            last = cut;
        }
    }
    // __final_insertion_sort
    if (n > 16) {
        insertion_sort(a, 0, 16, comp);
        for (int i = 16; i != n; ++i) unguarded_linear_insert(a, i, comp);
    } else {
        insertion_sort(a, 0, n, comp);
    }
}

// ---- the same std::sort, level-parallel ------------------------------------------------------------------
// libstdc++'s result for equal keys depends on every swap it makes, so the algorithm itself is replayed — but its
// independent pieces run in parallel:
//   * __introsort_loop: all sub-ranges of one recursion level are partitioned concurrently, one warp per range.
//     __unguarded_partition is a pure function of the range's ORIGINAL contents: with l_1 < l_2 < ... the positions
//     (from the left) whose key is <= the pivot's and r_1 > r_2 > ... the positions (from the right) whose key is >= it,
//     it swaps (l_k, r_k) for k = 1 .. K, K = #{k : l_k < r_k}, and returns cut = min(l_{K+1}, r_K) — ballots and
//     prefix counts instead of two dependent pointer walks.
//   * depth exhausted -> heap sort of that range by one lane (as rare as in the reference);
//   * __final_insertion_sort is a STABLE sort of whatever the loop left, i.e. rank(i) = #{key_j > key_i} + #{j < i :
//     key_j == key_i}: one pass of comparisons per element, all threads.
struct SortShared {
    long long *a;   // cap records (key << 32 | source index)
    long long *tmp; // cap records (stable-sort scatter target)
    int *lpos;      // cap: L-stopper positions of each range, stored at the range's own offset
    int *rpos;      // cap: R-stopper positions (left to right)
    int *rf, *rl, *rd;  // two range lists of max_ranges each: first / last / depth
};

__device__ void parallel_std_sort_desc(const SortShared &S, int n, int max_ranges, int *s_cnt /* 2 ints */) {
    DescByKey comp;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5, nwarp = blockDim.x >> 5;
    if (n <= 1) return;
    long long *a = S.a;
    if (n > 16) {
        if (tid == 0) {
            S.rf[0] = 0;
            S.rl[0] = n;
            S.rd[0] = 2 * (31 - __clz(n));
            s_cnt[0] = 1;
            s_cnt[1] = 0;
        }
        __syncthreads();
        for (int level = 0;; level++) {
            const int cur = level & 1, nxt = cur ^ 1;
            const int n_cur = s_cnt[cur];
            if (n_cur == 0) break;
            const int *cf = S.rf + cur * max_ranges, *cl = S.rl + cur * max_ranges, *cd = S.rd + cur * max_ranges;
            int *nf = S.rf + nxt * max_ranges, *nl = S.rl + nxt * max_ranges, *nd = S.rd + nxt * max_ranges;
            for (int ri = wid; ri < n_cur; ri += nwarp) {
                const int first = cf[ri], last = cl[ri];
                int depth = cd[ri];
                if (depth == 0) {
                    if (lane == 0) heap_sort(a + first, last - first, comp);
                    continue;
                }
                --depth;
                if (lane == 0) {  // __move_median_to_first(first, first + 1, mid, last - 1)
                    const int mid = first + (last - first) / 2;
                    const int ia = first + 1, ib = mid, ic = last - 1;
                    int m;
                    if (comp(a[ia], a[ib])) {
                        if (comp(a[ib], a[ic])) m = ib;
                        else if (comp(a[ia], a[ic])) m = ic;
                        else m = ia;
                    } else if (comp(a[ia], a[ic])) m = ia;
                    else if (comp(a[ib], a[ic])) m = ic;
                    else m = ib;
                    const long long t = a[first];
                    a[first] = a[m];
                    a[m] = t;
                }
                __syncwarp();
                const int pk = (int)(a[first] >> 32);
                int nL = 0, nR = 0;
                for (int p0 = first + 1; p0 < last; p0 += 32) {
                    const int p = p0 + lane;
                    const int key = p < last ? (int)(a[p] >> 32) : 0;
                    const bool isL = p < last && key <= pk, isR = p < last && key >= pk;
                    const unsigned bl = __ballot_sync(0xffffffffu, isL), br = __ballot_sync(0xffffffffu, isR);
                    const unsigned below = (1u << lane) - 1u;
                    if (isL) S.lpos[first + nL + __popc(bl & below)] = p;
                    if (isR) S.rpos[first + nR + __popc(br & below)] = p;
                    nL += __popc(bl);
                    nR += __popc(br);
                }
                __syncwarp();
                // k-th (1-based) stopper from the left: lpos[first + k - 1]; from the right: rpos[first + nR - k]
                int K = 0;
                for (int k = 1 + lane; k <= min(nL, nR); k += 32)
                    if (S.lpos[first + k - 1] < S.rpos[first + nR - k]) K++;
                K = warp_sum(K);
                for (int k = 1 + lane; k <= K; k += 32) {
                    const int l = S.lpos[first + k - 1], r = S.rpos[first + nR - k];
                    const long long t = a[l];
                    a[l] = a[r];
                    a[r] = t;
                }
                const int c1 = K < nL ? S.lpos[first + K] : 0x7fffffff, c2 = K >= 1 ? S.rpos[first + nR - K] : 0x7fffffff;
                const int cut = min(c1, c2);
                if (lane == 0) {
                    if (last - cut > 16) {
                        const int q = atomicAdd(&s_cnt[nxt], 1);
                        nf[q] = cut;
                        nl[q] = last;
                        nd[q] = depth;
                    }
                    if (cut - first > 16) {
                        const int q = atomicAdd(&s_cnt[nxt], 1);
                        nf[q] = first;
                        nl[q] = cut;
                        nd[q] = depth;
                    }
                }
            }
            __syncthreads();
            if (tid == 0) s_cnt[cur] = 0;
            __syncthreads();
        }
    }
    // __final_insertion_sort == stable sort of the current order
    __syncthreads();
    for (int i = tid; i < n; i += blockDim.x) {
        const int ki = (int)(a[i] >> 32);
        int rank = 0;
        for (int j = 0; j < n; j++) {
            const int kj = (int)(a[j] >> 32);
            rank += (kj > ki) || (kj == ki && j < i);
        }
        S.tmp[rank] = a[i];
    }
    __syncthreads();
    for (int i = tid; i < n; i += blockDim.x) a[i] = S.tmp[i];
    __syncthreads();
}

// setMask (feature_tracker.cpp:46-83): sort by track_cnt (std::sort order), then the greedy keep loop.  One CTA; the
// sort records, pixel coordinates and kept list live in shared memory; the keep loop is run by ONE warp (ballots, no
// block barriers) over coordinates and mask bits gathered in parallel beforehand.
__global__ void __launch_bounds__(BT) set_mask_select_kernel(TrackArrays in, TrackArrays out, int *n_dev, int cap,
                                                             const uint8_t *__restrict__ base_mask, int mask_pitch,
                                                             const int *__restrict__ half_width, int min_dist, int max_ranges) {
    extern __shared__ long long smem_ll[];
    __shared__ int s_cnt[2];
    __shared__ int s_nkept;
    SortShared S;
    S.a = smem_ll;
    S.tmp = S.a + cap;
    S.lpos = reinterpret_cast<int *>(S.tmp + cap);
    S.rpos = S.lpos + cap;
    S.rf = S.rpos + cap;
    S.rl = S.rf + 2 * max_ranges;
    S.rd = S.rl + 2 * max_ranges;
    int *s_px = S.rd + 2 * max_ranges, *s_py = s_px + cap, *s_ok = s_py + cap, *kept = s_ok + cap;  // kept: indices into the sorted order
    const int n = min(*n_dev, cap);
    for (int i = threadIdx.x; i < n; i += BT) S.a[i] = ((long long)in.cnt[i] << 32) | (unsigned)i;
    __syncthreads();
    parallel_std_sort_desc(S, n, max_ranges, s_cnt);
    for (int k = threadIdx.x; k < n; k += BT) {
        const int src = (int)(S.a[k] & 0xffffffffu);
        const float2 pt = in.forw[src];
        const int px = cv_round(pt.x), py = cv_round(pt.y);  // Mat::at<uchar>(Point2f -> Point)
        s_px[k] = px;
        s_py[k] = py;
        s_ok[k] = base_mask[(size_t)py * mask_pitch + px] == 255;
    }
    __syncthreads();
    if (n > 512) {
        // many tracks (3840x2880 / 2000 features): the whole CTA tests a candidate against the kept list
        int nk = 0;
        for (int k = 0; k < n; k++) {
            if (!s_ok[k]) continue;  // (uniform)
            const int px = s_px[k], py = s_py[k];
            int hit = 0;
            for (int j = threadIdx.x; j < nk; j += BT) {
                const int q = kept[j];
                const int dx = abs(px - s_px[q]), dy = abs(py - s_py[q]);
                if (dy <= min_dist && dx <= half_width[dy]) hit = 1;
            }
            if (!__syncthreads_or(hit)) {  // (uniform)
                if (threadIdx.x == 0) kept[nk] = k;
                nk++;
                __syncthreads();  // the new entry is read by everybody in the next round
            }
        }
        if (threadIdx.x == 0) s_nkept = nk;
    } else if (threadIdx.x < 32) {
        const int lane = threadIdx.x;
        int nk = 0;
        for (int k = 0; k < n; k++) {
            if (!s_ok[k]) continue;  // (uniform)
            const int px = s_px[k], py = s_py[k];
            int hit = 0;
            for (int j = lane; j < nk; j += 32) {
                const int q = kept[j];
                const int dx = abs(px - s_px[q]), dy = abs(py - s_py[q]);
                if (dy <= min_dist && dx <= half_width[dy]) hit = 1;
            }
            if (!__any_sync(0xffffffffu, hit)) {
                if (lane == 0) kept[nk] = k;
                nk++;
                __syncwarp();
            }
        }
        if (lane == 0) s_nkept = nk;
    }
    __syncthreads();
    const int nkept = s_nkept;
    for (int o = threadIdx.x; o < nkept; o += BT) copy_track(in, (int)(S.a[kept[o]] & 0xffffffffu), out, o, 0);
    __syncthreads();
    if (threadIdx.x == 0) *n_dev = nkept;
}

__global__ void __launch_bounds__(64) mask_raster_kernel(uint8_t *mask, int rows, int cols, int pitch,
                                                         const float2 *__restrict__ pts, const int *__restrict__ n_dev,
                                                         int cap, const int *__restrict__ half_width, int radius) {
    const int n = min(*n_dev, cap);
    const int i = blockIdx.x;
    if (i >= n) return;
    const float2 pt = pts[i];
    const int cx = cv_round(pt.x), cy = cv_round(pt.y);  // cv::circle(mask, Point2f -> Point, MIN_DIST, 0, -1)
    for (int r = threadIdx.x; r < 2 * radius + 1; r += blockDim.x) {
        const int dy = r - radius;
        const int y = cy + dy;
        if (y < 0 || y >= rows) continue;
        const int hw = half_width[abs(dy)];
        const int xa = max(cx - hw, 0), xb = min(cx + hw, cols - 1);
        for (int x = xa; x <= xb; x++) mask[(size_t)y * pitch + x] = 0;
    }
}

__global__ void fill_circle_kernel(uint8_t *img, int rows, int cols, int pitch, int cx, int cy, int radius,
                                   const int *__restrict__ half_width, uint8_t value) {
    const int r = blockIdx.x;  // row index within [-radius, radius]
    const int dy = r - radius;
    const int y = cy + dy;
    if (y < 0 || y >= rows) return;
    const int hw = half_width[abs(dy)];
    const int xa = max(cx - hw, 0), xb = min(cx + hw, cols - 1);
    for (int x = xa + threadIdx.x; x <= xb; x += blockDim.x) img[(size_t)y * pitch + x] = value;
}

__global__ void __launch_bounds__(256) undistort_kernel(OcamDev cam, const float2 *__restrict__ pts,
                                                        const int *__restrict__ ids, float *bearing_prev,
                                                        uint8_t *prev_valid, float *__restrict__ bearing_out,
                                                        float *__restrict__ vel_out, const int *__restrict__ n_dev, int cap,
                                                        double dt, const int *__restrict__ have_prev_in,
                                                        int *__restrict__ have_prev_out, const double *__restrict__ dt_dev) {
    if (dt_dev) dt = *dt_dev;
    const int n = min(*n_dev, cap);
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int have_prev = *have_prev_in;  // !prev_un_pts_map_3d.empty()
    if (i == 0) *have_prev_out = (n > 0) ? 1 : 0;
    if (i >= n) return;
    const float2 p = pts[i];
    const double3 b = ocam_lift(cam, (double)p.x, (double)p.y, true);  // b / b.norm() in f64 ...
    const float bx = (float)b.x, by = (float)b.y, bz = (float)b.z;     // ... then cv::Point3f
    float vx = 0.f, vy = 0.f, vz = 0.f;
    const int id = ids[i];
    if (have_prev && id != -1 && prev_valid[i]) {
        // (cur.x - prev.x) is a float subtraction, the division by dt is done in double, result cast to float
        vx = (float)((double)__fsub_rn(bx, bearing_prev[3 * i + 0]) / dt);
        vy = (float)((double)__fsub_rn(by, bearing_prev[3 * i + 1]) / dt);
        vz = (float)((double)__fsub_rn(bz, bearing_prev[3 * i + 2]) / dt);
    }
    bearing_out[3 * i + 0] = bx;
    bearing_out[3 * i + 1] = by;
    bearing_out[3 * i + 2] = bz;
    vel_out[3 * i + 0] = vx;
    vel_out[3 * i + 1] = vy;
    vel_out[3 * i + 2] = vz;
    // becomes prev_un_pts_map_3d: keyed by the id in force NOW (-1 for points added this frame)
    bearing_prev[3 * i + 0] = bx;
    bearing_prev[3 * i + 1] = by;
    bearing_prev[3 * i + 2] = bz;
    prev_valid[i] = (id != -1) ? 1 : 0;
}

__global__ void __launch_bounds__(BT) update_ids_kernel(int *ids, const int *__restrict__ n_dev, int cap, int *n_id_dev) {
    __shared__ int s_warp[33];
    const int n = min(*n_dev, cap);
    const int n_id_base = *n_id_dev;  // static int FeatureTracker::n_id
    __syncthreads();
    int carry = 0;
    for (int i0 = 0; i0 < n; i0 += BT) {
        const int i = i0 + threadIdx.x;
        const int isnew = (i < n && ids[i] == -1) ? 1 : 0;
        int total;
        const int ex = block_excl_scan(isnew, s_warp, &total);
        if (isnew) ids[i] = n_id_base + carry + ex;  // ids[i] = n_id++ in index order
        carry += total;
        __syncthreads();
    }
    if (threadIdx.x == 0) *n_id_dev = n_id_base + carry;
}

}  // namespace

int launch_compact(const TrackArrays &in, const TrackArrays &out, const uint8_t *status, int *n_dev, int cap, int cnt_add,
                   cudaStream_t s) {
    compact_kernel<<<1, BT, 0, s>>>(in, out, status, n_dev, cap, cnt_add);
    LFVIO_LAUNCH_CHECK();
    return LFVIO_OK;
}

int launch_set_mask_select(const TrackArrays &in, const TrackArrays &out, int *n_dev, int cap,
                           const uint8_t *base_mask, int mask_pitch, const int *half_width, int min_dist,
                           int *sort_scratch, cudaStream_t s) {
    (void)sort_scratch;  // (the sort records live in shared memory now)
    const int max_ranges = cap / 8 + 8;
    const size_t smem = sizeof(long long) * 2 * (size_t)cap + sizeof(int) * (2 * (size_t)cap + 6 * (size_t)max_ranges + 4 * (size_t)cap);
    LFVIO_REQUIRE(smem <= 200 * 1024, "setMask: too many tracks for the shared-memory sort (max_cnt <= ~5000)");
    LFVIO_TRY(ensure_dyn_smem(reinterpret_cast<const void *>(set_mask_select_kernel), 200 * 1024));
    set_mask_select_kernel<<<1, BT, smem, s>>>(in, out, n_dev, cap, base_mask, mask_pitch, half_width, min_dist, max_ranges);
    LFVIO_LAUNCH_CHECK();
    return LFVIO_OK;
}

int launch_mask_raster(uint8_t *mask, const uint8_t *base_mask, int rows, int cols, int pitch,
                       const float2 *pts, const int *n_dev, int cap, const int *half_width, int min_dist,
                       cudaStream_t s) {
    LFVIO_CUDA_TRY(cudaMemcpy2DAsync(mask, pitch, base_mask, pitch, cols, rows, cudaMemcpyDeviceToDevice, s));
    if (cap > 0) {
        mask_raster_kernel<<<cap, 64, 0, s>>>(mask, rows, cols, pitch, pts, n_dev, cap, half_width, min_dist);
        LFVIO_LAUNCH_CHECK();
    }
    return LFVIO_OK;
}

int launch_fill_circle(uint8_t *img, int rows, int cols, int pitch, int cx, int cy, int radius,
                       const int *half_width_dev, uint8_t value, cudaStream_t s) {
    fill_circle_kernel<<<2 * radius + 1, 128, 0, s>>>(img, rows, cols, pitch, cx, cy, radius, half_width_dev, value);
    LFVIO_LAUNCH_CHECK();
    return LFVIO_OK;
}

int launch_undistort(const OcamDev &cam, const float2 *pts, const int *ids, float *bearing_prev_inout,
                     uint8_t *prev_valid_inout, float *bearing_out, float *vel_out, const int *n_dev, int cap,
                     double dt, const int *have_prev_in, int *have_prev_out, cudaStream_t s, const double *dt_dev) {
    if (cap <= 0) cap = 1;  // still rotates the have_prev flag
    undistort_kernel<<<cdiv(cap, 256), 256, 0, s>>>(cam, pts, ids, bearing_prev_inout, prev_valid_inout, bearing_out,
                                                    vel_out, n_dev, cap, dt, have_prev_in, have_prev_out, dt_dev);
    LFVIO_LAUNCH_CHECK();
    return LFVIO_OK;
}

int launch_update_ids(int *ids, const int *n_dev, int cap, int *n_id_dev, cudaStream_t s) {
    update_ids_kernel<<<1, BT, 0, s>>>(ids, n_dev, cap, n_id_dev);
    LFVIO_LAUNCH_CHECK();
    return LFVIO_OK;
}

}  // namespace lfvio
