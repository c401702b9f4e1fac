// FeatureTracker on device: per-frame pipeline + the C-ABI entry points of include/lfvio.h (front end).
//
// Reference: class FeatureTracker, feature_tracker/src/feature_tracker.{h,cpp}; readImage :95-184.
// One handle == one camera == one CUDA stream.  All per-track state (positions, ids, track_cnt, previous
// bearings) lives in HBM as SoA; counts stay on device so the whole frame is enqueued without a host
// round trip; the host mirror is refreshed with one batch of small D2H copies at the end of the frame.
#include <cstdlib>
#include <cstring>
#include <map>
#include <vector>

#include "frontend.cuh"
#include "tma.cuh"

namespace lfvio {

namespace {

template <class T>
struct DevBuf {
    T *p = nullptr;
    size_t n = 0;
    int alloc(size_t count) {
        n = count;
        if (count == 0) count = 1;
        LFVIO_CUDA_TRY(cudaMalloc(&p, sizeof(T) * count));
        LFVIO_CUDA_TRY(cudaMemset(p, 0, sizeof(T) * count));
        return LFVIO_OK;
    }
    ~DevBuf() {
        if (p) cudaFree(p);
    }
};
template <class T>
struct PinBuf {
    T *p = nullptr;
    int alloc(size_t count) {
        if (count == 0) count = 1;
        LFVIO_CUDA_TRY(cudaMallocHost(&p, sizeof(T) * count));
        memset(p, 0, sizeof(T) * count);
        return LFVIO_OK;
    }
    ~PinBuf() {
        if (p) cudaFreeHost(p);
    }
};

struct TrackStore {
    DevBuf<float2> cur, forw;
    DevBuf<int> ids, cnt;
    DevBuf<float> bearing;
    DevBuf<uint8_t> prev_valid;
    int alloc(int cap) {
        LFVIO_TRY(cur.alloc(cap));
        LFVIO_TRY(forw.alloc(cap));
        LFVIO_TRY(ids.alloc(cap));
        LFVIO_TRY(cnt.alloc(cap));
        LFVIO_TRY(bearing.alloc(3 * (size_t)cap));
        LFVIO_TRY(prev_valid.alloc(cap));
        return LFVIO_OK;
    }
    TrackArrays view() { return TrackArrays{cur.p, forw.p, ids.p, cnt.p, bearing.p, prev_valid.p}; }
};

// Image-operator scratch shared by the tracker and the stage-level entry points.
struct CornerStore {
    DevBuf<unsigned long long> cand, cand_sorted;
    DevBuf<int> cell_count, cell_start, counters;
    DevBuf<uint32_t> hist, max_bits;
    DevBuf<uint8_t> state;
    DevBuf<float> eig;
    CornerScratch cs;
    int alloc(int rows, int cols, int cell) {
        const size_t px = (size_t)rows * cols;
        const int cap = (int)std::max<size_t>(px / 4, 1024);
        LFVIO_TRY(cand.alloc(cap));
        LFVIO_TRY(cand_sorted.alloc(cap));
        LFVIO_TRY(state.alloc(cap));
        cs.cell = cell < 1 ? 1 : cell;
        cs.gw = (cols + cs.cell - 1) / cs.cell;
        cs.gh = (rows + cs.cell - 1) / cs.cell;
        LFVIO_TRY(cell_count.alloc((size_t)cs.gw * cs.gh + 1));
        LFVIO_TRY(cell_start.alloc((size_t)cs.gw * cs.gh + 1));
        LFVIO_TRY(counters.alloc(8));
        LFVIO_TRY(hist.alloc(kCandHistBins));
        LFVIO_TRY(max_bits.alloc(1));
        LFVIO_TRY(eig.alloc(px));
        cs.cand = cand.p;
        cs.cand_sorted = cand_sorted.p;
        cs.cand_cap = cap;
        cs.cell_count = cell_count.p;
        cs.cell_start = cell_start.p;
        cs.hist = hist.p;
        cs.state = state.p;
        cs.counters = counters.p;
        return LFVIO_OK;
    }
};

int align_up(int v, int a) { return (v + a - 1) / a * a; }

int alloc_pyramid(Pyramid *p, std::vector<DevBuf<uint8_t>> *store, int rows, int cols, int max_level, int win) {
    make_pyramid_layout(p, rows, cols, max_level, win);
    store->resize(p->levels);
    for (int l = 0; l < p->levels; l++) {
        p->lvl[l].pitch = align_up(p->lvl[l].cols, 128);
        LFVIO_TRY((*store)[l].alloc((size_t)p->lvl[l].pitch * p->lvl[l].rows));
        p->lvl[l].data = (*store)[l].p;
    }
    return LFVIO_OK;
}

LkParams default_lk(int rows, int cols, int border_check) {
    LkParams prm;
    prm.win = 41;
    prm.max_iter = 30;                 // TermCriteria(COUNT+EPS, 30, 0.01)
    prm.eps2 = 0.01 * 0.01;            // criteria.epsilon *= criteria.epsilon
    prm.min_eig_thr = 1e-4;
    prm.border_check = border_check;
    prm.rows = rows;
    prm.cols = cols;
    return prm;
}

}  // namespace

}  // namespace lfvio

using namespace lfvio;

enum TrackerStage { TS_CLAHE = 0, TS_PYRAMID, TS_LK, TS_COMPACT, TS_RANSAC, TS_SETMASK, TS_MASK_RASTER, TS_MIN_EIG, TS_CANDIDATES,
                    TS_SELECT, TS_UNDISTORT, TS_UPDATE_IDS, TS_COUNT };
static const char *kTrackerStageNames[TS_COUNT] = {"clahe(3 kernels)", "pyr_down(3 kernels)", "lk_kernel", "compact_kernel",
                                                   "reject_with_f(4 kernels + compact)",
                                                   "set_mask_select_kernel", "mask_raster(copy+kernel)", "min_eig_kernel",
                                                   "candidates_kernel", "select_kernel", "undistort_kernel",
                                                   "update_ids_kernel"};

struct lfvio_tracker {
    StageProfiler prof;
    lfvio_tracker_cfg cfg;
    int device = 0, n_sm = 148, cap = 0, pitch = 0;
    cudaStream_t stream = nullptr;
    OcamDev cam;
    // images
    DevBuf<uint8_t> raw, mask, base_mask;
    std::vector<DevBuf<uint8_t>> pyr_store[2];
    Pyramid pyr[2];
    PyrTmaps pyr_tm[2];
    CUtensorMap eig_tm[2];     // level 0 of either pyramid slot with the 96 x 68 box of min_eig_kernel
    bool eig_tm_ok[2] = {false, false};
    int cur_slot = 0;
    long frames = 0;
    // CLAHE scratch
    DevBuf<uint32_t> clahe_hist;
    DevBuf<uint8_t> clahe_lut;
    // tracks (double buffered for compaction / reordering)
    TrackStore store[2];
    int a = 0;           // index of the live store
    bool swapped_cf[2] = {false, false};
    DevBuf<uint8_t> status;
    // rejectWithF scratch (cfg.ransac)
    DevBuf<double> rs_b1, rs_b2, rs_E, rs_E_best;
    DevBuf<float> rs_score;
    DevBuf<uint8_t> rs_flags;
    DevBuf<int> rs_best;
    RansacScratch ransac() {
        return RansacScratch{reinterpret_cast<double3 *>(rs_b1.p), reinterpret_cast<double3 *>(rs_b2.p), rs_E.p, rs_score.p, rs_flags.p,
                             rs_E_best.p, rs_best.p};
    }
    DevBuf<int> n_dev, have_prev, n_id_dev, sort_scratch, half_width;
    DevBuf<float> bearing_out, vel_out;
    CornerStore corners;
    // host side
    PinBuf<uint8_t> h_img;
    PinBuf<int> h_n;       // [0]=n, [1]=n_id, [2..3]=corner candidate counter / overflow flag
    PinBuf<float> h_pts, h_bearing, h_vel;
    PinBuf<int> h_ids, h_cnt;
    int n_host = 0, n_upper = 0, n_id_host = 0;
    bool mirror_fresh = true;
    double cur_time = 0, prev_time = 0;
    // Per-frame scalars (dt of undistortedPoints, the frame counter of the injected RANSAC sampler) travel through a small
    // device block so that the kernel launches of a frame carry no frame-dependent argument and the whole readImage can be
    // replayed as ONE cudaGraphLaunch.  A frame's launch sequence depends only on (which track store is live, which
    // pyramid slot is current, whether cur/forw of each store are swapped, frame parity, publish flag, ids on device):
    // one executable graph per reachable combination, captured the first time it occurs.
    struct FrameParams {
        double dt;
        unsigned frame, pad;
    };
    struct GraphEntry {
        cudaGraphExec_t exec = nullptr;
        int a = 0, cur_slot = 0, kernels = 0;  // kernels: launches inside the graph (for lfvio_launch_count)
        bool swapped[2] = {false, false};
    };
    DevBuf<FrameParams> fp_dev;
    PinBuf<FrameParams> fp_host;  // ring: a slot is rewritten only kFpRing frames later
    static constexpr int kFpRing = 256;
    unsigned fp_next = 0;
    bool use_graphs = true;
    std::map<unsigned, GraphEntry> graphs;
    long graph_launches = 0, graph_captures = 0;

    TrackArrays live() { return store[a].view(); }
    TrackArrays spare() { return store[a ^ 1].view(); }
};

namespace {

int tracker_init(lfvio_tracker *t, const lfvio_tracker_cfg *cfg, int device) {
    t->cfg = *cfg;
    t->device = device;
    LFVIO_CUDA_TRY(cudaSetDevice(device));
    t->n_sm = sm_count(device);
    t->cap = cfg->capacity > 0 ? cfg->capacity : cfg->max_cnt;
    if (t->cap < cfg->max_cnt) t->cap = cfg->max_cnt;
    LFVIO_CUDA_TRY(cudaStreamCreateWithFlags(&t->stream, cudaStreamNonBlocking));
    t->cam = make_ocam_dev(cfg->cam);
    const int rows = cfg->rows, cols = cfg->cols;
    t->pitch = align_up(cols, 128);
    const size_t img_bytes = (size_t)t->pitch * rows;
    LFVIO_TRY(t->raw.alloc(img_bytes));
    LFVIO_TRY(t->mask.alloc(img_bytes));
    LFVIO_TRY(t->base_mask.alloc(img_bytes));
    for (int k = 0; k < 2; k++) {
        LFVIO_TRY(alloc_pyramid(&t->pyr[k], &t->pyr_store[k], rows, cols, 3, 41));
        make_pyramid_tmaps(t->pyr[k], &t->pyr_tm[k]);
        t->eig_tm_ok[k] = encode_u8_tmap(&t->eig_tm[k], t->pyr[k].lvl[0].data, t->pyr[k].lvl[0].rows, t->pyr[k].lvl[0].cols, t->pyr[k].lvl[0].pitch, 96, 68);
    }
    LFVIO_TRY(t->clahe_hist.alloc(kClaheTiles * kClaheTiles * 256));
    LFVIO_TRY(t->clahe_lut.alloc(kClaheTiles * kClaheTiles * 256));
    for (int k = 0; k < 2; k++) LFVIO_TRY(t->store[k].alloc(t->cap));
    LFVIO_TRY(t->status.alloc(t->cap));
    if (t->cfg.ransac) {
        LFVIO_TRY(t->rs_b1.alloc(3 * (size_t)t->cap));
        LFVIO_TRY(t->rs_b2.alloc(3 * (size_t)t->cap));
        LFVIO_TRY(t->rs_E.alloc(100 * 9));
        LFVIO_TRY(t->rs_E_best.alloc(9));
        LFVIO_TRY(t->rs_score.alloc(100));
        LFVIO_TRY(t->rs_flags.alloc(100 * (size_t)t->cap));
        LFVIO_TRY(t->rs_best.alloc(1));
    }
    LFVIO_TRY(t->fp_dev.alloc(1));
    LFVIO_TRY(t->fp_host.alloc(lfvio_tracker::kFpRing));
    {
        const char *e = getenv("LFVIO_NO_GRAPH"), *e2 = getenv("LFVIO_NO_GRAPH_TRACKER");
        t->use_graphs = !(e && e[0] == '1') && !(e2 && e2[0] == '1');
    }
    LFVIO_TRY(t->n_dev.alloc(1));
    LFVIO_TRY(t->have_prev.alloc(2));
    LFVIO_TRY(t->n_id_dev.alloc(1));
    LFVIO_TRY(t->sort_scratch.alloc(4 * (size_t)t->cap));
    LFVIO_TRY(t->bearing_out.alloc(3 * (size_t)t->cap));
    LFVIO_TRY(t->vel_out.alloc(3 * (size_t)t->cap));
    LFVIO_TRY(t->corners.alloc(rows, cols, cfg->min_dist));
    LFVIO_TRY(t->h_img.alloc((size_t)cols * rows));
    LFVIO_TRY(t->h_n.alloc(4));
    LFVIO_TRY(t->h_pts.alloc(2 * (size_t)t->cap));
    LFVIO_TRY(t->h_bearing.alloc(3 * (size_t)t->cap));
    LFVIO_TRY(t->h_vel.alloc(3 * (size_t)t->cap));
    LFVIO_TRY(t->h_ids.alloc(t->cap));
    LFVIO_TRY(t->h_cnt.alloc(t->cap));
    // annulus (feature_tracker.cpp:52-56): cv::Point(CENTER_X, CENTER_Y) truncates the doubles
    const int rmax = std::max(std::max(cfg->max_r, cfg->min_r), cfg->min_dist);
    std::vector<int> hw(rmax + 1);
    DevBuf<int> hw_dev;
    LFVIO_TRY(hw_dev.alloc(rmax + 1));
    const int cx = (int)cfg->center_x, cy = (int)cfg->center_y;
    const int radii[2] = {cfg->max_r, cfg->min_r};
    const uint8_t vals[2] = {255, 0};
    for (int k = 0; k < 2; k++) {
        if (radii[k] < 0) continue;
        circle_half_widths(radii[k], hw.data());
        LFVIO_CUDA_TRY(cudaMemcpyAsync(hw_dev.p, hw.data(), sizeof(int) * (radii[k] + 1), cudaMemcpyHostToDevice, t->stream));
        LFVIO_TRY(launch_fill_circle(t->base_mask.p, rows, cols, t->pitch, cx, cy, radii[k], hw_dev.p, vals[k], t->stream));
        LFVIO_CUDA_TRY(cudaStreamSynchronize(t->stream));
    }
    LFVIO_TRY(t->half_width.alloc(cfg->min_dist + 1));
    circle_half_widths(cfg->min_dist, hw.data());
    LFVIO_CUDA_TRY(cudaMemcpy(t->half_width.p, hw.data(), sizeof(int) * (cfg->min_dist + 1), cudaMemcpyHostToDevice));
    LFVIO_CUDA_TRY(cudaMemcpy2D(t->mask.p, t->pitch, t->base_mask.p, t->pitch, cols, rows, cudaMemcpyDeviceToDevice));
    return LFVIO_OK;
}

// The launch sequence of one readImage on t->stream (directly, or into a stream capture).  src is device memory (u8, pitch
// src_pitch).  dt and the RANSAC frame counter are read from t->fp_dev by the kernels.
int frame_body(lfvio_tracker *t, const uint8_t *src, int src_pitch, double cur_time, int pub, bool ids_on_device) {
    const lfvio_tracker_cfg &c = t->cfg;
    cudaStream_t s = t->stream;
    t->cur_time = cur_time;
    const int forw_slot = (t->frames == 0) ? t->cur_slot : (t->cur_slot ^ 1);
    Pyramid &forw = t->pyr[forw_slot];
    if (c.equalize) {
        t->prof.begin(TS_CLAHE, s);
        LFVIO_TRY(launch_clahe(src, src_pitch, forw.lvl[0].data, forw.lvl[0].pitch, c.rows, c.cols, 3.0f,
                               t->clahe_hist.p, t->clahe_lut.p, t->n_sm, s));
        t->prof.end(s);
    } else {
        LFVIO_CUDA_TRY(cudaMemcpy2DAsync(forw.lvl[0].data, forw.lvl[0].pitch, src, src_pitch, c.cols, c.rows,
                                         cudaMemcpyDeviceToDevice, s));
    }
    t->prof.begin(TS_PYRAMID, s);
        LFVIO_TRY(launch_build_pyramid(forw, s));
        t->prof.end(s);
    Pyramid &cur = t->pyr[t->cur_slot];  // first frame: prev_img = cur_img = forw_img = img
    // forw_pts.clear(); if (cur_pts.size() > 0) calcOpticalFlowPyrLK ... reduceVector ...
    TrackArrays A = t->live(), B = t->spare();
    if (t->n_upper > 0) {
        const LkParams prm = default_lk(c.rows, c.cols, 1);
        t->prof.begin(TS_LK, s);
        LFVIO_TRY(launch_lk(cur, t->pyr_tm[t->cur_slot], forw, A.cur, A.forw, t->status.p, t->n_upper, t->n_dev.p, prm, s));
        t->prof.end(s);
        t->prof.begin(TS_COMPACT, s);
        LFVIO_TRY(launch_compact(A, B, t->status.p, t->n_dev.p, t->cap, 1, s));
        t->prof.end(s);  // + track_cnt++
        t->a ^= 1;
        A = t->live();
        B = t->spare();
    }
    if (pub) {
        if (c.ransac && t->n_upper > 0) {  // rejectWithF() :149, :329-386
            t->prof.begin(TS_RANSAC, s);
            LFVIO_TRY(launch_reject_with_f(t->cam, A.cur, A.forw, t->n_dev.p, t->n_upper, t->cap, c.ransac_seed, (unsigned)t->frames,
                                           t->ransac(), t->status.p, s, &t->fp_dev.p->frame));
            LFVIO_TRY(launch_compact(A, B, t->status.p, t->n_dev.p, t->cap, 0, s));
            t->prof.end(s);
            t->a ^= 1;
            A = t->live();
            B = t->spare();
        }
        t->prof.begin(TS_SETMASK, s);
        LFVIO_TRY(launch_set_mask_select(A, B, t->n_dev.p, t->cap, t->base_mask.p, t->pitch, t->half_width.p, c.min_dist,
                                         t->sort_scratch.p, s));
        t->prof.end(s);
        t->a ^= 1;
        A = t->live();
        t->prof.begin(TS_MASK_RASTER, s);
        LFVIO_TRY(launch_mask_raster(t->mask.p, t->base_mask.p, c.rows, c.cols, t->pitch, A.forw, t->n_dev.p, t->cap,
                                     t->half_width.p, c.min_dist, s));
        t->prof.end(s);
        // goodFeaturesToTrack(forw_img, n_pts, MAX_CNT - forw_pts.size(), 0.01, MIN_DIST, mask) + addPoints()
        t->prof.begin(TS_MIN_EIG, s);
        LFVIO_TRY(launch_min_eig(forw.lvl[0].data, forw.lvl[0].pitch, t->mask.p, t->pitch, t->corners.eig.p, c.rows, c.cols,
                                 t->corners.max_bits.p, s, t->eig_tm_ok[forw_slot] ? &t->eig_tm[forw_slot] : nullptr));
        t->prof.end(s);
        t->prof.begin(TS_CANDIDATES, s);
        LFVIO_TRY(launch_corner_candidates(t->corners.eig.p, t->mask.p, t->pitch, c.rows, c.cols, t->corners.max_bits.p,
                                           0.01, t->corners.cs, s));
        t->prof.end(s);
        t->prof.begin(TS_SELECT, s);
        LFVIO_TRY(launch_corner_select(t->corners.cs, c.cols, (float)c.min_dist, c.max_cnt, t->n_dev.p, A.forw, A.ids,
                                       A.cnt, A.prev_valid, t->cap, s));
        t->prof.end(s);
        t->n_upper = std::max(t->n_upper, c.max_cnt);
    }
    // prev_img = cur_img; cur_img = forw_img; cur_pts = forw_pts
    t->cur_slot = forw_slot;
    std::swap(t->store[t->a].cur.p, t->store[t->a].forw.p);
    t->swapped_cf[t->a] = !t->swapped_cf[t->a];
    A = t->live();
    const int hp = (int)(t->frames & 1);
    t->prof.begin(TS_UNDISTORT, s);
        LFVIO_TRY(launch_undistort(t->cam, A.cur, A.ids, A.bearing, A.prev_valid, t->bearing_out.p, t->vel_out.p,
                               t->n_dev.p, t->n_upper, cur_time - t->prev_time, t->have_prev.p + hp,
                               t->have_prev.p + (hp ^ 1), s, &t->fp_dev.p->dt));
        t->prof.end(s);
    t->prev_time = cur_time;
    t->frames++;
    if (ids_on_device) {
        t->prof.begin(TS_UPDATE_IDS, s);
        LFVIO_TRY(launch_update_ids(A.ids, t->n_dev.p, t->cap, t->n_id_dev.p, s));
        t->prof.end(s);
    }
    t->mirror_fresh = false;
    return LFVIO_OK;
}

// Enqueue one readImage on t->stream: stage the per-frame scalars, then either launch the frame's kernels one by one (first
// frame, stage profiling on, LFVIO_NO_GRAPH=1) or replay the executable graph of the current pointer state.
int enqueue_frame(lfvio_tracker *t, const uint8_t *src, int src_pitch, double cur_time, int pub, bool ids_on_device) {
    cudaStream_t s = t->stream;
    lfvio_tracker::FrameParams *hp = t->fp_host.p + (t->fp_next++ % lfvio_tracker::kFpRing);
    hp->dt = cur_time - t->prev_time;
    hp->frame = (unsigned)t->frames;
    LFVIO_CUDA_TRY(cudaMemcpyAsync(t->fp_dev.p, hp, sizeof(*hp), cudaMemcpyHostToDevice, s));
    if (!t->use_graphs || t->prof.on || t->frames == 0) return frame_body(t, src, src_pitch, cur_time, pub, ids_on_device);
    // grids sized by capacity, counts read on the device: the launch geometry is then the same for every frame
    t->n_upper = t->cap;
    if (src != t->raw.p)
        LFVIO_CUDA_TRY(cudaMemcpy2DAsync(t->raw.p, t->pitch, src, src_pitch, t->cfg.cols, t->cfg.rows, cudaMemcpyDeviceToDevice, s));
    const unsigned key = (unsigned)t->a | (unsigned)t->cur_slot << 1 | (unsigned)t->swapped_cf[0] << 2 | (unsigned)t->swapped_cf[1] << 3 |
                         (unsigned)(t->frames & 1) << 4 | (unsigned)(pub != 0) << 5 | (unsigned)ids_on_device << 6;
    auto it = t->graphs.find(key);
    if (it == t->graphs.end()) {  // first occurrence of this state: plain launches (also loads the kernels outside any capture)
        t->graphs.emplace(key, lfvio_tracker::GraphEntry());
        return frame_body(t, t->raw.p, t->pitch, cur_time, pub, ids_on_device);
    }
    if (it->second.exec == nullptr) {
        cudaGraph_t g = nullptr;
        LFVIO_CUDA_TRY(cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal));
        const unsigned long long l0 = thread_launch_count();
        const int st = frame_body(t, t->raw.p, t->pitch, cur_time, pub, ids_on_device);
        const int n_kernels = (int)(thread_launch_count() - l0);
        const cudaError_t ce = cudaStreamEndCapture(s, &g);
        if (st != LFVIO_OK) {
            if (g) cudaGraphDestroy(g);
            return st;
        }
        LFVIO_CUDA_TRY(ce);
        lfvio_tracker::GraphEntry e;
        LFVIO_CUDA_TRY(cudaGraphInstantiate(&e.exec, g, 0));
        cudaGraphDestroy(g);
        e.a = t->a, e.cur_slot = t->cur_slot, e.swapped[0] = t->swapped_cf[0], e.swapped[1] = t->swapped_cf[1];
        e.kernels = n_kernels;
        it->second = e;
        t->graph_captures++;
    } else {
        // the host-side bookkeeping frame_body would have done
        const lfvio_tracker::GraphEntry &e = it->second;
        for (int k = 0; k < 2; k++)
            if (e.swapped[k] != t->swapped_cf[k]) {
                std::swap(t->store[k].cur.p, t->store[k].forw.p);
                t->swapped_cf[k] = e.swapped[k];
            }
        t->a = e.a;
        t->cur_slot = e.cur_slot;
        t->cur_time = cur_time;
        t->prev_time = cur_time;
        t->frames++;
        t->mirror_fresh = false;
        count_launch(e.kernels);
    }
    LFVIO_CUDA_TRY(cudaGraphLaunch(it->second.exec, s));
    t->graph_launches++;
    return LFVIO_OK;
}

int download(lfvio_tracker *t) {
    cudaStream_t s = t->stream;
    TrackArrays A = t->live();
    const int m = t->n_upper;
    LFVIO_CUDA_TRY(cudaMemcpyAsync(t->h_n.p, t->n_dev.p, sizeof(int), cudaMemcpyDeviceToHost, s));
    LFVIO_CUDA_TRY(cudaMemcpyAsync(t->h_n.p + 1, t->n_id_dev.p, sizeof(int), cudaMemcpyDeviceToHost, s));
    if (m > 0) {
        LFVIO_CUDA_TRY(cudaMemcpyAsync(t->h_pts.p, A.cur, sizeof(float2) * m, cudaMemcpyDeviceToHost, s));
        LFVIO_CUDA_TRY(cudaMemcpyAsync(t->h_ids.p, A.ids, sizeof(int) * m, cudaMemcpyDeviceToHost, s));
        LFVIO_CUDA_TRY(cudaMemcpyAsync(t->h_cnt.p, A.cnt, sizeof(int) * m, cudaMemcpyDeviceToHost, s));
        LFVIO_CUDA_TRY(cudaMemcpyAsync(t->h_bearing.p, t->bearing_out.p, sizeof(float) * 3 * m, cudaMemcpyDeviceToHost, s));
        LFVIO_CUDA_TRY(cudaMemcpyAsync(t->h_vel.p, t->vel_out.p, sizeof(float) * 3 * m, cudaMemcpyDeviceToHost, s));
    }
    t->h_n.p[2] = t->h_n.p[3] = 0;
    if (t->corners.cs.counters)  // rides along with the other copies: no second round trip after the synchronisation
        LFVIO_CUDA_TRY(cudaMemcpyAsync(t->h_n.p + 2, t->corners.cs.counters, 2 * sizeof(int), cudaMemcpyDeviceToHost, s));
    LFVIO_CUDA_TRY(cudaStreamSynchronize(s));
    t->n_host = std::min(t->h_n.p[0], t->cap);
    t->n_id_host = t->h_n.p[1];
    t->n_upper = t->n_host;
    t->mirror_fresh = true;
    if (t->corners.cs.counters) {
        const int *flags = t->h_n.p + 2;
        if (flags[1]) {
            set_error("corner candidate buffer overflow (%d candidates > capacity %d)", flags[0], t->corners.cs.cand_cap);
            return LFVIO_ERR_CAPACITY;
        }
    }
    return LFVIO_OK;
}

}  // namespace

extern "C" int lfvio_reject_with_f(const lfvio_ocam *cam, const float *cur_pts, const float *forw_pts, int n,
                                   unsigned long long seed, unsigned frame, unsigned char *status, double *E_out,
                                   float *scores_out, int *best_out) {
    clear_error();
    LFVIO_REQUIRE(cam && cur_pts && forw_pts && status && n >= 0, "lfvio_reject_with_f: bad arguments");
    if (n == 0) return LFVIO_OK;
    OcamDev dc = make_ocam_dev(*cam);
    DevBuf<float2> c, f;
    DevBuf<double> b1, b2, E, Eb;
    DevBuf<float> score;
    DevBuf<uint8_t> flags, st;
    DevBuf<int> nd, best;
    LFVIO_TRY(c.alloc(n));
    LFVIO_TRY(f.alloc(n));
    LFVIO_TRY(b1.alloc(3 * (size_t)n));
    LFVIO_TRY(b2.alloc(3 * (size_t)n));
    LFVIO_TRY(E.alloc(900));
    LFVIO_TRY(Eb.alloc(9));
    LFVIO_TRY(score.alloc(100));
    LFVIO_TRY(flags.alloc(100 * (size_t)n));
    LFVIO_TRY(st.alloc(n));
    LFVIO_TRY(nd.alloc(1));
    LFVIO_TRY(best.alloc(1));
    LFVIO_CUDA_TRY(cudaMemcpy(c.p, cur_pts, sizeof(float2) * n, cudaMemcpyHostToDevice));
    LFVIO_CUDA_TRY(cudaMemcpy(f.p, forw_pts, sizeof(float2) * n, cudaMemcpyHostToDevice));
    LFVIO_CUDA_TRY(cudaMemcpy(nd.p, &n, sizeof(int), cudaMemcpyHostToDevice));
    LFVIO_CUDA_TRY(cudaMemset(Eb.p, 0, sizeof(double) * 9));
    LFVIO_CUDA_TRY(cudaMemset(score.p, 0, sizeof(float) * 100));
    LFVIO_CUDA_TRY(cudaMemset(best.p, 0xff, sizeof(int)));
    const RansacScratch rs{reinterpret_cast<double3 *>(b1.p), reinterpret_cast<double3 *>(b2.p), E.p, score.p, flags.p, Eb.p, best.p};
    LFVIO_TRY(launch_reject_with_f(dc, c.p, f.p, nd.p, n, n, seed, frame, rs, st.p, 0));
    LFVIO_CUDA_TRY(cudaMemcpy(status, st.p, n, cudaMemcpyDeviceToHost));
    if (E_out) LFVIO_CUDA_TRY(cudaMemcpy(E_out, Eb.p, sizeof(double) * 9, cudaMemcpyDeviceToHost));
    if (scores_out) LFVIO_CUDA_TRY(cudaMemcpy(scores_out, score.p, sizeof(float) * 100, cudaMemcpyDeviceToHost));
    if (best_out) LFVIO_CUDA_TRY(cudaMemcpy(best_out, best.p, sizeof(int), cudaMemcpyDeviceToHost));
    return LFVIO_OK;
}

extern "C" int lfvio_tracker_create(const lfvio_tracker_cfg *cfg, int device, lfvio_tracker **out) {
    clear_error();
    LFVIO_REQUIRE(cfg && out, "lfvio_tracker_create: null argument");
    LFVIO_REQUIRE(cfg->rows > 41 && cfg->cols > 41, "lfvio_tracker_create: image must exceed the 41x41 LK window");
    LFVIO_REQUIRE(cfg->max_cnt > 0 && cfg->min_dist >= 1, "lfvio_tracker_create: max_cnt > 0 and min_dist >= 1 required");
    LFVIO_REQUIRE(cfg->max_r >= 0 && cfg->min_r >= 0, "lfvio_tracker_create: bad annulus radii");
    lfvio_tracker *t = new (std::nothrow) lfvio_tracker();
    LFVIO_REQUIRE(t, "lfvio_tracker_create: out of host memory");
    int st = tracker_init(t, cfg, device);
    if (st != LFVIO_OK) {
        lfvio_tracker_destroy(t);
        return st;
    }
    *out = t;
    return LFVIO_OK;
}

extern "C" void lfvio_tracker_destroy(lfvio_tracker *t) {
    if (!t) return;
    cudaSetDevice(t->device);
    if (t->stream) {
        cudaStreamSynchronize(t->stream);
        for (auto &kv : t->graphs)
            if (kv.second.exec) cudaGraphExecDestroy(kv.second.exec);
        cudaStreamDestroy(t->stream);
    }
    delete t;
}

extern "C" int lfvio_tracker_read_image(lfvio_tracker *t, const uint8_t *img, int stride, double cur_time,
                                        int pub_this_frame) {
    clear_error();
    LFVIO_REQUIRE(t && img && stride >= t->cfg.cols, "lfvio_tracker_read_image: bad arguments");
    LFVIO_CUDA_TRY(cudaSetDevice(t->device));
    const int rows = t->cfg.rows, cols = t->cfg.cols;
    // The call is blocking (download() synchronises), so a caller buffer that is already page-locked (cudaHostAlloc /
    // cudaHostRegister — e.g. a ROS image transport that pins its pool) can be the DMA source directly.  Pageable memory is
    // staged through the handle's pinned buffer in row bands, so that the host copy of band i+1 overlaps the DMA of band i.
    cudaPointerAttributes attr{};
    const bool pinned = cudaPointerGetAttributes(&attr, img) == cudaSuccess && attr.type == cudaMemoryTypeHost;
    if (!pinned) (void)cudaGetLastError();  // (older drivers report unregistered memory as an error)
    if (pinned) {
        LFVIO_CUDA_TRY(cudaMemcpy2DAsync(t->raw.p, t->pitch, img, stride, cols, rows, cudaMemcpyHostToDevice, t->stream));
    } else {
        constexpr int kBands = 4;
        for (int b = 0; b < kBands; b++) {
            const int y0 = (int)((long long)rows * b / kBands), y1 = (int)((long long)rows * (b + 1) / kBands);
            for (int y = y0; y < y1; y++) memcpy(t->h_img.p + (size_t)y * cols, img + (size_t)y * stride, cols);
            LFVIO_CUDA_TRY(cudaMemcpy2DAsync(t->raw.p + (size_t)y0 * t->pitch, t->pitch, t->h_img.p + (size_t)y0 * cols, cols, cols, y1 - y0,
                                             cudaMemcpyHostToDevice, t->stream));
        }
    }
    LFVIO_TRY(enqueue_frame(t, t->raw.p, t->pitch, cur_time, pub_this_frame, false));
    return download(t);
}

extern "C" int lfvio_tracker_read_image_dev(lfvio_tracker *t, const uint8_t *img_dev, int stride, double cur_time,
                                            int pub_this_frame) {
    clear_error();
    LFVIO_REQUIRE(t && img_dev && stride >= t->cfg.cols, "lfvio_tracker_read_image_dev: bad arguments");
    LFVIO_CUDA_TRY(cudaSetDevice(t->device));
    return enqueue_frame(t, img_dev, stride, cur_time, pub_this_frame, true);
}

extern "C" int lfvio_tracker_sync(lfvio_tracker *t) {
    clear_error();
    LFVIO_REQUIRE(t, "lfvio_tracker_sync: null handle");
    LFVIO_CUDA_TRY(cudaSetDevice(t->device));
    return download(t);
}

extern "C" int lfvio_tracker_count(const lfvio_tracker *t) { return t ? t->n_host : 0; }

extern "C" int lfvio_tracker_get(const lfvio_tracker *t, lfvio_tracks *out) {
    clear_error();
    LFVIO_REQUIRE(t && out, "lfvio_tracker_get: null argument");
    LFVIO_REQUIRE(t->mirror_fresh, "lfvio_tracker_get: host mirror is stale; call lfvio_tracker_sync first");
    const int n = t->n_host;
    LFVIO_REQUIRE(out->capacity >= n, "lfvio_tracker_get: output capacity too small");
    out->n = n;
    if (out->cur_pts) memcpy(out->cur_pts, t->h_pts.p, sizeof(float) * 2 * n);
    if (out->ids) memcpy(out->ids, t->h_ids.p, sizeof(int) * n);
    if (out->track_cnt) memcpy(out->track_cnt, t->h_cnt.p, sizeof(int) * n);
    if (out->cur_un_pts_3d) memcpy(out->cur_un_pts_3d, t->h_bearing.p, sizeof(float) * 3 * n);
    if (out->pts_velocity_3d) memcpy(out->pts_velocity_3d, t->h_vel.p, sizeof(float) * 3 * n);
    return LFVIO_OK;
}

extern "C" int lfvio_tracker_update_ids(lfvio_tracker *t, int *n_id) {
    clear_error();
    LFVIO_REQUIRE(t && n_id, "lfvio_tracker_update_ids: null argument");
    LFVIO_REQUIRE(t->mirror_fresh, "lfvio_tracker_update_ids: host mirror is stale; call lfvio_tracker_sync first");
    LFVIO_CUDA_TRY(cudaSetDevice(t->device));
    if (*n_id != t->n_id_host) {
        t->n_id_host = *n_id;
        LFVIO_CUDA_TRY(cudaMemcpyAsync(t->n_id_dev.p, n_id, sizeof(int), cudaMemcpyHostToDevice, t->stream));
        LFVIO_CUDA_TRY(cudaStreamSynchronize(t->stream));
    }
    LFVIO_TRY(launch_update_ids(t->live().ids, t->n_dev.p, t->cap, t->n_id_dev.p, t->stream));  // stays async
    for (int i = 0; i < t->n_host; i++)
        if (t->h_ids.p[i] == -1) t->h_ids.p[i] = t->n_id_host++;
    *n_id = t->n_id_host;
    return LFVIO_OK;
}

extern "C" int lfvio_tracker_get_mask(lfvio_tracker *t, uint8_t *out, int out_stride) {
    clear_error();
    LFVIO_REQUIRE(t && out && out_stride >= t->cfg.cols, "lfvio_tracker_get_mask: bad arguments");
    LFVIO_CUDA_TRY(cudaSetDevice(t->device));
    LFVIO_CUDA_TRY(cudaStreamSynchronize(t->stream));
    LFVIO_CUDA_TRY(cudaMemcpy2D(out, out_stride, t->mask.p, t->pitch, t->cfg.cols, t->cfg.rows, cudaMemcpyDeviceToHost));
    return LFVIO_OK;
}

extern "C" int lfvio_tracker_get_image(lfvio_tracker *t, int level, uint8_t *out, int out_stride) {
    clear_error();
    LFVIO_REQUIRE(t && out, "lfvio_tracker_get_image: null argument");
    const Pyramid &p = t->pyr[t->cur_slot];
    LFVIO_REQUIRE(level >= 0 && level < p.levels && out_stride >= p.lvl[level].cols, "lfvio_tracker_get_image: bad level/stride");
    LFVIO_CUDA_TRY(cudaSetDevice(t->device));
    LFVIO_CUDA_TRY(cudaStreamSynchronize(t->stream));
    LFVIO_CUDA_TRY(cudaMemcpy2D(out, out_stride, p.lvl[level].data, p.lvl[level].pitch, p.lvl[level].cols, p.lvl[level].rows,
                                cudaMemcpyDeviceToHost));
    return LFVIO_OK;
}

extern "C" int lfvio_tracker_set_points(lfvio_tracker *t, const float *cur_pts, const int *ids, const int *track_cnt, int n) {
    clear_error();
    LFVIO_REQUIRE(t && n >= 0 && n <= t->cap, "lfvio_tracker_set_points: bad arguments");
    LFVIO_REQUIRE(n == 0 || (cur_pts && ids && track_cnt), "lfvio_tracker_set_points: null array");
    LFVIO_CUDA_TRY(cudaSetDevice(t->device));
    LFVIO_CUDA_TRY(cudaStreamSynchronize(t->stream));
    TrackArrays A = t->live();
    if (n > 0) {
        LFVIO_CUDA_TRY(cudaMemcpy(A.cur, cur_pts, sizeof(float) * 2 * n, cudaMemcpyHostToDevice));
        LFVIO_CUDA_TRY(cudaMemcpy(A.ids, ids, sizeof(int) * n, cudaMemcpyHostToDevice));
        LFVIO_CUDA_TRY(cudaMemcpy(A.cnt, track_cnt, sizeof(int) * n, cudaMemcpyHostToDevice));
        LFVIO_CUDA_TRY(cudaMemset(A.prev_valid, 0, n));
        memcpy(t->h_pts.p, cur_pts, sizeof(float) * 2 * n);
        memcpy(t->h_ids.p, ids, sizeof(int) * n);
        memcpy(t->h_cnt.p, track_cnt, sizeof(int) * n);
    }
    LFVIO_CUDA_TRY(cudaMemcpy(t->n_dev.p, &n, sizeof(int), cudaMemcpyHostToDevice));
    t->n_host = t->n_upper = n;
    return LFVIO_OK;
}

extern "C" void *lfvio_tracker_stream(lfvio_tracker *t) { return t ? (void *)t->stream : nullptr; }

extern "C" int lfvio_tracker_stage_count(void) { return TS_COUNT; }
extern "C" const char *lfvio_tracker_stage_name(int i) { return (i >= 0 && i < TS_COUNT) ? kTrackerStageNames[i] : ""; }
extern "C" int lfvio_tracker_profile_begin(lfvio_tracker *t, int max_frames) {
    clear_error();
    LFVIO_REQUIRE(t && max_frames > 0, "lfvio_tracker_profile_begin: bad arguments");
    LFVIO_CUDA_TRY(cudaSetDevice(t->device));
    return t->prof.enable(TS_COUNT, max_frames * TS_COUNT);
}
extern "C" int lfvio_tracker_profile_end(lfvio_tracker *t, float *stage_ms, int *stage_records) {
    clear_error();
    LFVIO_REQUIRE(t && stage_ms && stage_records, "lfvio_tracker_profile_end: bad arguments");
    LFVIO_CUDA_TRY(cudaSetDevice(t->device));
    LFVIO_CUDA_TRY(cudaStreamSynchronize(t->stream));
    LFVIO_TRY(t->prof.collect(stage_ms, stage_records));
    t->prof.disable();
    return LFVIO_OK;
}

// ================================================================================================
// Stage-level entry points (host buffers, synchronous)
// ================================================================================================
namespace {

struct DevImage {
    DevBuf<uint8_t> buf;
    int rows = 0, cols = 0, pitch = 0;
    int upload(const uint8_t *src, int r, int c, int stride) {
        rows = r;
        cols = c;
        pitch = align_up(c, 128);
        LFVIO_TRY(buf.alloc((size_t)pitch * r));
        if (src) LFVIO_CUDA_TRY(cudaMemcpy2D(buf.p, pitch, src, stride, c, r, cudaMemcpyHostToDevice));
        return LFVIO_OK;
    }
    int download(uint8_t *dst, int stride) {
        LFVIO_CUDA_TRY(cudaMemcpy2D(dst, stride, buf.p, pitch, cols, rows, cudaMemcpyDeviceToHost));
        return LFVIO_OK;
    }
};

}  // namespace

extern "C" int lfvio_clahe(const uint8_t *img, int rows, int cols, int stride, double clip_limit, uint8_t *out,
                           int out_stride, int device) {
    clear_error();
    LFVIO_REQUIRE(img && out && rows > 0 && cols > 0 && stride >= cols && out_stride >= cols, "lfvio_clahe: bad arguments");
    LFVIO_CUDA_TRY(cudaSetDevice(device));
    DevImage src, dst;
    LFVIO_TRY(src.upload(img, rows, cols, stride));
    LFVIO_TRY(dst.upload(nullptr, rows, cols, 0));
    DevBuf<uint32_t> hist;
    DevBuf<uint8_t> lut;
    LFVIO_TRY(hist.alloc(kClaheTiles * kClaheTiles * 256));
    LFVIO_TRY(lut.alloc(kClaheTiles * kClaheTiles * 256));
    LFVIO_TRY(launch_clahe(src.buf.p, src.pitch, dst.buf.p, dst.pitch, rows, cols, (float)clip_limit, hist.p, lut.p,
                           sm_count(device), 0));
    LFVIO_CUDA_TRY(cudaDeviceSynchronize());
    return dst.download(out, out_stride);
}

extern "C" int lfvio_pyr_down(const uint8_t *img, int rows, int cols, int stride, uint8_t *out, int out_stride, int device) {
    clear_error();
    LFVIO_REQUIRE(img && out && rows > 0 && cols > 0 && stride >= cols && out_stride >= (cols + 1) / 2, "lfvio_pyr_down: bad arguments");
    LFVIO_CUDA_TRY(cudaSetDevice(device));
    DevImage src, dst;
    LFVIO_TRY(src.upload(img, rows, cols, stride));
    LFVIO_TRY(dst.upload(nullptr, (rows + 1) / 2, (cols + 1) / 2, 0));
    ImgLevel a{src.buf.p, rows, cols, src.pitch}, b{dst.buf.p, dst.rows, dst.cols, dst.pitch};
    LFVIO_TRY(launch_pyr_down(a, b, 0));
    LFVIO_CUDA_TRY(cudaDeviceSynchronize());
    return dst.download(out, out_stride);
}

extern "C" int lfvio_calc_optical_flow_pyr_lk(const uint8_t *prev, const uint8_t *next, int rows, int cols, int stride,
                                              const float *prev_pts, int n, float *next_pts, uint8_t *status,
                                              int max_level, int device) {
    clear_error();
    LFVIO_REQUIRE(prev && next && rows > 41 && cols > 41 && stride >= cols && n >= 0, "lfvio_calc_optical_flow_pyr_lk: bad arguments");
    LFVIO_REQUIRE(max_level >= 0 && max_level < kMaxPyrLevels, "lfvio_calc_optical_flow_pyr_lk: bad max_level");
    if (n == 0) return LFVIO_OK;
    LFVIO_REQUIRE(prev_pts && next_pts && status, "lfvio_calc_optical_flow_pyr_lk: null array");
    LFVIO_CUDA_TRY(cudaSetDevice(device));
    Pyramid pa, pb;
    std::vector<DevBuf<uint8_t>> sa, sb;
    LFVIO_TRY(alloc_pyramid(&pa, &sa, rows, cols, max_level, 41));
    LFVIO_TRY(alloc_pyramid(&pb, &sb, rows, cols, max_level, 41));
    LFVIO_CUDA_TRY(cudaMemcpy2D(pa.lvl[0].data, pa.lvl[0].pitch, prev, stride, cols, rows, cudaMemcpyHostToDevice));
    LFVIO_CUDA_TRY(cudaMemcpy2D(pb.lvl[0].data, pb.lvl[0].pitch, next, stride, cols, rows, cudaMemcpyHostToDevice));
    LFVIO_TRY(launch_build_pyramid(pa, 0));
    LFVIO_TRY(launch_build_pyramid(pb, 0));
    DevBuf<float2> p0, p1;
    DevBuf<uint8_t> st;
    LFVIO_TRY(p0.alloc(n));
    LFVIO_TRY(p1.alloc(n));
    LFVIO_TRY(st.alloc(n));
    LFVIO_CUDA_TRY(cudaMemcpy(p0.p, prev_pts, sizeof(float2) * n, cudaMemcpyHostToDevice));
    const LkParams prm = default_lk(rows, cols, 0);
    PyrTmaps tma;
    make_pyramid_tmaps(pa, &tma);
    LFVIO_TRY(launch_lk(pa, tma, pb, p0.p, p1.p, st.p, n, nullptr, prm, 0));
    LFVIO_CUDA_TRY(cudaDeviceSynchronize());
    LFVIO_CUDA_TRY(cudaMemcpy(next_pts, p1.p, sizeof(float2) * n, cudaMemcpyDeviceToHost));
    LFVIO_CUDA_TRY(cudaMemcpy(status, st.p, n, cudaMemcpyDeviceToHost));
    return LFVIO_OK;
}

extern "C" int lfvio_corner_min_eigen_val(const uint8_t *img, int rows, int cols, int stride, float *out, int device) {
    clear_error();
    LFVIO_REQUIRE(img && out && rows > 0 && cols > 0 && stride >= cols, "lfvio_corner_min_eigen_val: bad arguments");
    LFVIO_CUDA_TRY(cudaSetDevice(device));
    DevImage src;
    LFVIO_TRY(src.upload(img, rows, cols, stride));
    DevBuf<float> eig;
    DevBuf<uint32_t> mx;
    LFVIO_TRY(eig.alloc((size_t)rows * cols));
    LFVIO_TRY(mx.alloc(1));
    CUtensorMap tm;
    const bool tm_ok = encode_u8_tmap(&tm, src.buf.p, rows, cols, src.pitch, 96, 68);
    LFVIO_TRY(launch_min_eig(src.buf.p, src.pitch, nullptr, 0, eig.p, rows, cols, mx.p, 0, tm_ok ? &tm : nullptr));
    LFVIO_CUDA_TRY(cudaDeviceSynchronize());
    LFVIO_CUDA_TRY(cudaMemcpy(out, eig.p, sizeof(float) * (size_t)rows * cols, cudaMemcpyDeviceToHost));
    return LFVIO_OK;
}

extern "C" int lfvio_good_features_to_track(const uint8_t *img, int rows, int cols, int stride, const uint8_t *mask,
                                            int mask_stride, int max_corners, double quality, double min_distance,
                                            float *corners, int *n_out, int device) {
    clear_error();
    LFVIO_REQUIRE(img && corners && n_out && rows > 0 && cols > 0 && stride >= cols && max_corners > 0,
                  "lfvio_good_features_to_track: bad arguments");
    LFVIO_REQUIRE(mask == nullptr || mask_stride >= cols, "lfvio_good_features_to_track: bad mask stride");
    LFVIO_CUDA_TRY(cudaSetDevice(device));
    DevImage src, msk;
    LFVIO_TRY(src.upload(img, rows, cols, stride));
    if (mask) LFVIO_TRY(msk.upload(mask, rows, cols, mask_stride));
    CornerStore cs;
    const int cell = (int)lrint(min_distance);  // cvRound(minDistance)
    LFVIO_TRY(cs.alloc(rows, cols, cell));
    DevBuf<float2> outp;
    LFVIO_TRY(outp.alloc(max_corners));
    CUtensorMap tm;
    const bool tm_ok = encode_u8_tmap(&tm, src.buf.p, rows, cols, src.pitch, 96, 68);
    LFVIO_TRY(launch_min_eig(src.buf.p, src.pitch, mask ? msk.buf.p : nullptr, msk.pitch, cs.eig.p, rows, cols, cs.max_bits.p, 0, tm_ok ? &tm : nullptr));
    LFVIO_TRY(launch_corner_candidates(cs.eig.p, mask ? msk.buf.p : nullptr, msk.pitch, rows, cols, cs.max_bits.p, quality, cs.cs, 0));
    LFVIO_TRY(launch_corner_select(cs.cs, cols, (float)min_distance, max_corners, nullptr, outp.p, nullptr, nullptr, nullptr,
                                   max_corners, 0));
    LFVIO_CUDA_TRY(cudaDeviceSynchronize());
    int counters[8];
    LFVIO_CUDA_TRY(cudaMemcpy(counters, cs.counters.p, sizeof(counters), cudaMemcpyDeviceToHost));
    if (counters[1]) {
        set_error("corner candidate buffer overflow (%d > %d)", counters[0], cs.cs.cand_cap);
        return LFVIO_ERR_CAPACITY;
    }
    *n_out = counters[2];
    if (*n_out > 0) LFVIO_CUDA_TRY(cudaMemcpy(corners, outp.p, sizeof(float2) * (*n_out), cudaMemcpyDeviceToHost));
    return LFVIO_OK;
}
