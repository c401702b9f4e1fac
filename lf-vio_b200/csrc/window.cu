// Estimator::optimization() handle: device residency, upload/download, launch sequence and the back-end
// C-ABI entry points of include/lfvio.h.
//
// Reference: vins_estimator/src/estimator.cpp:676-1009.  Layout in HBM: one contiguous "problem blob"
// (projection factor records | IMU factor records | landmark->factor index | prior x0, J0, r0 | initial state)
// uploaded with a single H2D copy from pinned memory; two state buffers and two evaluation buffers
// (current / candidate) so an accepted step is a flag flip; the reduced system and dogleg vectors in one
// scratch allocation.  The trust-region loop runs as a fixed launch sequence with a device-side control block
// (no host synchronisation until the result is read back).
#include <cstdlib>
#include <cstring>
#include <vector>

#include <map>

#include <time.h>

#include "backend.cuh"
#include "marg.cuh"

namespace lfvio {
int launch_solver_init(SolverCtl *ctl, const DevProblem &P, cudaStream_t s);
}  // namespace lfvio

using namespace lfvio;

namespace {

template <class T>
struct DBuf {
    T *p = nullptr;
    bool own = true;
    void view(void *q) {  // a range of another allocation (not freed here)
        p = static_cast<T *>(q);
        own = false;
    }
    int alloc(size_t n) {
        if (n == 0) n = 1;
        LFVIO_CUDA_TRY(cudaMalloc(&p, sizeof(T) * n));
        LFVIO_CUDA_TRY(cudaMemset(p, 0, sizeof(T) * n));
        return LFVIO_OK;
    }
    ~DBuf() {
        if (p && own) cudaFree(p);
    }
};

size_t align8(size_t v) { return (v + 7) & ~size_t(7); }
size_t align16(size_t v) { return (v + 15) & ~size_t(15); }

}  // namespace

enum WindowStage { WS_PREPARE = 0, WS_EVAL_INITIAL, WS_JACOBI_SCALE, WS_ACCUMULATE, WS_STEP, WS_EVALUATE, WS_MARGINALIZE, WS_COUNT };
static const char *kWindowStageNames[WS_COUNT] = {"prepare(init+imu_prepare+prior_prepare)", "evaluate_kernel(initial)",
                                                  "jacobi_scale(2 kernels)", "accumulate_kernel", "step_kernel",
                                                  "evaluate_kernel(candidate)", "marginalize"};

struct lfvio_window {
    StageProfiler prof;
    int device = 0, F = 0, Lmax = 0, Rmax = 0, nmax = 0;
    cudaStream_t stream = nullptr;
    cudaStream_t side = nullptr;  // second branch of the per-upload constants (prior_prepare next to imu_prepare)
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    // blob offsets (bytes)
    size_t off_proj = 0, off_P = 0, off_imu = 0, off_lm = 0, off_px0 = 0, off_pJ0 = 0, off_pr0 = 0, off_x = 0, blob_bytes = 0;
    DBuf<uint8_t> blob;
    uint8_t *h_blob = nullptr;  // pinned
    DBuf<double> x[2], x_out, imu_sqrt_info, H0;
    DBuf<double> rp[2], Jp[2], ri[2], Ji[2], rm[2];
    DBuf<double> scratch;
    SolverScratch w;
    DBuf<SolverCtl> ctl;
    DBuf<uint8_t> out_blob;      // [status | ctl | x_out | marg_out]: what a call returns, one D2H copy
    uint8_t *h_out = nullptr;    // its pinned mirror; the four pointers below look into it
    size_t out_marg_off = 0;
    SolverCtl *h_ctl = nullptr;  // pinned
    double *h_x = nullptr;       // pinned
    int *h_status = nullptr;     // pinned (tail of h_x): imu_prepare status words
    DBuf<int> status;
    DevProblem P;            // host mirror (launch geometry, sizes of the copies)
    DevProblem *P_dev = nullptr;  // what the kernels read: the header of the blob
    EvalBuf e[2];
    bool uploaded = false;
    // One executable graph per launch geometry (g_* buckets, step variant, marginalisation flag ...), captured the first
    // time that geometry occurs: a solve is then one cudaGraphLaunch instead of ~35 launches.
    struct GraphEntry {
        cudaGraphExec_t exec = nullptr;
        int kernels = 0;
    };
    std::map<unsigned long long, GraphEntry> graphs;
    bool use_graphs = true;
    long graph_launches = 0, graph_captures = 0;
    DBuf<MargPlan> plan_dev[2];
    MargPlan *h_plan = nullptr;  // pinned staging, 2 entries
    // marginalisation (K14)
    MargPlan plan[2];        // [LFVIO_MARGIN_OLD], [LFVIO_MARGIN_SECOND_NEW], rebuilt at every upload
    bool plan_valid[2] = {false, false};
    DBuf<int> lm_list[2];
    DBuf<double> margT, marg_work, marg_out;  // marg_out: [J0 n^2 | r0 n | x0] packed for the plan's n, so that one D2H copy returns the prior
    double *h_marg = nullptr;  // pinned: [J0 nmax^2 | r0 nmax | x0 9*blocks]
    int *h_lm[2] = {nullptr, nullptr};  // pinned staging of lm_list
    int last_marg_flag = -1;

    ~lfvio_window() {
        for (auto &kv : graphs)
            if (kv.second.exec) cudaGraphExecDestroy(kv.second.exec);
        if (h_plan) cudaFreeHost(h_plan);
        if (h_blob) cudaFreeHost(h_blob);
        if (h_out) cudaFreeHost(h_out);
        for (int k = 0; k < 2; k++)
            if (h_lm[k]) cudaFreeHost(h_lm[k]);
    }
};

namespace {

int window_init(lfvio_window *w, int window_size, int max_landmarks, int max_proj, int device) {
    w->device = device;
    LFVIO_CUDA_TRY(cudaSetDevice(device));
    {   // highest priority: the solve is a chain of small dependent kernels; a CTA of it should never queue behind the
        // image-sized grids of a tracker that shares the GPU
        int lo = 0, hi = 0;
        LFVIO_CUDA_TRY(cudaDeviceGetStreamPriorityRange(&lo, &hi));
        LFVIO_CUDA_TRY(cudaStreamCreateWithPriority(&w->stream, cudaStreamNonBlocking, hi));
        LFVIO_CUDA_TRY(cudaStreamCreateWithPriority(&w->side, cudaStreamNonBlocking, hi));
        LFVIO_CUDA_TRY(cudaEventCreateWithFlags(&w->ev_fork, cudaEventDisableTiming));
        LFVIO_CUDA_TRY(cudaEventCreateWithFlags(&w->ev_join, cudaEventDisableTiming));
    }
    const int F = window_size + 1;
    w->F = F;
    w->Lmax = max_landmarks;
    w->Rmax = max_proj;
    w->nmax = 15 * F + 7;
    const int NI = F - 1;
    StateLayout lay{F, max_landmarks};
    // One block, one DMA per upload.  The projection factors END where the fixed-size part begins (their start moves with their
    // count), the prior's n x n Jacobian follows it (a prefix is used): whatever the sizes, what an upload sends is ONE contiguous
    // range.  The problem header the kernels read (DevProblem) leads the fixed part.
    size_t o = 0;
    w->off_proj = o;  // start of the region (where a device FeatureManager writes its list); host lists end at off_P
    o = align16(o + sizeof(lfvio_proj_factor) * (size_t)max_proj);
    w->off_P = o;
    o = align16(o + sizeof(DevProblem));
    w->off_imu = o;
    o = align16(o + sizeof(lfvio_imu_factor) * (size_t)NI);
    w->off_px0 = o;
    o = align16(o + sizeof(double) * 9 * (size_t)kMaxPriorBlocks);
    w->off_pr0 = o;
    o = align16(o + sizeof(double) * (size_t)w->nmax);
    w->off_x = o;
    o = align16(o + sizeof(double) * (size_t)lay.size());
    w->off_lm = o;
    o = align16(o + sizeof(int) * (size_t)(max_landmarks + 1));
    w->off_pJ0 = o;
    o = align16(o + sizeof(double) * (size_t)w->nmax * w->nmax);
    w->blob_bytes = o;
    LFVIO_TRY(w->blob.alloc(o));
    w->P_dev = reinterpret_cast<DevProblem *>(w->blob.p + w->off_P);
    {   // everything a call returns, in one block (one DMA back): [status | control block | state | next prior (prefix used)]
        size_t q = 0;
        const size_t o_status = q;
        q = align16(q + 4 * sizeof(int));
        const size_t o_ctl = q;
        q = align16(q + sizeof(SolverCtl));
        const size_t o_x = q;
        q = align16(q + sizeof(double) * (size_t)lay.size());
        w->out_marg_off = q;
        q = align16(q + sizeof(double) * ((size_t)w->nmax * w->nmax + w->nmax + 9 * (size_t)kMaxPriorBlocks));
        LFVIO_TRY(w->out_blob.alloc(q));
        LFVIO_CUDA_TRY(cudaMallocHost(&w->h_out, q));
        memset(w->h_out, 0, q);
        w->status.view(w->out_blob.p + o_status);
        w->ctl.view(w->out_blob.p + o_ctl);
        w->x_out.view(w->out_blob.p + o_x);
        w->marg_out.view(w->out_blob.p + w->out_marg_off);
        w->h_status = reinterpret_cast<int *>(w->h_out + o_status);
        w->h_ctl = reinterpret_cast<SolverCtl *>(w->h_out + o_ctl);
        w->h_x = reinterpret_cast<double *>(w->h_out + o_x);
        w->h_marg = reinterpret_cast<double *>(w->h_out + w->out_marg_off);
    }
    LFVIO_CUDA_TRY(cudaMallocHost(&w->h_blob, o));
    memset(w->h_blob, 0, o);
    for (int k = 0; k < 2; k++) {
        LFVIO_TRY(w->x[k].alloc(lay.size()));
        LFVIO_TRY(w->rp[k].alloc(2 * (size_t)max_proj));
        LFVIO_TRY(w->Jp[k].alloc(40 * (size_t)max_proj));
        LFVIO_TRY(w->ri[k].alloc(15 * (size_t)NI));
        LFVIO_TRY(w->Ji[k].alloc(450 * (size_t)NI));
        LFVIO_TRY(w->rm[k].alloc(w->nmax));
        w->e[k] = EvalBuf{w->rp[k].p, w->Jp[k].p, w->ri[k].p, w->Ji[k].p, w->rm[k].p};
    }
    LFVIO_TRY(w->imu_sqrt_info.alloc(225 * (size_t)NI));
    LFVIO_TRY(w->H0.alloc((size_t)w->nmax * w->nmax));
    // scratch: [S D*D | rhs D | diag2 D | grad D | lm_c L | lm_g L] (zeroed per iteration) | lm_ws | proj_wj | scale | ...
    const size_t D = lay.D(), L = max_landmarks, N = D + L;
    const size_t n_scr = D * D + 3 * D + 2 * L + 13 * L + 6 * (size_t)max_proj + 5 * N + D + (L + (size_t)max_proj) / 2 + 2;
    LFVIO_TRY(w->scratch.alloc(n_scr));
    double *p = w->scratch.p;
    SolverScratch &s = w->w;
    s.S = p;          p += D * D;
    s.rhs = p;        p += D;
    s.diag2 = p;      p += D;
    s.grad = p;       p += D;
    s.lm_c = p;       p += L;
    s.lm_g = p;       p += L;
    s.zero_bytes = sizeof(double) * (size_t)(p - s.S);
    s.lm_ws = p;      p += 13 * L;
    s.proj_wj = p;    p += 6 * (size_t)max_proj;
    s.scale = p;      p += N;
    s.diagonal = p;   p += N;
    s.gradient = p;   p += N;
    s.gn = p;         p += N;
    s.step = p;       p += N;
    s.y = p;          p += D;
    s.lm_coli = reinterpret_cast<int *>(p);
    s.proj_colj = s.lm_coli + L;
    LFVIO_CUDA_TRY(cudaMallocHost(&w->h_plan, 2 * sizeof(MargPlan)));
    for (int k = 0; k < 2; k++) LFVIO_TRY(w->plan_dev[k].alloc(1));
    {
        const char *e = getenv("LFVIO_NO_GRAPH"), *e2 = getenv("LFVIO_NO_GRAPH_WINDOW");
        w->use_graphs = !(e && e[0] == '1') && !(e2 && e2[0] == '1');
    }
    {
        const size_t dim = (size_t)w->nmax + 15;
        for (int k = 0; k < 2; k++) {
            LFVIO_TRY(w->lm_list[k].alloc(max_landmarks));
            LFVIO_CUDA_TRY(cudaMallocHost(&w->h_lm[k], sizeof(int) * (size_t)max_landmarks));
        }
        LFVIO_TRY(w->margT.alloc(dim * dim + dim));
        LFVIO_TRY(w->marg_work.alloc(marg_finish_workspace_doubles(15, w->nmax)));
    }
    return LFVIO_OK;
}

// Launch geometry of the plan + its device copy (the kernels read the plan from device memory; an empty plan has n = 0 and
// makes both marginalisation kernels return at once).
int publish_plan(lfvio_window *w, int flag) {
    MargPlan &pl = w->plan[flag];
    const int cap_small = 6 * w->F + 10;  // 6 (F - 1) + 9 + 6 + 1: the largest prior the estimator itself produces
    pl.g_cap_n = (pl.n <= cap_small && w->P.prior.n <= cap_small) ? cap_small : w->nmax;
    pl.use_smem = marg_workspace_in_smem(pl.g_cap_n) ? 1 : 0;
    pl.g_items = std::min(w->Lmax + 1, (pl.n_lm + 1 + 63) / 64 * 64);
    pl.self_dev = w->plan_dev[flag].p;
    w->h_plan[flag] = pl;
    LFVIO_CUDA_TRY(cudaMemcpyAsync(w->plan_dev[flag].p, &w->h_plan[flag], sizeof(MargPlan), cudaMemcpyHostToDevice, w->stream));
    return LFVIO_OK;
}

// Which blocks are dropped / kept and in which column order (estimator.cpp:833-1005, marginalization_factor.cpp:174-205),
// canonical order by block id instead of the reference's address-keyed unordered_map.
int build_marg_plan(lfvio_window *w, int flag, const lfvio_factor_list *fl, const lfvio_solver_opts *o, int L,
                    const DevFactorSrc *src = nullptr) {
    MargPlan &pl = w->plan[flag];
    memset(&pl, 0, sizeof(pl));
    w->plan_valid[flag] = false;
    pl.flag = flag;
    pl.imu_index = -1;
    const int F = w->F;
    const StateLayout lay{F, L};
    for (int i = 0; i < kMaxPriorBlocks; i++) pl.col[i] = -1;
    std::map<int, int> is_drop;
    auto touch = [&](int id, bool drop) {
        auto it = is_drop.find(id);
        if (it == is_drop.end()) is_drop[id] = drop ? 1 : 0;
        else if (drop) it->second = 1;
    };
    const bool have_prior = fl->prior && fl->prior->n > 0;
    std::vector<int> lms;
    if (flag == LFVIO_MARGIN_OLD) {
        if (have_prior) {
            pl.use_prior = 1;
            for (int b = 0; b < fl->prior->n_blocks; b++) touch(fl->prior->block_id[b], fl->prior->block_id[b] <= 1);
        }
        for (int k = 0; k < fl->n_imu; k++)
            if (fl->imu[k].frame_i == 0 && fl->imu[k].sum_dt < 10.0) {
                pl.imu_index = k;
                touch(0, true);
                touch(1, true);
                touch(2, false);
                touch(3, false);
            }
        int last_l = -1;
        if (src) {  // the list was built on the device: which frames the landmarks hosted in frame 0 are seen from
            for (int j = 1; j < F; j++)
                if (src->frame_j_mask >> j & 1u) {
                    touch(0, true);
                    touch(2 * j, false);
                    touch(2 * F, false);
                    if (o->estimate_td) touch(2 * F + 1, false);
                }
        }
        for (int k = 0; !src && k < fl->n_proj; k++)
            if (fl->proj[k].frame_i == 0) {
                touch(0, true);
                touch(2 * fl->proj[k].frame_j, false);
                touch(2 * F, false);
                if (o->estimate_td) touch(2 * F + 1, false);
                if (fl->proj[k].landmark != last_l) lms.push_back(last_l = fl->proj[k].landmark);
            }
    } else {
        const int drop_id = 2 * (F - 2);  // para_Pose[WINDOW_SIZE - 1]
        bool found = false;
        if (have_prior)
            for (int b = 0; b < fl->prior->n_blocks; b++) found |= (fl->prior->block_id[b] == drop_id);
        if (!found) return publish_plan(w, flag);  // nothing to do (estimator.cpp:942-943)
        pl.use_prior = 1;
        for (int b = 0; b < fl->prior->n_blocks; b++) touch(fl->prior->block_id[b], fl->prior->block_id[b] == drop_id);
    }
    if (is_drop.empty()) return publish_plan(w, flag);
    int pos = 0;
    for (auto &kv : is_drop)
        if (kv.second) {
            pl.col[kv.first] = pos;
            pos += lay.block_lsize(kv.first);
        }
    pl.m_p = pos;
    LFVIO_REQUIRE(pl.m_p <= 15, "marginalize: more than one pose + speed-bias block to drop");
    int off = 0;
    for (auto &kv : is_drop)
        if (!kv.second) {
            pl.col[kv.first] = pos;
            pos += lay.block_lsize(kv.first);
            pl.keep_id[pl.n_keep] = kv.first;
            pl.keep_x0[pl.n_keep] = off;
            off += lay.block_gsize(kv.first);
            pl.n_keep++;
        }
    pl.n = pos - pl.m_p;
    pl.dim = pos;
    pl.n_lm = (src && flag == LFVIO_MARGIN_OLD) ? src->n_marg_lm : (int)lms.size();
    pl.lm_list = w->lm_list[flag].p;
    if (!lms.empty()) {  // staged through pinned memory: no synchronisation on the upload path
        memcpy(w->h_lm[flag], lms.data(), sizeof(int) * lms.size());
        LFVIO_CUDA_TRY(cudaMemcpyAsync(w->lm_list[flag].p, w->h_lm[flag], sizeof(int) * lms.size(), cudaMemcpyHostToDevice, w->stream));
    }
    w->plan_valid[flag] = pl.n > 0;
    return publish_plan(w, flag);
}

static MargBuffers marg_buffers(lfvio_window *w, int n) {
    double *base = w->marg_out.p;
    return MargBuffers{w->margT.p, w->marg_work.p, base, base + (size_t)n * n, base + (size_t)n * n + n};
}

// (an empty plan has n = 0 on the device and both kernels return at once: the launch sequence does not depend on it)
int enqueue_marginalize(lfvio_window *w, int flag, bool after_solve, const double *x0 = nullptr, const double *x1 = nullptr) {
    MargBuffers mb = marg_buffers(w, w->plan[flag].n);
    w->prof.begin(WS_MARGINALIZE, w->stream);
    LFVIO_TRY(launch_marginalize(w->P, w->e[0], w->e[1], x0 ? x0 : w->x[0].p, x1 ? x1 : w->x[1].p, after_solve ? w->ctl.p : nullptr,
                                 w->plan[flag], mb, w->stream));
    w->prof.end(w->stream);
    return LFVIO_OK;
}

// Run `body` (a launch sequence on w->stream whose arguments depend only on the handle and on `key`) as an executable
// graph: captured the first time the key occurs, replayed afterwards.  Stage profiling and LFVIO_NO_GRAPH=1 launch directly.
template <class Body>
int run_graphed(lfvio_window *w, unsigned long long key, Body body) {
    if (!w->use_graphs || w->prof.on) return body();
    cudaStream_t s = w->stream;
    auto it = w->graphs.find(key);
    if (it == w->graphs.end()) {  // first occurrence: plain launches (also loads the kernels outside any capture)
        w->graphs.emplace(key, lfvio_window::GraphEntry());
        return body();
    }
    if (it->second.exec == nullptr) {
        cudaGraph_t g = nullptr;
        LFVIO_CUDA_TRY(cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal));
        const unsigned long long l0 = thread_launch_count();
        const int st = body();
        const int n_kernels = (int)(thread_launch_count() - l0);
        const cudaError_t ce = cudaStreamEndCapture(s, &g);
        if (st != LFVIO_OK) {
            if (g) cudaGraphDestroy(g);
            return st;
        }
        LFVIO_CUDA_TRY(ce);
        lfvio_window::GraphEntry e;
        LFVIO_CUDA_TRY(cudaGraphInstantiate(&e.exec, g, 0));
        cudaGraphDestroy(g);
        e.kernels = n_kernels;
        it->second = e;
        w->graph_captures++;
    } else {
        count_launch(it->second.kernels);
    }
    LFVIO_CUDA_TRY(cudaGraphLaunch(it->second.exec, s));
    w->graph_launches++;
    return LFVIO_OK;
}

// everything a launch sequence of this handle depends on besides the handle itself
unsigned long long geometry_key(const lfvio_window *w, int variant, int flag, bool with_plan) {
    const DevProblem &P = w->P;
    unsigned long long k = (unsigned long long)variant;
    k = k * 4 + (unsigned)flag;
    k = k * 1024 + (unsigned)P.g_nbp;
    k = k * 65536 + (unsigned)P.g_L;
    k = k * 128 + (unsigned)P.o.max_num_iterations;
    k = k * 65536 + (with_plan ? (unsigned)w->plan[flag].g_items : 0u);
    k = k * 2 + (with_plan ? (unsigned)(w->plan[flag].g_cap_n == w->nmax) : 0u);
    return k;
}

// Pack the problem into the pinned blob, enqueue the single H2D copy and fill the device-side descriptor.
// defer_plans: the caller builds the marginalisation plan it needs itself AFTER it has enqueued the solve, so that this
// piece of host work (maps, a small H2D copy) overlaps the solve on the device instead of delaying its first kernel.
// Per-upload constants of the solve: square-root information of the IMU factors, H0 = J0^T J0 of the prior.
int enqueue_prepare(lfvio_window *w) {
    cudaStream_t s = w->stream;
    // two independent chains of dependent arithmetic (15 x 15 inverses + Cholesky per IMU factor; H0 = J0^T J0): side by side —
    // a fork / join through events, which a stream capture turns into two branches of the graph
    LFVIO_CUDA_TRY(cudaEventRecord(w->ev_fork, s));
    LFVIO_CUDA_TRY(cudaStreamWaitEvent(w->side, w->ev_fork, 0));
    LFVIO_TRY(launch_prior_prepare(w->P, w->H0.p, w->side));
    LFVIO_CUDA_TRY(cudaEventRecord(w->ev_join, w->side));
    LFVIO_CUDA_TRY(cudaMemsetAsync(w->status.p, 0, sizeof(int) * 4, s));
    LFVIO_TRY(launch_imu_prepare(w->P, w->imu_sqrt_info.p, w->status.p, s));
    LFVIO_CUDA_TRY(cudaStreamWaitEvent(s, w->ev_join, 0));
    return LFVIO_OK;
}

// defer_prepare: the caller enqueues enqueue_prepare() itself (as part of a replayed graph).
int upload(lfvio_window *w, const lfvio_window_state *st, const lfvio_factor_list *fl, const lfvio_solver_opts *o,
           bool defer_plans = false, bool defer_prepare = false, const DevFactorSrc *src = nullptr) {
    LFVIO_REQUIRE(st && fl && o, "window: null argument");
    LFVIO_REQUIRE(st->n_frames == w->F, "window: n_frames != window_size + 1 of the handle");
    const int n_proj = src ? src->n_proj : fl->n_proj;
    const int n_lm = src ? src->n_landmarks : st->n_landmarks;
    LFVIO_REQUIRE(n_lm >= 0 && n_lm <= w->Lmax, "window: too many landmarks for this handle");
    LFVIO_REQUIRE(n_proj >= 0 && n_proj <= w->Rmax, "window: too many projection factors for this handle");
    LFVIO_REQUIRE(fl->n_imu >= 0 && fl->n_imu <= w->F - 1, "window: too many IMU factors");
    LFVIO_REQUIRE(o->max_num_iterations >= 0 && o->max_num_iterations < LFVIO_MAX_ITER_LOG, "window: bad max_num_iterations");
    const int F = w->F, L = n_lm;
    StateLayout lay{F, L};
    uint8_t *hb = w->h_blob;
    const size_t proj_bytes = sizeof(lfvio_proj_factor) * (size_t)n_proj;
    const size_t proj_at = src ? w->off_proj : w->off_P - proj_bytes;  // a host list ends where the fixed part begins
    if (!src && n_proj) memcpy(hb + proj_at, fl->proj, proj_bytes);
    if (fl->n_imu) memcpy(hb + w->off_imu, fl->imu, sizeof(lfvio_imu_factor) * (size_t)fl->n_imu);
    // landmark -> factor ranges (the reference inserts factors landmark by landmark, estimator.cpp:725-772)
    int *lm = reinterpret_cast<int *>(hb + w->off_lm);
    if (!src) {  // (a list built by lfvio_fm_build_factors is grouped and in range by construction)
        int k = 0;
        for (int l = 0; l < L; l++) {
            lm[l] = k;
            while (k < fl->n_proj && fl->proj[k].landmark == l) {
                const lfvio_proj_factor &f = fl->proj[k];
                LFVIO_REQUIRE(f.frame_i >= 0 && f.frame_i < F && f.frame_j >= 0 && f.frame_j < F && f.frame_i != f.frame_j,
                              "window: projection factor frame index out of range");
                LFVIO_REQUIRE(f.frame_i == fl->proj[lm[l]].frame_i, "window: factors of one landmark must share the host frame");
                k++;
            }
            LFVIO_REQUIRE(k - lm[l] <= kMaxObs, "window: too many observations of one landmark");
        }
        lm[L] = k;
        LFVIO_REQUIRE(k == fl->n_proj, "window: projection factors must be grouped by ascending landmark index");
    }
    for (int k = 0; k < fl->n_imu; k++)
        LFVIO_REQUIRE(fl->imu[k].frame_i >= 0 && fl->imu[k].frame_i < F - 1, "window: IMU factor frame index out of range");
    DevProblem &P = w->P;
    memset(&P, 0, sizeof(P));
    P.lay = lay;
    P.o = *o;
    P.n_proj = n_proj;
    P.n_imu = fl->n_imu;
    P.proj = reinterpret_cast<const lfvio_proj_factor *>(w->blob.p + proj_at);
    P.imu = reinterpret_cast<const lfvio_imu_factor *>(w->blob.p + w->off_imu);
    P.lm_first = reinterpret_cast<const int *>(w->blob.p + w->off_lm);
    P.imu_sqrt_info = w->imu_sqrt_info.p;
    if (fl->prior && fl->prior->n > 0) {
        const lfvio_prior &p = *fl->prior;
        LFVIO_REQUIRE(p.n <= w->nmax && p.n_blocks <= kMaxPriorBlocks, "window: prior too large for this handle");
        P.prior.n = p.n;
        P.prior.n_blocks = p.n_blocks;
        int off = 0, lsum = 0;
        for (int b = 0; b < p.n_blocks; b++) {
            const int id = p.block_id[b];
            LFVIO_REQUIRE(id >= 0 && id <= 2 * F + 1, "window: prior block id out of range");
            P.prior.block_id[b] = id;
            P.prior.block_idx[b] = p.block_idx[b];
            P.prior.block_x0[b] = off;
            off += lay.block_gsize(id);
            lsum += lay.block_lsize(id);
            LFVIO_REQUIRE(p.block_idx[b] >= 0 && p.block_idx[b] + lay.block_lsize(id) <= p.n, "window: prior block offset out of range");
        }
        LFVIO_REQUIRE(lsum == p.n, "window: prior size does not match its blocks");
        memcpy(hb + w->off_px0, p.x0, sizeof(double) * off);
        memcpy(hb + w->off_pJ0, p.J0, sizeof(double) * (size_t)p.n * p.n);
        memcpy(hb + w->off_pr0, p.r0, sizeof(double) * p.n);
        P.prior.x0 = reinterpret_cast<const double *>(w->blob.p + w->off_px0);
        P.prior.J0 = reinterpret_cast<const double *>(w->blob.p + w->off_pJ0);
        P.prior.r0 = reinterpret_cast<const double *>(w->blob.p + w->off_pr0);
        P.prior.H0 = w->H0.p;
    }
    // launch geometry: counts rounded up to coarse buckets so that consecutive frames of a stream share one graph
    P.self_dev = w->P_dev;
    P.g_nbp = std::max(1, std::min(cdiv(w->Rmax, 128), (cdiv(n_proj, 128) + 7) / 8 * 8));
    P.g_L = std::max(1, std::min(w->Lmax, (L + 127) / 128 * 128));
    P.g_imu = F - 1;
    P.cap_L = w->Lmax;
    P.cap_n = w->nmax;
    memcpy(hb + w->off_P, &P, sizeof(DevProblem));
    double *hx = reinterpret_cast<double *>(hb + w->off_x);
    memcpy(hx + lay.pose(0), st->para_Pose, sizeof(double) * 7 * F);
    memcpy(hx + lay.sb(0), st->para_SpeedBias, sizeof(double) * 9 * F);
    memcpy(hx + lay.ex(), st->para_Ex_Pose, sizeof(double) * 7);
    hx[lay.td()] = st->para_Td[0];
    if (L && !src) memcpy(hx + lay.feat(0), st->para_Feature, sizeof(double) * L);
    // one DMA: [projection factors | header | IMU factors | prior x0, r0 | state | landmark ranges | prior J0 (n x n prefix)]
    cudaStream_t s = w->stream;
    auto copy = [&](size_t off, size_t bytes) -> int {
        if (bytes) LFVIO_CUDA_TRY(cudaMemcpyAsync(w->blob.p + off, hb + off, bytes, cudaMemcpyHostToDevice, s));
        return LFVIO_OK;
    };
    const size_t j0_bytes = P.prior.n ? sizeof(double) * (size_t)P.prior.n * P.prior.n : 0;
    if (src) {
        // ranges, inverse depths and factors are already in the blob (written there by the FeatureManager's kernel): the header
        // up to the fixed part of the state, then the prior's Jacobian
        LFVIO_TRY(copy(w->off_P, w->off_x + sizeof(double) * lay.feat(0) - w->off_P));
        LFVIO_TRY(copy(w->off_pJ0, j0_bytes));
    } else {
        const size_t end = j0_bytes ? w->off_pJ0 + j0_bytes : w->off_lm + sizeof(int) * ((size_t)L + 1);
        LFVIO_TRY(copy(proj_at, end - proj_at));
    }
    if (!defer_prepare) LFVIO_TRY(enqueue_prepare(w));
    if (defer_plans) {
        w->plan_valid[0] = w->plan_valid[1] = false;
    } else {
        LFVIO_TRY(build_marg_plan(w, LFVIO_MARGIN_OLD, fl, o, L, src));
        LFVIO_TRY(build_marg_plan(w, LFVIO_MARGIN_SECOND_NEW, fl, o, L, src));
    }
    w->uploaded = true;
    return LFVIO_OK;
}

// The whole ceres::Solve as a launch sequence (no host synchronisation).
int enqueue_solve(lfvio_window *w) {
    cudaStream_t s = w->stream;
    const DevProblem &P = w->P;
    const size_t xbytes = sizeof(double) * (size_t)StateLayout{w->F, w->Lmax}.size();  // capacity: same copy every frame
    w->prof.begin(WS_PREPARE, s);
    LFVIO_TRY(launch_solver_init(w->ctl.p, P, s));
    LFVIO_CUDA_TRY(cudaMemcpyAsync(w->x[0].p, w->blob.p + w->off_x, xbytes, cudaMemcpyDeviceToDevice, s));
    w->prof.end(s);
    w->prof.begin(WS_EVAL_INITIAL, s);
    LFVIO_TRY(launch_evaluate(P, w->x[0].p, w->x[1].p, w->e[0], w->e[1], w->ctl.p, 0, 1, s));
    w->prof.end(s);
    w->prof.begin(WS_JACOBI_SCALE, s);
    LFVIO_TRY(launch_jacobi_scale(P, w->e[0], w->w, s));
    w->prof.end(s);
    const bool self_clearing = step_zeroes_system(P);
    if (self_clearing) LFVIO_CUDA_TRY(cudaMemsetAsync(w->w.S, 0, w->w.zero_bytes, s));  // once; chol_kernel keeps it clear
    for (int it = 0; it < P.o.max_num_iterations; it++) {
        w->prof.begin(WS_ACCUMULATE, s);
        LFVIO_TRY(launch_accumulate(P, w->e[0], w->e[1], w->w, w->ctl.p, s, self_clearing));
        w->prof.end(s);
        w->prof.begin(WS_STEP, s);
        LFVIO_TRY(launch_step(P, w->e[0], w->e[1], w->w, w->x[0].p, w->x[1].p, w->ctl.p, s));
        w->prof.end(s);
        w->prof.begin(WS_EVALUATE, s);
        LFVIO_TRY(launch_evaluate(P, w->x[0].p, w->x[1].p, w->e[0], w->e[1], w->ctl.p, 1, 1, s));
        w->prof.end(s);
    }
    return LFVIO_OK;
}

__global__ void select_state_kernel(const DevProblem *__restrict__ Pd, const double *x0, const double *x1, const SolverCtl *ctl,
                                    double *out) {
    const int n = Pd->lay.size();
    const double *x = ctl->cur ? x1 : x0;
    for (int q = blockIdx.x * blockDim.x + threadIdx.x; q < n; q += gridDim.x * blockDim.x) out[q] = x[q];
}

// Estimator::double2vector() gauge re-anchoring (estimator.cpp:532-583) followed by vector2double() (:488-530), on device:
// the optimised window is rotated about gravity so that frame 0 keeps the yaw it had before the solve, and translated so
// that frame 0 keeps its position; poses go through the same quaternion -> matrix -> quaternion round trip as on the host.
__device__ DM3 dev_ypr2R(double y_deg, double p_deg, double r_deg) {
    const double y = y_deg / 180.0 * M_PI, p = p_deg / 180.0 * M_PI, r = r_deg / 180.0 * M_PI;
    DM3 Rz = dm3_identity(), Ry = dm3_identity(), Rx = dm3_identity();
    Rz.m[0][0] = cos(y); Rz.m[0][1] = -sin(y); Rz.m[1][0] = sin(y); Rz.m[1][1] = cos(y);
    Ry.m[0][0] = cos(p); Ry.m[0][2] = sin(p); Ry.m[2][0] = -sin(p); Ry.m[2][2] = cos(p);
    Rx.m[1][1] = cos(r); Rx.m[1][2] = -sin(r); Rx.m[2][1] = sin(r); Rx.m[2][2] = cos(r);
    return Rz * Ry * Rx;
}
__device__ D3 dev_R2ypr(const DM3 &R) {  // Utility::R2ypr, degrees
    const D3 n = d3(R.m[0][0], R.m[1][0], R.m[2][0]), o = d3(R.m[0][1], R.m[1][1], R.m[2][1]), a = d3(R.m[0][2], R.m[1][2], R.m[2][2]);
    const double y = atan2(n.y, n.x);
    const double p = atan2(-n.z, n.x * cos(y) + n.y * sin(y));
    const double r = atan2(a.x * sin(y) - a.y * cos(y), -o.x * sin(y) + o.y * cos(y));
    return d3(y / M_PI * 180.0, p / M_PI * 180.0, r / M_PI * 180.0);
}
__device__ void dev_rot_to_quat(const DM3 &m, double *q /* x y z w */) {  // Eigen::Quaterniond(Matrix3d)
    double t = m.m[0][0] + m.m[1][1] + m.m[2][2];
    if (t > 0) {
        t = sqrt(t + 1.0);
        q[3] = 0.5 * t;
        t = 0.5 / t;
        q[0] = (m.m[2][1] - m.m[1][2]) * t;
        q[1] = (m.m[0][2] - m.m[2][0]) * t;
        q[2] = (m.m[1][0] - m.m[0][1]) * t;
    } else {
        int i = 0;
        if (m.m[1][1] > m.m[0][0]) i = 1;
        if (m.m[2][2] > m.m[i][i]) i = 2;
        const int j = (i + 1) % 3, k = (j + 1) % 3;
        t = sqrt(m.m[i][i] - m.m[j][j] - m.m[k][k] + 1.0);
        q[i] = 0.5 * t;
        t = 0.5 / t;
        q[3] = (m.m[k][j] - m.m[j][k]) * t;
        q[j] = (m.m[j][i] + m.m[i][j]) * t;
        q[k] = (m.m[k][i] + m.m[i][k]) * t;
    }
}
__global__ void gauge_fix_kernel(const DevProblem *__restrict__ Pd, const double *x_init, const double *x0, const double *x1,
                                 const SolverCtl *ctl, double *out) {
    const StateLayout lay = Pd->lay;
    const double *x = ctl->cur ? x1 : x0;
    const int tid = blockIdx.x * blockDim.x + threadIdx.x, nth = gridDim.x * blockDim.x;
    for (int q = lay.ex() + tid; q < lay.size(); q += nth) out[q] = x[q];  // extrinsic, td, inverse depths: unchanged
    if (tid >= lay.F) return;
    D3 P0i, P0o;
    DQ Q0i, Q0o;
    load_pose(x_init + lay.pose(0), &P0i, &Q0i);
    load_pose(x + lay.pose(0), &P0o, &Q0o);
    const DM3 R0_init = qrot(Q0i), R00 = qrot(Q0o);
    const D3 origin_R0 = dev_R2ypr(R0_init), origin_R00 = dev_R2ypr(R00);
    DM3 rot_diff = dev_ypr2R(origin_R0.x - origin_R00.x, 0.0, 0.0);
    if (fabs(fabs(origin_R0.y) - 90) < 1.0 || fabs(fabs(origin_R00.y) - 90) < 1.0) rot_diff = R0_init * transpose3(R00);  // euler singular point
    const int f = tid;
    D3 p;
    DQ qf;
    load_pose(x + lay.pose(f), &p, &qf);
    const DM3 Rf = rot_diff * qrot(qnormalized(qf));
    const D3 Pf = rot_diff * (p - P0o) + P0i;
    const double *sb = x + lay.sb(f);
    const D3 Vf = rot_diff * d3(sb[0], sb[1], sb[2]);
    double *po = out + lay.pose(f), *so = out + lay.sb(f);
    po[0] = Pf.x; po[1] = Pf.y; po[2] = Pf.z;
    dev_rot_to_quat(Rf, po + 3);
    so[0] = Vf.x; so[1] = Vf.y; so[2] = Vf.z;
    for (int k = 3; k < 9; k++) so[k] = sb[k];
}

int download(lfvio_window *w, lfvio_window_state *st, lfvio_solve_summary *sum, bool select_state = true, size_t marg_doubles = 0) {
    cudaStream_t s = w->stream;
    const StateLayout lay = w->P.lay;
    if (select_state) {
        select_state_kernel<<<cdiv(lay.size(), 256), 256, 0, s>>>(w->P.self_dev, w->x[0].p, w->x[1].p, w->ctl.p, w->x_out.p);
        LFVIO_LAUNCH_CHECK();
    }
    // one DMA back: status, control block, state and (marg_doubles > 0) the next prior
    const size_t out_bytes = marg_doubles ? w->out_marg_off + sizeof(double) * marg_doubles
                                          : (size_t)(reinterpret_cast<uint8_t *>(w->h_x) - w->h_out) + sizeof(double) * lay.size();
    LFVIO_CUDA_TRY(cudaMemcpyAsync(w->h_out, w->out_blob.p, out_bytes, cudaMemcpyDeviceToHost, s));
    int *status = w->h_status;
    LFVIO_CUDA_TRY(cudaStreamSynchronize(s));
    if (st) {
        memcpy(st->para_Pose, w->h_x + lay.pose(0), sizeof(double) * 7 * lay.F);
        memcpy(st->para_SpeedBias, w->h_x + lay.sb(0), sizeof(double) * 9 * lay.F);
        memcpy(st->para_Ex_Pose, w->h_x + lay.ex(), sizeof(double) * 7);
        st->para_Td[0] = w->h_x[lay.td()];
        if (lay.L) memcpy(st->para_Feature, w->h_x + lay.feat(0), sizeof(double) * lay.L);
    }
    const SolverCtl &c = *w->h_ctl;
#ifdef LFVIO_CHOL_PROFILE
    if (getenv("LFVIO_DEBUG_STEP")) {
        long long sub[16];
        chol_profile_read(sub);
        static const char *sn[16] = {"w0 panel16", "w0 diag-upd", "w0 factor", "w0 wait", "-", "-", "-", "-", "wk panel-store", "wk wait-panel", "wk trailing", "wk wait", "wk panel-load", "wk panel-solve", "-", "-"};
        for (int i = 0; i < 16; i++) fprintf(stderr, "chol sub-phase %-12s %8.0f cycles\n", sn[i], (double)sub[i]);
    }
#endif
    if (getenv("LFVIO_DEBUG_STEP")) {
        static const char *names[6] = {"cauchy+xnorm", "copy+quadform", "cholesky", "tri-solve", "gn+backsub", "cand_tail"};
        for (int i = 0; i < 6; i++) fprintf(stderr, "step phase %-16s %8.2f us\n", names[i], (c.phase_clk[i + 1] - c.phase_clk[i]) / 1000.0);
        auto us = [&](int a, int b) { return (c.phase_clk[b] - c.phase_clk[a]) / 1000.0; };
        fprintf(stderr, "  cauchy: diag/gradient loop %.2f | gg,gmax sums %.2f | landmark dots %.2f | x_norm %.2f\n", us(0, 7), us(7, 8), us(8, 9), us(9, 1));
        fprintf(stderr, "  copy: panel zero + rhs %.2f | S rows %.2f | jj sum %.2f | ctl + done check %.2f\n", us(1, 10), us(10, 11), us(11, 12), us(12, 2));
        fprintf(stderr, "  gn: reduced columns %.2f | landmark backsub %.2f | sums %.2f\n", us(4, 13), us(13, 14), us(14, 5));
        fprintf(stderr, "  cand_tail: dogleg scalars + candidate %.2f | dx2 sum + publish %.2f\n", us(5, 15), us(15, 6));
    }
    if (sum) {
        memset(sum, 0, sizeof(*sum));
        sum->iterations = c.iter;
        sum->num_successful_steps = c.successful;
        sum->termination = c.done == 1 ? 1 : (c.done < 0 ? -1 : 0);
        sum->initial_cost = c.cost_log[0];
        sum->final_cost = c.cost;
        for (int i = 0; i <= c.iter && i < LFVIO_MAX_ITER_LOG; i++) {
            sum->cost[i] = c.cost_log[i];
            sum->radius[i] = c.radius_log[i];
            sum->accepted[i] = c.accepted_log[i];
        }
    }
    if (status[0]) {
        set_error("IMU covariance of a pre-integration is singular / not positive definite");
        return LFVIO_ERR_NUMERIC;
    }
    if (c.done < 0) {
        set_error("trust-region solve failed (non-finite cost or 5 invalid steps in a row)");
        return LFVIO_ERR_NUMERIC;
    }
    return LFVIO_OK;
}

}  // namespace

extern "C" void lfvio_solver_default_opts(lfvio_solver_opts *o) {
    if (!o) return;
    memset(o, 0, sizeof(*o));
    o->max_num_iterations = 8;   // config/mindvision/mindvision.yaml:134
    o->fixed_iterations = 0;
    o->estimate_td = 1;          // mindvision.yaml:151
    o->estimate_extrinsic = 1;   // mindvision.yaml:84
    o->row = 960;
    o->focal_length = 160.0;     // vins_estimator/src/parameters.h:11
    o->tr = 0.0;
    o->g[0] = 0.0;
    o->g[1] = 0.0;
    o->g[2] = 9.81007;           // mindvision.yaml:142
    o->cauchy_a = 1.0;           // estimator.cpp:681
    o->initial_trust_region_radius = 1e4;  // ceres::Solver::Options defaults (1.12.0)
    o->max_trust_region_radius = 1e16;
    o->min_trust_region_radius = 1e-32;
    o->min_relative_decrease = 1e-3;
    o->function_tolerance = 1e-6;
    o->gradient_tolerance = 1e-10;
    o->parameter_tolerance = 1e-8;
}

extern "C" int lfvio_window_create(int window_size, int max_landmarks, int max_proj, int device, lfvio_window **out) {
    clear_error();
    LFVIO_REQUIRE(out, "lfvio_window_create: null out");
    LFVIO_REQUIRE(window_size >= 2 && window_size + 1 <= kMaxFrames, "lfvio_window_create: window_size out of range");
    LFVIO_REQUIRE(max_landmarks > 0 && max_proj > 0, "lfvio_window_create: capacities must be positive");
    lfvio_window *w = new (std::nothrow) lfvio_window();
    LFVIO_REQUIRE(w, "lfvio_window_create: out of host memory");
    const int st = window_init(w, window_size, max_landmarks, max_proj, device);
    if (st != LFVIO_OK) {
        lfvio_window_destroy(w);
        return st;
    }
    *out = w;
    return LFVIO_OK;
}

extern "C" void lfvio_window_destroy(lfvio_window *w) {
    if (!w) return;
    cudaSetDevice(w->device);
    if (getenv("LFVIO_DEBUG_HOST"))
        fprintf(stderr, "lfvio_window: %d graph keys, %d captures, %d graph launches\n", (int)w->graphs.size(), (int)w->graph_captures, (int)w->graph_launches);
    if (w->stream) {
        cudaStreamSynchronize(w->stream);
        cudaStreamDestroy(w->stream);
    }
    if (w->side) {
        cudaStreamSynchronize(w->side);
        cudaStreamDestroy(w->side);
    }
    if (w->ev_fork) cudaEventDestroy(w->ev_fork);
    if (w->ev_join) cudaEventDestroy(w->ev_join);
    delete w;
}

extern "C" int lfvio_window_optimize(lfvio_window *w, lfvio_window_state *state, const lfvio_factor_list *factors,
                                     const lfvio_solver_opts *opts, lfvio_solve_summary *summary) {
    clear_error();
    LFVIO_REQUIRE(w, "lfvio_window_optimize: null handle");
    LFVIO_CUDA_TRY(cudaSetDevice(w->device));
    LFVIO_TRY(upload(w, state, factors, opts, false, /*defer_prepare=*/true));
    LFVIO_TRY(run_graphed(w, geometry_key(w, 1, 0, false), [&]() -> int {
        LFVIO_TRY(enqueue_prepare(w));
        return enqueue_solve(w);
    }));
    return download(w, state, summary);
}

// Estimator::optimization() in one call: solve -> double2vector/vector2double gauge fix -> next prior, one upload, one sync.
static int optimize_full_impl(lfvio_window *w, lfvio_window_state *state, const lfvio_factor_list *factors, const lfvio_solver_opts *opts,
                              int margin_flag, lfvio_solve_summary *summary, lfvio_prior *prior_out, const DevFactorSrc *src) {
    // LFVIO_DEBUG_HOST=1: where the host time of this call goes (averages printed every 200 calls)
    static const bool dbg_host = getenv("LFVIO_DEBUG_HOST") != nullptr;
    static double t_acc[6] = {0, 0, 0, 0, 0, 0};
    static int t_calls = 0;
    auto now_us = [] { timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec * 1e6 + ts.tv_nsec * 1e-3; };
    double t_prev = dbg_host ? now_us() : 0.0;
    auto lap = [&](int i) { if (dbg_host) { const double t = now_us(); t_acc[i] += t - t_prev; t_prev = t; } };
    LFVIO_REQUIRE(w && prior_out, "lfvio_window_optimize_full: null argument");
    LFVIO_REQUIRE(margin_flag == LFVIO_MARGIN_OLD || margin_flag == LFVIO_MARGIN_SECOND_NEW, "lfvio_window_optimize_full: bad margin_flag");
    LFVIO_CUDA_TRY(cudaSetDevice(w->device));
    // (same switch) device-side timeline of the call: events between its parts, read after the call's own synchronisation
    static cudaEvent_t dev_ev[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
    static double dev_acc[4] = {0, 0, 0, 0};
    if (dbg_host && !dev_ev[0])
        for (auto &e : dev_ev) cudaEventCreate(&e);
    if (dbg_host) cudaEventRecord(dev_ev[0], w->stream);
    LFVIO_TRY(upload(w, state, factors, opts, /*defer_plans=*/true, /*defer_prepare=*/true, src));
    if (dbg_host) cudaEventRecord(dev_ev[1], w->stream);
    lap(0);
    cudaStream_t s = w->stream;
    const DevProblem &P = w->P;
    const int size_cap = StateLayout{w->F, w->Lmax}.size();
    // graph 1: per-upload constants, the solve, the gauge re-anchoring
    LFVIO_TRY(run_graphed(w, geometry_key(w, 2, 0, false), [&]() -> int {
        LFVIO_TRY(enqueue_prepare(w));
        LFVIO_TRY(enqueue_solve(w));
        gauge_fix_kernel<<<cdiv(std::max(size_cap, w->F), 256), 256, 0, s>>>(P.self_dev, reinterpret_cast<const double *>(w->blob.p + w->off_x),
                                                                          w->x[0].p, w->x[1].p, w->ctl.p, w->x_out.p);
        LFVIO_LAUNCH_CHECK();
        return LFVIO_OK;
    }));
    if (dbg_host) cudaEventRecord(dev_ev[2], s);
    lap(1);
    LFVIO_TRY(build_marg_plan(w, margin_flag, factors, opts, w->P.lay.L, src));  // host work under the running solve
    lap(2);
    prior_out->n = 0;
    prior_out->n_blocks = 0;
    const bool do_marg = w->plan_valid[margin_flag];
    const MargPlan &pl = w->plan[margin_flag];
    size_t nn = 0;
    int x0_len = 0;
    if (do_marg) {
        LFVIO_REQUIRE(pl.n <= prior_out->cap_n && pl.n_keep <= prior_out->cap_blocks, "lfvio_window_optimize_full: output prior too small");
        // graph 2: preMarginalize() at the re-anchored state (evaluate into buffer 0), then eliminate
        LFVIO_TRY(run_graphed(w, geometry_key(w, 3, margin_flag, true), [&]() -> int {
            LFVIO_TRY(launch_evaluate(P, w->x_out.p, w->x_out.p, w->e[0], w->e[1], w->ctl.p, 2, 1, s));
            return enqueue_marginalize(w, margin_flag, false, w->x_out.p, w->x_out.p);
        }));
        w->last_marg_flag = margin_flag;
        nn = (size_t)pl.n * pl.n;
        for (int b = 0; b < pl.n_keep; b++) x0_len += P.lay.block_gsize(pl.keep_id[b]);
    }
    if (dbg_host) cudaEventRecord(dev_ev[3], s);
    lap(3);
    LFVIO_TRY(download(w, state, summary, false, do_marg ? nn + pl.n + x0_len : 0));
    lap(4);
    if (dbg_host) {
        cudaEventRecord(dev_ev[4], s);
        cudaEventSynchronize(dev_ev[4]);
        for (int i = 0; i < 4; i++) {
            float ms = 0;
            cudaEventElapsedTime(&ms, dev_ev[i], dev_ev[i + 1]);
            dev_acc[i] += 1e3 * ms;
        }
    }
    if (do_marg && getenv("LFVIO_DEBUG_STEP")) {
        long long mc[8];
        marg_profile_read(mc);
        static const char *names[6] = {"stage T, Amm", "Amm^+", "Tm, A', b'", "blocked chol+tail", "pivoted chol", "write-out"};
        for (int i = 0; i < 6; i++) fprintf(stderr, "marg_finish phase %-14s %8.2f us\n", names[i], (mc[i + 1] - mc[i]) / 1000.0);
        fprintf(stderr, "marg_finish fast path: %lld\n", mc[7]);
    }
    if (do_marg) {
        prior_out->n = pl.n;
        prior_out->n_blocks = pl.n_keep;
        const int F = w->F;
        for (int b = 0; b < pl.n_keep; b++) {
            const int id = pl.keep_id[b];
            int new_id = id;  // addr_shift (estimator.cpp:921-931 / :980-1000)
            if (id < 2 * F) {
                if (margin_flag == LFVIO_MARGIN_OLD) new_id = id - 2;
                else if ((id >> 1) == F - 1) new_id = id - 2;
            }
            prior_out->block_id[b] = new_id;
            prior_out->block_idx[b] = pl.col[id] - pl.m_p;
        }
        memcpy(prior_out->J0, w->h_marg, sizeof(double) * nn);
        memcpy(prior_out->r0, w->h_marg + nn, sizeof(double) * pl.n);
        memcpy(prior_out->x0, w->h_marg + nn + pl.n, sizeof(double) * x0_len);
    }
    lap(5);
    if (dbg_host && ++t_calls % 200 == 0) {
        static const char *nm[6] = {"upload (pack + enqueue H2D)", "enqueue graph 1 (solve)", "marginalisation plan (host)", "enqueue graph 2 + D2H",
                                    "download (wait for the stream + copy out)", "prior copy out"};
        for (int i = 0; i < 6; i++) fprintf(stderr, "optimize_full host %-44s %8.1f us\n", nm[i], t_acc[i] / t_calls);
        static const char *dn[4] = {"H2D (from the call's first enqueue)", "graph 1: prepare + solve + gauge", "plan copy + graph 2: marginalisation", "D2H"};
        for (int i = 0; i < 4; i++) fprintf(stderr, "optimize_full device %-42s %8.1f us\n", dn[i], dev_acc[i] / t_calls);
    }
    return LFVIO_OK;
}

extern "C" int lfvio_window_optimize_full(lfvio_window *w, lfvio_window_state *state, const lfvio_factor_list *factors,
                                          const lfvio_solver_opts *opts, int margin_flag, lfvio_solve_summary *summary,
                                          lfvio_prior *prior_out) {
    clear_error();
    return optimize_full_impl(w, state, factors, opts, margin_flag, summary, prior_out, nullptr);
}

// The same with the projection factors, their landmark ranges and para_Feature taken from a device FeatureManager: its build
// kernel writes them straight into this handle's problem block, the host sees three counts and a frame mask, and the optimised
// inverse depths go back into the manager (FeatureManager::setDepth) without leaving the device.
extern "C" int lfvio_window_optimize_full_fm(lfvio_window *w, lfvio_fm *fm, lfvio_window_state *state, const lfvio_factor_list *factors,
                                             const lfvio_solver_opts *opts, int margin_flag, lfvio_solve_summary *summary,
                                             lfvio_prior *prior_out) {
    clear_error();
    LFVIO_REQUIRE(w && fm && state && factors && opts, "lfvio_window_optimize_full_fm: null argument");
    LFVIO_REQUIRE(fm_window_size(fm) == w->F - 1 && fm_device(fm) == w->device, "lfvio_window_optimize_full_fm: the manager and the window differ in size or device");
    LFVIO_CUDA_TRY(cudaSetDevice(w->device));
    FmBuildDst dst;
    dst.proj = reinterpret_cast<lfvio_proj_factor *>(w->blob.p + w->off_proj);
    dst.lm_first = reinterpret_cast<int *>(w->blob.p + w->off_lm);
    dst.para_feature = reinterpret_cast<double *>(w->blob.p + w->off_x) + StateLayout{w->F, 0}.feat(0);
    dst.marg_lm = w->lm_list[LFVIO_MARGIN_OLD].p;
    dst.proj_cap = w->Rmax;
    dst.lm_cap = w->Lmax;
    DevFactorSrc src;
    LFVIO_TRY(fm_build_factors_into(fm, dst, &src));  // synchronises the manager's stream: the block is complete
    state->n_landmarks = src.n_landmarks;
    LFVIO_TRY(optimize_full_impl(w, state, factors, opts, margin_flag, summary, prior_out, &src));
    // double2vector(): f_manager.setDepth(getFeatureDepth) (estimator.cpp:560-563) from the re-anchored state on the device
    return fm_set_depth_from_dev(fm, w->x_out.p + w->P.lay.feat(0));
}

extern "C" int lfvio_window_upload(lfvio_window *w, const lfvio_window_state *state, const lfvio_factor_list *factors,
                                   const lfvio_solver_opts *opts) {
    clear_error();
    LFVIO_REQUIRE(w, "lfvio_window_upload: null handle");
    LFVIO_CUDA_TRY(cudaSetDevice(w->device));
    return upload(w, state, factors, opts);
}

extern "C" int lfvio_window_optimize_dev(lfvio_window *w, int with_marginalize, int margin_flag) {
    clear_error();
    LFVIO_REQUIRE(w && w->uploaded, "lfvio_window_optimize_dev: call lfvio_window_upload first");
    LFVIO_CUDA_TRY(cudaSetDevice(w->device));
    LFVIO_REQUIRE(margin_flag == LFVIO_MARGIN_OLD || margin_flag == LFVIO_MARGIN_SECOND_NEW, "lfvio_window_optimize_dev: bad margin_flag");
    LFVIO_TRY(run_graphed(w, geometry_key(w, with_marginalize ? 5 : 4, margin_flag, with_marginalize != 0), [&]() -> int {
        LFVIO_TRY(enqueue_solve(w));
        if (with_marginalize) LFVIO_TRY(enqueue_marginalize(w, margin_flag, true));
        return LFVIO_OK;
    }));
    if (with_marginalize) w->last_marg_flag = margin_flag;
    return LFVIO_OK;
}

extern "C" int lfvio_window_download(lfvio_window *w, lfvio_window_state *state, lfvio_solve_summary *summary) {
    clear_error();
    LFVIO_REQUIRE(w && w->uploaded, "lfvio_window_download: nothing uploaded");
    LFVIO_CUDA_TRY(cudaSetDevice(w->device));
    return download(w, state, summary);
}

extern "C" int lfvio_window_evaluate(lfvio_window *w, const lfvio_window_state *state, const lfvio_factor_list *factors,
                                     const lfvio_solver_opts *opts, int robust, double *res_proj, double *jac_proj,
                                     double *res_imu, double *jac_imu, double *res_prior, double *cost) {
    clear_error();
    LFVIO_REQUIRE(w, "lfvio_window_evaluate: null handle");
    LFVIO_CUDA_TRY(cudaSetDevice(w->device));
    LFVIO_TRY(upload(w, state, factors, opts));
    cudaStream_t s = w->stream;
    const DevProblem &P = w->P;
    LFVIO_TRY(launch_solver_init(w->ctl.p, P, s));
    LFVIO_CUDA_TRY(cudaMemcpyAsync(w->x[0].p, w->blob.p + w->off_x, sizeof(double) * P.lay.size(), cudaMemcpyDeviceToDevice, s));
    LFVIO_TRY(launch_evaluate(P, w->x[0].p, w->x[1].p, w->e[0], w->e[1], w->ctl.p, 2, robust, s));
    LFVIO_CUDA_TRY(cudaStreamSynchronize(s));
    if (res_proj && P.n_proj) LFVIO_CUDA_TRY(cudaMemcpy(res_proj, w->rp[0].p, sizeof(double) * 2 * P.n_proj, cudaMemcpyDeviceToHost));
    if (jac_proj && P.n_proj) LFVIO_CUDA_TRY(cudaMemcpy(jac_proj, w->Jp[0].p, sizeof(double) * 40 * P.n_proj, cudaMemcpyDeviceToHost));
    if (res_imu && P.n_imu) LFVIO_CUDA_TRY(cudaMemcpy(res_imu, w->ri[0].p, sizeof(double) * 15 * P.n_imu, cudaMemcpyDeviceToHost));
    if (jac_imu && P.n_imu) LFVIO_CUDA_TRY(cudaMemcpy(jac_imu, w->Ji[0].p, sizeof(double) * 450 * P.n_imu, cudaMemcpyDeviceToHost));
    if (res_prior && P.prior.n) LFVIO_CUDA_TRY(cudaMemcpy(res_prior, w->rm[0].p, sizeof(double) * P.prior.n, cudaMemcpyDeviceToHost));
    if (cost) {
        SolverCtl c;
        LFVIO_CUDA_TRY(cudaMemcpy(&c, w->ctl.p, sizeof(c), cudaMemcpyDeviceToHost));
        *cost = c.acc[7];
    }
    int status[4];
    LFVIO_CUDA_TRY(cudaMemcpy(status, w->status.p, sizeof(status), cudaMemcpyDeviceToHost));
    if (status[0]) {
        set_error("IMU covariance of a pre-integration is singular / not positive definite");
        return LFVIO_ERR_NUMERIC;
    }
    return LFVIO_OK;
}

extern "C" int lfvio_window_marginalize(lfvio_window *w, const lfvio_window_state *state, const lfvio_factor_list *factors,
                                        const lfvio_solver_opts *opts, int margin_flag, lfvio_prior *out) {
    clear_error();
    LFVIO_REQUIRE(w && out, "lfvio_window_marginalize: null argument");
    LFVIO_REQUIRE(margin_flag == LFVIO_MARGIN_OLD || margin_flag == LFVIO_MARGIN_SECOND_NEW, "lfvio_window_marginalize: bad margin_flag");
    LFVIO_CUDA_TRY(cudaSetDevice(w->device));
    LFVIO_TRY(upload(w, state, factors, opts));
    out->n = 0;
    out->n_blocks = 0;
    if (!w->plan_valid[margin_flag]) return LFVIO_OK;
    cudaStream_t s = w->stream;
    const DevProblem &P = w->P;
    const MargPlan &pl = w->plan[margin_flag];
    LFVIO_REQUIRE(pl.n <= out->cap_n && pl.n_keep <= out->cap_blocks, "lfvio_window_marginalize: output prior too small");
    // preMarginalize(): evaluate the dropped factors (with the loss correction) at the given state
    LFVIO_TRY(launch_solver_init(w->ctl.p, P, s));
    LFVIO_CUDA_TRY(cudaMemcpyAsync(w->x[0].p, w->blob.p + w->off_x, sizeof(double) * P.lay.size(), cudaMemcpyDeviceToDevice, s));
    LFVIO_TRY(launch_evaluate(P, w->x[0].p, w->x[1].p, w->e[0], w->e[1], w->ctl.p, 2, 1, s));
    LFVIO_TRY(enqueue_marginalize(w, margin_flag, false));
    w->last_marg_flag = margin_flag;
    const size_t nn = (size_t)pl.n * pl.n;
    int x0_len = 0;
    for (int b = 0; b < pl.n_keep; b++) x0_len += P.lay.block_gsize(pl.keep_id[b]);
    LFVIO_CUDA_TRY(cudaMemcpyAsync(w->h_marg, w->marg_out.p, sizeof(double) * (nn + pl.n + x0_len), cudaMemcpyDeviceToHost, s));
    int *status = w->h_status;  // pinned: a D2H copy into pageable memory would be staged synchronously
    LFVIO_CUDA_TRY(cudaMemcpyAsync(status, w->status.p, 4 * sizeof(int), cudaMemcpyDeviceToHost, s));
    LFVIO_CUDA_TRY(cudaStreamSynchronize(s));
    if (status[0]) {
        set_error("IMU covariance of a pre-integration is singular / not positive definite");
        return LFVIO_ERR_NUMERIC;
    }
    out->n = pl.n;
    out->n_blocks = pl.n_keep;
    const int F = w->F;
    for (int b = 0; b < pl.n_keep; b++) {
        const int id = pl.keep_id[b];
        int new_id = id;  // addr_shift (estimator.cpp:921-931 / :980-1000)
        if (id < 2 * F) {
            if (margin_flag == LFVIO_MARGIN_OLD) new_id = id - 2;
            else if ((id >> 1) == F - 1) new_id = id - 2;
        }
        out->block_id[b] = new_id;
        out->block_idx[b] = pl.col[id] - pl.m_p;
    }
    memcpy(out->J0, w->h_marg, sizeof(double) * nn);
    memcpy(out->r0, w->h_marg + nn, sizeof(double) * pl.n);
    memcpy(out->x0, w->h_marg + nn + pl.n, sizeof(double) * x0_len);
    return LFVIO_OK;
}

extern "C" void *lfvio_window_stream(lfvio_window *w) { return w ? (void *)w->stream : nullptr; }

extern "C" int lfvio_window_stage_count(void) { return WS_COUNT; }
extern "C" const char *lfvio_window_stage_name(int i) { return (i >= 0 && i < WS_COUNT) ? kWindowStageNames[i] : ""; }
extern "C" int lfvio_window_profile_begin(lfvio_window *w, int max_solves) {
    clear_error();
    LFVIO_REQUIRE(w && max_solves > 0, "lfvio_window_profile_begin: bad arguments");
    LFVIO_CUDA_TRY(cudaSetDevice(w->device));
    return w->prof.enable(WS_COUNT, max_solves * (4 + 3 * LFVIO_MAX_ITER_LOG));
}
extern "C" int lfvio_window_profile_end(lfvio_window *w, float *stage_ms, int *stage_records) {
    clear_error();
    LFVIO_REQUIRE(w && stage_ms && stage_records, "lfvio_window_profile_end: bad arguments");
    LFVIO_CUDA_TRY(cudaSetDevice(w->device));
    LFVIO_CUDA_TRY(cudaStreamSynchronize(w->stream));
    LFVIO_TRY(w->prof.collect(stage_ms, stage_records));
    w->prof.disable();
    return LFVIO_OK;
}
