// Shared host/device helpers for the lfvio CUDA library (sm_100a only).
#pragma once
#include <cstdlib>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string>
#include <utility>

#include "../../include/lfvio.h"

namespace lfvio {

// thread-local last error text (lfvio_last_error)
void set_error(const char *fmt, ...);
void clear_error();
// process-wide kernel launch counter (lfvio_launch_count)
void count_launch(int n = 1);
// launches counted on the calling thread (delta around a stream capture = kernels inside the captured graph)
uint64_t thread_launch_count();

#define LFVIO_CUDA_TRY(expr)                                                              \
    do {                                                                                  \
        cudaError_t _e = (expr);                                                          \
        if (_e != cudaSuccess) {                                                          \
            ::lfvio::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr,              \
                               cudaGetErrorString(_e));                                   \
            return LFVIO_ERR_CUDA;                                                        \
        }                                                                                 \
    } while (0)

#define LFVIO_TRY(expr)                 \
    do {                                \
        int _s = (expr);                \
        if (_s != LFVIO_OK) return _s;  \
    } while (0)

#define LFVIO_REQUIRE(cond, msg)                                        \
    do {                                                                \
        if (!(cond)) {                                                  \
            ::lfvio::set_error("%s:%d: %s", __FILE__, __LINE__, msg);   \
            return LFVIO_ERR_ARG;                                       \
        }                                                               \
    } while (0)

// Check the launch itself (configuration errors); execution errors surface at the next sync.
#define LFVIO_LAUNCH_CHECK()                      \
    do {                                          \
        ::lfvio::count_launch();                  \
        LFVIO_CUDA_TRY(cudaPeekAtLastError());    \
    } while (0)

static inline int cdiv(int a, int b) { return (a + b - 1) / b; }

#ifdef __CUDACC__
// Programmatic dependent launch (sm_90+): a kernel launched with launch_pdl() may be scheduled while its predecessor in
// the stream is still running; pdl_prologue() at the top of BOTH kernels (a) lets the successor's CTAs become resident
// right away and (b) blocks until the predecessor has completed and its writes are visible.  For chains of short,
// strictly dependent kernels (the trust-region loop) this removes the launch/scheduling gap between them.
// Optional per-CTA timeline (make EXTRA=-DLFVIO_CTA_PROFILE): globaltimer stamps of up to 4 points per CTA of the last launch.
#ifdef LFVIO_CTA_PROFILE
#define CTA_PROF_DECL(name) static __device__ long long name[4096][4]
#define CTA_PROF(name, i)                                                                              \
    do {                                                                                               \
        if (threadIdx.x == 0 && blockIdx.x < 4096) {                                                   \
            unsigned long long t_;                                                                     \
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_));                                      \
            name[blockIdx.x][i] = (long long)t_;                                                       \
        }                                                                                              \
    } while (0)
#define CTA_PROF_READER(fn, name)                                                                      \
    extern "C" int fn(long long *out, int n_cta) {                                                     \
        return cudaMemcpyFromSymbol(out, name, sizeof(long long) * 4 * (size_t)n_cta) == cudaSuccess ? 0 : -1; \
    }
#else
#define CTA_PROF_DECL(name)
#define CTA_PROF(name, i)
#define CTA_PROF_READER(fn, name)
#endif
__device__ __forceinline__ void pdl_prologue() {
    asm volatile("griddepcontrol.launch_dependents;");
    asm volatile("griddepcontrol.wait;" ::: "memory");
}
template <class... KArgs, class... Args>
static inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, Args &&...args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = s;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    // Off by default (LFVIO_PDL=1 turns it on): measured on a B200, a dependent grid that was launched early keeps the OTHER
    // stream's kernels from being scheduled (tracker stream 0.27 -> 0.76 ms/frame next to the solver chain) and, with the
    // chain replayed from a CUDA graph, gains the chain itself only 1.6 % (profiles/r2_streams.md).
    static const bool pdl = [] { const char *e = getenv("LFVIO_PDL"); return e && e[0] == '1'; }();
    cfg.numAttrs = (s == nullptr || !pdl) ? 0 : 1;  // stage-level entry points run on the default stream: plain launch there
    return cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...);
}
#endif

// cudaFuncAttributeMaxDynamicSharedMemorySize is a per-device (per-context) property of a kernel: set it once per
// (kernel, device).  Handles on several devices may live in one process (one host thread per stream).
int ensure_dyn_smem(const void *kernel, int bytes);

// B200: 148 SMs.  Queried once per device and cached.
int sm_count(int device);

// Optional per-stage device timing (CUDA events on the handle's own stream), used by bench.py to measure the
// average launch duration of each kernel group live, without a profiler attached.
struct StageProfiler {
    bool on = false;
    int n_stages = 0;
    size_t used = 0;
    cudaEvent_t *ev = nullptr;  // 2 * cap events
    int *stage = nullptr;
    size_t cap = 0;
    int enable(int stages, int max_records);
    void disable();
    void begin(int st, cudaStream_t s) {
        if (on && used < cap) {
            stage[used] = st;
            cudaEventRecord(ev[2 * used], s);
        }
    }
    void end(cudaStream_t s) {
        if (on && used < cap) {
            cudaEventRecord(ev[2 * used + 1], s);
            used++;
        }
    }
    // after a stream synchronise: per-stage total milliseconds and record counts; resets the record list
    int collect(float *ms, int *counts);
    ~StageProfiler() { disable(); }
};

}  // namespace lfvio

#ifdef __CUDACC__
namespace lfvio {

// OpenCV cvRound(float/double) == lrint == round-half-to-even.
__device__ __forceinline__ int cv_round(float v) { return __float2int_rn(v); }
__device__ __forceinline__ int cv_round(double v) { return __double2int_rn(v); }
__device__ __forceinline__ int cv_floor(float v) { return __float2int_rd(v); }

// cv::borderInterpolate(p, len, BORDER_REFLECT_101) for |overshoot| < len-1 repeated.
__device__ __forceinline__ int reflect101(int p, int len) {
    if (len == 1) return 0;
    while (p < 0 || p >= len) {
        if (p < 0) p = -p;
        else p = 2 * (len - 1) - p;
    }
    return p;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ long long warp_sum(long long v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ int warp_sum(int v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

}  // namespace lfvio
#endif
