// Error reporting, device probing and launch accounting for the lfvio C-ABI.
#include <atomic>
#include <cstdarg>
#include <cstring>

#include "common.cuh"

#include <mutex>
#include <utility>
#include <vector>

namespace lfvio {

static thread_local char g_err[1024] = "";
static std::atomic<uint64_t> g_launches{0};

void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
void clear_error() { g_err[0] = 0; }
int ensure_dyn_smem(const void *kernel, int bytes) {
    static std::mutex mu;
    static std::vector<std::pair<const void *, unsigned long long>> seen;  // kernel -> bit per device
    int dev = 0;
    LFVIO_CUDA_TRY(cudaGetDevice(&dev));
    std::lock_guard<std::mutex> lk(mu);
    unsigned long long *mask = nullptr;
    for (auto &p : seen)
        if (p.first == kernel) mask = &p.second;
    if (!mask) {
        seen.emplace_back(kernel, 0ull);
        mask = &seen.back().second;
    }
    if ((*mask >> (dev & 63)) & 1ull) return LFVIO_OK;
    LFVIO_CUDA_TRY(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
    *mask |= 1ull << (dev & 63);
    return LFVIO_OK;
}

static thread_local uint64_t tl_launches = 0;
void count_launch(int n) {
    g_launches.fetch_add((uint64_t)n, std::memory_order_relaxed);
    tl_launches += (uint64_t)n;
}
uint64_t thread_launch_count() { return tl_launches; }

int sm_count(int device) {
    static int cached[64] = {0};
    if (device < 0 || device >= 64) return 148;
    if (cached[device] == 0) {
        int v = 0;
        if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, device) != cudaSuccess || v <= 0) v = 148;
        cached[device] = v;
    }
    return cached[device];
}

int StageProfiler::enable(int stages, int max_records) {
    disable();
    n_stages = stages;
    cap = (size_t)max_records;
    ev = new cudaEvent_t[2 * cap];
    stage = new int[cap];
    for (size_t i = 0; i < 2 * cap; i++) LFVIO_CUDA_TRY(cudaEventCreate(&ev[i]));
    used = 0;
    on = true;
    return LFVIO_OK;
}
void StageProfiler::disable() {
    if (ev) {
        for (size_t i = 0; i < 2 * cap; i++) cudaEventDestroy(ev[i]);
        delete[] ev;
        delete[] stage;
    }
    ev = nullptr;
    stage = nullptr;
    cap = used = 0;
    on = false;
}
int StageProfiler::collect(float *ms, int *counts) {
    for (int i = 0; i < n_stages; i++) {
        ms[i] = 0.f;
        counts[i] = 0;
    }
    for (size_t r = 0; r < used; r++) {
        float t = 0.f;
        LFVIO_CUDA_TRY(cudaEventElapsedTime(&t, ev[2 * r], ev[2 * r + 1]));
        ms[stage[r]] += t;
        counts[stage[r]] += 1;
    }
    used = 0;
    return LFVIO_OK;
}

}  // namespace lfvio

extern "C" const char *lfvio_last_error(void) { return lfvio::g_err; }

extern "C" int lfvio_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();  // clear sticky "no device" state
        return 0;
    }
    return n;
}

extern "C" const char *lfvio_version(void) { return "lfvio-b200 0.1 (sm_100a)"; }

extern "C" uint64_t lfvio_launch_count(void) { return lfvio::g_launches.load(); }
