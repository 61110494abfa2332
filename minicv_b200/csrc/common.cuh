// common.cuh -- host-side plumbing shared by every translation unit of libMiniCVNative.so:
// per-thread error string, pooled per-device contexts (stream set + grow-only workspaces) leased per export
// call, launch accounting.  No OpenCV, no torch types.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include <string.h>
#include <atomic>
#include <vector>
#include <mutex>

#include "../../include/MiniCVNative.h"

namespace mcv {

// ---- error reporting (reference convention: failure value + untouched outputs) ----
void set_error(const char* fmt, ...);
void clear_error();
const char* get_error();          // the calling thread's error string ("" when none)

#define MCV_CUDA_TRY(expr)                                                          \
    do {                                                                            \
        cudaError_t _e = (expr);                                                    \
        if (_e != cudaSuccess) {                                                    \
            ::mcv::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e),\
                             __FILE__, __LINE__);                                   \
            return -1;                                                              \
        }                                                                           \
    } while (0)

extern std::atomic<int64_t> g_launches;
inline void count_launch(int n = 1) { g_launches.fetch_add(n, std::memory_order_relaxed); }

#define MCV_LAUNCH_CHECK()                                                          \
    do {                                                                            \
        ::mcv::count_launch();                                                      \
        cudaError_t _e = cudaGetLastError();                                        \
        if (_e != cudaSuccess) {                                                    \
            ::mcv::set_error("kernel launch failed: %s (%s:%d)",                    \
                             cudaGetErrorString(_e), __FILE__, __LINE__);           \
            return -1;                                                              \
        }                                                                           \
    } while (0)

// Grow-only device buffer.
struct DevBuf {
    void*  p = nullptr;
    size_t cap = 0;
    int reserve(size_t bytes);
    void release();
    template <class T> T* as() const { return reinterpret_cast<T*>(p); }
};

// Grow-only pinned host buffer (staging for small control arrays and results).
struct PinBuf {
    void*  p = nullptr;
    size_t cap = 0;
    int reserve(size_t bytes);
    void release();
    template <class T> T* as() const { return reinterpret_cast<T*>(p); }
};

// Contexts (stream set + grow-only workspaces + pinned staging) live in a process-wide pool per device.  Every
// compute export opens an ApiScope; get_context() leases a context to the calling thread until the outermost scope
// of that thread closes, so the exports are re-entrant from any number of host threads (SURVEY.md section 8b,
// threading row) while device memory scales with the number of CONCURRENT calls (capped, MINICV_MAX_CONTEXTS per
// device, default 8; further callers wait), not with the number of threads that ever called.  A thread that
// retires leaves nothing behind.  mcvTrim() frees the workspaces of every idle context.
//
// Asynchronous (*Device) exports leave work queued on the caller's stream when they return: the scope records a
// "last use" event on that stream, and whoever leases the context next makes its streams wait on it, so scratch
// buffers are never overwritten under a running kernel whichever stream / thread comes next.
struct DeviceState;
struct Context {
    int          device = -1;
    int          sm_count = 148;
    DeviceState* dev = nullptr;
    cudaStream_t stream = nullptr;       // private stream for host-pointer entry points
    cudaStream_t stream2 = nullptr;      // host->device copy stream (copy/compute overlap)
    cudaStream_t stream3 = nullptr;      // device->host copy stream
    cudaEvent_t  ev[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    DevBuf       ws[96];                 // numbered workspaces, meaning fixed per call site
    PinBuf       pin[8];
    // tensor-core matcher control arrays, double-buffered so that the host can prepare launch k+1
    // while the copy of launch k is still queued behind launch k-1
    PinBuf       ctlPin[2];
    DevBuf       ctlWs[2];
    cudaEvent_t  ctlEv[2] = {nullptr, nullptr};
    int          ctlTurn = 0;
    // growable pool of (timing-disabled) events for the chunked host-buffer entry points
    cudaEvent_t* evPool = nullptr;
    int          evPoolSize = 0;
    bool         mwc_ready = false;      // ws[28] holds OpenCV's raw RNG stream (ransac.cu)
    // pending asynchronous use by the previous lease (see above)
    cudaEvent_t  lastUse = nullptr;
    bool         lastUsePending = false;
    // pinned staging ring for pageable caller buffers (staging.cu)
    struct StageRing* ring = nullptr;
};
// n events from the pool (grown on demand); nullptr on failure
inline cudaEvent_t* event_pool(Context* c, int n) {
    if (n > c->evPoolSize) {
        cudaEvent_t* ev = new cudaEvent_t[n];
        for (int i = 0; i < c->evPoolSize; ++i) ev[i] = c->evPool[i];
        for (int i = c->evPoolSize; i < n; ++i)
            if (cudaEventCreateWithFlags(&ev[i], cudaEventDisableTiming) != cudaSuccess) { delete[] ev; return nullptr; }
        delete[] c->evPool;
        c->evPool = ev;
        c->evPoolSize = n;
    }
    return c->evPool;
}
// Bracket the dominant kernel of a matcher call on its launching stream (no-ops unless mcvKernelTimingStart
// switched the device's timing on).
// Stage tags (mcvKernelTimingStopTagged): 0 = the dominant matcher kernel (tcgen05 GEMM + top-2, or XOR/POPC).
enum KtTag {
    KT_MATCH = 0, KT_EXPAND = 1, KT_KNN_FINALIZE = 2, KT_L2 = 3,
    KT_PREP = 10, KT_SAMPLE = 11, KT_SOLVE = 12, KT_COMPACT = 13, KT_SCORE = 14, KT_SELECT = 15, KT_FINALIZE = 16,
    KT_POSE = 17, KT_GATHER = 18, KT_SCALE = 19
};
void kt_begin(Context* c, cudaStream_t s, int tag = KT_MATCH);
void kt_end(Context* c, cudaStream_t s);

// RAII guard of one export call.  The outermost scope of a thread remembers the caller's current device; when it
// closes it returns every context the thread leased inside it and restores that device.
struct ApiScope {
    ApiScope();
    ~ApiScope();
    ApiScope(const ApiScope&) = delete;
    ApiScope& operator=(const ApiScope&) = delete;
};
#define MCV_API_ENTER() ::mcv::ApiScope _mcv_scope; ::mcv::clear_error()

// The calling thread's context for `device` (-1: the caller's current device), leased for the current ApiScope;
// nullptr (and an error string) when no CUDA device is usable -- there is no CPU fallback behind this.
// get_context: for exports that synchronise before they return (work goes to the context's own streams).
// get_context_async: for *Device exports that enqueue on `stream` and return.
Context* get_context(int device);
Context* get_context_async(int device, cudaStream_t stream);

inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

}  // namespace mcv
