// common.cuh -- host-side plumbing shared by every translation unit of libMiniCVNative.so:
// per-thread error string, per-(thread, device) context with its own stream and
// grow-only workspaces, launch accounting.  No OpenCV, no torch types.
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
    template <class T> T* as() const { return reinterpret_cast<T*>(p); }
};

// Grow-only pinned host buffer (staging for small control arrays and results).
struct PinBuf {
    void*  p = nullptr;
    size_t cap = 0;
    int reserve(size_t bytes);
    template <class T> T* as() const { return reinterpret_cast<T*>(p); }
};

// One context per (host thread, device): the exports are re-entrant and may be called
// from any .NET thread concurrently (threading row), so nothing
// here is shared between threads except the launch counter.
struct Context {
    int          device = -1;
    int          sm_count = 148;
    cudaStream_t stream = nullptr;       // private stream for host-pointer entry points
    cudaStream_t stream2 = nullptr;      // host->device copy stream (copy/compute overlap)
    cudaStream_t stream3 = nullptr;      // device->host copy stream
    cudaEvent_t  ev[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    DevBuf       ws[72];                 // numbered workspaces, meaning fixed per call site
    PinBuf       pin[6];
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
    // optional CUDA-event timing of the dominant matcher kernel (mcvKernelTimingStart / Stop)
    bool         kt_on = false;
    int          kt_n = 0, kt_cap = 0;
    cudaEvent_t* kt_ev = nullptr;        // 2 * kt_cap events
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
// Bracket the dominant kernel of a matcher call on its launching stream (no-ops unless enabled).
inline void kt_begin(Context* c, cudaStream_t s) { if (c->kt_on && c->kt_n < c->kt_cap) cudaEventRecord(c->kt_ev[2 * c->kt_n], s); }
inline void kt_end(Context* c, cudaStream_t s) { if (c->kt_on && c->kt_n < c->kt_cap) { cudaEventRecord(c->kt_ev[2 * c->kt_n + 1], s); ++c->kt_n; } }

// Returns the calling thread's context for `device` (-1: cudaGetDevice()), creating it on
// first use; nullptr (and an error string) when no CUDA device is usable -- there is no
// CPU fallback behind this.
Context* get_context(int device);

inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

}  // namespace mcv
