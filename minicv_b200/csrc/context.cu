// context.cu -- error string, pooled per-device contexts leased per export call, diagnostics exports.
#include "common.cuh"
#include "knn.cuh"

#include <condition_variable>
#include <stdlib.h>

namespace mcv {

static thread_local char t_err[512] = {0};
std::atomic<int64_t> g_launches{0};
static std::atomic<int64_t> g_dev_bytes{0};      // device memory held by workspaces (all contexts)
static std::atomic<int64_t> g_pin_bytes{0};      // pinned host memory held by staging buffers

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(t_err, sizeof(t_err), fmt, ap);
    va_end(ap);
}
void clear_error() { t_err[0] = 0; }
const char* get_error() { return t_err; }

int DevBuf::reserve(size_t bytes) {
    if (bytes <= cap) return 0;
    release();
    size_t want = align_up(bytes + bytes / 4, 1 << 20);
    cudaError_t e = cudaMalloc(&p, want);
    if (e != cudaSuccess) {
        p = nullptr;
        cudaGetLastError();
        set_error("cudaMalloc(%zu) failed: %s", want, cudaGetErrorString(e));
        return -1;
    }
    cap = want;
    g_dev_bytes.fetch_add((int64_t)cap);
    return 0;
}
void DevBuf::release() {
    if (p) { cudaFree(p); g_dev_bytes.fetch_sub((int64_t)cap); }
    p = nullptr; cap = 0;
}

int PinBuf::reserve(size_t bytes) {
    if (bytes <= cap) return 0;
    release();
    size_t want = align_up(bytes + bytes / 4, 1 << 16);
    cudaError_t e = cudaMallocHost(&p, want);
    if (e != cudaSuccess) {
        p = nullptr;
        cudaGetLastError();
        set_error("cudaMallocHost(%zu) failed: %s", want, cudaGetErrorString(e));
        return -1;
    }
    cap = want;
    g_pin_bytes.fetch_add((int64_t)cap);
    return 0;
}
void PinBuf::release() {
    if (p) { cudaFreeHost(p); g_pin_bytes.fetch_sub((int64_t)cap); }
    p = nullptr; cap = 0;
}

// ---- per-device pool ---------------------------------------------------------------------------------------
struct DeviceState {
    std::mutex              mu;
    std::condition_variable cv;
    std::vector<Context*>   all;         // owned; never destroyed (CUDA teardown order at process exit is undefined)
    std::vector<Context*>   idle;        // LIFO: the most recently used context has the warm workspaces
    int                     sm_count = 0;
    bool                    probed = false, usable = false;
    // optional CUDA-event timing of the dominant matcher kernel (mcvKernelTimingStart / Stop)
    std::mutex              ktMu;
    std::atomic<bool>       ktOn{false};
    int                     ktN = 0, ktCap = 0;
    cudaEvent_t*            ktEv = nullptr;   // 2 * ktCap events
    int*                    ktTag = nullptr;  // ktCap stage tags (common.cuh KT_*)
};

static constexpr int kMaxDevices = 64;
// never destroyed: threads of the host application may still be inside an export while static destructors run
static DeviceState* const g_dev = new DeviceState[kMaxDevices];
static std::atomic<int64_t> g_ws_limit{-1};      // bytes per context kept across calls; < 0: unlimited

static int max_contexts() {
    static const int n = [] {
        const char* e = getenv("MINICV_MAX_CONTEXTS");
        int v = e ? atoi(e) : 8;
        return v < 1 ? 1 : v > 256 ? 256 : v;
    }();
    return n;
}

void kt_begin(Context* c, cudaStream_t s, int tag) {
    DeviceState* d = c->dev;
    if (!d->ktOn.load(std::memory_order_relaxed)) return;
    std::lock_guard<std::mutex> lk(d->ktMu);
    if (d->ktOn && d->ktN < d->ktCap) { cudaEventRecord(d->ktEv[2 * d->ktN], s); d->ktTag[d->ktN] = tag; }
}
void kt_end(Context* c, cudaStream_t s) {
    DeviceState* d = c->dev;
    if (!d->ktOn.load(std::memory_order_relaxed)) return;
    std::lock_guard<std::mutex> lk(d->ktMu);
    if (d->ktOn && d->ktN < d->ktCap) { cudaEventRecord(d->ktEv[2 * d->ktN + 1], s); ++d->ktN; }
}

static size_t context_bytes(const Context* c) {
    size_t t = 0;
    for (const DevBuf& b : c->ws) t += b.cap;
    for (const DevBuf& b : c->ctlWs) t += b.cap;
    return t;
}
void stage_ring_release(Context* c);     // staging.cu
static void context_free_buffers(Context* c) {
    for (DevBuf& b : c->ws) b.release();
    for (DevBuf& b : c->ctlWs) b.release();
    for (PinBuf& b : c->pin) b.release();
    for (PinBuf& b : c->ctlPin) b.release();
    stage_ring_release(c);
    c->mwc_ready = false;
}

static Context* create_context(int device, DeviceState* d) {
    if (cudaSetDevice(device) != cudaSuccess) {
        cudaGetLastError();
        set_error("cudaSetDevice(%d) failed", device);
        return nullptr;
    }
    if (!d->probed) {
        cudaDeviceProp prop;
        if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) {
            cudaGetLastError();
            set_error("cudaGetDeviceProperties(%d) failed", device);
            return nullptr;
        }
        d->sm_count = prop.multiProcessorCount;
        d->usable = prop.major >= 10;
        d->probed = true;
        if (!d->usable) {
            set_error("device %d is sm_%d%d; this library carries sm_100a code only", device, prop.major, prop.minor);
            return nullptr;
        }
    }
    if (!d->usable) { set_error("device %d cannot run sm_100a code", device); return nullptr; }
    Context* c = new Context();
    c->device = device;
    c->dev = d;
    c->sm_count = d->sm_count;
    bool ok = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) == cudaSuccess &&
              cudaStreamCreateWithFlags(&c->stream2, cudaStreamNonBlocking) == cudaSuccess &&
              cudaStreamCreateWithFlags(&c->stream3, cudaStreamNonBlocking) == cudaSuccess;
    for (int i = 0; ok && i < 8; ++i) ok = cudaEventCreateWithFlags(&c->ev[i], cudaEventDisableTiming) == cudaSuccess;
    for (int i = 0; ok && i < 2; ++i) ok = cudaEventCreateWithFlags(&c->ctlEv[i], cudaEventDisableTiming) == cudaSuccess;
    ok = ok && cudaEventCreateWithFlags(&c->lastUse, cudaEventDisableTiming) == cudaSuccess;
    if (!ok) {
        set_error("stream / event creation failed on device %d: %s", device, cudaGetErrorString(cudaGetLastError()));
        delete c;      // the few handles created so far are leaked; this only happens when the device is gone
        return nullptr;
    }
    return c;
}

// ---- leases of the calling thread ----------------------------------------------------------------------------
struct Lease { Context* ctx; cudaStream_t asyncStream; bool async; };
struct ThreadScope {
    int depth = 0;
    int callerDevice = -1;
    bool deviceChanged = false;
    std::vector<Lease> leases;
};
static thread_local ThreadScope t_scope;

static void release_lease(const Lease& l) {
    Context* c = l.ctx;
    cudaSetDevice(c->device);
    if (l.async) {
        // work may still be queued on the caller's stream: the next lease waits for it on the device
        c->lastUsePending = cudaEventRecord(c->lastUse, l.asyncStream) == cudaSuccess;
        if (!c->lastUsePending) { cudaGetLastError(); cudaStreamSynchronize(l.asyncStream); }
    }
    // host-pointer exports synchronise before returning; an error path may have left copies or kernels in flight on the
    // context's own streams -- drain them so the caller's buffers and our scratch are quiescent
    if (cudaStreamQuery(c->stream) != cudaSuccess) { cudaGetLastError(); cudaStreamSynchronize(c->stream); }
    if (cudaStreamQuery(c->stream2) != cudaSuccess) { cudaGetLastError(); cudaStreamSynchronize(c->stream2); }
    if (cudaStreamQuery(c->stream3) != cudaSuccess) { cudaGetLastError(); cudaStreamSynchronize(c->stream3); }
    cudaGetLastError();
    const int64_t lim = g_ws_limit.load(std::memory_order_relaxed);
    if (lim >= 0 && (int64_t)context_bytes(c) > lim) {
        if (c->lastUsePending) { cudaEventSynchronize(c->lastUse); c->lastUsePending = false; }
        context_free_buffers(c);
    }
    DeviceState* d = c->dev;
    {
        std::lock_guard<std::mutex> lk(d->mu);
        d->idle.push_back(c);
    }
    d->cv.notify_one();
}

ApiScope::ApiScope() {
    ThreadScope& t = t_scope;
    if (t.depth++ == 0) {
        t.deviceChanged = false;
        t.callerDevice = -1;
    }
}
ApiScope::~ApiScope() {
    ThreadScope& t = t_scope;
    if (--t.depth > 0) return;
    for (const Lease& l : t.leases) release_lease(l);
    t.leases.clear();
    if (t.deviceChanged && t.callerDevice >= 0) cudaSetDevice(t.callerDevice);
    cudaGetLastError();
}

static Context* lease(int device, bool async, cudaStream_t stream) {
    ThreadScope& t = t_scope;
    if (t.depth <= 0) { set_error("internal: context requested outside an export scope"); return nullptr; }
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n <= 0) {
        cudaGetLastError();
        set_error("no usable CUDA device (%s); libMiniCVNative has no CPU fallback",
                  e != cudaSuccess ? cudaGetErrorString(e) : "device count 0");
        return nullptr;
    }
    if (t.callerDevice < 0) {
        if (cudaGetDevice(&t.callerDevice) != cudaSuccess) { cudaGetLastError(); t.callerDevice = 0; }
    }
    if (device < 0) device = t.callerDevice;
    if (device >= n || device >= kMaxDevices) {
        set_error("device %d out of range (have %d)", device, n);
        return nullptr;
    }
    Context* c = nullptr;
    for (Lease& l : t.leases)
        if (l.ctx->device == device) {
            c = l.ctx;
            if (async && !l.async) { l.async = true; l.asyncStream = stream; }
            break;
        }
    const bool fresh = (c == nullptr);
    if (fresh) {
        DeviceState* d = &g_dev[device];
        std::unique_lock<std::mutex> lk(d->mu);
        for (;;) {
            if (!d->idle.empty()) { c = d->idle.back(); d->idle.pop_back(); break; }
            if ((int)d->all.size() < max_contexts()) {
                c = create_context(device, d);
                if (!c) { cudaSetDevice(t.callerDevice); return nullptr; }
                d->all.push_back(c);
                break;
            }
            d->cv.wait(lk);
        }
        lk.unlock();
        t.leases.push_back(Lease{c, stream, async});
    }
    if (device != t.callerDevice) t.deviceChanged = true;
    if (cudaSetDevice(device) != cudaSuccess) { cudaGetLastError(); set_error("cudaSetDevice(%d) failed", device); return nullptr; }
    if (fresh && c->lastUsePending) {
        // the previous lease left asynchronous work behind: order every stream this call can use after it
        cudaStreamWaitEvent(c->stream, c->lastUse, 0);
        cudaStreamWaitEvent(c->stream2, c->lastUse, 0);
        cudaStreamWaitEvent(c->stream3, c->lastUse, 0);
        if (async) cudaStreamWaitEvent(stream, c->lastUse, 0);
        c->lastUsePending = false;
    } else if (!fresh && async) {
        // nested async use inside one scope (an export built on a *Device export): the streams are the scope's own
    }
    return c;
}

Context* get_context(int device) { return lease(device, false, nullptr); }
Context* get_context_async(int device, cudaStream_t stream) { return lease(device, true, stream); }

}  // namespace mcv

extern "C" {

MCV_API const char* mcvLastError(void) { return mcv::t_err; }

MCV_API int mcvDeviceCount(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

MCV_API int64_t mcvKernelLaunchCount(void) { return mcv::g_launches.load(); }

MCV_API int mcvKernelTimingStart(int device, int capacity) {
    MCV_API_ENTER();
    mcv::Context* c = mcv::get_context(device);
    if (!c || capacity <= 0) return -1;
    mcv::DeviceState* d = c->dev;
    std::lock_guard<std::mutex> lk(d->ktMu);
    if (capacity > d->ktCap) {
        cudaEvent_t* ev = new cudaEvent_t[2 * (size_t)capacity];
        for (int i = 0; i < 2 * capacity; ++i) cudaEventCreate(&ev[i]);
        for (int i = 0; i < 2 * d->ktCap; ++i) cudaEventDestroy(d->ktEv[i]);
        delete[] d->ktEv;
        delete[] d->ktTag;
        d->ktEv = ev;
        d->ktTag = new int[(size_t)capacity];
        d->ktCap = capacity;
    }
    d->ktN = 0;
    d->ktOn = true;
    return 0;
}

MCV_API int mcvKernelTimingStopTagged(int device, float* ms, int* tags, int capacity) {
    MCV_API_ENTER();
    mcv::Context* c = mcv::get_context(device);
    if (!c || !ms) return -1;
    mcv::DeviceState* d = c->dev;
    std::lock_guard<std::mutex> lk(d->ktMu);
    d->ktOn = false;
    int n = 0;
    for (int i = 0; i < d->ktN && n < capacity; ++i) {
        if (!tags && d->ktTag[i] != 0) continue;        // untagged form: the dominant matcher kernel only
        if (cudaEventSynchronize(d->ktEv[2 * i + 1]) != cudaSuccess ||
            cudaEventElapsedTime(&ms[n], d->ktEv[2 * i], d->ktEv[2 * i + 1]) != cudaSuccess) {
            cudaGetLastError();
            break;
        }
        if (tags) tags[n] = d->ktTag[i];
        ++n;
    }
    d->ktN = 0;
    return n;
}

MCV_API int mcvKernelTimingStop(int device, float* ms, int capacity) {
    return mcvKernelTimingStopTagged(device, ms, nullptr, capacity);
}

MCV_API int mcvSetHammingEngine(int engine) {
    if (engine < 0 || engine > 3) return -1;
    mcv::g_hamming_engine.store(engine);
    return 0;
}

MCV_API int mcvSetL2Engine(int engine) {
    if (engine < 0 || engine > 2) return -1;
    mcv::g_l2_engine.store(engine);
    return 0;
}

MCV_API int mcvTrim(void) {
    // frees the workspaces and staging buffers of every idle context (leased ones are left alone)
    int freed = 0;
    int prev = -1;
    if (cudaGetDevice(&prev) != cudaSuccess) { cudaGetLastError(); prev = -1; }
    for (int dv = 0; dv < mcv::kMaxDevices; ++dv) {
        mcv::DeviceState* d = &mcv::g_dev[dv];
        std::vector<mcv::Context*> mine;
        {
            std::lock_guard<std::mutex> lk(d->mu);
            mine.swap(d->idle);
        }
        if (mine.empty()) continue;
        cudaSetDevice(dv);
        for (mcv::Context* c : mine) {
            if (c->lastUsePending) { cudaEventSynchronize(c->lastUse); c->lastUsePending = false; }
            mcv::context_free_buffers(c);
            ++freed;
        }
        {
            std::lock_guard<std::mutex> lk(d->mu);
            for (mcv::Context* c : mine) d->idle.push_back(c);
        }
        d->cv.notify_all();
    }
    if (prev >= 0) cudaSetDevice(prev);
    cudaGetLastError();
    return freed;
}

MCV_API int64_t mcvWorkspaceBytes(int kind) {
    return kind == 1 ? mcv::g_pin_bytes.load() : mcv::g_dev_bytes.load();
}

MCV_API int mcvSetWorkspaceLimit(int64_t bytesPerContext) {
    mcv::g_ws_limit.store(bytesPerContext);
    return 0;
}

MCV_API int mcvContextCount(int device) {
    if (device < 0 || device >= mcv::kMaxDevices) return 0;
    std::lock_guard<std::mutex> lk(mcv::g_dev[device].mu);
    return (int)mcv::g_dev[device].all.size();
}

MCV_API const char* mcvBuildInfo(void) { return "minicv_b200 sm_100a " __DATE__ " " __TIME__; }

}
