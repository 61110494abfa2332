// context.cu -- error string, per-thread contexts, diagnostics exports.
#include "common.cuh"
#include "knn.cuh"

#include <map>
#include <memory>

namespace mcv {

static thread_local char t_err[512] = {0};
std::atomic<int64_t> g_launches{0};

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(t_err, sizeof(t_err), fmt, ap);
    va_end(ap);
}
void clear_error() { t_err[0] = 0; }

int DevBuf::reserve(size_t bytes) {
    if (bytes <= cap) return 0;
    if (p) { cudaFree(p); p = nullptr; cap = 0; }
    size_t want = align_up(bytes + bytes / 4, 1 << 20);
    cudaError_t e = cudaMalloc(&p, want);
    if (e != cudaSuccess) {
        p = nullptr;
        set_error("cudaMalloc(%zu) failed: %s", want, cudaGetErrorString(e));
        return -1;
    }
    cap = want;
    return 0;
}

int PinBuf::reserve(size_t bytes) {
    if (bytes <= cap) return 0;
    if (p) { cudaFreeHost(p); p = nullptr; cap = 0; }
    size_t want = align_up(bytes + bytes / 4, 1 << 16);
    cudaError_t e = cudaMallocHost(&p, want);
    if (e != cudaSuccess) {
        p = nullptr;
        set_error("cudaMallocHost(%zu) failed: %s", want, cudaGetErrorString(e));
        return -1;
    }
    cap = want;
    return 0;
}

struct ThreadContexts {
    std::map<int, std::unique_ptr<Context>> by_device;
    // Contexts are intentionally leaked at thread exit: CUDA teardown order at process
    // exit is undefined and freeing after the driver unloaded would crash the host app.
    ~ThreadContexts() { for (auto& kv : by_device) kv.second.release(); }
};
static thread_local ThreadContexts t_ctx;

Context* get_context(int device) {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n <= 0) {
        cudaGetLastError();
        set_error("no usable CUDA device (%s); libMiniCVNative has no CPU fallback",
                  e != cudaSuccess ? cudaGetErrorString(e) : "device count 0");
        return nullptr;
    }
    if (device < 0) {
        if (cudaGetDevice(&device) != cudaSuccess) device = 0;
    }
    if (device >= n) {
        set_error("device %d out of range (have %d)", device, n);
        return nullptr;
    }
    auto it = t_ctx.by_device.find(device);
    if (it != t_ctx.by_device.end()) {
        cudaSetDevice(device);
        return it->second.get();
    }
    if (cudaSetDevice(device) != cudaSuccess) {
        set_error("cudaSetDevice(%d) failed", device);
        return nullptr;
    }
    std::unique_ptr<Context> c(new Context());
    c->device = device;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) == cudaSuccess) {
        c->sm_count = prop.multiProcessorCount;
        if (prop.major < 10) {
            set_error("device %d is sm_%d%d; this library carries sm_100a code only",
                      device, prop.major, prop.minor);
            return nullptr;
        }
    }
    if (cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaStreamCreateWithFlags(&c->stream2, cudaStreamNonBlocking) != cudaSuccess ||
        cudaStreamCreateWithFlags(&c->stream3, cudaStreamNonBlocking) != cudaSuccess) {
        set_error("cudaStreamCreate failed: %s", cudaGetErrorString(cudaGetLastError()));
        return nullptr;
    }
    for (int i = 0; i < 8; ++i) cudaEventCreateWithFlags(&c->ev[i], cudaEventDisableTiming);
    for (int i = 0; i < 2; ++i) cudaEventCreateWithFlags(&c->ctlEv[i], cudaEventDisableTiming);
    Context* raw = c.get();
    t_ctx.by_device[device] = std::move(c);
    return raw;
}

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
    mcv::clear_error();
    mcv::Context* c = mcv::get_context(device);
    if (!c || capacity <= 0) return -1;
    if (capacity > c->kt_cap) {
        cudaEvent_t* ev = new cudaEvent_t[2 * (size_t)capacity];
        for (int i = 0; i < 2 * capacity; ++i) cudaEventCreate(&ev[i]);
        for (int i = 0; i < 2 * c->kt_cap; ++i) cudaEventDestroy(c->kt_ev[i]);
        delete[] c->kt_ev;
        c->kt_ev = ev;
        c->kt_cap = capacity;
    }
    c->kt_n = 0;
    c->kt_on = true;
    return 0;
}

MCV_API int mcvKernelTimingStop(int device, float* ms, int capacity) {
    mcv::clear_error();
    mcv::Context* c = mcv::get_context(device);
    if (!c) return -1;
    c->kt_on = false;
    int n = 0;
    for (; n < c->kt_n && n < capacity; ++n) {
        if (cudaEventSynchronize(c->kt_ev[2 * n + 1]) != cudaSuccess ||
            cudaEventElapsedTime(&ms[n], c->kt_ev[2 * n], c->kt_ev[2 * n + 1]) != cudaSuccess) {
            cudaGetLastError();
            break;
        }
    }
    c->kt_n = 0;
    return n;
}

MCV_API int mcvSetHammingEngine(int engine) {
    if (engine < 0 || engine > 3) return -1;
    mcv::g_hamming_engine.store(engine);
    return 0;
}

MCV_API const char* mcvBuildInfo(void) { return "minicv_b200 sm_100a " __DATE__ " " __TIME__; }

}
