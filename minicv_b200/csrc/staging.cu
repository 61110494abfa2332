// staging.cu -- host-buffer plumbing of the C ABI.
//
// A P/Invoke caller hands the exports GC-pinned *pageable* arrays (the F# wrappers pin V2d[] / byte[] with `fixed`,
// reference src/MiniCV/OpenCV.fs:855-920).  cudaMemcpyAsync from pageable memory is staged by the driver through one
// small bounce buffer on the calling thread, far below the PCIe rate.  The exports therefore look at every large host
// pointer (cudaPointerGetAttributes): CUDA-pinned / registered memory is copied directly; pageable memory goes through
// a library-owned ring of pinned slots, filled by a small pool of copy threads and drained by cudaMemcpyAsync, so the
// host memcpy of piece i+1.. overlaps the DMA of piece i.  Results travel the other way through a pinned bounce
// buffer and are handed to the caller's pageable arrays by the same threads.
//
// Also here: pinned allocation / registration exports for callers that want the direct path, and the copy-bandwidth
// probe bench.py uses to measure the box's host<->device ceiling.
#include "common.cuh"
#include "staging.cuh"

#include <condition_variable>
#include <functional>
#include <thread>
#include <vector>
#include <stdlib.h>

namespace mcv {

// ---- copy-thread pool ---------------------------------------------------------------------------------------
namespace {

struct Job {
    std::function<void(int)> fn;     // fn(participant index)
    int want = 0;                    // helpers wanted
    int taken = 0;                   // helpers that picked the job up
    int done = 0;
};

struct CopyPool {
    std::mutex mu;
    std::condition_variable cvWork, cvDone;
    std::vector<Job*> queue;
    int nThreads = 0;
    bool started = false;

    void start() {
        if (started) return;
        started = true;
        const char* e = getenv("MINICV_COPY_THREADS");
        int n = e ? atoi(e) : 0;
        if (n <= 0) {
            const unsigned hc = std::thread::hardware_concurrency();
            // measured on the 16-vCPU single-GPU box (C2 pageable, pairs/s): 4 threads 51.5 k, 6: 56.9 k, 8: 66.5 k, 12: 52.3 k
            n = hc >= 16 ? 8 : hc >= 8 ? 3 : hc >= 4 ? 2 : 1;
        }
        n = std::min(n, 32);
        nThreads = n;
        for (int i = 0; i < n - 1; ++i) {       // the calling thread is participant 0
            std::thread t([this] { loop(); });
            t.detach();                          // never joined: CUDA teardown order at process exit is undefined
        }
    }
    void loop() {
        std::unique_lock<std::mutex> lk(mu);
        for (;;) {
            Job* j = nullptr;
            for (Job* q : queue) if (q->taken < q->want) { j = q; break; }
            if (!j) { cvWork.wait(lk); continue; }
            const int me = ++j->taken;           // participants 1..want
            lk.unlock();
            j->fn(me);
            lk.lock();
            if (++j->done == j->want) cvDone.notify_all();
        }
    }
    // runs fn(0) on the caller and fn(1..helpers) on pool threads; returns when all are done
    void run(int helpers, const std::function<void(int)>& fn) {
        {
            std::lock_guard<std::mutex> lk(mu);
            start();
        }
        helpers = std::max(0, std::min(helpers, nThreads - 1));
        if (helpers == 0) { fn(0); return; }
        Job j;
        j.fn = fn; j.want = helpers;
        {
            std::lock_guard<std::mutex> lk(mu);
            queue.push_back(&j);
        }
        cvWork.notify_all();
        fn(0);
        std::unique_lock<std::mutex> lk(mu);
        // helpers that have not started yet are not needed any more once the caller ran out of pieces
        j.want = j.taken;
        cvDone.wait(lk, [&] { return j.done >= j.want; });
        for (size_t i = 0; i < queue.size(); ++i) if (queue[i] == &j) { queue.erase(queue.begin() + i); break; }
    }
};
// Leaked on purpose: the pool's threads are detached and sleep on its condition variable for the life of the
// process; destroying a condition variable that still has waiters blocks inside glibc (process exit would hang).
CopyPool& g_pool = *new CopyPool();

}  // namespace

int copy_threads() {
    std::lock_guard<std::mutex> lk(g_pool.mu);
    g_pool.start();
    return g_pool.nThreads;
}

void parallel_memcpy(void* dst, const void* src, size_t bytes) {
    const size_t piece = 1u << 20;
    if (bytes <= 2 * piece) { memcpy(dst, src, bytes); return; }
    const size_t n = (bytes + piece - 1) / piece;
    std::atomic<size_t> next{0};
    g_pool.run((int)std::min<size_t>(n - 1, 64), [&](int) {
        for (;;) {
            const size_t i = next.fetch_add(1);
            if (i >= n) break;
            const size_t o = i * piece;
            memcpy((char*)dst + o, (const char*)src + o, std::min(piece, bytes - o));
        }
    });
}

// ---- pinned ring --------------------------------------------------------------------------------------------
struct StageRing {
    static constexpr int SLOTS = 16;
    static constexpr size_t SLOT_BYTES = 2u << 20;
    uint64_t cursor = 0;                     // pieces staged so far (next slot = cursor % SLOTS)
    void*       slot[SLOTS] = {};
    cudaEvent_t ev[SLOTS] = {};
    std::atomic<uint64_t> turn[SLOTS];       // number of pieces that went through the slot
    bool ok = false;
};

static StageRing* ring_of(Context* c) {
    if (c->ring) return c->ring->ok ? c->ring : nullptr;
    StageRing* r = new StageRing();
    r->ok = true;
    for (int i = 0; i < StageRing::SLOTS; ++i) {
        r->turn[i].store(0);
        if (cudaMallocHost(&r->slot[i], StageRing::SLOT_BYTES) != cudaSuccess ||
            cudaEventCreateWithFlags(&r->ev[i], cudaEventDisableTiming) != cudaSuccess) {
            cudaGetLastError();
            r->ok = false;
            break;
        }
    }
    c->ring = r;
    if (!r->ok) { set_error("pinned staging ring allocation failed"); return nullptr; }
    return r;
}

void stage_ring_release(Context* c) {
    StageRing* r = c->ring;
    if (!r) return;
    for (int i = 0; i < StageRing::SLOTS; ++i) {
        if (r->ev[i]) { cudaEventSynchronize(r->ev[i]); cudaEventDestroy(r->ev[i]); }
        if (r->slot[i]) cudaFreeHost(r->slot[i]);
    }
    cudaGetLastError();
    delete r;
    c->ring = nullptr;
}

bool host_is_pageable(const void* p) {
    if (!p) return false;
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, p) != cudaSuccess) { cudaGetLastError(); return true; }
    return at.type == cudaMemoryTypeUnregistered;
}

static std::atomic<int> g_stage_mode{0};     // 0 = auto (ring for pageable), 1 = never stage (plain cudaMemcpyAsync)

int h2d_copy_segs(Context* c, const H2dSeg* segs, int nSegs, cudaStream_t s) {
    // direct DMA for pinned / small segments; everything else is cut into slot-sized pieces
    struct Piece { char* dst; const char* src; size_t len; };
    std::vector<Piece> pieces;
    const size_t SB = StageRing::SLOT_BYTES;
    const bool noStage = g_stage_mode.load(std::memory_order_relaxed) == 1;
    for (int i = 0; i < nSegs; ++i) {
        const H2dSeg& g = segs[i];
        if (g.bytes == 0) continue;
        if (!g.pageable || g.bytes < (256u << 10) || noStage) {
            MCV_CUDA_TRY(cudaMemcpyAsync(g.dst, g.src, g.bytes, cudaMemcpyHostToDevice, s));
            continue;
        }
        for (size_t o = 0; o < g.bytes; o += SB)
            pieces.push_back(Piece{(char*)g.dst + o, (const char*)g.src + o, std::min(SB, g.bytes - o)});
    }
    const size_t n = pieces.size();
    if (n == 0) return 0;
    StageRing* r = ring_of(c);
    if (!r) return -1;
    // piece i takes slot (cursor + i) % SLOTS as that slot's next turn: the ring is walked round-robin across calls,
    // and turns hand a slot from piece to piece in order
    const uint64_t cursor = r->cursor;
    uint64_t base[StageRing::SLOTS];
    for (int k = 0; k < StageRing::SLOTS; ++k) base[k] = r->turn[k].load();
    std::atomic<size_t> next{0};
    std::atomic<int> err{0};
    const int device = c->device;
    g_pool.run((int)std::min<size_t>(n - 1, 64), [&](int who) {
        if (who != 0) cudaSetDevice(device);
        for (;;) {
            const size_t i = next.fetch_add(1);
            if (i >= n) break;
            const int k = (int)((cursor + i) % StageRing::SLOTS);
            // pieces of this call that used slot k before piece i: those with index i - SLOTS, i - 2 SLOTS, ...
            const uint64_t myTurn = base[k] + i / StageRing::SLOTS;
            while (r->turn[k].load(std::memory_order_acquire) != myTurn) std::this_thread::yield();
            if (myTurn > 0 && cudaEventSynchronize(r->ev[k]) != cudaSuccess) err.store(1);
            memcpy(r->slot[k], pieces[i].src, pieces[i].len);
            if (cudaMemcpyAsync(pieces[i].dst, r->slot[k], pieces[i].len, cudaMemcpyHostToDevice, s) != cudaSuccess ||
                cudaEventRecord(r->ev[k], s) != cudaSuccess)
                err.store(1);
            r->turn[k].store(myTurn + 1, std::memory_order_release);
        }
    });
    r->cursor = cursor + n;
    if (err.load()) {
        set_error("staged host->device copy failed: %s", cudaGetErrorString(cudaGetLastError()));
        return -1;
    }
    return 0;
}

int h2d_copy(Context* c, void* dDst, const void* hSrc, size_t bytes, cudaStream_t s, bool pageable) {
    const H2dSeg g{dDst, hSrc, bytes, pageable};
    return h2d_copy_segs(c, &g, 1, s);
}

}  // namespace mcv

extern "C" {

MCV_API void* mcvAllocPinned(int64_t bytes) {
    mcv::clear_error();
    void* p = nullptr;
    if (bytes <= 0) return nullptr;
    if (cudaHostAlloc(&p, (size_t)bytes, cudaHostAllocPortable) != cudaSuccess) {
        mcv::set_error("cudaHostAlloc(%lld) failed: %s", (long long)bytes, cudaGetErrorString(cudaGetLastError()));
        return nullptr;
    }
    return p;
}

MCV_API void mcvFreePinned(void* p) {
    if (p) { cudaFreeHost(p); cudaGetLastError(); }
}

MCV_API int mcvHostRegister(void* p, int64_t bytes) {
    mcv::clear_error();
    if (!p || bytes <= 0) { mcv::set_error("mcvHostRegister: bad arguments"); return -1; }
    if (cudaHostRegister(p, (size_t)bytes, cudaHostRegisterPortable) != cudaSuccess) {
        mcv::set_error("cudaHostRegister failed: %s", cudaGetErrorString(cudaGetLastError()));
        return -1;
    }
    return 0;
}

MCV_API int mcvHostUnregister(void* p) {
    mcv::clear_error();
    if (!p || cudaHostUnregister(p) != cudaSuccess) {
        mcv::set_error("cudaHostUnregister failed: %s", cudaGetErrorString(cudaGetLastError()));
        return -1;
    }
    return 0;
}

MCV_API int mcvSetStagingMode(int mode) {
    if (mode < 0 || mode > 1) return -1;
    mcv::g_stage_mode.store(mode);
    return 0;
}

MCV_API int mcvCopyThreads(void) { return mcv::copy_threads(); }

// Host<->device copy rate of `device` in GB/s: `reps` copies of `bytes` each way as asked, timed with CUDA events on a
// private stream after one untimed copy.  direction 0 = host->device, 1 = device->host, 2 = both at once (two streams;
// the figure is the sum).  hostKind 0 = CUDA-pinned host memory, 1 = pageable host memory through the staging ring,
// 2 = pageable through plain cudaMemcpyAsync.  < 0 on error.
MCV_API double mcvMeasureCopyBandwidth(int device, int64_t bytes, int reps, int direction, int hostKind) {
    MCV_API_ENTER();
    using namespace mcv;
    if (bytes <= 0 || reps <= 0 || direction < 0 || direction > 2 || hostKind < 0 || hostKind > 2) {
        set_error("mcvMeasureCopyBandwidth: bad arguments");
        return -1;
    }
    Context* c = get_context(device);
    if (!c) return -1;
    void *h0 = nullptr, *h1 = nullptr, *d0 = nullptr, *d1 = nullptr;
    cudaEvent_t e0 = nullptr, e1 = nullptr, f0 = nullptr, f1 = nullptr;
    double result = -1;
    const bool pinned = hostKind == 0;
    const size_t nb = (size_t)bytes;
    auto halloc = [&](void** p) {
        if (pinned) return cudaMallocHost(p, nb) == cudaSuccess;
        *p = malloc(nb);
        return *p != nullptr;
    };
    do {
        if (!halloc(&h0) || !halloc(&h1)) { set_error("host allocation failed"); break; }
        memset(h0, 1, nb); memset(h1, 2, nb);
        if (cudaMalloc(&d0, nb) != cudaSuccess || cudaMalloc(&d1, nb) != cudaSuccess) { set_error("cudaMalloc failed"); break; }
        cudaEventCreate(&e0); cudaEventCreate(&e1); cudaEventCreate(&f0); cudaEventCreate(&f1);
        const bool up = direction != 1, down = direction != 0;
        auto h2d = [&]() {
            if (hostKind == 1) return h2d_copy(c, d0, h0, nb, c->stream2, true);
            return cudaMemcpyAsync(d0, h0, nb, cudaMemcpyHostToDevice, c->stream2) == cudaSuccess ? 0 : -1;
        };
        auto d2h = [&]() { return cudaMemcpyAsync(h1, d1, nb, cudaMemcpyDeviceToHost, c->stream3) == cudaSuccess ? 0 : -1; };
        bool bad = false;
        if (up) bad |= h2d() != 0;
        if (down) bad |= d2h() != 0;
        cudaStreamSynchronize(c->stream2); cudaStreamSynchronize(c->stream3);
        if (up) cudaEventRecord(e0, c->stream2);
        if (down) cudaEventRecord(f0, c->stream3);
        for (int r = 0; r < reps && !bad; ++r) {
            if (up) bad |= h2d() != 0;
            if (down) bad |= d2h() != 0;
        }
        if (up) cudaEventRecord(e1, c->stream2);
        if (down) cudaEventRecord(f1, c->stream3);
        if (cudaStreamSynchronize(c->stream2) != cudaSuccess || cudaStreamSynchronize(c->stream3) != cudaSuccess || bad) {
            set_error("copy failed: %s", cudaGetErrorString(cudaGetLastError()));
            break;
        }
        double gbs = 0;
        float ms = 0;
        if (up && cudaEventElapsedTime(&ms, e0, e1) == cudaSuccess && ms > 0) gbs += (double)nb * reps / (ms * 1e6);
        if (down && cudaEventElapsedTime(&ms, f0, f1) == cudaSuccess && ms > 0) gbs += (double)nb * reps / (ms * 1e6);
        result = gbs;
    } while (false);
    if (e0) cudaEventDestroy(e0);
    if (e1) cudaEventDestroy(e1);
    if (f0) cudaEventDestroy(f0);
    if (f1) cudaEventDestroy(f1);
    if (d0) cudaFree(d0);
    if (d1) cudaFree(d1);
    if (pinned) { if (h0) cudaFreeHost(h0); if (h1) cudaFreeHost(h1); }
    else { free(h0); free(h1); }
    cudaGetLastError();
    return result;
}

// Two-phase form of the same probe, for measuring SEVERAL devices at once (one caller thread or process per device):
// Create allocates the pinned host and the device blocks, the callers synchronise among themselves (bench.py: a
// barrier), Run then keeps both copy directions busy for about `seconds` of wall time and returns what it moved per
// second (direction: 0 host->device, 1 device->host, 2 both concurrently, 3 both in the 4 : 1 byte mix of the C2
// host-buffer path; out[0] = host->device GB/s, out[1] = device->host GB/s).  Long, simultaneous windows are what makes the per-device figures add up to the box's ceiling.
struct CopyProbe {
    uint32_t magic = 0x4D435650u;
    int device = 0;
    size_t bytes = 0;
    void *h0 = nullptr, *h1 = nullptr, *d0 = nullptr, *d1 = nullptr;
    cudaStream_t sUp = nullptr, sDown = nullptr;
    cudaEvent_t e0 = nullptr, e1 = nullptr, f0 = nullptr, f1 = nullptr;
};

MCV_API void mcvCopyProbeDestroy(void* handle) {
    CopyProbe* p = static_cast<CopyProbe*>(handle);
    if (!p || p->magic != 0x4D435650u) return;
    int prev = -1;
    cudaGetDevice(&prev);
    cudaSetDevice(p->device);
    if (p->sUp) { cudaStreamSynchronize(p->sUp); cudaStreamDestroy(p->sUp); }
    if (p->sDown) { cudaStreamSynchronize(p->sDown); cudaStreamDestroy(p->sDown); }
    for (cudaEvent_t e : {p->e0, p->e1, p->f0, p->f1}) if (e) cudaEventDestroy(e);
    if (p->d0) cudaFree(p->d0);
    if (p->d1) cudaFree(p->d1);
    if (p->h0) cudaFreeHost(p->h0);
    if (p->h1) cudaFreeHost(p->h1);
    p->magic = 0;
    delete p;
    if (prev >= 0) cudaSetDevice(prev);
    cudaGetLastError();
}

MCV_API void* mcvCopyProbeCreate(int device, int64_t bytes) {
    MCV_API_ENTER();
    using namespace mcv;
    if (bytes <= 0) { set_error("mcvCopyProbeCreate: bad size"); return nullptr; }
    Context* c = get_context(device);      // validates the device and makes it current
    if (!c) return nullptr;
    CopyProbe* p = new CopyProbe();
    p->device = c->device;
    p->bytes = (size_t)bytes;
    bool ok = cudaMallocHost(&p->h0, p->bytes) == cudaSuccess && cudaMallocHost(&p->h1, p->bytes) == cudaSuccess &&
              cudaMalloc(&p->d0, p->bytes) == cudaSuccess && cudaMalloc(&p->d1, p->bytes) == cudaSuccess &&
              cudaStreamCreateWithFlags(&p->sUp, cudaStreamNonBlocking) == cudaSuccess &&
              cudaStreamCreateWithFlags(&p->sDown, cudaStreamNonBlocking) == cudaSuccess &&
              cudaEventCreate(&p->e0) == cudaSuccess && cudaEventCreate(&p->e1) == cudaSuccess &&
              cudaEventCreate(&p->f0) == cudaSuccess && cudaEventCreate(&p->f1) == cudaSuccess;
    if (ok) {
        memset(p->h0, 1, p->bytes);
        memset(p->h1, 2, p->bytes);
        ok = cudaMemcpyAsync(p->d0, p->h0, p->bytes, cudaMemcpyHostToDevice, p->sUp) == cudaSuccess &&
             cudaMemcpyAsync(p->h1, p->d1, p->bytes, cudaMemcpyDeviceToHost, p->sDown) == cudaSuccess &&
             cudaStreamSynchronize(p->sUp) == cudaSuccess && cudaStreamSynchronize(p->sDown) == cudaSuccess;
    }
    if (!ok) {
        set_error("mcvCopyProbeCreate: allocation / warm-up failed: %s", cudaGetErrorString(cudaGetLastError()));
        mcvCopyProbeDestroy(p);
        return nullptr;
    }
    return p;
}

MCV_API int mcvCopyProbeRun(void* handle, double seconds, int direction, double* outGbs) {
    MCV_API_ENTER();
    using namespace mcv;
    CopyProbe* p = static_cast<CopyProbe*>(handle);
    if (!p || p->magic != 0x4D435650u || direction < 0 || direction > 3 || !outGbs || !(seconds > 0)) {
        set_error("mcvCopyProbeRun: bad arguments");
        return -1;
    }
    if (!get_context(p->device)) return -1;
    const bool up = direction != 1, down = direction != 0;
    // calibrate: copies per batch so that a batch lasts ~10 ms, then batches until the wall time is spent
    const int per = 4 * (int)std::max<double>(1.0, 0.010 * 40e9 / (double)p->bytes / 4);
    int64_t nUp = 0, nDown = 0;
    if (up) cudaEventRecord(p->e0, p->sUp);
    if (down) cudaEventRecord(p->f0, p->sDown);
    timespec t0, t1;
    clock_gettime(CLOCK_MONOTONIC, &t0);
    for (;;) {
        for (int i = 0; i < per; ++i) {
            if (up) { MCV_CUDA_TRY(cudaMemcpyAsync(p->d0, p->h0, p->bytes, cudaMemcpyHostToDevice, p->sUp)); ++nUp; }
            // direction 3: the byte mix of the C2 host-buffer path, one block back for every four blocks up
            if (down && (direction != 3 || (i & 3) == 3)) {
                MCV_CUDA_TRY(cudaMemcpyAsync(p->h1, p->d1, p->bytes, cudaMemcpyDeviceToHost, p->sDown));
                ++nDown;
            }
        }
        // one batch in flight at a time: the loop tracks the wall clock and both directions stay concurrent
        if (up) MCV_CUDA_TRY(cudaStreamSynchronize(p->sUp));
        if (down) MCV_CUDA_TRY(cudaStreamSynchronize(p->sDown));
        clock_gettime(CLOCK_MONOTONIC, &t1);
        if ((t1.tv_sec - t0.tv_sec) + 1e-9 * (t1.tv_nsec - t0.tv_nsec) >= seconds) break;
    }
    if (up) cudaEventRecord(p->e1, p->sUp);
    if (down) cudaEventRecord(p->f1, p->sDown);
    MCV_CUDA_TRY(cudaStreamSynchronize(p->sUp));
    MCV_CUDA_TRY(cudaStreamSynchronize(p->sDown));
    outGbs[0] = outGbs[1] = 0;
    float ms = 0;
    if (up && cudaEventElapsedTime(&ms, p->e0, p->e1) == cudaSuccess && ms > 0) outGbs[0] = (double)p->bytes * nUp / (ms * 1e6);
    if (down && cudaEventElapsedTime(&ms, p->f0, p->f1) == cudaSuccess && ms > 0) outGbs[1] = (double)p->bytes * nDown / (ms * 1e6);
    return 0;
}

}  // extern "C"
