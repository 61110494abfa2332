// shard.cu -- large-database matching (config C5) behind the C ABI.
//
// The TRAIN set is sharded by contiguous index range over the GPUs, the queries are replicated.  Every GPU emits its
// local top-2 per query as packed keys (dist << 32) | globalIdx (hamming_knn.cu / hamming_tc.cu, KNN_EMIT_KEYS), one
// ncclAllGather of nQuery x 2 x 8 bytes per rank over NVLink / NVSwitch collects them, and the two smallest keys per
// query are the global result -- BFMatcher's lowest-index tie-break survives because the index is the low word.
//
// Two forms:
//   (i)  in-process: cvLargeDbCreate / cvLargeDbMatch / cvLargeDbDestroy keep the shards resident on the devices of a
//        mask; one call from the reference's side of the boundary drives all of them (ncclCommInitAll, grouped
//        all-gather issued by the calling thread);
//   (ii) one process per GPU (torchrun, MPI, ...): mcvShardCommUniqueId / mcvShardCommInit build the communicator
//        from an id the caller distributes, cvMatchKnn2ShardedDevice = local keys -> ncclAllGather -> merge on the
//        caller's stream.
// NCCL is loaded at run time (dlopen "libnccl.so.2", $MINICV_NCCL_LIB overrides), so the library itself has no link
// dependency on it and processes that already carry an NCCL (torch) share that copy.  Without a loadable NCCL the
// in-process form moves the keys with peer copies (cudaMemcpyPeerAsync) to the first device instead; form (ii) fails.
#include "knn.cuh"
#include "staging.cuh"

#include <dlfcn.h>
#include <stdlib.h>
#include <algorithm>
#include <vector>
#include <string>

namespace mcv {
namespace {

// the slice of nccl.h this file needs (NCCL 2.x ABI)
typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
typedef int ncclResult_t;
constexpr int ncclUint64 = 5;

struct Nccl {
    void* h = nullptr;
    bool tried = false;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommInitAll)(ncclComm_t*, int, const int*) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    ncclResult_t (*GetVersion)(int*) = nullptr;
};
Nccl g_nccl;
std::mutex g_nccl_mu;

Nccl* nccl() {
    std::lock_guard<std::mutex> lk(g_nccl_mu);
    if (g_nccl.tried) return g_nccl.h ? &g_nccl : nullptr;
    g_nccl.tried = true;
    const char* env = getenv("MINICV_NCCL_LIB");
    const char* names[] = {env, "libnccl.so.2", "libnccl.so"};
    void* h = nullptr;
    for (const char* n : names) {
        if (!n || !*n) continue;
        h = dlopen(n, RTLD_NOW | RTLD_LOCAL);          // every entry point is taken with dlsym: nothing leaks into the host's namespace
        if (h) break;
    }
    if (!h) return nullptr;
    bool ok = true;
    auto sym = [&](const char* n) { void* p = dlsym(h, n); if (!p) ok = false; return p; };
    g_nccl.GetUniqueId = reinterpret_cast<decltype(g_nccl.GetUniqueId)>(sym("ncclGetUniqueId"));
    g_nccl.CommInitRank = reinterpret_cast<decltype(g_nccl.CommInitRank)>(sym("ncclCommInitRank"));
    g_nccl.CommInitAll = reinterpret_cast<decltype(g_nccl.CommInitAll)>(sym("ncclCommInitAll"));
    g_nccl.CommDestroy = reinterpret_cast<decltype(g_nccl.CommDestroy)>(sym("ncclCommDestroy"));
    g_nccl.AllGather = reinterpret_cast<decltype(g_nccl.AllGather)>(sym("ncclAllGather"));
    g_nccl.GroupStart = reinterpret_cast<decltype(g_nccl.GroupStart)>(sym("ncclGroupStart"));
    g_nccl.GroupEnd = reinterpret_cast<decltype(g_nccl.GroupEnd)>(sym("ncclGroupEnd"));
    g_nccl.GetErrorString = reinterpret_cast<decltype(g_nccl.GetErrorString)>(sym("ncclGetErrorString"));
    g_nccl.GetVersion = reinterpret_cast<decltype(g_nccl.GetVersion)>(sym("ncclGetVersion"));
    if (!ok) { dlclose(h); return nullptr; }
    g_nccl.h = h;
    return &g_nccl;
}

#define MCV_NCCL_TRY(n, expr)                                                                           \
    do {                                                                                                \
        ncclResult_t _r = (expr);                                                                       \
        if (_r != 0) {                                                                                  \
            ::mcv::set_error("%s failed: %s (%s:%d)", #expr, (n)->GetErrorString(_r), __FILE__, __LINE__); \
            return -1;                                                                                  \
        }                                                                                               \
    } while (0)

constexpr uint32_t SHARD_COMM_MAGIC = 0x4D435643u, LARGE_DB_MAGIC = 0x4D43564Cu;

struct ShardComm {          // form (ii)
    uint32_t magic = SHARD_COMM_MAGIC;
    ncclComm_t comm = nullptr;
    int nRanks = 1, rank = 0, device = 0;
    DevBuf keys, gathered;
};

struct DbShard {
    int device = 0;
    int64_t base = 0, rows = 0;
    uint8_t* dTrain = nullptr;
    ncclComm_t comm = nullptr;
    cudaStream_t stream = nullptr;     // private: the shard outlives any leased context
    cudaEvent_t done = nullptr;
    DevBuf query, keys, gathered;
};
struct LargeDb {            // form (i)
    uint32_t magic = LARGE_DB_MAGIC;
    int descBytes = 0;
    int64_t nTrain = 0;
    bool useNccl = false;
    std::vector<DbShard> shards;
    DevBuf idx64, dist, pass;          // on shards[0].device
    PinBuf out;
    std::mutex mu;                     // one match at a time per database
};

int shard_keys(Context* ctx, cudaStream_t st, const uint8_t* dQuery, int nQuery, const uint8_t* dTrain, int64_t nTrain,
               int64_t base, int descBytes, uint64_t* dKeys) {
    if (nTrain > 0x7FFFFFFFll) { set_error("shard of %lld rows exceeds the 31-bit local index", (long long)nTrain); return -1; }
    if (base < 0 || base + nTrain > (1ll << 32)) {
        set_error("global train index %lld + %lld does not fit the 32-bit index field of the packed key", (long long)base, (long long)nTrain);
        return -1;
    }
    const int64_t qo[2] = {0, nQuery}, to[2] = {0, nTrain};
    KnnArgs a;
    a.dQuery = dQuery; a.dTrain = dTrain; a.hQOff = qo; a.hTOff = to;
    a.nPairs = 1; a.descBytes = descBytes; a.emit = KNN_EMIT_KEYS;
    a.globalIdxBase = base; a.dKeys = dKeys;
    return launch_hamming_knn2(ctx, st, a);
}

}  // namespace
}  // namespace mcv

using namespace mcv;

extern "C" {

MCV_API int mcvNcclVersion(void) {
    Nccl* n = nccl();
    int v = 0;
    if (!n || n->GetVersion(&v) != 0) return 0;
    return v;
}

// ---- form (ii): one process per GPU ---------------------------------------------------------------------------
MCV_API int mcvShardCommUniqueId(uint8_t* id128) {
    MCV_API_ENTER();
    Nccl* n = nccl();
    if (!n) { set_error("NCCL is not loadable (libnccl.so.2; set MINICV_NCCL_LIB)"); return -1; }
    if (!id128) { set_error("mcvShardCommUniqueId: null buffer"); return -1; }
    ncclUniqueId id;
    MCV_NCCL_TRY(n, n->GetUniqueId(&id));
    memcpy(id128, id.internal, 128);
    return 0;
}

MCV_API void* mcvShardCommInit(const uint8_t* id128, int nRanks, int rank, int device) {
    MCV_API_ENTER();
    Nccl* n = nccl();
    if (!n) { set_error("NCCL is not loadable (libnccl.so.2; set MINICV_NCCL_LIB)"); return nullptr; }
    if (!id128 || nRanks < 1 || rank < 0 || rank >= nRanks) { set_error("mcvShardCommInit: bad arguments"); return nullptr; }
    Context* ctx = get_context(device);
    if (!ctx) return nullptr;
    ShardComm* c = new ShardComm();
    c->nRanks = nRanks; c->rank = rank; c->device = ctx->device;
    ncclUniqueId id;
    memcpy(id.internal, id128, 128);
    const ncclResult_t r = n->CommInitRank(&c->comm, nRanks, id, rank);
    if (r != 0) {
        set_error("ncclCommInitRank failed: %s", n->GetErrorString(r));
        delete c;
        return nullptr;
    }
    return c;
}

MCV_API void mcvShardCommDestroy(void* comm) {
    MCV_API_ENTER();
    ShardComm* c = static_cast<ShardComm*>(comm);
    if (!c || c->magic != SHARD_COMM_MAGIC) return;
    Nccl* n = nccl();
    int prev = -1;
    cudaGetDevice(&prev);
    cudaSetDevice(c->device);
    cudaDeviceSynchronize();
    if (n && c->comm) n->CommDestroy(c->comm);
    c->keys.release(); c->gathered.release();
    c->magic = 0;
    delete c;
    if (prev >= 0) cudaSetDevice(prev);
    cudaGetLastError();
}

MCV_API int cvMatchKnn2ShardedDevice(void* comm, const uint8_t* dQuery, int nQuery, const uint8_t* dTrainShard,
                                     int64_t nTrainShard, int64_t globalIdxBase, int descBytes, float ratio,
                                     int64_t* dTrainIdx64, float* dDist, uint8_t* dRatioPass, void* stream) {
    MCV_API_ENTER();
    ShardComm* c = static_cast<ShardComm*>(comm);
    if (!c || c->magic != SHARD_COMM_MAGIC) { set_error("cvMatchKnn2ShardedDevice: bad communicator handle"); return -1; }
    if (nQuery < 0 || nTrainShard < 0 || !dTrainIdx64 || !dDist) { set_error("cvMatchKnn2ShardedDevice: bad arguments"); return -1; }
    if (nQuery == 0) return 0;
    Nccl* n = nccl();
    if (!n) { set_error("NCCL is not loadable"); return -1; }
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    Context* ctx = get_context_async(c->device, st);
    if (!ctx) return -1;
    const size_t kb = (size_t)nQuery * 16;
    if (c->keys.reserve(kb) || c->gathered.reserve(kb * c->nRanks)) return -1;
    if (shard_keys(ctx, st, dQuery, nQuery, dTrainShard, nTrainShard, globalIdxBase, descBytes, c->keys.as<uint64_t>())) return -1;
    const uint64_t* all = c->keys.as<uint64_t>();
    if (c->nRanks > 1) {
        MCV_NCCL_TRY(n, n->AllGather(c->keys.p, c->gathered.p, (size_t)nQuery * 2, ncclUint64, c->comm, st));
        all = c->gathered.as<uint64_t>();
    }
    return launch_merge_keys(ctx, st, all, c->nRanks, nQuery, ratio, dTrainIdx64, dDist, dRatioPass);
}

// ---- form (i): in-process, shards resident on the devices of a mask -----------------------------------------------
MCV_API void cvLargeDbDestroy(void* db);

MCV_API void* cvLargeDbCreate(const uint8_t* train, int64_t nTrain, int descBytes, uint32_t deviceMask) {
    MCV_API_ENTER();
    if (!train || nTrain <= 0 || descBytes <= 0 || descBytes > 128) { set_error("cvLargeDbCreate: bad arguments"); return nullptr; }
    if (nTrain > (1ll << 32)) { set_error("cvLargeDbCreate: at most 2^32 train rows (32-bit index field of the packed key)"); return nullptr; }
    int nDevAll = 0;
    if (cudaGetDeviceCount(&nDevAll) != cudaSuccess || nDevAll <= 0) {
        cudaGetLastError();
        set_error("no usable CUDA device; libMiniCVNative has no CPU fallback");
        return nullptr;
    }
    if (deviceMask == 0) deviceMask = nDevAll >= 32 ? 0xFFFFFFFFu : ((1u << nDevAll) - 1u);
    std::vector<int> devs;
    for (int d = 0; d < 32; ++d)
        if (deviceMask & (1u << d)) {
            if (d >= nDevAll) { set_error("deviceMask names device %d but only %d are visible", d, nDevAll); return nullptr; }
            devs.push_back(d);
        }
    const int nS = (int)std::min<int64_t>((int64_t)devs.size(), nTrain);
    LargeDb* db = new LargeDb();
    db->descBytes = descBytes; db->nTrain = nTrain;
    db->shards.resize(nS);
    const bool pageable = host_is_pageable(train);
    bool ok = true;
    const int64_t per = nTrain / nS, rem = nTrain % nS;
    for (int s = 0; s < nS && ok; ++s) {
        DbShard& sh = db->shards[s];
        sh.device = devs[s];
        sh.base = s * per + std::min<int64_t>(s, rem);
        sh.rows = per + (s < rem ? 1 : 0);
        Context* ctx = get_context(sh.device);       // also makes the device current
        if (!ctx) { ok = false; break; }
        const size_t bytes = (size_t)sh.rows * descBytes;
        if (cudaMalloc(&sh.dTrain, bytes + 16) != cudaSuccess ||
            cudaStreamCreateWithFlags(&sh.stream, cudaStreamNonBlocking) != cudaSuccess ||
            cudaEventCreateWithFlags(&sh.done, cudaEventDisableTiming) != cudaSuccess) {
            set_error("cvLargeDbCreate: allocation on device %d failed: %s", sh.device, cudaGetErrorString(cudaGetLastError()));
            ok = false;
            break;
        }
        if (h2d_copy(ctx, sh.dTrain, train + (size_t)sh.base * descBytes, bytes, sh.stream, pageable)) { ok = false; break; }
    }
    for (int s = 0; s < nS && ok; ++s) {
        cudaSetDevice(db->shards[s].device);
        if (cudaStreamSynchronize(db->shards[s].stream) != cudaSuccess) { set_error("cvLargeDbCreate: upload failed"); ok = false; }
    }
    if (ok && nS > 1) {
        const char* tr = getenv("MINICV_SHARD_TRANSPORT");
        Nccl* n = (tr && !strcmp(tr, "p2p")) ? nullptr : nccl();
        if (n) {
            std::vector<ncclComm_t> comms(nS);
            std::vector<int> dl(nS);
            for (int s = 0; s < nS; ++s) dl[s] = db->shards[s].device;
            const ncclResult_t r = n->CommInitAll(comms.data(), nS, dl.data());
            if (r != 0) { set_error("ncclCommInitAll failed: %s", n->GetErrorString(r)); ok = false; }
            else {
                for (int s = 0; s < nS; ++s) db->shards[s].comm = comms[s];
                db->useNccl = true;
            }
        } else {
            // peer copies to the first device instead
            cudaSetDevice(db->shards[0].device);
            for (int s = 1; s < nS; ++s) {
                int can = 0;
                cudaDeviceCanAccessPeer(&can, db->shards[0].device, db->shards[s].device);
                if (can) { cudaDeviceEnablePeerAccess(db->shards[s].device, 0); cudaGetLastError(); }
            }
        }
    }
    if (!ok) {
        const std::string keep = get_error();
        cvLargeDbDestroy(db);
        set_error("%s", keep.c_str());
        return nullptr;
    }
    return db;
}

MCV_API void cvLargeDbDestroy(void* handle) {
    LargeDb* db = static_cast<LargeDb*>(handle);
    if (!db || db->magic != LARGE_DB_MAGIC) return;
    int prev = -1;
    cudaGetDevice(&prev);
    Nccl* n = db->useNccl ? nccl() : nullptr;
    for (DbShard& sh : db->shards) {
        cudaSetDevice(sh.device);
        if (sh.stream) cudaStreamSynchronize(sh.stream);
        if (n && sh.comm) n->CommDestroy(sh.comm);
        if (sh.dTrain) cudaFree(sh.dTrain);
        if (sh.done) cudaEventDestroy(sh.done);
        if (sh.stream) cudaStreamDestroy(sh.stream);
        sh.query.release(); sh.keys.release(); sh.gathered.release();
    }
    if (!db->shards.empty()) cudaSetDevice(db->shards[0].device);
    db->idx64.release(); db->dist.release(); db->pass.release(); db->out.release();
    db->magic = 0;
    delete db;
    if (prev >= 0) cudaSetDevice(prev);
    cudaGetLastError();
}

MCV_API int cvLargeDbInfo(void* handle, int* nShards, int* usesNccl, int64_t* nTrain) {
    LargeDb* db = static_cast<LargeDb*>(handle);
    if (!db || db->magic != LARGE_DB_MAGIC) return -1;
    if (nShards) *nShards = (int)db->shards.size();
    if (usesNccl) *usesNccl = db->useNccl ? 1 : 0;
    if (nTrain) *nTrain = db->nTrain;
    return 0;
}

static int large_db_match(LargeDb* db, const uint8_t* query, int nQuery, float ratio, int64_t* trainIdx64,
                          float* dist, uint8_t* ratioPass) {
    const int nS = (int)db->shards.size();
    const size_t qb = (size_t)nQuery * db->descBytes, kb = (size_t)nQuery * 16;
    Nccl* n = db->useNccl ? nccl() : nullptr;
    // local top-2 keys on every shard
    for (int s = 0; s < nS; ++s) {
        DbShard& sh = db->shards[s];
        Context* ctx = get_context(sh.device);
        if (!ctx) return -1;
        if (sh.query.reserve(qb + 16) || sh.keys.reserve(kb) || ((db->useNccl || s == 0) && sh.gathered.reserve(kb * nS))) return -1;
        MCV_CUDA_TRY(cudaMemcpyAsync(sh.query.p, query, qb, cudaMemcpyHostToDevice, sh.stream));
        uint64_t* keys = (nS > 1 && !db->useNccl && s == 0) ? sh.gathered.as<uint64_t>() : sh.keys.as<uint64_t>();
        if (shard_keys(ctx, sh.stream, sh.query.as<uint8_t>(), nQuery, sh.dTrain, sh.rows, sh.base, db->descBytes, keys)) return -1;
    }
    // exchange
    DbShard& s0 = db->shards[0];
    const uint64_t* all = s0.keys.as<uint64_t>();
    if (nS > 1 && db->useNccl) {
        MCV_NCCL_TRY(n, n->GroupStart());
        for (int s = 0; s < nS; ++s) {
            DbShard& sh = db->shards[s];
            const ncclResult_t r = n->AllGather(sh.keys.p, sh.gathered.p, (size_t)nQuery * 2, ncclUint64, sh.comm, sh.stream);
            if (r != 0) { n->GroupEnd(); set_error("ncclAllGather failed: %s", n->GetErrorString(r)); return -1; }
        }
        MCV_NCCL_TRY(n, n->GroupEnd());
        all = s0.gathered.as<uint64_t>();
    } else if (nS > 1) {
        for (int s = 1; s < nS; ++s) {
            DbShard& sh = db->shards[s];
            MCV_CUDA_TRY(cudaSetDevice(sh.device));
            MCV_CUDA_TRY(cudaMemcpyPeerAsync(s0.gathered.as<uint8_t>() + kb * s, s0.device, sh.keys.p, sh.device, kb, sh.stream));
            MCV_CUDA_TRY(cudaEventRecord(sh.done, sh.stream));
        }
        MCV_CUDA_TRY(cudaSetDevice(s0.device));
        for (int s = 1; s < nS; ++s) MCV_CUDA_TRY(cudaStreamWaitEvent(s0.stream, db->shards[s].done, 0));
        all = s0.gathered.as<uint64_t>();
    }
    // merge + ratio test on the first device, results to the host
    Context* c0 = get_context(s0.device);
    if (!c0) return -1;
    const size_t nq = (size_t)nQuery;
    if (db->idx64.reserve(nq * 16) || db->dist.reserve(nq * 8) || db->pass.reserve(nq) || db->out.reserve(nq * 25 + 64)) return -1;
    if (launch_merge_keys(c0, s0.stream, all, nS, nQuery, ratio, db->idx64.as<int64_t>(), db->dist.as<float>(), db->pass.as<uint8_t>()))
        return -1;
    uint8_t* ho = db->out.as<uint8_t>();
    MCV_CUDA_TRY(cudaMemcpyAsync(ho, db->idx64.p, nq * 16, cudaMemcpyDeviceToHost, s0.stream));
    MCV_CUDA_TRY(cudaMemcpyAsync(ho + nq * 16, db->dist.p, nq * 8, cudaMemcpyDeviceToHost, s0.stream));
    MCV_CUDA_TRY(cudaMemcpyAsync(ho + nq * 24, db->pass.p, nq, cudaMemcpyDeviceToHost, s0.stream));
    MCV_CUDA_TRY(cudaStreamSynchronize(s0.stream));
    if (db->useNccl)
        for (int s = 1; s < nS; ++s) {      // every rank of the collective has finished before the buffers are reused
            MCV_CUDA_TRY(cudaSetDevice(db->shards[s].device));
            MCV_CUDA_TRY(cudaStreamSynchronize(db->shards[s].stream));
        }
    memcpy(trainIdx64, ho, nq * 16);
    memcpy(dist, ho + nq * 16, nq * 8);
    int pass = 0;
    const uint8_t* hp = ho + nq * 24;
    for (size_t i = 0; i < nq; ++i) pass += hp[i];
    if (ratioPass) memcpy(ratioPass, hp, nq);
    return pass;
}

MCV_API int cvLargeDbMatch(void* handle, const uint8_t* query, int nQuery, float ratio, int64_t* trainIdx64,
                           float* dist, uint8_t* ratioPass) {
    MCV_API_ENTER();
    LargeDb* db = static_cast<LargeDb*>(handle);
    if (!db || db->magic != LARGE_DB_MAGIC) { set_error("cvLargeDbMatch: bad database handle"); return -1; }
    if (nQuery < 0 || (nQuery > 0 && (!query || !trainIdx64 || !dist))) { set_error("cvLargeDbMatch: bad arguments"); return -1; }
    if (nQuery == 0) return 0;
    std::lock_guard<std::mutex> lk(db->mu);
    const int rc = large_db_match(db, query, nQuery, ratio, trainIdx64, dist, ratioPass);
    if (rc < 0)
        for (DbShard& sh : db->shards) {     // nothing of a failed call may still be running on leased scratch
            cudaSetDevice(sh.device);
            cudaStreamSynchronize(sh.stream);
            cudaGetLastError();
        }
    return rc;
}

}  // extern "C"
