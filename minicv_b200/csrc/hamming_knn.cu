// hamming_knn.cu -- brute-force Hamming kNN (k = 2) + Lowe ratio test on the integer pipe.
//
// Replaces (nothing in the reference -- SURVEY.md section 0: MiniCV ships no matcher); the
// contract is cv2.BFMatcher(NORM_HAMMING).knnMatch(q, t, 2) over the descriptor blob that
// cvDetectFeatures returns (reference src/MiniCVNative/MiniCVNative.cpp:316-341).
//
// Mapping: one query descriptor per thread, held in registers for the whole sweep; the
// train set streams through shared memory in 8 KB stages (cp.async, double-buffered) and
// is read with 128-bit broadcast LDS, so each train word is fetched from HBM/L2 once per
// 256 queries.  Distances are XOR + POPC; the running top-2 lives in two registers as
// packed keys (dist << IDX) | trainIdx, so one unsigned min/max pair updates both ranks
// and the lowest train index wins every tie exactly like BFMatcher's stable sort.
// Bound: integer/POPC pipe (W popc per distance), not HBM: ~100 distances per byte moved.
#include "knn.cuh"

#include <algorithm>
#include <float.h>

namespace mcv {
namespace {

constexpr int QT = 256;  // queries per block, one per thread

template <int W> struct Cfg {
    static constexpr int TT = 2048 / W;                 // train rows per 8 KB stage
    static constexpr int IDX = (W <= 16) ? 22 : 21;     // index bits of the packed key
    static constexpr uint32_t MASK = (1u << IDX) - 1u;
};

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
    unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N> __device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;\n" ::"n"(N));
}

__device__ __forceinline__ void top2_update(uint32_t& m1, uint32_t& m2, uint32_t key) {
    uint32_t hi = max(m1, key);
    m1 = min(m1, key);
    m2 = min(m2, hi);
}

__device__ __forceinline__ void top2_update64(uint64_t& m1, uint64_t& m2, uint64_t key) {
    uint64_t hi = m1 > key ? m1 : key;
    m1 = m1 < key ? m1 : key;
    m2 = m2 < hi ? m2 : hi;
}

// Writes one query's final result.  d1 < ratio*d2 is evaluated in fp32 with the product
// rounded to fp32 (the user-side Lowe test on DMatch.distance floats).
__device__ __forceinline__ bool emit_match(int64_t q, uint64_t k1, uint64_t k2, float ratio,
                                           int* trainIdx, float* dist, uint8_t* ratioPass) {
    int i1 = -1, i2 = -1;
    float d1 = FLT_MAX, d2 = FLT_MAX;
    if (k1 != ~0ull) { i1 = (int)(uint32_t)k1; d1 = (float)(uint32_t)(k1 >> 32); }
    if (k2 != ~0ull) { i2 = (int)(uint32_t)k2; d2 = (float)(uint32_t)(k2 >> 32); }
    reinterpret_cast<int2*>(trainIdx)[q] = make_int2(i1, i2);
    reinterpret_cast<float2*>(dist)[q] = make_float2(d1, d2);
    bool pass = (i2 >= 0) && (d1 < __fmul_rn(ratio, d2));
    if (ratioPass) ratioPass[q] = pass ? 1 : 0;
    return pass;
}

template <int W>
__global__ void __launch_bounds__(QT, 2)
hamming_knn2_kernel(const uint4* __restrict__ query, const uint4* __restrict__ train,
                    const int64_t* __restrict__ qOff, const int64_t* __restrict__ tOff,
                    const int* __restrict__ tileStart, int nPairs, int splitRows, int nSplits,
                    int64_t totalQ, float ratio, int* __restrict__ trainIdx,
                    float* __restrict__ dist, uint8_t* __restrict__ ratioPass,
                    int* __restrict__ passCount, uint64_t* __restrict__ splitKeys) {
    constexpr int TT = Cfg<W>::TT;
    constexpr int IDX = Cfg<W>::IDX;
    constexpr int V = W / 4;  // uint4 per row
    __shared__ uint4 sT[2][TT * V];

    const int tid = threadIdx.x;
    // locate the pair this query tile belongs to
    int lo = 0, hi = nPairs;
    const int tile = blockIdx.x;
    while (hi - lo > 1) {
        int mid = (lo + hi) >> 1;
        if (tileStart[mid] <= tile) lo = mid; else hi = mid;
    }
    const int p = lo;
    const int64_t q0 = qOff[p], qEnd = qOff[p + 1];
    const int64_t tRow0 = tOff[p];
    const int tCount = (int)(tOff[p + 1] - tRow0);
    const int64_t qRow = q0 + (int64_t)(tile - tileStart[p]) * QT + tid;
    const bool valid = qRow < qEnd;

    uint32_t q[W];
#pragma unroll
    for (int j = 0; j < V; ++j) {
        uint4 v = valid ? __ldg(query + qRow * V + j) : make_uint4(0, 0, 0, 0);
        q[4 * j] = v.x; q[4 * j + 1] = v.y; q[4 * j + 2] = v.z; q[4 * j + 3] = v.w;
    }

    const int split = blockIdx.y;
    const int t0 = min(tCount, split * splitRows);
    const int t1 = min(tCount, t0 + splitRows);
    const int nT = t1 - t0;
    const uint4* gT = train + (tRow0 + t0) * V;

    uint32_t m1 = 0xFFFFFFFFu, m2 = 0xFFFFFFFFu;
    const int nTiles = (nT + TT - 1) / TT;

    auto load_tile = [&](int buf, int it) {
        const int rows = min(TT, nT - it * TT);
        const int chunks = rows * V;
        const uint4* g = gT + (int64_t)it * TT * V;
        for (int c = tid; c < chunks; c += QT) cp_async16(&sT[buf][c], g + c);
    };

    if (nTiles > 0) { load_tile(0, 0); }
    cp_async_commit();
    for (int it = 0; it < nTiles; ++it) {
        const int buf = it & 1;
        if (it + 1 < nTiles) load_tile(buf ^ 1, it + 1);
        cp_async_commit();
        cp_async_wait<1>();
        __syncthreads();
        const int rows = min(TT, nT - it * TT);
        const uint4* s = sT[buf];
        const uint32_t base = (uint32_t)(it * TT);
#pragma unroll 4
        for (int r = 0; r < rows; ++r) {
            int d = 0;
#pragma unroll
            for (int j = 0; j < V; ++j) {
                const uint4 v = s[r * V + j];
                d += __popc(q[4 * j] ^ v.x) + __popc(q[4 * j + 1] ^ v.y) +
                     __popc(q[4 * j + 2] ^ v.z) + __popc(q[4 * j + 3] ^ v.w);
            }
            top2_update(m1, m2, ((uint32_t)d << IDX) | (base + (uint32_t)r));
        }
        __syncthreads();
    }
    cp_async_wait<0>();

    // widen to (dist << 32) | pair-local index
    uint64_t k1 = ~0ull, k2 = ~0ull;
    if (m1 != 0xFFFFFFFFu)
        k1 = ((uint64_t)(m1 >> IDX) << 32) | (uint64_t)((m1 & Cfg<W>::MASK) + (uint32_t)t0);
    if (m2 != 0xFFFFFFFFu)
        k2 = ((uint64_t)(m2 >> IDX) << 32) | (uint64_t)((m2 & Cfg<W>::MASK) + (uint32_t)t0);

    if (splitKeys) {
        if (valid) {
            uint64_t* o = splitKeys + ((int64_t)split * totalQ + qRow) * 2;
            reinterpret_cast<ulonglong2*>(o)[0] = make_ulonglong2(k1, k2);
        }
        return;
    }
    bool pass = false;
    if (valid) pass = emit_match(qRow, k1, k2, ratio, trainIdx, dist, ratioPass);
    if (passCount) {
        unsigned b = __ballot_sync(0xFFFFFFFFu, pass);
        if ((tid & 31) == 0 && b) atomicAdd(passCount, __popc(b));
    }
}

// Merges nSplits key pairs per query; emits matches or (shard mode) global keys.
__global__ void __launch_bounds__(256)
knn2_finalize_kernel(const uint64_t* __restrict__ splitKeys, int nSplits, int64_t totalQ,
                     float ratio, int emitKeys, int64_t globalIdxBase, int* __restrict__ trainIdx,
                     float* __restrict__ dist, uint8_t* __restrict__ ratioPass,
                     int* __restrict__ passCount, uint64_t* __restrict__ outKeys) {
    const int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    bool pass = false;
    if (q < totalQ) {
        uint64_t m1 = ~0ull, m2 = ~0ull;
        for (int s = 0; s < nSplits; ++s) {
            const ulonglong2 k =
                reinterpret_cast<const ulonglong2*>(splitKeys + ((int64_t)s * totalQ + q) * 2)[0];
            top2_update64(m1, m2, k.x);
            top2_update64(m1, m2, k.y);
        }
        if (emitKeys) {
            if (m1 != ~0ull) m1 += (uint64_t)globalIdxBase;
            if (m2 != ~0ull) m2 += (uint64_t)globalIdxBase;
            reinterpret_cast<ulonglong2*>(outKeys + q * 2)[0] = make_ulonglong2(m1, m2);
        } else {
            pass = emit_match(q, m1, m2, ratio, trainIdx, dist, ratioPass);
        }
    }
    if (passCount && !emitKeys) {
        unsigned b = __ballot_sync(0xFFFFFFFFu, pass);
        if ((threadIdx.x & 31) == 0 && b) atomicAdd(passCount, __popc(b));
    }
}

// Global top-2 over gathered shard keys [nShards][nQuery][2] (C5, after ncclAllGather).
__global__ void __launch_bounds__(256)
merge_shard_keys_kernel(const uint64_t* __restrict__ keys, int nShards, int nQuery, float ratio,
                        int64_t* __restrict__ trainIdx64, float* __restrict__ dist,
                        uint8_t* __restrict__ ratioPass) {
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= nQuery) return;
    uint64_t m1 = ~0ull, m2 = ~0ull;
    for (int s = 0; s < nShards; ++s) {
        const ulonglong2 k =
            reinterpret_cast<const ulonglong2*>(keys + ((int64_t)s * nQuery + q) * 2)[0];
        top2_update64(m1, m2, k.x);
        top2_update64(m1, m2, k.y);
    }
    int64_t i1 = -1, i2 = -1;
    float d1 = FLT_MAX, d2 = FLT_MAX;
    if (m1 != ~0ull) { i1 = (int64_t)(uint32_t)m1; d1 = (float)(uint32_t)(m1 >> 32); }
    if (m2 != ~0ull) { i2 = (int64_t)(uint32_t)m2; d2 = (float)(uint32_t)(m2 >> 32); }
    trainIdx64[2 * (int64_t)q] = i1;
    trainIdx64[2 * (int64_t)q + 1] = i2;
    dist[2 * (int64_t)q] = d1;
    dist[2 * (int64_t)q + 1] = d2;
    if (ratioPass) ratioPass[q] = ((i2 >= 0) && (d1 < __fmul_rn(ratio, d2))) ? 1 : 0;
}

// Re-packs rows of `srcBytes` bytes into zero-padded rows of `dstWords` 32-bit words
// (61-byte AKAZE rows -> 64 bytes; unaligned blobs -> 16-byte aligned rows).
__global__ void __launch_bounds__(256)
repack_rows_kernel(const uint8_t* __restrict__ src, int srcBytes, uint32_t* __restrict__ dst,
                   int dstWords, int64_t nRows) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nRows * dstWords) return;
    const int64_t row = i / dstWords;
    const int w = (int)(i - row * dstWords);
    const uint8_t* s = src + row * srcBytes;
    uint32_t v = 0;
#pragma unroll
    for (int b = 0; b < 4; ++b) {
        const int o = w * 4 + b;
        if (o < srcBytes) v |= (uint32_t)s[o] << (8 * b);
    }
    dst[i] = v;
}

template <int W>
int launch_kernel_w(cudaStream_t stream, dim3 grid, const uint4* q, const uint4* t,
                    const int64_t* qOff, const int64_t* tOff, const int* tileStart, int nPairs,
                    int splitRows, int nSplits, int64_t totalQ, float ratio, int* trainIdx,
                    float* dist, uint8_t* ratioPass, int* passCount, uint64_t* splitKeys) {
    hamming_knn2_kernel<W><<<grid, QT, 0, stream>>>(q, t, qOff, tOff, tileStart, nPairs, splitRows,
                                                    nSplits, totalQ, ratio, trainIdx, dist,
                                                    ratioPass, passCount, splitKeys);
    MCV_LAUNCH_CHECK();
    return 0;
}

}  // namespace

std::atomic<int> g_hamming_engine{0};

int launch_knn2_finalize(cudaStream_t stream, const uint64_t* dSplitKeys, int nSplits, int64_t totalQ,
                         const KnnArgs& a) {
    knn2_finalize_kernel<<<(unsigned)((totalQ + 255) / 256), 256, 0, stream>>>(
        dSplitKeys, nSplits, totalQ, a.ratio, a.emit == KNN_EMIT_KEYS ? 1 : 0, a.globalIdxBase,
        a.dTrainIdx, a.dDist, a.dRatioPass, a.dPassCount, a.dKeys);
    MCV_LAUNCH_CHECK();
    return 0;
}

int launch_hamming_knn2(Context* ctx, cudaStream_t stream, const KnnArgs& a) {
    if (a.nPairs <= 0) return 0;
    {
        // engine choice: the tensor-core path pays a bit-expansion pass, so small problems stay
        // on the integer pipe
        const int eng = g_hamming_engine.load(std::memory_order_relaxed);
        bool tc = false;
        if (eng == 3 && hamming_tc2_supported(a.descBytes)) return launch_hamming_knn2_tc2(ctx, stream, a);
        if (eng == 2 || eng == 3) tc = hamming_tc_supported(a.descBytes);
        else if (eng == 0 && hamming_tc_supported(a.descBytes)) {
            const double q = (double)(a.hQOff[a.nPairs] - a.hQOff[0]);
            const double t = (double)(a.hTOff[a.nPairs] - a.hTOff[0]) / std::max(a.nPairs, 1);
            tc = q * t >= 64.0 * 1024 * 1024;
        }
        if (tc) return launch_hamming_knn2_tc(ctx, stream, a);
    }
    if (a.descBytes <= 0 || a.descBytes > 128) {
        set_error("Hamming matcher supports 1..128-byte descriptors (got %d)", a.descBytes);
        return -1;
    }
    const int W = a.descBytes <= 16 ? 4 : a.descBytes <= 32 ? 8 : a.descBytes <= 64 ? 16 : 32;
    const int TT = 2048 / W;
    const int idxBits = (W <= 16) ? 22 : 21;

    const int64_t totalQ = a.hQOff[a.nPairs] - a.hQOff[0];
    const int64_t totalT = a.hTOff[a.nPairs] - a.hTOff[0];
    if (a.hQOff[0] != 0 || a.hTOff[0] != 0) {
        set_error("offset arrays must start at 0");
        return -1;
    }
    if (totalQ == 0) return 0;

    // control arrays: qOff, tOff (int64) and tileStart (int32), staged through pinned memory
    const size_t n1 = (size_t)a.nPairs + 1;
    const size_t ctlBytes = n1 * (8 + 8 + 4);
    if (ctx->pin[0].reserve(ctlBytes) || ctx->ws[0].reserve(ctlBytes)) return -1;
    // the previous call's async upload from this pinned block must have drained
    MCV_CUDA_TRY(cudaEventSynchronize(ctx->ev[0]));
    int64_t* hQ = ctx->pin[0].as<int64_t>();
    int64_t* hT = hQ + n1;
    int* hTile = reinterpret_cast<int*>(hT + n1);
    int64_t tiles = 0;
    int64_t maxT = 0;
    for (int p = 0; p < a.nPairs; ++p) {
        const int64_t nq = a.hQOff[p + 1] - a.hQOff[p];
        const int64_t nt = a.hTOff[p + 1] - a.hTOff[p];
        if (nq < 0 || nt < 0 || nt > 0x7FFFFFFFll) {
            set_error("bad offsets at pair %d", p);
            return -1;
        }
        hQ[p] = a.hQOff[p];
        hT[p] = a.hTOff[p];
        hTile[p] = (int)tiles;
        tiles += (nq + QT - 1) / QT;
        maxT = std::max(maxT, nt);
    }
    hQ[a.nPairs] = a.hQOff[a.nPairs];
    hT[a.nPairs] = a.hTOff[a.nPairs];
    hTile[a.nPairs] = (int)tiles;
    if (tiles > 0x7FFFFFFFll) { set_error("too many query tiles"); return -1; }
    MCV_CUDA_TRY(cudaMemcpyAsync(ctx->ws[0].p, hQ, ctlBytes, cudaMemcpyHostToDevice, stream));
    MCV_CUDA_TRY(cudaEventRecord(ctx->ev[0], stream));
    const int64_t* dQOff = ctx->ws[0].as<int64_t>();
    const int64_t* dTOff = dQOff + n1;
    const int* dTile = reinterpret_cast<const int*>(dTOff + n1);

    // row layout: W 32-bit words, 16-byte aligned; repack when the blob is not already so
    const uint8_t* dQ = a.dQuery;
    const uint8_t* dT = a.dTrain;
    const bool packed = (a.descBytes == W * 4) && ((reinterpret_cast<uintptr_t>(dQ) & 15) == 0) &&
                        ((reinterpret_cast<uintptr_t>(dT) & 15) == 0);
    if (!packed) {
        if (ctx->ws[1].reserve((size_t)totalQ * W * 4 + 16) ||
            ctx->ws[2].reserve((size_t)std::max<int64_t>(totalT, 1) * W * 4 + 16))
            return -1;
        const int64_t nq = totalQ * W, nt = totalT * W;
        repack_rows_kernel<<<(unsigned)((nq + 255) / 256), 256, 0, stream>>>(
            a.dQuery, a.descBytes, ctx->ws[1].as<uint32_t>(), W, totalQ);
        MCV_LAUNCH_CHECK();
        if (nt > 0) {
            repack_rows_kernel<<<(unsigned)((nt + 255) / 256), 256, 0, stream>>>(
                a.dTrain, a.descBytes, ctx->ws[2].as<uint32_t>(), W, totalT);
            MCV_LAUNCH_CHECK();
        }
        dQ = ctx->ws[1].as<uint8_t>();
        dT = ctx->ws[2].as<uint8_t>();
    }

    // split the train range when the query tiles alone cannot fill the machine, and
    // always enough that a split's local index fits the packed key
    int64_t wantSplits = 1;
    const int64_t fill = 4ll * ctx->sm_count;
    if (tiles < fill) wantSplits = (fill + tiles - 1) / tiles;
    const int64_t maxTiles = std::max<int64_t>(1, (maxT + TT - 1) / TT);
    wantSplits = std::min<int64_t>(wantSplits, maxTiles);
    const int64_t keyCap = (1ll << idxBits);
    wantSplits = std::max<int64_t>(wantSplits, (maxT + keyCap - 1) / keyCap);
    int64_t splitRows = (std::max<int64_t>(maxT, 1) + wantSplits - 1) / wantSplits;
    splitRows = (splitRows + TT - 1) / TT * TT;
    int nSplits = (int)std::max<int64_t>(1, (maxT + splitRows - 1) / splitRows);
    if (nSplits > 65535) { set_error("train set too large for one launch"); return -1; }

    const bool twoStage = nSplits > 1 || a.emit == KNN_EMIT_KEYS;
    uint64_t* dSplitKeys = nullptr;
    if (twoStage) {
        if (ctx->ws[3].reserve((size_t)nSplits * totalQ * 16)) return -1;
        dSplitKeys = ctx->ws[3].as<uint64_t>();
    }
    dim3 grid((unsigned)tiles, (unsigned)nSplits);
    const uint4* q4 = reinterpret_cast<const uint4*>(dQ);
    const uint4* t4 = reinterpret_cast<const uint4*>(dT);
    int rc;
#define MCV_W(WW)                                                                             \
    rc = launch_kernel_w<WW>(stream, grid, q4, t4, dQOff, dTOff, dTile, a.nPairs,             \
                             (int)splitRows, nSplits, totalQ, a.ratio, a.dTrainIdx, a.dDist,  \
                             a.dRatioPass, a.dPassCount, dSplitKeys)
    kt_begin(ctx, stream);
    switch (W) {
        case 4: MCV_W(4); break;
        case 8: MCV_W(8); break;
        case 16: MCV_W(16); break;
        default: MCV_W(32); break;
    }
#undef MCV_W
    kt_end(ctx, stream);
    if (rc) return rc;
    if (twoStage) return launch_knn2_finalize(stream, dSplitKeys, nSplits, totalQ, a);
    return 0;
}

int launch_merge_keys(Context* ctx, cudaStream_t stream, const uint64_t* dKeys, int nShards,
                      int nQuery, float ratio, int64_t* dTrainIdx64, float* dDist,
                      uint8_t* dRatioPass) {
    (void)ctx;
    if (nQuery <= 0) return 0;
    merge_shard_keys_kernel<<<(nQuery + 255) / 256, 256, 0, stream>>>(
        dKeys, nShards, nQuery, ratio, dTrainIdx64, dDist, dRatioPass);
    MCV_LAUNCH_CHECK();
    return 0;
}

}  // namespace mcv
