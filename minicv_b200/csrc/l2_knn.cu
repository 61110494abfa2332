// l2_knn.cu -- brute-force L2 kNN (k = 2) + ratio test for float32 descriptors (SIFT-style).
//
// Contract: cv2.BFMatcher(NORM_L2).knnMatch(q, t, 2) (the reference
// ships no matcher), bit for bit -- which for floating point means OpenCV's accumulation ORDER, not just its
// formula.  BFMatcher -> batchDistance -> sqrt(hal::normL2Sqr_) (core/src/norm.cpp), the universal-intrinsic loop at
// 128-bit width without FMA (restated and pinned in oracle/bf_knn.py l2_matrix_cv against
// tests/golden/knn_l2_general.npz):
//     element e of the first 16 * floor(D / 16) goes to partial sum (accumulator a = (e / 4) % 4, lane l = e % 4),
//     s[l] = ((acc[0][l] + acc[1][l]) + acc[2][l]) + acc[3][l],   d = (s[0] + s[2]) + (s[1] + s[3]),
//     then the remaining D % 16 elements are added one by one; dist = sqrtf(d).
// l2_sqr_cv_order() below is that order with ONE partial sum live at a time (the 16 strided partial sums are
// independent, so they can be finished one after the other), which keeps a 4 x 4 register micro-tile affordable.
//
// Mapping: TLT queries x TLT train rows per block step, 16 x 16 threads with a (TLT/16)^2 register
// micro-tile each; both operand tiles sit in shared memory with ALL dimensions, transposed to [k][row], because
// the order walks the dimensions with stride 16.  Every thread keeps the top-2 of its queries as packed keys
// (float bits << 32 | trainIdx) -- non-negative floats order like their bit patterns, so one unsigned min/max
// pair updates both ranks with the lowest index winning ties; the 16 column threads of a query merge through
// shared memory.  This is the fp32 CUDA-core engine (FP32-pipe bound: 3 flops per dimension, no FMA); the tcgen05
// GEMM + exact re-scoring engine (l2_tc.cu) replaces it for large batches.
#include "knn.cuh"
#include "l2_order.cuh"

#include <algorithm>
#include <float.h>

namespace mcv {
namespace {


__device__ __forceinline__ void top2_u64(uint64_t& m1, uint64_t& m2, uint64_t key) {
    const uint64_t hi = m1 > key ? m1 : key;
    m1 = m1 < key ? m1 : key;
    m2 = m2 < hi ? m2 : hi;
}

template <int TLT>
__global__ void __launch_bounds__(256)
l2_knn2_kernel(const float* __restrict__ query, const float* __restrict__ train,
               const int64_t* __restrict__ qOff, const int64_t* __restrict__ tOff,
               const int* __restrict__ tileStart, int nPairs, int dim, int splitRows, int64_t totalQ,
               uint64_t* __restrict__ splitKeys) {
    constexpr int MT = TLT / 16;                 // micro-tile edge
    constexpr int LD = TLT + 1;                  // padded row length of the transposed tiles
    extern __shared__ __align__(16) float smem[];
    float* sQ = smem;                            // [dim][LD]
    float* sT = smem + (size_t)dim * LD;         // [dim][LD]
    __shared__ uint64_t sM[TLT][16][2];
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    int lo = 0, hi = nPairs;
    const int tile = blockIdx.x;
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (tileStart[mid] <= tile) lo = mid; else hi = mid;
    }
    const int p = lo;
    const int64_t q0 = qOff[p] + (int64_t)(tile - tileStart[p]) * TLT, qEnd = qOff[p + 1];
    const int64_t tRow0 = tOff[p];
    const int tCount = (int)(tOff[p + 1] - tRow0);
    const int split = blockIdx.y;
    const int t0 = min(tCount, split * splitRows), t1 = min(tCount, t0 + splitRows);

    uint64_t m1[MT], m2[MT];
#pragma unroll
    for (int i = 0; i < MT; ++i) { m1[i] = ~0ull; m2[i] = ~0ull; }

    // the query tile: all dimensions, once
    for (int e = threadIdx.x; e < TLT * dim; e += 256) {
        const int r = e / dim, k = e - r * dim;
        const int64_t qr = q0 + r;
        sQ[k * LD + r] = qr < qEnd ? query[qr * dim + k] : 0.f;
    }
    for (int tb = t0; tb < t1; tb += TLT) {
        __syncthreads();
        for (int e = threadIdx.x; e < TLT * dim; e += 256) {
            const int r = e / dim, k = e - r * dim;
            const int tr = tb + r;
            sT[k * LD + r] = tr < t1 ? train[(tRow0 + tr) * dim + k] : 0.f;
        }
        __syncthreads();
        float d[MT][MT];
        l2_sqr_cv_order<MT>(sQ + ty * MT, sT + tx * MT, LD, dim, d);
#pragma unroll
        for (int i = 0; i < MT; ++i)
#pragma unroll
            for (int j = 0; j < MT; ++j) {
                const int tr = tb + tx * MT + j;
                if (tr < t1) {
                    const float dist = __fsqrt_rn(d[i][j]);
                    top2_u64(m1[i], m2[i], ((uint64_t)__float_as_uint(dist) << 32) | (uint32_t)tr);
                }
            }
    }
    // merge the 16 column threads of each query
#pragma unroll
    for (int i = 0; i < MT; ++i) { sM[ty * MT + i][tx][0] = m1[i]; sM[ty * MT + i][tx][1] = m2[i]; }
    __syncthreads();
    if (threadIdx.x < TLT) {
        const int r = threadIdx.x;
        uint64_t a = ~0ull, b = ~0ull;
        for (int c = 0; c < 16; ++c) { top2_u64(a, b, sM[r][c][0]); top2_u64(a, b, sM[r][c][1]); }
        const int64_t qr = q0 + r;
        if (qr < qEnd)
            reinterpret_cast<ulonglong2*>(splitKeys + ((int64_t)split * totalQ + qr) * 2)[0] = make_ulonglong2(a, b);
    }
}

// keys carry float bits: same merge as the Hamming finalize, but the distance word is a float
__global__ void __launch_bounds__(256)
l2_finalize_kernel(const uint64_t* __restrict__ splitKeys, int nSplits, int64_t totalQ, float ratio,
                   int* __restrict__ trainIdx, float* __restrict__ dist, uint8_t* __restrict__ ratioPass,
                   int* __restrict__ passCount) {
    const int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    bool pass = false;
    if (q < totalQ) {
        uint64_t m1 = ~0ull, m2 = ~0ull;
        for (int s = 0; s < nSplits; ++s) {
            const ulonglong2 k = reinterpret_cast<const ulonglong2*>(splitKeys + ((int64_t)s * totalQ + q) * 2)[0];
            top2_u64(m1, m2, k.x);
            top2_u64(m1, m2, k.y);
        }
        int i1 = -1, i2 = -1;
        float d1 = FLT_MAX, d2 = FLT_MAX;
        if (m1 != ~0ull) { i1 = (int)(uint32_t)m1; d1 = __uint_as_float((uint32_t)(m1 >> 32)); }
        if (m2 != ~0ull) { i2 = (int)(uint32_t)m2; d2 = __uint_as_float((uint32_t)(m2 >> 32)); }
        reinterpret_cast<int2*>(trainIdx)[q] = make_int2(i1, i2);
        reinterpret_cast<float2*>(dist)[q] = make_float2(d1, d2);
        pass = (i2 >= 0) && (d1 < __fmul_rn(ratio, d2));
        if (ratioPass) ratioPass[q] = pass ? 1 : 0;
    }
    if (passCount) {
        const unsigned b = __ballot_sync(0xFFFFFFFFu, pass);
        if ((threadIdx.x & 31) == 0 && b) atomicAdd(passCount, __popc(b));
    }
}

}  // namespace

std::atomic<int> g_l2_engine{0};

int launch_l2_knn2(Context* ctx, cudaStream_t stream, const KnnArgs& a) {
    if (a.nPairs <= 0) return 0;
    if (a.descBytes <= 0 || a.descBytes % 4 != 0) { set_error("L2 matcher needs float32 rows"); return -1; }
    if (a.emit != KNN_EMIT_MATCHES) { set_error("L2 matcher has no shard-key mode"); return -1; }
    {
        // engine switch (mcvSetL2Engine): 1 = CUDA cores, 2 = tensor cores, 0 = auto
        const int eng = g_l2_engine.load(std::memory_order_relaxed);
        bool tcEng = false;
        if (eng >= 2) tcEng = l2_tc_supported(a.descBytes);
        else if (eng == 0 && l2_tc_supported(a.descBytes)) {
            const double q = (double)a.hQOff[a.nPairs];
            const double t = (double)a.hTOff[a.nPairs] / std::max(a.nPairs, 1);
            tcEng = q * t >= 16.0 * 1024 * 1024;
        }
        if (tcEng && (reinterpret_cast<uintptr_t>(a.dQuery) & 3) == 0 && (reinterpret_cast<uintptr_t>(a.dTrain) & 3) == 0)
            return launch_l2_knn2_tc(ctx, stream, a);
    }
    if ((reinterpret_cast<uintptr_t>(a.dQuery) & 3) || (reinterpret_cast<uintptr_t>(a.dTrain) & 3)) {
        set_error("L2 matcher needs 4-byte aligned descriptor blobs");
        return -1;
    }
    const int dim = a.descBytes / 4;
    // both operand tiles keep every dimension in shared memory (the reference order strides through them): pick the
    // largest tile edge that fits
    const size_t smemCap = 200 * 1024;
    const int TL = 2 * (size_t)dim * 65 * 4 <= smemCap ? 64 : 2 * (size_t)dim * 33 * 4 <= smemCap ? 32 : 16;
    if (2 * (size_t)dim * 17 * 4 > smemCap) { set_error("L2 matcher supports descriptors of up to 1500 floats"); return -1; }
    const int64_t totalQ = a.hQOff[a.nPairs];
    if (a.hQOff[0] != 0 || a.hTOff[0] != 0) { set_error("offset arrays must start at 0"); return -1; }
    if (totalQ == 0) return 0;
    const size_t n1 = (size_t)a.nPairs + 1;
    const size_t ctlBytes = n1 * (8 + 8 + 4);
    if (ctx->pin[0].reserve(ctlBytes) || ctx->ws[0].reserve(ctlBytes)) return -1;
    MCV_CUDA_TRY(cudaEventSynchronize(ctx->ev[0]));
    int64_t* hQ = ctx->pin[0].as<int64_t>();
    int64_t* hT = hQ + n1;
    int* hTile = reinterpret_cast<int*>(hT + n1);
    int64_t tiles = 0, maxT = 0;
    for (int p = 0; p < a.nPairs; ++p) {
        const int64_t nq = a.hQOff[p + 1] - a.hQOff[p], nt = a.hTOff[p + 1] - a.hTOff[p];
        if (nq < 0 || nt < 0 || nt > 0x7FFFFFFFll) { set_error("bad offsets at pair %d", p); return -1; }
        hQ[p] = a.hQOff[p]; hT[p] = a.hTOff[p]; hTile[p] = (int)tiles;
        tiles += (nq + TL - 1) / TL;
        maxT = std::max(maxT, nt);
    }
    hQ[a.nPairs] = a.hQOff[a.nPairs]; hT[a.nPairs] = a.hTOff[a.nPairs]; hTile[a.nPairs] = (int)tiles;
    if (tiles > 0x7FFFFFFFll) { set_error("too many query tiles"); return -1; }
    MCV_CUDA_TRY(cudaMemcpyAsync(ctx->ws[0].p, hQ, ctlBytes, cudaMemcpyHostToDevice, stream));
    MCV_CUDA_TRY(cudaEventRecord(ctx->ev[0], stream));
    const int64_t* dQOff = ctx->ws[0].as<int64_t>();
    const int64_t* dTOff = dQOff + n1;
    const int* dTile = reinterpret_cast<const int*>(dTOff + n1);

    int64_t wantSplits = 1;
    const int64_t fill = 4ll * ctx->sm_count;
    if (tiles < fill) wantSplits = (fill + tiles - 1) / tiles;
    wantSplits = std::min<int64_t>(wantSplits, std::max<int64_t>(1, (maxT + TL - 1) / TL));
    int64_t splitRows = (std::max<int64_t>(maxT, 1) + wantSplits - 1) / wantSplits;
    splitRows = (splitRows + TL - 1) / TL * TL;
    const int nSplits = (int)std::max<int64_t>(1, (maxT + splitRows - 1) / splitRows);
    if (nSplits > 65535) { set_error("train set too large for one launch"); return -1; }
    if (ctx->ws[3].reserve((size_t)nSplits * totalQ * 16)) return -1;
    uint64_t* dKeys = ctx->ws[3].as<uint64_t>();
    dim3 grid((unsigned)tiles, (unsigned)nSplits);
    {
        const float* fq = reinterpret_cast<const float*>(a.dQuery);
        const float* ft = reinterpret_cast<const float*>(a.dTrain);
        const size_t smem = 2 * (size_t)dim * (TL + 1) * sizeof(float);
#define MCV_L2(T)                                                                                              \
        do {                                                                                                   \
            auto kern = l2_knn2_kernel<T>;                                                                     \
            if (smem > 48 * 1024)                                                                              \
                MCV_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
            kern<<<grid, 256, smem, stream>>>(fq, ft, dQOff, dTOff, dTile, a.nPairs, dim, (int)splitRows, totalQ, dKeys); \
        } while (0)
        if (TL == 64) MCV_L2(64); else if (TL == 32) MCV_L2(32); else MCV_L2(16);
#undef MCV_L2
    }
    MCV_LAUNCH_CHECK();
    l2_finalize_kernel<<<(unsigned)((totalQ + 255) / 256), 256, 0, stream>>>(dKeys, nSplits, totalQ, a.ratio,
                                                                            a.dTrainIdx, a.dDist, a.dRatioPass,
                                                                            a.dPassCount);
    MCV_LAUNCH_CHECK();
    return 0;
}

}  // namespace mcv
