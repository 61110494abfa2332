// hamming_tc.cu -- Hamming kNN (k = 2) on the 5th-generation tensor cores (tcgen05, sm_100a).
//
// Formulation: bits are expanded to int8 +-1 (one pass, HBM-bound), so for K-bit descriptors
//     a . b = K - 2 * hamming(a, b)          (exact in int32)
// and the distance matrix of a pair is a K-major int8 GEMM  D = Q * T^T  issued as
// tcgen05.mma.kind::i8 (M = 128, N = 128, K = 32 per instruction), operands staged in shared
// memory by TMA (SWIZZLE_128B), accumulators in TMEM (2 query tiles x 2 stages x 128 columns
// = all 512 columns).  The epilogue warps read the accumulators with tcgen05.ld (one query row
// per thread, 32 train columns per load) and keep the running top-2 per row in registers:
// a group maximum filters each 8-column group against the current second best, so the exact
// (index-carrying, lowest-index-wins) update only runs for groups that can change the result.
//
// Warp roles (320 threads): warp 0 = TMA producer, warp 1 = TMEM allocator + MMA issuer,
// warps 2..9 = epilogue (TMEM lane quarter = warp % 4; the two warps of a quarter take one
// query tile each, or one half of the columns each when there is a single query tile, and
// software-pipeline tcgen05.ld of the next 32 columns under the top-2 work of the current 32).  One CTA sweeps the train rows of its
// split for 256 queries (two M = 128 tiles share every train stage, halving L2 -> SM traffic).
// Results leave the kernel as packed keys (dist << 32) | trainIdx and go through the same
// finalize kernel as the integer-pipe engine (hamming_knn.cu), so tie-breaking is identical.
#include "knn.cuh"
#include "tc_ptx.cuh"

#include <cuda.h>
#include <algorithm>
#include <mutex>

namespace mcv {
namespace {

constexpr int EPI_WARPS = 16;      // four epilogue warps per TMEM lane quarter (latency hiding)
constexpr int TC_THREADS = 64 + EPI_WARPS * 32;
using namespace tc;

// ---------------------------------------------------------------------------------------------
// bit expansion: packed descriptor bits -> int8 +-1 rows of KC*128 bytes (zero bits -> -1;
// padding bits are "zero" on both sides, so they never contribute to the distance)
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
expand_bits_kernel(const uint8_t* __restrict__ src, int srcBytes, int64_t nRows, int rowBytesOut,
                   uint4* __restrict__ dst) {
    // one thread -> 16 output bytes (16 bits = 2 source bytes)
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int per = rowBytesOut / 16;
    if (i >= nRows * per) return;
    const int64_t row = i / per;
    const int g = (int)(i - row * per);
    const uint8_t* s = src + row * srcBytes;
    const int b0 = 2 * g, b1 = 2 * g + 1;
    const uint32_t bits = (b0 < srcBytes ? (uint32_t)s[b0] : 0u) | ((b1 < srcBytes ? (uint32_t)s[b1] : 0u) << 8);
    uint32_t w[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const uint32_t nib = (bits >> (4 * q)) & 0xFu;
        // spread 4 bits to 4 bytes of 0/1, then 1 -> 0x01 (+1), 0 -> 0xFF (-1)
        const uint32_t sp = (nib * 0x00204081u) & 0x01010101u;
        w[q] = 0xFFFFFFFFu ^ (sp * 0xFEu);
    }
    dst[i] = make_uint4(w[0], w[1], w[2], w[3]);
}

// ---------------------------------------------------------------------------------------------
// main kernel
// ---------------------------------------------------------------------------------------------
struct TcSmem {      // carved from dynamic shared memory (1024-byte aligned)
    static constexpr int BAR_BYTES = 256;
};

__device__ __forceinline__ void top2_insert(int x, int col, int& b1, int& i1, int& b2, int& i2) {
    if (x > b2) {
        if (x > b1) { b2 = b1; i2 = i1; b1 = x; i1 = col; }
        else { b2 = x; i2 = col; }
    }
}

template <int KC, int ATILES, int STAGES>
__global__ void __launch_bounds__(TC_THREADS, 1)
hamming_tc_kernel(const __grid_constant__ CUtensorMap mapQ, const __grid_constant__ CUtensorMap mapT,
                  const int64_t* __restrict__ qOff, const int64_t* __restrict__ tOff,
                  const int* __restrict__ tileStart, int nPairs, int splitRows, int64_t totalQ,
                  uint64_t* __restrict__ splitKeys) {
    constexpr int QROWS = ATILES * TILE_M;
    constexpr int A_BYTES = ATILES * KC * CHUNK_BYTES;
    constexpr int B_BYTES = KC * CHUNK_BYTES;
    constexpr int KBITS = KC * 128;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* sA = smem;
    uint8_t* sB = smem + A_BYTES;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + A_BYTES + STAGES * B_BYTES);
    uint64_t* a_full = bars;                 // 1
    uint64_t* b_full = bars + 1;             // STAGES
    uint64_t* b_empty = b_full + STAGES;     // STAGES
    uint64_t* acc_full = b_empty + STAGES;   // 2
    uint64_t* acc_empty = acc_full + 2;      // 2
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    // locate pair / query tile / train split
    int lo = 0, hi = nPairs;
    const int tile = blockIdx.x;
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (tileStart[mid] <= tile) lo = mid; else hi = mid;
    }
    const int p = lo;
    const int64_t q0 = qOff[p], qEnd = qOff[p + 1];
    const int64_t tRow0 = tOff[p];
    const int tCount = (int)(tOff[p + 1] - tRow0);
    const int64_t qTile0 = q0 + (int64_t)(tile - tileStart[p]) * QROWS;
    const int split = blockIdx.y;
    const int t0 = min(tCount, split * splitRows);
    const int t1 = min(tCount, t0 + splitRows);
    const int nT = t1 - t0;
    const int nStagesTotal = (nT + TILE_N - 1) / TILE_N;

    if (threadIdx.x == 0) {
        mbar_init(a_full, 1);
        for (int s = 0; s < STAGES; ++s) { mbar_init(b_full + s, 1); mbar_init(b_empty + s, 1); }
        for (int s = 0; s < 2; ++s) { mbar_init(acc_full + s, 1); mbar_init(acc_empty + s, EPI_WARPS * 32); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ===== TMA producer =====
        if (lane == 0) {
            mbar_expect_tx(a_full, A_BYTES);
            for (int t = 0; t < ATILES; ++t)
                for (int kc = 0; kc < KC; ++kc)
                    tma_load_2d(sA + (t * KC + kc) * CHUNK_BYTES, &mapQ, kc * 128, (int)(qTile0 + t * TILE_M), a_full);
            for (int j = 0; j < nStagesTotal; ++j) {
                const int s = j % STAGES;
                mbar_wait(b_empty + s, ((j / STAGES) & 1) ^ 1);
                mbar_expect_tx(b_full + s, B_BYTES);
                for (int kc = 0; kc < KC; ++kc)
                    tma_load_2d(sB + s * B_BYTES + kc * CHUNK_BYTES, &mapT, kc * 128,
                                (int)(tRow0 + t0 + (int64_t)j * TILE_N), b_full + s);
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer =====
        // instruction descriptor: D = s32, A = B = signed int8, K-major both, N = 128, M = 128
        constexpr uint32_t idesc = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(TILE_N >> 3) << 17) | ((uint32_t)(TILE_M >> 4) << 24);
        if (lane == 0) {
            mbar_wait(a_full, 0);
            for (int j = 0; j < nStagesTotal; ++j) {
                const int s = j % STAGES;
                const int as = j & 1;
                mbar_wait(b_full + s, (j / STAGES) & 1);
                mbar_wait(acc_empty + as, ((j >> 1) & 1) ^ 1);
                tc_fence_after();
                const uint32_t bBase = smem_u32(sB + s * B_BYTES);
#pragma unroll
                for (int t = 0; t < ATILES; ++t) {
                    const uint32_t aBase = smem_u32(sA + t * KC * CHUNK_BYTES);
                    const uint32_t dcol = tmem_base + (uint32_t)(as * (ATILES * TILE_N) + t * TILE_N);
#pragma unroll
                    for (int kc = 0; kc < KC; ++kc)
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            const uint64_t da = umma_desc_sw128(aBase + kc * CHUNK_BYTES + k * 32);
                            const uint64_t db = umma_desc_sw128(bBase + kc * CHUNK_BYTES + k * 32);
                            tc_mma_i8(dcol, da, db, idesc, (kc | k) ? 1u : 0u);
                        }
                }
                tc_commit(b_empty + s);      // frees the train stage once these MMAs retire
                tc_commit(acc_full + as);    // accumulators of this stage are complete
            }
        }
    } else {
        // ===== epilogue: warps 2..9, TMEM lane quarter = warp % 4 =====
        const int quarter = warp & 3;
        const int grp = (warp - 2) >> 2;                 // 0 .. 3
        const int row = quarter * 32 + lane;             // accumulator row within an M tile
        // The four warps of a lane quarter split a stage's work; every warp reports its own
        // partial top-2 as a separate key slot (merged exactly by the finalize kernel).
        // ATILES == 2: query tile = grp & 1, column half = grp >> 1 (two 32-column chunks).
        // ATILES == 1: one 32-column chunk each.
        const int myTile = ATILES == 2 ? (grp & 1) : 0;
        constexpr int NCHUNK = ATILES == 2 ? 2 : 1;
        constexpr int SLOTS = ATILES == 2 ? 2 : 4;       // key slots per split
        const int colPart = ATILES == 2 ? (grp >> 1) : grp;
        const int chunk0 = colPart * NCHUNK;
        int b1 = INT_MIN, i1 = -1, b2 = INT_MIN, i2 = -1;
        const uint32_t tlane = tmem_base + ((uint32_t)(quarter * 32) << 16);

        auto process = [&](int (&v)[32], int nv, int cb) {
            if (nv < 32) {
#pragma unroll
                for (int i = 0; i < 32; ++i) if (i >= nv) v[i] = INT_MIN;
            }
            int g[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int m01 = max(v[q * 8 + 0], v[q * 8 + 1]), m23 = max(v[q * 8 + 2], v[q * 8 + 3]);
                const int m45 = max(v[q * 8 + 4], v[q * 8 + 5]), m67 = max(v[q * 8 + 6], v[q * 8 + 7]);
                g[q] = max(max(m01, m23), max(m45, m67));
            }
            const int gm = max(max(g[0], g[1]), max(g[2], g[3]));
            if (gm > b2) {
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    if (g[q] > b2) {
#pragma unroll
                        for (int i = 0; i < 8; ++i) top2_insert(v[q * 8 + i], cb + q * 8 + i, b1, i1, b2, i2);
                    }
                }
            }
        };

        for (int j = 0; j < nStagesTotal; ++j) {
            const int as = j & 1;
            mbar_wait(acc_full + as, (j >> 1) & 1);
            tc_fence_after();
            const int colBase = j * TILE_N;               // within the split
            const int valid = min(TILE_N, nT - colBase);  // train rows of this stage that exist
            const uint32_t tcol = tlane + (uint32_t)(as * (ATILES * TILE_N) + myTile * TILE_N + chunk0 * 32);
            const int cb0 = t0 + colBase + chunk0 * 32;
            int va[32], vb[32];
            // software pipeline: the load of chunk c+1 is in flight while chunk c is processed
            tc_ld32(tcol, va);
            tc_wait_ld();
            if (NCHUNK == 2) {
                tc_ld32(tcol + 32, vb);
                process(va, valid - chunk0 * 32, cb0);
                tc_wait_ld();
                tc_fence_before();
                mbar_arrive(acc_empty + as);                 // all of this warp's loads have landed
                process(vb, valid - (chunk0 + 1) * 32, cb0 + 32);
            } else {
                tc_fence_before();
                mbar_arrive(acc_empty + as);
                process(va, valid - chunk0 * 32, cb0);
            }
        }
        // emit packed keys (dist << 32) | pair-local train index
        {
            const int64_t qRow = qTile0 + myTile * TILE_M + row;
            const int slot = split * SLOTS + colPart;
            if (qRow < qEnd) {
                uint64_t k1 = ~0ull, k2 = ~0ull;
                if (i1 >= 0) k1 = ((uint64_t)(uint32_t)((KBITS - b1) >> 1) << 32) | (uint64_t)(uint32_t)i1;
                if (i2 >= 0) k2 = ((uint64_t)(uint32_t)((KBITS - b2) >> 1) << 32) | (uint64_t)(uint32_t)i2;
                uint64_t* o = splitKeys + ((int64_t)slot * totalQ + qRow) * 2;
                reinterpret_cast<ulonglong2*>(o)[0] = make_ulonglong2(k1, k2);
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
    }
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
template <int KC, int ATILES, int STAGES>
int launch_cfg(cudaStream_t stream, dim3 grid, const CUtensorMap& mq, const CUtensorMap& mt, const int64_t* qOff,
               const int64_t* tOff, const int* tileStart, int nPairs, int splitRows, int64_t totalQ,
               uint64_t* splitKeys) {
    constexpr int SMEM = ATILES * KC * CHUNK_BYTES + STAGES * KC * CHUNK_BYTES + TcSmem::BAR_BYTES + 1024;
    auto kern = hamming_tc_kernel<KC, ATILES, STAGES>;
    MCV_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
    kern<<<grid, TC_THREADS, SMEM, stream>>>(mq, mt, qOff, tOff, tileStart, nPairs, splitRows, totalQ, splitKeys);
    MCV_LAUNCH_CHECK();
    return 0;
}

}  // namespace

bool hamming_tc_supported(int descBytes) { return descBytes >= 1 && descBytes <= 64; }

// Tensor-core engine.  Same contract as launch_hamming_knn2 (knn.cuh); uses ctx workspaces
// 0 (control), 3 (split keys), 24 / 25 (expanded query / train rows).
int launch_hamming_knn2_tc(Context* ctx, cudaStream_t stream, const KnnArgs& a) {
    if (a.nPairs <= 0) return 0;
    if (!hamming_tc_supported(a.descBytes)) { set_error("tensor-core matcher supports 1..64-byte descriptors"); return -1; }
    const int KC = a.descBytes <= 16 ? 1 : a.descBytes <= 32 ? 2 : 4;
    const int ATILES = KC == 4 ? 1 : 2;
    const int QROWS = ATILES * TILE_M;
    const int rowBytes = KC * 128;
    const int64_t totalQ = a.hQOff[a.nPairs], totalT = a.hTOff[a.nPairs];
    if (a.hQOff[0] != 0 || a.hTOff[0] != 0) { set_error("offset arrays must start at 0"); return -1; }
    if (totalQ == 0) return 0;

    // control arrays (same staging protocol as the integer-pipe engine)
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
        tiles += (nq + QROWS - 1) / QROWS;
        maxT = std::max(maxT, nt);
    }
    hQ[a.nPairs] = totalQ; hT[a.nPairs] = totalT; hTile[a.nPairs] = (int)tiles;
    if (tiles > 0x7FFFFFFFll || totalQ + 2 * QROWS > 0x7FFFFFFFll || totalT + 2 * TILE_N > 0x7FFFFFFFll) {
        set_error("batch too large for 32-bit TMA row coordinates");
        return -1;
    }
    MCV_CUDA_TRY(cudaMemcpyAsync(ctx->ws[0].p, hQ, ctlBytes, cudaMemcpyHostToDevice, stream));
    MCV_CUDA_TRY(cudaEventRecord(ctx->ev[0], stream));
    const int64_t* dQOff = ctx->ws[0].as<int64_t>();
    const int64_t* dTOff = dQOff + n1;
    const int* dTile = reinterpret_cast<const int*>(dTOff + n1);

    // expand bits -> int8 +-1
    if (ctx->ws[24].reserve((size_t)totalQ * rowBytes + 1024) ||
        ctx->ws[25].reserve((size_t)std::max<int64_t>(totalT, 1) * rowBytes + 1024))
        return -1;
    {
        const int64_t nq = totalQ * (rowBytes / 16), nt = totalT * (rowBytes / 16);
        expand_bits_kernel<<<(unsigned)((nq + 255) / 256), 256, 0, stream>>>(a.dQuery, a.descBytes, totalQ, rowBytes,
                                                                             ctx->ws[24].as<uint4>());
        MCV_LAUNCH_CHECK();
        if (nt > 0) {
            expand_bits_kernel<<<(unsigned)((nt + 255) / 256), 256, 0, stream>>>(a.dTrain, a.descBytes, totalT, rowBytes,
                                                                                 ctx->ws[25].as<uint4>());
            MCV_LAUNCH_CHECK();
        }
    }
    CUtensorMap mq, mt;
    if (make_map(&mq, ctx->ws[24].p, totalQ, rowBytes) || make_map(&mt, ctx->ws[25].p, totalT, rowBytes)) return -1;

    // train splits: fill the machine when there are few query tiles
    int64_t wantSplits = 1;
    const int64_t fill = 2ll * ctx->sm_count;
    if (tiles < fill) wantSplits = (fill + tiles - 1) / tiles;
    const int64_t maxStages = std::max<int64_t>(1, (maxT + TILE_N - 1) / TILE_N);
    wantSplits = std::min<int64_t>(wantSplits, std::max<int64_t>(1, maxStages / 4));
    int64_t splitRows = (std::max<int64_t>(maxT, 1) + wantSplits - 1) / wantSplits;
    splitRows = (splitRows + TILE_N - 1) / TILE_N * TILE_N;
    const int nSplits = (int)std::max<int64_t>(1, (maxT + splitRows - 1) / splitRows);
    if (nSplits > 65535) { set_error("train set too large for one launch"); return -1; }
    const int nSlots = (ATILES == 2 ? 2 : 4) * nSplits;   // every epilogue warp of a lane quarter reports its own slot
    if (ctx->ws[3].reserve((size_t)nSlots * totalQ * 16)) return -1;
    uint64_t* dSplitKeys = ctx->ws[3].as<uint64_t>();

    dim3 grid((unsigned)tiles, (unsigned)nSplits);
    int rc;
    if (KC == 1) rc = launch_cfg<1, 2, 4>(stream, grid, mq, mt, dQOff, dTOff, dTile, a.nPairs, (int)splitRows, totalQ, dSplitKeys);
    else if (KC == 2) rc = launch_cfg<2, 2, 4>(stream, grid, mq, mt, dQOff, dTOff, dTile, a.nPairs, (int)splitRows, totalQ, dSplitKeys);
    else rc = launch_cfg<4, 1, 2>(stream, grid, mq, mt, dQOff, dTOff, dTile, a.nPairs, (int)splitRows, totalQ, dSplitKeys);
    if (rc) return rc;
    return launch_knn2_finalize(stream, dSplitKeys, nSlots, totalQ, a);
}

}  // namespace mcv
