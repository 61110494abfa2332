// hamming_tc.cu -- Hamming kNN (k = 2) on the 5th-generation tensor cores (tcgen05, sm_100a).
//
// Formulation: bits are expanded to int8 (one pass, HBM-bound): query bits to +-1, train bits to
// +-S (S = 64, or 32 for 512-bit rows), so for K-bit descriptors
//     a . b = S * (K - 2 * hamming(a, b))     (exact in int32, and it fits in int16)
// and the distance matrix of a pair is a K-major int8 GEMM  D = Q * T^T  issued as
// tcgen05.mma.kind::i8 (M = 128, N = 128, K = 32 per instruction), operands staged in shared
// memory by TMA (SWIZZLE_128B), accumulators in TMEM (four 128-column buffers = all 512 columns:
// 2 query tiles x 2 stages, or 1 tile x 4 stages for 512-bit rows).
//
// Epilogue = the binding resource (the ALU pipe issues 64 min/max per clock per SM, while
// tcgen05.ld delivers > 1 KB/clk/SM, tools/tmem_bench.cu), so it is built to spend ~1.5 ALU
// lane-operations per distance and never diverge:
//   * tcgen05.ld.32x32b.x32.pack::16b returns the low 16 bits of two adjacent accumulator
//     columns per register (the values fit in int16), so every DPX instruction
//     (VIADDMNMX.S16x2 / VIMNMX3.S16x2) works on two distances;
//   * a . b is a multiple of 2S, so adding (CH - 1 - column) < 2S to it makes a 16-bit key that
//     carries its own column: max() of keys = smallest distance, lowest column on ties.  The add
//     is fused into the min/max (VIADDMNMX), and the exact top-2 of a chunk of CH columns is the
//     branch-free recurrence  t = min(b1, x); b1 = max(b1, x); b2 = max(b2, t);
//   * once per chunk the four surviving candidates are widened to 32-bit keys
//     (value, 0xFFFF - chunk number, column) and merged into the row's running top-2 with six
//     more min/max -- again no branch, so a warp never serialises on the rows that improve.
//
// Warp roles (64 + (12 + 4) * 32 = 576 threads): warp 0 = TMA producer, warp 1 = TMEM allocator + MMA issuer,
// warps 2..13 = epilogue (three groups of four warps, one group per 128-column accumulator buffer; TMEM lane quarter =
// warp % 4), warps 14..17 = stagers that expand the NEXT work item's packed query rows into TMEM (columns 0..127)
// while the current item is swept.  TMEM: query tiles in the first 128 columns, NBUF = 3 accumulator buffers in the
// remaining 384.  One CTA is persistent (one per SM) and walks (query tile, train split) work items; per item it
// sweeps the train rows of its split for 256 queries (two M = 128 tiles share every train stage, halving L2 -> SM
// traffic).  Results leave the kernel as packed keys
// (dist << 32) | trainIdx and go through the same finalize kernel as the integer-pipe engine
// (hamming_knn.cu), so tie-breaking is identical.
#include "knn.cuh"
#include "tc_ptx.cuh"

#include <cuda.h>
#include <algorithm>
#include <vector>
#include <mutex>

namespace mcv {
namespace {

constexpr int NBUF = 3;            // 128-column accumulator buffers in TMEM (columns 128..511)
constexpr int EPI_WARPS = 4 * NBUF; // one group of four warps (one per TMEM lane quarter) per buffer
constexpr int STAGER_WARPS = 4;     // warps that write the query tiles into TMEM (one per lane quarter)
constexpr int TC_THREADS = 64 + (EPI_WARPS + STAGER_WARPS) * 32;
using namespace tc;

// ---------------------------------------------------------------------------------------------
// bit expansion: packed descriptor bits -> int8 +-scale rows of KC*128 bytes (zero bits ->
// -scale; padding bits are "zero" on both sides, so they never contribute to the distance)
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
expand_bits_kernel(const uint8_t* __restrict__ src, int srcBytes, int64_t nRows, int rowBytesOut,
                   uint32_t scale, uint4* __restrict__ dst) {
    // one thread -> 16 output bytes (16 bits = 2 source bytes)
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int per = rowBytesOut / 16;
    if (i >= nRows * per) return;
    const int64_t row = i / per;
    const int g = (int)(i - row * per);
    const uint8_t* s = src + row * srcBytes;
    const int b0 = 2 * g, b1 = 2 * g + 1;
    const uint32_t bits = (b0 < srcBytes ? (uint32_t)s[b0] : 0u) | ((b1 < srcBytes ? (uint32_t)s[b1] : 0u) << 8);
    const uint32_t neg = (256u - scale) * 0x01010101u;          // -scale in every byte
    const uint32_t flip = (scale ^ (256u - scale)) & 0xFFu;     // -scale ^ flip = +scale
    uint32_t w[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const uint32_t nib = (bits >> (4 * q)) & 0xFu;
        // spread 4 bits to 4 bytes of 0/1, then 1 -> +scale, 0 -> -scale
        const uint32_t sp = (nib * 0x00204081u) & 0x01010101u;
        w[q] = neg ^ (sp * flip);
    }
    dst[i] = make_uint4(w[0], w[1], w[2], w[3]);
}

// ---------------------------------------------------------------------------------------------
// main kernel
// ---------------------------------------------------------------------------------------------
#ifndef MCV_TC_WARP_ARRIVE
#define MCV_TC_WARP_ARRIVE 0
#endif
#ifndef MCV_TC_ABLATE
#define MCV_TC_ABLATE 0      // profiling builds only: 1 = epilogue without the top-2 math, 2 = also without tcgen05.ld, 3 = also without TMA after the pipeline fill
#endif
#ifdef MCV_TC_PROF
__device__ unsigned long long g_tc_prof[16];

#define PROF_T0() const long long _pt = clock64()
#if MCV_TC_PROF >= 2     // fine-grained wait / work counters (intrusive)
#define PROF_ADD(i) atomicAdd(&g_tc_prof[i], (unsigned long long)(clock64() - _pt))
#else
#define PROF_ADD(i) (void)_pt
#endif
#else
#define PROF_T0()
#define PROF_ADD(i)
#endif

struct alignas(16) TileDesc {   // work item of one CTA column: QROWS query rows of one pair against its train rows
    int64_t qTile0;             // first query row (global row index in the packed batch)
    int64_t qEnd;               // end of this pair's query rows
    int64_t tRow0;              // first train row of the pair
    int32_t tCount;             // train rows of the pair
    int32_t pad;
};
struct alignas(16) PairDesc {   // host-built, one per pair (pinned, read by build_tiles_kernel over PCIe)
    int64_t q0, q1;             // query row range
    int64_t t0;                 // first train row
    int32_t tCount;             // train rows
    int32_t tile0;              // first tile index of the pair
};

// one block per pair: writes the pair's tile descriptors into HBM
__global__ void __launch_bounds__(128)
build_tiles_kernel(const PairDesc* __restrict__ hostPairs, int qRows, TileDesc* __restrict__ tiles) {
    __shared__ PairDesc pd;
    if (threadIdx.x < 2) reinterpret_cast<uint4*>(&pd)[threadIdx.x] = reinterpret_cast<const uint4*>(hostPairs + blockIdx.x)[threadIdx.x];
    __syncthreads();
    const int n = (int)((pd.q1 - pd.q0 + qRows - 1) / qRows);
    for (int i = threadIdx.x; i < n; i += blockDim.x)
        tiles[pd.tile0 + i] = TileDesc{pd.q0 + (int64_t)i * qRows, pd.q1, pd.t0, pd.tCount, 0};
}

struct TcSmem {      // carved from dynamic shared memory (1024-byte aligned)
    static constexpr int BAR_BYTES = 256;
};

// Per-configuration key layout.  A chunk is CH accumulator columns whose 16-bit keys are
//     key16 = S * (a . b) + (CH - 1 - column)         S * (a . b) is a multiple of 2S >= CH
// i.e. key16 = (m << IDXB) | (CH - 1 - column) with m = (a . b) / 2 = K/2 - hamming.
template <int KC> struct KeyCfg {
    static constexpr int SCALE = KC == 4 ? 32 : 64;      // train-side int8 magnitude
    static constexpr int IDXB = KC == 4 ? 6 : 7;         // column bits inside a key
    static constexpr int CH = 1 << IDXB;                 // columns per chunk (64 or 128)
    static_assert(SCALE * KC * 128 + CH - 1 <= 32767, "keys must fit in int16");
    static_assert(2 * SCALE >= CH, "column bits must fit under the value step");
};
constexpr uint32_t PACK_NONE = 0x80008000u;              // two int16 minima: "no candidate"
constexpr bool PACK_LOW_IS_EVEN = true;                  // .pack::16b: even column in the low half

// packed index constants for register r of a chunk whose first column is `base`
template <int CH> __device__ __forceinline__ constexpr uint32_t idx_const(int base, int r) {
    const uint32_t e = (uint32_t)(CH - 1 - (base + 2 * r)), o = (uint32_t)(CH - 1 - (base + 2 * r + 1));
    return PACK_LOW_IS_EVEN ? ((o << 16) | e) : ((e << 16) | o);
}

// exact top-2 (per 16-bit lane) of b1 >= b2 and two more registers x0, x1 (+ their constants)
__device__ __forceinline__ void top2_step(uint32_t x0, uint32_t c0, uint32_t x1, uint32_t c1, uint32_t& b1, uint32_t& b2) {
    const uint32_t t0 = __viaddmin_s16x2(x0, c0, b1);
    b1 = __viaddmax_s16x2(x0, c0, b1);
    const uint32_t t1 = __viaddmin_s16x2(x1, c1, b1);
    b1 = __viaddmax_s16x2(x1, c1, b1);
    b2 = __vimax3_s16x2(b2, t0, t1);
}
// (a1 >= a2) U (b1 >= b2) -> top-2, per 16-bit lane
__device__ __forceinline__ void top2_merge16(uint32_t& a1, uint32_t& a2, uint32_t b1, uint32_t b2) {
    const uint32_t lo = __vmins2(a1, b1);
    a1 = __vmaxs2(a1, b1);
    a2 = __vimax3_s16x2(lo, a2, b2);
}

// Widen the two 16-bit keys of `b` to 32-bit keys (m, ctr, column) and return them.
template <int IDXB> __device__ __forceinline__ void widen(uint32_t b, uint32_t ctrShifted, int& lo, int& hi) {
    constexpr uint32_t MASKM = ~((1u << (16 + IDXB)) - 1u);
    constexpr uint32_t MASKI = (1u << IDXB) - 1u;
    hi = (int)(((b & MASKM) | ctrShifted) | ((b >> 16) & MASKI));
    lo = (int)((((b << 16) & MASKM) | ctrShifted) | (b & MASKI));
}

// One chunk of CH columns held in NREG = CH/2 packed registers -> merged into (g1 >= g2).
template <int IDXB, int NREG>
__device__ __forceinline__ void chunk_top2(const uint32_t* v, uint32_t chunkNo, int& g1, int& g2) {
    constexpr int CH = 1 << IDXB;
    static_assert(NREG * 2 == CH && NREG % 8 == 0, "chunk shape");
    constexpr int NCHAIN = 4, PER = NREG / NCHAIN;       // four independent dependency chains
    uint32_t b1[NCHAIN], b2[NCHAIN];
#pragma unroll
    for (int c = 0; c < NCHAIN; ++c) { b1[c] = PACK_NONE; b2[c] = PACK_NONE; }
#pragma unroll
    for (int i = 0; i < PER; i += 2)
#pragma unroll
        for (int c = 0; c < NCHAIN; ++c) {
            const int r = c * PER + i;
            top2_step(v[r], idx_const<CH>(0, r), v[r + 1], idx_const<CH>(0, r + 1), b1[c], b2[c]);
        }
    top2_merge16(b1[0], b2[0], b1[1], b2[1]);
    top2_merge16(b1[2], b2[2], b1[3], b2[3]);
    top2_merge16(b1[0], b2[0], b1[2], b2[2]);
    // four candidates (top-2 of the even columns, top-2 of the odd columns) -> 32-bit keys
    const uint32_t ctrShifted = (0xFFFFu - chunkNo) << IDXB;
    int hA, hB, lA, lB;
    widen<IDXB>(b1[0], ctrShifted, hA, hB);
    widen<IDXB>(b2[0], ctrShifted, lA, lB);
    const int H = max(hA, hB);
    const int S = __vimax3_s32(min(hA, hB), lA, lB);
    g2 = __vimax3_s32(g2, min(g1, H), S);
    g1 = max(g1, H);
}

// columns >= nv of a 128-column row (64 packed registers) become "no candidate"
__device__ __forceinline__ void mask_tail(uint32_t (&v)[64], int nv) {
#pragma unroll
    for (int r = 0; r < 64; ++r) {
        const int ce = PACK_LOW_IS_EVEN ? 2 * r : 2 * r + 1;      // column in the low half
        const int co = PACK_LOW_IS_EVEN ? 2 * r + 1 : 2 * r;      // column in the high half
        if (ce >= nv) v[r] = (v[r] & 0xFFFF0000u) | 0x8000u;
        if (co >= nv) v[r] = (v[r] & 0x0000FFFFu) | 0x80000000u;
    }
}

// 128 descriptor bits (16 bytes, zero-padded past descBytes) of one query row, as four words
__device__ __forceinline__ void load_query_chunk(const uint8_t* __restrict__ rowPtr, int descBytes, int kc,
                                                 bool rowValid, uint32_t (&src)[4]) {
    src[0] = src[1] = src[2] = src[3] = 0u;
    if (!rowValid) return;
    if (((descBytes | (int)(reinterpret_cast<uintptr_t>(rowPtr) & 3)) & 3) == 0) {
#pragma unroll
        for (int w = 0; w < 4; ++w)
            if (kc * 16 + w * 4 < descBytes) src[w] = __ldg(reinterpret_cast<const uint32_t*>(rowPtr) + kc * 4 + w);
    } else {
#pragma unroll
        for (int i = 0; i < 16; ++i)
            if (kc * 16 + i < descBytes) src[i >> 2] |= (uint32_t)__ldg(rowPtr + kc * 16 + i) << (8 * (i & 3));
    }
}
// ... expanded to 32 words of int8 +-1 and written to 32 TMEM columns of this thread's lane: the A
// operand of the MMAs (rows past the end of the batch are all-zero)
__device__ __forceinline__ void store_query_chunk(const uint32_t (&src)[4], bool rowValid, uint32_t taddr) {
    uint32_t v[32];
#pragma unroll
    for (int w = 0; w < 4; ++w)
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            const uint32_t nib = (src[w] >> (4 * q)) & 0xFu;
            const uint32_t sp = (nib * 0x00204081u) & 0x01010101u;      // 4 bits -> 4 bytes of 0/1
            v[w * 8 + q] = rowValid ? (0xFFFFFFFFu ^ (sp * 0xFEu)) : 0u;   // 1 -> +1, 0 -> -1
        }
    tc_st32(taddr, v);
}

// One work item = one TileDesc (QROWS query rows of a pair) x one train split.  The kernel is
// persistent: CTA c walks items c, c + gridDim.x, ... and every role keeps its own cursor over the same
// item sequence, so the TMA ring, the accumulator buffers and the barrier phases run on across items.
struct WorkItem {
    int64_t qTile0, qEnd, tRow0;
    int t0, nT, nStages, split;
};
__device__ __forceinline__ WorkItem load_item(const TileDesc* __restrict__ tileDescs, int nTiles, int splitRows, int w) {
    const int split = w / nTiles, tile = w - split * nTiles;
    const int4 td0 = __ldg(reinterpret_cast<const int4*>(tileDescs + tile));
    const int4 td1 = __ldg(reinterpret_cast<const int4*>(tileDescs + tile) + 1);
    WorkItem it;
    it.qTile0 = ((int64_t)td0.y << 32) | (uint32_t)td0.x;
    it.qEnd = ((int64_t)td0.w << 32) | (uint32_t)td0.z;
    it.tRow0 = ((int64_t)td1.y << 32) | (uint32_t)td1.x;
    const int tCount = td1.z;
    it.t0 = min(tCount, split * splitRows);
    it.nT = min(tCount, it.t0 + splitRows) - it.t0;
    it.nStages = (it.nT + TILE_N - 1) / TILE_N;
    it.split = split;
    return it;
}

template <int KC, int ATILES, int STAGES>
__global__ void __launch_bounds__(TC_THREADS, 1)
hamming_tc_kernel(const uint8_t* __restrict__ query, int descBytes, const __grid_constant__ CUtensorMap mapT,
                  const TileDesc* __restrict__ tileDescs, int nTiles, int nWork, int splitRows, int64_t tRowBase,
                  int64_t totalQ, uint64_t* __restrict__ splitKeys) {
    constexpr int B_BYTES = KC * CHUNK_BYTES;
    constexpr int KBITS = KC * 128;
    constexpr int A_COLS = KC * 32;                      // TMEM columns of one query tile (4 int8 per column)
    constexpr int ACC_COL0 = 128;                        // accumulator buffers start here (ATILES * A_COLS <= 128)
    static_assert(ATILES * A_COLS <= ACC_COL0, "query tiles must fit under the accumulators");
    constexpr int IDXB = KeyCfg<KC>::IDXB;
    constexpr int CH = KeyCfg<KC>::CH;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* sB = smem;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * B_BYTES);
    uint64_t* a_full = bars;                 // 1: query tiles of the current item are in TMEM
    uint64_t* a_free = bars + 1;             // 1: every MMA of the current item has retired
    uint64_t* b_full = bars + 2;             // STAGES
    uint64_t* b_empty = b_full + STAGES;     // STAGES
    uint64_t* acc_full = b_empty + STAGES;   // NBUF
    uint64_t* acc_empty = acc_full + NBUF;   // NBUF
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + NBUF);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#ifdef MCV_TC_PROF
    const long long _cta0 = clock64();
#endif

    if (threadIdx.x == 0) {
        mbar_init(a_full, STAGER_WARPS);             // one arrival per staging warp
        mbar_init(a_free, 1);
        for (int s = 0; s < STAGES; ++s) { mbar_init(b_full + s, 1); mbar_init(b_empty + s, 1); }
        for (int s = 0; s < NBUF; ++s) { mbar_init(acc_full + s, 1); mbar_init(acc_empty + s, 128); }
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
        // ===== TMA producer (train stages only; the query tiles live in TMEM) =====
        // one elected lane; the stage ring runs on across work items, so the first stages of the next
        // item are in flight while the current one drains
        if (elect_one()) {
            int s = 0, sphase = 0;
            for (int w = blockIdx.x; w < nWork; w += gridDim.x) {
                const WorkItem it = load_item(tileDescs, nTiles, splitRows, w);
                for (int j = 0; j < it.nStages; ++j) {
                    mbar_wait(b_empty + s, sphase ^ 1);
                    mbar_expect_tx(b_full + s, B_BYTES);
#pragma unroll
                    for (int kc = 0; kc < KC; ++kc)
                        tma_load_2d(sB + s * B_BYTES + kc * CHUNK_BYTES, &mapT, kc * 128,
                                    (int)(it.tRow0 - tRowBase + it.t0 + (int64_t)j * TILE_N), b_full + s);
                    if (++s == STAGES) { s = 0; sphase ^= 1; }
                }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        // ===== MMA issuer =====
        // The tensor pipe queues only ~2-3 MMAs (tools/mma_bench.cu: an issuer that is away for more
        // than ~150 cycles starves it), so the barrier waits of unit n+1 are taken in the middle of
        // unit n, while its last MMAs are still queued.
        // instruction descriptor: D = s32, A = B = signed int8, K-major both, N = 128, M = 128
        constexpr uint32_t idesc = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(TILE_N >> 3) << 17) | ((uint32_t)(TILE_M >> 4) << 24);
        constexpr int NMMA = KC * 4;                 // MMAs per unit
        constexpr int TAIL = 2;                      // MMAs issued after the look-ahead waits
        if (elect_one()) {
            int buf = 0, use = 0;                    // global unit counter: buffer = g % NBUF, use = g / NBUF
            int s = 0, sphase = 0;                   // smem stage ring
            int item = 0;
            for (int w = blockIdx.x; w < nWork; w += gridDim.x, ++item) {
                const WorkItem it = load_item(tileDescs, nTiles, splitRows, w);
                const int nUnits = it.nStages * ATILES;
                mbar_wait(a_full, item & 1);
                if (nUnits > 0) {
                    mbar_wait(b_full + s, sphase);
                    mbar_wait(acc_empty + buf, (use & 1) ^ 1);
                }
                tc_fence_after();
                int t = 0;
                for (int n = 0; n < nUnits; ++n) {
                    const uint32_t bBase = smem_u32(sB + s * B_BYTES);
                    const uint32_t acol = tmem_base + (uint32_t)(t * A_COLS);
                    const uint32_t dcol = tmem_base + (uint32_t)(ACC_COL0 + buf * TILE_N);
#pragma unroll
                    for (int i = 0; i < NMMA - TAIL; ++i) {
                        const uint64_t db = umma_desc_sw128(bBase + (i >> 2) * CHUNK_BYTES + (i & 3) * 32);
                        tc_mma_i8_ts(dcol, acol + (uint32_t)(i * 8), db, idesc, i ? 1u : 0u);
                    }
                    // look ahead: everything unit n+1 of this item needs
                    int nbuf = buf + 1, nuse = use;
                    if (nbuf == NBUF) { nbuf = 0; ++nuse; }
                    const bool lastTile = t == ATILES - 1;
                    int ns = s, nsphase = sphase;
                    if (lastTile) { if (++ns == STAGES) { ns = 0; nsphase ^= 1; } }
                    if (n + 1 < nUnits) {
                        if (lastTile) mbar_wait(b_full + ns, nsphase);
                        mbar_wait(acc_empty + nbuf, (nuse & 1) ^ 1);
                        tc_fence_after();
                    }
#pragma unroll
                    for (int i = NMMA - TAIL; i < NMMA; ++i) {
                        const uint64_t db = umma_desc_sw128(bBase + (i >> 2) * CHUNK_BYTES + (i & 3) * 32);
                        tc_mma_i8_ts(dcol, acol + (uint32_t)(i * 8), db, idesc, 1u);
                    }
                    tc_commit(acc_full + buf);           // this unit's accumulators are complete
                    if (lastTile) { tc_commit(b_empty + s); t = 0; } else ++t;   // frees the train stage
                    buf = nbuf; use = nuse; s = ns; sphase = nsphase;
                }
                tc_commit(a_free);                       // the query tiles may be replaced once these MMAs retire
            }
        }
        __syncwarp();
    } else if (warp >= 2 + EPI_WARPS) {
        // ===== query stagers: warps 14..17, TMEM lane quarter = warp % 4 =====
        // expand this lane's query rows (one per tile) from the packed blob into the A region of TMEM;
        // the rows of the next item are fetched while the current item computes
        const int quarter = warp & 3;
        const int row = quarter * 32 + lane;
        const uint32_t tlane = tmem_base + ((uint32_t)(quarter * 32) << 16);
        int item = 0;
        for (int w = blockIdx.x; w < nWork; w += gridDim.x, ++item) {
            const WorkItem it = load_item(tileDescs, nTiles, splitRows, w);
            uint32_t src[ATILES][KC][4];
            bool valid[ATILES];
#pragma unroll
            for (int t = 0; t < ATILES; ++t) {
                const int64_t qRow = it.qTile0 + t * TILE_M + row;
                valid[t] = qRow < totalQ;
                const uint8_t* rowPtr = query + (valid[t] ? qRow : 0) * (int64_t)descBytes;
#pragma unroll
                for (int kc = 0; kc < KC; ++kc) load_query_chunk(rowPtr, descBytes, kc, valid[t], src[t][kc]);
            }
            if (item > 0) {
                mbar_wait(a_free, (item - 1) & 1);       // previous item's MMAs have all read their A
                tc_fence_after();
            }
#pragma unroll
            for (int t = 0; t < ATILES; ++t)
#pragma unroll
                for (int kc = 0; kc < KC; ++kc)
                    store_query_chunk(src[t][kc], valid[t], tlane + (uint32_t)(t * A_COLS + kc * 32));
            tc_wait_st();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(a_full);
        }
    } else {
        // ===== epilogue: warps 2..13, TMEM lane quarter = warp % 4 =====
        const int quarter = warp & 3;
        const int grp = (warp - 2) >> 2;                 // 0 .. 2 = the accumulator buffer this warp owns
        const int row = quarter * 32 + lane;             // accumulator row within an M tile
        const uint32_t tcol = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(ACC_COL0 + grp * TILE_N);
        int use = 0;                                     // uses of this group's buffer so far
        int gu = grp;                                    // global index of this group's next unit
        int unitBase = 0;                                // global index of the current item's first unit
        for (int w = blockIdx.x; w < nWork; w += gridDim.x) {
            const WorkItem it = load_item(tileDescs, nTiles, splitRows, w);
            const int nUnits = it.nStages * ATILES;
            // running top-2 of this item as (m, ctr, column) keys, per query tile
            int k1[ATILES], k2[ATILES];
#pragma unroll
            for (int t = 0; t < ATILES; ++t) { k1[t] = INT_MIN; k2[t] = INT_MIN; }
            for (; gu < unitBase + nUnits; gu += NBUF) {
                const int n = gu - unitBase;
                const int j = ATILES == 2 ? (n >> 1) : n;
                const int t = ATILES == 2 ? (n & 1) : 0;
                { PROF_T0(); mbar_wait(acc_full + grp, use & 1); if (lane == 0 && warp == 2) PROF_ADD(2); }
                ++use;
                tc_fence_after();
                PROF_T0();
                uint32_t v[64];
                tc_ld32_pack16(tcol, v);                     // columns 0..63 of the buffer
                tc_ld32_pack16(tcol + 64, v + 32);           // columns 64..127
                tc_wait_ld();
                tc_fence_before();
                mbar_arrive(acc_empty + grp);                // the buffer can be overwritten now
                if (lane == 0 && warp == 2) PROF_ADD(3);
                const int valid = it.nT - j * TILE_N;        // train rows of this stage that exist
                if (valid < TILE_N) mask_tail(v, valid);
                int g1 = k1[0], g2 = k2[0];
                if (ATILES == 2 && t) { g1 = k1[ATILES - 1]; g2 = k2[ATILES - 1]; }
                if constexpr (CH == 128) {
                    chunk_top2<IDXB, 64>(v, (uint32_t)j, g1, g2);
                } else {
                    chunk_top2<IDXB, 32>(v, (uint32_t)(2 * j), g1, g2);
                    chunk_top2<IDXB, 32>(v + 32, (uint32_t)(2 * j + 1), g1, g2);
                }
                if (ATILES == 2 && t) { k1[ATILES - 1] = g1; k2[ATILES - 1] = g2; } else { k1[0] = g1; k2[0] = g2; }
                if (lane == 0 && warp == 2) PROF_ADD(4);
            }
            unitBase += nUnits;
            // emit packed keys (dist << 32) | pair-local train index; one slot per (split, group)
            auto decode = [&](int g) -> uint64_t {
                const int m = g >> (16 + IDXB);                        // K/2 - hamming
                const int dist = KBITS / 2 - m;
                if (dist > KBITS) return ~0ull;                        // sentinel / masked column
                const int chunkNo = 0xFFFF - ((g >> IDXB) & 0xFFFF);
                const int col = (CH - 1) - (g & (CH - 1));
                return ((uint64_t)(uint32_t)dist << 32) | (uint64_t)(uint32_t)(it.t0 + chunkNo * CH + col);
            };
            const int slot = it.split * NBUF + grp;
#pragma unroll
            for (int t = 0; t < ATILES; ++t) {
                const int64_t qRow = it.qTile0 + t * TILE_M + row;
                if (qRow < it.qEnd) {
                    uint64_t* o = splitKeys + ((int64_t)slot * totalQ + qRow) * 2;
                    reinterpret_cast<ulonglong2*>(o)[0] = make_ulonglong2(decode(k1[t]), decode(k2[t]));
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
#ifdef MCV_TC_PROF
    if (threadIdx.x == 0) { atomicAdd(&g_tc_prof[5], (unsigned long long)(clock64() - _cta0)); atomicAdd(&g_tc_prof[6], 1ull); }
#endif
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
    }
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
template <int KC, int ATILES, int STAGES>
int launch_cfg(Context* ctx, cudaStream_t stream, dim3 grid, const uint8_t* query, int descBytes, const CUtensorMap& mt,
               const TileDesc* tileDescs, int nTiles, int nWork, int splitRows, int64_t tRowBase, int64_t totalQ,
               uint64_t* splitKeys) {
    constexpr int SMEM = STAGES * KC * CHUNK_BYTES + TcSmem::BAR_BYTES + 1024;
    auto kern = hamming_tc_kernel<KC, ATILES, STAGES>;
    MCV_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
    kt_begin(ctx, stream);
    kern<<<grid, TC_THREADS, SMEM, stream>>>(query, descBytes, mt, tileDescs, nTiles, nWork, splitRows, tRowBase, totalQ, splitKeys);
    kt_end(ctx, stream);
    MCV_LAUNCH_CHECK();
    return 0;
}

}  // namespace

#ifdef MCV_TC_PROF
extern "C" __attribute__((visibility("default"))) void mcvDebugTcProf(unsigned long long* out, int reset) {
    cudaDeviceSynchronize();
    cudaMemcpyFromSymbol(out, g_tc_prof, sizeof(unsigned long long) * 16);
    if (reset) { unsigned long long z[16] = {}; cudaMemcpyToSymbol(g_tc_prof, z, sizeof(z)); }
}

#endif

bool hamming_tc_supported(int descBytes) { return descBytes >= 1 && descBytes <= 64; }

// Tensor-core engine.  Same contract as launch_hamming_knn2 (knn.cuh); uses ctx workspaces
// 3 (split keys), 25 (expanded train rows) and the double-buffered control arrays.
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

    // control array: one descriptor per query tile (a CTA finds its work with a single 32-byte load)
    int64_t tiles = 0, maxT = 0;
    for (int p = 0; p < a.nPairs; ++p) {
        const int64_t nq = a.hQOff[p + 1] - a.hQOff[p], nt = a.hTOff[p + 1] - a.hTOff[p];
        if (nq < 0 || nt < 0 || nt > 0x7FFFFFFFll) { set_error("bad offsets at pair %d", p); return -1; }
        tiles += (nq + QROWS - 1) / QROWS;
        maxT = std::max(maxT, nt);
    }
    if (tiles > 0x7FFFFFFFll || totalQ + 2 * QROWS > 0x7FFFFFFFll || totalT + 2 * TILE_N > 0x7FFFFFFFll) {
        set_error("batch too large for 32-bit TMA row coordinates");
        return -1;
    }
    if (tiles == 0) return 0;
    // per-pair control rows (q0, q1, t0, nT, first tile) go through a pinned buffer; a small kernel
    // reads them over PCIe and writes one TileDesc per query tile.  Not cudaMemcpyAsync: a small copy
    // queued on the compute stream would sit in the copy engine's in-order queue and hold back the
    // bulk descriptor copies of the following chunks (cvMatchKnn2Batch) until this stream catches up.
    const size_t pairBytes = (size_t)a.nPairs * sizeof(PairDesc);
    const int turn = ctx->ctlTurn ^= 1;
    if (ctx->ctlPin[turn].reserve(pairBytes) || ctx->ctlWs[turn].reserve((size_t)tiles * sizeof(TileDesc))) return -1;
    MCV_CUDA_TRY(cudaEventSynchronize(ctx->ctlEv[turn]));     // the pinned buffer's use two launches ago
    PairDesc* hPairs = ctx->ctlPin[turn].as<PairDesc>();
    {
        int64_t k = 0;
        for (int p = 0; p < a.nPairs; ++p) {
            hPairs[p] = PairDesc{a.hQOff[p], a.hQOff[p + 1], a.hTOff[p], (int)(a.hTOff[p + 1] - a.hTOff[p]), (int)k};
            k += (a.hQOff[p + 1] - a.hQOff[p] + QROWS - 1) / QROWS;
        }
    }
    build_tiles_kernel<<<a.nPairs, 128, 0, stream>>>(hPairs, QROWS, ctx->ctlWs[turn].as<TileDesc>());
    MCV_LAUNCH_CHECK();
    MCV_CUDA_TRY(cudaEventRecord(ctx->ctlEv[turn], stream));
    const TileDesc* dTiles = ctx->ctlWs[turn].as<TileDesc>();

    // train splits: fill the machine when there are few query tiles
    int64_t wantSplits = 1;
    const int64_t fill = 2ll * ctx->sm_count;
    if (tiles < fill) wantSplits = (fill + tiles - 1) / tiles;
    const int64_t maxStages = std::max<int64_t>(1, (maxT + TILE_N - 1) / TILE_N);
    wantSplits = std::min<int64_t>(wantSplits, std::max<int64_t>(1, maxStages / 4));
    // the 16-bit chunk counter inside the epilogue keys covers 2^22 train rows per split
    wantSplits = std::max<int64_t>(wantSplits, (maxT + (1ll << 22) - 1) >> 22);
    int64_t splitRows = (std::max<int64_t>(maxT, 1) + wantSplits - 1) / wantSplits;
    splitRows = (splitRows + TILE_N - 1) / TILE_N * TILE_N;
    const int nSplits = (int)std::max<int64_t>(1, (maxT + splitRows - 1) / splitRows);
    if (nSplits > 65535) { set_error("train set too large for one launch"); return -1; }
    const int nSlots = NBUF * nSplits;   // one key slot per epilogue warp group and split
    if (ctx->ws[3].reserve((size_t)nSlots * totalQ * 16)) return -1;
    uint64_t* dSplitKeys = ctx->ws[3].as<uint64_t>();

    // The batch is processed in chunks of whole pairs: expand the chunk's train bits to int8 +-scale
    // (the query rows are expanded inside the kernel, straight into TMEM), then one persistent launch
    // (one CTA per SM walking the chunk's (tile, split) work items c, c + grid, ...).  A chunk's
    // expanded rows (<= 96 MB) are still in the 126 MB L2 when the matcher streams them (measured:
    // 68.8 k instead of 73.7 k cycles per 64-stage work item), and chunk ends are chosen where the item
    // count is closest to a multiple of the SM count so that no launch ends with a thin last wave.
    const int64_t budgetRows = std::max<int64_t>((96ll << 20) / rowBytes, 1);
    const uint32_t trainScale = KC == 4 ? KeyCfg<4>::SCALE : KeyCfg<2>::SCALE;
    int64_t maxChunkRows = 0;
    std::vector<int> ends;                 // chunk end (pair index, exclusive)
    {
        const int sm = std::max(ctx->sm_count, 1);
        int p0 = 0;
        while (p0 < a.nPairs) {
            int64_t rows = 0, tl = 0, bestWaste = -1;
            int best = p0 + 1;
            for (int p = p0; p < a.nPairs; ++p) {
                const int64_t nt = a.hTOff[p + 1] - a.hTOff[p];
                if (p > p0 && rows + nt > budgetRows) break;
                rows += nt;
                tl += (a.hQOff[p + 1] - a.hQOff[p] + QROWS - 1) / QROWS;
                const int64_t items = tl * nSplits;
                const int64_t waves = (items + sm - 1) / sm;
                // idle CTA-slots in the last wave, relative to the chunk's work (per mille)
                const int64_t waste = items > 0 ? (waves * sm - items) * 1000 / (waves * sm) : 0;
                if (rows * 2 >= budgetRows || p + 1 == a.nPairs) {      // only ends in the upper half of the budget
                    if (bestWaste < 0 || waste <= bestWaste) { bestWaste = waste; best = p + 1; }
                } else if (bestWaste < 0) best = p + 1;
            }
            if (bestWaste < 0) {           // a single pair larger than half the budget, or the tail of the batch
                best = p0 + 1;
                int64_t r = a.hTOff[p0 + 1] - a.hTOff[p0];
                while (best < a.nPairs && r + (a.hTOff[best + 1] - a.hTOff[best]) <= budgetRows) { r += a.hTOff[best + 1] - a.hTOff[best]; ++best; }
            }
            ends.push_back(best);
            maxChunkRows = std::max<int64_t>(maxChunkRows, a.hTOff[best] - a.hTOff[p0]);
            p0 = best;
        }
    }
    if (ctx->ws[25].reserve((size_t)std::max<int64_t>(maxChunkRows, 1) * rowBytes + 1024)) return -1;
    int p0 = 0;
    int64_t tile0 = 0;
    for (size_t c = 0; c < ends.size(); ++c) {
        const int p1 = ends[c];
        const int64_t tA = a.hTOff[p0], tB = a.hTOff[p1];
        int64_t chunkTiles = 0;
        for (int p = p0; p < p1; ++p) chunkTiles += (a.hQOff[p + 1] - a.hQOff[p] + QROWS - 1) / QROWS;
        if (chunkTiles > 0) {
            const int64_t n16 = (tB - tA) * (rowBytes / 16);
            if (n16 > 0) {
                kt_begin(ctx, stream, KT_EXPAND);
                expand_bits_kernel<<<(unsigned)((n16 + 255) / 256), 256, 0, stream>>>(
                    a.dTrain + tA * a.descBytes, a.descBytes, tB - tA, rowBytes, trainScale, ctx->ws[25].as<uint4>());
                kt_end(ctx, stream);
                MCV_LAUNCH_CHECK();
            }
            CUtensorMap mt;
            if (make_map(&mt, ctx->ws[25].p, std::max<int64_t>(tB - tA, 1), rowBytes)) return -1;
            const int64_t nWork64 = chunkTiles * nSplits;
            if (nWork64 > 0x7FFFFFFFll) { set_error("too many work items for one launch"); return -1; }
            const int nWork = (int)nWork64;
            dim3 grid((unsigned)std::min<int64_t>(nWork64, ctx->sm_count));
            const TileDesc* td = dTiles + tile0;
            int rc;
            if (KC == 1) rc = launch_cfg<1, 2, 8>(ctx, stream, grid, a.dQuery, a.descBytes, mt, td, (int)chunkTiles, nWork, (int)splitRows, tA, totalQ, dSplitKeys);
            else if (KC == 2) rc = launch_cfg<2, 2, 6>(ctx, stream, grid, a.dQuery, a.descBytes, mt, td, (int)chunkTiles, nWork, (int)splitRows, tA, totalQ, dSplitKeys);
            else rc = launch_cfg<4, 1, 3>(ctx, stream, grid, a.dQuery, a.descBytes, mt, td, (int)chunkTiles, nWork, (int)splitRows, tA, totalQ, dSplitKeys);
            if (rc) return rc;
        }
        tile0 += chunkTiles;
        p0 = p1;
    }
    kt_begin(ctx, stream, KT_KNN_FINALIZE);
    const int rcf = launch_knn2_finalize(stream, dSplitKeys, nSlots, totalQ, a);
    kt_end(ctx, stream);
    return rcf;
}

}  // namespace mcv
