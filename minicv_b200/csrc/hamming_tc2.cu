// hamming_tc2.cu -- tensor-core Hamming kNN, packed-accumulator variant ("two distances per
// TMEM word").
//
// Built when the first int8 engine was believed to be bound by TMEM read-out; this variant halves the
// read-out per distance.  tools/tmem_bench.cu later measured tcgen05.ld at 1-4 KB/clk/SM -- the real limit
// was the ALU pipe in the epilogue, which hamming_tc.cu now addresses with packed 16-bit keys.
// Operands are FP8: queries e4m3 +-1, train rows e5m2, where the rows of the FIRST half of a
// split carry +-1 ("lo" stream) and the rows of the SECOND half carry +-1024 ("hi" stream).
// Two chained MMAs (kind::f8f6f4, fp32 accumulate) put both dot products into ONE accumulator:
//        D = a . b_lo  +  1024 * (a . b_hi)        |a . b| <= 256, so D is an exact integer
// The epilogue reads each accumulator once and recovers both: lo = low 10 bits (sign-extended by
// a left shift of the magic-number-biased integer image), hi orders like D itself.  Each query
// row keeps an independent running top-2 for the lo stream and for the hi stream (each stream
// visits train indices in ascending order, so "first seen wins" keeps BFMatcher's lowest-index
// tie-break); both are emitted as packed keys and merged by the common finalize kernel.
//
// Per distance: tensor work unchanged (34.4 GOP / 256-bit pair), TMEM read 2 B instead of 4 B,
// ~1.5 ALU-pipe ops + 0.5 FMA-pipe ops in the filter instead of 1.
#include "knn.cuh"
#include "tc_ptx.cuh"

#include <algorithm>

// Built only with -DMCV_WITH_TC2 (make EXTRA=-DMCV_WITH_TC2): the engine was evaluated and rejected (41.5 ms per C2
// step); the default library answers "not supported" for engine 3 and falls back to the int8 engine.
#ifdef MCV_WITH_TC2
namespace mcv {
namespace {

using namespace tc;

constexpr int EPI_WARPS2 = 8;
constexpr int TC2_THREADS = 64 + EPI_WARPS2 * 32;
constexpr float MAGIC = 12582912.0f;          // 1.5 * 2^23: integer image of an exact-integer float
constexpr int MAGIC_BITS = 0x4B400000;

// ---------------------------------------------------------------------------------------------
// bit expansion to FP8.  Queries: e4m3 (+1 = 0x38, -1 = 0xB8).  Train: e5m2, +-1 (0x3C / 0xBC)
// for rows in the first half of their split, +-1024 (0x64 / 0xE4) for rows in the second half.
// Padding bits are zero bits on both sides: they multiply to +1 / +1024 and cancel in K - s.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
expand_fp8_kernel(const uint8_t* __restrict__ src, int srcBytes, int64_t nRows, int rowBytesOut,
                  int isTrain, const int64_t* __restrict__ off, const int64_t* __restrict__ offPad,
                  int nPairs, int splitRows, uint4* __restrict__ dst) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int per = rowBytesOut / 16;
    if (i >= nRows * per) return;
    const int64_t row = i / per;
    const int g = (int)(i - row * per);
    uint32_t base = 0xB8B8B8B8u;
    int64_t dstRow = row;
    if (isTrain) {
        int lo = 0, hi = nPairs;
        while (hi - lo > 1) {
            const int mid = (lo + hi) >> 1;
            if (off[mid] <= row) lo = mid; else hi = mid;
        }
        const int64_t r = row - off[lo];
        const bool isHi = (int)(r % splitRows) >= (splitRows >> 1);
        base = isHi ? 0xE4E4E4E4u : 0xBCBCBCBCu;
        dstRow = offPad[lo] + r;       // train blocks are padded to whole splits with zero rows
    }
    const uint8_t* s = src + row * srcBytes;
    const int b0 = 2 * g, b1 = 2 * g + 1;
    const uint32_t bits = (b0 < srcBytes ? (uint32_t)s[b0] : 0u) | ((b1 < srcBytes ? (uint32_t)s[b1] : 0u) << 8);
    uint32_t w[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const uint32_t nib = (bits >> (4 * q)) & 0xFu;
        const uint32_t sp = (nib * 0x00204081u) & 0x01010101u;   // 4 bits -> 4 bytes of 0/1
        w[q] = base ^ (sp * 0x80u);                               // set bit clears the sign
    }
    dst[dstRow * per + g] = make_uint4(w[0], w[1], w[2], w[3]);
}

template <int KC, int STAGES>
__global__ void __launch_bounds__(TC2_THREADS, 1)
hamming_tc2_kernel(const __grid_constant__ CUtensorMap mapQ, const __grid_constant__ CUtensorMap mapT,
                   const int64_t* __restrict__ qOff, const int64_t* __restrict__ tOff,
                   const int64_t* __restrict__ tOffPad, const int* __restrict__ tileStart, int nPairs,
                   int splitRows, int64_t totalQ, uint64_t* __restrict__ splitKeys) {
    constexpr int ATILES = 2;
    constexpr int QROWS = ATILES * TILE_M;
    constexpr int A_BYTES = ATILES * KC * CHUNK_BYTES;
    constexpr int HALF_BYTES = KC * CHUNK_BYTES;       // one 128-row operand block
    constexpr int B_BYTES = 2 * HALF_BYTES;            // lo block + hi block
    constexpr int KBITS = KC * 128;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* sA = smem;
    uint8_t* sB = smem + A_BYTES;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + A_BYTES + STAGES * B_BYTES);
    uint64_t* a_full = bars;
    uint64_t* b_full = bars + 1;
    uint64_t* b_empty = b_full + STAGES;
    uint64_t* acc_full = b_empty + STAGES;
    uint64_t* acc_empty = acc_full + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    int lo = 0, hi = nPairs;
    const int tile = blockIdx.x;
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (tileStart[mid] <= tile) lo = mid; else hi = mid;
    }
    const int p = lo;
    const int64_t q0 = qOff[p], qEnd = qOff[p + 1];
    const int64_t tRow0 = tOffPad[p];                   // row in the padded, expanded train array
    const int tCount = (int)(tOff[p + 1] - tOff[p]);
    const int64_t qTile0 = q0 + (int64_t)(tile - tileStart[p]) * QROWS;
    const int split = blockIdx.y;
    const int H = splitRows >> 1;                       // multiple of 128
    const int t0 = min(tCount, split * splitRows);
    const int t1 = min(tCount, t0 + splitRows);
    const int nT = t1 - t0;
    const int loCount = min(nT, H), hiCount = nT - loCount;
    const int nStagesTotal = (loCount + TILE_N - 1) / TILE_N;

    if (threadIdx.x == 0) {
        mbar_init(a_full, 1);
        for (int s = 0; s < STAGES; ++s) { mbar_init(b_full + s, 1); mbar_init(b_empty + s, 1); }
        for (int s = 0; s < 2; ++s) { mbar_init(acc_full + s, 1); mbar_init(acc_empty + s, EPI_WARPS2 * 32); }
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
                const int64_t rLo = tRow0 + t0 + (int64_t)j * TILE_N;
                for (int kc = 0; kc < KC; ++kc) {
                    tma_load_2d(sB + s * B_BYTES + kc * CHUNK_BYTES, &mapT, kc * 128, (int)rLo, b_full + s);
                    tma_load_2d(sB + s * B_BYTES + HALF_BYTES + kc * CHUNK_BYTES, &mapT, kc * 128, (int)(rLo + H), b_full + s);
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer =====
        // D = f32, A = e4m3, B = e5m2, K-major both, N = 128, M = 128
        constexpr uint32_t idesc = (1u << 4) | (0u << 7) | (1u << 10) | ((uint32_t)(TILE_N >> 3) << 17) | ((uint32_t)(TILE_M >> 4) << 24);
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
                    for (int half = 0; half < 2; ++half)
#pragma unroll
                        for (int kc = 0; kc < KC; ++kc)
#pragma unroll
                            for (int k = 0; k < 4; ++k) {
                                const uint64_t da = umma_desc_sw128(aBase + kc * CHUNK_BYTES + k * 32);
                                const uint64_t db = umma_desc_sw128(bBase + half * HALF_BYTES + kc * CHUNK_BYTES + k * 32);
                                tc_mma_f8(dcol, da, db, idesc, (half | kc | k) ? 1u : 0u);
                            }
                }
                tc_commit(b_empty + s);
                tc_commit(acc_full + as);
            }
        }
    } else {
        // ===== epilogue: warps 2..9; quarter = warp % 4, query tile = (warp - 2) / 4 =====
        const int quarter = warp & 3;
        const int myTile = (warp - 2) >> 2;
        const int row = quarter * 32 + lane;
        // lo stream: values kept as exact-integer floats; hi stream: integers + a float threshold
        float lb1 = -3.0e38f, lb2 = -3.0e38f;
        int li1 = -1, li2 = -1;
        int hb1 = INT_MIN, hi1 = -1, hb2 = INT_MIN, hi2 = -1;
        float hiThr = -3.0e38f;            // D > hiThr  <=>  hi(D) > hb2
        const uint32_t tlane = tmem_base + ((uint32_t)(quarter * 32) << 16);

        // one 32-column chunk: split every accumulator D = 1024*hi + lo on the FMA pipe
        // (hi = rint(D/1024) via the magic-number add, lo = D - 1024*hi, both exact), filter
        // 8-column groups by their maxima, and run the exact insert only where it can matter
        auto process = [&](const int (&raw)[32], int nvLo, int nvHi, int cbLo, int cbHi) {
            float f[32], lf[32];
#pragma unroll
            for (int i = 0; i < 32; ++i) {
                f[i] = __int_as_float(raw[i]);
                const float hf = __fsub_rn(__fmaf_rn(f[i], 0.0009765625f, MAGIC), MAGIC);
                lf[i] = __fmaf_rn(hf, -1024.0f, f[i]);
            }
            if (nvLo < 32 || nvHi < 32) {          // tail stage only
#pragma unroll
                for (int i = 0; i < 32; ++i) {
                    if (i >= nvLo) lf[i] = -3.0e38f;
                    if (i >= nvHi) f[i] = -3.0e38f;
                }
            }
            float gl[4], gf[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const float a01 = fmaxf(lf[q * 8 + 0], lf[q * 8 + 1]), a23 = fmaxf(lf[q * 8 + 2], lf[q * 8 + 3]);
                const float a45 = fmaxf(lf[q * 8 + 4], lf[q * 8 + 5]), a67 = fmaxf(lf[q * 8 + 6], lf[q * 8 + 7]);
                gl[q] = fmaxf(fmaxf(a01, a23), fmaxf(a45, a67));
                const float f01 = fmaxf(f[q * 8 + 0], f[q * 8 + 1]), f23 = fmaxf(f[q * 8 + 2], f[q * 8 + 3]);
                const float f45 = fmaxf(f[q * 8 + 4], f[q * 8 + 5]), f67 = fmaxf(f[q * 8 + 6], f[q * 8 + 7]);
                gf[q] = fmaxf(fmaxf(f01, f23), fmaxf(f45, f67));
            }
            const float gml = fmaxf(fmaxf(gl[0], gl[1]), fmaxf(gl[2], gl[3]));
            const float gmf = fmaxf(fmaxf(gf[0], gf[1]), fmaxf(gf[2], gf[3]));
            // exact updates: one code instance per stream, the triggered group is copied into x[]
            if (gml > lb2) {
#pragma unroll 1
                for (int q = 0; q < 4; ++q) {
                    const float g = q == 0 ? gl[0] : q == 1 ? gl[1] : q == 2 ? gl[2] : gl[3];
                    if (g > lb2) {
                        float x[8];
                        if (q == 0) { _Pragma("unroll") for (int i = 0; i < 8; ++i) x[i] = lf[i]; }
                        else if (q == 1) { _Pragma("unroll") for (int i = 0; i < 8; ++i) x[i] = lf[8 + i]; }
                        else if (q == 2) { _Pragma("unroll") for (int i = 0; i < 8; ++i) x[i] = lf[16 + i]; }
                        else { _Pragma("unroll") for (int i = 0; i < 8; ++i) x[i] = lf[24 + i]; }
                        const int cb = cbLo + q * 8;
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            if (x[i] > lb2) {
                                const bool top = x[i] > lb1;
                                lb2 = top ? lb1 : x[i]; li2 = top ? li1 : cb + i;
                                lb1 = top ? x[i] : lb1; li1 = top ? cb + i : li1;
                            }
                        }
                    }
                }
            }
            if (gmf > hiThr) {
#pragma unroll 1
                for (int q = 0; q < 4; ++q) {
                    const float g = q == 0 ? gf[0] : q == 1 ? gf[1] : q == 2 ? gf[2] : gf[3];
                    if (g > hiThr) {
                        float x[8];
                        if (q == 0) { _Pragma("unroll") for (int i = 0; i < 8; ++i) x[i] = f[i]; }
                        else if (q == 1) { _Pragma("unroll") for (int i = 0; i < 8; ++i) x[i] = f[8 + i]; }
                        else if (q == 2) { _Pragma("unroll") for (int i = 0; i < 8; ++i) x[i] = f[16 + i]; }
                        else { _Pragma("unroll") for (int i = 0; i < 8; ++i) x[i] = f[24 + i]; }
                        const int cb = cbHi + q * 8;
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            if (x[i] > hiThr) {
                                const int d = __float_as_int(__fadd_rn(x[i], MAGIC)) - MAGIC_BITS;
                                const int hv = (d + 512) >> 10;
                                const bool top = hv > hb1;
                                hb2 = top ? hb1 : hv; hi2 = top ? hi1 : cb + i;
                                hb1 = top ? hv : hb1; hi1 = top ? cb + i : hi1;
                                hiThr = hb2 == INT_MIN ? -3.0e38f : (float)(hb2 * 1024 + 512);
                            }
                        }
                    }
                }
            }
        };

        for (int j = 0; j < nStagesTotal; ++j) {
            const int as = j & 1;
            mbar_wait(acc_full + as, (j >> 1) & 1);
            tc_fence_after();
            const int colBase = j * TILE_N;
            const int validLo = min(TILE_N, loCount - colBase);
            const int validHi = max(0, min(TILE_N, hiCount - colBase));
            const uint32_t tcol = tlane + (uint32_t)(as * (ATILES * TILE_N) + myTile * TILE_N);
            const int cbLo = t0 + colBase, cbHi = t0 + H + colBase;
            int cur[32], nxt[32];
            tc_ld32(tcol, cur);
            tc_wait_ld();
            // one code instance of process(): the next chunk's tcgen05.ld is in flight under it
#pragma unroll 1
            for (int c = 0; c < 4; ++c) {
                if (c < 3) tc_ld32(tcol + (c + 1) * 32, nxt);
                process(cur, validLo - c * 32, validHi - c * 32, cbLo + c * 32, cbHi + c * 32);
                tc_wait_ld();
                if (c == 2) { tc_fence_before(); mbar_arrive(acc_empty + as); }   // every load has landed
#pragma unroll
                for (int i = 0; i < 32; ++i) cur[i] = nxt[i];
            }
        }
        // emit: lo stream -> slot 2*split, hi stream -> slot 2*split + 1
        const int64_t qRow = qTile0 + myTile * TILE_M + row;
        if (qRow < qEnd) {
            uint64_t k1 = ~0ull, k2 = ~0ull, k3 = ~0ull, k4 = ~0ull;
            if (li1 >= 0) k1 = ((uint64_t)(uint32_t)((KBITS - (int)lb1) >> 1) << 32) | (uint64_t)(uint32_t)li1;
            if (li2 >= 0) k2 = ((uint64_t)(uint32_t)((KBITS - (int)lb2) >> 1) << 32) | (uint64_t)(uint32_t)li2;
            if (hi1 >= 0) k3 = ((uint64_t)(uint32_t)((KBITS - hb1) >> 1) << 32) | (uint64_t)(uint32_t)hi1;
            if (hi2 >= 0) k4 = ((uint64_t)(uint32_t)((KBITS - hb2) >> 1) << 32) | (uint64_t)(uint32_t)hi2;
            uint64_t* o = splitKeys + ((int64_t)(2 * split) * totalQ + qRow) * 2;
            reinterpret_cast<ulonglong2*>(o)[0] = make_ulonglong2(k1, k2);
            reinterpret_cast<ulonglong2*>(o + totalQ * 2)[0] = make_ulonglong2(k3, k4);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
    }
}

template <int KC, int STAGES>
int launch_cfg2(cudaStream_t stream, dim3 grid, const CUtensorMap& mq, const CUtensorMap& mt, const int64_t* qOff,
                const int64_t* tOff, const int64_t* tOffPad, const int* tileStart, int nPairs, int splitRows,
                int64_t totalQ, uint64_t* splitKeys) {
    constexpr int SMEM = 2 * KC * CHUNK_BYTES + STAGES * 2 * KC * CHUNK_BYTES + 256 + 1024;
    auto kern = hamming_tc2_kernel<KC, STAGES>;
    MCV_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
    kern<<<grid, TC2_THREADS, SMEM, stream>>>(mq, mt, qOff, tOff, tOffPad, tileStart, nPairs, splitRows, totalQ, splitKeys);
    MCV_LAUNCH_CHECK();
    return 0;
}

}  // namespace

bool hamming_tc2_supported(int descBytes) { return descBytes >= 1 && descBytes <= 32; }

// Packed-accumulator tensor-core engine (<= 256-bit descriptors).  Same contract as
// launch_hamming_knn2; uses ctx workspaces 0 (control), 3 (keys), 24 / 25 (expanded rows).
int launch_hamming_knn2_tc2(Context* ctx, cudaStream_t stream, const KnnArgs& a) {
    if (a.nPairs <= 0) return 0;
    if (!hamming_tc2_supported(a.descBytes)) { set_error("packed tensor-core matcher supports 1..32-byte descriptors"); return -1; }
    const int KC = a.descBytes <= 16 ? 1 : 2;
    const int QROWS = 2 * TILE_M;
    const int rowBytes = KC * 128;
    const int64_t totalQ = a.hQOff[a.nPairs], totalT = a.hTOff[a.nPairs];
    if (a.hQOff[0] != 0 || a.hTOff[0] != 0) { set_error("offset arrays must start at 0"); return -1; }
    if (totalQ == 0) return 0;

    const size_t n1 = (size_t)a.nPairs + 1;
    const size_t ctlBytes = n1 * (8 + 8 + 8 + 4);
    if (ctx->pin[0].reserve(ctlBytes) || ctx->ws[0].reserve(ctlBytes)) return -1;
    MCV_CUDA_TRY(cudaEventSynchronize(ctx->ev[0]));
    int64_t* hQ = ctx->pin[0].as<int64_t>();
    int64_t* hT = hQ + n1;
    int64_t* hTP = hT + n1;
    int* hTile = reinterpret_cast<int*>(hTP + n1);
    int64_t tiles = 0, maxT = 0;
    for (int p = 0; p < a.nPairs; ++p) {
        const int64_t nq = a.hQOff[p + 1] - a.hQOff[p], nt = a.hTOff[p + 1] - a.hTOff[p];
        if (nq < 0 || nt < 0 || nt > 0x3FFFFFFFll) { set_error("bad offsets at pair %d", p); return -1; }
        hQ[p] = a.hQOff[p]; hT[p] = a.hTOff[p]; hTile[p] = (int)tiles;
        tiles += (nq + QROWS - 1) / QROWS;
        maxT = std::max(maxT, nt);
    }
    hQ[a.nPairs] = totalQ; hT[a.nPairs] = totalT; hTile[a.nPairs] = (int)tiles;

    // split geometry first: the train expansion depends on it (second half of a split = hi stream)
    int64_t wantSplits = 1;
    const int64_t fill = 2ll * ctx->sm_count;
    if (tiles < fill) wantSplits = (fill + tiles - 1) / tiles;
    const int64_t maxBlocks = std::max<int64_t>(1, (maxT + 2 * TILE_N - 1) / (2 * TILE_N));
    wantSplits = std::min<int64_t>(wantSplits, std::max<int64_t>(1, maxBlocks / 4));
    int64_t splitRows = (std::max<int64_t>(maxT, 1) + wantSplits - 1) / wantSplits;
    splitRows = (splitRows + 2 * TILE_N - 1) / (2 * TILE_N) * (2 * TILE_N);      // halves are multiples of 128
    const int nSplits = (int)std::max<int64_t>(1, (maxT + splitRows - 1) / splitRows);
    if (nSplits > 32767) { set_error("train set too large for one launch"); return -1; }
    // every pair's train block is padded to whole splits in the expanded array (zero rows = +0.0
    // in fp8), so an out-of-range lo/hi row never leaks another pair's values into an accumulator
    int64_t totalTPad = 0;
    for (int p = 0; p < a.nPairs; ++p) {
        hTP[p] = totalTPad;
        const int64_t nt = a.hTOff[p + 1] - a.hTOff[p];
        totalTPad += (nt + splitRows - 1) / splitRows * splitRows;
    }
    hTP[a.nPairs] = totalTPad;
    if (tiles > 0x7FFFFFFFll || totalQ + 2 * QROWS > 0x7FFFFFFFll || totalTPad + 2 * TILE_N > 0x7FFFFFFFll) {
        set_error("batch too large for 32-bit TMA row coordinates");
        return -1;
    }
    MCV_CUDA_TRY(cudaMemcpyAsync(ctx->ws[0].p, hQ, ctlBytes, cudaMemcpyHostToDevice, stream));
    MCV_CUDA_TRY(cudaEventRecord(ctx->ev[0], stream));
    const int64_t* dQOff = ctx->ws[0].as<int64_t>();
    const int64_t* dTOff = dQOff + n1;
    const int64_t* dTOffPad = dTOff + n1;
    const int* dTile = reinterpret_cast<const int*>(dTOffPad + n1);

    if (ctx->ws[24].reserve((size_t)totalQ * rowBytes + 1024) ||
        ctx->ws[25].reserve((size_t)std::max<int64_t>(totalTPad, 1) * rowBytes + 1024))
        return -1;
    if (totalTPad != totalT)
        MCV_CUDA_TRY(cudaMemsetAsync(ctx->ws[25].p, 0, (size_t)totalTPad * rowBytes, stream));
    {
        const int64_t nq = totalQ * (rowBytes / 16), nt = totalT * (rowBytes / 16);
        expand_fp8_kernel<<<(unsigned)((nq + 255) / 256), 256, 0, stream>>>(a.dQuery, a.descBytes, totalQ, rowBytes, 0,
                                                                            dQOff, dQOff, a.nPairs, (int)splitRows,
                                                                            ctx->ws[24].as<uint4>());
        MCV_LAUNCH_CHECK();
        if (nt > 0) {
            expand_fp8_kernel<<<(unsigned)((nt + 255) / 256), 256, 0, stream>>>(a.dTrain, a.descBytes, totalT, rowBytes, 1,
                                                                                dTOff, dTOffPad, a.nPairs, (int)splitRows,
                                                                                ctx->ws[25].as<uint4>());
            MCV_LAUNCH_CHECK();
        }
    }
    CUtensorMap mq, mt;
    if (make_map(&mq, ctx->ws[24].p, totalQ, rowBytes) || make_map(&mt, ctx->ws[25].p, totalTPad, rowBytes)) return -1;

    const int nSlots = 2 * nSplits;
    if (ctx->ws[3].reserve((size_t)nSlots * totalQ * 16)) return -1;
    uint64_t* dSplitKeys = ctx->ws[3].as<uint64_t>();
    dim3 grid((unsigned)tiles, (unsigned)nSplits);
    int rc;
    if (KC == 1) rc = launch_cfg2<1, 4>(stream, grid, mq, mt, dQOff, dTOff, dTOffPad, dTile, a.nPairs, (int)splitRows, totalQ, dSplitKeys);
    else rc = launch_cfg2<2, 2>(stream, grid, mq, mt, dQOff, dTOff, dTOffPad, dTile, a.nPairs, (int)splitRows, totalQ, dSplitKeys);
    if (rc) return rc;
    return launch_knn2_finalize(stream, dSplitKeys, nSlots, totalQ, a);
}

}  // namespace mcv

#else
#include "knn.cuh"
namespace mcv {
bool hamming_tc2_supported(int) { return false; }
int launch_hamming_knn2_tc2(Context*, cudaStream_t, const KnnArgs&) { set_error("the packed-accumulator engine is not built (make EXTRA=-DMCV_WITH_TC2)"); return -1; }
}  // namespace mcv
#endif