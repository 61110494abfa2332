// l2_tc.cu -- L2 kNN (k = 2) for float32 descriptors on the tensor cores (north_star kernel 2):
// a TMA-fed tcgen05 GEMM for  ||a - b||^2 = ||a||^2 + ||b||^2 - 2 a.b  followed by exact fp32
// re-scoring of the top candidates.
//
// Every float is split into two bf16 pieces (x = hi + lo, |lo| <= 2^-8 |x|); the GEMM accumulates
// hi.hi + hi.lo + lo.hi in fp32 (kind::f16, bf16 inputs), which reproduces a.b to ~2^-16 relative
// -- exact for OpenCV's integer-valued SIFT descriptors (<= 255), whose lo pieces are zero.
// The epilogue ranks train rows by  2 a.b - ||b||^2  (||a||^2 is constant per query), keeps the
// four best per query row and split, and a re-scoring kernel recomputes those candidates with the
// oracle's arithmetic (sequential fp32, separate multiply and add, sqrt) and takes the best two by
// (distance, index).  The reported distances are therefore bit-identical to l2_knn.cu and to cv2.  The
// indices are certified too: l2_merge_kernel bounds what the approximate ranking could have dropped and sends
// the (rare) queries it cannot certify to an exact brute-force kernel, so both engines return the same result
// on any input (tests/test_knn_l2_gpu.py runs both against the oracle and the cv2 goldens).
#include "knn.cuh"
#include "tc_ptx.cuh"
#include "l2_order.cuh"

#include <algorithm>
#include <float.h>

namespace mcv {
namespace {

using namespace tc;

constexpr int L2_EPI_WARPS = 8;
constexpr int L2_THREADS = 64 + L2_EPI_WARPS * 32;
constexpr int L2_ROW_BYTES = 512;           // 128 bf16 hi + 128 bf16 lo
constexpr int L2_CHUNKS = 4;                // 128-byte chunks per row: hi0 hi1 lo0 lo1
constexpr int L2_CAND = 4;                  // candidates kept per query row and slot

__device__ __forceinline__ uint16_t f2bf(float x) {          // round-to-nearest-even bf16
    uint32_t u = __float_as_uint(x);
    if ((u & 0x7F800000u) == 0x7F800000u) return (uint16_t)(u >> 16);
    u += 0x7FFFu + ((u >> 16) & 1u);
    return (uint16_t)(u >> 16);
}
__device__ __forceinline__ float bf2f(uint16_t b) { return __uint_as_float((uint32_t)b << 16); }

// float rows -> [hi(128 bf16) | lo(128 bf16)] rows of 512 bytes (zero padded past dim), and for
// train rows also ||b||^2 (fp32, accumulated in double)
// stats[0] = bit pattern of the largest ||row||^2 seen (non-negative floats order like their bits),
// stats[1] != 0 when the GEMM's score is not exact for some row: a lo piece is non-zero past bf16's reach, or a
// squared norm is too large for exact fp32 accumulation (integer-valued SIFT rows leave it 0)
__global__ void __launch_bounds__(128)
expand_bf16_kernel(const float* __restrict__ src, int dim, int64_t nRows, uint16_t* __restrict__ dst,
                   float* __restrict__ norms, unsigned* __restrict__ stats) {
    const int64_t row = blockIdx.x;
    if (row >= nRows) return;
    const int k = threadIdx.x;            // 128 threads = 128 dimensions
    const float x = k < dim ? src[row * dim + k] : 0.f;
    const uint16_t h = f2bf(x);
    const float rest = x - bf2f(h);
    const uint16_t l = f2bf(rest);
    dst[row * 256 + k] = h;
    dst[row * 256 + 128 + k] = l;
    const bool inexact = (rest - bf2f(l)) != 0.f || !(x == x) || x != truncf(x);
    double s = (double)x * (double)x;
    __shared__ double sm[4];
    __shared__ int sInexact;
    if (k == 0) sInexact = 0;
    __syncthreads();
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xFFFFFFFFu, s, o);
    if ((k & 31) == 0) sm[k >> 5] = s;
    if (__any_sync(0xFFFFFFFFu, inexact) && (k & 31) == 0) sInexact = 1;
    __syncthreads();
    if (k == 0) {
        const float n2 = (float)(sm[0] + sm[1] + sm[2] + sm[3]);
        norms[row] = n2;
        atomicMax(&stats[0], __float_as_uint(n2 >= 0.f ? n2 : __int_as_float(0x7F800000)));
        if (sInexact || !(n2 < 4194304.f)) stats[1] = 1u;     // sums of integer products stay exact below 2^24
    }
}

__device__ __forceinline__ void cand_insert(float x, int col, float (&b)[L2_CAND], int (&ix)[L2_CAND]) {
    // sorted descending by score; strict compare keeps the first-seen (lowest) index on ties
    if (x > b[3]) {
        if (x > b[1]) {
            b[3] = b[2]; ix[3] = ix[2];
            b[2] = b[1]; ix[2] = ix[1];
            if (x > b[0]) { b[1] = b[0]; ix[1] = ix[0]; b[0] = x; ix[0] = col; }
            else { b[1] = x; ix[1] = col; }
        } else if (x > b[2]) { b[3] = b[2]; ix[3] = ix[2]; b[2] = x; ix[2] = col; }
        else { b[3] = x; ix[3] = col; }
    }
}

template <int STAGES>
__global__ void __launch_bounds__(L2_THREADS, 1)
l2_tc_kernel(const __grid_constant__ CUtensorMap mapQ, const __grid_constant__ CUtensorMap mapT,
             const float* __restrict__ trainNorms, const int64_t* __restrict__ qOff,
             const int64_t* __restrict__ tOff, const int* __restrict__ tileStart, int nPairs, int splitRows,
             int64_t totalQ, int4* __restrict__ cands) {
    constexpr int A_BYTES = L2_CHUNKS * CHUNK_BYTES;          // one 128-query tile
    constexpr int B_BYTES = L2_CHUNKS * CHUNK_BYTES;
    constexpr int NB_BYTES = TILE_N * 4;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* sA = smem;
    uint8_t* sB = smem + A_BYTES;
    float* sNb = reinterpret_cast<float*>(smem + A_BYTES + STAGES * B_BYTES);      // 2 x 128 norms
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + A_BYTES + STAGES * B_BYTES + 2 * NB_BYTES);
    uint64_t* a_full = bars;
    uint64_t* b_full = bars + 1;
    uint64_t* b_empty = b_full + STAGES;          // MMA done with the operand stage
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
    const int64_t tRow0 = tOff[p];
    const int tCount = (int)(tOff[p + 1] - tRow0);
    const int64_t qTile0 = q0 + (int64_t)(tile - tileStart[p]) * TILE_M;
    const int split = blockIdx.y;
    const int t0 = min(tCount, split * splitRows);
    const int t1 = min(tCount, t0 + splitRows);
    const int nT = t1 - t0;
    const int nStagesTotal = (nT + TILE_N - 1) / TILE_N;

    if (threadIdx.x == 0) {
        mbar_init(a_full, 1);
        for (int s = 0; s < STAGES; ++s) { mbar_init(b_full + s, 1); mbar_init(b_empty + s, 1); }
        for (int s = 0; s < 2; ++s) { mbar_init(acc_full + s, 1); mbar_init(acc_empty + s, L2_EPI_WARPS * 32); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(256u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (elect_one()) {
            mbar_expect_tx(a_full, A_BYTES);
            for (int kc = 0; kc < L2_CHUNKS; ++kc)
                tma_load_2d(sA + kc * CHUNK_BYTES, &mapQ, kc * 128, (int)qTile0, a_full);
            for (int j = 0; j < nStagesTotal; ++j) {
                const int s = j % STAGES;
                mbar_wait(b_empty + s, ((j / STAGES) & 1) ^ 1);
                const int64_t r = tRow0 + t0 + (int64_t)j * TILE_N;
                mbar_expect_tx(b_full + s, B_BYTES);
                for (int kc = 0; kc < L2_CHUNKS; ++kc)
                    tma_load_2d(sB + s * B_BYTES + kc * CHUNK_BYTES, &mapT, kc * 128, (int)r, b_full + s);
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        // D = f32, A = B = bf16, K-major, N = 128, M = 128.  One elected lane issues (uniform datapath,
        // MMAs back to back); the waits of stage j+1 are taken before the last two MMAs of stage j
        // because the tensor pipe queues only ~2-3 MMAs (tools/mma_bench.cu).
        constexpr uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(TILE_N >> 3) << 17) | ((uint32_t)(TILE_M >> 4) << 24);
        if (elect_one() && nStagesTotal > 0) {
            mbar_wait(a_full, 0);
            mbar_wait(b_full, 0);
            tc_fence_after();
            const uint32_t aBase = smem_u32(sA);
            for (int j = 0; j < nStagesTotal; ++j) {
                const int s = j % STAGES;
                const int as = j & 1;
                const uint32_t bBase = smem_u32(sB + s * B_BYTES);
                const uint32_t dcol = tmem_base + (uint32_t)(as * TILE_N);
                // (A part, B part): hi.hi, hi.lo, lo.hi  -- part p occupies chunks 2p, 2p+1
#pragma unroll
                for (int i = 0; i < 24; ++i) {
                    const int t = i >> 3, kc = (i >> 2) & 1, k = i & 3;
                    const int pa = t == 2 ? 1 : 0, pb = t == 1 ? 1 : 0;
                    if (i == 22 && j + 1 < nStagesTotal) {
                        const int jn = j + 1;
                        mbar_wait(b_full + jn % STAGES, (jn / STAGES) & 1);
                        mbar_wait(acc_empty + (jn & 1), ((jn >> 1) & 1) ^ 1);
                        tc_fence_after();
                    }
                    const uint64_t da = umma_desc_sw128(aBase + (pa * 2 + kc) * CHUNK_BYTES + k * 32);
                    const uint64_t db = umma_desc_sw128(bBase + (pb * 2 + kc) * CHUNK_BYTES + k * 32);
                    tc_mma_f16(dcol, da, db, idesc, i ? 1u : 0u);
                }
                tc_commit(b_empty + s);
                tc_commit(acc_full + as);
            }
        }
        __syncwarp();
    } else {
        // epilogue warps 2..9: quarter = warp % 4, column half = (warp - 2) / 4 (two 32-column chunks)
        const int quarter = warp & 3;
        const int half = (warp - 2) >> 2;
        const int row = quarter * 32 + lane;
        float b[L2_CAND];
        int ix[L2_CAND];
#pragma unroll
        for (int i = 0; i < L2_CAND; ++i) { b[i] = -FLT_MAX; ix[i] = -1; }
        const uint32_t tlane = tmem_base + ((uint32_t)(quarter * 32) << 16);

        auto process = [&](const int (&raw)[32], const float* nb, int nv, int cb) {
            float sc[32];
#pragma unroll
            for (int i = 0; i < 32; i += 4) {
                const float4 n4 = *reinterpret_cast<const float4*>(nb + i);
                sc[i + 0] = fmaf(__int_as_float(raw[i + 0]), 2.f, -n4.x);
                sc[i + 1] = fmaf(__int_as_float(raw[i + 1]), 2.f, -n4.y);
                sc[i + 2] = fmaf(__int_as_float(raw[i + 2]), 2.f, -n4.z);
                sc[i + 3] = fmaf(__int_as_float(raw[i + 3]), 2.f, -n4.w);
            }
            if (nv < 32) {
#pragma unroll
                for (int i = 0; i < 32; ++i) if (i >= nv) sc[i] = -FLT_MAX;
            }
            float g[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const float m01 = fmaxf(sc[q * 8 + 0], sc[q * 8 + 1]), m23 = fmaxf(sc[q * 8 + 2], sc[q * 8 + 3]);
                const float m45 = fmaxf(sc[q * 8 + 4], sc[q * 8 + 5]), m67 = fmaxf(sc[q * 8 + 6], sc[q * 8 + 7]);
                g[q] = fmaxf(fmaxf(m01, m23), fmaxf(m45, m67));
            }
            const float gm = fmaxf(fmaxf(g[0], g[1]), fmaxf(g[2], g[3]));
            if (gm > b[3]) {
#pragma unroll
                for (int q = 0; q < 4; ++q)
                    if (g[q] > b[3]) {
#pragma unroll
                        for (int i = 0; i < 8; ++i) cand_insert(sc[q * 8 + i], cb + q * 8 + i, b, ix);
                    }
            }
        };

        const int et = threadIdx.x - 64;                  // 0..255 among the epilogue threads
        for (int j = 0; j < nStagesTotal; ++j) {
            const int as = j & 1;
            // ||b||^2 of this stage's 128 train rows: fetched while the MMAs still run, parked in a
            // double-buffered shared array, published by one named barrier of the epilogue warps
            // (the barrier of stage j+1 orders every read of buffer j&1 before its reuse at j+2)
            float myNorm = 0.f;
            if (et < TILE_N) myNorm = trainNorms[tRow0 + t0 + (int64_t)j * TILE_N + et];   // padded array
            mbar_wait(acc_full + as, (j >> 1) & 1);
            tc_fence_after();
            if (et < TILE_N) sNb[as * TILE_N + et] = myNorm;
            asm volatile("bar.sync 1, %0;" ::"n"(L2_EPI_WARPS * 32) : "memory");
            const int colBase = j * TILE_N;
            const int valid = min(TILE_N, nT - colBase);
            const uint32_t tcol = tlane + (uint32_t)(as * TILE_N + half * 64);
            const float* nb = sNb + as * TILE_N + half * 64;
            const int cb0 = t0 + colBase + half * 64;
            int va[32], vb[32];
            tc_ld32(tcol, va);
            tc_wait_ld();
            tc_ld32(tcol + 32, vb);
            process(va, nb, valid - half * 64, cb0);
            tc_wait_ld();
            tc_fence_before();
            mbar_arrive(acc_empty + as);
            process(vb, nb + 32, valid - half * 64 - 32, cb0 + 32);
        }
        const int64_t qRow = qTile0 + row;
        if (qRow < qEnd) cands[((int64_t)(split * 2 + half)) * totalQ + qRow] = make_int4(ix[0], ix[1], ix[2], ix[3]);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(256u) : "memory");
    }
}

// exact re-scoring: OpenCV's arithmetic AND accumulation order (l2_order.cuh: hal::normL2Sqr_ at 128-bit width, 16
// strided partial sums, no FMA) on every candidate, so the distances returned are cv2's bit for bit.
// One thread per (query, slot): its four candidates run as independent dependency chains over 128-bit loads (a
// float4 is exactly one SIMD lane group of the reference loop); adjacent threads share the query row (broadcast
// loads).  Emits the slot's best two as packed keys (float bits << 32) | index.
__global__ void __launch_bounds__(128)
l2_rescore_kernel(const float* __restrict__ query, const float* __restrict__ train,
                  const int64_t* __restrict__ qOff, const int64_t* __restrict__ tOff, int nPairs, int dim,
                  const int4* __restrict__ cands, int nSlots, int64_t totalQ, uint64_t* __restrict__ slotKeys,
                  float* __restrict__ slotMaxSq) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= totalQ * nSlots) return;
    const int64_t q = idx / nSlots;
    const int s = (int)(idx - q * nSlots);
    int lo = 0, hi = nPairs;
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (qOff[mid] <= q) lo = mid; else hi = mid;
    }
    const int64_t tRow0 = tOff[lo];
    const int4 c = cands[(int64_t)s * totalQ + q];
    const int id[4] = {c.x, c.y, c.z, c.w};
    const float* a = query + q * dim;
    const float* b[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) b[k] = id[k] < 0 ? a : train + (tRow0 + id[k]) * dim;   // empty slot: any valid row
    float dsq[4];
    const bool vec = (dim & 3) == 0 && ((reinterpret_cast<uintptr_t>(query) | reinterpret_cast<uintptr_t>(train)) & 15) == 0;   // rows are 16-byte aligned
    if (vec) {
        const int nb = dim >> 4;
        float acc[4][4][4];     // [candidate][accumulator][lane]
#pragma unroll
        for (int k = 0; k < 4; ++k)
#pragma unroll
            for (int ac = 0; ac < 4; ++ac)
#pragma unroll
                for (int l = 0; l < 4; ++l) acc[k][ac][l] = 0.f;
        for (int blk = 0; blk < nb; ++blk) {
#pragma unroll
            for (int ac = 0; ac < 4; ++ac) {
                const int d = 16 * blk + 4 * ac;
                const float4 av = __ldg(reinterpret_cast<const float4*>(a + d));
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const float4 bv = __ldg(reinterpret_cast<const float4*>(b[k] + d));
                    float df = __fsub_rn(av.x, bv.x); acc[k][ac][0] = __fadd_rn(__fmul_rn(df, df), acc[k][ac][0]);
                    df = __fsub_rn(av.y, bv.y); acc[k][ac][1] = __fadd_rn(__fmul_rn(df, df), acc[k][ac][1]);
                    df = __fsub_rn(av.z, bv.z); acc[k][ac][2] = __fadd_rn(__fmul_rn(df, df), acc[k][ac][2]);
                    df = __fsub_rn(av.w, bv.w); acc[k][ac][3] = __fadd_rn(__fmul_rn(df, df), acc[k][ac][3]);
                }
            }
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            float sl[4];
#pragma unroll
            for (int l = 0; l < 4; ++l)
                sl[l] = __fadd_rn(__fadd_rn(__fadd_rn(acc[k][0][l], acc[k][1][l]), acc[k][2][l]), acc[k][3][l]);
            float d = __fadd_rn(__fadd_rn(sl[0], sl[2]), __fadd_rn(sl[1], sl[3]));
            for (int e = nb << 4; e < dim; ++e) {
                const float df = __fsub_rn(a[e], b[k][e]);
                d = __fadd_rn(d, __fmul_rn(df, df));
            }
            dsq[k] = d;
        }
    } else {
#pragma unroll
        for (int k = 0; k < 4; ++k) dsq[k] = l2_sqr_cv_order_rows(a, b[k], dim);
    }
    uint64_t m1 = ~0ull, m2 = ~0ull;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        if (id[k] < 0) continue;
        const uint64_t key = ((uint64_t)__float_as_uint(__fsqrt_rn(dsq[k])) << 32) | (uint32_t)id[k];
        const uint64_t hi2 = m1 > key ? m1 : key;
        m1 = m1 < key ? m1 : key;
        m2 = m2 < hi2 ? m2 : hi2;
    }
    reinterpret_cast<ulonglong2*>(slotKeys + idx * 2)[0] = make_ulonglong2(m1, m2);
    // a slot with an empty entry saw fewer than L2_CAND rows: nothing of it was dropped
    const bool full = id[0] >= 0 && id[1] >= 0 && id[2] >= 0 && id[3] >= 0;
    slotMaxSq[idx] = full ? fmaxf(fmaxf(dsq[0], dsq[1]), fmaxf(dsq[2], dsq[3])) : -1.f;
}

// best two over a query's slots by (distance, index), ratio test in fp32 -- with a guarantee.  The GEMM ranks rows by
// an approximate score (bf16 hi/lo split without the lo.lo term, fp32 accumulation): |score error| <= eps in
// squared-distance units with eps = 2^-12 (||a||^2 + max ||b||^2), far above the real error, and 0 when every row
// is integer valued (lo pieces zero, sums exact).  A row the epilogue DROPPED from a slot scored no better than the
// slot's worst kept candidate, so its exact squared distance is at least  max_kept_exact - 4 eps.  If that bound does
// not clear the runner-up for some full slot, the query's result cannot be certified and the query is handed to
// l2_exact_fallback_kernel (brute force in the reference order); certified queries are final here.
__global__ void __launch_bounds__(256)
l2_merge_kernel(const uint64_t* __restrict__ slotKeys, const float* __restrict__ slotMaxSq, int nSlots, int64_t totalQ,
                const float* __restrict__ qNorms, const unsigned* __restrict__ stats, float ratio,
                int* __restrict__ trainIdx, float* __restrict__ dist, uint8_t* __restrict__ ratioPass,
                int* __restrict__ passCount, int* __restrict__ fbCount, int* __restrict__ fbList) {
    const int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    bool pass = false;
    if (q < totalQ) {
        uint64_t m1 = ~0ull, m2 = ~0ull;
        for (int s = 0; s < nSlots; ++s) {
            const ulonglong2 k = reinterpret_cast<const ulonglong2*>(slotKeys + (q * nSlots + s) * 2)[0];
            const uint64_t keys[2] = {k.x, k.y};
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                const uint64_t hi2 = m1 > keys[i] ? m1 : keys[i];
                m1 = m1 < keys[i] ? m1 : keys[i];
                m2 = m2 < hi2 ? m2 : hi2;
            }
        }
        int i1 = -1, i2 = -1;
        float d1 = FLT_MAX, d2 = FLT_MAX;
        if (m1 != ~0ull) { i1 = (int)(uint32_t)m1; d1 = __uint_as_float((uint32_t)(m1 >> 32)); }
        if (m2 != ~0ull) { i2 = (int)(uint32_t)m2; d2 = __uint_as_float((uint32_t)(m2 >> 32)); }
        bool certain = true;
        if (i2 >= 0) {
            const float eps = stats[1] ? 0.000244140625f * (qNorms[q] + __uint_as_float(stats[0])) : 0.f;
            const float d2sq = d2 * d2 * 1.000001f + 4.f * eps;          // upper bound of the runner-up's squared distance + margin
            for (int s = 0; s < nSlots; ++s) {
                const float mx = slotMaxSq[q * nSlots + s];
                if (mx >= 0.f && !(mx > d2sq)) certain = false;          // NaN distances are never certified
            }
        }
        if (certain) {
            reinterpret_cast<int2*>(trainIdx)[q] = make_int2(i1, i2);
            reinterpret_cast<float2*>(dist)[q] = make_float2(d1, d2);
            pass = (i2 >= 0) && (d1 < __fmul_rn(ratio, d2));
            if (ratioPass) ratioPass[q] = pass ? 1 : 0;
        } else {
            fbList[atomicAdd(fbCount, 1)] = (int)q;
        }
    }
    if (passCount) {
        const unsigned bal = __ballot_sync(0xFFFFFFFFu, pass);
        if ((threadIdx.x & 31) == 0 && bal) atomicAdd(passCount, __popc(bal));
    }
}

// Uncertified queries: exact distances to every train row of the pair, in the reference order; one block per query
// (grid-stride over the list), 256 threads stride the rows, top-2 by (distance, index) merged through shared memory.
__global__ void __launch_bounds__(256)
l2_exact_fallback_kernel(const float* __restrict__ query, const float* __restrict__ train,
                         const int64_t* __restrict__ qOff, const int64_t* __restrict__ tOff, int nPairs, int dim,
                         const int* __restrict__ fbCount, const int* __restrict__ fbList, float ratio,
                         int* __restrict__ trainIdx, float* __restrict__ dist, uint8_t* __restrict__ ratioPass,
                         int* __restrict__ passCount) {
    __shared__ uint64_t sK[256][2];
    const int n = *fbCount;
    for (int e = blockIdx.x; e < n; e += gridDim.x) {
        const int64_t q = fbList[e];
        int lo = 0, hi = nPairs;
        while (hi - lo > 1) {
            const int mid = (lo + hi) >> 1;
            if (qOff[mid] <= q) lo = mid; else hi = mid;
        }
        const int64_t tRow0 = tOff[lo];
        const int tCount = (int)(tOff[lo + 1] - tRow0);
        const float* a = query + q * dim;
        uint64_t m1 = ~0ull, m2 = ~0ull;
        for (int r = threadIdx.x; r < tCount; r += 256) {
            const float d = __fsqrt_rn(l2_sqr_cv_order_rows(a, train + (tRow0 + r) * dim, dim));
            const uint64_t key = ((uint64_t)__float_as_uint(d) << 32) | (uint32_t)r;
            const uint64_t hi2 = m1 > key ? m1 : key;
            m1 = m1 < key ? m1 : key;
            m2 = m2 < hi2 ? m2 : hi2;
        }
        sK[threadIdx.x][0] = m1; sK[threadIdx.x][1] = m2;
        __syncthreads();
        for (int w = 128; w > 0; w >>= 1) {
            if (threadIdx.x < w) {
                uint64_t a1 = sK[threadIdx.x][0], a2 = sK[threadIdx.x][1];
                const uint64_t b1 = sK[threadIdx.x + w][0], b2 = sK[threadIdx.x + w][1];
                const uint64_t keys[2] = {b1, b2};
#pragma unroll
                for (int i = 0; i < 2; ++i) {
                    const uint64_t hi2 = a1 > keys[i] ? a1 : keys[i];
                    a1 = a1 < keys[i] ? a1 : keys[i];
                    a2 = a2 < hi2 ? a2 : hi2;
                }
                sK[threadIdx.x][0] = a1; sK[threadIdx.x][1] = a2;
            }
            __syncthreads();
        }
        if (threadIdx.x == 0) {
            m1 = sK[0][0]; m2 = sK[0][1];
            int i1 = -1, i2 = -1;
            float d1 = FLT_MAX, d2 = FLT_MAX;
            if (m1 != ~0ull) { i1 = (int)(uint32_t)m1; d1 = __uint_as_float((uint32_t)(m1 >> 32)); }
            if (m2 != ~0ull) { i2 = (int)(uint32_t)m2; d2 = __uint_as_float((uint32_t)(m2 >> 32)); }
            reinterpret_cast<int2*>(trainIdx)[q] = make_int2(i1, i2);
            reinterpret_cast<float2*>(dist)[q] = make_float2(d1, d2);
            const bool pass = (i2 >= 0) && (d1 < __fmul_rn(ratio, d2));
            if (ratioPass) ratioPass[q] = pass ? 1 : 0;
            if (passCount && pass) atomicAdd(passCount, 1);
        }
        __syncthreads();
    }
}

}  // namespace

bool l2_tc_supported(int descBytes) { return descBytes >= 4 && descBytes <= 512 && descBytes % 4 == 0; }

// uses ctx workspaces 0 (control), 3 (candidates), 72 (stats, query norms, fallback list), 73 (slot maxima),
// 24 / 25 (bf16 rows), 26 (train norms), 27 (slot keys)
int launch_l2_knn2_tc(Context* ctx, cudaStream_t stream, const KnnArgs& a) {
    if (a.nPairs <= 0) return 0;
    const int dim = a.descBytes / 4;
    const int64_t totalQ = a.hQOff[a.nPairs], totalT = a.hTOff[a.nPairs];
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
        tiles += (nq + TILE_M - 1) / TILE_M;
        maxT = std::max(maxT, nt);
    }
    hQ[a.nPairs] = totalQ; hT[a.nPairs] = totalT; hTile[a.nPairs] = (int)tiles;
    if (tiles > 0x7FFFFFFFll || totalQ + 2 * TILE_M > 0x7FFFFFFFll || totalT + 2 * TILE_N > 0x7FFFFFFFll) {
        set_error("batch too large for 32-bit TMA row coordinates");
        return -1;
    }
    MCV_CUDA_TRY(cudaMemcpyAsync(ctx->ws[0].p, hQ, ctlBytes, cudaMemcpyHostToDevice, stream));
    MCV_CUDA_TRY(cudaEventRecord(ctx->ev[0], stream));
    const int64_t* dQOff = ctx->ws[0].as<int64_t>();
    const int64_t* dTOff = dQOff + n1;
    const int* dTile = reinterpret_cast<const int*>(dTOff + n1);

    if (ctx->ws[24].reserve((size_t)totalQ * L2_ROW_BYTES + 1024) ||
        ctx->ws[25].reserve((size_t)std::max<int64_t>(totalT, 1) * L2_ROW_BYTES + 1024) ||
        ctx->ws[26].reserve((size_t)(totalT + 2 * TILE_N) * 4 + 64) ||
        ctx->ws[72].reserve((size_t)totalQ * 8 + 64))
        return -1;
    float* dNorms = ctx->ws[26].as<float>();
    // ws[22]: [0..1] stats (max norm bits, inexact flag), [2] fallback count, then query norms, then the fallback list
    unsigned* dStats = ctx->ws[72].as<unsigned>();
    int* dFbCount = reinterpret_cast<int*>(dStats + 2);
    float* dQNorms = reinterpret_cast<float*>(dStats + 4);
    int* dFbList = reinterpret_cast<int*>(dQNorms + totalQ);
    MCV_CUDA_TRY(cudaMemsetAsync(dStats, 0, 16, stream));
    MCV_CUDA_TRY(cudaMemsetAsync(dNorms + totalT, 0, (size_t)2 * TILE_N * 4, stream));     // read-ahead padding
    expand_bf16_kernel<<<(unsigned)totalQ, 128, 0, stream>>>(reinterpret_cast<const float*>(a.dQuery), dim, totalQ,
                                                              ctx->ws[24].as<uint16_t>(), dQNorms, dStats);
    MCV_LAUNCH_CHECK();
    if (totalT > 0) {
        expand_bf16_kernel<<<(unsigned)totalT, 128, 0, stream>>>(reinterpret_cast<const float*>(a.dTrain), dim, totalT,
                                                                  ctx->ws[25].as<uint16_t>(), dNorms, dStats);
        MCV_LAUNCH_CHECK();
    }
    CUtensorMap mq, mt;
    if (make_map(&mq, ctx->ws[24].p, totalQ, L2_ROW_BYTES) || make_map(&mt, ctx->ws[25].p, totalT, L2_ROW_BYTES)) return -1;

    int64_t wantSplits = 1;
    const int64_t fill = 2ll * ctx->sm_count;
    if (tiles < fill) wantSplits = (fill + tiles - 1) / tiles;
    const int64_t maxStages = std::max<int64_t>(1, (maxT + TILE_N - 1) / TILE_N);
    wantSplits = std::min<int64_t>(wantSplits, std::max<int64_t>(1, maxStages / 2));
    int64_t splitRows = (std::max<int64_t>(maxT, 1) + wantSplits - 1) / wantSplits;
    splitRows = (splitRows + TILE_N - 1) / TILE_N * TILE_N;
    const int nSplits = (int)std::max<int64_t>(1, (maxT + splitRows - 1) / splitRows);
    if (nSplits > 32767) { set_error("train set too large for one launch"); return -1; }
    const int nSlots = 2 * nSplits;
    if (ctx->ws[3].reserve((size_t)nSlots * totalQ * 16)) return -1;
    int4* dCands = ctx->ws[3].as<int4>();

    constexpr int STAGES = 2;
    constexpr int SMEM = L2_CHUNKS * CHUNK_BYTES + STAGES * L2_CHUNKS * CHUNK_BYTES + 2 * TILE_N * 4 + 256 + 1024;
    auto kern = l2_tc_kernel<STAGES>;
    MCV_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
    dim3 grid((unsigned)tiles, (unsigned)nSplits);
    kern<<<grid, L2_THREADS, SMEM, stream>>>(mq, mt, dNorms, dQOff, dTOff, dTile, a.nPairs, (int)splitRows, totalQ, dCands);
    MCV_LAUNCH_CHECK();
    if (ctx->ws[27].reserve((size_t)nSlots * totalQ * 16) || ctx->ws[73].reserve((size_t)nSlots * totalQ * 4)) return -1;
    uint64_t* dSlotKeys = ctx->ws[27].as<uint64_t>();
    float* dSlotMax = ctx->ws[73].as<float>();
    l2_rescore_kernel<<<(unsigned)((totalQ * nSlots + 127) / 128), 128, 0, stream>>>(
        reinterpret_cast<const float*>(a.dQuery), reinterpret_cast<const float*>(a.dTrain), dQOff, dTOff, a.nPairs,
        dim, dCands, nSlots, totalQ, dSlotKeys, dSlotMax);
    MCV_LAUNCH_CHECK();
    l2_merge_kernel<<<(unsigned)((totalQ + 255) / 256), 256, 0, stream>>>(
        dSlotKeys, dSlotMax, nSlots, totalQ, dQNorms, dStats, a.ratio, a.dTrainIdx, a.dDist, a.dRatioPass, a.dPassCount,
        dFbCount, dFbList);
    MCV_LAUNCH_CHECK();
    // queries whose top-2 the approximate ranking could not certify (rare): exact brute force
    l2_exact_fallback_kernel<<<(unsigned)std::min<int64_t>(totalQ, 8ll * ctx->sm_count), 256, 0, stream>>>(
        reinterpret_cast<const float*>(a.dQuery), reinterpret_cast<const float*>(a.dTrain), dQOff, dTOff, a.nPairs, dim,
        dFbCount, dFbList, a.ratio, a.dTrainIdx, a.dDist, a.dRatioPass, a.dPassCount);
    MCV_LAUNCH_CHECK();
    return 0;
}

}  // namespace mcv
