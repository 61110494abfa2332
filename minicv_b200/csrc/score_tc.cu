// score_tc.cu -- the scoring stage of the H / F / E RANSAC engine with a tensor-core first-level filter.
//
// score_kernel<M> (ransac.cu) spends ~16 issue slots per (model, correspondence) evaluation, 10.5 of them in an fp32
// filter whose only job is to prove "outlier" for the 98-99 % of evaluations that are outliers by a wide margin.  The
// epipolar residual (F, E) and the reprojection residual's diagonal component (H: filter0.cuh, "homography"; its survivors
// get the float32 exact test directly, no classifier) are bilinear in (model, correspondence), so the residuals of a whole tile are a small-K GEMM
// (filter0.cuh): here tcgen05.mma.kind::tf32 computes 128 models x 64 correspondences per instruction group into TMEM
// (split-TF32 operands, K = 32) and the epilogue spends ONE comparison |r~| > T (+ one mask update) per evaluation, T
// constant per (model, stage).  What survives level 0 (~3 %: near misses, inliers, anything non-finite) goes through a
// two-sided fp32 classifier in the pair's coordinates (certainly outlier / certainly inlier; model and constants in
// registers, coordinates from the stage's shared-memory image): the first three survivors of every lane in place, the
// rest compacted ACROSS the warp into a shared-memory queue and drained 32 at a time.  The sliver the classifier cannot
// decide (~1 % of the survivors) is queued once more and gets the reference's exact test of cvmath.cuh -- the counts
// are the exact test's, bit for bit, as before.
//
//   score_tc_features_kernel   one block per image pair: finite coordinate ranges -> centre / scale (PairNorm); counting
//                              sort of the correspondences by image-1 cell (32 x 32 Morton grid) so that a stage of 64
//                              is spatially compact; every correspondence -> one 128-byte split-TF32 row, stored as the
//                              shared-memory image of the SWIZZLE_128B K-major operand (8-row groups of 1 KB), followed
//                              by the stage's 64 recentred coordinate quadruples (fp32), so that a stage is ONE 9 KB bulk
//                              copy; the sorted original coordinates (for the exact test); the image-1 box of every stage
//   score_tc_kernel            CTA = 256 models (two 128-row MMA tiles) x all correspondences of the pair.
//                              warp 0: cp.async.bulk producer (5 stages of 64 rows); warp 1: TMEM allocation + MMA issue
//                              (two accumulator buffers of 2 x 64 columns); warps 2..9: epilogue, warp w owns TMEM lanes
//                              32 (w % 4) of tile (w - 2) / 4, one model per lane.  Model rows are converted in place at
//                              CTA start (fp64 transform into the pair's coordinates, threshold constants, split-TF32
//                              row written in the swizzled layout, fence.proxy.async).  Two CTAs per SM (256 TMEM
//                              columns each).
//   score_pilot_kernel        32 models x 128 correspondences per pair through the exact test: the device-side gate that
//                              decides which scoring kernel takes the chunk (ransac.cuh, score_tc_mode)
// Models with many survivors in a stage (a good model keeps its 30 % inliers) are resolved by the whole warp, one
// correspondence per lane, as in filter_pass.cuh.
#include "ransac.cuh"
#include "tc_ptx.cuh"
#include "filter0.cuh"

#include <stdlib.h>
#include <algorithm>

namespace mcv {

// MINICV_SCORE_ENGINE (0 / 1 / 2) presets the engine: lets a whole test-suite run through one of them unchanged
static int score_engine_from_env() {
    const char* e = getenv("MINICV_SCORE_ENGINE");
    const int v = e ? atoi(e) : 0;
    return v >= 0 && v <= 2 ? v : 0;
}
std::atomic<int> g_score_engine{score_engine_from_env()};

namespace {

using namespace tc;
using namespace cvm;

constexpr int ST_NT = 64;                           // correspondences per stage (= MMA N)
constexpr int ST_MT = 2;                            // 128-model tiles per CTA
constexpr int ST_ROWS = 128 * ST_MT;
#ifndef ST_STAGES_V
#define ST_STAGES_V 5
#endif
constexpr int ST_STAGES = ST_STAGES_V;
constexpr int ST_EPI_WARPS = 4 * ST_MT;
constexpr int ST_THREADS = 64 + 32 * ST_EPI_WARPS;  // 320
#ifndef ST_DENSE_V
#define ST_DENSE_V 12
#endif
constexpr int ST_DENSE_MIN = ST_DENSE_V;                     // survivors (of 64) from which a model is resolved by the whole warp
constexpr int ST_QCAP = 32 * (ST_DENSE_MIN - 1) + 32;   // leftovers of one stage on top of < 32 waiting items
constexpr int ST_STAGE_BYTES = ST_NT * 128 + ST_NT * 16;  // operand rows + recentred coordinates = 9 KB
constexpr int ST_ROW_BYTES = F0_K * 4;              // 128
constexpr int ST_TMEM_COLS = 2 * ST_MT * ST_NT;     // 256
constexpr int ST_CELLS = 1024;
#ifndef ST_INPLACE_V
#define ST_INPLACE_V 3
#endif
constexpr int ST_INPLACE = ST_INPLACE_V;           // survivors per lane and stage classified in place (the rest is queued)

static_assert(ST_TMEM_COLS == 256, "TMEM allocation is a power of two");
static_assert(ST_ROW_BYTES == 128, "one SWIZZLE_128B row per operand row");

// Stages of pair p start at stage stage_base(off[p], p) of the stage-image, sorted-coordinate (64 rows per stage) and box
// buffers (every pair padded to whole stages).
__host__ __device__ inline int64_t stage_base(int64_t offP, int64_t p) { return (offP + ST_NT - 1) / ST_NT + p * 2; }

__device__ __forceinline__ bool finite_d(double x) { return fabs(x) <= DBL_MAX; }

template <int MODEL>
__global__ void __launch_bounds__(256)
score_tc_features_kernel(const float4* __restrict__ pf, const double* __restrict__ pd, const int64_t* __restrict__ off,
                         int64_t pairBase, PairNorm* __restrict__ norms, float* __restrict__ feat,
                         void* __restrict__ sortedRaw, StageBox* __restrict__ boxes, const int* __restrict__ gate) {
    constexpr int isE = MODEL == RM_ESSENTIAL ? 1 : 0;
    constexpr bool isH = MODEL == RM_HOMOGRAPHY;
    if (gate && score_gate_classic(gate)) return;             // the fp32-filter kernel scores this chunk
    const int p = blockIdx.x;
    const int64_t base = off[p];
    const int n = (int)(off[p + 1] - base);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    __shared__ double sLo[8][4], sHi[8][4];
    __shared__ int sMx[4];
    __shared__ double sRed[8][2];
    __shared__ PairNorm sPn;
    __shared__ int sHist[ST_CELLS];
    __shared__ int sWarpSum[8];
    auto load = [&](int j, double& x1, double& y1, double& x2, double& y2) {
        if (isE) {
            const double4 q = reinterpret_cast<const double4*>(pd)[base + j];
            x1 = q.x; y1 = q.y; x2 = q.z; y2 = q.w;
        } else {
            const float4 q = pf[base + j];
            x1 = q.x; y1 = q.y; x2 = q.z; y2 = q.w;
        }
    };
    // ---- ranges of the finite coordinates, maxima of the fp32 view (a NaN sorts above everything as a bit pattern)
    double lo[4] = {DBL_MAX, DBL_MAX, DBL_MAX, DBL_MAX}, hi[4] = {-DBL_MAX, -DBL_MAX, -DBL_MAX, -DBL_MAX};
    int mx[4] = {0, 0, 0, 0};
    if (tid < 4) sMx[tid] = 0;
    for (int i = tid; i < ST_CELLS; i += 256) sHist[i] = 0;
    for (int j = tid; j < n; j += 256) {
        double c[4];
        load(j, c[0], c[1], c[2], c[3]);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            if (finite_d(c[i])) { lo[i] = fmin(lo[i], c[i]); hi[i] = fmax(hi[i], c[i]); }
            mx[i] = max(mx[i], __float_as_int(fabsf((float)c[i])));
        }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        for (int o = 16; o > 0; o >>= 1) {
            lo[i] = fmin(lo[i], __shfl_xor_sync(0xFFFFFFFFu, lo[i], o));
            hi[i] = fmax(hi[i], __shfl_xor_sync(0xFFFFFFFFu, hi[i], o));
        }
        mx[i] = __reduce_max_sync(0xFFFFFFFFu, mx[i]);
        if (lane == 0) { sLo[warp][i] = lo[i]; sHi[warp][i] = hi[i]; }
    }
    __syncthreads();
    if (lane == 0) {
#pragma unroll
        for (int i = 0; i < 4; ++i) atomicMax(&sMx[i], mx[i]);
    }
    __syncthreads();
    if (tid == 0) {
        double l[4], h[4];
        for (int i = 0; i < 4; ++i) {
            l[i] = sLo[0][i]; h[i] = sHi[0][i];
            for (int w = 1; w < 8; ++w) { l[i] = fmin(l[i], sLo[w][i]); h[i] = fmax(h[i], sHi[w][i]); }
        }
        PairNorm pn;
        pair_norm_init(l, h, pn);
        pn.X1 = __int_as_float(sMx[0]); pn.Y1 = __int_as_float(sMx[1]);
        pn.X2 = __int_as_float(sMx[2]); pn.Y2 = __int_as_float(sMx[3]);
        pn.psiMax = 0.f; pn.v2Max = 0.f;
        sPn = pn;
    }
    __syncthreads();
    const PairNorm pn = sPn;
    // ---- counting sort by image-1 cell: histogram, exclusive scan, scatter
    for (int j = tid; j < n; j += 256) {
        double x1, y1, x2, y2;
        load(j, x1, y1, x2, y2);
        atomicAdd(&sHist[hat_cell(hat_coords(pn, x1, y1, x2, y2))], 1);
    }
    __syncthreads();
    {
        int v[4], s = 0;
#pragma unroll
        for (int i = 0; i < 4; ++i) { v[i] = sHist[4 * tid + i]; s += v[i]; }
        int incl = s;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(0xFFFFFFFFu, incl, o);
            if (lane >= o) incl += t;
        }
        if (lane == 31) sWarpSum[warp] = incl;
        __syncthreads();
        int wbase = 0;
        for (int w = 0; w < warp; ++w) wbase += sWarpSum[w];
        int run = wbase + incl - s;
#pragma unroll
        for (int i = 0; i < 4; ++i) { sHist[4 * tid + i] = run; run += v[i]; }
    }
    __syncthreads();
    const int64_t stage0 = stage_base(base, pairBase + p);
    const int64_t row0 = stage0 * ST_NT;
    uint8_t* img = reinterpret_cast<uint8_t*>(feat) + stage0 * ST_STAGE_BYTES;
    const int npad = (n + ST_NT - 1) / ST_NT * ST_NT;
    double pmax = 0, vmax = 0;
    for (int j = tid; j < n; j += 256) {
        double x1, y1, x2, y2;
        load(j, x1, y1, x2, y2);
        const Hat h = hat_coords(pn, x1, y1, x2, y2);
        const int pos = atomicAdd(&sHist[hat_cell(h)], 1);
        float row[F0_K];
        double v2 = 0;
        const double n2 = isH ? filter0_features_h(h, row) : filter0_features(h, row, &v2);
        if (n2 > pmax) pmax = n2;                        // NaN: not taken (the pair is switched off through X1..Y2)
        if (v2 > vmax) vmax = v2;
        uint8_t* st = img + (int64_t)(pos / ST_NT) * ST_STAGE_BYTES;
        const int rs = pos % ST_NT, r = rs & 7;
        float4* out = reinterpret_cast<float4*>(st + (rs >> 3) * 1024 + r * 128);
#pragma unroll
        for (int c = 0; c < 8; ++c) out[c ^ r] = make_float4(row[4 * c], row[4 * c + 1], row[4 * c + 2], row[4 * c + 3]);
        // coordinate block of the stage: recentred (F / E: the classifier works in the pair's coordinates) or original
        // (H: the survivors get the float32 exact test directly)
        reinterpret_cast<float4*>(st + ST_NT * 128)[rs] = isH ? make_float4((float)x1, (float)y1, (float)x2, (float)y2)
                                                              : make_float4((float)h.u1, (float)h.v1, (float)h.u2, (float)h.v2);
        if (isE) reinterpret_cast<double4*>(sortedRaw)[row0 + pos] = make_double4(x1, y1, x2, y2);
        else reinterpret_cast<float4*>(sortedRaw)[row0 + pos] = make_float4((float)x1, (float)y1, (float)x2, (float)y2);
    }
    for (int j = n + tid; j < npad; j += 256) {          // padding rows of the last stage
        uint8_t* st = img + (int64_t)(j / ST_NT) * ST_STAGE_BYTES;
        const int rs = j % ST_NT;
        float4* out = reinterpret_cast<float4*>(st + (rs >> 3) * 1024 + (rs & 7) * 128);
#pragma unroll
        for (int c = 0; c < 8; ++c) out[c] = make_float4(0.f, 0.f, 0.f, 0.f);
        reinterpret_cast<float4*>(st + ST_NT * 128)[rs] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (isE) reinterpret_cast<double4*>(sortedRaw)[row0 + j] = make_double4(0, 0, 0, 0);
        else reinterpret_cast<float4*>(sortedRaw)[row0 + j] = make_float4(0, 0, 0, 0);
    }
    for (int o = 16; o > 0; o >>= 1) {
        pmax = fmax(pmax, __shfl_xor_sync(0xFFFFFFFFu, pmax, o));
        vmax = fmax(vmax, __shfl_xor_sync(0xFFFFFFFFu, vmax, o));
    }
    if (lane == 0) { sRed[warp][0] = pmax; sRed[warp][1] = vmax; }
    __syncthreads();                                     // also orders the sorted rows before the box pass below
    if (tid == 0) {
        for (int w = 1; w < 8; ++w) { pmax = fmax(pmax, sRed[w][0]); vmax = fmax(vmax, sRed[w][1]); }
        PairNorm o = pn;
        o.psiMax = (float)(sqrt(pmax) * (1.0 + 1e-6)) * 1.0000002f;
        o.v2Max = (float)(vmax * (1.0 + 1e-6)) * 1.0000002f;
        norms[p] = o;
    }
    // ---- image-1 box of every stage (one warp per stage, two rows per lane)
    StageBox* bx = boxes + stage0;
    for (int s = warp; s * ST_NT < n; s += 8) {
        double ulo = DBL_MAX, uhi = -DBL_MAX, vlo = DBL_MAX, vhi = -DBL_MAX;
        double us[2], vs[2];
        bool ok[2];
#pragma unroll
        for (int h2 = 0; h2 < 2; ++h2) {
            const int j = s * ST_NT + h2 * 32 + lane;
            ok[h2] = j < n;
            us[h2] = 0; vs[h2] = 0;
            if (ok[h2]) {
                double x1, y1;
                if (isE) { const double4 q = reinterpret_cast<const double4*>(sortedRaw)[row0 + j]; x1 = q.x; y1 = q.y; }
                else { const float4 q = reinterpret_cast<const float4*>(sortedRaw)[row0 + j]; x1 = q.x; y1 = q.y; }
                us[h2] = (x1 - pn.c1x) * pn.k; vs[h2] = (y1 - pn.c1y) * pn.k;
                ulo = fmin(ulo, us[h2]); uhi = fmax(uhi, us[h2]); vlo = fmin(vlo, vs[h2]); vhi = fmax(vhi, vs[h2]);
            }
        }
        for (int o = 16; o > 0; o >>= 1) {
            ulo = fmin(ulo, __shfl_xor_sync(0xFFFFFFFFu, ulo, o)); uhi = fmax(uhi, __shfl_xor_sync(0xFFFFFFFFu, uhi, o));
            vlo = fmin(vlo, __shfl_xor_sync(0xFFFFFFFFu, vlo, o)); vhi = fmax(vhi, __shfl_xor_sync(0xFFFFFFFFu, vhi, o));
        }
        const float cu = (float)(0.5 * ulo + 0.5 * uhi), cv = (float)(0.5 * vlo + 0.5 * vhi);
        double pu = 0, pv = 0;
#pragma unroll
        for (int h2 = 0; h2 < 2; ++h2)
            if (ok[h2]) { pu = fmax(pu, fabs(us[h2] - (double)cu)); pv = fmax(pv, fabs(vs[h2] - (double)cv)); }
        for (int o = 16; o > 0; o >>= 1) {
            pu = fmax(pu, __shfl_xor_sync(0xFFFFFFFFu, pu, o)); pv = fmax(pv, __shfl_xor_sync(0xFFFFFFFFu, pv, o));
        }
        if (lane == 0) {
            StageBox b;
            b.cu = cu; b.cv = cv;
            b.pu = (float)(pu * (1.0 + 1e-6)) * 1.0000002f; b.pv = (float)(pv * (1.0 + 1e-6)) * 1.0000002f;
            bx[s] = b;
        }
    }
}

struct EpiModel { double m[9]; };

// bit i = "column i is NOT certainly an outlier", i.e. !(|v_i| > T).  One FADD and one funnel shift per value: the sign of
// T - |v| is set exactly when |v| > T; a NaN (non-finite value or threshold: CUDA's canonical NaN has the sign bit clear)
// and +0 (equality) leave it clear, so those survive.  Two independent shift chains.
__device__ __forceinline__ uint32_t survivors32(const int (&v)[32], float T) {
    uint32_t ra = 0, rb = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        ra = __funnelshift_l(__float_as_uint(T - fabsf(__int_as_float(v[i]))), ra, 1);
        rb = __funnelshift_l(__float_as_uint(T - fabsf(__int_as_float(v[16 + i]))), rb, 1);
    }
    // ra bit (15 - i) = reject_i, rb bit (15 - i) = reject_(16 + i)
    return ~__brev((ra << 16) | rb);
}

template <int MODEL, bool AUDIT>
__global__ void __launch_bounds__(ST_THREADS, 2)
score_tc_kernel(const int64_t* __restrict__ off, int64_t pairBase, const PairNorm* __restrict__ norms,
                const float* __restrict__ feat, const void* __restrict__ sortedRaw, const StageBox* __restrict__ boxes,
                const double* __restrict__ models, const int* __restrict__ slotList, const int* __restrict__ nTotal,
                int maxSlots, float thr2, int* __restrict__ counts, unsigned long long* __restrict__ audit,
                const int* __restrict__ gate, const int* __restrict__ nBegin) {
    constexpr int isE = MODEL == RM_ESSENTIAL ? 1 : 0;
    constexpr bool isH = MODEL == RM_HOMOGRAPHY;
    constexpr int A_BYTES = ST_ROWS * ST_ROW_BYTES;           // 32 KB
    if (gate && score_gate_classic(gate)) return;
    constexpr int B_BYTES = ST_STAGE_BYTES;                   // 9 KB per stage: 64 operand rows + 64 coordinate quadruples
    const int p = blockIdx.y;
    const int total = nTotal[p];
    const int begin = nBegin ? nBegin[p] : 0;               // models [begin, total) of the compacted list
    const int m0 = blockIdx.x * ST_ROWS;
    if (begin >= total || m0 >= total || m0 + ST_ROWS <= begin) return;

    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* sA = smem;
    uint8_t* sB = sA + A_BYTES;
    float* sM0 = reinterpret_cast<float*>(sB + ST_STAGES * B_BYTES);               // [12][ST_ROWS]: g[9], K2, K1, KD
    int* sSlot = reinterpret_cast<int*>(sM0 + 12 * ST_ROWS);                       // [ST_ROWS] model slot (fp64 model for the exact test)
    int* sCnt = sSlot + ST_ROWS;
    uint32_t* sQ = reinterpret_cast<uint32_t*>(sCnt + ST_ROWS);                    // [ST_EPI_WARPS][ST_QCAP] level-0 leftovers
    uint32_t* sQ2 = sQ + ST_EPI_WARPS * ST_QCAP;                                   // [ST_EPI_WARPS][64] undecided items
    uint64_t* bars = reinterpret_cast<uint64_t*>(sQ2 + ST_EPI_WARPS * 64);
    uint64_t* b_full = bars;
    uint64_t* b_empty = b_full + ST_STAGES;
    uint64_t* acc_full = b_empty + ST_STAGES;
    uint64_t* acc_empty = acc_full + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t base = off[p];
    const int n = (int)(off[p + 1] - base);
    const int nStages = (n + ST_NT - 1) / ST_NT;
    const PairNorm pn = norms[p];
    const int64_t stage0 = stage_base(base, pairBase + p);
    const int64_t row0 = stage0 * ST_NT;
    const uint8_t* img = reinterpret_cast<const uint8_t*>(feat) + stage0 * ST_STAGE_BYTES;

    if (threadIdx.x == 0) {
        // an operand stage is free again when its MMAs have completed AND every epilogue warp has read its coordinates
        for (int s = 0; s < ST_STAGES; ++s) { mbar_init(b_full + s, 1); mbar_init(b_empty + s, 1 + ST_EPI_WARPS); }
        for (int s = 0; s < 2; ++s) { mbar_init(acc_full + s, 1); mbar_init(acc_empty + s, ST_EPI_WARPS * 32); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        // the first operand stages are requested right away: they land while the epilogue threads convert their models
        // (nothing but this thread has touched the barriers yet)
        for (int j = 0; j < min(nStages, ST_STAGES); ++j) {
            mbar_expect_tx(b_full + j, B_BYTES);
            bulk_load_1d(sB + j * B_BYTES, img + (int64_t)j * B_BYTES, B_BYTES, b_full + j);
        }
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"((uint32_t)ST_TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    // ---- epilogue threads: their own model row -> split-TF32 operand row, threshold and classifier constants
    const int ew = warp - 2;                         // 0..7 for the epilogue warps
    const int tileM = ew >> 2;
    const int quarter = warp & 3;
    const int row = tileM * 128 + quarter * 32 + lane;          // row of the CTA's 256 (valid for warps >= 2)
    Model0 mc;
    mc.base = INFINITY;
    float Hf[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};      // homography: the float32 matrix of the exact test
    bool active = false;
    double fhA[AUDIT ? 9 : 1];
    double emA = 0;
    if (warp >= 2) {
        active = m0 + row < total && m0 + row >= begin;
        double m[9];
        int slot = 0;
        if (active) {
            slot = slotList[(int64_t)p * maxSlots + m0 + row];
            const double* src = models + ((int64_t)p * maxSlots + slot) * 9;
#pragma unroll
            for (int j = 0; j < 9; ++j) m[j] = src[j];
        } else {
#pragma unroll
            for (int j = 0; j < 9; ++j) m[j] = 0;
        }
        float frow[F0_K];
        if (isH) {
#pragma unroll
            for (int j = 0; j < 8; ++j) Hf[j] = (float)m[j];
            filter0_model_h(pn, Hf, thr2, frow, mc, AUDIT ? fhA : nullptr, AUDIT ? &emA : nullptr);
        } else {
            filter0_model(isE, pn, m, thr2, frow, mc, AUDIT ? fhA : nullptr, AUDIT ? &emA : nullptr);
        }
        const int r = row & 7;
        float4* arow = reinterpret_cast<float4*>(sA + (row >> 3) * 1024 + r * 128);
#pragma unroll
        for (int c = 0; c < 8; ++c) arow[c ^ r] = make_float4(frow[4 * c], frow[4 * c + 1], frow[4 * c + 2], frow[4 * c + 3]);
#pragma unroll
        for (int j = 0; j < 9; ++j) sM0[j * ST_ROWS + row] = isH ? (j < 8 ? Hf[j] : 0.f) : mc.g[j];
        sSlot[row] = slot;
        sM0[9 * ST_ROWS + row] = mc.K2; sM0[10 * ST_ROWS + row] = mc.K1; sM0[11 * ST_ROWS + row] = mc.KD;
        sCnt[row] = 0;
        fence_async_smem();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (elect_one()) {
            for (int j = ST_STAGES; j < nStages; ++j) {   // stages 0 .. ST_STAGES - 1 were requested at CTA start
                const int s = j % ST_STAGES;
                mbar_wait_backoff(b_empty + s, ((j / ST_STAGES) & 1) ^ 1);
                mbar_expect_tx(b_full + s, B_BYTES);
                bulk_load_1d(sB + s * B_BYTES, img + (int64_t)j * B_BYTES, B_BYTES, b_full + s);
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        // D = f32, A = B = TF32, K-major, M = 128, N = 64; K = 32 as four instructions 32 bytes apart inside the 128-byte rows
        constexpr uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(ST_NT >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        if (elect_one()) {
            const uint32_t aBase = smem_u32(sA);
            for (int j = 0; j < nStages; ++j) {
                const int s = j % ST_STAGES, buf = j & 1;
                mbar_wait_backoff(b_full + s, (j / ST_STAGES) & 1);
                mbar_wait_backoff(acc_empty + buf, ((j >> 1) & 1) ^ 1);
                tc_fence_after();
                const uint32_t bBase = smem_u32(sB + s * B_BYTES);
#pragma unroll
                for (int t = 0; t < ST_MT; ++t) {
                    const uint32_t dcol = tmem_base + (uint32_t)(buf * ST_MT * ST_NT + t * ST_NT);
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        tc_mma_tf32(dcol, umma_desc_sw128(aBase + t * 128 * ST_ROW_BYTES + k * 32), umma_desc_sw128(bBase + k * 32),
                                    idesc, (uint32_t)k);
                }
                tc_commit(b_empty + s);
                tc_commit(acc_full + buf);
            }
        }
        __syncwarp();
    } else {
        constexpr unsigned FULL = 0xFFFFFFFFu;
        uint32_t* q = sQ + ew * ST_QCAP;
        uint32_t* q2 = sQ2 + ew * 64;
        const int rowBase = row - lane;
        int qn = 0, q2n = 0;                           // items queued (warp-uniform)
        int cnt = 0;                                   // inliers counted in registers (this lane's model)
        const uint32_t tlane = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(tileM * ST_NT);
        const StageBox* bx = boxes + stage0;
        const double thrk = (double)thr2 * pn.k * pn.k;
        const float thrkM = (float)(thrk * (double)CLASS_MARGIN), thrkL = (float)(thrk * (double)CLASS_LOW);
        float auditMax = 0.f;

        auto load_model = [&](int r) {                 // fp64 model of a row, for the (rare) exact test
            EpiModel mm;
            const double* src = models + ((int64_t)p * maxSlots + sSlot[r]) * 9;
#pragma unroll
            for (int j = 0; j < 9; ++j) mm.m[j] = src[j];
            return mm;
        };
        auto load_m0 = [&](int r) {
            Model0 mm;
#pragma unroll
            for (int j = 0; j < 9; ++j) mm.g[j] = sM0[j * ST_ROWS + r];
            mm.K2 = sM0[9 * ST_ROWS + r]; mm.K1 = sM0[10 * ST_ROWS + r]; mm.KD = sM0[11 * ST_ROWS + r];
            return mm;
        };
        // exact test of one undecided item (row within the warp's 32, sorted correspondence index)
        auto exact_item = [&](uint32_t item) {
            const int r = rowBase + (int)(item & 31u);
            const int j = (int)(item >> 5);
            const EpiModel mm = load_model(r);
            bool in;
            if (isE) {
                const double4 c = reinterpret_cast<const double4*>(sortedRaw)[row0 + j];
                in = e_inlier(mm.m, c.x, c.y, c.z, c.w, thr2);
            } else {
                const float4 c = __ldg(reinterpret_cast<const float4*>(sortedRaw) + row0 + j);
                in = f_inlier(mm.m, (double)c.x, (double)c.y, (double)c.z, (double)c.w, thr2);
            }
            if (in) atomicAdd(&sCnt[r], 1);
        };
        // all lanes: queue the items the classifier could not decide (und = this lane has one), exact test per full warp
        auto push_undecided = [&](bool und, uint32_t item) {
            const unsigned b = __ballot_sync(FULL, und);
            if (b) {
                if (und) q2[q2n + __popc(b & ((1u << lane) - 1u))] = item;
                q2n += __popc(b);
                __syncwarp();
                if (q2n >= 32) {
                    q2n -= 32;
                    exact_item(q2[q2n + lane]);
                    __syncwarp();
                }
            }
        };
        // one (model, correspondence): 0 outlier, 1 inlier, 2 undecided.  F / E: the two-sided classifier in the pair's
        // coordinates; H: the float32 exact test itself on the original coordinates (35 instructions, nothing to decide later)
        auto classify = [&](const Model0& mm, const float4& h) {
            if (isH) return h_error(mm.g, h.x, h.y, h.z, h.w) <= thr2 ? 1 : 0;
            return hat_classify(isE, mm, thrkM, thrkL, h.x, h.y, h.z, h.w);
        };
        // 32 queued leftovers, one per lane (all lanes call): model constants from shared memory, coordinates from the
        // stage image in global memory (its shared-memory copy may be gone)
        auto resolve_item = [&](uint32_t item, bool valid) {
            int cls = 0;
            if (valid) {
                const int r = rowBase + (int)(item & 31u);
                const int j = (int)(item >> 5);
                const Model0 mm = load_m0(r);
                const float4 h = __ldg(reinterpret_cast<const float4*>(img + (int64_t)(j / ST_NT) * ST_STAGE_BYTES + ST_NT * 128) + (j % ST_NT));
                cls = classify(mm, h);
                if (cls == 1) atomicAdd(&sCnt[r], 1);
            }
            if (!isH) push_undecided(cls == 2, item);
        };

        for (int j = 0; j < nStages; ++j) {
            const int s = j % ST_STAGES, buf = j & 1;
            const int c0 = j * ST_NT;
            const int valid = min(ST_NT, n - c0);
            const float4 b4 = __ldg(reinterpret_cast<const float4*>(bx + j));
            StageBox sb;
            sb.cu = b4.x; sb.cv = b4.y; sb.pu = b4.z; sb.pv = b4.w;
            const float T = isH ? filter0_threshold_h(mc, sb) : filter0_threshold(isE, mc, sb);
            // the stage's coordinate block of the global image (8 lines) into L1: the queue drains gather from it
            if (lane < 8) asm volatile("prefetch.global.L1 [%0];" ::"l"(img + (int64_t)j * ST_STAGE_BYTES + ST_NT * 128 + lane * 128));
            mbar_wait(acc_full + buf, (j >> 1) & 1);
            tc_fence_after();
            const uint32_t tcol = tlane + (uint32_t)(buf * ST_MT * ST_NT);
            uint32_t mlo = 0, mhi = 0;
            {
                int v[32];
                tc_ld32(tcol, v);
                tc_wait_ld();
                mlo = survivors32(v, T);
                if (AUDIT) {
#pragma unroll 1
                    for (int i = 0; i < 32 && i < valid; ++i) {
                        double x1, y1, x2, y2;
                        if (isE) { const double4 c = reinterpret_cast<const double4*>(sortedRaw)[row0 + c0 + i]; x1 = c.x; y1 = c.y; x2 = c.z; y2 = c.w; }
                        else { const float4 c = reinterpret_cast<const float4*>(sortedRaw)[row0 + c0 + i]; x1 = c.x; y1 = c.y; x2 = c.z; y2 = c.w; }
                        const Hat h = hat_coords(pn, x1, y1, x2, y2);
                        const double tt = h.u2 + h.v2;
                        const double ref = isH ? fhA[0] * h.u1 + fhA[1] * h.v1 + fhA[2] - tt * (fhA[3] * h.u1 + fhA[4] * h.v1 + fhA[5])
                                               : fhA[0] * h.u2 * h.u1 + fhA[1] * h.u2 * h.v1 + fhA[2] * h.u2 + fhA[3] * h.v2 * h.u1 +
                                                     fhA[4] * h.v2 * h.v1 + fhA[5] * h.v2 + fhA[6] * h.u1 + fhA[7] * h.v1 + fhA[8];
                        const float ratio = (float)(fabs((double)__int_as_float(v[i]) - ref) / emA);
                        if (active && mc.base < INFINITY && ratio > auditMax) auditMax = ratio;
                    }
                }
                tc_ld32(tcol + 32, v);
                tc_wait_ld();
                tc_fence_before();
                mbar_arrive(acc_empty + buf);
                mhi = survivors32(v, T);
            }
            if (valid < ST_NT) {
                mlo &= valid >= 32 ? FULL : (1u << valid) - 1u;
                mhi &= valid <= 32 ? 0u : (1u << (valid - 32)) - 1u;
            }
            if (!active) { mlo = 0; mhi = 0; }
            // the stage's recentred coordinates (shared-memory image, valid until this warp arrives on b_empty below)
            const float4* sHat = reinterpret_cast<const float4*>(sB + s * B_BYTES + ST_NT * 128);
            // ---- models that kept many (a good model keeps its inliers; any model keeps the correspondences around its
            //      epipole, where the residual vanishes for every x2): the whole warp classifies them, one / two
            //      correspondences per lane, model broadcast from shared memory
            unsigned dense = __ballot_sync(FULL, __popc(mlo) + __popc(mhi) >= ST_DENSE_MIN);
            while (dense) {
                const int dm = __ffs(dense) - 1;
                dense &= dense - 1;
                const uint32_t dlo = __shfl_sync(FULL, mlo, dm), dhi = __shfl_sync(FULL, mhi, dm);
                const Model0 mm = load_m0(rowBase + dm);
                const float4 h0 = sHat[lane], h1 = sHat[32 + lane];
                int c0cls = classify(mm, h0);
                int c1cls = classify(mm, h1);
                if (!((dlo >> lane) & 1u)) c0cls = 0;
                if (!((dhi >> lane) & 1u)) c1cls = 0;
                const int c = __popc(__ballot_sync(FULL, c0cls == 1)) + __popc(__ballot_sync(FULL, c1cls == 1));
                if (!isH && __any_sync(FULL, c0cls == 2 || c1cls == 2)) {
                    push_undecided(c0cls == 2, ((uint32_t)(c0 + lane) << 5) | (uint32_t)dm);
                    push_undecided(c1cls == 2, ((uint32_t)(c0 + 32 + lane) << 5) | (uint32_t)dm);
                }
                if (lane == dm) { cnt += c; mlo = 0; mhi = 0; }
            }
            // ---- everybody's first ST_INPLACE survivors in place, as ONE straight-line block (independent dependency chains;
            //      a lane without a survivor classifies column 0 and drops the result): own model in registers, coordinates
            //      from shared memory
            {
                int bsel[ST_INPLACE], cls[ST_INPLACE];
#pragma unroll
                for (int t = 0; t < ST_INPLACE; ++t) {           // lowest set bit of (mhi:mlo), -1 if none, cleared
                    const bool lo = mlo != 0;
                    const uint32_t cur = lo ? mlo : mhi;
                    bsel[t] = cur ? __ffs(cur) - 1 + (lo ? 0 : 32) : -1;
                    const uint32_t rest = cur & (cur - 1);
                    mhi = lo ? mhi : rest;
                    mlo = lo ? rest : 0u;
                }
                bool anyUnd = false;
#pragma unroll
                for (int t = 0; t < ST_INPLACE; ++t) {
                    const float4 h = sHat[bsel[t] & 63];
                    cls[t] = isH ? (h_error(Hf, h.x, h.y, h.z, h.w) <= thr2 ? 1 : 0) : hat_classify(isE, mc, thrkM, thrkL, h.x, h.y, h.z, h.w);
                    if (bsel[t] < 0) cls[t] = 0;
                    cnt += cls[t] == 1 ? 1 : 0;
                    anyUnd |= cls[t] == 2;
                }
                if (!isH && __any_sync(FULL, anyUnd)) {
#pragma unroll
                    for (int t = 0; t < ST_INPLACE; ++t) push_undecided(cls[t] == 2, ((uint32_t)(c0 + bsel[t]) << 5) | (uint32_t)lane);
                }
            }
            // ---- the rest: compacted across the warp, drained 32 at a time
            const int mine = __popc(mlo) + __popc(mhi);
            if (__any_sync(FULL, mine != 0)) {
                int incl = mine;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const int t = __shfl_up_sync(FULL, incl, o);
                    if (lane >= o) incl += t;
                }
                int pos = qn + incl - mine;
                const uint32_t tag = ((uint32_t)c0 << 5) | (uint32_t)lane;
                while (mlo | mhi) {                              // one trip per leftover of the fullest lane
                    const bool lo = mlo != 0;
                    const uint32_t cur = lo ? mlo : mhi;
                    q[pos++] = tag + ((uint32_t)(__ffs(cur) - 1 + (lo ? 0 : 32)) << 5);
                    const uint32_t rest = cur & (cur - 1);
                    mhi = lo ? mhi : rest;
                    mlo = lo ? rest : 0u;
                }
                qn += __shfl_sync(FULL, incl, 31);
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(b_empty + s);     // this warp is done with the stage's shared-memory image
            while (qn >= 32) {
                qn -= 32;
                resolve_item(q[qn + lane], true);
            }
            __syncwarp();
        }
        resolve_item(q[lane], lane < qn);
        if (!isH && lane < q2n) exact_item(q2[lane]);
        __syncwarp();
        if (active) counts[(int64_t)p * maxSlots + m0 + row] = cnt + sCnt[row];
        if (AUDIT) {
            for (int o = 16; o > 0; o >>= 1) auditMax = fmaxf(auditMax, __shfl_xor_sync(FULL, auditMax, o));
            if (lane == 0) atomicMax(audit, (unsigned long long)__float_as_uint(auditMax));
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)ST_TMEM_COLS) : "memory");
    }
}

constexpr size_t ST_SMEM = 1024 + ST_ROWS * ST_ROW_BYTES + ST_STAGES * ST_STAGE_BYTES + 12 * ST_ROWS * 4 +
                           2 * ST_ROWS * 4 + ST_EPI_WARPS * (ST_QCAP + 64) * 4 + (2 * ST_STAGES + 4) * 8 + 16;

template <int MODEL>
int run(Context* ctx, cudaStream_t st, const float4* pf, const double* pd, const int64_t* dOff, const int64_t* hOff,
        int64_t pairBase, int np, const double* dModels, const int* dSlotList, const int* dNTotal, int slots,
        float thr2, int* dCounts, float* auditOut, const int* gate, const int* dBegin, bool haveFeatures) {
    // buffers of the pairs [0, np) of this chunk (dOff / hOff already point at the chunk's first pair), addressed relative
    // to the chunk's first pair
    const int64_t s0 = stage_base(hOff[0], pairBase), s1 = stage_base(hOff[np], pairBase + np) + 2;
    const size_t rawBytes = MODEL == RM_ESSENTIAL ? 32 : 16;
    if (ctx->ws[80].reserve((size_t)(s1 - s0) * ST_STAGE_BYTES) || ctx->ws[81].reserve((size_t)np * sizeof(PairNorm)) ||
        ctx->ws[82].reserve((size_t)(s1 - s0) * ST_NT * rawBytes) || ctx->ws[83].reserve((size_t)(s1 - s0) * sizeof(StageBox)) ||
        ctx->ws[84].reserve(64))
        return -1;
    float* feat = reinterpret_cast<float*>(ctx->ws[80].as<uint8_t>() - s0 * ST_STAGE_BYTES);
    void* raw = ctx->ws[82].as<uint8_t>() - s0 * ST_NT * rawBytes;
    StageBox* boxes = ctx->ws[83].as<StageBox>() - s0;
    PairNorm* norms = ctx->ws[81].as<PairNorm>();
    unsigned long long* audit = ctx->ws[84].as<unsigned long long>();
    if (!haveFeatures) {
        score_tc_features_kernel<MODEL><<<np, 256, 0, st>>>(pf, pd, dOff, pairBase, norms, feat, raw, boxes, gate);
        MCV_LAUNCH_CHECK();
    }
    dim3 g((slots + ST_ROWS - 1) / ST_ROWS, np);
    if (auditOut) {
        MCV_CUDA_TRY(cudaMemsetAsync(audit, 0, 8, st));
        auto kern = score_tc_kernel<MODEL, true>;
        MCV_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ST_SMEM));
        kern<<<g, ST_THREADS, ST_SMEM, st>>>(dOff, pairBase, norms, feat, raw, boxes, dModels, dSlotList, dNTotal, slots, thr2, dCounts, audit, gate, dBegin);
        MCV_LAUNCH_CHECK();
        unsigned long long h = 0;
        MCV_CUDA_TRY(cudaMemcpyAsync(&h, audit, 8, cudaMemcpyDeviceToHost, st));
        MCV_CUDA_TRY(cudaStreamSynchronize(st));
        const uint32_t bits = (uint32_t)h;
        float f;
        memcpy(&f, &bits, 4);
        *auditOut = std::max(*auditOut, f);
    } else {
        auto kern = score_tc_kernel<MODEL, false>;
        MCV_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ST_SMEM));
        kern<<<g, ST_THREADS, ST_SMEM, st>>>(dOff, pairBase, norms, feat, raw, boxes, dModels, dSlotList, dNTotal, slots, thr2, dCounts, audit, gate, dBegin);
        MCV_LAUNCH_CHECK();
    }
    return 0;
}

}  // namespace

std::atomic<int> g_score_audit{0};
std::atomic<uint32_t> g_score_audit_max{0};

namespace {

// 32 models x 128 correspondences per pair through the exact test: how many of the sampled models fit (keep >= 1/8)?
template <int MODEL>
__global__ void __launch_bounds__(32)
score_pilot_kernel(const float4* __restrict__ pf, const double* __restrict__ pd, const int64_t* __restrict__ off,
                   const double* __restrict__ models, const int* __restrict__ slotList, const int* __restrict__ nTotal,
                   int maxSlots, float thr2, int* __restrict__ gate) {
    const int p = blockIdx.x, lane = threadIdx.x;
    const int total = nTotal[p];
    const int64_t base = off[p];
    const int n = (int)min((int64_t)128, off[p + 1] - base);
    const bool active = lane < total && n >= 16;
    double m[9];
#pragma unroll
    for (int j = 0; j < 9; ++j) m[j] = 0;
    if (active) {
        const double* src = models + ((int64_t)p * maxSlots + slotList[(int64_t)p * maxSlots + lane]) * 9;
#pragma unroll
        for (int j = 0; j < 9; ++j) m[j] = src[j];
    }
    float Hf[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) Hf[j] = (float)m[j];
    int cnt = 0;
    if (active) {
        for (int j = 0; j < n; ++j) {
            if (MODEL == RM_ESSENTIAL) {
                const double4 c = reinterpret_cast<const double4*>(pd)[base + j];
                cnt += e_inlier(m, c.x, c.y, c.z, c.w, thr2) ? 1 : 0;
            } else {
                const float4 c = pf[base + j];
                if (MODEL == RM_HOMOGRAPHY) cnt += h_error(Hf, c.x, c.y, c.z, c.w) <= thr2 ? 1 : 0;
                else cnt += f_inlier(m, (double)c.x, (double)c.y, (double)c.z, (double)c.w, thr2) ? 1 : 0;
            }
        }
    }
    const unsigned looked = __ballot_sync(0xFFFFFFFFu, active), fit = __ballot_sync(0xFFFFFFFFu, active && cnt * 8 >= n);
    if (lane == 0 && looked) { atomicAdd(gate, __popc(fit)); atomicAdd(gate + 1, __popc(looked)); }
}

}  // namespace

int score_tc_mode(int model, int64_t maxN, int slots, int np, int smCount) {
    const int eng = g_score_engine.load(std::memory_order_relaxed);
    if (eng == 1) return 0;
    if (maxN >= (1 << 26) || maxN < 1) return 0;
    if (eng == 2) return 1;
    // auto: enough correspondences per pair to amortise the per-CTA model conversion, and enough (pair, 256-model) work
    // items: for H enough to fill the machine twice (below that the fp32-filter kernel, which can also split a pair's
    // correspondences over blocks, is as fast or faster: tools/gate_probe.py at 1 .. 8 pairs), then the device decides
    // (ransac.cuh); for F / E 64 items are enough -- at 2 pairs x 44 items the tensor engine costs 0.09 instead of 0.07 ms
    // at 30 % inliers and 0.18 instead of 0.49 ms at 70 %, where the other kernel's dense lanes take the fp64 exact test
    const int64_t items = (int64_t)np * ((slots + ST_ROWS - 1) / ST_ROWS);
    if (maxN < 256 || slots < 256) return 0;
    if (model == RM_HOMOGRAPHY) return items >= 2ll * smCount ? 2 : 0;
    return items >= 64 ? 1 : 0;
}

int* score_tc_pilot(Context* ctx, cudaStream_t st, int model, const float4* pf, const double* pd, const int64_t* dOff, int np,
                    const double* dModels, const int* dSlotList, const int* dNTotal, int slots, float thr2) {
    if (ctx->ws[85].reserve(64)) return nullptr;
    int* gate = ctx->ws[85].as<int>();
    if (cudaMemsetAsync(gate, 0, 8, st) != cudaSuccess) { set_error("cudaMemsetAsync failed"); return nullptr; }
    if (model == RM_HOMOGRAPHY)
        score_pilot_kernel<RM_HOMOGRAPHY><<<np, 32, 0, st>>>(pf, pd, dOff, dModels, dSlotList, dNTotal, slots, thr2, gate);
    else if (model == RM_FUNDAMENTAL)
        score_pilot_kernel<RM_FUNDAMENTAL><<<np, 32, 0, st>>>(pf, pd, dOff, dModels, dSlotList, dNTotal, slots, thr2, gate);
    else
        score_pilot_kernel<RM_ESSENTIAL><<<np, 32, 0, st>>>(pf, pd, dOff, dModels, dSlotList, dNTotal, slots, thr2, gate);
    count_launch();
    if (cudaGetLastError() != cudaSuccess) { set_error("score_pilot_kernel launch failed"); return nullptr; }
    return gate;
}

int score_tc_run(Context* ctx, cudaStream_t st, int model, const float4* pf, const double* pd, const int64_t* dOff,
                 const int64_t* hOff, int64_t pairBase, int np, const double* dModels, const int* dSlotList,
                 const int* dNTotal, int slots, float thr2, int* dCounts, const int* gate, const int* dBegin,
                 bool haveFeatures) {
    float am = 0.f;
    float* ap = g_score_audit.load(std::memory_order_relaxed) ? &am : nullptr;
    int rc;
    if (model == RM_HOMOGRAPHY)
        rc = run<RM_HOMOGRAPHY>(ctx, st, pf, pd, dOff, hOff, pairBase, np, dModels, dSlotList, dNTotal, slots, thr2, dCounts, ap, gate, dBegin, haveFeatures);
    else if (model == RM_FUNDAMENTAL)
        rc = run<RM_FUNDAMENTAL>(ctx, st, pf, pd, dOff, hOff, pairBase, np, dModels, dSlotList, dNTotal, slots, thr2, dCounts, ap, gate, dBegin, haveFeatures);
    else
        rc = run<RM_ESSENTIAL>(ctx, st, pf, pd, dOff, hOff, pairBase, np, dModels, dSlotList, dNTotal, slots, thr2, dCounts, ap, gate, dBegin, haveFeatures);
    if (ap) {
        uint32_t bits;
        memcpy(&bits, &am, 4);
        uint32_t cur = g_score_audit_max.load();
        while (bits > cur && !g_score_audit_max.compare_exchange_weak(cur, bits)) {}
    }
    return rc;
}

}  // namespace mcv

/* Scoring engine of the F / E RANSAC stage: 0 = auto (default), 1 = fp32-filter kernel, 2 = tcgen05 level-0 filter. */
MCV_API int mcvSetScoreEngine(int engine) {
    if (engine < 0 || engine > 2) return -1;
    mcv::g_score_engine.store(engine);
    return 0;
}
/* Audit mode (tests): while on, the tensor-core scoring kernel also recomputes every level-0 value in fp64 and records
 * the largest |r~ - r| / e_m it saw (e_m = the arithmetic error budget of filter0.cuh); mcvScoreAuditMax returns it. */
MCV_API int mcvSetScoreAudit(int on) {
    mcv::g_score_audit.store(on ? 1 : 0);
    if (on) mcv::g_score_audit_max.store(0);
    return 0;
}
MCV_API float mcvScoreAuditMax(void) {
    const uint32_t bits = mcv::g_score_audit_max.load();
    float f;
    memcpy(&f, &bits, 4);
    return f;
}
