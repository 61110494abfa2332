// pnp.cu -- batched PnP RANSAC with the AP3P minimal solver and the EPnP refit: the GPU form of
// cvSolvePnPRansac (reference src/MiniCVNative/MiniCVNative.cpp:93-139 -> cv::solvePnPRansac) for
// solverKind 5 (SOLVEPNP_AP3P), and of the reference's own solveAp3p export (ap3p.cpp:282-317).
// "solvePnPRansac on the same batched-RANSAC engine".
//
// Same staging as ransac.cu (whose compact / select stages it shares), with 3-D -> 2-D correspondences:
//   prep      world / image points -> float32 (cv::solvePnPRansac converts float64 input to float32)
//   sample    OpenCV's MWC stream (4 distinct indices per iteration, no degeneracy check for PnP), a
//             counter-based hash RNG (thread per hypothesis) or injected index sets
//   solve     one hypothesis per thread: undistort 4 points, AP3P on three, the fourth picks the pose,
//             Rodrigues -> (rvec, tvec), exactly the 3x2 model OpenCV's callback stores
//   compact   ransac_compact
//   score     one model per thread x all correspondences through shared memory: Rodrigues back to R once,
//             then per point the projectPoints arithmetic in float64, rounded to float32, squared float32
//             residual against the threshold (bit-exact restatement, no FMA contraction)   [FP64 pipe]
//   select    ransac_select (OpenCV's accept rule + RANSACUpdateNumIters replayed)
//   finalize  one CTA per problem: mask + ordered inlier list of the winner, then EPnP over the inliers
//             (block-wide reductions for the sums over points, thread 0 for the small dense algebra)
#include "ransac.cuh"
#include "filter_pass.cuh"
#include "pnpmath.cuh"

#include <algorithm>

namespace mcv {
namespace {

using namespace cvm;

constexpr int MP = 4;       // AP3P inside solvePnPRansac: model_points = 4
constexpr int MODEL_D = 6;  // rvec, tvec

__device__ __forceinline__ PnpCamera load_cam(const double* cam) {
    PnpCamera c;
    c.fx = cam[0]; c.fy = cam[1]; c.cx = cam[2]; c.cy = cam[3];
    c.k1 = cam[4]; c.k2 = cam[5]; c.p1 = cam[6]; c.p2 = cam[7];
    return c;
}

// ---------------------------------------------------------------------------------------------
// prep
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
pnp_prep_kernel(const double* __restrict__ obj, const double* __restrict__ img, int64_t total,
                float4* __restrict__ of, float2* __restrict__ imf) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    of[i] = make_float4((float)obj[3 * i], (float)obj[3 * i + 1], (float)obj[3 * i + 2], 0.f);
    imf[i] = make_float2((float)img[2 * i], (float)img[2 * i + 1]);
}

// ---------------------------------------------------------------------------------------------
// sampling
// ---------------------------------------------------------------------------------------------
// OpenCV's stream: RNG((uint64)-1) per call; PnPRansacCallback has no checkSubset, so an iteration is
// four draws plus the redraws of duplicates.  The stream is sequential by construction: one thread per
// problem (a few thousand multiply-with-carry steps).
__global__ void __launch_bounds__(64)
pnp_sample_mwc_kernel(const int64_t* __restrict__ off, int nProblems, int maxIters, int* __restrict__ samples,
                      int* __restrict__ nIter, int* __restrict__ flags) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= nProblems) return;
    const int n = (int)(off[p + 1] - off[p]);
    int* out = samples + (int64_t)p * maxIters * MP;
    if (n < MP) { nIter[p] = 0; flags[p] = PAIR_FAIL; return; }
    if (n == MP) {
        for (int i = 0; i < MP; ++i) out[i] = i;
        nIter[p] = 1; flags[p] = PAIR_DIRECT; return;
    }
    Mwc rng;
    for (int it = 0; it < maxIters; ++it) {
        int idx[MP];
        for (int i = 0; i < MP; ++i) {
            int v;
            bool dup;
            do {
                v = rng.uniform(n);
                dup = false;
                for (int k = 0; k < i; ++k) dup |= (idx[k] == v);
            } while (dup);
            idx[i] = v;
        }
        for (int i = 0; i < MP; ++i) out[it * MP + i] = idx[i];
    }
    nIter[p] = maxIters;
    flags[p] = 0;
}

__global__ void __launch_bounds__(128)
pnp_sample_counter_kernel(const int64_t* __restrict__ off, int maxIters, uint64_t seed, int64_t problemBase,
                          int* __restrict__ samples, int* __restrict__ nIter, int* __restrict__ flags) {
    const int it = blockIdx.x * blockDim.x + threadIdx.x;
    const int p = blockIdx.y;
    if (it >= maxIters) return;
    const int n = (int)(off[p + 1] - off[p]);
    int* out = samples + ((int64_t)p * maxIters + it) * MP;
    if (n < MP) { if (it == 0) { nIter[p] = 0; flags[p] = PAIR_FAIL; } return; }
    if (n == MP) {
        if (it == 0) { for (int i = 0; i < MP; ++i) out[i] = i; nIter[p] = 1; flags[p] = PAIR_DIRECT; }
        return;
    }
    if (it == 0) { nIter[p] = maxIters; flags[p] = 0; }
    uint64_t ctr = mix64(seed ^ mix64((uint64_t)(problemBase + p) * 0x100000001B3ull + (uint64_t)it));
    int idx[MP];
    for (int i = 0; i < MP; ++i) {
        int v;
        bool dup;
        do {
            ctr = mix64(ctr);
            v = (int)((uint32_t)(ctr >> 32) % (uint32_t)n);
            dup = false;
            for (int k = 0; k < i; ++k) dup |= (idx[k] == v);
        } while (dup);
        idx[i] = v;
    }
    for (int i = 0; i < MP; ++i) out[i] = idx[i];
}

__global__ void __launch_bounds__(256)
pnp_injected_setup_kernel(const int64_t* __restrict__ off, int nProblems, int maxIters, int* __restrict__ nIter,
                          int* __restrict__ flags) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= nProblems) return;
    const int n = (int)(off[p + 1] - off[p]);
    if (n < MP) { nIter[p] = 0; flags[p] = PAIR_FAIL; }
    else { nIter[p] = maxIters; flags[p] = 0; }
}

// ---------------------------------------------------------------------------------------------
// solve: one hypothesis per thread
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(64)
pnp_solve_kernel(const float4* __restrict__ of, const float2* __restrict__ imf, const int64_t* __restrict__ off,
                 const double* __restrict__ cams, int maxIters, const int* __restrict__ samples,
                 const int* __restrict__ nIter, double* __restrict__ models, uint8_t* __restrict__ nModels) {
    const int it = blockIdx.x * blockDim.x + threadIdx.x;
    const int p = blockIdx.y;
    if (it >= maxIters) return;
    uint8_t* nm = nModels + (int64_t)p * maxIters + it;
    if (it >= nIter[p]) { *nm = 0; return; }
    const int64_t base = off[p];
    const int n = (int)(off[p + 1] - base);
    const int* idx = samples + ((int64_t)p * maxIters + it) * MP;
    float o[12], im[8];
    bool bad = false;
    for (int i = 0; i < MP; ++i) bad |= (idx[i] < 0 || idx[i] >= n);
    if (bad) { *nm = 0; return; }
    for (int i = 0; i < MP; ++i) {
        const float4 q = of[base + idx[i]];
        const float2 m = imf[base + idx[i]];
        o[3 * i] = q.x; o[3 * i + 1] = q.y; o[3 * i + 2] = q.z;
        im[2 * i] = m.x; im[2 * i + 1] = m.y;
    }
    const PnpCamera cam = load_cam(cams + 8 * p);
    double rv[3], tv[3];
    bool ok = pnp_solve4(cam, o, im, rv, tv);
    // a pose with non-finite entries can never collect an inlier; drop it here
    for (int k = 0; k < 3; ++k) ok = ok && isfinite(rv[k]) && isfinite(tv[k]);
    if (ok) {
        double* out = models + ((int64_t)p * maxIters + it) * MODEL_D;
        for (int k = 0; k < 3; ++k) { out[k] = rv[k]; out[3 + k] = tv[k]; }
    }
    *nm = ok ? 1 : 0;
}

__global__ void __launch_bounds__(256)
pnp_inject_models_kernel(const double* __restrict__ inj, int64_t nHyp, double* __restrict__ models,
                         uint8_t* __restrict__ nModels) {
    const int64_t h = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (h >= nHyp) return;
    for (int k = 0; k < MODEL_D; ++k) models[h * MODEL_D + k] = inj[h * MODEL_D + k];
    nModels[h] = 1;
}

// ---------------------------------------------------------------------------------------------
// score: one model per thread, correspondences through shared memory
// ---------------------------------------------------------------------------------------------
constexpr int SCORE_TP = 512;     // points per stage: 512 x (16 + 8) B = 12 KB

__global__ void __launch_bounds__(128)
pnp_score_kernel(const float4* __restrict__ of, const float2* __restrict__ imf, const int64_t* __restrict__ off,
                 const double* __restrict__ cams, const double* __restrict__ models,
                 const int* __restrict__ slotList, const int* __restrict__ nTotal, int maxSlots, float thr2,
                 int chunkPts, int* __restrict__ counts) {
    __shared__ float4 sO[SCORE_TP];
    __shared__ float2 sI[SCORE_TP];
    __shared__ float2 sC[SCORE_TP];      // (u - cx, v - cy) in float32, for the filter
    __shared__ int smax[7];
    const int p = blockIdx.y;
    const int total = nTotal[p];
    const int k = blockIdx.x * 128 + threadIdx.x;
    if (blockIdx.x * 128 >= total) return;
    const bool active = k < total;
    const int64_t base = off[p];
    const int n = (int)(off[p + 1] - base);
    const int i0 = blockIdx.z * chunkPts;
    const int i1 = min(n, i0 + chunkPts);
    const PnpCamera cam = load_cam(cams + 8 * p);
    const bool dist = cam.distorted();

    double R[9], t[3] = {0, 0, 0};
    if (active) {
        const int slot = slotList[(int64_t)p * maxSlots + k];
        const double* src = models + ((int64_t)p * maxSlots + slot) * MODEL_D;
        const double rv[3] = {src[0], src[1], src[2]};
        rodrigues_v2m(rv, R);
        t[0] = src[3]; t[1] = src[4]; t[2] = src[5];
    } else {
#pragma unroll
        for (int j = 0; j < 9; ++j) R[j] = 0;
    }
    PnpFilter pfil;
    pnp_filter_init(R, t, cam, dist, pfil);
    int cnt = 0;
    bool skipFilter = false;
    for (int s0 = i0; s0 < i1; s0 += SCORE_TP) {
        const int rows = min(SCORE_TP, i1 - s0);
        // tile + the maxima the filter's error bounds need (bit patterns of non-negative floats order like integers;
        // a NaN sorts above everything and switches the filter off for the tile)
        if (threadIdx.x < 7) smax[threadIdx.x] = 0;
        __syncthreads();
        {
            int mx[7] = {0, 0, 0, 0, 0, 0, 0};
            for (int j = threadIdx.x; j < rows; j += 128) {
                const float4 q = of[base + s0 + j];
                const float2 m = imf[base + s0 + j];
                const float2 c = make_float2((float)((double)m.x - cam.cx), (float)((double)m.y - cam.cy));
                sO[j] = q; sI[j] = m; sC[j] = c;
                const float v[7] = {q.x, q.y, q.z, c.x, c.y, m.x, m.y};
#pragma unroll
                for (int e = 0; e < 7; ++e) mx[e] = max(mx[e], __float_as_int(fabsf(v[e])));
            }
#pragma unroll
            for (int e = 0; e < 7; ++e) {
                const int r = __reduce_max_sync(0xFFFFFFFFu, mx[e]);
                if ((threadIdx.x & 31) == 0) atomicMax(&smax[e], r);
            }
        }
        __syncthreads();
        // filtered exact scoring: fp32 filter over 64 correspondences (pnpmath.cuh, pnp_certain_outlier), the survivors
        // collected in a bit mask and given the exact fp64 projection afterwards
        PnpFilterTile ft;
        pnp_filter_tile(pfil, __int_as_float(smax[0]), __int_as_float(smax[1]), __int_as_float(smax[2]),
                        __int_as_float(smax[3]), __int_as_float(smax[4]), __int_as_float(smax[5]),
                        __int_as_float(smax[6]), thr2, ft);
        struct MP { double R[9], t[3]; };
        MP mine;
#pragma unroll
        for (int j = 0; j < 9; ++j) mine.R[j] = R[j];
#pragma unroll
        for (int j = 0; j < 3; ++j) mine.t[j] = t[j];
        auto shfl = [](const MP& a, int src) {
            MP r;
#pragma unroll
            for (int j = 0; j < 9; ++j) r.R[j] = __shfl_sync(0xFFFFFFFFu, a.R[j], src);
#pragma unroll
            for (int j = 0; j < 3; ++j) r.t[j] = __shfl_sync(0xFFFFFFFFu, a.t[j], src);
            return r;
        };
        for (int j0 = 0; j0 < rows; j0 += 64) {
            const int jn = min(64, rows - j0);
            auto test = [&](const MP& mm, int j) {
                const float4 q = sO[j0 + j];
                const float2 m = sI[j0 + j];
                return project_error(mm.R, mm.t, cam, dist, q.x, q.y, q.z, m.x, m.y) <= thr2;
            };
            unsigned long long pending = 0;
            if (skipFilter) {
            } else if (jn == 64) {
#pragma unroll 1
                for (int o = 0; o < 4; ++o) {
                    unsigned w = 0;
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        const float4 q = sO[j0 + 16 * o + j];
                        const float2 c = sC[j0 + 16 * o + j];
                        if (!pnp_certain_outlier(pfil, ft, q.x, q.y, q.z, c.x, c.y)) w |= 1u << j;
                    }
                    pending |= (unsigned long long)w << (16 * o);
                }
            } else {
                for (int j = 0; j < jn; ++j) {
                    const float4 q = sO[j0 + j];
                    const float2 c = sC[j0 + j];
                    if (!pnp_certain_outlier(pfil, ft, q.x, q.y, q.z, c.x, c.y)) pending |= 1ull << j;
                }
            }
            if (!active) pending = 0;
            if (!skipFilter) skipFilter = warp_too_dense(pending);
            if (skipFilter) {
#pragma unroll 2
                for (int j = 0; j < jn; ++j) cnt += test(mine, j) ? 1 : 0;
            } else {
                cnt += resolve_pending(pending, mine, shfl, test);
            }
        }
        __syncthreads();
    }
    if (active) {
        int* c = counts + (int64_t)p * maxSlots + k;
        if (gridDim.z == 1) *c = cnt; else atomicAdd(c, cnt);
    }
}

// ---------------------------------------------------------------------------------------------
// finalize
// ---------------------------------------------------------------------------------------------
constexpr int FB = 64;      // finalize block size: small blocks keep more problems resident (the serial thread-0 phases of EPnP dominate)

template <int K>
__device__ __forceinline__ void block_sum(double (&v)[K], double* sred /*[K * FB/32]*/, double* sout /*[K]*/) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < K; ++k) {
        double x = v[k];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xFFFFFFFFu, x, o);
        if (lane == 0) sred[warp * K + k] = x;
    }
    __syncthreads();
    if (threadIdx.x < K) {
        double s = 0;
        for (int w = 0; w < FB / 32; ++w) s += sred[w * K + threadIdx.x];
        sout[threadIdx.x] = s;
    }
    __syncthreads();
}

__global__ void __launch_bounds__(FB, 7)
pnp_finalize_kernel(const float4* __restrict__ of, const float2* __restrict__ imf, const int64_t* __restrict__ off,
                    const double* __restrict__ cams, const double* __restrict__ models, int maxSlots,
                    const int* __restrict__ best, const int* __restrict__ flags, float thr2, int refine,
                    int* __restrict__ inlierScratch, double* __restrict__ outR, double* __restrict__ outT,
                    uint8_t* __restrict__ masks, int* __restrict__ inlierIdxOut, int* __restrict__ inlierCounts,
                    uint8_t* __restrict__ ok, int* __restrict__ itersRun, int* __restrict__ bestIter,
                    double* __restrict__ minimalOut) {
    const int p = blockIdx.x;
    const int tid = threadIdx.x;
    const int64_t base = off[p];
    const int n = (int)(off[p + 1] - base);
    const int* b = best + 4 * p;
    const int fl = flags[p];
    const bool direct = (fl & PAIR_DIRECT) != 0;
    const int slot = b[3];
    if (b[0] < 0 || slot < 0) {
        for (int i = tid; i < n; i += FB) masks[base + i] = 0;
        if (tid == 0) {
            ok[p] = 0;
            if (inlierCounts) inlierCounts[p] = 0;
            if (itersRun) itersRun[p] = b[2];
            if (bestIter) bestIter[p] = -1;
        }
        return;
    }
    __shared__ double sred[40 * (FB / 32)];
    __shared__ double sout[40];
    __shared__ double sR[9], sT[3];
    __shared__ double sCws[12], sCi[9], sCcs[36], sRt[36], sPc0[9], sPw0[3];
    __shared__ int sScan[FB / 32 + 1];
    __shared__ int sBase;

    const PnpCamera cam = load_cam(cams + 8 * p);
    const bool dist = cam.distorted();
    const double* mdl = models + ((int64_t)p * maxSlots + slot) * MODEL_D;
    if (tid == 0) {
        const double rv[3] = {mdl[0], mdl[1], mdl[2]};
        double R[9];
        rodrigues_v2m(rv, R);
        for (int k = 0; k < 9; ++k) sR[k] = R[k];
        sT[0] = mdl[3]; sT[1] = mdl[4]; sT[2] = mdl[5];
        sBase = 0;
    }
    __syncthreads();

    // ---- mask of the winner + ordered inlier list ----
    int* list = inlierScratch + base;
    {
        double R[9], t[3];
        for (int k = 0; k < 9; ++k) R[k] = sR[k];
        for (int k = 0; k < 3; ++k) t[k] = sT[k];
        const int lane = tid & 31, warp = tid >> 5;
        for (int i0 = 0; i0 < n; i0 += FB) {
            const int i = i0 + tid;
            bool in = false;
            if (i < n) {
                if (direct) in = true;          // N == 4: OpenCV returns every point as an inlier
                else {
                    const float4 q = of[base + i];
                    const float2 m = imf[base + i];
                    in = project_error(R, t, cam, dist, q.x, q.y, q.z, m.x, m.y) <= thr2;
                }
                masks[base + i] = in ? 1 : 0;
            }
            const unsigned bal = __ballot_sync(0xFFFFFFFFu, in);
            if (lane == 0) sScan[warp] = __popc(bal);
            __syncthreads();
            int before = sBase;
            for (int w = 0; w < warp; ++w) before += sScan[w];
            if (in) list[before + __popc(bal & ((1u << lane) - 1))] = i;
            __syncthreads();
            if (tid == 0) { int s = 0; for (int w = 0; w < FB / 32; ++w) s += sScan[w]; sBase += s; }
            __syncthreads();
        }
    }
    const int ni = sBase;
    if (inlierIdxOut)
        for (int i = tid; i < ni; i += FB) inlierIdxOut[base + i] = list[i];
    if (tid == 0) {
        ok[p] = 1;
        if (inlierCounts) inlierCounts[p] = ni;
        if (itersRun) itersRun[p] = b[2];
        if (bestIter) bestIter[p] = slot;
        if (minimalOut) for (int k = 0; k < 6; ++k) minimalOut[6 * p + k] = mdl[k];
    }
    if (!refine || direct || ni < 4) {
        if (tid == 0) for (int k = 0; k < 3; ++k) { outR[3 * p + k] = mdl[k]; outT[3 * p + k] = mdl[3 + k]; }
        return;
    }

    // ---- EPnP over the inliers (solvePnP(..., SOLVEPNP_EPNP) on the float32 values promoted to double) ----
    const double fu = cam.fx, fv = cam.fy, uc = cam.cx, vc = cam.cy;
    auto world = [&](int j, double* pw) {
        const float4 q = of[base + list[j]];
        pw[0] = q.x; pw[1] = q.y; pw[2] = q.z;
    };
    auto pixel = [&](int j, double* u, double* v) {      // us = undistorted * f + c
        const float2 m = imf[base + list[j]];
        double x, y;
        undistort_point(cam, (double)m.x, (double)m.y, &x, &y);
        *u = x * fu + uc; *v = y * fv + vc;
    };
    // centroid
    {
        double v[3] = {0, 0, 0};
        for (int j = tid; j < ni; j += FB) { double pw[3]; world(j, pw); v[0] += pw[0]; v[1] += pw[1]; v[2] += pw[2]; }
        block_sum<3>(v, sred, sout);
        if (tid == 0) for (int k = 0; k < 3; ++k) { sCws[k] = sout[k] / ni; sPw0[k] = sout[k] / ni; }
        __syncthreads();
    }
    // scatter matrix -> control points, inverse basis
    {
        double v[6] = {0, 0, 0, 0, 0, 0};
        const double c0 = sCws[0], c1 = sCws[1], c2 = sCws[2];
        for (int j = tid; j < ni; j += FB) {
            double pw[3]; world(j, pw);
            const double d0 = pw[0] - c0, d1 = pw[1] - c1, d2 = pw[2] - c2;
            v[0] += d0 * d0; v[1] += d0 * d1; v[2] += d0 * d2; v[3] += d1 * d1; v[4] += d1 * d2; v[5] += d2 * d2;
        }
        block_sum<6>(v, sred, sout);
        if (tid == 0) {
            const double cov[9] = {sout[0], sout[1], sout[2], sout[1], sout[3], sout[4], sout[2], sout[4], sout[5]};
            double cws[12], ci[9];
            for (int k = 0; k < 3; ++k) cws[k] = sCws[k];
            epnp_control_points(ni, cov, cws, ci);
            for (int k = 0; k < 12; ++k) sCws[k] = cws[k];
            for (int k = 0; k < 9; ++k) sCi[k] = ci[k];
        }
        __syncthreads();
    }
    double cws[12], ci[9];
    for (int k = 0; k < 12; ++k) cws[k] = sCws[k];
    for (int k = 0; k < 9; ++k) ci[k] = sCi[k];
    // M^T M
    {
        double acc[40];
#pragma unroll
        for (int k = 0; k < 40; ++k) acc[k] = 0;
        for (int j = tid; j < ni; j += FB) {
            double pw[3], a[4], u, v;
            world(j, pw);
            pixel(j, &u, &v);
            epnp_alphas(cws, ci, pw, a);
            epnp_accum(a, uc - u, vc - v, acc);
        }
        block_sum<40>(acc, sred, sout);
        if (tid == 0) {
            double a40[40], MtM[144], ut4[48], betas[12];
            for (int k = 0; k < 40; ++k) a40[k] = sout[k];
            epnp_mtm(a40, fu, fv, MtM);
            epnp_betas(MtM, cws, ut4, betas);
            // camera-frame control points per candidate, sign fixed by the first inlier's depth
            double pw[3], a[4];
            world(0, pw);
            epnp_alphas(cws, ci, pw, a);
            for (int c = 0; c < 3; ++c) {
                double ccs[12], pc[3];
                epnp_ccs(ut4, betas + 4 * c, ccs);
                epnp_pc(a, ccs, pc);
                const double sg = pc[2] < 0.0 ? -1.0 : 1.0;
                for (int k = 0; k < 12; ++k) sCcs[12 * c + k] = sg * ccs[k];
            }
        }
        __syncthreads();
    }
    // camera-frame centroids of the three candidates
    {
        double v[9];
#pragma unroll
        for (int k = 0; k < 9; ++k) v[k] = 0;
        for (int j = tid; j < ni; j += FB) {
            double pw[3], a[4], pc[3];
            world(j, pw);
            epnp_alphas(cws, ci, pw, a);
            for (int c = 0; c < 3; ++c) {
                epnp_pc(a, sCcs + 12 * c, pc);
                v[3 * c] += pc[0]; v[3 * c + 1] += pc[1]; v[3 * c + 2] += pc[2];
            }
        }
        block_sum<9>(v, sred, sout);
        if (tid < 9) sPc0[tid] = sout[tid] / ni;
        __syncthreads();
    }
    // correlation matrices -> R, t per candidate
    {
        double v[27];
#pragma unroll
        for (int k = 0; k < 27; ++k) v[k] = 0;
        for (int j = tid; j < ni; j += FB) {
            double pw[3], a[4], pc[3];
            world(j, pw);
            epnp_alphas(cws, ci, pw, a);
            const double w0 = pw[0] - sPw0[0], w1 = pw[1] - sPw0[1], w2 = pw[2] - sPw0[2];
            for (int c = 0; c < 3; ++c) {
                epnp_pc(a, sCcs + 12 * c, pc);
                for (int r = 0; r < 3; ++r) {
                    const double d = pc[r] - sPc0[3 * c + r];
                    v[9 * c + 3 * r] += d * w0; v[9 * c + 3 * r + 1] += d * w1; v[9 * c + 3 * r + 2] += d * w2;
                }
            }
        }
        block_sum<27>(v, sred, sout);
        if (tid < 3) {
            double abt[9], R[9], t[3], pc0[3], pw0[3];
            for (int k = 0; k < 9; ++k) abt[k] = sout[9 * tid + k];
            for (int k = 0; k < 3; ++k) { pc0[k] = sPc0[3 * tid + k]; pw0[k] = sPw0[k]; }
            epnp_rt(abt, pc0, pw0, R, t);
            for (int k = 0; k < 9; ++k) sRt[12 * tid + k] = R[k];
            for (int k = 0; k < 3; ++k) sRt[12 * tid + 9 + k] = t[k];
        }
        __syncthreads();
    }
    // reprojection error of the three candidates, the smallest wins (ties keep the earlier one)
    {
        double v[3] = {0, 0, 0};
        for (int j = tid; j < ni; j += FB) {
            double pw[3], pu, pv;
            world(j, pw);
            pixel(j, &pu, &pv);
            for (int c = 0; c < 3; ++c) v[c] += epnp_reproj(sRt + 12 * c, sRt + 12 * c + 9, pw, fu, fv, uc, vc, pu, pv);
        }
        block_sum<3>(v, sred, sout);
        if (tid == 0) {
            int N = 0;
            if (sout[1] < sout[0]) N = 1;
            if (sout[2] < sout[N]) N = 2;
            double R[9], rv[3];
            for (int k = 0; k < 9; ++k) R[k] = sRt[12 * N + k];
            rodrigues_m2v(R, rv);
            for (int k = 0; k < 3; ++k) { outR[3 * p + k] = rv[k]; outT[3 * p + k] = sRt[12 * N + 9 + k]; }
        }
    }
}

// debug taps: expand compacted counts back to [maxIters] (-1 = no pose)
__global__ void __launch_bounds__(256)
pnp_dbg_counts_kernel(const int* __restrict__ iterStart, const int* __restrict__ counts, int maxIters,
                      int* __restrict__ out) {
    const int p = blockIdx.y;
    const int it = blockIdx.x * blockDim.x + threadIdx.x;
    if (it >= maxIters) return;
    const int* is = iterStart + (int64_t)p * (maxIters + 1);
    const int k0 = is[it], k1 = is[it + 1];
    out[(int64_t)p * maxIters + it] = (k1 > k0) ? counts[(int64_t)p * maxIters + k0] : -1;
}

__global__ void __launch_bounds__(64)
ap3p_batch_kernel(const double* __restrict__ img, const double* __restrict__ obj, const double* __restrict__ cam4,
                  int n, double* __restrict__ Rs, double* __restrict__ ts, int* __restrict__ counts) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double* c = cam4 + 4 * i;
    double b[3][3], w[3][3];
    for (int j = 0; j < 3; ++j) {
        ap3p_bearing(img[6 * i + 2 * j], img[6 * i + 2 * j + 1], c[0], c[1], c[2], c[3], b[j], nullptr);
        for (int k = 0; k < 3; ++k) w[j][k] = obj[9 * i + 3 * j + k];
    }
    double R[36], t[12];
    const int m = ap3p_poses(b[0], b[1], b[2], w[0], w[1], w[2], false, nullptr, nullptr, R, t);
    for (int k = 0; k < m * 9; ++k) Rs[36 * (int64_t)i + k] = R[k];
    for (int k = 0; k < m * 3; ++k) ts[12 * (int64_t)i + k] = t[k];
    counts[i] = m;
}

}  // namespace

int ap3p_batch_device(Context* ctx, cudaStream_t stream, const double* dImg, const double* dObj, const double* dCam4,
                      int n, double* dRs, double* dTs, int* dCounts) {
    (void)ctx;
    if (n <= 0) return 0;
    ap3p_batch_kernel<<<(n + 63) / 64, 64, 0, stream>>>(dImg, dObj, dCam4, n, dRs, dTs, dCounts);
    MCV_LAUNCH_CHECK();
    return 0;
}

int pnp_ransac_batch_device(Context* ctx, cudaStream_t st, const PnpParams& prm, const PnpIO& io) {
    const int nProblems = io.nProblems;
    if (nProblems <= 0) return 0;
    if (!io.hOff || io.hOff[0] != 0) { set_error("pnp: offsets must start at 0"); return -1; }
    if (!io.dRvec || !io.dTvec || !io.dMasks || !io.dOk || !io.dCam) { set_error("pnp: null outputs / cameras"); return -1; }
    if (!(prm.confidence > 0 && prm.confidence < 1) && prm.stopMode == MCV_STOP_ADAPTIVE) {
        set_error("pnp: confidence must be in (0,1) (OpenCV asserts the same)");
        return -1;
    }
    if (prm.maxIters > (1 << 20)) { set_error("pnp: maxIters too large"); return -1; }
    const int maxIters = std::max(prm.maxIters, 1);
    const int64_t total = io.hOff[nProblems];
    const float thr2 = (float)(prm.threshold * prm.threshold);

    const size_t n1 = (size_t)nProblems + 1;
    if (ctx->pin[1].reserve(n1 * 8) || ctx->ws[48].reserve(n1 * 8)) return -1;
    MCV_CUDA_TRY(cudaEventSynchronize(ctx->ev[1]));
    memcpy(ctx->pin[1].p, io.hOff, n1 * 8);
    MCV_CUDA_TRY(cudaMemcpyAsync(ctx->ws[48].p, ctx->pin[1].p, n1 * 8, cudaMemcpyHostToDevice, st));
    MCV_CUDA_TRY(cudaEventRecord(ctx->ev[1], st));
    const int64_t* dOff = ctx->ws[48].as<int64_t>();

    const size_t tb = (size_t)std::max<int64_t>(total, 1);
    if (ctx->ws[49].reserve(tb * 16) || ctx->ws[50].reserve(tb * 8) || ctx->ws[51].reserve(tb * 4)) return -1;
    float4* of = ctx->ws[49].as<float4>();
    float2* imf = ctx->ws[50].as<float2>();
    int* dInlierScratch = ctx->ws[51].as<int>();
    if (total > 0) {
        pnp_prep_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(io.dObj, io.dImg, total, of, imf);
        MCV_LAUNCH_CHECK();
    }

    const size_t slots = (size_t)maxIters;
    const size_t perProblem = slots * (MODEL_D * 8 + MP * 4 + 16);
    const int chunk = std::min(65535,     // the problem index is gridDim.y of several stage kernels
                               (int)std::max<size_t>(1, std::min<size_t>((size_t)nProblems, ((size_t)1 << 30) / std::max<size_t>(perProblem, 1))));
    if (ctx->ws[52].reserve((size_t)chunk * slots * MP * 4) ||          // samples
        ctx->ws[53].reserve((size_t)chunk * 8 + 64) ||                  // nIter, flags
        ctx->ws[54].reserve((size_t)chunk * slots * MODEL_D * 8) ||     // models
        ctx->ws[55].reserve((size_t)chunk * slots) ||                   // nModels
        ctx->ws[56].reserve((size_t)chunk * (slots + 1) * 4) ||         // iterStart
        ctx->ws[57].reserve((size_t)chunk * slots * 4) ||               // slotList
        ctx->ws[58].reserve((size_t)chunk * slots * 4) ||               // counts
        ctx->ws[59].reserve((size_t)chunk * 4 + 64) ||                  // nTotal
        ctx->ws[60].reserve((size_t)chunk * 16))                        // best
        return -1;
    int* dSamples = ctx->ws[52].as<int>();
    int* dNIter = ctx->ws[53].as<int>();
    int* dFlags = dNIter + chunk;
    double* dModels = ctx->ws[54].as<double>();
    uint8_t* dNModels = ctx->ws[55].as<uint8_t>();
    int* dIterStart = ctx->ws[56].as<int>();
    int* dSlotList = ctx->ws[57].as<int>();
    int* dCounts = ctx->ws[58].as<int>();
    int* dNTotal = ctx->ws[59].as<int>();
    int* dBest = ctx->ws[60].as<int>();

    for (int p0 = 0; p0 < nProblems; p0 += chunk) {
        const int np = std::min(chunk, nProblems - p0);
        const int64_t* off = dOff + p0;
        const double* cams = io.dCam + (size_t)p0 * 8;
        int64_t maxN = 0;
        for (int p = p0; p < p0 + np; ++p) maxN = std::max(maxN, io.hOff[p + 1] - io.hOff[p]);

        // ---- sample
        if (io.dInjModels) {
            pnp_injected_setup_kernel<<<(np + 255) / 256, 256, 0, st>>>(off, np, maxIters, dNIter, dFlags);
            MCV_LAUNCH_CHECK();
        } else if (prm.sampleMode == MCV_SAMPLE_INJECTED) {
            if (!io.dInjSamples) { set_error("MCV_SAMPLE_INJECTED without sample indices"); return -1; }
            MCV_CUDA_TRY(cudaMemcpyAsync(dSamples, io.dInjSamples + (size_t)p0 * maxIters * MP,
                                         (size_t)np * maxIters * MP * 4, cudaMemcpyDeviceToDevice, st));
            pnp_injected_setup_kernel<<<(np + 255) / 256, 256, 0, st>>>(off, np, maxIters, dNIter, dFlags);
            MCV_LAUNCH_CHECK();
        } else if (prm.sampleMode == MCV_SAMPLE_COUNTER) {
            dim3 g((maxIters + 127) / 128, np);
            pnp_sample_counter_kernel<<<g, 128, 0, st>>>(off, maxIters, prm.seed, prm.problemIndexBase + p0, dSamples, dNIter, dFlags);
            MCV_LAUNCH_CHECK();
        } else {
            pnp_sample_mwc_kernel<<<(np + 63) / 64, 64, 0, st>>>(off, np, maxIters, dSamples, dNIter, dFlags);
            MCV_LAUNCH_CHECK();
        }
        // ---- solve
        if (io.dInjModels) {
            const int64_t nh = (int64_t)np * maxIters;
            pnp_inject_models_kernel<<<(unsigned)((nh + 255) / 256), 256, 0, st>>>(
                io.dInjModels + (size_t)p0 * maxIters * MODEL_D, nh, dModels, dNModels);
            MCV_LAUNCH_CHECK();
        } else {
            dim3 g((maxIters + 63) / 64, np);
            pnp_solve_kernel<<<g, 64, 0, st>>>(of, imf, off, cams, maxIters, dSamples, dNIter, dModels, dNModels);
            MCV_LAUNCH_CHECK();
        }
        // ---- compact
        if (ransac_compact(st, np, dNModels, dNIter, maxIters, 1, dIterStart, dSlotList, dNTotal)) return -1;
        // ---- score
        {
            const int tiles = (int)((slots + 127) / 128);
            int zs = 1;
            const int64_t blocksXY = (int64_t)tiles * np;
            const int64_t fill = 4ll * ctx->sm_count;
            if (blocksXY < fill) zs = (int)std::min<int64_t>((fill + blocksXY - 1) / blocksXY, std::max<int64_t>(1, (maxN + SCORE_TP - 1) / SCORE_TP));
            int chunkPts = (int)((std::max<int64_t>(maxN, 1) + zs - 1) / zs);
            zs = (int)((std::max<int64_t>(maxN, 1) + chunkPts - 1) / chunkPts);
            if (zs > 1) MCV_CUDA_TRY(cudaMemsetAsync(dCounts, 0, (size_t)np * slots * 4, st));
            dim3 g(tiles, np, zs);
            pnp_score_kernel<<<g, 128, 0, st>>>(of, imf, off, cams, dModels, dSlotList, dNTotal, (int)slots, thr2,
                                                chunkPts, dCounts);
            MCV_LAUNCH_CHECK();
        }
        // ---- select
        if (ransac_select(st, off, np, maxIters, 1, MP, prm.stopMode == MCV_STOP_ADAPTIVE, prm.confidence, dNIter,
                          dFlags, dIterStart, dSlotList, dCounts, dBest))
            return -1;
        // ---- finalize
        pnp_finalize_kernel<<<np, FB, 0, st>>>(
            of, imf, off, cams, dModels, (int)slots, dBest, dFlags, thr2, prm.refine, dInlierScratch,
            io.dRvec + 3 * (size_t)p0, io.dTvec + 3 * (size_t)p0, io.dMasks, io.dInlierIdx,
            io.dInlierCounts ? io.dInlierCounts + p0 : nullptr, io.dOk + p0, io.dItersRun ? io.dItersRun + p0 : nullptr,
            io.dBestIter ? io.dBestIter + p0 : nullptr, io.dMinimal ? io.dMinimal + 6 * (size_t)p0 : nullptr);
        MCV_LAUNCH_CHECK();
        // ---- debug taps
        if (io.dDbgCounts) {
            dim3 g((maxIters + 255) / 256, np);
            pnp_dbg_counts_kernel<<<g, 256, 0, st>>>(dIterStart, dCounts, maxIters, io.dDbgCounts + (size_t)p0 * slots);
            MCV_LAUNCH_CHECK();
        }
        if (io.dDbgModels)
            MCV_CUDA_TRY(cudaMemcpyAsync(io.dDbgModels + (size_t)p0 * slots * MODEL_D, dModels,
                                         (size_t)np * slots * MODEL_D * 8, cudaMemcpyDeviceToDevice, st));
        if (io.dDbgSamples && !io.dInjModels)
            MCV_CUDA_TRY(cudaMemcpyAsync(io.dDbgSamples + (size_t)p0 * maxIters * MP, dSamples,
                                         (size_t)np * maxIters * MP * 4, cudaMemcpyDeviceToDevice, st));
    }
    return 0;
}

}  // namespace mcv
