// scale_search.cu -- CameraPose.findScaled on the GPU (reference:
// src/MiniCV/CameraPose.fs:39-134 with Camera.transformedView / project1, src/MiniCV/Camera.fs:61-84).
//
// The reference walks the observations once; each proposes two candidate scales and every candidate is
// costed against ALL observations (mean squared reprojection error of the points that project): O(2 n^2)
// projections per problem, strictly sequential on one thread.  Here the three stages are data-parallel over
// (problem, observation) / (problem, candidate):
//   scale_candidates_kernel  per problem the zero-scale view + translation direction (thread 0), then one
//                            observation per thread -> its two candidate scales or "tiny" (CameraPose.fs:97-113)
//   scale_cost_kernel        one candidate per thread: builds the camera that scale produces in registers, then
//                            walks the observations (smem tiles, broadcast 128-bit loads) IN ORDER, so the sum
//                            is the reference's left-to-right sum bit for bit.  FP64-pipe bound: 29 fp64
//                            mul/add + 2 IEEE divisions per (candidate, observation); 40 B per observation read
//                            once per block of candidates.
//   scale_select_kernel      first strictly smallest cost in candidate order (CameraPose.fs:118-124) and the
//                            scaled pose (CameraPose.scale).
// Arithmetic is scalemath.cuh (separately rounded operations, no FMA contraction).
#include "ransac.cuh"
#include "scalemath.cuh"

#include <algorithm>
#include <math_constants.h>

namespace mcv {

using namespace scm;

namespace {

constexpr int CAND_THREADS = 256;
constexpr int COST_TILE = 256;        // observations per shared-memory tile
constexpr int SELECT_THREADS = 256;

__device__ __forceinline__ void load_cam(const double* c, Cam& cam) {
    for (int i = 0; i < 3; ++i) { cam.loc[i] = c[i]; cam.fwd[i] = c[3 + i]; cam.up[i] = c[6 + i]; cam.right[i] = c[9 + i]; }
    cam.focal[0] = c[12]; cam.focal[1] = c[13];
}

__global__ void __launch_bounds__(CAND_THREADS)
scale_candidates_kernel(const double* __restrict__ cams, const double* __restrict__ rot, const double* __restrict__ trans,
                        const double* __restrict__ world, const double* __restrict__ obs,
                        const int64_t* __restrict__ off, double* __restrict__ cand, uint8_t* __restrict__ candOk) {
    const int p = blockIdx.y;
    const int64_t o0 = off[p];
    const int n = (int)(off[p + 1] - o0);
    const int i0 = blockIdx.x * CAND_THREADS;
    if (i0 >= n) return;
    __shared__ ScaleSetup S;
    if (threadIdx.x == 0) {
        Cam src;
        load_cam(cams + 14 * (size_t)p, src);
        scale_setup(src, rot + 9 * (size_t)p, trans + 3 * (size_t)p, S);
    }
    __syncthreads();
    const int i = i0 + threadIdx.x;
    if (i >= n) return;
    const double w[3] = {world[3 * (o0 + i)], world[3 * (o0 + i) + 1], world[3 * (o0 + i) + 2]};
    const double ob[2] = {obs[2 * (o0 + i)], obs[2 * (o0 + i) + 1]};
    double sx = 0.0, sy = 0.0;
    const bool ok = candidate_scales(S, w, ob, sx, sy);
    cand[2 * (o0 + i)] = sx;
    cand[2 * (o0 + i) + 1] = sy;
    candOk[2 * (o0 + i)] = ok ? 1 : 0;
    candOk[2 * (o0 + i) + 1] = ok ? 1 : 0;
}

// One candidate scale per thread against every observation of its problem, in order.
__global__ void __launch_bounds__(128)
scale_cost_kernel(const double* __restrict__ cams, const double* __restrict__ rot, const double* __restrict__ trans,
                  const double* __restrict__ world, const double* __restrict__ obs, const int64_t* __restrict__ off,
                  const double* __restrict__ scales, const uint8_t* __restrict__ scaleOk,
                  const int64_t* __restrict__ scaleOff, double* __restrict__ costs) {
    const int p = blockIdx.y;
    const int64_t s0 = scaleOff[p];
    const int ns = (int)(scaleOff[p + 1] - s0);
    const int c0 = blockIdx.x * blockDim.x;
    if (c0 >= ns) return;
    const int64_t o0 = off[p];
    const int n = (int)(off[p + 1] - o0);

    __shared__ __align__(16) double tile[COST_TILE * 6];      // wx wy | wz ox | oy -- : three 128-bit loads

    const int c = c0 + threadIdx.x;
    const bool active = c < ns && (scaleOk == nullptr || scaleOk[s0 + c] != 0);
    Cam cam;
    if (active) {
        Cam src;
        load_cam(cams + 14 * (size_t)p, src);
        double basis[16];
        camera_basis(src, basis);
        camera_for_scale(basis, src.focal, rot + 9 * (size_t)p, trans + 3 * (size_t)p, scales[s0 + c], cam);
    } else {
        for (int i = 0; i < 3; ++i) { cam.loc[i] = 0.0; cam.fwd[i] = 0.0; cam.up[i] = 0.0; cam.right[i] = 0.0; }
        cam.focal[0] = cam.focal[1] = 0.0;
    }
    double sum = 0.0;
    int cnt = 0;
    for (int t0 = 0; t0 < n; t0 += COST_TILE) {
        const int tn = min(COST_TILE, n - t0);
        __syncthreads();
        for (int i = threadIdx.x; i < tn; i += blockDim.x) {
            const double* w = world + 3 * (o0 + t0 + i);
            const double* ob = obs + 2 * (o0 + t0 + i);
            tile[6 * i + 0] = w[0]; tile[6 * i + 1] = w[1]; tile[6 * i + 2] = w[2];
            tile[6 * i + 3] = ob[0]; tile[6 * i + 4] = ob[1];
        }
        __syncthreads();
        if (active) {
            // four observations at a time: the fast-path chains are independent and interleave; the generic
            // division (denormal / huge quotients) is decided once per group and is rare
            int i = 0;
            for (; i + 4 <= tn; i += 4) {
                double term[4];
                bool ok[4];
                bool exact = true;
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const double2 a = *reinterpret_cast<const double2*>(&tile[6 * (i + k)]);
                    const double2 b = *reinterpret_cast<const double2*>(&tile[6 * (i + k) + 2]);
                    exact &= project_term_t<true>(cam, a.x, a.y, b.x, b.y, tile[6 * (i + k) + 4], term[k], ok[k]);
                }
                if (!exact) {
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        ok[k] = project_term(cam, tile[6 * (i + k)], tile[6 * (i + k) + 1], tile[6 * (i + k) + 2],
                                             tile[6 * (i + k) + 3], tile[6 * (i + k) + 4], term[k]);
                }
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    if (ok[k]) { sum = dadd(sum, term[k]); ++cnt; }
            }
            for (; i < tn; ++i) {
                double term;
                if (project_term(cam, tile[6 * i], tile[6 * i + 1], tile[6 * i + 2], tile[6 * i + 3], tile[6 * i + 4], term)) {
                    sum = dadd(sum, term);
                    ++cnt;
                }
            }
        }
    }
    if (c < ns) costs[s0 + c] = (active && cnt > 0) ? ddiv(sum, (double)cnt) : CUDART_INF;
}

// first strictly smallest cost in candidate order; the scaled pose
__global__ void __launch_bounds__(SELECT_THREADS)
scale_select_kernel(const double* __restrict__ trans, const double* __restrict__ scales,
                    const int64_t* __restrict__ scaleOff, const double* __restrict__ costs,
                    double* __restrict__ outCost, double* __restrict__ outScale, double* __restrict__ outT,
                    int* __restrict__ outSign) {
    const int p = blockIdx.x;
    const int64_t s0 = scaleOff[p];
    const int ns = (int)(scaleOff[p + 1] - s0);
    double best = CUDART_INF;
    int bi = 0x7FFFFFFF;
    for (int c = threadIdx.x; c < ns; c += SELECT_THREADS) {
        const double v = costs[s0 + c];
        if (v < best) { best = v; bi = c; }             // per thread c ascends: strict < keeps the first
    }
    __shared__ double sb[SELECT_THREADS];
    __shared__ int si[SELECT_THREADS];
    sb[threadIdx.x] = best; si[threadIdx.x] = bi;
    __syncthreads();
    for (int s = SELECT_THREADS / 2; s > 0; s >>= 1) {
        if (threadIdx.x < s) {
            const double v = sb[threadIdx.x + s];
            const int j = si[threadIdx.x + s];
            if (v < sb[threadIdx.x] || (v == sb[threadIdx.x] && j < si[threadIdx.x])) { sb[threadIdx.x] = v; si[threadIdx.x] = j; }
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        const bool found = si[0] != 0x7FFFFFFF;
        const double s = found ? scales[s0 + si[0]] : 0.0;
        outCost[p] = found ? sb[0] : CUDART_INF;
        outScale[p] = s;
        for (int i = 0; i < 3; ++i) outT[3 * p + i] = dmul(s, trans[3 * (size_t)p + i]);
        outSign[p] = (s > 0.0 ? 1 : 0) - (s < 0.0 ? 1 : 0);
    }
}

__global__ void double_offsets_kernel(const int64_t* __restrict__ off, int n1, int64_t* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n1) out[i] = 2 * off[i];
}

}  // namespace

int scale_costs_device(Context* ctx, cudaStream_t st, const double* dCams, const double* dRot, const double* dTrans,
                       const double* dWorld, const double* dObs, const int64_t* dOff, int nProblems,
                       const double* dScales, const uint8_t* dScaleOk, const int64_t* dScaleOff, int64_t maxScales,
                       double* dCosts) {
    if (nProblems <= 0 || maxScales <= 0) return 0;
    if (nProblems > 65535) { set_error("scale search: at most 65535 problems per call (split the batch)"); return -1; }
    // the sum over the observations cannot be split (its order is the result), so parallelism is candidates x
    // problems only: narrow blocks when that does not fill the machine
    int threads = 128;
    while (threads > 32 && ((maxScales + threads - 1) / threads) * (int64_t)nProblems < 2ll * ctx->sm_count) threads >>= 1;
    dim3 g((unsigned)((maxScales + threads - 1) / threads), nProblems);
    kt_begin(ctx, st, KT_SCALE);
    scale_cost_kernel<<<g, threads, 0, st>>>(dCams, dRot, dTrans, dWorld, dObs, dOff, dScales, dScaleOk, dScaleOff, dCosts);
    kt_end(ctx, st);
    MCV_LAUNCH_CHECK();
    return 0;
}

int find_scaled_batch_device(Context* ctx, cudaStream_t st, const ScaleIO& io) {
    const int np = io.nProblems;
    if (np <= 0) return 0;
    if (!io.hOff || io.hOff[0] != 0) { set_error("findScaled: offsets must start at 0"); return -1; }
    if (!io.dCams || !io.dRot || !io.dTrans || !io.dCost || !io.dBestScale || !io.dScaledT || !io.dSign) {
        set_error("findScaled: null cameras / poses / outputs");
        return -1;
    }
    const int64_t total = io.hOff[np];
    int64_t maxN = 0;
    for (int p = 0; p < np; ++p) maxN = std::max(maxN, io.hOff[p + 1] - io.hOff[p]);
    const size_t n1 = (size_t)np + 1;
    const size_t tb = (size_t)std::max<int64_t>(total, 1);
    if (ctx->pin[4].reserve(n1 * 8) || ctx->ws[64].reserve(n1 * 16) || ctx->ws[65].reserve(tb * 2 * (8 + 8 + 1))) return -1;
    MCV_CUDA_TRY(cudaEventSynchronize(ctx->ev[5]));
    memcpy(ctx->pin[4].p, io.hOff, n1 * 8);
    int64_t* dOff = ctx->ws[64].as<int64_t>();
    int64_t* dOff2 = dOff + n1;
    MCV_CUDA_TRY(cudaMemcpyAsync(dOff, ctx->pin[4].p, n1 * 8, cudaMemcpyHostToDevice, st));
    MCV_CUDA_TRY(cudaEventRecord(ctx->ev[5], st));
    double_offsets_kernel<<<(unsigned)((n1 + 255) / 256), 256, 0, st>>>(dOff, (int)n1, dOff2);
    MCV_LAUNCH_CHECK();
    double* dCand = io.dCand ? io.dCand : ctx->ws[65].as<double>();
    double* dCandCost = io.dCandCost ? io.dCandCost : ctx->ws[65].as<double>() + 2 * tb;
    uint8_t* dCandOk = io.dCandOk ? io.dCandOk : reinterpret_cast<uint8_t*>(ctx->ws[65].as<double>() + 4 * tb);
    if (maxN > 0) {
        dim3 g((unsigned)((maxN + CAND_THREADS - 1) / CAND_THREADS), np);
        scale_candidates_kernel<<<g, CAND_THREADS, 0, st>>>(io.dCams, io.dRot, io.dTrans, io.dWorld, io.dObs, dOff,
                                                            dCand, dCandOk);
        MCV_LAUNCH_CHECK();
        if (scale_costs_device(ctx, st, io.dCams, io.dRot, io.dTrans, io.dWorld, io.dObs, dOff, np, dCand, dCandOk,
                               dOff2, 2 * maxN, dCandCost))
            return -1;
    }
    scale_select_kernel<<<np, SELECT_THREADS, 0, st>>>(io.dTrans, dCand, dOff2, dCandCost, io.dCost, io.dBestScale,
                                                       io.dScaledT, io.dSign);
    MCV_LAUNCH_CHECK();
    return 0;
}

}  // namespace mcv
