// pose.cu -- decomposeEssentialMat + recoverPose for a batch, and cvFivePoint on the device.
//
// Replaces the tail of the reference's exports: cvRecoverPoses -> decomposeEssentialMat
// (src/MiniCVNative/MiniCVNative.cpp:184-188), cvRecoverPose -> cv::recoverPose (:212), and
// cvFivePoint -> runFivepoint (:368-382, fivepoint.cpp:233-339).
#include "ransac.cuh"
#include "cvmath.cuh"

#include <algorithm>

namespace mcv {
namespace {

using namespace cvm;

__global__ void __launch_bounds__(64)
decompose_kernel(const double* __restrict__ E, const uint8_t* __restrict__ ok, int nPairs,
                 double* __restrict__ R1, double* __restrict__ R2, double* __restrict__ T) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= nPairs) return;
    double r1[9], r2[9], t[3], e[9];
    if (ok && !ok[p]) {
        for (int k = 0; k < 9; ++k) { R1[9 * p + k] = 0; R2[9 * p + k] = 0; }
        for (int k = 0; k < 3; ++k) T[3 * p + k] = 0;
        return;
    }
    for (int k = 0; k < 9; ++k) e[k] = E[9 * p + k];
    decompose_essential(e, r1, r2, t);
    for (int k = 0; k < 9; ++k) { R1[9 * p + k] = r1[k]; R2[9 * p + k] = r2[k]; }
    for (int k = 0; k < 3; ++k) T[3 * p + k] = t[k];
}

__global__ void __launch_bounds__(256)
normalize_kernel(const double2* __restrict__ a, const double2* __restrict__ b, int64_t total,
                 double invf, double nx, double ny, double* __restrict__ pd) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const double2 pa = a[i], pb = b[i];
    double* o = pd + 4 * i;
    o[0] = dadd(dmul(pa.x, invf), nx);
    o[1] = dadd(dmul(pa.y, invf), ny);
    o[2] = dadd(dmul(pb.x, invf), nx);
    o[3] = dadd(dmul(pb.y, invf), ny);
}

// The points recoverPose triangulates are those of the input mask (the RANSAC inliers, typically a quarter of the
// correspondences).  pose_compact_kernel lists them per pair (ordered), so that the triangulation kernel below runs on
// full blocks of masked points instead of blocks in which three warps of four have nothing to do.
__global__ void __launch_bounds__(256)
pose_compact_kernel(const int64_t* __restrict__ off, const uint8_t* __restrict__ ok, const uint8_t* __restrict__ maskIn,
                    int* __restrict__ list /*[total]*/, int* __restrict__ counts /*[nPairs]*/, uint8_t* __restrict__ cand) {
    const int p = blockIdx.x;
    const int64_t base = off[p];
    const int n = (int)(off[p + 1] - base);
    if (ok && !ok[p]) { if (threadIdx.x == 0) counts[p] = 0; return; }
    __shared__ int sWarp[8];
    __shared__ int sBase;
    if (threadIdx.x == 0) sBase = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int c0 = 0; c0 < n; c0 += 256) {
        const int i = c0 + threadIdx.x;
        const bool want = i < n && (!maskIn || maskIn[base + i]);
        if (i < n) cand[base + i] = 0;
        const unsigned bal = __ballot_sync(0xFFFFFFFFu, want);
        if (lane == 0) sWarp[warp] = __popc(bal);
        __syncthreads();
        int before = 0, total = 0;
        for (int w = 0; w < 8; ++w) { if (w < warp) before += sWarp[w]; total += sWarp[w]; }
        if (want) list[base + sBase + before + __popc(bal & ((1u << lane) - 1u))] = i;
        __syncthreads();
        if (threadIdx.x == 0) sBase += total;
        __syncthreads();
    }
    if (threadIdx.x == 0) counts[p] = sBase;
}

// one masked correspondence per thread, four pose candidates each
__global__ void __launch_bounds__(128)
cheirality_kernel(const double* __restrict__ pd, const int64_t* __restrict__ off,
                  const uint8_t* __restrict__ ok, const double* __restrict__ R1,
                  const double* __restrict__ R2, const double* __restrict__ T,
                  const int* __restrict__ list, const int* __restrict__ counts, uint8_t* __restrict__ cand /*[total] bit k*/,
                  int* __restrict__ good4) {
    const int p = blockIdx.x;            // pairs on grid.x (no 65535 limit), tiles of 128 masked points on grid.y
    if (ok && !ok[p]) return;
    const int cnt = min(128, counts[p] - (int)blockIdx.y * 128);
    if (cnt <= 0) return;
    const int64_t base = off[p];
    __shared__ double sR[18], sT[3];
    if (threadIdx.x < 9) { sR[threadIdx.x] = R1[9 * p + threadIdx.x]; sR[9 + threadIdx.x] = R2[9 * p + threadIdx.x]; }
    if (threadIdx.x < 3) sT[threadIdx.x] = T[3 * p + threadIdx.x];
    __syncthreads();
    // scratch of the 4 x 4 Jacobi SVDs in thread-interleaved shared memory (as local arrays the resident threads'
    // frames exceed L2 and are written back to DRAM)
    __shared__ double sScr[32 * 128];
    double* scr = sScr + threadIdx.x;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned bits = 0;
    if ((int)threadIdx.x < cnt) {
        const int j = list[base + blockIdx.y * 128 + threadIdx.x];
        const double* q = pd + 4 * (base + j);
        const double tp[3] = {sT[0], sT[1], sT[2]}, tn[3] = {-sT[0], -sT[1], -sT[2]};
        if (cheirality_ok_at<128>(sR, tp, q[0], q[1], q[2], q[3], 50.0, scr)) bits |= 1;
        if (cheirality_ok_at<128>(sR + 9, tp, q[0], q[1], q[2], q[3], 50.0, scr)) bits |= 2;
        if (cheirality_ok_at<128>(sR, tn, q[0], q[1], q[2], q[3], 50.0, scr)) bits |= 4;
        if (cheirality_ok_at<128>(sR + 9, tn, q[0], q[1], q[2], q[3], 50.0, scr)) bits |= 8;
        cand[base + j] = (uint8_t)bits;
    }
    if (warp * 32 < cnt) {                                   // warp-uniform: this warp holds compacted points
        for (int k = 0; k < 4; ++k) {
            const unsigned b = __ballot_sync(0xFFFFFFFFu, (bits >> k) & 1);
            if (lane == 0 && b) atomicAdd(good4 + 4 * p + k, __popc(b));
        }
    }
}

__global__ void __launch_bounds__(64)
pose_select_kernel(int nPairs, const uint8_t* __restrict__ ok, const double* __restrict__ R1,
                   const double* __restrict__ R2, const double* __restrict__ T,
                   const int* __restrict__ good4, double* __restrict__ R, double* __restrict__ Tsel,
                   int* __restrict__ good, int* __restrict__ choice) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= nPairs) return;
    if (ok && !ok[p]) { if (good) good[p] = 0; choice[p] = -1; return; }
    const int g1 = good4[4 * p], g2 = good4[4 * p + 1], g3 = good4[4 * p + 2], g4 = good4[4 * p + 3];
    int k;
    if (g1 >= g2 && g1 >= g3 && g1 >= g4) k = 0;
    else if (g2 >= g1 && g2 >= g3 && g2 >= g4) k = 1;
    else if (g3 >= g1 && g3 >= g2 && g3 >= g4) k = 2;
    else k = 3;
    const double* Rs = (k & 1) ? R2 + 9 * p : R1 + 9 * p;
    const double sg = (k >= 2) ? -1.0 : 1.0;
    if (R) for (int j = 0; j < 9; ++j) R[9 * p + j] = Rs[j];
    if (Tsel) for (int j = 0; j < 3; ++j) Tsel[3 * p + j] = sg * T[3 * p + j];
    if (good) good[p] = good4[4 * p + k];
    choice[p] = k;
}

__global__ void __launch_bounds__(256)
pose_mask_kernel(const int64_t* __restrict__ off, const int* __restrict__ choice,
                 const uint8_t* __restrict__ cand, uint8_t* __restrict__ maskOut) {
    const int p = blockIdx.x;
    const int64_t base = off[p];
    const int n = (int)(off[p + 1] - base);
    const int i = blockIdx.y * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int k = choice[p];
    maskOut[base + i] = (k >= 0 && ((cand[base + i] >> k) & 1)) ? 1 : 0;
}

}  // namespace

int recover_pose_batch_device(Context* ctx, cudaStream_t st, const PoseIO& io) {
    const int nPairs = io.nPairs;
    if (nPairs <= 0) return 0;
    const int64_t total = io.hOff[nPairs];
    const size_t n1 = (size_t)nPairs + 1;
    if (ctx->pin[2].reserve(n1 * 8) || ctx->ws[22].reserve(n1 * 8 + (size_t)nPairs * (9 + 9 + 3) * 8 + (size_t)nPairs * 24 + 256))
        return -1;
    MCV_CUDA_TRY(cudaEventSynchronize(ctx->ev[6]));
    memcpy(ctx->pin[2].p, io.hOff, n1 * 8);
    int64_t* dOff = ctx->ws[22].as<int64_t>();
    MCV_CUDA_TRY(cudaMemcpyAsync(dOff, ctx->pin[2].p, n1 * 8, cudaMemcpyHostToDevice, st));
    MCV_CUDA_TRY(cudaEventRecord(ctx->ev[6], st));
    double* wR1 = reinterpret_cast<double*>(dOff + n1);
    double* wR2 = wR1 + 9 * (size_t)nPairs;
    double* wT = wR2 + 9 * (size_t)nPairs;
    int* wGood4 = reinterpret_cast<int*>(wT + 3 * (size_t)nPairs);
    int* wChoice = wGood4 + 4 * (size_t)nPairs;
    double* R1 = io.dR1 ? io.dR1 : wR1;
    double* R2 = io.dR2 ? io.dR2 : wR2;
    double* T = io.dT ? io.dT : wT;
    kt_begin(ctx, st, KT_POSE);
    decompose_kernel<<<(nPairs + 63) / 64, 64, 0, st>>>(io.dE, io.dOk, nPairs, R1, R2, T);
    MCV_LAUNCH_CHECK();
    const bool wantPose = io.dR || io.dTsel || io.dGood || io.dGood4 || io.dMaskOut;
    if (!wantPose) { kt_end(ctx, st); return 0; }
    if (ctx->ws[21].reserve((size_t)std::max<int64_t>(total, 1) * 32) ||
        ctx->ws[23].reserve((size_t)std::max<int64_t>(total, 1)) ||
        ctx->ws[47].reserve((size_t)std::max<int64_t>(total, 1) * 4 + (size_t)nPairs * 4 + 64))
        return -1;
    double* pd = ctx->ws[21].as<double>();
    uint8_t* cand = ctx->ws[23].as<uint8_t>();
    int* dCounts = ctx->ws[47].as<int>();
    int* dList = dCounts + nPairs;
    int64_t maxN = 0;
    for (int p = 0; p < nPairs; ++p) maxN = std::max(maxN, io.hOff[p + 1] - io.hOff[p]);
    int* good4 = io.dGood4 ? io.dGood4 : wGood4;
    MCV_CUDA_TRY(cudaMemsetAsync(good4, 0, (size_t)nPairs * 16, st));
    if (total > 0) {
        const double invf = 1.0 / io.focal;
        normalize_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(
            reinterpret_cast<const double2*>(io.dA), reinterpret_cast<const double2*>(io.dB), total,
            invf, -io.ppx * invf, -io.ppy * invf, pd);
        MCV_LAUNCH_CHECK();
        pose_compact_kernel<<<nPairs, 256, 0, st>>>(dOff, io.dOk, io.dMaskIn, dList, dCounts, cand);
        MCV_LAUNCH_CHECK();
        if ((maxN + 127) / 128 > 65535) { set_error("recoverPose: more than 8.3 M correspondences in one pair"); return -1; }
        dim3 g(nPairs, (unsigned)((maxN + 127) / 128));
        cheirality_kernel<<<g, 128, 0, st>>>(pd, dOff, io.dOk, R1, R2, T, dList, dCounts, cand, good4);
        MCV_LAUNCH_CHECK();
    }
    pose_select_kernel<<<(nPairs + 63) / 64, 64, 0, st>>>(nPairs, io.dOk, R1, R2, T, good4, io.dR,
                                                          io.dTsel, io.dGood, wChoice);
    MCV_LAUNCH_CHECK();
    if (io.dMaskOut && total > 0) {
        dim3 g(nPairs, (unsigned)((maxN + 255) / 256));
        pose_mask_kernel<<<g, 256, 0, st>>>(dOff, wChoice, cand, io.dMaskOut);
        MCV_LAUNCH_CHECK();
    }
    kt_end(ctx, st);
    return 0;
}

// five_point_device: ransac.cu (the 5-point solver runs as the three stage kernels there)

}  // namespace mcv
