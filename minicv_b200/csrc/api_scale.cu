// api_scale.cu -- C-ABI entry points of the scale-search row:
//   cvFindScaled / cvFindScaledBatch / cvFindScaledBatchDevice == CameraPose.findScaled
//   (src/MiniCV/CameraPose.fs:39-134) and cvScaleCosts == its inner avgReprojectionError (:75-90).
// Thin host layer over scale_search.cu: packs the blittable structs, stages buffers, never computes on the CPU.
#include "ransac.cuh"

#include <algorithm>
#include <math.h>
#include <string.h>

using namespace mcv;

namespace {

struct Carver {
    uint8_t* base = nullptr;
    size_t used = 0;
    template <class T> T* take(size_t count) {
        T* p = reinterpret_cast<T*>(base + used);
        used += align_up(std::max<size_t>(count, 1) * sizeof(T), 256);
        return p;
    }
};

static_assert(sizeof(McvCamera) == 112, "McvCamera must be 14 packed doubles");
static_assert(sizeof(McvCameraPose) == 112, "McvCameraPose layout");

int check_offsets(const int64_t* off, int nProblems) {
    if (nProblems < 0 || !off) { set_error("findScaled: bad batch arguments"); return -1; }
    if (off[0] != 0) { set_error("findScaled: offsets must start at 0"); return -1; }
    for (int p = 0; p < nProblems; ++p)
        if (off[p + 1] < off[p] || off[p + 1] - off[p] > 0x3FFFFFFF) { set_error("findScaled: bad offsets at problem %d", p); return -1; }
    return 0;
}

}  // namespace

extern "C" {

MCV_API int cvFindScaledBatchDevice(const double* dCams, const double* dRot, const double* dTrans,
                                    const double* dWorld, const double* dObs, const int64_t* offsets,
                                    int nProblems, double* dCosts, double* dBestScales, double* dScaledT,
                                    int32_t* dSigns, int device, void* stream) {
    MCV_API_ENTER();
    if (check_offsets(offsets, nProblems)) return -1;
    if (nProblems == 0) return 0;
    if (offsets[nProblems] > 0 && (!dWorld || !dObs)) { set_error("findScaled: null observation arrays"); return -1; }
    Context* ctx = get_context_async(device, static_cast<cudaStream_t>(stream));
    if (!ctx) return -1;
    ScaleIO io;
    io.dCams = dCams; io.dRot = dRot; io.dTrans = dTrans; io.dWorld = dWorld; io.dObs = dObs;
    io.hOff = offsets; io.nProblems = nProblems;
    io.dCost = dCosts; io.dBestScale = dBestScales; io.dScaledT = dScaledT; io.dSign = dSigns;
    return find_scaled_batch_device(ctx, static_cast<cudaStream_t>(stream), io);
}

MCV_API int cvFindScaledBatch(const McvCamera* srcCams, const McvCameraPose* poses, const McvPoint3d* world,
                              const McvPoint2d* obs, const int64_t* offsets, int nProblems, double* costs,
                              McvCameraPose* scaled, double* bestScales, double* candidates, uint8_t* candOk,
                              double* candCosts, int device) {
    MCV_API_ENTER();
    if (check_offsets(offsets, nProblems)) return -1;
    if (nProblems == 0) return 0;
    if (!srcCams || !poses || !costs || !scaled) { set_error("cvFindScaledBatch: null cameras / poses / outputs"); return -1; }
    const int64_t total = offsets[nProblems];
    if (total > 0 && (!world || !obs)) { set_error("cvFindScaledBatch: null observation arrays"); return -1; }
    Context* ctx = get_context(device);
    if (!ctx) return -1;
    cudaStream_t st = ctx->stream;
    const size_t np = (size_t)nProblems, tb = (size_t)std::max<int64_t>(total, 1);

    // pose rotations / translations packed next to the results in one pinned staging block
    const size_t hostBytes = np * (9 + 3 + 1 + 1 + 3) * 8 + np * 4 + 1024;
    if (ctx->pin[5].reserve(hostBytes)) return -1;
    MCV_CUDA_TRY(cudaStreamSynchronize(st));                       // the staging block is reused call to call
    double* hRot = ctx->pin[5].as<double>();
    double* hTrans = hRot + 9 * np;
    double* hCost = hTrans + 3 * np;
    double* hScale = hCost + np;
    double* hScaledT = hScale + np;
    int32_t* hSign = reinterpret_cast<int32_t*>(hScaledT + 3 * np);
    for (size_t p = 0; p < np; ++p) {
        memcpy(hRot + 9 * p, poses[p].Rotation, 72);
        memcpy(hTrans + 3 * p, poses[p].Translation, 24);
    }

    size_t bytes = 8192 + tb * (24 + 16) + np * (112 + 72 + 24 + 8 + 8 + 24 + 4);
    if (candidates) bytes += tb * 16 + 256;
    if (candCosts) bytes += tb * 16 + 256;
    if (candOk) bytes += tb * 2 + 256;
    if (ctx->ws[66].reserve(bytes)) return -1;
    Carver cv; cv.base = ctx->ws[66].as<uint8_t>();
    double* dWorld = cv.take<double>(tb * 3);
    double* dObs = cv.take<double>(tb * 2);
    double* dCams = cv.take<double>(np * 14);
    double* dRot = cv.take<double>(np * 9);
    double* dTrans = cv.take<double>(np * 3);
    double* dCost = cv.take<double>(np);
    double* dScale = cv.take<double>(np);
    double* dScaledT = cv.take<double>(np * 3);
    int* dSign = cv.take<int>(np);
    ScaleIO io;
    if (candidates) io.dCand = cv.take<double>(tb * 2);
    if (candCosts) io.dCandCost = cv.take<double>(tb * 2);
    if (candOk) io.dCandOk = cv.take<uint8_t>(tb * 2);
    if (cv.used > bytes) { set_error("internal: staging carve overflow"); return -1; }
    if (total > 0) {
        MCV_CUDA_TRY(cudaMemcpyAsync(dWorld, world, (size_t)total * 24, cudaMemcpyHostToDevice, st));
        MCV_CUDA_TRY(cudaMemcpyAsync(dObs, obs, (size_t)total * 16, cudaMemcpyHostToDevice, st));
    }
    MCV_CUDA_TRY(cudaMemcpyAsync(dCams, srcCams, np * 112, cudaMemcpyHostToDevice, st));
    MCV_CUDA_TRY(cudaMemcpyAsync(dRot, hRot, np * 72, cudaMemcpyHostToDevice, st));
    MCV_CUDA_TRY(cudaMemcpyAsync(dTrans, hTrans, np * 24, cudaMemcpyHostToDevice, st));
    io.dCams = dCams; io.dRot = dRot; io.dTrans = dTrans; io.dWorld = dWorld; io.dObs = dObs;
    io.hOff = offsets; io.nProblems = nProblems;
    io.dCost = dCost; io.dBestScale = dScale; io.dScaledT = dScaledT; io.dSign = dSign;
    if (find_scaled_batch_device(ctx, st, io)) return -1;
    MCV_CUDA_TRY(cudaMemcpyAsync(hCost, dCost, np * 8, cudaMemcpyDeviceToHost, st));
    MCV_CUDA_TRY(cudaMemcpyAsync(hScale, dScale, np * 8, cudaMemcpyDeviceToHost, st));
    MCV_CUDA_TRY(cudaMemcpyAsync(hScaledT, dScaledT, np * 24, cudaMemcpyDeviceToHost, st));
    MCV_CUDA_TRY(cudaMemcpyAsync(hSign, dSign, np * 4, cudaMemcpyDeviceToHost, st));
    if (total > 0) {
        if (candidates) MCV_CUDA_TRY(cudaMemcpyAsync(candidates, io.dCand, (size_t)total * 16, cudaMemcpyDeviceToHost, st));
        if (candCosts) MCV_CUDA_TRY(cudaMemcpyAsync(candCosts, io.dCandCost, (size_t)total * 16, cudaMemcpyDeviceToHost, st));
        if (candOk) MCV_CUDA_TRY(cudaMemcpyAsync(candOk, io.dCandOk, (size_t)total * 2, cudaMemcpyDeviceToHost, st));
    }
    MCV_CUDA_TRY(cudaStreamSynchronize(st));
    for (size_t p = 0; p < np; ++p) {
        costs[p] = hCost[p];
        if (bestScales) bestScales[p] = hScale[p];
        McvCameraPose out;
        memset(&out, 0, sizeof(out));
        if (offsets[p + 1] > offsets[p]) {                        // CameraPose.scale bestScale pose
            out.RotationIndex = poses[p].RotationIndex;
            out.ScaleSign = hSign[p] * poses[p].ScaleSign;
            memcpy(out.Rotation, poses[p].Rotation, 72);
            memcpy(out.Translation, hScaledT + 3 * p, 24);
            out.IsInverse = poses[p].IsInverse;
        }                                                         // else CameraPose(): all zero (CameraPose.fs:42)
        scaled[p] = out;
    }
    return 0;
}

MCV_API double cvFindScaled(double inlierThreshold, const McvCamera* srcCam, const McvPoint3d* world,
                            const McvPoint2d* obs, int n, const McvCameraPose* pose, McvCameraPose* scaled) {
    (void)inlierThreshold;                                        // unused by the reference as well (countInliers is dead)
    MCV_API_ENTER();
    if (n < 0 || !srcCam || !pose || !scaled) { set_error("cvFindScaled: bad arguments"); return NAN; }
    const int64_t off[2] = {0, n};
    double cost = NAN;
    McvCameraPose out;
    if (cvFindScaledBatch(srcCam, pose, world, obs, off, 1, &cost, &out, nullptr, nullptr, nullptr, nullptr, -1)) return NAN;
    *scaled = out;
    return cost;
}

MCV_API int cvScaleCosts(const McvCamera* srcCam, const McvCameraPose* pose, const McvPoint3d* world,
                         const McvPoint2d* obs, int n, const double* scales, int nScales, double* costs) {
    MCV_API_ENTER();
    if (n < 0 || nScales < 0 || !srcCam || !pose || (n > 0 && (!world || !obs)) || (nScales > 0 && (!scales || !costs))) {
        set_error("cvScaleCosts: bad arguments");
        return -1;
    }
    if (nScales == 0) return 0;
    Context* ctx = get_context(-1);
    if (!ctx) return -1;
    cudaStream_t st = ctx->stream;
    const size_t tb = (size_t)std::max(n, 1), ns = (size_t)nScales;
    const size_t bytes = 8192 + tb * 40 + ns * 16 + 112 + 96 + 64;
    if (ctx->ws[67].reserve(bytes)) return -1;
    Carver cv; cv.base = ctx->ws[67].as<uint8_t>();
    double* dWorld = cv.take<double>(tb * 3);
    double* dObs = cv.take<double>(tb * 2);
    double* dCam = cv.take<double>(14);
    double* dRot = cv.take<double>(9);
    double* dTrans = cv.take<double>(3);
    double* dScales = cv.take<double>(ns);
    double* dCosts = cv.take<double>(ns);
    int64_t* dOffs = cv.take<int64_t>(4);
    if (cv.used > bytes) { set_error("internal: staging carve overflow"); return -1; }
    const int64_t offs[4] = {0, n, 0, nScales};
    if (n > 0) {
        MCV_CUDA_TRY(cudaMemcpyAsync(dWorld, world, (size_t)n * 24, cudaMemcpyHostToDevice, st));
        MCV_CUDA_TRY(cudaMemcpyAsync(dObs, obs, (size_t)n * 16, cudaMemcpyHostToDevice, st));
    }
    MCV_CUDA_TRY(cudaMemcpyAsync(dCam, srcCam, 112, cudaMemcpyHostToDevice, st));
    MCV_CUDA_TRY(cudaMemcpyAsync(dRot, pose->Rotation, 72, cudaMemcpyHostToDevice, st));
    MCV_CUDA_TRY(cudaMemcpyAsync(dTrans, pose->Translation, 24, cudaMemcpyHostToDevice, st));
    MCV_CUDA_TRY(cudaMemcpyAsync(dScales, scales, ns * 8, cudaMemcpyHostToDevice, st));
    MCV_CUDA_TRY(cudaMemcpyAsync(dOffs, offs, 32, cudaMemcpyHostToDevice, st));
    MCV_CUDA_TRY(cudaStreamSynchronize(st));                       // offs lives on this stack frame
    if (scale_costs_device(ctx, st, dCam, dRot, dTrans, dWorld, dObs, dOffs, 1, dScales, nullptr, dOffs + 2, nScales, dCosts))
        return -1;
    MCV_CUDA_TRY(cudaMemcpyAsync(costs, dCosts, ns * 8, cudaMemcpyDeviceToHost, st));
    MCV_CUDA_TRY(cudaStreamSynchronize(st));
    return 0;
}

}  // extern "C"
