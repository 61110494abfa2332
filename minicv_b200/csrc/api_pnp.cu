// api_pnp.cu -- C-ABI entry points of the PnP-RANSAC row:
//   cvSolvePnPRansac (reference signature, src/MiniCVNative/MiniCVNative.cpp:93-139), native for the AP3P
//   solver kind; solveAp3p (reference signature, src/MiniCVNative/ap3p.cpp:282-317); and the batched /
//   injected-hypotheses forms.  Thin host layer over pnp.cu: stages buffers, never computes on the CPU.
#include "ransac.cuh"
#include "staging.cuh"

#include <algorithm>
#include <string.h>

namespace mcv {
bool forward_solve_pnp_ransac(const McvPoint2d* img, const McvPoint3d* world, const int N, const McvMatx33d K,
                              const double* dist, const int solverKind, const int iterationsCount,
                              const float reprojectionError, const double confidence, McvVec3d* tVec,
                              McvVec3d* rVec, int* inlierCount, int* outInliers);
}

using namespace mcv;

namespace {

// carve one workspace into aligned pieces
struct Carver {
    uint8_t* base = nullptr;
    size_t used = 0;
    template <class T> T* take(size_t count) {
        T* p = reinterpret_cast<T*>(base + used);
        used += align_up(std::max<size_t>(count, 1) * sizeof(T), 256);
        return p;
    }
};

PnpParams to_params(const McvPnpParams* p) {
    PnpParams r;
    r.sampleMode = p->sampleMode; r.stopMode = p->stopMode; r.maxIters = p->maxIters; r.refine = p->refine;
    r.threshold = p->threshold; r.confidence = p->confidence; r.seed = p->seed;
    r.problemIndexBase = p->problemIndexBase;
    return r;
}

int check_args(const void* obj, const void* img, const int64_t* off, int nProblems, const void* cams) {
    if (nProblems < 0 || !off) { set_error("bad batch arguments"); return -1; }
    if (off[0] != 0) { set_error("offsets must start at 0"); return -1; }
    for (int p = 0; p < nProblems; ++p)
        if (off[p + 1] < off[p] || off[p + 1] - off[p] > 0x3FFFFFFF) { set_error("bad offsets at problem %d", p); return -1; }
    if (nProblems > 0 && off[nProblems] > 0 && (!obj || !img)) { set_error("null point arrays"); return -1; }
    if (nProblems > 0 && !cams) { set_error("null camera array"); return -1; }
    return 0;
}

}  // namespace

extern "C" {

MCV_API int cvSolvePnPRansacBatchDevice(const McvPnpParams* prm, const double* dObj, const double* dImg,
                                        const int64_t* offsets, int nProblems, const double* dCams,
                                        double* dRvecs, double* dTvecs, uint8_t* dMasks, int32_t* dInlierCounts,
                                        uint8_t* dOk, int device, void* stream) {
    MCV_API_ENTER();
    if (!prm || check_args(dObj, dImg, offsets, nProblems, dCams)) return -1;
    if (nProblems == 0) return 0;
    Context* ctx = get_context_async(device, static_cast<cudaStream_t>(stream));
    if (!ctx) return -1;
    PnpIO io;
    io.dObj = dObj; io.dImg = dImg; io.hOff = offsets; io.nProblems = nProblems; io.dCam = dCams;
    io.dRvec = dRvecs; io.dTvec = dTvecs; io.dMasks = dMasks; io.dInlierCounts = dInlierCounts; io.dOk = dOk;
    return pnp_ransac_batch_device(ctx, static_cast<cudaStream_t>(stream), to_params(prm), io);
}

MCV_API int cvSolvePnPRansacBatch(const McvPnpParams* prm, const McvPoint3d* obj, const McvPoint2d* img,
                                  const int64_t* offsets, int nProblems, const double* cams,
                                  const int32_t* injSamples, const double* injModels, McvVec3d* rvecs,
                                  McvVec3d* tvecs, uint8_t* masks, int32_t* inlierCounts, uint8_t* ok,
                                  int32_t* itersRun, int32_t* bestIter, double* minimal, int32_t* dbgCounts,
                                  double* dbgModels, int32_t* dbgSamples, int device) {
    MCV_API_ENTER();
    if (!prm || check_args(obj, img, offsets, nProblems, cams)) return -1;
    if (nProblems == 0) return 0;
    if (!rvecs || !tvecs || !ok) { set_error("cvSolvePnPRansacBatch: null outputs"); return -1; }
    Context* ctx = get_context(device);
    if (!ctx) return -1;
    cudaStream_t st = ctx->stream;
    const PnpParams pp = to_params(prm);
    const int maxIters = std::max(pp.maxIters, 1);
    const int64_t total = offsets[nProblems];
    const size_t tb = (size_t)std::max<int64_t>(total, 1), np = (size_t)nProblems, nh = np * maxIters;

    size_t bytes = 4096 + 256 * 24;
    bytes += tb * (24 + 16 + 1 + 4) + np * (64 + 24 + 24 + 4 + 1 + 4 + 4 + 48);
    if (injSamples) bytes += nh * 16;
    if (injModels) bytes += nh * 48;
    if (dbgCounts) bytes += nh * 4;
    if (dbgModels) bytes += nh * 48;
    if (dbgSamples) bytes += nh * 16;
    if (ctx->ws[61].reserve(bytes)) return -1;
    Carver cv; cv.base = ctx->ws[61].as<uint8_t>();
    double* dObj = cv.take<double>(tb * 3);
    double* dImg = cv.take<double>(tb * 2);
    double* dCam = cv.take<double>(np * 8);
    double* dR = cv.take<double>(np * 3);
    double* dT = cv.take<double>(np * 3);
    uint8_t* dMasks = cv.take<uint8_t>(tb);
    int* dCnt = cv.take<int>(np);
    uint8_t* dOk = cv.take<uint8_t>(np);
    int* dIters = cv.take<int>(np);
    int* dBestIt = cv.take<int>(np);
    double* dMin = cv.take<double>(np * 6);
    if (total > 0) {
        if (h2d_copy(ctx, dObj, obj, (size_t)total * 24, st, host_is_pageable(obj)) ||
            h2d_copy(ctx, dImg, img, (size_t)total * 16, st, host_is_pageable(img)))
            return -1;
    }
    MCV_CUDA_TRY(cudaMemcpyAsync(dCam, cams, np * 64, cudaMemcpyHostToDevice, st));
    PnpIO io;
    io.dObj = dObj; io.dImg = dImg; io.hOff = offsets; io.nProblems = nProblems; io.dCam = dCam;
    io.dRvec = dR; io.dTvec = dT; io.dMasks = dMasks; io.dInlierCounts = dCnt; io.dOk = dOk;
    io.dItersRun = dIters; io.dBestIter = dBestIt; io.dMinimal = dMin;
    if (injSamples) {
        int* d = cv.take<int>(nh * 4);
        MCV_CUDA_TRY(cudaMemcpyAsync(d, injSamples, nh * 16, cudaMemcpyHostToDevice, st));
        io.dInjSamples = d;
    }
    if (injModels) {
        double* d = cv.take<double>(nh * 6);
        MCV_CUDA_TRY(cudaMemcpyAsync(d, injModels, nh * 48, cudaMemcpyHostToDevice, st));
        io.dInjModels = d;
    }
    if (dbgCounts) io.dDbgCounts = cv.take<int>(nh);
    if (dbgModels) io.dDbgModels = cv.take<double>(nh * 6);
    if (dbgSamples) io.dDbgSamples = cv.take<int>(nh * 4);
    if (cv.used > bytes) { set_error("internal: staging carve overflow"); return -1; }
    if (pnp_ransac_batch_device(ctx, st, pp, io)) return -1;
    MCV_CUDA_TRY(cudaMemcpyAsync(rvecs, dR, np * 24, cudaMemcpyDeviceToHost, st));
    MCV_CUDA_TRY(cudaMemcpyAsync(tvecs, dT, np * 24, cudaMemcpyDeviceToHost, st));
    MCV_CUDA_TRY(cudaMemcpyAsync(ok, dOk, np, cudaMemcpyDeviceToHost, st));
    if (masks && total > 0) MCV_CUDA_TRY(cudaMemcpyAsync(masks, dMasks, (size_t)total, cudaMemcpyDeviceToHost, st));
    if (inlierCounts) MCV_CUDA_TRY(cudaMemcpyAsync(inlierCounts, dCnt, np * 4, cudaMemcpyDeviceToHost, st));
    if (itersRun) MCV_CUDA_TRY(cudaMemcpyAsync(itersRun, dIters, np * 4, cudaMemcpyDeviceToHost, st));
    if (bestIter) MCV_CUDA_TRY(cudaMemcpyAsync(bestIter, dBestIt, np * 4, cudaMemcpyDeviceToHost, st));
    if (minimal) MCV_CUDA_TRY(cudaMemcpyAsync(minimal, dMin, np * 48, cudaMemcpyDeviceToHost, st));
    if (dbgCounts) MCV_CUDA_TRY(cudaMemcpyAsync(dbgCounts, io.dDbgCounts, nh * 4, cudaMemcpyDeviceToHost, st));
    if (dbgModels) MCV_CUDA_TRY(cudaMemcpyAsync(dbgModels, io.dDbgModels, nh * 48, cudaMemcpyDeviceToHost, st));
    if (dbgSamples) MCV_CUDA_TRY(cudaMemcpyAsync(dbgSamples, io.dDbgSamples, nh * 16, cudaMemcpyDeviceToHost, st));
    MCV_CUDA_TRY(cudaStreamSynchronize(st));
    int n = 0;
    for (int i = 0; i < nProblems; ++i) n += ok[i] ? 1 : 0;
    return n;
}

MCV_API bool cvSolvePnPRansac(const McvPoint2d* imgPoints, const McvPoint3d* worldPoints, const int N,
                              const McvMatx33d K, const double* distortionCoeffs, const int solverKind,
                              const int iterationsCount, const float reprojectionError, const double confidence,
                              McvVec3d* tVec, McvVec3d* rVec, int* inlierCount, int* outInliers) {
    if (solverKind != 5)
        return forward_solve_pnp_ransac(imgPoints, worldPoints, N, K, distortionCoeffs, solverKind, iterationsCount,
                                        reprojectionError, confidence, tVec, rVec, inlierCount, outInliers);
    MCV_API_ENTER();
    if (inlierCount) *inlierCount = 0;
    if (N < 4 || !imgPoints || !worldPoints || !tVec || !rVec) {
        if (N >= 0 && N < 4) return false;      // cv::solvePnPRansac asserts npoints >= 4; the shim would throw
        set_error("cvSolvePnPRansac: bad arguments");
        return false;
    }
    if (!(confidence > 0 && confidence < 1)) { set_error("cvSolvePnPRansac: confidence must be in (0,1)"); return false; }
    Context* ctx = get_context(-1);
    if (!ctx) return false;
    cudaStream_t st = ctx->stream;
    const size_t tb = (size_t)N;
    const size_t bytes = 8192 + tb * (24 + 16 + 1 + 4);
    if (ctx->ws[61].reserve(bytes)) return false;
    Carver cv; cv.base = ctx->ws[61].as<uint8_t>();
    double* dObj = cv.take<double>(tb * 3);
    double* dImg = cv.take<double>(tb * 2);
    double* dCam = cv.take<double>(8);
    double* dR = cv.take<double>(3);
    double* dT = cv.take<double>(3);
    uint8_t* dMasks = cv.take<uint8_t>(tb);
    int* dIdx = cv.take<int>(tb);
    int* dCnt = cv.take<int>(1);
    uint8_t* dOk = cv.take<uint8_t>(1);
    const double cam[8] = {K.M[0], K.M[4], K.M[2], K.M[5],
                           distortionCoeffs ? distortionCoeffs[0] : 0.0, distortionCoeffs ? distortionCoeffs[1] : 0.0,
                           distortionCoeffs ? distortionCoeffs[2] : 0.0, distortionCoeffs ? distortionCoeffs[3] : 0.0};
    auto fail = [&]() { return false; };
    if (cudaMemcpyAsync(dObj, worldPoints, tb * 24, cudaMemcpyHostToDevice, st) != cudaSuccess ||
        cudaMemcpyAsync(dImg, imgPoints, tb * 16, cudaMemcpyHostToDevice, st) != cudaSuccess ||
        cudaMemcpyAsync(dCam, cam, 64, cudaMemcpyHostToDevice, st) != cudaSuccess) {
        set_error("cvSolvePnPRansac: host->device copy failed");
        return fail();
    }
    const int64_t off[2] = {0, N};
    PnpParams pp;
    pp.maxIters = iterationsCount; pp.threshold = (double)reprojectionError; pp.confidence = confidence;
    PnpIO io;
    io.dObj = dObj; io.dImg = dImg; io.hOff = off; io.nProblems = 1; io.dCam = dCam;
    io.dRvec = dR; io.dTvec = dT; io.dMasks = dMasks; io.dInlierIdx = dIdx; io.dInlierCounts = dCnt; io.dOk = dOk;
    if (pnp_ransac_batch_device(ctx, st, pp, io)) return fail();
    uint8_t ok = 0;
    int cnt = 0;
    double r[3], t[3];
    if (cudaMemcpyAsync(&ok, dOk, 1, cudaMemcpyDeviceToHost, st) != cudaSuccess ||
        cudaMemcpyAsync(&cnt, dCnt, 4, cudaMemcpyDeviceToHost, st) != cudaSuccess ||
        cudaMemcpyAsync(r, dR, 24, cudaMemcpyDeviceToHost, st) != cudaSuccess ||
        cudaMemcpyAsync(t, dT, 24, cudaMemcpyDeviceToHost, st) != cudaSuccess ||
        cudaStreamSynchronize(st) != cudaSuccess) {
        set_error("cvSolvePnPRansac: device->host copy failed");
        return fail();
    }
    if (!ok) return false;
    if (outInliers && cnt > 0) {
        if (cudaMemcpyAsync(outInliers, dIdx, (size_t)cnt * 4, cudaMemcpyDeviceToHost, st) != cudaSuccess ||
            cudaStreamSynchronize(st) != cudaSuccess) {
            set_error("cvSolvePnPRansac: device->host copy failed");
            return fail();
        }
    }
    memcpy(rVec->V, r, 24);
    memcpy(tVec->V, t, 24);
    if (inlierCount) *inlierCount = cnt;
    return true;
}

MCV_API int cvSolveAp3pBatch(const McvPoint2d* img, const McvPoint3d* obj, const double* cam4, int n,
                             McvMatx33d* Rs, McvVec3d* ts, int32_t* counts) {
    MCV_API_ENTER();
    if (n < 0 || (n > 0 && (!img || !obj || !cam4 || !Rs || !ts || !counts))) { set_error("cvSolveAp3pBatch: bad arguments"); return -1; }
    if (n == 0) return 0;
    Context* ctx = get_context(-1);
    if (!ctx) return -1;
    cudaStream_t st = ctx->stream;
    const size_t m = (size_t)n;
    if (ctx->ws[62].reserve(4096 + m * (48 + 72 + 32 + 288 + 96 + 4))) return -1;
    Carver cv; cv.base = ctx->ws[62].as<uint8_t>();
    double* dImg = cv.take<double>(m * 6);
    double* dObj = cv.take<double>(m * 9);
    double* dCam = cv.take<double>(m * 4);
    double* dRs = cv.take<double>(m * 36);
    double* dTs = cv.take<double>(m * 12);
    int* dCnt = cv.take<int>(m);
    MCV_CUDA_TRY(cudaMemcpyAsync(dImg, img, m * 48, cudaMemcpyHostToDevice, st));
    MCV_CUDA_TRY(cudaMemcpyAsync(dObj, obj, m * 72, cudaMemcpyHostToDevice, st));
    MCV_CUDA_TRY(cudaMemcpyAsync(dCam, cam4, m * 32, cudaMemcpyHostToDevice, st));
    if (ap3p_batch_device(ctx, st, dImg, dObj, dCam, n, dRs, dTs, dCnt)) return -1;
    MCV_CUDA_TRY(cudaMemcpyAsync(counts, dCnt, m * 4, cudaMemcpyDeviceToHost, st));
    MCV_CUDA_TRY(cudaStreamSynchronize(st));
    // only the poses that exist are written, like the reference (the rest of the caller's buffers is untouched)
    for (int i = 0; i < n; ++i) {
        const int c = counts[i];
        if (c <= 0) continue;
        MCV_CUDA_TRY(cudaMemcpyAsync(Rs + 4 * (size_t)i, dRs + 36 * (size_t)i, (size_t)c * 72, cudaMemcpyDeviceToHost, st));
        MCV_CUDA_TRY(cudaMemcpyAsync(ts + 4 * (size_t)i, dTs + 12 * (size_t)i, (size_t)c * 24, cudaMemcpyDeviceToHost, st));
    }
    MCV_CUDA_TRY(cudaStreamSynchronize(st));
    return 0;
}

MCV_API int solveAp3p(McvMatx33d* R, McvVec3d* t, double mu0, double mv0, double X0, double Y0, double Z0,
                      double mu1, double mv1, double X1, double Y1, double Z1, double mu2, double mv2,
                      double X2, double Y2, double Z2, double inv_fx, double inv_fy, double cx_fx, double cy_fy) {
    if (!R || !t) { set_error("solveAp3p: null outputs"); return 0; }
    const McvPoint2d img[3] = {{mu0, mv0}, {mu1, mv1}, {mu2, mv2}};
    const McvPoint3d obj[3] = {{X0, Y0, Z0}, {X1, Y1, Z1}, {X2, Y2, Z2}};
    const double cam4[4] = {inv_fx, inv_fy, cx_fx, cy_fy};
    int32_t n = 0;
    if (cvSolveAp3pBatch(img, obj, cam4, 1, R, t, &n)) return 0;
    return n;
}

}  // extern "C"
