// api_estimate.cu -- C-ABI entry points of the two-view estimators.
//   group (1): cvRecoverPose / cvRecoverPoses / cvFivePoint with the reference's signatures
//              (src/MiniCVNative/MiniCVNative.cpp:165-215, 368-382)
//   group (2): findHomography / findFundamentalMat and the batched + injected forms
// Thin host layer over ransac.cu / pose.cu: stages buffers, never computes on the CPU.
#include "ransac.cuh"
#include "staging.cuh"

#include <vector>
#include <string.h>
#include <float.h>

using namespace mcv;

namespace {

struct Staged {     // device staging for one host-pointer call
    Context* ctx = nullptr;
    cudaStream_t st = nullptr;
    double* dA = nullptr; double* dB = nullptr;
    double* dModels = nullptr; uint8_t* dMasks = nullptr; int* dCnt = nullptr; uint8_t* dOk = nullptr;
    int* dIters = nullptr; int* dBest = nullptr;
};

RansacParams to_params(const McvRansacParams* p) {
    RansacParams r;
    r.model = p->model; r.sampleMode = p->sampleMode; r.stopMode = p->stopMode;
    r.maxIters = p->maxIters; r.threshold = p->threshold; r.confidence = p->confidence;
    r.seed = p->seed; r.focal = p->focal; r.ppx = p->ppx; r.ppy = p->ppy; r.refine = p->refine;
    r.pairIndexBase = p->pairIndexBase;
    return r;
}

int stage_alloc(Staged& s, int64_t total, int nPairs) {
    Context* c = s.ctx;
    const size_t tb = (size_t)std::max<int64_t>(total, 1);
    if (c->ws[30].reserve(tb * 16) || c->ws[31].reserve(tb * 16) || c->ws[32].reserve((size_t)nPairs * 72 + 72) ||
        c->ws[33].reserve(tb) || c->ws[34].reserve((size_t)nPairs * 4 + 16) || c->ws[35].reserve((size_t)nPairs + 16) ||
        c->ws[36].reserve((size_t)nPairs * 4 + 16) || c->ws[37].reserve((size_t)nPairs * 8 + 16))
        return -1;
    s.dA = c->ws[30].as<double>(); s.dB = c->ws[31].as<double>();
    s.dModels = c->ws[32].as<double>(); s.dMasks = c->ws[33].as<uint8_t>();
    s.dCnt = c->ws[34].as<int>(); s.dOk = c->ws[35].as<uint8_t>();
    s.dIters = c->ws[36].as<int>(); s.dBest = c->ws[37].as<int>();
    MCV_CUDA_TRY(cudaMemsetAsync(s.dModels, 0, (size_t)nPairs * 72, s.st));     // pairs without a model report a zero matrix
    return 0;
}

int stage_points(Staged& s, const McvPoint2d* a, const McvPoint2d* b, int64_t total, int nPairs) {
    if (stage_alloc(s, total, nPairs)) return -1;
    if (total > 0) {
        // pageable caller arrays (P/Invoke) go through the pinned ring, pinned ones are copied directly
        if (h2d_copy(s.ctx, s.dA, a, (size_t)total * 16, s.st, host_is_pageable(a)) ||
            h2d_copy(s.ctx, s.dB, b, (size_t)total * 16, s.st, host_is_pageable(b)))
            return -1;
    }
    return 0;
}

// Host-buffer batches in PAGEABLE caller memory (what P/Invoke passes) are cut into up to four runs of whole pairs with
// about the same number of correspondences: the points of run k + 1 go through the pinned staging ring while run k is
// estimated (pairs are independent, so the result is the same).  Pinned arrays are copied in one piece: their copy is a
// tenth of the call and four smaller launches cost more than the overlap returns (C3-H 133 k -> 116 k pairs/s when tried).
// bounds: pair indices, front() = 0, back() = nPairs; one run when the batch is small.
std::vector<int> pair_runs(const int64_t* off, int nPairs, bool pageable) {
    const int64_t total = off[nPairs];
    int runs = 4;
    if (!pageable || nPairs < 8 || total * 32 < (8ll << 20)) runs = 1;
    std::vector<int> b{0};
    for (int r = 1; r < runs; ++r) {
        const int64_t want = total * r / runs;
        int p = b.back();
        while (p < nPairs && off[p] < want) ++p;
        if (p > b.back() && p < nPairs) b.push_back(p);
    }
    b.push_back(nPairs);
    return b;
}
// the points of pairs [p0, p1) to the device on the context's copy stream; `ev` marks their arrival
int stage_run(Staged& s, const McvPoint2d* a, const McvPoint2d* b, const int64_t* off, int p0, int p1, cudaEvent_t ev) {
    const int64_t o0 = off[p0], o1 = off[p1];
    cudaStream_t cs = s.ctx->stream2;
    if (o1 > o0) {
        if (h2d_copy(s.ctx, s.dA + 2 * o0, a + o0, (size_t)(o1 - o0) * 16, cs, host_is_pageable(a)) ||
            h2d_copy(s.ctx, s.dB + 2 * o0, b + o0, (size_t)(o1 - o0) * 16, cs, host_is_pageable(b)))
            return -1;
    }
    MCV_CUDA_TRY(cudaEventRecord(ev, cs));
    MCV_CUDA_TRY(cudaStreamWaitEvent(s.st, ev, 0));
    return 0;
}

int check_batch_args(const void* a, const void* b, const int64_t* off, int nPairs) {
    if (nPairs < 0 || !off) { set_error("bad batch arguments"); return -1; }
    if (off[0] != 0) { set_error("offsets must start at 0"); return -1; }
    for (int p = 0; p < nPairs; ++p)
        if (off[p + 1] < off[p] || off[p + 1] - off[p] > 0x3FFFFFFF) { set_error("bad offsets at pair %d", p); return -1; }
    if (nPairs > 0 && off[nPairs] > 0 && (!a || !b)) { set_error("null point arrays"); return -1; }
    return 0;
}

// single-pair RANSAC with OpenCV semantics; returns 1/0/-1
int single_ransac(const RansacParams& prm, const McvPoint2d* a, const McvPoint2d* b, int N,
                  McvMatx33d* model, uint8_t* mask, int* inliers) {
    if (N < 0 || !a || !b || !model) { set_error("bad arguments"); return -1; }
    if (N < model_points(prm.model)) return 0;
    Context* ctx = get_context(-1);
    if (!ctx) return -1;
    Staged s; s.ctx = ctx; s.st = ctx->stream;
    const int64_t off[2] = {0, N};
    if (stage_points(s, a, b, N, 1)) return -1;
    RansacIO io;
    io.dA = s.dA; io.dB = s.dB; io.hOff = off; io.nPairs = 1;
    io.dModels = s.dModels; io.dMasks = s.dMasks; io.dInlierCounts = s.dCnt; io.dOk = s.dOk;
    if (ransac_batch_device(ctx, s.st, prm, io)) return -1;
    uint8_t ok = 0;
    int cnt = 0;
    double m[9];
    MCV_CUDA_TRY(cudaMemcpyAsync(&ok, s.dOk, 1, cudaMemcpyDeviceToHost, s.st));
    MCV_CUDA_TRY(cudaMemcpyAsync(&cnt, s.dCnt, 4, cudaMemcpyDeviceToHost, s.st));
    MCV_CUDA_TRY(cudaMemcpyAsync(m, s.dModels, 72, cudaMemcpyDeviceToHost, s.st));
    MCV_CUDA_TRY(cudaStreamSynchronize(s.st));
    if (!ok) return 0;
    if (mask) {
        MCV_CUDA_TRY(cudaMemcpyAsync(mask, s.dMasks, (size_t)N, cudaMemcpyDeviceToHost, s.st));
        MCV_CUDA_TRY(cudaStreamSynchronize(s.st));
    }
    memcpy(model->M, m, 72);
    if (inliers) *inliers = cnt;
    return 1;
}

}  // namespace

extern "C" {

MCV_API int cvFindHomography(const McvPoint2d* a, const McvPoint2d* b, int N, double thr, int maxIters,
                             double conf, McvMatx33d* H, uint8_t* mask) {
    MCV_API_ENTER();
    RansacParams prm;
    prm.model = RM_HOMOGRAPHY; prm.threshold = thr; prm.maxIters = maxIters; prm.confidence = conf;
    prm.refine = 1;
    const int rc = single_ransac(prm, a, b, N, H, mask, nullptr);
    return rc < 0 ? 0 : rc;
}

MCV_API int cvFindFundamentalMat(const McvPoint2d* a, const McvPoint2d* b, int N, double thr, double conf,
                                 int maxIters, McvMatx33d* F, uint8_t* mask) {
    MCV_API_ENTER();
    // cv::findFundamentalMat's dispatch (calib3d/src/fundam.cpp): fewer than 7 points -> nothing; exactly 7 -> the
    // 7-point kernel on all of them, every real solution returned (up to three, stacked), mask all ones; a
    // non-positive threshold becomes 3, a confidence outside (eps, 1 - eps) becomes 0.99; below 15 points the
    // FM_RANSAC method runs LMedS, from 15 points on RANSAC.
    if (N < 0 || !a || !b || !F) { set_error("cvFindFundamentalMat: bad arguments"); return 0; }
    if (N < 7) return 0;
    if (thr <= 0) thr = 3;
    if (conf < DBL_EPSILON || conf > 1 - DBL_EPSILON) conf = 0.99;
    RansacParams prm;
    prm.model = RM_FUNDAMENTAL; prm.threshold = thr; prm.maxIters = maxIters; prm.confidence = conf;
    if (N == 7) {
        McvRansacParams q = {};
        q.model = RM_FUNDAMENTAL; q.sampleMode = MCV_SAMPLE_OPENCV_MWC; q.stopMode = MCV_STOP_FIXED; q.maxIters = 1;
        q.threshold = thr; q.confidence = conf;
        const int64_t off[2] = {0, 7};
        McvMatx33d best;
        uint8_t ok = 0;
        int32_t counts[3] = {-1, -1, -1};
        double models[27];
        if (cvRansacBatch(&q, a, b, off, 1, nullptr, nullptr, &best, nullptr, nullptr, &ok, nullptr, nullptr, counts,
                          models, nullptr, -1) < 0 || !ok)
            return 0;
        int k = 0;
        for (int m = 0; m < 3; ++m)
            if (counts[m] >= 0) memcpy(F[k++].M, models + 9 * m, 72);
        if (k == 0) return 0;
        if (mask) memset(mask, 1, 7);
        return k;
    }
    if (N < 15) prm.stopMode = MCV_STOP_LMEDS;
    const int rc = single_ransac(prm, a, b, N, F, mask, nullptr);
    return rc < 0 ? 0 : rc;
}

MCV_API int cvRansacBatchDevice(const McvRansacParams* p, const double* dA, const double* dB,
                                const int64_t* offsets, int nPairs, double* dModels, uint8_t* dMasks,
                                int32_t* dInlierCounts, uint8_t* dOk, int device, void* stream) {
    MCV_API_ENTER();
    if (!p || check_batch_args(dA, dB, offsets, nPairs)) return -1;
    if (nPairs == 0) return 0;
    Context* ctx = get_context_async(device, static_cast<cudaStream_t>(stream));
    if (!ctx) return -1;
    RansacIO io;
    io.dA = dA; io.dB = dB; io.hOff = offsets; io.nPairs = nPairs;
    io.dModels = dModels; io.dMasks = dMasks; io.dInlierCounts = dInlierCounts; io.dOk = dOk;
    return ransac_batch_device(ctx, static_cast<cudaStream_t>(stream), to_params(p), io);
}

MCV_API int cvRansacBatch(const McvRansacParams* p, const McvPoint2d* a, const McvPoint2d* b,
                          const int64_t* offsets, int nPairs, const int32_t* injSamples,
                          const double* injModels, McvMatx33d* models, uint8_t* masks,
                          int32_t* inlierCounts, uint8_t* ok, int32_t* itersRun, int32_t* best,
                          int32_t* dbgCounts, double* dbgModels, int32_t* dbgSamples, int device) {
    MCV_API_ENTER();
    if (!p || check_batch_args(a, b, offsets, nPairs)) return -1;
    if (nPairs == 0) return 0;
    if (!models || !ok) { set_error("cvRansacBatch: null outputs"); return -1; }
    Context* ctx = get_context(device);
    if (!ctx) return -1;
    const RansacParams prm = to_params(p);
    if (prm.model < 0 || prm.model > 2) { set_error("unknown model"); return -1; }
    const int mp = model_points(prm.model), mm = max_models(prm.model);
    const int maxIters = std::max(prm.maxIters, 1);
    const int64_t total = offsets[nPairs];
    Staged s; s.ctx = ctx; s.st = ctx->stream;
    const bool taps = injSamples || injModels || dbgCounts || dbgModels || dbgSamples;
    const std::vector<int> runs = taps ? std::vector<int>{0, nPairs} : pair_runs(offsets, nPairs, host_is_pageable(a) || host_is_pageable(b));
    const int nRuns = (int)runs.size() - 1;
    if (nRuns > 1) {
        // copy of run k + 1 under the estimation of run k
        if (stage_alloc(s, total, nPairs)) return -1;
        cudaEvent_t* ev = event_pool(ctx, nRuns);
        if (!ev) { set_error("event pool"); return -1; }
        std::vector<int64_t> roff;
        for (int r = 0; r < nRuns; ++r) {
            const int p0 = runs[r], p1 = runs[r + 1];
            if (stage_run(s, a, b, offsets, p0, p1, ev[r])) return -1;
            roff.assign(offsets + p0, offsets + p1 + 1);
            for (auto& o : roff) o -= offsets[p0];
            RansacIO io;
            io.dA = s.dA + 2 * offsets[p0]; io.dB = s.dB + 2 * offsets[p0]; io.hOff = roff.data(); io.nPairs = p1 - p0;
            io.dModels = s.dModels + 9 * (size_t)p0; io.dMasks = s.dMasks + offsets[p0]; io.dInlierCounts = s.dCnt + p0;
            io.dOk = s.dOk + p0; io.dItersRun = s.dIters + p0; io.dBest = s.dBest + 2 * (size_t)p0;
            RansacParams pr = prm;
            pr.pairIndexBase += p0;                       // counter-based streams are keyed by the global pair index
            if (ransac_batch_device(ctx, s.st, pr, io)) return -1;
        }
    }
    RansacIO io;
    if (nRuns <= 1 && stage_points(s, a, b, total, nPairs)) return -1;
    io.dA = s.dA; io.dB = s.dB; io.hOff = offsets; io.nPairs = nPairs;
    io.dModels = s.dModels; io.dMasks = s.dMasks; io.dInlierCounts = s.dCnt; io.dOk = s.dOk;
    io.dItersRun = s.dIters; io.dBest = s.dBest;
    const size_t nh = (size_t)nPairs * maxIters;
    if (injSamples) {
        if (ctx->ws[38].reserve(nh * mp * 4)) return -1;
        MCV_CUDA_TRY(cudaMemcpyAsync(ctx->ws[38].p, injSamples, nh * mp * 4, cudaMemcpyHostToDevice, s.st));
        io.dInjSamples = ctx->ws[38].as<int>();
    }
    if (injModels) {
        if (ctx->ws[39].reserve(nh * 72)) return -1;
        MCV_CUDA_TRY(cudaMemcpyAsync(ctx->ws[39].p, injModels, nh * 72, cudaMemcpyHostToDevice, s.st));
        io.dInjModels = ctx->ws[39].as<double>();
    }
    if (dbgCounts) { if (ctx->ws[40].reserve(nh * mm * 4)) return -1; io.dDbgCounts = ctx->ws[40].as<int>(); }
    if (dbgModels) { if (ctx->ws[41].reserve(nh * mm * 72)) return -1; io.dDbgModels = ctx->ws[41].as<double>(); }
    if (dbgSamples) { if (ctx->ws[42].reserve(nh * mp * 4)) return -1; io.dDbgSamples = ctx->ws[42].as<int>(); }
    if (nRuns <= 1 && ransac_batch_device(ctx, s.st, prm, io)) return -1;
    MCV_CUDA_TRY(cudaMemcpyAsync(models, s.dModels, (size_t)nPairs * 72, cudaMemcpyDeviceToHost, s.st));
    MCV_CUDA_TRY(cudaMemcpyAsync(ok, s.dOk, (size_t)nPairs, cudaMemcpyDeviceToHost, s.st));
    if (masks && total > 0) MCV_CUDA_TRY(cudaMemcpyAsync(masks, s.dMasks, (size_t)total, cudaMemcpyDeviceToHost, s.st));
    if (inlierCounts) MCV_CUDA_TRY(cudaMemcpyAsync(inlierCounts, s.dCnt, (size_t)nPairs * 4, cudaMemcpyDeviceToHost, s.st));
    if (itersRun) MCV_CUDA_TRY(cudaMemcpyAsync(itersRun, s.dIters, (size_t)nPairs * 4, cudaMemcpyDeviceToHost, s.st));
    if (best) MCV_CUDA_TRY(cudaMemcpyAsync(best, s.dBest, (size_t)nPairs * 8, cudaMemcpyDeviceToHost, s.st));
    if (dbgCounts) MCV_CUDA_TRY(cudaMemcpyAsync(dbgCounts, io.dDbgCounts, nh * mm * 4, cudaMemcpyDeviceToHost, s.st));
    if (dbgModels) MCV_CUDA_TRY(cudaMemcpyAsync(dbgModels, io.dDbgModels, nh * mm * 72, cudaMemcpyDeviceToHost, s.st));
    if (dbgSamples) MCV_CUDA_TRY(cudaMemcpyAsync(dbgSamples, io.dDbgSamples, nh * mp * 4, cudaMemcpyDeviceToHost, s.st));
    MCV_CUDA_TRY(cudaStreamSynchronize(s.st));
    int n = 0;
    for (int i = 0; i < nPairs; ++i) n += ok[i] ? 1 : 0;
    return n;
}

MCV_API int cvRecoverPoseBatchDevice(const RecoverPoseConfig* cfg, const double* dA, const double* dB,
                                     const int64_t* offsets, int nPairs, int sampleMode, int maxIters,
                                     double* dE, double* dR, double* dT, uint8_t* dMs, int32_t* dGood,
                                     uint8_t* dOk, int device, void* stream) {
    MCV_API_ENTER();
    if (!cfg || check_batch_args(dA, dB, offsets, nPairs)) return -1;
    if (nPairs == 0) return 0;
    if (!dE || !dMs || !dOk) { set_error("cvRecoverPoseBatchDevice: null outputs"); return -1; }
    Context* ctx = get_context_async(device, static_cast<cudaStream_t>(stream));
    if (!ctx) return -1;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    RansacParams prm;
    prm.model = RM_ESSENTIAL; prm.sampleMode = sampleMode; prm.stopMode = MCV_STOP_ADAPTIVE;
    prm.maxIters = maxIters > 0 ? maxIters : 1000;
    prm.threshold = cfg->InlierThreshold; prm.confidence = cfg->Probability;
    prm.focal = cfg->FocalLength; prm.ppx = cfg->PrincipalPoint.X; prm.ppy = cfg->PrincipalPoint.Y;
    RansacIO io;
    io.dA = dA; io.dB = dB; io.hOff = offsets; io.nPairs = nPairs;
    io.dModels = dE; io.dMasks = dMs; io.dOk = dOk;
    if (ransac_batch_device(ctx, st, prm, io)) return -1;
    PoseIO po;
    po.dE = dE; po.dOk = dOk; po.dA = dA; po.dB = dB; po.hOff = offsets; po.nPairs = nPairs;
    po.dMaskIn = dMs; po.focal = prm.focal; po.ppx = prm.ppx; po.ppy = prm.ppy;
    po.dR = dR; po.dTsel = dT; po.dGood = dGood;
    return recover_pose_batch_device(ctx, st, po);
}

MCV_API int cvRecoverPoseBatch(const RecoverPoseConfig* cfg, const McvPoint2d* a, const McvPoint2d* b,
                               const int64_t* offsets, int nPairs, int sampleMode, int maxIters,
                               McvMatx33d* R, McvVec3d* t, uint8_t* ms, int32_t* good, uint8_t* ok,
                               int device) {
    MCV_API_ENTER();
    if (!cfg || check_batch_args(a, b, offsets, nPairs)) return -1;
    if (nPairs == 0) return 0;
    if (!R || !t || !ok) { set_error("cvRecoverPoseBatch: null outputs"); return -1; }
    Context* ctx = get_context(device);
    if (!ctx) return -1;
    const int64_t total = offsets[nPairs];
    Staged s; s.ctx = ctx; s.st = ctx->stream;
    if (ctx->ws[43].reserve((size_t)nPairs * (72 + 24 + 4) + 64)) return -1;
    double* dR = ctx->ws[43].as<double>();
    double* dT = dR + 9 * (size_t)nPairs;
    int* dGood = reinterpret_cast<int*>(dT + 3 * (size_t)nPairs);
    // (one piece: cut into runs of pairs like cvRansacBatch, the pageable C4 batch got slower, 33.7 k -> 31.3 k pairs/s --
    // many more, shorter kernels per run than the H / F path)
    if (stage_points(s, a, b, total, nPairs)) return -1;
    if (cvRecoverPoseBatchDevice(cfg, s.dA, s.dB, offsets, nPairs, sampleMode, maxIters, s.dModels, dR, dT,
                                 s.dMasks, dGood, s.dOk, ctx->device, s.st))
        return -1;
    MCV_CUDA_TRY(cudaMemcpyAsync(ok, s.dOk, (size_t)nPairs, cudaMemcpyDeviceToHost, s.st));
    MCV_CUDA_TRY(cudaMemcpyAsync(R, dR, (size_t)nPairs * 72, cudaMemcpyDeviceToHost, s.st));
    MCV_CUDA_TRY(cudaMemcpyAsync(t, dT, (size_t)nPairs * 24, cudaMemcpyDeviceToHost, s.st));
    if (ms && total > 0) MCV_CUDA_TRY(cudaMemcpyAsync(ms, s.dMasks, (size_t)total, cudaMemcpyDeviceToHost, s.st));
    if (good) MCV_CUDA_TRY(cudaMemcpyAsync(good, dGood, (size_t)nPairs * 4, cudaMemcpyDeviceToHost, s.st));
    MCV_CUDA_TRY(cudaStreamSynchronize(s.st));
    int n = 0;
    for (int i = 0; i < nPairs; ++i) n += ok[i] ? 1 : 0;
    return n;
}

MCV_API int cvRecoverPose(const RecoverPoseConfig* cfg, int N, const McvPoint2d* pa, const McvPoint2d* pb,
                          McvMatx33d* rMat, McvVec3d* tVec, uint8_t* ms) {
    MCV_API_ENTER();
    if (!cfg || !pa || !pb || !rMat || !tVec || N < 5) {
        // the reference asserts inside OpenCV for N < 5 (process abort); we return 0 instead
        set_error("cvRecoverPose: bad arguments or N < 5");
        return 0;
    }
    const int64_t off[2] = {0, N};
    McvMatx33d R; McvVec3d t; int good = 0; uint8_t ok = 0;
    std::vector<uint8_t> mask((size_t)N);
    const int rc = cvRecoverPoseBatch(cfg, pa, pb, off, 1, MCV_SAMPLE_OPENCV_MWC, 1000, &R, &t, mask.data(),
                                      &good, &ok, -1);
    if (rc < 0 || !ok) return 0;
    *rMat = R; *tVec = t;
    if (ms) memcpy(ms, mask.data(), (size_t)N);
    return good;
}

MCV_API int32_t cvRecoverPoses(const RecoverPoseConfig* cfg, int N, const McvPoint2d* pa, const McvPoint2d* pb,
                               McvMatx33d* rMat1, McvMatx33d* rMat2, McvVec3d* tVec, uint8_t* ms) {
    MCV_API_ENTER();
    if (N < 5) return 0;   // reference: `if (N < 5) return false;` (MiniCVNative.cpp:167)
    if (!cfg || !pa || !pb || !rMat1 || !rMat2 || !tVec) { set_error("cvRecoverPoses: null arguments"); return 0; }
    Context* ctx = get_context(-1);
    if (!ctx) return 0;
    Staged s; s.ctx = ctx; s.st = ctx->stream;
    const int64_t off[2] = {0, N};
    if (stage_points(s, pa, pb, N, 1)) return 0;
    RansacParams prm;
    prm.model = RM_ESSENTIAL; prm.maxIters = 1000;
    prm.threshold = cfg->InlierThreshold; prm.confidence = cfg->Probability;
    prm.focal = cfg->FocalLength; prm.ppx = cfg->PrincipalPoint.X; prm.ppy = cfg->PrincipalPoint.Y;
    RansacIO io;
    io.dA = s.dA; io.dB = s.dB; io.hOff = off; io.nPairs = 1;
    io.dModels = s.dModels; io.dMasks = s.dMasks; io.dOk = s.dOk;
    if (ransac_batch_device(ctx, s.st, prm, io)) return 0;
    if (ctx->ws[43].reserve(256)) return 0;
    double* dR1 = ctx->ws[43].as<double>();
    double* dR2 = dR1 + 9;
    double* dT = dR2 + 9;
    PoseIO po;
    po.dE = s.dModels; po.dOk = s.dOk; po.hOff = off; po.nPairs = 1; po.dR1 = dR1; po.dR2 = dR2; po.dT = dT;
    if (recover_pose_batch_device(ctx, s.st, po)) return 0;
    uint8_t ok = 0;
    double out[21];
    if (cudaMemcpyAsync(&ok, s.dOk, 1, cudaMemcpyDeviceToHost, s.st) != cudaSuccess ||
        cudaMemcpyAsync(out, dR1, sizeof(out), cudaMemcpyDeviceToHost, s.st) != cudaSuccess ||
        cudaStreamSynchronize(s.st) != cudaSuccess) { set_error("cvRecoverPoses: copy failed"); return 0; }
    // reference copies the mask before checking E (MiniCVNative.cpp:179-183)
    if (ms) {
        if (cudaMemcpyAsync(ms, s.dMasks, (size_t)N, cudaMemcpyDeviceToHost, s.st) != cudaSuccess ||
            cudaStreamSynchronize(s.st) != cudaSuccess) { set_error("cvRecoverPoses: copy failed"); return 0; }
    }
    if (!ok) return 0;
    memcpy(rMat1->M, out, 72); memcpy(rMat2->M, out + 9, 72); memcpy(tVec->V, out + 18, 24);
    return 1;
}

MCV_API int cvFivePointBatch(const McvPoint2d* pa, const McvPoint2d* pb, int nSets, McvMatx33d* Es,
                             int32_t* counts) {
    MCV_API_ENTER();
    if (nSets < 0 || (nSets > 0 && (!pa || !pb || !Es || !counts))) { set_error("cvFivePointBatch: bad arguments"); return -1; }
    if (nSets == 0) return 0;
    Context* ctx = get_context(-1);
    if (!ctx) return -1;
    cudaStream_t st = ctx->stream;
    const size_t n = (size_t)nSets;
    if (ctx->ws[44].reserve(n * 80 * 2 + n * 720 + n * 4 + 64)) return -1;
    double* dA = ctx->ws[44].as<double>();
    double* dB = dA + 10 * n;
    double* dEs = dB + 10 * n;
    int* dC = reinterpret_cast<int*>(dEs + 90 * n);
    MCV_CUDA_TRY(cudaMemcpyAsync(dA, pa, n * 80, cudaMemcpyHostToDevice, st));
    MCV_CUDA_TRY(cudaMemcpyAsync(dB, pb, n * 80, cudaMemcpyHostToDevice, st));
    if (five_point_device(ctx, st, dA, dB, nSets, dEs, dC)) return -1;
    MCV_CUDA_TRY(cudaMemcpyAsync(counts, dC, n * 4, cudaMemcpyDeviceToHost, st));
    std::vector<double> tmp(90 * n);
    MCV_CUDA_TRY(cudaMemcpyAsync(tmp.data(), dEs, n * 720, cudaMemcpyDeviceToHost, st));
    MCV_CUDA_TRY(cudaStreamSynchronize(st));
    int total = 0;
    for (size_t i = 0; i < n; ++i) {
        memcpy(Es[10 * i].M, tmp.data() + 90 * i, (size_t)counts[i] * 72);
        total += counts[i];
    }
    return total;
}

MCV_API int cvFivePoint(const McvPoint2d* pa, const McvPoint2d* pb, McvMatx33d* Es) {
    if (!pa || !pb || !Es) { clear_error(); set_error("cvFivePoint: null arguments"); return 0; }
    McvMatx33d tmp[10];
    int32_t cnt = 0;
    if (cvFivePointBatch(pa, pb, 1, tmp, &cnt) < 0) return 0;
    for (int i = 0; i < cnt; ++i) Es[i] = tmp[i];
    return cnt;
}

MCV_API int cvDecomposeEssentialMat(const McvMatx33d* E, McvMatx33d* R1, McvMatx33d* R2, McvVec3d* t) {
    MCV_API_ENTER();
    if (!E || !R1 || !R2 || !t) { set_error("cvDecomposeEssentialMat: null arguments"); return 0; }
    Context* ctx = get_context(-1);
    if (!ctx) return 0;
    cudaStream_t st = ctx->stream;
    if (ctx->ws[45].reserve(512)) return 0;
    double* dE = ctx->ws[45].as<double>();
    const int64_t off[2] = {0, 0};
    if (cudaMemcpyAsync(dE, E->M, 72, cudaMemcpyHostToDevice, st) != cudaSuccess) return 0;
    PoseIO po;
    po.dE = dE; po.hOff = off; po.nPairs = 1; po.dR1 = dE + 9; po.dR2 = dE + 18; po.dT = dE + 27;
    if (recover_pose_batch_device(ctx, st, po)) return 0;
    double out[21];
    if (cudaMemcpyAsync(out, dE + 9, sizeof(out), cudaMemcpyDeviceToHost, st) != cudaSuccess ||
        cudaStreamSynchronize(st) != cudaSuccess) return 0;
    memcpy(R1->M, out, 72); memcpy(R2->M, out + 9, 72); memcpy(t->V, out + 18, 24);
    return 1;
}

MCV_API int cvRecoverPoseFromE(const McvMatx33d* E, const McvPoint2d* a, const McvPoint2d* b, int N,
                               double focal, double ppx, double ppy, uint8_t* mask, McvMatx33d* R,
                               McvVec3d* t, int32_t* good4) {
    MCV_API_ENTER();
    if (!E || !a || !b || N <= 0 || !R || !t) { set_error("cvRecoverPoseFromE: bad arguments"); return -1; }
    Context* ctx = get_context(-1);
    if (!ctx) return -1;
    Staged s; s.ctx = ctx; s.st = ctx->stream;
    const int64_t off[2] = {0, N};
    if (stage_points(s, a, b, N, 1)) return -1;
    if (ctx->ws[45].reserve(1024 + (size_t)N)) return -1;
    double* dE = ctx->ws[45].as<double>();
    double* dR = dE + 9;
    double* dT = dR + 9;
    int* dGood = reinterpret_cast<int*>(dT + 3);
    int* dGood4 = dGood + 1;
    uint8_t* dMaskOut = reinterpret_cast<uint8_t*>(dE) + 1024;
    MCV_CUDA_TRY(cudaMemcpyAsync(dE, E->M, 72, cudaMemcpyHostToDevice, s.st));
    if (mask) MCV_CUDA_TRY(cudaMemcpyAsync(s.dMasks, mask, (size_t)N, cudaMemcpyHostToDevice, s.st));
    PoseIO po;
    po.dE = dE; po.dA = s.dA; po.dB = s.dB; po.hOff = off; po.nPairs = 1;
    po.dMaskIn = mask ? s.dMasks : nullptr; po.focal = focal; po.ppx = ppx; po.ppy = ppy;
    po.dR = dR; po.dTsel = dT; po.dGood = dGood; po.dGood4 = dGood4; po.dMaskOut = dMaskOut;
    if (recover_pose_batch_device(ctx, s.st, po)) return -1;
    double out[12];
    int g[5];
    MCV_CUDA_TRY(cudaMemcpyAsync(out, dR, 96, cudaMemcpyDeviceToHost, s.st));
    MCV_CUDA_TRY(cudaMemcpyAsync(g, dGood, 20, cudaMemcpyDeviceToHost, s.st));
    if (mask) MCV_CUDA_TRY(cudaMemcpyAsync(mask, dMaskOut, (size_t)N, cudaMemcpyDeviceToHost, s.st));
    MCV_CUDA_TRY(cudaStreamSynchronize(s.st));
    memcpy(R->M, out, 72); memcpy(t->V, out + 9, 24);
    if (good4) for (int k = 0; k < 4; ++k) good4[k] = g[1 + k];
    return g[0];
}

}  // extern "C"
