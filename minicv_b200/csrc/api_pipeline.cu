// api_pipeline.cu -- the north-star pipeline "kNN match + ratio test -> homography RANSAC" as ONE host-buffer export.
//
// A caller on the reference's side of the boundary holds, per image pair, the two descriptor blobs and the two
// keypoint arrays `detectFeatures` returned (reference src/MiniCV/OpenCV.fs:440-535, DetectorResult layout
// src/MiniCVNative/MiniCVNative.h:23-29).  With the separate exports it would match (descriptors up, 16 bytes per query
// back), gather the matched V2d pairs on the host, and call the estimator (correspondences up again).  Here the
// hand-off stays on the device: descriptors and keypoints go up once, chunked so that
// the copies overlap the matcher; the ratio-passing matches are compacted into correspondence lists by gather.cu;
// findHomography's RANSAC + refit runs on them; only H, the counts and -- if asked for -- the match lists and inlier
// masks come back.
#include "knn.cuh"
#include "ransac.cuh"
#include "staging.cuh"

#include <algorithm>
#include <thread>
#include <string>
#include <vector>

using namespace mcv;

extern "C" {

MCV_API int cvMatchHomographyBatch(const uint8_t* query, const int64_t* queryOffsets, const uint8_t* train,
                                   const int64_t* trainOffsets, const McvPoint2d* kpQuery, const McvPoint2d* kpTrain,
                                   int nPairs, int descBytes, float ratio, double thr, int maxIters, double conf,
                                   McvMatx33d* H, uint8_t* ok, int32_t* nMatches, int32_t* nInliers,
                                   int32_t* matchQuery, int32_t* matchTrain, uint8_t* inlierMask, int device) {
    MCV_API_ENTER();
    if (nPairs < 0 || !queryOffsets || !trainOffsets) { set_error("cvMatchHomographyBatch: bad arguments"); return -1; }
    if (nPairs == 0) return 0;
    if (!H || !ok) { set_error("cvMatchHomographyBatch: null outputs"); return -1; }
    if (queryOffsets[0] != 0 || trainOffsets[0] != 0) { set_error("offsets must start at 0"); return -1; }
    for (int p = 0; p < nPairs; ++p)
        if (queryOffsets[p + 1] < queryOffsets[p] || trainOffsets[p + 1] < trainOffsets[p]) { set_error("bad offsets at pair %d", p); return -1; }
    const int64_t totalQ = queryOffsets[nPairs], totalT = trainOffsets[nPairs];
    if ((totalQ > 0 && (!query || !kpQuery)) || (totalT > 0 && (!train || !kpTrain))) { set_error("cvMatchHomographyBatch: null inputs"); return -1; }
    if (descBytes <= 0) { set_error("cvMatchHomographyBatch: bad descriptor size"); return -1; }
    Context* ctx = get_context(device);
    if (!ctx) return -1;
    const size_t np1 = (size_t)nPairs;
    if (totalQ == 0) {
        memset(ok, 0, np1);
        if (nMatches) memset(nMatches, 0, np1 * 4);
        if (nInliers) memset(nInliers, 0, np1 * 4);
        return 0;
    }
    cudaStream_t sK = ctx->stream, sIn = ctx->stream2;
    // 1. descriptors up + matching, results stay on the device (ws 6 / 8)
    if (match_batch_host(ctx, query, queryOffsets, train, trainOffsets, nPairs, descBytes, MCV_NORM_HAMMING, ratio, nullptr,
                         nullptr, nullptr, nullptr))
        return -1;
    // 2. keypoints up (queued behind the descriptor chunks on the copy stream, so the matcher starts first)
    const size_t tq = (size_t)std::max<int64_t>(totalQ, 1), tt = (size_t)std::max<int64_t>(totalT, 1);
    if (ctx->ws[75].reserve(tq * 16) || ctx->ws[76].reserve(tt * 16) || ctx->ws[77].reserve(tq * 16) ||
        ctx->ws[78].reserve(tq * 16) || ctx->ws[79].reserve(np1 * 72 + tq + np1 * 4 + np1 + 256))
        return -1;
    double* dKq = ctx->ws[75].as<double>();
    double* dKt = ctx->ws[76].as<double>();
    double* dA = ctx->ws[77].as<double>();
    double* dB = ctx->ws[78].as<double>();
    double* dModels = ctx->ws[79].as<double>();
    int* dCnt = reinterpret_cast<int*>(dModels + 9 * np1);
    uint8_t* dOk = reinterpret_cast<uint8_t*>(dCnt + np1);
    uint8_t* dMasks = dOk + ((np1 + 15) & ~(size_t)15);
    MCV_CUDA_TRY(cudaMemsetAsync(dModels, 0, np1 * 72, sK));       // pairs without a model report a zero matrix
    if (h2d_copy(ctx, dKq, kpQuery, (size_t)totalQ * 16, sIn, host_is_pageable(kpQuery))) return -1;
    if (totalT > 0 && h2d_copy(ctx, dKt, kpTrain, (size_t)totalT * 16, sIn, host_is_pageable(kpTrain))) return -1;
    MCV_CUDA_TRY(cudaEventRecord(ctx->ev[4], sIn));
    MCV_CUDA_TRY(cudaStreamWaitEvent(sK, ctx->ev[4], 0));
    // 3. ratio-passing matches -> packed correspondence lists (one host synchronisation: the list offsets)
    std::vector<int64_t> cOff(np1 + 1);
    const int nCorr = cvGatherMatchesDevice(dKq, dKt, queryOffsets, trainOffsets, nPairs, ctx->ws[6].as<int>(),
                                            ctx->ws[8].as<uint8_t>(), dA, dB, cOff.data(), ctx->device, sK);
    if (nCorr < 0) return -1;
    // 4. findHomography semantics on every pair (OpenCV's sampling stream, adaptive stop, refit + LM, re-mask)
    McvRansacParams prm = {};
    prm.model = RM_HOMOGRAPHY; prm.sampleMode = MCV_SAMPLE_OPENCV_MWC; prm.stopMode = MCV_STOP_ADAPTIVE;
    prm.maxIters = maxIters; prm.threshold = thr; prm.confidence = conf; prm.refine = 1;
    if (cvRansacBatchDevice(&prm, dA, dB, cOff.data(), nPairs, dModels, dMasks, dCnt, dOk, ctx->device, sK)) return -1;
    // 5. results
    MCV_CUDA_TRY(cudaMemcpyAsync(H, dModels, np1 * 72, cudaMemcpyDeviceToHost, sK));
    MCV_CUDA_TRY(cudaMemcpyAsync(ok, dOk, np1, cudaMemcpyDeviceToHost, sK));
    if (nInliers) MCV_CUDA_TRY(cudaMemcpyAsync(nInliers, dCnt, np1 * 4, cudaMemcpyDeviceToHost, sK));
    if (inlierMask && nCorr > 0) MCV_CUDA_TRY(cudaMemcpyAsync(inlierMask, dMasks, (size_t)cOff[np1], cudaMemcpyDeviceToHost, sK));
    std::vector<int> hIdx;
    std::vector<uint8_t> hPass;
    if (matchQuery || matchTrain) {
        hIdx.resize((size_t)totalQ * 2);
        hPass.resize((size_t)totalQ);
        MCV_CUDA_TRY(cudaMemcpyAsync(hIdx.data(), ctx->ws[6].p, (size_t)totalQ * 8, cudaMemcpyDeviceToHost, sK));
        MCV_CUDA_TRY(cudaMemcpyAsync(hPass.data(), ctx->ws[8].p, (size_t)totalQ, cudaMemcpyDeviceToHost, sK));
    }
    MCV_CUDA_TRY(cudaStreamSynchronize(sK));
    if (nMatches) for (size_t p = 0; p < np1; ++p) nMatches[p] = (int32_t)(cOff[p + 1] - cOff[p]);
    if (matchQuery || matchTrain) {
        // correspondence k of pair p = its k-th ratio-passing query (the order gather.cu uses); indices local to the pair
        for (size_t p = 0; p < np1; ++p) {
            int64_t k = cOff[p];
            for (int64_t i = queryOffsets[p]; i < queryOffsets[p + 1]; ++i)
                if (hPass[(size_t)i]) {
                    if (matchQuery) matchQuery[k] = (int32_t)(i - queryOffsets[p]);
                    if (matchTrain) matchTrain[k] = hIdx[2 * (size_t)i];
                    ++k;
                }
        }
    }
    int n = 0;
    for (size_t p = 0; p < np1; ++p) n += ok[p] ? 1 : 0;
    return n;
}

MCV_API int cvMatchHomographyBatchMulti(const uint8_t* query, const int64_t* queryOffsets, const uint8_t* train,
                                        const int64_t* trainOffsets, const McvPoint2d* kpQuery, const McvPoint2d* kpTrain,
                                        int nPairs, int descBytes, float ratio, double thr, int maxIters, double conf,
                                        McvMatx33d* H, uint8_t* ok, int32_t* nMatches, int32_t* nInliers,
                                        uint32_t deviceMask) {
    MCV_API_ENTER();
    if (nPairs < 0 || !queryOffsets || !trainOffsets) { set_error("cvMatchHomographyBatchMulti: bad arguments"); return -1; }
    if (nPairs == 0) return 0;
    if (!H || !ok) { set_error("cvMatchHomographyBatchMulti: null outputs"); return -1; }
    int nAll = 0;
    if (cudaGetDeviceCount(&nAll) != cudaSuccess || nAll <= 0) {
        cudaGetLastError();
        set_error("no usable CUDA device; libMiniCVNative has no CPU fallback");
        return -1;
    }
    if (deviceMask == 0) deviceMask = nAll >= 32 ? 0xFFFFFFFFu : ((1u << nAll) - 1u);
    std::vector<int> devs;
    for (int d = 0; d < 32; ++d)
        if (deviceMask & (1u << d)) {
            if (d >= nAll) { set_error("deviceMask names device %d but only %d are visible", d, nAll); return -1; }
            devs.push_back(d);
        }
    const int nDev = (int)std::min<size_t>(devs.size(), (size_t)nPairs);
    // contiguous blocks of equal matching work
    std::vector<int> bounds(nDev + 1, nPairs);
    {
        double total = 0;
        std::vector<double> w((size_t)nPairs);
        for (int p = 0; p < nPairs; ++p) {
            w[p] = (double)(queryOffsets[p + 1] - queryOffsets[p]) * (double)(trainOffsets[p + 1] - trainOffsets[p]) + 1.0;
            total += w[p];
        }
        bounds[0] = 0;
        double acc = 0;
        int p = 0;
        for (int d = 1; d < nDev; ++d) {
            while (p < nPairs && acc + w[p] * 0.5 < total * d / nDev) acc += w[p++];
            bounds[d] = p;
        }
    }
    std::vector<int> rc(nDev, 0);
    std::vector<std::string> errs(nDev);
    const size_t rb = (size_t)descBytes;
    auto body = [&](int s) {
        clear_error();
        const int p0 = bounds[s], p1 = bounds[s + 1];
        if (p1 <= p0) return;
        std::vector<int64_t> qo((size_t)(p1 - p0) + 1), to((size_t)(p1 - p0) + 1);
        for (int p = p0; p <= p1; ++p) { qo[p - p0] = queryOffsets[p] - queryOffsets[p0]; to[p - p0] = trainOffsets[p] - trainOffsets[p0]; }
        const int64_t q0 = queryOffsets[p0], t0 = trainOffsets[p0];
        rc[s] = cvMatchHomographyBatch(query ? query + q0 * rb : nullptr, qo.data(), train ? train + t0 * rb : nullptr, to.data(),
                                       kpQuery ? kpQuery + q0 : nullptr, kpTrain ? kpTrain + t0 : nullptr, p1 - p0, descBytes,
                                       ratio, thr, maxIters, conf, H + p0, ok + p0, nMatches ? nMatches + p0 : nullptr,
                                       nInliers ? nInliers + p0 : nullptr, nullptr, nullptr, nullptr, devs[s]);
        if (rc[s] < 0) errs[s] = get_error();
    };
    std::vector<std::thread> th;
    for (int s = 1; s < nDev; ++s) th.emplace_back(body, s);
    body(0);
    for (auto& t : th) t.join();
    int total = 0;
    for (int s = 0; s < nDev; ++s) {
        if (rc[s] < 0) { set_error("device slot %d: %s", s, errs[s].c_str()); return -1; }
        total += rc[s];
    }
    return total;
}

}  // extern "C"
