// api_match.cu -- C-ABI entry points of the descriptor matcher (new exports; the reference
// has no matcher, see include/MiniCVNative.h group (2)).  Thin host layer: stages buffers,
// drives the kernels in hamming_knn.cu / l2_knn.cu, never computes on the CPU.
#include "knn.cuh"
#include "staging.cuh"

#include <algorithm>
#include <vector>

using namespace mcv;

namespace {

int check_norm(int normType, int descBytes) {
    if (normType == MCV_NORM_HAMMING) return 0;
    if (normType == MCV_NORM_L2) {
        if (descBytes % 4 != 0) { set_error("L2 rows must be float32 (descBytes %% 4 == 0)"); return -1; }
        return 0;
    }
    set_error("unsupported normType %d (6 = HAMMING, 4 = L2)", normType);
    return -1;
}

int dispatch(Context* ctx, cudaStream_t s, int normType, const KnnArgs& a) {
    return normType == MCV_NORM_HAMMING ? launch_hamming_knn2(ctx, s, a) : launch_l2_knn2(ctx, s, a);
}

}  // namespace

namespace mcv {

// The chunked host-buffer matcher behind cvMatchKnn2Batch and cvMatchHomographyBatch: copy-in, matching and copy-out
// of consecutive chunks of pairs overlap on three streams.  Device results stay in ctx workspaces 6 (trainIdx),
// 7 (dist), 8 (ratioPass); the matching kernels run on ctx->stream.  Host outputs are optional (all of trainIdx /
// dist null: nothing is copied back); countOut (optional) receives the number of ratio-passing queries and makes the
// call wait for the kernels.
int match_batch_host(Context* ctx, const uint8_t* query, const int64_t* queryOffsets, const uint8_t* train,
                     const int64_t* trainOffsets, int nPairs, int descBytes, int normType, float ratio, int* trainIdx,
                     float* dist, uint8_t* ratioPass, int* countOut) {
    const int64_t totalQ = queryOffsets[nPairs], totalT = trainOffsets[nPairs];
    const size_t rb = (size_t)descBytes;
    if (ctx->ws[4].reserve((size_t)totalQ * rb + 16) || ctx->ws[5].reserve((size_t)totalT * rb + 16) ||
        ctx->ws[6].reserve((size_t)totalQ * 8) || ctx->ws[7].reserve((size_t)totalQ * 8) ||
        ctx->ws[8].reserve((size_t)totalQ) || ctx->ws[9].reserve(256))
        return -1;
    uint8_t* dQ = ctx->ws[4].as<uint8_t>();
    uint8_t* dT = ctx->ws[5].as<uint8_t>();
    int* dIdx = ctx->ws[6].as<int>();
    float* dDist = ctx->ws[7].as<float>();
    uint8_t* dPass = ctx->ws[8].as<uint8_t>();
    int* dCount = ctx->ws[9].as<int>();

    cudaStream_t sIn = ctx->stream2, sK = ctx->stream, sOut = ctx->stream3;
    MCV_CUDA_TRY(cudaMemsetAsync(dCount, 0, sizeof(int), sK));

    // Caller buffers: CUDA-pinned / registered memory is copied directly.  Pageable memory (what a P/Invoke caller
    // passes) goes through the pinned ring on the way in (staging.cu) and through a pinned bounce buffer on the way
    // out, handed to the caller's arrays chunk by chunk while later chunks are still being matched.
    const bool pageQ = host_is_pageable(query), pageT = host_is_pageable(train);
    const bool wantOut = trainIdx != nullptr && dist != nullptr;
    const size_t outBytes = (size_t)totalQ * (8 + 8 + (ratioPass ? 1 : 0));
    const bool bounceOut = wantOut && outBytes >= (1u << 20) && host_is_pageable(trainIdx);
    int* hIdx = trainIdx; float* hDist = dist; uint8_t* hPass = ratioPass;
    if (bounceOut) {
        if (ctx->pin[6].reserve(outBytes + 64)) return -1;
        hIdx = ctx->pin[6].as<int>();
        hDist = reinterpret_cast<float*>(hIdx + 2 * totalQ);
        hPass = ratioPass ? reinterpret_cast<uint8_t*>(hDist + 2 * totalQ) : nullptr;
    }

    // chunk the pair list so that copy-in, compute and copy-out overlap on three streams
    const size_t chunkTarget = 32u << 20;
    std::vector<int> bounds(1, 0);
    for (int p0 = 0; p0 < nPairs;) {
        int p1 = p0;
        size_t bytes = 0;
        while (p1 < nPairs && (p1 == p0 || bytes < chunkTarget)) {
            bytes += (size_t)(queryOffsets[p1 + 1] - queryOffsets[p1] + trainOffsets[p1 + 1] -
                              trainOffsets[p1]) * rb;
            ++p1;
        }
        bounds.push_back(p1);
        p0 = p1;
    }
    const int nChunks = (int)bounds.size() - 1;
    cudaEvent_t* ev = event_pool(ctx, 3 * nChunks);
    if (!ev) { set_error("cvMatchKnn2Batch: cudaEventCreate failed"); return -1; }
    std::vector<int64_t> qo, to;
    int drained = 0;
    auto drain = [&](int c) {      // chunk c's results: pinned bounce buffer -> the caller's arrays
        const int64_t q0 = queryOffsets[bounds[c]], q1 = queryOffsets[bounds[c + 1]];
        if (q1 == q0) return;
        parallel_memcpy(trainIdx + 2 * q0, hIdx + 2 * q0, (size_t)(q1 - q0) * 8);
        parallel_memcpy(dist + 2 * q0, hDist + 2 * q0, (size_t)(q1 - q0) * 8);
        if (ratioPass) memcpy(ratioPass + q0, hPass + q0, (size_t)(q1 - q0));
    };
    for (int c = 0; c < nChunks; ++c) {
        const int p0 = bounds[c], p1 = bounds[c + 1];
        const int64_t q0 = queryOffsets[p0], q1 = queryOffsets[p1];
        const int64_t t0 = trainOffsets[p0], t1 = trainOffsets[p1];
        // copy-in of chunk c, then its (small) control copy right behind it: the copy engine serves
        // its queue in submission order, so control copies must not sit behind later bulk copies
        {
            const H2dSeg segs[2] = {{dQ + q0 * rb, query + q0 * rb, (size_t)(q1 - q0) * rb, pageQ},
                                    {dT + t0 * rb, train + t0 * rb, (size_t)(t1 - t0) * rb, pageT}};
            if (h2d_copy_segs(ctx, segs, 2, sIn)) return -1;
        }
        MCV_CUDA_TRY(cudaEventRecord(ev[3 * c], sIn));
        MCV_CUDA_TRY(cudaStreamWaitEvent(sK, ev[3 * c], 0));
        if (q1 > q0) {
            qo.resize(p1 - p0 + 1);
            to.resize(p1 - p0 + 1);
            for (int p = p0; p <= p1; ++p) { qo[p - p0] = queryOffsets[p] - q0; to[p - p0] = trainOffsets[p] - t0; }
            KnnArgs a;
            a.dQuery = dQ + q0 * rb; a.dTrain = dT + t0 * rb;
            a.hQOff = qo.data(); a.hTOff = to.data();
            a.nPairs = p1 - p0; a.descBytes = descBytes; a.ratio = ratio;
            a.dTrainIdx = dIdx + 2 * q0; a.dDist = dDist + 2 * q0; a.dRatioPass = dPass + q0;
            a.dPassCount = dCount;
            if (dispatch(ctx, sK, normType, a)) return -1;
            if (wantOut) {
                MCV_CUDA_TRY(cudaEventRecord(ev[3 * c + 1], sK));
                MCV_CUDA_TRY(cudaStreamWaitEvent(sOut, ev[3 * c + 1], 0));
                MCV_CUDA_TRY(cudaMemcpyAsync(hIdx + 2 * q0, dIdx + 2 * q0, (size_t)(q1 - q0) * 8,
                                             cudaMemcpyDeviceToHost, sOut));
                MCV_CUDA_TRY(cudaMemcpyAsync(hDist + 2 * q0, dDist + 2 * q0, (size_t)(q1 - q0) * 8,
                                             cudaMemcpyDeviceToHost, sOut));
                if (ratioPass)
                    MCV_CUDA_TRY(cudaMemcpyAsync(hPass + q0, dPass + q0, (size_t)(q1 - q0),
                                                 cudaMemcpyDeviceToHost, sOut));
            }
        }
        MCV_CUDA_TRY(cudaEventRecord(ev[3 * c + 2], sOut));
        // hand finished chunks to the caller's pageable arrays while the copy engines and the matcher work on later ones
        if (bounceOut) {
            while (drained < c && cudaEventQuery(ev[3 * drained + 2]) == cudaSuccess) drain(drained++);
            cudaGetLastError();      // cudaErrorNotReady is an answer, not a failure
        }
    }
    if (bounceOut) {
        for (; drained < nChunks; ++drained) {
            MCV_CUDA_TRY(cudaEventSynchronize(ev[3 * drained + 2]));
            drain(drained);
        }
    }
    int count = 0;
    if (countOut) {
        MCV_CUDA_TRY(cudaMemcpyAsync(&count, dCount, sizeof(int), cudaMemcpyDeviceToHost, sK));
        MCV_CUDA_TRY(cudaStreamSynchronize(sK));
        *countOut = count;
    }
    if (wantOut) MCV_CUDA_TRY(cudaStreamSynchronize(sOut));
    return 0;
}


}  // namespace mcv

extern "C" {

MCV_API int cvMatchKnn2BatchDevice(const uint8_t* dQuery, const int64_t* queryOffsets,
                                   const uint8_t* dTrain, const int64_t* trainOffsets, int nPairs,
                                   int descBytes, int normType, float ratio, int* dTrainIdx,
                                   float* dDist, uint8_t* dRatioPass, int device, void* stream) {
    MCV_API_ENTER();
    if (nPairs < 0 || !queryOffsets || !trainOffsets || !dTrainIdx || !dDist) {
        set_error("cvMatchKnn2BatchDevice: bad arguments");
        return -1;
    }
    if (nPairs == 0) return 0;
    if (check_norm(normType, descBytes)) return -1;
    Context* ctx = get_context_async(device, static_cast<cudaStream_t>(stream));
    if (!ctx) return -1;
    KnnArgs a;
    a.dQuery = dQuery; a.dTrain = dTrain;
    a.hQOff = queryOffsets; a.hTOff = trainOffsets;
    a.nPairs = nPairs; a.descBytes = descBytes; a.ratio = ratio;
    a.dTrainIdx = dTrainIdx; a.dDist = dDist; a.dRatioPass = dRatioPass;
    return dispatch(ctx, static_cast<cudaStream_t>(stream), normType, a);
}

MCV_API int cvMatchKnn2Batch(const uint8_t* query, const int64_t* queryOffsets,
                             const uint8_t* train, const int64_t* trainOffsets, int nPairs,
                             int descBytes, int normType, float ratio, int* trainIdx, float* dist,
                             uint8_t* ratioPass, int device) {
    MCV_API_ENTER();
    if (nPairs < 0 || !queryOffsets || !trainOffsets) {
        set_error("cvMatchKnn2Batch: bad arguments");
        return -1;
    }
    if (nPairs == 0 || queryOffsets[nPairs] == 0) return 0;
    if (!trainIdx || !dist) { set_error("cvMatchKnn2Batch: null output buffers"); return -1; }
    if (check_norm(normType, descBytes)) return -1;
    if (queryOffsets[0] != 0 || trainOffsets[0] != 0) { set_error("offsets must start at 0"); return -1; }
    const int64_t totalQ = queryOffsets[nPairs], totalT = trainOffsets[nPairs];
    if (totalQ < 0 || totalT < 0 || (totalQ > 0 && !query) || (totalT > 0 && !train)) {
        set_error("cvMatchKnn2Batch: bad offsets / null data");
        return -1;
    }
    if (totalQ == 0) return 0;
    Context* ctx = get_context(device);
    if (!ctx) return -1;

    int count = 0;
    if (match_batch_host(ctx, query, queryOffsets, train, trainOffsets, nPairs, descBytes, normType, ratio, trainIdx, dist,
                         ratioPass, &count))
        return -1;
    return count;
}

MCV_API int cvMatchKnn2(const uint8_t* query, int nQuery, const uint8_t* train, int nTrain,
                        int descBytes, int normType, float ratio, int* trainIdx, float* dist,
                        uint8_t* ratioPass) {
    if (nQuery < 0 || nTrain < 0) { clear_error(); set_error("cvMatchKnn2: negative count"); return -1; }
    const int64_t qo[2] = {0, nQuery}, to[2] = {0, nTrain};
    return cvMatchKnn2Batch(query, qo, train, to, 1, descBytes, normType, ratio, trainIdx, dist,
                            ratioPass, -1);
}

MCV_API int cvMatchFeatures(const McvDetectorResult* query, const McvDetectorResult* train,
                            float ratio, int* trainIdx, float* dist, uint8_t* ratioPass) {
    MCV_API_ENTER();
    if (!query || !train) { set_error("cvMatchFeatures: null DetectorResult"); return -1; }
    if (query->PointCount <= 0) return 0;
    if (query->DescriptorElementType != train->DescriptorElementType && train->PointCount > 0) {
        set_error("cvMatchFeatures: descriptor element types differ");
        return -1;
    }
    // element sizes by OpenCV depth code (reference src/MiniCV/OpenCV.fs:319-327)
    static const int esz[] = {1, 1, 2, 2, 4, 4, 8};
    const int ty = query->DescriptorElementType;
    if (ty != 0 && ty != 5) { set_error("cvMatchFeatures: only u8 (binary) and f32 descriptors"); return -1; }
    const int dimQ = query->DescriptorEntries / query->PointCount;
    if (train->PointCount > 0 && train->DescriptorEntries / train->PointCount != dimQ) {
        set_error("cvMatchFeatures: descriptor dimensions differ");
        return -1;
    }
    return cvMatchKnn2(query->Descriptors, query->PointCount, train->Descriptors,
                       train->PointCount, dimQ * esz[ty], ty == 0 ? MCV_NORM_HAMMING : MCV_NORM_L2,
                       ratio, trainIdx, dist, ratioPass);
}

MCV_API int cvMatchKnn2ShardKeysDevice(const uint8_t* dQuery, int nQuery, const uint8_t* dTrainShard,
                                       int nTrainShard, int64_t globalIdxBase, int descBytes,
                                       uint64_t* dKeys, int device, void* stream) {
    MCV_API_ENTER();
    if (nQuery < 0 || nTrainShard < 0 || !dKeys) { set_error("cvMatchKnn2ShardKeysDevice: bad arguments"); return -1; }
    if (nQuery == 0) return 0;
    if (globalIdxBase < 0 || globalIdxBase + (int64_t)nTrainShard > (1ll << 32)) {
        set_error("cvMatchKnn2ShardKeysDevice: globalIdxBase + nTrainShard exceeds the 32-bit index field of the packed key");
        return -1;
    }
    Context* ctx = get_context_async(device, static_cast<cudaStream_t>(stream));
    if (!ctx) return -1;
    const int64_t qo[2] = {0, nQuery}, to[2] = {0, nTrainShard};
    KnnArgs a;
    a.dQuery = dQuery; a.dTrain = dTrainShard; a.hQOff = qo; a.hTOff = to;
    a.nPairs = 1; a.descBytes = descBytes; a.emit = KNN_EMIT_KEYS;
    a.globalIdxBase = globalIdxBase; a.dKeys = dKeys;
    return launch_hamming_knn2(ctx, static_cast<cudaStream_t>(stream), a);
}

MCV_API int cvMergeShardKeysDevice(const uint64_t* dKeys, int nShards, int nQuery, float ratio,
                                   int64_t* dTrainIdx64, float* dDist, uint8_t* dRatioPass,
                                   int device, void* stream) {
    MCV_API_ENTER();
    if (!dKeys || nShards <= 0 || nQuery < 0 || !dTrainIdx64 || !dDist) {
        set_error("cvMergeShardKeysDevice: bad arguments");
        return -1;
    }
    Context* ctx = get_context_async(device, static_cast<cudaStream_t>(stream));
    if (!ctx) return -1;
    return launch_merge_keys(ctx, static_cast<cudaStream_t>(stream), dKeys, nShards, nQuery, ratio,
                             dTrainIdx64, dDist, dRatioPass);
}

}  // extern "C"
