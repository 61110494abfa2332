// knn.cuh -- internal interface of the descriptor kNN (k=2) engines.
#pragma once
#include "common.cuh"

namespace mcv {

// What the launcher should emit.
enum KnnEmit {
    KNN_EMIT_MATCHES = 0,  // trainIdx(int32) / dist / ratioPass, indices local to the pair
    KNN_EMIT_KEYS    = 1   // packed u64 keys (dist << 32) | (globalIdxBase + idx), 2 per query
};

struct KnnArgs {
    const uint8_t* dQuery = nullptr;   // device, row-major
    const uint8_t* dTrain = nullptr;   // device, row-major
    const int64_t* hQOff = nullptr;    // host, nPairs+1 row offsets
    const int64_t* hTOff = nullptr;    // host, nPairs+1 row offsets
    int     nPairs = 0;
    int     descBytes = 0;
    float   ratio = 0.f;
    int     emit = KNN_EMIT_MATCHES;
    int64_t globalIdxBase = 0;         // KNN_EMIT_KEYS only
    int*      dTrainIdx = nullptr;     // 2 per query
    float*    dDist = nullptr;         // 2 per query
    uint8_t*  dRatioPass = nullptr;    // 1 per query, may be null
    uint64_t* dKeys = nullptr;         // KNN_EMIT_KEYS: 2 per query
    int*      dPassCount = nullptr;    // device counter (pre-zeroed by the launcher) or null
};

// Hamming kNN-2 on the integer pipe (XOR + POPC, shared-memory tiled train set).
// Asynchronous on `stream`; uses ctx workspaces 0..3 and pinned buffer 0.
int launch_hamming_knn2(Context* ctx, cudaStream_t stream, const KnnArgs& a);

// Same contract on the tensor cores (tcgen05 int8 +-1 GEMM, hamming_tc.cu).
bool hamming_tc_supported(int descBytes);
int launch_hamming_knn2_tc(Context* ctx, cudaStream_t stream, const KnnArgs& a);
// Packed-accumulator tensor-core engine (two distances per TMEM word, <= 256 bit), hamming_tc2.cu.
bool hamming_tc2_supported(int descBytes);
int launch_hamming_knn2_tc2(Context* ctx, cudaStream_t stream, const KnnArgs& a);
// Engine choice: 0 = auto, 1 = integer pipe (XOR + POPC), 2 = tensor cores (int8),
// 3 = tensor cores with packed accumulators (fp8).
extern std::atomic<int> g_hamming_engine;
// L2 engine choice (mcvSetL2Engine): 0 = auto, 1 = fp32 CUDA cores, 2 = tcgen05 bf16 split + exact re-scoring.
extern std::atomic<int> g_l2_engine;
// Shared second stage: merges nSplits key pairs per query and emits matches or keys.
int launch_knn2_finalize(cudaStream_t stream, const uint64_t* dSplitKeys, int nSplits, int64_t totalQ,
                         const KnnArgs& a);

// L2 kNN-2 for float32 descriptors: fp32 CUDA-core engine (difference form) and the tcgen05
// bf16x2-split GEMM + exact re-scoring engine (l2_tc.cu); launch_l2_knn2 picks by the engine switch.
int launch_l2_knn2(Context* ctx, cudaStream_t stream, const KnnArgs& a);
bool l2_tc_supported(int descBytes);
int launch_l2_knn2_tc(Context* ctx, cudaStream_t stream, const KnnArgs& a);

// Chunked host-buffer matcher (api_match.cu): copy-in, matching and copy-out of consecutive chunks of pairs overlap
// on the context's three streams.  Device results stay in ctx workspaces 6 (trainIdx), 7 (dist), 8 (ratioPass), the
// kernels run on ctx->stream.  Host outputs are optional; countOut (optional) makes the call wait for the kernels.
int match_batch_host(Context* ctx, const uint8_t* query, const int64_t* queryOffsets, const uint8_t* train,
                     const int64_t* trainOffsets, int nPairs, int descBytes, int normType, float ratio, int* trainIdx,
                     float* dist, uint8_t* ratioPass, int* countOut);

// Merge nShards x nQuery x 2 packed keys into global top-2 (+ ratio test).
int launch_merge_keys(Context* ctx, cudaStream_t stream, const uint64_t* dKeys, int nShards,
                      int nQuery, float ratio, int64_t* dTrainIdx64, float* dDist,
                      uint8_t* dRatioPass);

}  // namespace mcv
