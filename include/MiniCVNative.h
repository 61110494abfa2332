/*
 * MiniCVNative.h -- C ABI of the B200-native drop-in for MiniCV's native library.
 *
 * Library identity: the F# layer binds `[<Literal>] lib = "MiniCVNative"`
 * (reference src/MiniCV/OpenCV.fs:340-341), so this header describes
 * libMiniCVNative.so.  Symbol convention follows the reference's DllExport macro
 * (src/MiniCVNative/stdafx.h:20-26): extern "C", default visibility, unmangled,
 * everything else hidden.  All buffers are caller-allocated and caller-owned; the
 * callee reads inputs during the call and writes outputs before it returns; no
 * pointer is retained.  No export throws; failure is 0 / -1 with outputs untouched
 * (the reference's convention).  Every export may be called from any
 * number of host threads at once, leaves the caller's current CUDA device as it found it, and
 * the asynchronous *Device exports may be issued on any streams in any order: scratch memory is
 * leased per call and ordered on the device behind its previous user (common.cuh).
 *
 * Three groups of exports:
 *   (1) EXISTING exports on the hot path, re-implemented on the GPU with their
 *       reference signatures unchanged: cvRecoverPose, cvRecoverPoses, cvFivePoint.
 *   (2) NEW exports the north-star path needs and the reference does not ship
 *       (no matcher, no findHomography/findFundamentalMat in the reference):
 *       cvMatchKnn2*, cvFindHomography*, cvFindFundamentalMat*, batched forms and
 *       the test-only *WithHypotheses forms used by the parity contract.
 *   (3) The ten remaining reference exports (PnP, detectors, fiducials, cvTest) are
 *       out of scope for acceleration; they are forwarded to a side-by-side original
 *       library so every DllImport keeps resolving.
 *
 * There is NO CPU fallback: every compute export returns the failure value and
 * records an error string (mcvLastError) when no CUDA device is usable.
 */
#ifndef MINICV_NATIVE_H
#define MINICV_NATIVE_H

#include <stdint.h>
#include <stdbool.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(_MSC_VER)
#define MCV_API __declspec(dllexport)
#else
#define MCV_API __attribute__((visibility("default")))
#endif

/* ---- blittable types shared with the F# side -------------------------------- */

/* cv::Point2d / Aardvark V2d: 16 bytes. */
typedef struct { double X, Y; } McvPoint2d;
/* cv::Matx33d / Aardvark M33d: 9 doubles, row-major, 72 bytes. */
typedef struct { double M[9]; } McvMatx33d;
/* cv::Vec3d / Aardvark V3d: 24 bytes. */
typedef struct { double V[3]; } McvVec3d;

/* Reference: src/MiniCVNative/MiniCVNative.cpp:39-46 <-> src/MiniCV/OpenCV.fs:16-36.
 * 40 bytes, sequential, no padding. */
typedef struct {
    double     FocalLength;
    McvPoint2d PrincipalPoint;
    double     Probability;
    double     InlierThreshold;
} RecoverPoseConfig;

/* Reference: src/MiniCVNative/MiniCVNative.h:14-29 (KeyPoint2d, DetectorResult):
 * the descriptor wire format a matcher consumes. */
typedef struct {
    float ptX, ptY;
    float size;
    float angle;
    float response;
    int   octave;
    int   class_id;
} McvKeyPoint2d;

typedef struct {
    int            PointCount;
    int            DescriptorEntries;      /* rows * cols of the descriptor Mat  */
    int            DescriptorElementType;  /* OpenCV depth code: 0 = u8, 5 = f32 */
    McvKeyPoint2d* Points;
    uint8_t*       Descriptors;            /* row-major PointCount x descBytes   */
} McvDetectorResult;

/* OpenCV norm codes accepted by the matcher. */
enum { MCV_NORM_L2 = 4, MCV_NORM_HAMMING = 6 };

/* RANSAC sampling / termination modes for the batched estimators. */
enum {
    MCV_SAMPLE_OPENCV_MWC = 0, /* OpenCV's per-call RNG(-1) multiply-with-carry stream   */
    MCV_SAMPLE_COUNTER    = 1, /* counter-based hash RNG, one hypothesis per thread      */
    MCV_SAMPLE_INJECTED   = 2  /* caller supplies the sample index sets                  */
};
enum {
    MCV_STOP_ADAPTIVE = 0,     /* OpenCV's RANSACUpdateNumIters semantics (confidence)   */
    MCV_STOP_FIXED    = 1,     /* score exactly maxIters hypotheses                      */
    MCV_STOP_LMEDS    = 2      /* not RANSAC: OpenCV's least-median-of-squares registrator (what
                                  cv::findFundamentalMat runs below 15 points): iteration count fixed up
                                  front from `confidence`, median-of-errors score, threshold ignored;
                                  fundamental matrices, at most 32 correspondences per pair              */
};

/* ---- diagnostics ------------------------------------------------------------ */

/* Last error message of the calling thread ("" when none).  Never NULL. */
MCV_API const char* mcvLastError(void);
/* Number of usable CUDA devices (0 when none); never throws. */
MCV_API int mcvDeviceCount(void);
/* Kernels launched by this library since load (all threads); used by bench.py. */
MCV_API int64_t mcvKernelLaunchCount(void);
/* CUDA-event timing of the dominant matcher kernel (the tcgen05 GEMM + top-2 kernel, or the
 * XOR/POPC kernel) of every subsequent matcher launch on `device` (any thread): events
 * are recorded on the launching stream around that kernel only.  Start allocates `capacity`
 * event pairs; Stop synchronises, writes up to `capacity` durations in milliseconds (launch
 * order) and returns how many.  Used by bench.py for the roofline figure. */
MCV_API int mcvKernelTimingStart(int device, int capacity);
MCV_API int mcvKernelTimingStop(int device, float* ms, int capacity);
/* Same, for every instrumented stage: tags[i] names the stage of ms[i] (0 matcher kernel, 1 train-bit expansion,
 * 2 kNN finalize, 3 L2 matcher, 10 point preparation, 11 sampling, 12 minimal solvers, 13 compaction, 14 scoring,
 * 15 selection, 16 finalize / refit, 17 pose recovery, 18 match gather, 19 scale search), in launch order. */
MCV_API int mcvKernelTimingStopTagged(int device, float* ms, int* tags, int capacity);
/* Hamming matcher engine: 0 = auto (default), 1 = integer pipe (XOR + POPC), 2 = tcgen05
 * tensor cores.  Both engines are bit-identical in their results; returns 0, or -1 if bad. */
MCV_API int mcvSetHammingEngine(int engine);
/* L2 matcher engine: 0 = auto (default), 1 = fp32 CUDA cores, 2 = tcgen05 bf16-split GEMM + exact re-scoring. */
MCV_API int mcvSetL2Engine(int engine);
/* Scoring engine of the F / E RANSAC stage: 0 = auto (default), 1 = fp32-filter kernel (one model per thread),
 * 2 = tcgen05: the epipolar residuals of a tile as a TF32 GEMM, one comparison per evaluation in the epilogue, survivors
 * to the fp32 filter and the exact test.  The inlier counts are the exact test's with either engine. */
MCV_API int mcvSetScoreEngine(int engine);
/* Audit mode of the tcgen05 scoring engine (tests): while on, every level-0 value is also recomputed in fp64 and the
 * largest |error| / budget seen since it was switched on is kept; mcvScoreAuditMax returns it (must stay below 1). */
MCV_API int mcvSetScoreAudit(int on);
MCV_API float mcvScoreAuditMax(void);
/* Contexts (stream set + grow-only workspaces) live in a per-device pool and are leased per call, so device
 * memory scales with the number of concurrent calls (at most $MINICV_MAX_CONTEXTS per device, default 8), not
 * with the number of host threads.  mcvTrim frees the workspaces and staging buffers of every idle context and
 * returns how many it trimmed; mcvWorkspaceBytes reports the device (kind 0) or pinned-host (kind 1) bytes
 * currently held; mcvSetWorkspaceLimit makes every call free its context's workspaces on return when they
 * exceed `bytesPerContext` (< 0: keep them, the default); mcvContextCount = contexts created on `device`. */
MCV_API int mcvTrim(void);
MCV_API int64_t mcvWorkspaceBytes(int kind);
MCV_API int mcvSetWorkspaceLimit(int64_t bytesPerContext);
MCV_API int mcvContextCount(int device);

/* ---- host buffers ----------------------------------------------------------------------------
 * The host-pointer exports accept any host memory.  CUDA-pinned or registered buffers are copied by DMA directly;
 * pageable buffers (what a P/Invoke caller's `fixed` arrays are) go through a library-owned ring of pinned slots
 * filled by a small pool of copy threads ($MINICV_COPY_THREADS).  A caller that wants the direct path can take
 * its buffers from mcvAllocPinned or register long-lived arrays once with mcvHostRegister. */
MCV_API void* mcvAllocPinned(int64_t bytes);
MCV_API void mcvFreePinned(void* p);
MCV_API int mcvHostRegister(void* p, int64_t bytes);
MCV_API int mcvHostUnregister(void* p);
/* 0 = stage pageable buffers through the pinned ring (default), 1 = plain cudaMemcpyAsync for everything. */
MCV_API int mcvSetStagingMode(int mode);
MCV_API int mcvCopyThreads(void);
/* Host<->device copy rate of `device` in GB/s: `reps` timed copies of `bytes` (CUDA events, private stream, one
 * untimed copy first).  direction 0 = host->device, 1 = device->host, 2 = both at once (sum).  hostKind 0 = pinned,
 * 1 = pageable through the staging ring, 2 = pageable through plain cudaMemcpyAsync.  < 0 on error.  bench.py runs
 * it on every rank at once to measure the box's aggregate ceiling. */
MCV_API double mcvMeasureCopyBandwidth(int device, int64_t bytes, int reps, int direction, int hostKind);
/* Two-phase form for measuring several devices AT ONCE (one thread or process per device): Create allocates pinned
 * host and device blocks of `bytes`; after the callers have synchronised among themselves, Run keeps the copy
 * engines busy for about `seconds` of wall time (direction 0 host->device, 1 device->host, 2 both, 3 both in the
 * 4 : 1 byte mix of the C2 host-buffer path) and writes outGbs[0] = host->device, outGbs[1] = device->host GB/s. */
MCV_API void* mcvCopyProbeCreate(int device, int64_t bytes);
MCV_API int mcvCopyProbeRun(void* probe, double seconds, int direction, double* outGbs);
MCV_API void mcvCopyProbeDestroy(void* probe);
/* Library build tag: "minicv_b200 sm_100a <date>". */
MCV_API const char* mcvBuildInfo(void);

/* ---- (2) descriptor matching -------------------------------------------------
 * New exports (absent in the reference).  Semantics == cv2.BFMatcher(norm)
 * .knnMatch(query, train, k=2) followed by the Lowe test d1 < ratio*d2 in float32.
 *   query/train : row-major nQuery/nTrain x descBytes (HAMMING: bytes; L2: descBytes
 *                 is the row size in BYTES of float32 rows)
 *   trainIdx    : 2*nQuery ints  (k-th neighbour of query i at [2*i+k]; -1 if absent)
 *   dist        : 2*nQuery floats (Hamming distance stored as float like DMatch;
 *                 FLT_MAX if absent)
 *   ratioPass   : nQuery bytes (1 = passes the ratio test); may be NULL
 * Returns the number of queries passing the ratio test, or -1 on error. */
MCV_API int cvMatchKnn2(const uint8_t* query, int nQuery,
                        const uint8_t* train, int nTrain,
                        int descBytes, int normType, float ratio,
                        int* trainIdx, float* dist, uint8_t* ratioPass);

/* Batched form: pair p matches query rows [queryOffsets[p], queryOffsets[p+1]) against
 * train rows [trainOffsets[p], trainOffsets[p+1]).  Offsets have nPairs+1 entries and
 * count ROWS.  trainIdx is relative to the pair's own train block.  Outputs are indexed
 * by global query row.  device = CUDA ordinal (-1: current/default).
 * Returns total ratio-passing queries or -1. */
MCV_API int cvMatchKnn2Batch(const uint8_t* query, const int64_t* queryOffsets,
                             const uint8_t* train, const int64_t* trainOffsets,
                             int nPairs, int descBytes, int normType, float ratio,
                             int* trainIdx, float* dist, uint8_t* ratioPass, int device);

/* Same, with every pointer (except the offsets, which stay on the host) already in
 * device memory of `device` and the work enqueued on CUDA stream `stream`
 * (cudaStream_t, may be NULL = legacy default stream).  Asynchronous; returns 0 or -1.
 * Rows must be 16-byte aligned when descBytes is a multiple of 16. */
MCV_API int cvMatchKnn2BatchDevice(const uint8_t* dQuery, const int64_t* queryOffsets,
                                   const uint8_t* dTrain, const int64_t* trainOffsets,
                                   int nPairs, int descBytes, int normType, float ratio,
                                   int* dTrainIdx, float* dDist, uint8_t* dRatioPass,
                                   int device, void* stream);

/* Multi-GPU form (SURVEY.md section 8b "optional deviceMask", 8e): the pair list is cut into contiguous blocks of
 * equal work, one per device whose bit is set in deviceMask (bit d = CUDA ordinal d; 0 = every visible device);
 * one host thread and one stream set per device, disjoint output slices, no collective.  Host pointers as in
 * cvMatchKnn2Batch.  Returns the total number of ratio-passing queries or -1. */
MCV_API int cvMatchKnn2BatchMulti(const uint8_t* query, const int64_t* queryOffsets,
                                  const uint8_t* train, const int64_t* trainOffsets,
                                  int nPairs, int descBytes, int normType, float ratio,
                                  int* trainIdx, float* dist, uint8_t* ratioPass, uint32_t deviceMask);

/* Matcher over two detector results in the reference's own container
 * (src/MiniCVNative/MiniCVNative.h:23-29) -- SURVEY.md section 8f next-1. */
MCV_API int cvMatchFeatures(const McvDetectorResult* query, const McvDetectorResult* train,
                            float ratio, int* trainIdx, float* dist, uint8_t* ratioPass);

/* Device-resident hand-off matcher -> estimators: builds the packed
 * correspondence lists of a batch from its match results.  Pair p yields one correspondence per
 * ratio-passing query, in query order: a = query keypoint, b = keypoint of its nearest train row.
 *   dKpQuery / dKpTrain : [totalQ][2] / [totalT][2] doubles (device), rows as in the matcher call
 *   dTrainIdx, dRatioPass: the matcher's device outputs
 *   dA, dB               : [totalQ][2] doubles capacity (device)
 *   outOffsets           : nPairs+1 correspondence offsets (HOST, written before return)
 * Returns the total number of correspondences or -1.  Synchronises `stream` once (the offsets). */
MCV_API int cvGatherMatchesDevice(const double* dKpQuery, const double* dKpTrain,
                                  const int64_t* queryOffsets, const int64_t* trainOffsets, int nPairs,
                                  const int* dTrainIdx, const uint8_t* dRatioPass, double* dA, double* dB,
                                  int64_t* outOffsets, int device, void* stream);

/* The north-star pipeline as ONE host-buffer call: kNN match + ratio test -> homography RANSAC, the hand-off kept on
 * the device (descriptors and keypoints go up once, H and the counts come back).  Per pair p: query / train descriptor
 * rows and keypoints kpQuery / kpTrain (the (x, y) of KeyPoint2d as doubles = the V2d arrays a MiniCV caller holds),
 * rows [offsets[p], offsets[p+1]).  Correspondence k of pair p = its k-th ratio-passing query, in query order:
 * a = that query's keypoint, b = the keypoint of its nearest train row; then == cv2.findHomography(a, b, RANSAC, thr,
 * maxIters=, confidence=).  Outputs: H, ok [nPairs]; optional nMatches, nInliers [nPairs]; optional matchQuery,
 * matchTrain (indices local to the pair) and inlierMask, packed pair after pair with room for one entry per query
 * row.  Returns the number of pairs with ok = 1, or -1. */
MCV_API int cvMatchHomographyBatch(const uint8_t* query, const int64_t* queryOffsets,
                                   const uint8_t* train, const int64_t* trainOffsets,
                                   const McvPoint2d* kpQuery, const McvPoint2d* kpTrain, int nPairs,
                                   int descBytes, float ratio, double thr, int maxIters, double conf,
                                   McvMatx33d* H, uint8_t* ok, int32_t* nMatches, int32_t* nInliers,
                                   int32_t* matchQuery, int32_t* matchTrain, uint8_t* inlierMask, int device);
/* Multi-GPU form (see cvMatchKnn2BatchMulti). */
MCV_API int cvMatchHomographyBatchMulti(const uint8_t* query, const int64_t* queryOffsets,
                                        const uint8_t* train, const int64_t* trainOffsets,
                                        const McvPoint2d* kpQuery, const McvPoint2d* kpTrain, int nPairs,
                                        int descBytes, float ratio, double thr, int maxIters, double conf,
                                        McvMatx33d* H, uint8_t* ok, int32_t* nMatches, int32_t* nInliers,
                                        uint32_t deviceMask);

/* Sharded large-database matching: local top-2 of
 * `nQuery` queries against this rank's train shard, emitted as packed keys
 * (dist << 32) | (globalIdxBase + localIdx), 2 per query (UINT64_MAX if absent);
 * device pointers, asynchronous on `stream`. */
MCV_API int cvMatchKnn2ShardKeysDevice(const uint8_t* dQuery, int nQuery,
                                       const uint8_t* dTrainShard, int nTrainShard,
                                       int64_t globalIdxBase, int descBytes,
                                       uint64_t* dKeys, int device, void* stream);
/* Merge `nShards` gathered key lists ([nShards][nQuery][2], e.g. the ncclAllGather
 * output) into the global top-2 and apply the ratio test.  Device pointers. */
MCV_API int cvMergeShardKeysDevice(const uint64_t* dKeys, int nShards, int nQuery, float ratio,
                                   int64_t* dTrainIdx64, float* dDist, uint8_t* dRatioPass,
                                   int device, void* stream);

/* Large-database matching with the collective inside the library (config C5).
 * (i) In-process: the train set is sharded by contiguous index range over the devices of deviceMask and stays
 * resident; cvLargeDbMatch replicates the queries, takes every shard's local top-2, all-gathers the packed keys with
 * NCCL (loaded at run time: libnccl.so.2, $MINICV_NCCL_LIB; peer copies to the first device when NCCL cannot be
 * loaded or $MINICV_SHARD_TRANSPORT=p2p) and merges them.  trainIdx64 [2 nQuery] are GLOBAL train indices (-1 if
 * absent), dist / ratioPass as in cvMatchKnn2.  Returns the number of ratio-passing queries or -1. */
MCV_API void* cvLargeDbCreate(const uint8_t* train, int64_t nTrain, int descBytes, uint32_t deviceMask);
MCV_API int cvLargeDbMatch(void* db, const uint8_t* query, int nQuery, float ratio, int64_t* trainIdx64,
                           float* dist, uint8_t* ratioPass);
MCV_API int cvLargeDbInfo(void* db, int* nShards, int* usesNccl, int64_t* nTrain);
MCV_API void cvLargeDbDestroy(void* db);
/* (ii) One process per GPU: rank 0 obtains a 128-byte id (mcvShardCommUniqueId) and distributes it by any means;
 * every rank calls mcvShardCommInit (collective), then cvMatchKnn2ShardedDevice = local keys -> ncclAllGather ->
 * merge, asynchronous on `stream`, device pointers, every rank receives the global result. */
MCV_API int mcvNcclVersion(void);
MCV_API int mcvShardCommUniqueId(uint8_t* id128);
MCV_API void* mcvShardCommInit(const uint8_t* id128, int nRanks, int rank, int device);
MCV_API void mcvShardCommDestroy(void* comm);
MCV_API int cvMatchKnn2ShardedDevice(void* comm, const uint8_t* dQuery, int nQuery, const uint8_t* dTrainShard,
                                     int64_t nTrainShard, int64_t globalIdxBase, int descBytes, float ratio,
                                     int64_t* dTrainIdx64, float* dDist, uint8_t* dRatioPass, void* stream);


/* ---- (1) existing hot-path exports, re-implemented on the GPU -------------------------
 * Signatures are the reference's (C++ references are pointers at the ABI).  cvRecoverPoses
 * returns a C++ bool in the reference, read by .NET as a 4-byte BOOL (no MarshalAs on the
 * import, src/MiniCV/OpenCV.fs:347): the implementation returns a zero-extended int32 0/1. */

/* Reference: src/MiniCVNative/MiniCVNative.cpp:197-215.  findEssentialMat(RANSAC, prob, thr,
 * maxIters 1000) -> ms = RANSAC inlier mask (N bytes, 0/1) -> recoverPose -> R, t; returns the
 * number of points passing the cheirality check.  On failure returns 0, outputs untouched. */
MCV_API int cvRecoverPose(const RecoverPoseConfig* config, int N, const McvPoint2d* pa,
                          const McvPoint2d* pb, McvMatx33d* rMat, McvVec3d* tVec, uint8_t* ms);

/* Reference: src/MiniCVNative/MiniCVNative.cpp:165-194.  Same RANSAC; R1, R2, t from
 * decomposeEssentialMat; false if N < 5 or no unique E. */
MCV_API int32_t cvRecoverPoses(const RecoverPoseConfig* config, int N, const McvPoint2d* pa,
                               const McvPoint2d* pb, McvMatx33d* rMat1, McvMatx33d* rMat2,
                               McvVec3d* tVec, uint8_t* ms);

/* Reference: src/MiniCVNative/MiniCVNative.cpp:368-382 -> fivepoint.cpp:233-339.  All real
 * essential matrices through 5 correspondences; Es has room for 10; returns the count. */
MCV_API int cvFivePoint(const McvPoint2d* pa, const McvPoint2d* pb, McvMatx33d* Es);

/* ---- (2) two-view estimation: new exports ---------------------------------------------- */

typedef struct {
    int32_t  model;        /* 0 homography (4-pt), 1 fundamental (7-pt), 2 essential (5-pt)      */
    int32_t  sampleMode;   /* MCV_SAMPLE_*                                                        */
    int32_t  stopMode;     /* MCV_STOP_*                                                          */
    int32_t  maxIters;     /* hypotheses per pair (OpenCV: maxIters)                              */
    double   threshold;    /* reprojection / epipolar threshold in the units of the points        */
    double   confidence;   /* (0,1) exclusive, as OpenCV asserts                                  */
    uint64_t seed;         /* MCV_SAMPLE_COUNTER only                                             */
    double   focal, ppx, ppy; /* essential only (findEssentialMat(focal, pp) overload)            */
    int32_t  refine;       /* homography: 1 = findHomography's refit + 10 LM iterations + re-mask */
    int32_t  pairIndexBase;/* MCV_SAMPLE_COUNTER: global index of pair 0 (the multi-GPU forms set it per block,
                              so a pair draws the same stream whichever device it lands on); else 0         */
} McvRansacParams;

/* == cv2.findHomography(a, b, RANSAC, thr, maxIters=, confidence=): H (h33 = 1) and mask.
 * Returns 1 on success, 0 on failure (outputs untouched). */
MCV_API int cvFindHomography(const McvPoint2d* a, const McvPoint2d* b, int N, double thr,
                             int maxIters, double conf, McvMatx33d* H, uint8_t* mask);
/* == cv2.findFundamentalMat(a, b, FM_RANSAC, thr, conf, maxIters), including OpenCV's dispatch: N < 7 -> 0;
 * N == 7 -> the 7-point kernel on all points, EVERY real solution written to F (room for 3 matrices required),
 * mask all ones, returns the number of solutions (1..3); 8 <= N < 15 -> LMedS (MCV_STOP_LMEDS); N >= 15 -> RANSAC,
 * best 7-point model, no refit.  thr <= 0 is read as 3, a confidence outside (DBL_EPSILON, 1 - DBL_EPSILON) as
 * 0.99.  Returns 1 (or the solution count) on success, 0 on failure (outputs untouched). */
MCV_API int cvFindFundamentalMat(const McvPoint2d* a, const McvPoint2d* b, int N, double thr,
                                 double conf, int maxIters, McvMatx33d* F, uint8_t* mask);

/* Batched estimator over packed correspondences: pair p owns rows [offsets[p], offsets[p+1]).
 * Covers the batched forms and the test-only *WithHypotheses forms of SURVEY.md section 8b:
 *   injSamples : [nPairs][maxIters][modelPoints] sample indices (MCV_SAMPLE_INJECTED) or NULL
 *   injModels  : [nPairs][maxIters][9] model matrices scored as given (skips the solver) or NULL
 * Outputs (host): models [nPairs], masks [total rows], inlierCounts/ok/itersRun [nPairs] and
 * best [nPairs][2] = (iteration, model index) of the winner; optional parity taps
 * dbgCounts [nPairs][maxIters][maxModels] (-1 = no model), dbgModels [...][9],
 * dbgSamples [nPairs][maxIters][modelPoints].  Returns the number of pairs with ok = 1, or -1. */
MCV_API int cvRansacBatch(const McvRansacParams* prm, const McvPoint2d* a, const McvPoint2d* b,
                          const int64_t* offsets, int nPairs, const int32_t* injSamples,
                          const double* injModels, McvMatx33d* models, uint8_t* masks,
                          int32_t* inlierCounts, uint8_t* ok, int32_t* itersRun, int32_t* best,
                          int32_t* dbgCounts, double* dbgModels, int32_t* dbgSamples, int device);
/* Multi-GPU form (see cvMatchKnn2BatchMulti): no injected hypotheses / parity taps. */
MCV_API int cvRansacBatchMulti(const McvRansacParams* prm, const McvPoint2d* a, const McvPoint2d* b,
                               const int64_t* offsets, int nPairs, McvMatx33d* models, uint8_t* masks,
                               int32_t* inlierCounts, uint8_t* ok, int32_t* itersRun, uint32_t deviceMask);
/* Device-resident form (offsets on the host, everything else in device memory; asynchronous). */
MCV_API int cvRansacBatchDevice(const McvRansacParams* prm, const double* dA, const double* dB,
                                const int64_t* offsets, int nPairs, double* dModels, uint8_t* dMasks,
                                int32_t* dInlierCounts, uint8_t* dOk, int device, void* stream);

/* Batched cvRecoverPose (config C4): essential RANSAC + recoverPose per pair.
 * R [nPairs], t [nPairs], ms [total rows] (RANSAC masks), good [nPairs] (cheirality counts),
 * ok [nPairs].  Returns pairs with ok = 1, or -1. */
MCV_API int cvRecoverPoseBatch(const RecoverPoseConfig* config, const McvPoint2d* a,
                               const McvPoint2d* b, const int64_t* offsets, int nPairs,
                               int sampleMode, int maxIters, McvMatx33d* R, McvVec3d* t,
                               uint8_t* ms, int32_t* good, uint8_t* ok, int device);
MCV_API int cvRecoverPoseBatchMulti(const RecoverPoseConfig* config, const McvPoint2d* a,
                                    const McvPoint2d* b, const int64_t* offsets, int nPairs,
                                    int sampleMode, int maxIters, McvMatx33d* R, McvVec3d* t,
                                    uint8_t* ms, int32_t* good, uint8_t* ok, uint32_t deviceMask);
MCV_API int cvRecoverPoseBatchDevice(const RecoverPoseConfig* config, const double* dA,
                                     const double* dB, const int64_t* offsets, int nPairs,
                                     int sampleMode, int maxIters, double* dE, double* dR, double* dT,
                                     uint8_t* dMs, int32_t* dGood, uint8_t* dOk, int device,
                                     void* stream);

/* == cv2.recoverPose(E, a, b, focal=, pp=, mask=) for a given E (parity tap): mask is in/out
 * (NULL = all points), good4 receives the four candidate counts.  Returns the count. */
MCV_API int cvRecoverPoseFromE(const McvMatx33d* E, const McvPoint2d* a, const McvPoint2d* b, int N,
                               double focal, double ppx, double ppy, uint8_t* mask,
                               McvMatx33d* R, McvVec3d* t, int32_t* good4);
/* == cv2.decomposeEssentialMat(E). */
MCV_API int cvDecomposeEssentialMat(const McvMatx33d* E, McvMatx33d* R1, McvMatx33d* R2, McvVec3d* t);
/* Batched cvFivePoint: nSets x 5 correspondences -> Es [nSets][10], counts [nSets]. */
MCV_API int cvFivePointBatch(const McvPoint2d* pa, const McvPoint2d* pb, int nSets,
                             McvMatx33d* Es, int32_t* counts);


/* ---- (2b) PnP RANSAC --------------------------------------------
 * cvSolvePnPRansac and solveAp3p keep the reference's signatures and are re-implemented on the GPU. */
typedef struct { double X, Y, Z; } McvPoint3d;

/* Reference: src/MiniCVNative/MiniCVNative.cpp:93-139 -> cv::solvePnPRansac(world, image, K, dist(4), rvec,
 * tvec, false, iterationsCount, reprojectionError, confidence, inliers, kind).  Native for solverKind 5
 * (SOLVEPNP_AP3P): 4-point AP3P hypotheses from OpenCV's sampling stream, reprojection-error scoring,
 * adaptive stop, final pose = EPnP over the inliers; outInliers (room for N) receives the ascending
 * indices of the best minimal model's inliers.  Every other solverKind is forwarded to the original shim
 * (section 3).  Returns false with *inlierCount = 0 and tVec / rVec untouched on failure. */
MCV_API bool cvSolvePnPRansac(const McvPoint2d* imgPoints, const McvPoint3d* worldPoints, const int N,
                              const McvMatx33d K, const double* distortionCoeffs, const int solverKind,
                              const int iterationsCount, const float reprojectionError,
                              const double confidence, McvVec3d* tVec, McvVec3d* rVec,
                              int* inlierCount, int* outInliers);

/* Reference: src/MiniCVNative/ap3p.cpp:282-317.  Up to four poses through three correspondences;
 * R (room for 4) in the reference's layout (camera point = R^T X + t), returns the count. */
MCV_API int solveAp3p(McvMatx33d* R, McvVec3d* t, double mu0, double mv0, double X0, double Y0, double Z0,
                      double mu1, double mv1, double X1, double Y1, double Z1, double mu2, double mv2,
                      double X2, double Y2, double Z2, double inv_fx, double inv_fy, double cx_fx, double cy_fy);
/* Batched solveAp3p: img [n][3], obj [n][3], cam4 [n][4] = inv_fx, inv_fy, cx_fx, cy_fy ->
 * Rs [n][4], ts [n][4], counts [n].  Returns 0, or -1 on error. */
MCV_API int cvSolveAp3pBatch(const McvPoint2d* img, const McvPoint3d* obj, const double* cam4, int n,
                             McvMatx33d* Rs, McvVec3d* ts, int32_t* counts);

typedef struct {
    int32_t  sampleMode;   /* MCV_SAMPLE_*                                                          */
    int32_t  stopMode;     /* MCV_STOP_*                                                            */
    int32_t  maxIters;     /* iterationsCount                                                       */
    int32_t  refine;       /* 1 = final pose is EPnP over the inliers (cv::solvePnPRansac); 0 = the
                              best minimal model                                                    */
    double   threshold;    /* reprojectionError in pixels                                           */
    double   confidence;   /* (0,1) exclusive                                                       */
    uint64_t seed;         /* MCV_SAMPLE_COUNTER only                                               */
    int64_t  problemIndexBase; /* MCV_SAMPLE_COUNTER: global index of problem 0 (see McvRansacParams)       */
} McvPnpParams;

/* Batched form over packed correspondences: problem p owns rows [offsets[p], offsets[p+1]) of obj / img
 * and the camera cams[8 p ..] = fx, fy, cx, cy, k1, k2, p1, p2.  Also the test-only *WithHypotheses form:
 *   injSamples : [nProblems][maxIters][4] sample indices (MCV_SAMPLE_INJECTED) or NULL
 *   injModels  : [nProblems][maxIters][6] (rvec, tvec) scored as given (skips the solver) or NULL
 * Outputs (host; all but rvecs / tvecs / masks / ok optional): rvecs, tvecs [nProblems]; masks [total rows]
 * (RANSAC mask of the best minimal model); inlierCounts, ok, itersRun, bestIter [nProblems]; minimal
 * [nProblems][6] (the winning minimal model); parity taps dbgCounts [nProblems][maxIters] (-1 = no pose),
 * dbgModels [nProblems][maxIters][6], dbgSamples [nProblems][maxIters][4].
 * Returns the number of problems with ok = 1, or -1. */
MCV_API int cvSolvePnPRansacBatch(const McvPnpParams* prm, const McvPoint3d* obj, const McvPoint2d* img,
                                  const int64_t* offsets, int nProblems, const double* cams,
                                  const int32_t* injSamples, const double* injModels, McvVec3d* rvecs,
                                  McvVec3d* tvecs, uint8_t* masks, int32_t* inlierCounts, uint8_t* ok,
                                  int32_t* itersRun, int32_t* bestIter, double* minimal, int32_t* dbgCounts,
                                  double* dbgModels, int32_t* dbgSamples, int device);
/* Multi-GPU form (see cvMatchKnn2BatchMulti). */
MCV_API int cvSolvePnPRansacBatchMulti(const McvPnpParams* prm, const McvPoint3d* obj, const McvPoint2d* img,
                                       const int64_t* offsets, int nProblems, const double* cams, McvVec3d* rvecs,
                                       McvVec3d* tvecs, uint8_t* masks, int32_t* inlierCounts, uint8_t* ok,
                                       int32_t* itersRun, uint32_t deviceMask);
/* Device-resident form (offsets on the host, everything else in device memory; asynchronous). */
MCV_API int cvSolvePnPRansacBatchDevice(const McvPnpParams* prm, const double* dObj, const double* dImg,
                                        const int64_t* offsets, int nProblems, const double* dCams,
                                        double* dRvecs, double* dTvecs, uint8_t* dMasks,
                                        int32_t* dInlierCounts, uint8_t* dOk, int device, void* stream);

/* ---- (2c) scale search -------------------------------------------
 * CameraPose.findScaled (managed code in the reference: src/MiniCV/CameraPose.fs:39-134, with
 * Camera.transformedView / project1, src/MiniCV/Camera.fs:61-84) moved behind the native boundary: every
 * observation proposes two candidate scales for the pose's unit translation and every candidate is costed
 * against all observations -- O(2 n^2) projections, one candidate per GPU thread, sums in the reference's
 * order.  New exports (the reference has no native counterpart); blittable mirrors of the F# types: */
typedef struct {               /* F# record Camera (Camera.fs:7-14), field order kept */
    double location[3];
    double forward[3];
    double up[3];
    double right[3];
    double focal[2];
} McvCamera;                   /* 112 bytes */
typedef struct {               /* F# struct CameraPose (CameraPose.fs:7-16), sequential layout */
    int32_t RotationIndex;
    int32_t ScaleSign;
    double  Rotation[9];       /* M33d, row-major */
    double  Translation[3];
    int32_t IsInverse;         /* bool marshalled as a 4-byte BOOL */
    int32_t _pad;
} McvCameraPose;               /* 112 bytes */

/* == CameraPose.findScaled inlierThreshold srcCam worldObservations pose: returns the cost and writes
 * CameraPose.scale bestScale pose to *scaled.  n = 0 -> +inf and an all-zero pose (CameraPose.fs:42);
 * NaN on error (outputs untouched).  inlierThreshold is accepted and unused, as in the reference. */
MCV_API double cvFindScaled(double inlierThreshold, const McvCamera* srcCam, const McvPoint3d* world,
                            const McvPoint2d* obs, int n, const McvCameraPose* pose, McvCameraPose* scaled);
/* Batched form: problem p owns rows [offsets[p], offsets[p+1]) of world / obs, camera srcCams[p] and pose
 * poses[p].  Outputs (host): costs, scaled [nProblems]; optional bestScales [nProblems] and parity taps over
 * candidate c = 2 i + {0: x-equation, 1: y-equation} of observation i: candidates, candOk (0 where the
 * reference's Fun.IsTiny test rejects the observation), candCosts (+inf where rejected) [2 * total rows].
 * Returns 0, or -1 on error. */
MCV_API int cvFindScaledBatch(const McvCamera* srcCams, const McvCameraPose* poses, const McvPoint3d* world,
                              const McvPoint2d* obs, const int64_t* offsets, int nProblems, double* costs,
                              McvCameraPose* scaled, double* bestScales, double* candidates, uint8_t* candOk,
                              double* candCosts, int device);
/* Device-resident form (offsets on the host; asynchronous on `stream`): dCams [nProblems][14], dRot
 * [nProblems][9], dTrans [nProblems][3], dWorld [total][3], dObs [total][2] -> dCosts, dBestScales
 * [nProblems], dScaledT [nProblems][3], dSigns [nProblems] (sign of the best scale). */
MCV_API int cvFindScaledBatchDevice(const double* dCams, const double* dRot, const double* dTrans,
                                    const double* dWorld, const double* dObs, const int64_t* offsets,
                                    int nProblems, double* dCosts, double* dBestScales, double* dScaledT,
                                    int32_t* dSigns, int device, void* stream);
/* == avgReprojectionError s (CameraPose.fs:75-90) for nScales given scales of one problem (parity tap for
 * injected scales): costs[k] = mean squared reprojection error under scale scales[k], +inf when no
 * observation projects.  Returns 0, or -1. */
MCV_API int cvScaleCosts(const McvCamera* srcCam, const McvCameraPose* pose, const McvPoint3d* world,
                         const McvPoint2d* obs, int n, const double* scales, int nScales, double* costs);

/* ---- (3) reference exports that are not on the path: forwarded, not implemented -----------
 * Signatures are the reference's (src/MiniCVNative/MiniCVNative.cpp:48,140,152,221,347,384,
 * 439,504; cvSolvePnPRansac for solver kinds other than AP3P).  Each call is forwarded to a side-by-side build of the original shim
 * ($MINICV_ORIGINAL_NATIVE, default "libMiniCVNativeCv.so"); without it the reference's failure
 * value is returned.  See INTEGRATION.md section 3. */
MCV_API bool cvSolvePnP(const McvPoint2d* imgPoints, const McvPoint3d* worldPoints, const int N,
                        const McvMatx33d K, const double* distortionCoeffs, const int solverKind,
                        McvVec3d* tVec, McvVec3d* rVec);
MCV_API void cvRefinePnPLM(const McvPoint2d* imgPoints, const McvPoint3d* worldPoints, const int N,
                           const McvMatx33d K, const double* distortionCoeffs, McvVec3d* tVec, McvVec3d* rVec);
MCV_API void cvRefinePnPVVS(const McvPoint2d* imgPoints, const McvPoint3d* worldPoints, const int N,
                            const McvMatx33d K, const double* distortionCoeffs, McvVec3d* tVec, McvVec3d* rVec);
MCV_API McvDetectorResult* cvDetectFeatures(char* data, int width, int height, int channels, int mode, void* config);
MCV_API void cvFreeFeatures(McvDetectorResult* res);
MCV_API bool cvDetectQRCode(char* data, int width, int height, int channels, int* positions, int* count);
MCV_API bool cvDetectArucoMarkers(char* data, int width, int height, int channels, int* infoCount, void* infos);
MCV_API void cvTest(void);
#ifdef __cplusplus
}
#endif
#endif /* MINICV_NATIVE_H */
