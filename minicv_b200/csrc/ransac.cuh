// ransac.cuh -- internal interface of the batched RANSAC engine (H / F / E) and pose recovery.
#pragma once
#include "common.cuh"

namespace mcv {

enum RansacModel { RM_HOMOGRAPHY = 0, RM_FUNDAMENTAL = 1, RM_ESSENTIAL = 2 };

struct RansacParams {
    int      model = RM_HOMOGRAPHY;
    int      sampleMode = MCV_SAMPLE_OPENCV_MWC;
    int      stopMode = MCV_STOP_ADAPTIVE;
    int      maxIters = 2000;
    double   threshold = 3.0;
    double   confidence = 0.995;
    uint64_t seed = 0;
    double   focal = 1.0, ppx = 0.0, ppy = 0.0;   // essential only
    int      refine = 1;                          // homography: OpenCV's refit + LM + re-mask
    int64_t  pairIndexBase = 0;                   // counter-based sampling: global index of pair 0
};

inline int model_points(int model) { return model == RM_HOMOGRAPHY ? 4 : model == RM_FUNDAMENTAL ? 7 : 5; }
inline int max_models(int model) { return model == RM_HOMOGRAPHY ? 1 : model == RM_FUNDAMENTAL ? 3 : 10; }

// Everything device-resident except the offsets (host).  All outputs optional except models/masks.
struct RansacIO {
    const double* dA = nullptr;            // [total][2] points of image 1 (double)
    const double* dB = nullptr;            // [total][2]
    const int64_t* hOff = nullptr;         // host, nPairs+1
    int nPairs = 0;
    const int* dInjSamples = nullptr;      // [nPairs][maxIters][MP] (MCV_SAMPLE_INJECTED)
    const double* dInjModels = nullptr;    // [nPairs][maxIters][9]  (skips the minimal solver)
    double*  dModels = nullptr;            // [nPairs][9]
    uint8_t* dMasks = nullptr;             // [total]
    int*     dInlierCounts = nullptr;      // [nPairs]
    uint8_t* dOk = nullptr;                // [nPairs]
    int*     dItersRun = nullptr;          // [nPairs] iterations OpenCV's loop would have run
    int*     dBest = nullptr;              // [nPairs][2] (iteration, model index) of the winner
    // debug / parity taps (device, optional)
    int*     dDbgCounts = nullptr;         // [nPairs][maxIters][MAXM], -1 where no model
    double*  dDbgModels = nullptr;         // [nPairs][maxIters][MAXM][9]
    int*     dDbgSamples = nullptr;        // [nPairs][maxIters][MP]
};

// per-problem flags written by the samplers and read by select / finalize
constexpr int PAIR_FAIL = 1;     // sampling failed at iteration 0 / N < modelPoints
constexpr int PAIR_DIRECT = 2;   // N == modelPoints: kernel on all points, mask all ones

// Model-independent stages (ransac.cu), shared with the PnP engine (pnp.cu):
//   compact: nModels [np][maxIters] -> iterStart [np][maxIters+1], slotList [np][maxIters*maxm] (dense list of
//            iteration*maxm+model in OpenCV's evaluation order), nTotal [np]
//   select : replays OpenCV's accept rule and RANSACUpdateNumIters over counts [np][maxIters*maxm] (compacted
//            order); best [np][4] = compacted index, inlier count, iterations run, slot
int ransac_compact(cudaStream_t st, int np, const uint8_t* dNModels, const int* dNIter, int maxIters, int maxm,
                   int* dIterStart, int* dSlotList, int* dNTotal);
int ransac_select(cudaStream_t st, const int64_t* dOff, int np, int maxIters, int maxm, int mp, bool adaptive,
                  double confidence, const int* dNIter, const int* dFlags, const int* dIterStart,
                  const int* dSlotList, const int* dCounts, int* dBest);

// Scoring with the tensor-core first-level filter (score_tc.cu): counts [np][slots] over the compacted model lists,
// identical to score_kernel's.  dOff / hOff point at the chunk's first pair (absolute offsets), pairBase = its index.
// Uses ctx workspaces 80..85 (86: the two-phase scoring of ransac.cu).
//   score_tc_mode: 0 = not for this chunk, 1 = tensor cores, 2 = decide on the device (homography only): score_tc_pilot
//   scores the first 32 models of every pair on 128 correspondences and leaves {models that fit, models looked at} in the
//   gate; both scoring kernels are launched and the one the gate does not select returns at once.  Measured
//   (tools/gate_probe.py, scoring stage of 256 pairs x 2000 x 4096, fp32-filter kernel vs tcgen05, by inlier ratio): H 0.3:
//   1.61 / 1.01 ms, 0.5: 2.16 / 1.68, 0.6: 2.29 / 2.12, 0.7: 2.30 / 2.62, 0.85: 2.30 / 3.33 -- the crossover is where every
//   sixth 4-point sample is clean; F 0.3: 3.87 / 2.69, 0.6: 8.51 / 4.50, 0.85: 19.0 / 8.57 -- the tensor engine at every
//   ratio (its survivors are classified in fp32, the other kernel's dense lanes take the fp64 exact test), likewise E.
//   The same gate, with a lower bar (1 fitting model in 32), tells the two-phase adaptive stop whether the horizon is
//   going to shrink at all (ransac.cu).
int score_tc_mode(int model, int64_t maxN, int slots, int np, int smCount);
int* score_tc_pilot(Context* ctx, cudaStream_t st, int model, const float4* pf, const double* pd, const int64_t* dOff, int np,
                    const double* dModels, const int* dSlotList, const int* dNTotal, int slots, float thr2);
// Models [dBegin[p], dNTotal[p]) of every pair's compacted list are scored (dBegin nullptr: from 0); haveFeatures: the
// chunk's operand stages are already in the workspace (second call for the same chunk).
int score_tc_run(Context* ctx, cudaStream_t st, int model, const float4* pf, const double* pd, const int64_t* dOff,
                 const int64_t* hOff, int64_t pairBase, int np, const double* dModels, const int* dSlotList,
                 const int* dNTotal, int slots, float thr2, int* dCounts, const int* gate, const int* dBegin,
                 bool haveFeatures);
// gate = {models that fit, models looked at}
__device__ __forceinline__ bool score_gate_classic(const int* gate) { return gate[0] * 6 > gate[1]; }    // engine (H)
__device__ __forceinline__ bool score_gate_split(const int* gate) { return gate[0] * 32 > gate[1]; }     // two-phase stop

// Asynchronous on `stream`.  Uses ctx workspaces 10..21 and pinned buffer 1.
int ransac_batch_device(Context* ctx, cudaStream_t stream, const RansacParams& prm, const RansacIO& io);

// decomposeEssentialMat + recoverPose for a batch (device resident).
//   dE [nPairs][9], dOk [nPairs]; points as in RansacIO; dMaskIn [total] (RANSAC mask)
//   outputs: dR1,dR2 [nPairs][9], dT [nPairs][3] (decomposition), dR,dTsel (chosen pose),
//            dGood [nPairs] (cheirality count).  Any output may be null.
struct PoseIO {
    const double* dE = nullptr;
    const uint8_t* dOk = nullptr;
    const double* dA = nullptr;
    const double* dB = nullptr;
    const int64_t* hOff = nullptr;
    int nPairs = 0;
    const uint8_t* dMaskIn = nullptr;
    double focal = 1.0, ppx = 0.0, ppy = 0.0;
    double* dR1 = nullptr; double* dR2 = nullptr; double* dT = nullptr;
    double* dR = nullptr; double* dTsel = nullptr; int* dGood = nullptr;
    int* dGood4 = nullptr;                 // [nPairs][4] candidate counts
    uint8_t* dMaskOut = nullptr;           // [total] recoverPose's output mask (optional)
};
int recover_pose_batch_device(Context* ctx, cudaStream_t stream, const PoseIO& io);

// cvFivePoint on the device: 5 correspondences -> <= 10 essential matrices.
int five_point_device(Context* ctx, cudaStream_t stream, const double* dA, const double* dB,
                      int nSets, double* dEs /*[nSets][10][9]*/, int* dCounts /*[nSets]*/);

}  // namespace mcv

namespace mcv {

// ---- PnP RANSAC (pnp.cu): cv::solvePnPRansac(SOLVEPNP_AP3P) semantics, batched ------------------------
struct PnpParams {
    int      sampleMode = MCV_SAMPLE_OPENCV_MWC;
    int      stopMode = MCV_STOP_ADAPTIVE;
    int      maxIters = 100;
    double   threshold = 8.0;      // reprojectionError (pixels)
    double   confidence = 0.99;
    uint64_t seed = 0;
    int      refine = 1;           // final pose = EPnP over the inliers (what solvePnPRansac returns)
    int64_t  problemIndexBase = 0; // counter-based sampling: global index of problem 0
};

struct PnpIO {
    const double* dObj = nullptr;          // [total][3] world points
    const double* dImg = nullptr;          // [total][2] image points (pixels)
    const int64_t* hOff = nullptr;         // host, nProblems+1
    int nProblems = 0;
    const double* dCam = nullptr;          // [nProblems][8]: fx, fy, cx, cy, k1, k2, p1, p2
    const int* dInjSamples = nullptr;      // [nProblems][maxIters][4]
    const double* dInjModels = nullptr;    // [nProblems][maxIters][6] (rvec, tvec), scored as given
    double*  dRvec = nullptr;              // [nProblems][3]
    double*  dTvec = nullptr;              // [nProblems][3]
    uint8_t* dMasks = nullptr;             // [total] RANSAC mask of the best minimal model
    int*     dInlierIdx = nullptr;         // [total]: problem p's inlier indices, ascending, from hOff[p] (optional)
    int*     dInlierCounts = nullptr;      // [nProblems]
    uint8_t* dOk = nullptr;                // [nProblems]
    int*     dItersRun = nullptr;          // [nProblems] (optional)
    int*     dBestIter = nullptr;          // [nProblems] iteration of the winner (optional)
    double*  dMinimal = nullptr;           // [nProblems][6] the winning minimal model (optional)
    int*     dDbgCounts = nullptr;         // [nProblems][maxIters], -1 where the solver returned no pose
    double*  dDbgModels = nullptr;         // [nProblems][maxIters][6]
    int*     dDbgSamples = nullptr;        // [nProblems][maxIters][4]
};
// Asynchronous on `stream`.  Uses ctx workspaces 48..61 and pinned buffer 1.
int pnp_ransac_batch_device(Context* ctx, cudaStream_t stream, const PnpParams& prm, const PnpIO& io);

// The reference's solveAp3p (ap3p.cpp:282) for a batch: img [n][3][2] pixels, obj [n][3][3], cam4 [n][4] =
// inv_fx, inv_fy, cx_fx, cy_fy -> Rs [n][4][9], ts [n][4][3], counts [n].  All device pointers.
int ap3p_batch_device(Context* ctx, cudaStream_t stream, const double* dImg, const double* dObj, const double* dCam4,
                      int n, double* dRs, double* dTs, int* dCounts);

}  // namespace mcv

namespace mcv {

// ---- scale search (scale_search.cu): CameraPose.findScaled (src/MiniCV/CameraPose.fs:39-134), batched ----
struct ScaleIO {
    const double* dCams = nullptr;         // [nProblems][14]: location, forward, up, right, focal (Camera.fs:7-14)
    const double* dRot = nullptr;          // [nProblems][9]  pose.Rotation, row-major
    const double* dTrans = nullptr;        // [nProblems][3]  pose.Translation
    const double* dWorld = nullptr;        // [total][3]
    const double* dObs = nullptr;          // [total][2]
    const int64_t* hOff = nullptr;         // host, nProblems+1
    int nProblems = 0;
    double* dCost = nullptr;               // [nProblems] best cost (+inf when no candidate projects anything)
    double* dBestScale = nullptr;          // [nProblems]
    double* dScaledT = nullptr;            // [nProblems][3] bestScale * Translation
    int*    dSign = nullptr;               // [nProblems] sign(bestScale)
    // parity taps (device, optional): candidate c = 2 i + {0: from x, 1: from y} of observation i
    double*  dCand = nullptr;              // [2 total]
    uint8_t* dCandOk = nullptr;            // [2 total] 0 where Fun.IsTiny rejected the observation
    double*  dCandCost = nullptr;          // [2 total] (+inf where rejected)
};
// Asynchronous on `stream`.  Uses ctx workspaces 64..65, pinned buffer 4, event 5.
int find_scaled_batch_device(Context* ctx, cudaStream_t stream, const ScaleIO& io);
// avgReprojectionError (CameraPose.fs:75-90) of given scales: problem p owns scales [dScaleOff[p], dScaleOff[p+1]).
int scale_costs_device(Context* ctx, cudaStream_t st, const double* dCams, const double* dRot, const double* dTrans,
                       const double* dWorld, const double* dObs, const int64_t* dOff, int nProblems,
                       const double* dScales, const uint8_t* dScaleOk, const int64_t* dScaleOff, int64_t maxScales,
                       double* dCosts);

}  // namespace mcv
