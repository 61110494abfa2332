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
