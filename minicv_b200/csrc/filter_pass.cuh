// filter_pass.cuh -- the second half of "filtered exact scoring" (score_kernel<M> in ransac.cu, pnp_score_kernel in
// pnp.cu): each lane of a warp holds one model and a 64-bit mask of the correspondences of the current block that its
// fp32 filter could not reject; those need the exact test.  Wrong models keep ~1 % of a block, a good model keeps its
// inliers (20-60 %), so lane-serial resolution would let one good model stall its 31 neighbours.  Instead:
//   * a lane with at least DENSE_MIN survivors is resolved by the whole warp: its model is broadcast with shuffles and
//     lane k tests correspondences k and k + 32;
//   * the remaining sparse masks are resolved lane-serially (a handful of iterations);
//   * a warp with many dense lanes (every model fits: e.g. scoring after a ratio-tested match) stops filtering for the
//     rest of the launch and runs the plain broadcast loop (warp_too_dense).
#pragma once
#include <cuda_runtime.h>

namespace mcv {

#ifndef FP_DENSE_MIN_V
#define FP_DENSE_MIN_V 6
#endif
constexpr int FP_DENSE_MIN = FP_DENSE_MIN_V;   // survivors (of 64) from which a lane is resolved cooperatively
constexpr int FP_DENSE_LANES = 8;       // dense lanes (of 32) from which the warp gives up filtering for the tile

// Warp-uniform: at least FP_DENSE_LANES lanes of the warp have dense masks -> the caller should stop filtering (the models
// do not change, so it stays that way for the rest of the launch) and run the plain broadcast loop, this block included.
__device__ __forceinline__ bool warp_too_dense(unsigned long long pending) {
    return __popc(__ballot_sync(0xFFFFFFFFu, __popcll(pending) >= FP_DENSE_MIN)) >= FP_DENSE_LANES;
}

// Model: trivially copyable register struct; shfl(model, srcLane) -> Model; test(model, j) -> bool (exact inlier test
// of correspondence j of the block).  Returns this lane's inlier count.
template <class Model, class Shfl, class Test>
__device__ __forceinline__ int resolve_pending(unsigned long long pending, const Model& mine, Shfl shfl, Test test) {
    constexpr unsigned FULL = 0xFFFFFFFFu;
    const int lane = threadIdx.x & 31;
    int cnt = 0;
    unsigned dense = __ballot_sync(FULL, __popcll(pending) >= FP_DENSE_MIN);
    while (dense) {
        const int m = __ffs(dense) - 1;
        dense &= dense - 1;
        const Model mm = shfl(mine, m);
        const unsigned long long pm = __shfl_sync(FULL, pending, m);
        const bool in0 = ((pm >> lane) & 1ull) && test(mm, lane);
        const bool in1 = ((pm >> (lane + 32)) & 1ull) && test(mm, lane + 32);
        const int c = __popc(__ballot_sync(FULL, in0)) + __popc(__ballot_sync(FULL, in1));
        if (lane == m) { cnt += c; pending = 0; }
    }
    while (__any_sync(FULL, pending != 0)) {
        if (pending) {
            const int j = __ffsll((long long)pending) - 1;
            pending &= pending - 1;
            cnt += test(mine, j) ? 1 : 0;
        }
    }
    return cnt;
}

}  // namespace mcv
