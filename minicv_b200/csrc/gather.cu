// gather.cu -- device-resident hand-off between the matcher and the estimators: turns ratio-passing
// matches into packed correspondence lists (SURVEY.md section 8f "next" row 1: keep the data on the
// device between the stages either side of the matcher; in the reference pipeline sketch this is the
// "gather matched V2d pairs" step the F# caller would do on the host).
#include "common.cuh"

#include <vector>

namespace mcv {
namespace {

__global__ void __launch_bounds__(256)
count_pass_kernel(const uint8_t* __restrict__ pass, const int64_t* __restrict__ qOff, int* __restrict__ counts) {
    const int p = blockIdx.x;
    const int64_t q0 = qOff[p], q1 = qOff[p + 1];
    int c = 0;
    for (int64_t i = q0 + threadIdx.x; i < q1; i += blockDim.x) c += pass[i] ? 1 : 0;
    __shared__ int s[8];
    for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xFFFFFFFFu, c, o);
    if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = c;
    __syncthreads();
    if (threadIdx.x == 0) { int t = 0; for (int w = 0; w < 8; ++w) t += s[w]; counts[p] = t; }
}

// ordered compaction: correspondence k of pair p = k-th ratio-passing query (query order kept)
__global__ void __launch_bounds__(256)
gather_kernel(const double2* __restrict__ kq, const double2* __restrict__ kt, const int64_t* __restrict__ qOff,
              const int64_t* __restrict__ tOff, const int64_t* __restrict__ oOff, const int* __restrict__ trainIdx,
              const uint8_t* __restrict__ pass, double2* __restrict__ a, double2* __restrict__ b) {
    const int p = blockIdx.x;
    const int64_t q0 = qOff[p], q1 = qOff[p + 1], t0 = tOff[p];
    __shared__ int sWarp[8];
    __shared__ int sBase;
    if (threadIdx.x == 0) sBase = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int64_t c0 = q0; c0 < q1; c0 += 256) {
        const int64_t i = c0 + threadIdx.x;
        const bool ok = i < q1 && pass[i];
        const unsigned bal = __ballot_sync(0xFFFFFFFFu, ok);
        if (lane == 0) sWarp[warp] = __popc(bal);
        __syncthreads();
        int before = 0, total = 0;
        for (int w = 0; w < 8; ++w) { if (w < warp) before += sWarp[w]; total += sWarp[w]; }
        if (ok) {
            const int64_t dst = oOff[p] + sBase + before + __popc(bal & ((1u << lane) - 1u));
            a[dst] = kq[i];
            b[dst] = kt[t0 + trainIdx[2 * i]];
        }
        __syncthreads();
        if (threadIdx.x == 0) sBase += total;
        __syncthreads();
    }
}

}  // namespace
}  // namespace mcv

using namespace mcv;

extern "C" {

/* see include/MiniCVNative.h */
MCV_API int cvGatherMatchesDevice(const double* dKpQuery, const double* dKpTrain, const int64_t* queryOffsets,
                                  const int64_t* trainOffsets, int nPairs, const int* dTrainIdx,
                                  const uint8_t* dRatioPass, double* dA, double* dB, int64_t* outOffsets,
                                  int device, void* stream) {
    MCV_API_ENTER();
    if (nPairs < 0 || !queryOffsets || !trainOffsets || !outOffsets) { set_error("cvGatherMatchesDevice: bad arguments"); return -1; }
    outOffsets[0] = 0;
    if (nPairs == 0) return 0;
    if (!dKpQuery || !dKpTrain || !dTrainIdx || !dRatioPass || !dA || !dB) { set_error("cvGatherMatchesDevice: null device buffers"); return -1; }
    Context* ctx = get_context_async(device, static_cast<cudaStream_t>(stream));
    if (!ctx) return -1;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const size_t n1 = (size_t)nPairs + 1;
    if (ctx->ws[46].reserve(n1 * 8 * 3 + (size_t)nPairs * 4 + 64) || ctx->pin[3].reserve(n1 * 8 * 3 + (size_t)nPairs * 4 + 64)) return -1;
    MCV_CUDA_TRY(cudaEventSynchronize(ctx->ev[7]));
    int64_t* hQ = ctx->pin[3].as<int64_t>();
    int64_t* hT = hQ + n1;
    int64_t* hO = hT + n1;
    int* hC = reinterpret_cast<int*>(hO + n1);
    memcpy(hQ, queryOffsets, n1 * 8);
    memcpy(hT, trainOffsets, n1 * 8);
    int64_t* dQ = ctx->ws[46].as<int64_t>();
    int64_t* dT = dQ + n1;
    int64_t* dO = dT + n1;
    int* dC = reinterpret_cast<int*>(dO + n1);
    MCV_CUDA_TRY(cudaMemcpyAsync(dQ, hQ, n1 * 16, cudaMemcpyHostToDevice, st));
    count_pass_kernel<<<nPairs, 256, 0, st>>>(dRatioPass, dQ, dC);
    MCV_LAUNCH_CHECK();
    MCV_CUDA_TRY(cudaMemcpyAsync(hC, dC, (size_t)nPairs * 4, cudaMemcpyDeviceToHost, st));
    MCV_CUDA_TRY(cudaStreamSynchronize(st));
    int64_t acc = 0;
    for (int p = 0; p < nPairs; ++p) { hO[p] = acc; acc += hC[p]; }
    hO[nPairs] = acc;
    memcpy(outOffsets, hO, n1 * 8);
    MCV_CUDA_TRY(cudaMemcpyAsync(dO, hO, n1 * 8, cudaMemcpyHostToDevice, st));
    kt_begin(ctx, st, KT_GATHER);
    gather_kernel<<<nPairs, 256, 0, st>>>(reinterpret_cast<const double2*>(dKpQuery), reinterpret_cast<const double2*>(dKpTrain),
                                          dQ, dT, dO, dTrainIdx, dRatioPass, reinterpret_cast<double2*>(dA),
                                          reinterpret_cast<double2*>(dB));
    kt_end(ctx, st);
    MCV_LAUNCH_CHECK();
    MCV_CUDA_TRY(cudaEventRecord(ctx->ev[7], st));
    return (int)std::min<int64_t>(acc, 0x7FFFFFFF);
}

}  // extern "C"
