// mma_bench.cu -- issue-rate floor of tcgen05.mma.kind::i8 (M = 128, K = 32 per instruction) per SM
// for the operand arrangements the matcher can use: A from shared memory (SS) or from TMEM (TS),
// N = 128 or 256, one accumulator (every MMA depends on the previous one) or several round-robin.
// One CTA per SM, one issuing thread, no TMA and no epilogue: the pure tensor-pipe / operand-fetch rate.
// The issuing warp runs a warp-uniform loop with one elected lane, as hamming_tc.cu does.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I minicv_b200/csrc -o build/mma_bench tools/mma_bench.cu
#include "tc_ptx.cuh"
#include <stdio.h>
using namespace mcv::tc;

constexpr int ITERS = 4096;      // MMAs per run

// MODE 0 = SS, 1 = TS;  N = 128 / 256;  NACC = accumulators used round-robin; LDW = warps streaming tcgen05.ld
template <int MODE, int N, int NACC, int LDW, int COMMIT = 0>
__global__ void __launch_bounds__(64 + 32 * 16, 1) mma_kernel(long long* cycles) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    __shared__ uint64_t bar;
    __shared__ uint64_t dummy[2];
    __shared__ uint64_t ready;
    __shared__ uint32_t slot;
    __shared__ volatile int done;
    const int warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) { mbar_init(&bar, 1); mbar_init(&dummy[0], 1 << 20); mbar_init(&dummy[1], 1 << 20); mbar_init(&ready, 1); mbar_arrive(&ready); done = 0; asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    for (int i = threadIdx.x; i < (128 + 256) * 128 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x01FF01FFu;
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tb = slot;
    constexpr uint32_t idesc = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    if (warp == 1) {
        // warp-uniform loop, one elected lane issues (the pattern hamming_tc.cu uses)
        const uint32_t aBase = smem_u32(smem), bBase = smem_u32(smem + 128 * 128);
        const long long t0 = clock64();
        for (int i0 = 0; i0 < ITERS; i0 += 16) {
            if (COMMIT == 4) { mbar_wait(&ready, 0); tc_fence_after(); }    // a wait that never blocks + the fence
            if (COMMIT == 6) mbar_wait(&ready, 0);
            if (COMMIT == 7) tc_fence_after();
            if (COMMIT >= 8) {      // the issuer is busy elsewhere for a while between groups of 16 MMAs
                const long long w0 = clock64();
                while (clock64() - w0 < (COMMIT - 7) * 100) { }
            }
            if (elect_one()) {
#pragma unroll
                for (int u = 0; u < 16; ++u) {
                    const int i = i0 + u;
                    const int k = u & 3;
                    if (COMMIT == 5 && (u & 7) == 0) tc_fence_after();
                    const uint32_t d = tb + (MODE == 1 ? 64u : 0u) + (uint32_t)(((i >> 3) % NACC) * N);
                    const uint64_t db = umma_desc_sw128(bBase + k * 32);
                    const uint32_t acc = (COMMIT && (u & 7) == 0) ? 0u : 1u;
                    if (MODE == 0) tc_mma_i8(d, umma_desc_sw128(aBase + k * 32), db, idesc, acc);
                    else tc_mma_i8_ts(d, tb + (uint32_t)(k * 8), db, idesc, acc);
                    if ((COMMIT == 1 || COMMIT == 2) && (u & 7) == 7) tc_commit(&dummy[0]);
                    if ((COMMIT == 2 || COMMIT == 3) && u == 15) tc_commit(&dummy[1]);
                }
            }
            __syncwarp();
        }
        if (elect_one()) tc_commit(&bar);
        __syncwarp();
        mbar_wait(&bar, 0);
        const long long t1 = clock64();
        if (threadIdx.x == 32) { cycles[blockIdx.x] = t1 - t0; done = 1; }
    } else if (warp >= 2 && warp < 2 + LDW) {
        // background accumulator read-out traffic, as the epilogue produces it
        uint32_t v[32];
        const uint32_t base = tb + ((uint32_t)((warp & 3) * 32) << 16);
        uint32_t acc = 0;
        while (!done) {
            tc_ld32_pack16(base + 64u + (uint32_t)(((warp >> 2) * 64) & 255), v);
            tc_wait_ld();
            acc ^= v[0];
        }
        if (acc == 0x12345678u) cycles[0] = 0;
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tb), "r"(512u) : "memory");
}

template <int MODE, int N, int NACC, int LDW, int COMMIT = 0> void run(int sms, const char* name) {
    long long* cyc; cudaMalloc(&cyc, sms * 8);
    auto kern = mma_kernel<MODE, N, NACC, LDW, COMMIT>;
    const int SMEM = (128 + 256) * 128 + 2048;
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM);
    kern<<<sms, 64 + 32 * 16, SMEM>>>(cyc);
    cudaDeviceSynchronize();
    kern<<<sms, 64 + 32 * 16, SMEM>>>(cyc);
    cudaError_t e = cudaDeviceSynchronize();
    long long* h = new long long[sms];
    cudaMemcpy(h, cyc, sms * 8, cudaMemcpyDeviceToHost);
    double avg = 0; for (int i = 0; i < sms; ++i) avg += h[i]; avg /= sms;
    const double per = avg / ITERS;
    printf("%-44s %7.1f clk/MMA  %7.0f int8 MAC/clk/SM  (ideal %d clk)  %s\n", name, per, 128.0 * N * 32 / per, N / 2,
           e == cudaSuccess ? "" : cudaGetErrorString(e));
    delete[] h; cudaFree(cyc);
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    const int s = p.multiProcessorCount;
    printf("%s  SMs=%d\n", p.name, s);
    run<0, 128, 1, 0>(s, "SS N=128, 1 accumulator");
    run<0, 128, 2, 0>(s, "SS N=128, 2 accumulators round-robin");
    run<0, 128, 4, 0>(s, "SS N=128, 4 accumulators round-robin");
    run<0, 256, 1, 0>(s, "SS N=256, 1 accumulator");
    run<0, 256, 2, 0>(s, "SS N=256, 2 accumulators round-robin");
    run<1, 128, 1, 0>(s, "TS N=128, 1 accumulator");
    run<1, 128, 3, 0>(s, "TS N=128, 3 accumulators round-robin");
    run<1, 256, 1, 0>(s, "TS N=256, 1 accumulator");
    run<0, 128, 2, 8>(s, "SS N=128, 2 acc + 8 warps tcgen05.ld");
    run<0, 128, 2, 16>(s, "SS N=128, 2 acc + 16 warps tcgen05.ld");
    run<1, 128, 3, 12>(s, "TS N=128, 3 acc + 12 warps tcgen05.ld");
    run<0, 256, 1, 8>(s, "SS N=256, 1 acc + 8 warps tcgen05.ld");
    run<1, 128, 3, 0, 1>(s, "TS N=128, 3 acc, commit every 8 MMAs");
    run<1, 128, 3, 0, 3>(s, "TS N=128, 3 acc, commit every 16 MMAs");
    run<1, 128, 3, 0, 2>(s, "TS N=128, 3 acc, commit/8 + second commit/16");
    run<1, 128, 3, 0, 4>(s, "TS N=128, 3 acc, wait(ready)+fence every 16 MMAs");
    run<1, 128, 3, 0, 6>(s, "TS N=128, 3 acc, wait(ready) every 16 MMAs");
    run<1, 128, 3, 0, 7>(s, "TS N=128, 3 acc, fence (all lanes) every 16 MMAs");
    run<1, 128, 3, 0, 5>(s, "TS N=128, 3 acc, fence (elected lane) every 8 MMAs");
    run<1, 128, 3, 0, 8>(s, "TS N=128, issuer away 100 clk per 16 MMAs");
    run<1, 128, 3, 0, 9>(s, "TS N=128, issuer away 200 clk per 16 MMAs");
    run<1, 128, 3, 0, 11>(s, "TS N=128, issuer away 400 clk per 16 MMAs");
    run<1, 128, 3, 0, 15>(s, "TS N=128, issuer away 800 clk per 16 MMAs");
    run<1, 128, 3, 12, 1>(s, "TS N=128, 3 acc, commit/8 + 12 warps ld");
    run<1, 128, 3, 12, 2>(s, "TS N=128, 3 acc, commit/8 + commit/16 + 12 warps ld");
    return 0;
}
