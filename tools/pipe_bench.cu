// pipe_bench.cu -- micro-benchmark of the SM pipes that bound this repo's kernels
// (POPC, LOP3/IADD3/VIMNMX on the ALU pipe, FP32 and FP64 FMA/MUL/ADD), in results per
// clock per SM.  Run on the B200 box; output feeds the roofline denominators in DESIGN.md.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/pipe_bench tools/pipe_bench.cu
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdint.h>

constexpr int ITERS = 4096;
constexpr int ILP = 8;

template <int OP>
__global__ void __launch_bounds__(256) k(uint32_t* out, uint32_t seed, long long* cycles) {
    uint32_t a[ILP];
    float f[ILP];
    double d[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) {
        a[i] = seed + threadIdx.x * 7 + i;
        f[i] = (float)(seed + i) * 1.0001f;
        d[i] = (double)(seed + i) * 1.0001;
    }
    const float fm = __int_as_float(0x3f800001 + (seed & 1));
    const double dm = 1.0 + 1e-9 * seed;
    long long t0 = clock64();
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) {
            if (OP == 0) a[i] = __popc(a[i]) + (a[i] << 3);            // POPC + LEA-ish
            if (OP == 1) a[i] = (a[i] ^ seed) & (a[i] >> 1 | seed);       // LOP3 + SHF
            if (OP == 2) a[i] = min(a[i] + seed, max(a[i], seed * 3u));  // VIMNMX + IADD
            if (OP == 3) f[i] = __fmaf_rn(f[i], fm, 0.5f);               // FFMA
            if (OP == 4) f[i] = __fadd_rn(__fmul_rn(f[i], fm), 0.5f);    // FMUL + FADD
            if (OP == 5) d[i] = __fma_rn(d[i], dm, 0.5);                 // DFMA
            if (OP == 6) d[i] = __dadd_rn(__dmul_rn(d[i], dm), 0.5);     // DMUL + DADD
            if (OP == 7) a[i] = __popc(a[i] ^ seed);                       // pure POPC chain (+LOP3)
            if (OP == 8) a[i] = a[i] * 3u + seed;                          // IMAD
        }
    }
    long long t1 = clock64();
    uint32_t r = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) r += a[i] + __float_as_uint(f[i]) + (uint32_t)__double2loint(d[i]);
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int OP> void run(const char* name, int opsPerIter, int sms) {
    uint32_t* out; long long* cyc;
    const int blocks = sms * 8;     // 8 x 256 threads = 64 warps per SM
    cudaMalloc(&out, blocks * 256 * 4);
    cudaMalloc(&cyc, blocks * 8);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<OP><<<blocks, 256>>>(out, 12345, cyc);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    k<OP><<<blocks, 256>>>(out, 12345, cyc);
    cudaEventRecord(e1);
    cudaDeviceSynchronize();
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    long long* h = new long long[blocks];
    cudaMemcpy(h, cyc, blocks * 8, cudaMemcpyDeviceToHost);
    double avg = 0; for (int i = 0; i < blocks; ++i) avg += h[i]; avg /= blocks;
    // every block runs ITERS*ILP dependent groups; 8 blocks share an SM for the whole kernel
    double perThread = (double)ITERS * ILP * opsPerIter;
    double perSMperClk = perThread * 256 * 8 / avg;   // all 8 co-resident blocks overlap
    double gops = perThread * 256.0 * blocks / (ms * 1e-3) / 1e9;
    printf("%-28s %8.2f results/clk/SM (block-cycle based, %d source ops/iter)  %10.1f Gop/s  %.3f ms\n",
           name, perSMperClk, opsPerIter, gops, ms);
    delete[] h; cudaFree(out); cudaFree(cyc);
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    printf("%s  SMs=%d  clock=%d kHz\n", p.name, p.multiProcessorCount, p.clockRate);
    int s = p.multiProcessorCount;
    run<7>("POPC(+LOP3)", 1, s);
    run<0>("POPC+SHL+IADD", 1, s);
    run<1>("LOP3 x2 + SHF", 1, s);
    run<2>("VIMNMX x2 + IADD/IMAD", 1, s);
    run<8>("IMAD", 1, s);
    run<3>("FFMA", 1, s);
    run<4>("FMUL+FADD", 1, s);
    run<5>("DFMA", 1, s);
    run<6>("DMUL+DADD", 1, s);
    return 0;
}
