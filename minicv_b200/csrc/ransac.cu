// ransac.cu -- batched RANSAC for homography / fundamental / essential matrices.
//
// The reference exports only the essential-matrix form (cvRecoverPose / cvRecoverPoses,
// src/MiniCVNative/MiniCVNative.cpp:165-215 -> cv::findEssentialMat(..., RANSAC, ...)); the
// H and F forms are the new exports the north-star path adds.  All three run through one
// engine that restates OpenCV's RANSACPointSetRegistrator::run as data-parallel stages:
//
//   prep      points -> float32 (H, F) or intrinsics-normalised float64 (E)        [HBM-bound]
//   sample    all maxIters subsets up front: OpenCV's MWC stream (one thread per pair, the
//             stream is sequential by construction), a counter-based hash RNG (one thread per
//             hypothesis) or injected index sets; degeneracy checks as in checkSubset
//   solve     one hypothesis per thread: 4-point H / 7-point F / 5-point E (cvmath.cuh)
//   compact   per pair prefix over the model counts -> dense model list in OpenCV's order
//   score     one model per thread x all correspondences staged through shared memory and
//             read as 128-bit broadcasts; inlier count in a register              [FP pipe bound]
//   select    replays OpenCV's sequential accept / RANSACUpdateNumIters logic over the counts,
//             so the winner is the model OpenCV's loop would have stopped with
//   finalize  inlier mask of the winner; homography: refit on inliers (normalised DLT) +
//             10 Levenberg-Marquardt iterations + mask from the refined H (findHomography)
#include "ransac.cuh"
#include "filter_pass.cuh"
#include "cvmath.cuh"

#include <stdlib.h>
#include <algorithm>
#include <vector>

namespace mcv {
namespace {

using namespace cvm;

template <int MODEL> struct MT;
template <> struct MT<RM_HOMOGRAPHY>  { static constexpr int MP = 4, MAXM = 1; };
template <> struct MT<RM_FUNDAMENTAL> { static constexpr int MP = 7, MAXM = 3; };
template <> struct MT<RM_ESSENTIAL>   { static constexpr int MP = 5, MAXM = 10; };


// ---------------------------------------------------------------------------------------------
// prep: double (x,y) pairs -> working layout
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
prep_points_kernel(const double2* __restrict__ a, const double2* __restrict__ b, int64_t total,
                   int isE, double invf, double nx, double ny, float4* __restrict__ pf,
                   double* __restrict__ pd) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const double2 pa = a[i], pb = b[i];
    if (!isE) {
        pf[i] = make_float4((float)pa.x, (float)pa.y, (float)pb.x, (float)pb.y);
    } else {
        // (x - pp) / f as x*(1/f) + (-pp*(1/f)), unfused (oracle/essential.py normalize_points)
        double* o = pd + 4 * i;
        o[0] = dadd(dmul(pa.x, invf), nx);
        o[1] = dadd(dmul(pa.y, invf), ny);
        o[2] = dadd(dmul(pb.x, invf), nx);
        o[3] = dadd(dmul(pb.y, invf), ny);
    }
}

// ---------------------------------------------------------------------------------------------
// sampling
// ---------------------------------------------------------------------------------------------
template <int MODEL>
__device__ __forceinline__ bool check_subset(const float4* pf, int64_t base, const int* idx) {
    if (MODEL == RM_ESSENTIAL) return true;
    constexpr int MP = MT<MODEL>::MP;
    float s[2 * MP], d[2 * MP];
#pragma unroll
    for (int i = 0; i < MP; ++i) {
        const float4 p = pf[base + idx[i]];
        s[2 * i] = p.x; s[2 * i + 1] = p.y; d[2 * i] = p.z; d[2 * i + 1] = p.w;
    }
    if (MODEL == RM_HOMOGRAPHY) return h_check_subset(s, d);
    return !(have_collinear(s, MP) || have_collinear(d, MP));
}

// OpenCV's stream: RNG((uint64)-1) per call, subsets drawn one after another (getSubset).
// One warp per pair: the warp stages the pair's points in shared memory (the degeneracy checks of
// every draw would otherwise be chains of dependent global loads), then lane 0 replays the
// sequential stream.
template <int MODEL>
__device__ __forceinline__ bool check_subset_pts(const float4* pts, const int* idx) {
    if (MODEL == RM_ESSENTIAL) return true;
    constexpr int MP = MT<MODEL>::MP;
    float s[2 * MP], d[2 * MP];
#pragma unroll
    for (int i = 0; i < MP; ++i) {
        const float4 p = pts[idx[i]];
        s[2 * i] = p.x; s[2 * i + 1] = p.y; d[2 * i] = p.z; d[2 * i + 1] = p.w;
    }
    if (MODEL == RM_HOMOGRAPHY) return h_check_subset(s, d);
    return !(have_collinear(s, MP) || have_collinear(d, MP));
}

template <int MODEL>
__global__ void __launch_bounds__(32)
sample_mwc_kernel(const float4* __restrict__ pf, const int64_t* __restrict__ off, int nPairs,
                  int maxIters, int smemPoints, int* __restrict__ samples, int* __restrict__ nIter,
                  int* __restrict__ flags, const uint8_t* __restrict__ needSeq) {
    constexpr int MP = MT<MODEL>::MP;
    extern __shared__ float4 sPts[];
    const int p = blockIdx.x;
    if (needSeq && !needSeq[p]) return;      // the parallel sampler already did this pair
    const int64_t base = off[p];
    const int n = (int)(off[p + 1] - base);
    const bool staged = MODEL != RM_ESSENTIAL && n <= smemPoints;
    if (staged) for (int i = threadIdx.x; i < n; i += 32) sPts[i] = pf[base + i];
    __syncwarp();
    if (threadIdx.x != 0) return;
    const float4* pts = staged ? sPts : (MODEL == RM_ESSENTIAL ? nullptr : pf + base);
    int* out = samples + (int64_t)p * maxIters * MP;
    if (n < MP) { nIter[p] = 0; flags[p] = PAIR_FAIL; return; }
    if (n == MP) {
        for (int i = 0; i < MP; ++i) out[i] = i;
        nIter[p] = 1; flags[p] = PAIR_DIRECT; return;
    }
    Mwc rng;
    int it = 0;
    int fl = 0;
    for (; it < maxIters; ++it) {
        int idx[MP];
        bool found = false;
        for (int attempt = 0; attempt < 10000; ++attempt) {
            for (int i = 0; i < MP; ++i) {
                int v;
                bool dup;
                do {
                    v = rng.uniform(n);
                    dup = false;
                    for (int k = 0; k < i; ++k) dup |= (idx[k] == v);
                } while (dup);
                idx[i] = v;
            }
            if (check_subset_pts<MODEL>(pts, idx)) { found = true; break; }
        }
        if (!found) { if (it == 0) fl = PAIR_FAIL; break; }
        for (int i = 0; i < MP; ++i) out[it * MP + i] = idx[i];
    }
    nIter[p] = it;
    flags[p] = fl;
}

// The same stream, replayed in parallel.  OpenCV seeds RNG((uint64)-1) at the start of every call, so
// the raw 32-bit draws are one fixed sequence (mwc_table_kernel, once per context); only `% n`, the
// duplicate redraws and the degeneracy checks depend on the pair.  One CTA per pair:
//   A  every thread simulates one *attempt* (draw MP distinct indices, check the subset) from every
//      start position of a block of the stream: attempt length and outcome per position;
//   B  one thread follows the chain position -> position + length, which is exactly the sequential
//      loop (an iteration = attempts until one passes, at most 10 000);
//   C  all threads copy the subsets of the iterations found to the sample list.
// Pairs the table cannot serve (more than 65 536 points, > 64 draws in one attempt, stream longer than
// the table) are flagged in needSeq and replayed by sample_mwc_kernel.
constexpr int MWC_TABLE = 1 << 16;
constexpr int MWC_BLOCK = 4096;
constexpr int MWC_MAXDRAW = 64;

__global__ void mwc_table_kernel(uint32_t* __restrict__ raw) {
    if (blockIdx.x != 0 || threadIdx.x != 0) return;
    Mwc rng;
    for (int i = 0; i < MWC_TABLE; ++i) raw[i] = rng.next();
}

template <int MODEL>
__global__ void __launch_bounds__(256)
sample_mwc_par_kernel(const float4* __restrict__ pf, const int64_t* __restrict__ off, int nPairs,
                      int maxIters, int smemPoints, const uint32_t* __restrict__ raw,
                      int* __restrict__ samples, int* __restrict__ nIter, int* __restrict__ flags,
                      uint8_t* __restrict__ needSeq) {
    constexpr int MP = MT<MODEL>::MP;
    extern __shared__ float4 sDyn[];
    const int p = blockIdx.x;
    const int64_t base = off[p];
    const int n = (int)(off[p + 1] - base);
    int* out = samples + (int64_t)p * maxIters * MP;
    if (n <= MP || n > 65536) {
        if (threadIdx.x == 0) {
            needSeq[p] = n > 65536 ? 1 : 0;
            if (n < MP) { nIter[p] = 0; flags[p] = PAIR_FAIL; }
            else if (n == MP) { for (int i = 0; i < MP; ++i) out[i] = i; nIter[p] = 1; flags[p] = PAIR_DIRECT; }
        }
        return;
    }
    const bool staged = MODEL != RM_ESSENTIAL && n <= smemPoints;
    float4* sPts = sDyn;
    uint32_t* sRaw = reinterpret_cast<uint32_t*>(sDyn + (MODEL == RM_ESSENTIAL ? 0 : smemPoints));
    uint16_t* aIdx = reinterpret_cast<uint16_t*>(sRaw + MWC_BLOCK + MWC_MAXDRAW);
    uint16_t* itPos = aIdx + MWC_BLOCK * MP;
    uint8_t* aLen = reinterpret_cast<uint8_t*>(itPos + MWC_BLOCK / MP + 8);
    __shared__ int sState[4];                 // found this round, stop code, next position, attempts
    if (staged) for (int i = threadIdx.x; i < n; i += blockDim.x) sPts[i] = pf[base + i];
    const float4* pts = staged ? sPts : (MODEL == RM_ESSENTIAL ? nullptr : pf + base);

    int pos = 0, it = 0, attempts = 0, stop = 0;    // stop: 1 = out of hypotheses, 2 = hand over, 3 = no valid subset
    while (it < maxIters && !stop) {
        if (pos + MWC_BLOCK + MWC_MAXDRAW > MWC_TABLE) { stop = 2; break; }
        __syncthreads();
        for (int i = threadIdx.x; i < MWC_BLOCK + MWC_MAXDRAW; i += blockDim.x) sRaw[i] = raw[pos + i];
        __syncthreads();
        // A: one attempt from every start position of the block
        for (int q = threadIdx.x; q < MWC_BLOCK; q += blockDim.x) {
            int idx[MP];
            int c = 0;
            bool over = false;
#pragma unroll
            for (int i = 0; i < MP; ++i) {
                int v = 0;
                bool dup;
                do {
                    if (c >= MWC_MAXDRAW) { over = true; break; }
                    v = (int)(sRaw[q + c] % (uint32_t)n);
                    ++c;
                    dup = false;
#pragma unroll
                    for (int k = 0; k < MP; ++k) dup |= (k < i) && (idx[k] == v);
                } while (dup);
                idx[i] = v;
            }
            bool ok = false;
            if (!over) ok = check_subset_pts<MODEL>(pts, idx);
#pragma unroll
            for (int i = 0; i < MP; ++i) aIdx[q * MP + i] = (uint16_t)idx[i];
            aLen[q] = over ? 0xFF : (uint8_t)((ok ? 0x80 : 0) | c);
        }
        __syncthreads();
        // B: follow the chain
        if (threadIdx.x == 0) {
            int q = 0, k = 0, st = 0, att = attempts;
            while (q < MWC_BLOCK && it + k < maxIters) {
                const int len = aLen[q];
                if (len == 0xFF) { st = 2; break; }
                if (len & 0x80) { itPos[k++] = (uint16_t)q; att = 0; }
                else if (++att >= 10000) { st = 3; break; }
                q += len & 0x7F;
            }
            sState[0] = k; sState[1] = st; sState[2] = pos + q; sState[3] = att;
        }
        __syncthreads();
        const int k = sState[0];
        stop = sState[1]; pos = sState[2]; attempts = sState[3];
        // C: subsets of the iterations found in this round
        for (int i = threadIdx.x; i < k * MP; i += blockDim.x)
            out[(int64_t)it * MP + i] = aIdx[itPos[i / MP] * MP + (i % MP)];
        it += k;
    }
    if (threadIdx.x == 0) {
        if (stop == 2) { needSeq[p] = 1; }
        else {
            needSeq[p] = 0;
            nIter[p] = it;
            flags[p] = (stop == 3 && it == 0) ? PAIR_FAIL : 0;
        }
    }
}

// Native mode: hypothesis (pair, it) draws from its own counter-based stream -- no sequential
// dependency, same duplicate / degeneracy rules.
// A subset that fails the degeneracy test is redrawn from the same stream (up to 64 attempts), and how often that happens
// differs wildly between hypotheses (homography quadruples at 30 % inliers: a geometric tail; with one hypothesis per
// thread the kernel ran with 9 of 32 lanes active).  So the lanes of a block share a pool of SC_POOL hypotheses: a lane
// makes ONE attempt per trip and, when its hypothesis is settled, takes the next one from the pool.  Every hypothesis
// still sees exactly its own stream and attempts, only who runs it changes.
constexpr int SC_POOL = 1024;
template <int MODEL>
__global__ void __launch_bounds__(128)
sample_counter_kernel(const float4* __restrict__ pf, const int64_t* __restrict__ off, int nPairs,
                      int maxIters, uint64_t seed, int64_t pairBase, int* __restrict__ samples,
                      int* __restrict__ nIter, int* __restrict__ flags) {
    constexpr int MP = MT<MODEL>::MP;
    const int p = blockIdx.y;
    const int it0 = blockIdx.x * SC_POOL, it1 = min(maxIters, it0 + SC_POOL);
    const int64_t base = off[p];
    const int n = (int)(off[p + 1] - base);
    int* outP = samples + (int64_t)p * maxIters * MP;
    if (n <= MP) {
        if (blockIdx.x == 0 && threadIdx.x == 0) {
            if (n < MP) { nIter[p] = 0; flags[p] = PAIR_FAIL; }
            else { for (int i = 0; i < MP; ++i) outP[i] = i; nIter[p] = 1; flags[p] = PAIR_DIRECT; }
        }
        return;
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) { nIter[p] = maxIters; flags[p] = 0; }
    __shared__ int sNext;
    if (threadIdx.x == 0) sNext = it0 + (int)blockDim.x;
    __syncthreads();
    int it = it0 + threadIdx.x, attempt = 0;
    uint64_t ctr = 0;
    if (it < it1) ctr = mix64(seed ^ mix64((uint64_t)(pairBase + p) * 0x100000001B3ull + (uint64_t)it));
    while (it < it1) {
        int idx[MP];
        for (int i = 0; i < MP; ++i) {
            int v;
            bool dup;
            do {
                ctr = mix64(ctr);
                v = (int)((uint32_t)(ctr >> 32) % (uint32_t)n);
                dup = false;
                for (int k = 0; k < i; ++k) dup |= (idx[k] == v);
            } while (dup);
            idx[i] = v;
        }
        ++attempt;
        if (check_subset<MODEL>(pf, base, idx) || attempt == 64) {       // after 64 tries keep the last subset
            int* out = outP + (int64_t)it * MP;
            for (int i = 0; i < MP; ++i) out[i] = idx[i];
            it = atomicAdd(&sNext, 1);
            attempt = 0;
            if (it < it1) ctr = mix64(seed ^ mix64((uint64_t)(pairBase + p) * 0x100000001B3ull + (uint64_t)it));
        }
    }
}

__global__ void __launch_bounds__(256)
injected_setup_kernel(const int64_t* __restrict__ off, int nPairs, int maxIters, int mp,
                      int* __restrict__ nIter, int* __restrict__ flags) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= nPairs) return;
    const int n = (int)(off[p + 1] - off[p]);
    if (n < mp) { nIter[p] = 0; flags[p] = PAIR_FAIL; }
    else { nIter[p] = maxIters; flags[p] = 0; }
}

// ---------------------------------------------------------------------------------------------
// solve: one hypothesis per thread
// ---------------------------------------------------------------------------------------------
template <int MODEL>
__global__ void __launch_bounds__(64)
solve_kernel(const float4* __restrict__ pf, const double* __restrict__ pd,
             const int64_t* __restrict__ off, int maxIters, const int* __restrict__ samples,
             const int* __restrict__ nIter, double* __restrict__ models,
             uint8_t* __restrict__ nModels) {
    constexpr int MP = MT<MODEL>::MP, MAXM = MT<MODEL>::MAXM;
    const int it = blockIdx.x * blockDim.x + threadIdx.x;
    const int p = blockIdx.y;
    if (it >= maxIters) return;
    uint8_t* nm = nModels + (int64_t)p * maxIters + it;
    if (it >= nIter[p]) { *nm = 0; return; }
    const int64_t base = off[p];
    const int n = (int)(off[p + 1] - base);
    const int* idx = samples + ((int64_t)p * maxIters + it) * MP;
    double* out = models + ((int64_t)p * maxIters + it) * MAXM * 9;
    int cnt = 0;
    bool bad = false;
    for (int i = 0; i < MP; ++i) bad |= (idx[i] < 0 || idx[i] >= n);
    if (bad) { *nm = 0; return; }
    // (the 5-point solver does not run here: as one kernel it needed 255 registers and ran at 18.7 of 32 active lanes,
    // and the single-function composition e_solve5, compiled for the device, missed cv2 on 40 of 184 golden roots by > 1e-6
    // while the same stages as separate kernels match -- see solve_e_*_kernel below)
    static_assert(MODEL != RM_ESSENTIAL, "essential matrices are solved by the three solve_e_* kernels");
    {
        float s[2 * MP], d[2 * MP];
        for (int i = 0; i < MP; ++i) {
            const float4 q = pf[base + idx[i]];
            s[2 * i] = q.x; s[2 * i + 1] = q.y; d[2 * i] = q.z; d[2 * i + 1] = q.w;
        }
        if (MODEL == RM_HOMOGRAPHY) {
            double H[9];
            cnt = h_solve4(s, d, H) ? 1 : 0;             // closed form, registers only (cvmath.cuh)
            if (cnt) for (int k = 0; k < 9; ++k) out[k] = H[k];
        } else {
            // the 7 x 9 design matrix of the Jacobi SVD lives in shared memory, interleaved by thread (element e of
            // thread t at [e * 64 + t]: conflict-free); as a local array it thrashed L1 and was written back to DRAM
            __shared__ double sAt[(MODEL == RM_FUNDAMENTAL) ? 63 * 64 : 1];
            double F[27];
            cnt = f_solve7_at<64>(s, d, F, sAt + threadIdx.x);
            if (cnt < 0 || cnt > 3) cnt = 0;
            for (int k = 0; k < cnt * 9; ++k) out[k] = F[k];
        }
    }
    *nm = (uint8_t)cnt;
}

// 5-point solver as three kernels (cvmath.cuh: e5_stage1 / real_roots10 / e5_root_to_E).  As one kernel it needed 255
// registers (8 warps per SM) and ran with 18.7 of 32 lanes active: the number of real roots and the Laguerre iteration
// counts differ per hypothesis, and every root carries a 3x3 Jacobi SVD.  Split: (1) null space, elimination and
// det B(z) per hypothesis; (2) root finding per hypothesis; (3) one thread per (hypothesis, root), the valid roots
// compacted in root order inside a 16-lane group, so the models land exactly where the single kernel put them.
constexpr int E5_SCR = 96;          // doubles per hypothesis: EE 36 | b 39 | coef 11 | roots 10

// Stage 1a -- null space (one thread per hypothesis): the 5 x 9 epipolar system's rows are orthonormalised and OpenCV's
// basis completion run on them (cvmath.cuh nullspace_rows: the vt rows 5..8 of cv::SVD, without the Jacobi sweeps); the
// 9 x 9 working matrix sits in shared memory, interleaved by thread (element e of thread t at [e * 64 + t]), the current
// row in registers.  Writes EE (4 x 9) to the hypothesis' scratch record.
// Hypothesis source: RANSAC samples (pd / off / samples / nIter, grid.y = pair) or explicit 5-point sets (dirA / dirB).
__global__ void __launch_bounds__(64)
solve_e_null_kernel(const double* __restrict__ pd, const int64_t* __restrict__ off, int maxIters,
                    const int* __restrict__ samples, const int* __restrict__ nIter, const double* __restrict__ dirA,
                    const double* __restrict__ dirB, int64_t nDirect, double* __restrict__ scr, int* __restrict__ nRoots) {
    constexpr int MP = 5;
    __shared__ double sAt[45 * 64];
    const int it = blockIdx.x * blockDim.x + threadIdx.x;
    int64_t h;
    double a[2 * MP], b[2 * MP];
    if (dirA) {
        h = it;
        if (h >= nDirect) return;
        nRoots[h] = -1;
        for (int k = 0; k < 10; ++k) { a[k] = dirA[10 * h + k]; b[k] = dirB[10 * h + k]; }
    } else {
        const int p = blockIdx.y;
        if (it >= maxIters) return;
        h = (int64_t)p * maxIters + it;
        nRoots[h] = -1;                                      // -1: no system (skipped iteration / bad sample / singular)
        if (it >= nIter[p]) return;
        const int64_t base = off[p];
        const int n = (int)(off[p + 1] - base);
        const int* idx = samples + h * MP;
        bool bad = false;
        for (int i = 0; i < MP; ++i) bad |= (idx[i] < 0 || idx[i] >= n);
        if (bad) return;
        for (int i = 0; i < MP; ++i) {
            const double* q = pd + 4 * (base + idx[i]);
            a[2 * i] = q[0]; a[2 * i + 1] = q[1]; b[2 * i] = q[2]; b[2 * i + 1] = q[3];
        }
    }
    double* At = sAt + threadIdx.x;
    for (int i = 0; i < 5; ++i) {
        const double x1 = a[2 * i], y1 = a[2 * i + 1], x2 = b[2 * i], y2 = b[2 * i + 1];
        double* q = At + 9 * i * 64;
        q[0] = x1 * x2; q[1 * 64] = y1 * x2; q[2 * 64] = x2; q[3 * 64] = x1 * y2; q[4 * 64] = y1 * y2; q[5 * 64] = y2;
        q[6 * 64] = x1; q[7 * 64] = y1; q[8 * 64] = 1.0;
    }
    double EE[36];
    nullspace_rows<9, 5, 9, 64>(At, EE);
    double* o = scr + h * E5_SCR;
    for (int k = 0; k < 36; ++k) o[k] = EE[k];
    nRoots[h] = 0;
}

// Stage 1b -- constraints, elimination, det B(z): ten lanes per hypothesis, three hypotheses per warp.
//   * lane r of a group builds row r of the 10 x 20 constraint matrix in registers (row 0 = det E, rows 1..9 =
//     2 E E^T E - tr(E E^T) E), from the group's EE and the six E E^T polynomials in shared memory;
//   * Gauss-Jordan with partial pivoting runs ACROSS the lanes: the pivot is an arg-max over the group (four rotation
//     shuffles of a key that carries the lane), the pivot row is broadcast by shuffles and every other lane eliminates
//     its own row in registers -- no local memory, no shared-memory matrix;
//   * the lanes that ended up holding rows 4..9 hand them to lanes 0..2, which form the three rows of B(z); lanes 0..5
//     expand one permutation term of det B(z) each and lane 0 adds them in the fixed order of the serial code.
// (As one thread per hypothesis this stage held a 3.3 KB local frame, wrote 10x its useful bytes back to DRAM and
// issued on 9 % of the cycles, see profiles/r01_filtered_scoring_ncu.md.)
constexpr int E5_HPB = 12;        // hypotheses per 128-thread block: 4 warps x 3 groups of 10 lanes

__device__ __forceinline__ double shfl_d(double v, int src) { return __shfl_sync(0xFFFFFFFFu, v, src); }

__global__ void __launch_bounds__(128)
solve_e_elim_kernel(int64_t nHyp, double* __restrict__ scr, int* __restrict__ nRoots) {
    __shared__ double sEE[E5_HPB][36];
    __shared__ double sEet[E5_HPB][6][10];
    __shared__ double sB[E5_HPB][39];
    __shared__ double sTerm[E5_HPB][6][11];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = lane / 10, pos = lane - 10 * g;            // lanes 30, 31: g == 3, idle
    const int slot = warp * 3 + (g < 3 ? g : 0);
    const int64_t h = ((int64_t)blockIdx.x * 4 + warp) * 3 + g;
    const bool live = g < 3 && h < nHyp && nRoots[h] == 0;
    const int gbase = 10 * (g < 3 ? g : 0);
    double (&EE)[36] = sEE[slot];
    if (live) for (int k = pos; k < 36; k += 10) EE[k] = scr[h * E5_SCR + k];
    __syncwarp();
    // E[i][j] as a degree-1 polynomial in (x, y, z, 1): coefficient of basis b at EE[b * 9 + (i * 3 + j)]
    auto e1 = [&](int k, double* o) { o[0] = EE[k]; o[1] = EE[9 + k]; o[2] = EE[18 + k]; o[3] = EE[27 + k]; };
    if (live && pos < 6) {
        const int i = pos < 3 ? 0 : (pos < 5 ? 1 : 2), j = pos < 3 ? pos : (pos < 5 ? pos - 2 : 2);
        double acc[10];
#pragma unroll
        for (int k = 0; k < 10; ++k) acc[k] = 0;
        for (int m = 0; m < 3; ++m) {
            double u[4], v[4], t[10];
            e1(i * 3 + m, u); e1(j * 3 + m, v);
            p5::mul11(u, v, t);
            // same association as the serial code: (t0 + t1) + t2
#pragma unroll
            for (int k = 0; k < 10; ++k) acc[k] = m == 0 ? t[k] : acc[k] + t[k];
        }
#pragma unroll
        for (int k = 0; k < 10; ++k) sEet[slot][pos][k] = acc[k];
    }
    __syncwarp();
    double row[20];
#pragma unroll
    for (int k = 0; k < 20; ++k) row[k] = 0;
    if (live) {
        if (pos == 0) {
            double m[10], u[4], v[4], w[4];
            e1(4, u); e1(8, v); p5::mul11(u, v, m); e1(0, w); p5::mul21_acc(m, w, 1., row);
            e1(5, u); e1(7, v); p5::mul11(u, v, m); p5::mul21_acc(m, w, -1., row);
            e1(3, u); e1(8, v); p5::mul11(u, v, m); e1(1, w); p5::mul21_acc(m, w, -1., row);
            e1(5, u); e1(6, v); p5::mul11(u, v, m); p5::mul21_acc(m, w, 1., row);
            e1(3, u); e1(7, v); p5::mul11(u, v, m); e1(2, w); p5::mul21_acc(m, w, 1., row);
            e1(4, u); e1(6, v); p5::mul11(u, v, m); p5::mul21_acc(m, w, -1., row);
        } else {
            const int i = (pos - 1) / 3, j = (pos - 1) - 3 * i;
            auto pairIdx = [](int a, int b) { const int lo = a < b ? a : b, hi = a < b ? b : a; return lo == 0 ? hi : (lo == 1 ? 2 + hi : 5); };
            for (int k = 0; k < 3; ++k) {
                double pe[10], u[4];
                const double* src = sEet[slot][pairIdx(i, k)];
#pragma unroll
                for (int q = 0; q < 10; ++q) pe[q] = src[q];
                e1(k * 3 + j, u);
                p5::mul21_acc(pe, u, 2., row);
            }
            double tr[10], u[4];
#pragma unroll
            for (int q = 0; q < 10; ++q) tr[q] = sEet[slot][0][q] + sEet[slot][3][q] + sEet[slot][5][q];
            e1(i * 3 + j, u);
            p5::mul21_acc(tr, u, -1., row);
        }
    }
    // ---- Gauss-Jordan on the ten leading columns, rows spread over the group's lanes
    bool used = false, singular = false;
    int myCol = -1;
#pragma unroll
    for (int c = 0; c < 10; ++c) {
        // arg-max of |row[c]| over the lanes that are not pivots yet; ties and near-ties (below 2^-48) go to the lower lane
        unsigned long long key = 0;
        if (live && !used) key = ((unsigned long long)__double_as_longlong(fabs(row[c])) & ~15ull) | (unsigned long long)(15 - pos);
#pragma unroll
        for (int d = 1; d <= 4; d <<= 1) {
            const int src = gbase + (pos + d >= 10 ? pos + d - 10 : pos + d);
            const unsigned long long o = __shfl_sync(0xFFFFFFFFu, key, src);
            key = o > key ? o : key;
        }
        {   // windows of 8 consecutive lanes so far: one more rotation by 2 covers all ten
            const int src = gbase + (pos + 2 >= 10 ? pos - 8 : pos + 2);
            const unsigned long long o = __shfl_sync(0xFFFFFFFFu, key, src);
            key = o > key ? o : key;
        }
        const int pl = gbase + (15 - (int)(key & 15ull));
        if ((key >> 4) == 0) singular = true;                // no usable pivot (the serial code returns false)
        const bool isPivot = live && lane == pl && !used;
        const double pv = shfl_d(row[c], pl);
        const double inv = 1 / pv;
        const double f = row[c];
#pragma unroll
        for (int k = c + 1; k < 20; ++k) {
            const double t = shfl_d(row[k], pl) * inv;       // the normalised pivot row
            row[k] = isPivot ? t : row[k] - f * t;
        }
        row[c] = isPivot ? 1.0 : 0.0;
        if (isPivot) { used = true; myCol = c; }
    }
    // ---- rows 4..9 (by pivot column) -> B(z): lane i in 0..2 takes rows 2 i + 4 and 2 i + 5
    int src1 = gbase, src2 = gbase;
#pragma unroll
    for (int c = 4; c < 10; ++c) {
        const unsigned m = __ballot_sync(0xFFFFFFFFu, myCol == c);
        const unsigned mg = (m >> gbase) & 0x3FFu;
        const int who = gbase + (mg ? __ffs(mg) - 1 : 0);
        if (c == 2 * pos + 4) src1 = who;
        if (c == 2 * pos + 5) src2 = who;
    }
    double r1[10], r2[10];
#pragma unroll
    for (int k = 0; k < 10; ++k) { r1[k] = shfl_d(row[10 + k], src1); r2[k] = shfl_d(row[10 + k], src2); }
    if (live && pos < 3) {
        double row1[13], row2[13];
#pragma unroll
        for (int k = 0; k < 13; ++k) { row1[k] = 0; row2[k] = 0; }
        row1[1] = r1[0]; row1[2] = r1[1]; row1[3] = r1[2];
        row1[5] = r1[3]; row1[6] = r1[4]; row1[7] = r1[5];
        row1[9] = r1[6]; row1[10] = r1[7]; row1[11] = r1[8]; row1[12] = r1[9];
        row2[0] = r2[0]; row2[1] = r2[1]; row2[2] = r2[2];
        row2[4] = r2[3]; row2[5] = r2[4]; row2[6] = r2[5];
        row2[8] = r2[6]; row2[9] = r2[7]; row2[10] = r2[8]; row2[11] = r2[9];
#pragma unroll
        for (int k = 0; k < 13; ++k) sB[slot][13 * pos + k] = row1[k] - row2[k];
    }
    __syncwarp();
    if (live && pos < 6) {
        // permutation terms of det B(z): columns 0 -> x-part (deg 3), 1 -> y-part (deg 3), 2 -> 1-part (deg 4)
        const int p0 = pos == 0 ? 0 : pos == 1 ? 1 : pos == 2 ? 2 : pos == 3 ? 0 : pos == 4 ? 1 : 2;
        const int p1 = pos == 0 ? 1 : pos == 1 ? 2 : pos == 2 ? 0 : pos == 3 ? 2 : pos == 4 ? 0 : 1;
        const int p2 = pos == 0 ? 2 : pos == 1 ? 0 : pos == 2 ? 1 : pos == 3 ? 1 : pos == 4 ? 2 : 0;
        double xy[7], full[11];
        upoly_mul(sB[slot] + 13 * p0 + 0, 4, sB[slot] + 13 * p1 + 4, 4, xy);
        upoly_mul(xy, 7, sB[slot] + 13 * p2 + 8, 5, full);
#pragma unroll
        for (int k = 0; k < 11; ++k) sTerm[slot][pos][k] = full[k];
    }
    __syncwarp();
    // a group is singular when any of its lanes saw an empty pivot column
    const unsigned sm = __ballot_sync(0xFFFFFFFFu, singular);
    const bool groupSing = ((sm >> gbase) & 0x3FFu) != 0;
    if (live) {
        double* o = scr + h * E5_SCR;
        if (groupSing) { if (pos == 0) nRoots[h] = -1; }
        else {
            for (int k = pos; k < 39; k += 10) o[36 + k] = sB[slot][k];
            if (pos == 0) {
                for (int k = 0; k < 11; ++k) {
                    double cf = 0;
                    for (int t = 0; t < 6; ++t) cf += (t < 3 ? 1. : -1.) * sTerm[slot][t][k];
                    o[75 + k] = cf;
                }
            }
        }
    }
}

// Stage 2 -- real roots of det B(z) (degree 10), one thread per hypothesis.  Same arithmetic as cvmath.cuh real_roots10
// (Laguerre iterations with deflation by real linear / quadratic factors, every root polished by Newton steps on the
// original polynomial, |Im| <= 1e-10 keeps a root), but run as a FLAT state machine: one trip of the warp-uniform loop
// is one polynomial evaluation followed by one Laguerre or one Newton update of whatever root the lane is working on.
// A lane that finishes a root starts its next one in the following trip instead of idling until the slowest lane of the
// warp has finished the same root -- the nested-loop form ran with 15 of 32 lanes active.  Coefficients, working
// polynomial and roots live in registers (all loops are unrolled with predicates on the current degree).
__global__ void __launch_bounds__(128)
solve_e_roots_kernel(int64_t nHyp, double* __restrict__ scr, int* __restrict__ nRoots) {
    const int64_t h = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    bool active = h < nHyp && nRoots[h] >= 0;
    double a0[11], w[11], out[10];
    int deg = 10, n = 0;
#pragma unroll
    for (int k = 0; k < 11; ++k) a0[k] = 0;
    if (active) {
        const double* c = scr + h * E5_SCR + 75;
#pragma unroll
        for (int k = 0; k < 11; ++k) a0[k] = c[k];
        // strip leading zero coefficients (degenerate systems)
        for (int guard = 0; guard < 10 && deg > 0 && a0[0] == 0; ++guard) {
#pragma unroll
            for (int k = 0; k < 10; ++k) a0[k] = a0[k + 1];
            a0[10] = 0;
            --deg;
        }
        if (deg == 0) active = false;
    }
#pragma unroll
    for (int k = 0; k < 11; ++k) w[k] = a0[k];
#pragma unroll
    for (int k = 0; k < 10; ++k) out[k] = 0;
    const int deg0 = deg;
    // per-root state
    bool lag = true;                 // true: Laguerre on the working polynomial, false: Newton polish on the original
    double xr = 0, xi = 0;           // Laguerre estimate (also what the working polynomial is deflated by)
    double pr = 0, pi = 0;           // polished estimate
    int it = 0;
    if (active && deg == 1) { xr = -w[1] / w[0]; xi = 0; pr = xr; pi = 0; lag = false; it = 0; }
    while (__any_sync(0xFFFFFFFFu, active)) {
        if (!active) continue;
        const double er = lag ? xr : pr, ei = lag ? xi : pi;
        const int dsel = lag ? deg : deg0;
        double br = lag ? w[0] : a0[0], bi = 0, dr = 0, di = 0, fr = 0, fi = 0;
        double err = fabs(br);
        const double ax = sqrt(er * er + ei * ei);
#pragma unroll
        for (int j = 1; j <= 10; ++j) {
            if (j <= dsel) {
                const double cj = lag ? w[j] : a0[j];
                const double nfr = er * fr - ei * fi + dr, nfi = er * fi + ei * fr + di;
                const double ndr = er * dr - ei * di + br, ndi = er * di + ei * dr + bi;
                const double nbr = er * br - ei * bi + cj, nbi = er * bi + ei * br;
                fr = nfr; fi = nfi; dr = ndr; di = ndi; br = nbr; bi = nbi;
                err = (fabs(br) + fabs(bi)) + ax * err;           // |b| <= |Re| + |Im|: a bound, no square root per term
            }
        }
        bool rootDone = false;
        if (lag) {
            bool conv = false;
            const double b2 = br * br + bi * bi;
            const double floor_ = err * 2e-16;
            if (b2 <= floor_ * floor_) conv = true;               // at the rounding floor of p(x)
            else {
                const double ib2 = 1 / b2;
                const double gr = (dr * br + di * bi) * ib2, gi = (di * br - dr * bi) * ib2;
                const double g2r = gr * gr - gi * gi, g2i = 2 * gr * gi;
                const double qr = (fr * br + fi * bi) * ib2, qi = (fi * br - fr * bi) * ib2;
                const double hr = g2r - 2 * qr, hi = g2i - 2 * qi;
                const double tr = (deg - 1) * (deg * hr - g2r), ti = (deg - 1) * (deg * hi - g2i);
                const double tm = sqrt(tr * tr + ti * ti);
                double sr = sqrt(fmax(0.0, 0.5 * (tm + tr))), si = sqrt(fmax(0.0, 0.5 * (tm - tr)));
                if (ti < 0) si = -si;
                double p1r = gr + sr, p1i = gi + si;
                const double p2r = gr - sr, p2i = gi - si;
                const double m1 = p1r * p1r + p1i * p1i, m2 = p2r * p2r + p2i * p2i;
                if (m2 > m1) { p1r = p2r; p1i = p2i; }
                const double pm = fmax(m1, m2);
                double dxr, dxi;
                if (pm > 0) { const double ipm = deg / pm; dxr = p1r * ipm; dxi = -p1i * ipm; }
                else { dxr = (1 + ax) * cos((double)it); dxi = (1 + ax) * sin((double)it); }
                const double nxr = xr - dxr, nxi = xi - dxi;
                if (nxr == xr && nxi == xi) conv = true;
                else {
                    xr = nxr; xi = nxi;
                    if (fabs(dxr) + fabs(dxi) <= 1e-16 * (fabs(xr) + fabs(xi))) conv = true;
                }
            }
            ++it;
            if (conv || it >= 80) {
                pr = xr; pi = xi; it = 0;
                // a clearly complex estimate cannot polish into a real root (|Im| <= 1e-10): skip its Newton steps on
                // the full-degree polynomial -- it is only deflated away (same rule in cvmath.cuh real_roots10)
                if (fabs(xi) > 1e-4 * (1 + fabs(xr))) rootDone = true;
                else lag = false;
            }
        } else {
            // one Newton step on the original polynomial (at most four), kept only while it is a small correction
            bool stop = false;
            const double d2 = dr * dr + di * di;
            if (d2 == 0) stop = true;
            else {
                const double id2 = 1 / d2;
                const double sr = (br * dr + bi * di) * id2, si = (bi * dr - br * di) * id2;
                if (fabs(sr) + fabs(si) > 1e-3 * (fabs(pr) + fabs(pi)) + 1e-300) stop = true;
                else {
                    pr -= sr; pi -= si;
                    if (fabs(sr) + fabs(si) <= 1e-16 * (fabs(pr) + fabs(pi))) stop = true;
                }
            }
            ++it;
            if (stop || it >= 4) rootDone = true;
        }
        if (rootDone) {
            if (fabs(pi) <= 1e-10) {
#pragma unroll
                for (int k = 0; k < 10; ++k) if (k == n) out[k] = pr;
                ++n;
            }
            // deflate the working polynomial by a real factor
            if (fabs(xi) <= 1e-12 * (1 + fabs(xr)) || deg == 1) {
                double rem = w[0];
#pragma unroll
                for (int j = 1; j < 10; ++j)
                    if (j < deg) { const double t = w[j] + rem * xr; w[j] = t; rem = t; }
                deg -= 1;
            } else {
                const double pcoef = 2 * xr, qcoef = -(xr * xr + xi * xi);
                double b0 = w[0], b1 = w[1] + pcoef * b0;
                w[1] = b1;
#pragma unroll
                for (int j = 2; j <= 8; ++j)
                    if (j <= deg - 2) { const double t = w[j] + pcoef * b1 + qcoef * b0; b0 = b1; b1 = t; w[j] = t; }
                deg -= 2;
            }
            if (deg < 1) active = false;
            else if (deg == 1) { xr = -w[1] / w[0]; xi = 0; pr = xr; pi = 0; lag = false; it = 0; }
            else { xr = 0; xi = 0; lag = true; it = 0; }
        }
    }
    if (!(h < nHyp) || nRoots[h] < 0) return;
    // ascending order (deterministic model order), then drop estimates that polished onto the same root
    const double inf = __longlong_as_double(0x7FF0000000000000ll);
#pragma unroll
    for (int k = 0; k < 10; ++k) if (k >= n) out[k] = inf;
#pragma unroll
    for (int i = 1; i < 10; ++i) {
#pragma unroll
        for (int j = i; j > 0; --j) {
            const double lo = fmin(out[j - 1], out[j]), hi2 = fmax(out[j - 1], out[j]);
            out[j - 1] = lo; out[j] = hi2;
        }
    }
    double* o = scr + h * E5_SCR + 86;
    int m = 0;
    double last = 0;
#pragma unroll
    for (int i = 0; i < 10; ++i) {
        if (i < n && (m == 0 || fabs(out[i] - last) > 1e-13 * (fabs(out[i]) + fabs(last)))) { o[m] = out[i]; last = out[i]; ++m; }
    }
    nRoots[h] = m;
}

__global__ void __launch_bounds__(128)
solve_e_models_kernel(int64_t nHyp, const double* __restrict__ scr, const int* __restrict__ nRoots,
                      double* __restrict__ models, uint8_t* __restrict__ nModels) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t h = t >> 4;                                // 16 lanes per hypothesis, lane r = root r
    const int r = (int)(t & 15);
    const bool live = h < nHyp;
    const int nr = live ? nRoots[h] : -1;
    double E[9];
    bool valid = false;
    if (r < nr) {
        const double* s = scr + h * E5_SCR;
        double EE[36], bb[39];
        for (int k = 0; k < 36; ++k) EE[k] = s[k];
        for (int k = 0; k < 39; ++k) bb[k] = s[36 + k];
        valid = e5_root_to_E(EE, bb, s[86 + r], E);
    }
    const unsigned all = __ballot_sync(0xFFFFFFFFu, valid);
    const unsigned grp = (all >> (threadIdx.x & 16)) & 0xFFFFu;
    if (valid) {
        const int slot = __popc(grp & ((1u << r) - 1u));
        double* o = models + (h * 10 + slot) * 9;
        for (int k = 0; k < 9; ++k) o[k] = E[k];
    }
    if (live && r == 0) nModels[h] = (uint8_t)__popc(grp);
}

#ifdef MCV_DEBUG_COMPOSED_E5
// Debug build only: the whole 5-point solver as ONE device function per thread (cvmath.cuh e_solve5), the form round 1
// saw disagree with the host build of the same source; kept buildable for the sanitizer runs recorded in profiles/.
__global__ void __launch_bounds__(64)
five_point_composed_kernel(const double* __restrict__ a, const double* __restrict__ b, int nSets,
                           double* __restrict__ Es, int* __restrict__ counts) {
    const int h = blockIdx.x * blockDim.x + threadIdx.x;
    if (h >= nSets) return;
    double pa[10], pb[10], E[90];
    for (int k = 0; k < 10; ++k) { pa[k] = a[10 * (int64_t)h + k]; pb[k] = b[10 * (int64_t)h + k]; }
    const int n = e_solve5(pa, pb, E);
    for (int k = 0; k < 9 * n; ++k) Es[90 * (int64_t)h + k] = E[k];
    counts[h] = n;
#ifdef MCV_DEBUG_E5_TRACE
    if (h == 0) {
        double EE[36], bb[39], coef[11], roots[10];
        e5_stage1(pa, pb, EE, bb, coef);
        const int nr = real_roots10(coef, roots);
        printf("composed n=%d staged-in-thread nr=%d\n", n, nr);
        for (int r = 0; r < nr; ++r) {
            double Er[9];
            const bool okr = e5_root_to_E(EE, bb, roots[r], Er);
            printf("  root %d z=%.17g ok=%d E0=%.15g E8=%.15g | composed E0=%.15g E8=%.15g\n", r, roots[r], (int)okr, Er[0], Er[8],
                   r < n ? E[9 * r] : 0., r < n ? E[9 * r + 8] : 0.);
        }
    }
#endif
}
#endif

__global__ void __launch_bounds__(256)
widen_counts_kernel(const uint8_t* __restrict__ in, int n, int* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = in[i];
}

// injected models: one model per iteration, copied into slot 0
__global__ void __launch_bounds__(256)
inject_models_kernel(const double* __restrict__ inj, int64_t nHyp, int maxm,
                     double* __restrict__ models, uint8_t* __restrict__ nModels) {
    const int64_t h = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (h >= nHyp) return;
    for (int k = 0; k < 9; ++k) models[h * maxm * 9 + k] = inj[h * 9 + k];
    nModels[h] = 1;
}

// ---------------------------------------------------------------------------------------------
// compact: dense model list per pair in (iteration, model) order
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
compact_kernel(const uint8_t* __restrict__ nModels, const int* __restrict__ nIter, int maxIters,
               int maxm, int* __restrict__ iterStart, int* __restrict__ slotList,
               int* __restrict__ nTotal) {
    const int p = blockIdx.x;
    const uint8_t* nm = nModels + (int64_t)p * maxIters;
    int* is = iterStart + (int64_t)p * (maxIters + 1);
    int* sl = slotList + (int64_t)p * maxIters * maxm;
    __shared__ int sPart[256];
    __shared__ int sBase;
    const int tid = threadIdx.x;
    const int ni = nIter[p];
    if (tid == 0) sBase = 0;
    __syncthreads();
    for (int c0 = 0; c0 < maxIters; c0 += 256) {
        const int it = c0 + tid;
        const int v = (it < ni) ? nm[it] : 0;
        sPart[tid] = v;
        __syncthreads();
        for (int o = 1; o < 256; o <<= 1) {          // inclusive Hillis-Steele scan
            const int t = tid >= o ? sPart[tid - o] : 0;
            __syncthreads();
            sPart[tid] += t;
            __syncthreads();
        }
        const int start = sBase + sPart[tid] - v;
        if (it < maxIters) {
            is[it] = start;
            for (int m = 0; m < v; ++m) sl[start + m] = it * maxm + m;
        }
        __syncthreads();
        if (tid == 255) sBase += sPart[255];
        __syncthreads();
    }
    if (tid == 0) { is[maxIters] = sBase; nTotal[p] = sBase; }
}

// Filter tile layout: correspondences in pairs, 8 floats per pair = x1 x1' | y1 y1' | x2 x2' | y2 y2', so that two
// 128-bit shared-memory loads deliver the four packed (f32x2) operands of two correspondences.
__device__ __forceinline__ void store_pair_slot(float* spair, int j, const float4& q) {
    float* d = spair + 8 * (j >> 1) + (j & 1);
    d[0] = q.x; d[2] = q.y; d[4] = q.z; d[6] = q.w;
}
// 64 correspondences (jn of them valid) through the packed filter; bit j of the result = "needs the exact test"
template <int MODEL>
__device__ __forceinline__ unsigned long long filter_block64(const FFilter2& p, const float* blk, int jn) {
    const ulonglong2* b2 = reinterpret_cast<const ulonglong2*>(blk);
    if (jn == 64) {
        // 8 pairs (16 correspondences) per fully unrolled step: mask bits are immediates, and the live state stays
        // small enough for 5 blocks per SM
        unsigned long long pend = 0;
#pragma unroll 1
        for (int o = 0; o < 4; ++o) {
            unsigned w = 0;
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                const ulonglong2 u = b2[16 * o + 2 * q], v = b2[16 * o + 2 * q + 1];
                bool r0, r1;
                if (MODEL == RM_FUNDAMENTAL) f_certain_outlier2(p, F2{u.x}, F2{u.y}, F2{v.x}, F2{v.y}, r0, r1);
                else e_certain_outlier2(p, F2{u.x}, F2{u.y}, F2{v.x}, F2{v.y}, r0, r1);
                if (!r0) w |= 1u << (2 * q);
                if (!r1) w |= 2u << (2 * q);
            }
            pend |= (unsigned long long)w << (16 * o);
        }
        return pend;
    }
    unsigned long long pend = 0;
    for (int q = 0; 2 * q < jn; ++q) {
        const ulonglong2 u = b2[2 * q], v = b2[2 * q + 1];
        bool r0, r1;
        if (MODEL == RM_FUNDAMENTAL) f_certain_outlier2(p, F2{u.x}, F2{u.y}, F2{v.x}, F2{v.y}, r0, r1);
        else e_certain_outlier2(p, F2{u.x}, F2{u.y}, F2{v.x}, F2{v.y}, r0, r1);
        if (!r0) pend |= 1ull << (2 * q);
        if (!r1 && 2 * q + 1 < jn) pend |= 2ull << (2 * q);
    }
    return pend;
}

// ---------------------------------------------------------------------------------------------
// score: one model per thread, correspondences through shared memory
// ---------------------------------------------------------------------------------------------
template <int MODEL>
__global__ void __launch_bounds__(128)
score_kernel(const float4* __restrict__ pf, const double* __restrict__ pd,
             const int64_t* __restrict__ off, const double* __restrict__ models,
             const int* __restrict__ slotList, const int* __restrict__ nTotal, int maxSlots,
             float thr2, int chunkPts, int* __restrict__ counts, const int* __restrict__ gate,
             const int* __restrict__ nBegin) {
    constexpr int TP = (MODEL == RM_ESSENTIAL) ? 512 : 1024;   // 16 KB per stage
    if (gate && !score_gate_classic(gate)) return;             // the tensor-core kernel scores this chunk
    __shared__ __align__(16) unsigned char sraw[16384];
    __shared__ int smax[4];
    __shared__ __align__(16) float spair_e[(MODEL == RM_ESSENTIAL) ? 4 * 512 : 4];    // essential: fp32 rounding of the tile
    float* const spair = spair_e;
    const int p = blockIdx.y;
    const int total = nTotal[p];
    const int k = blockIdx.x * 128 + threadIdx.x;
    const int begin = nBegin ? nBegin[p] : 0;                  // models [begin, total) of the compacted list
    if (begin >= total || blockIdx.x * 128 >= total || blockIdx.x * 128 + 128 <= begin) return;
    const bool active = k < total && k >= begin;
    const int64_t base = off[p];
    const int n = (int)(off[p + 1] - base);
    const int i0 = blockIdx.z * chunkPts;
    const int i1 = min(n, i0 + chunkPts);

    double m[9];
    float Hf[8];
    if (active) {
        const int slot = slotList[(int64_t)p * maxSlots + k];
        const double* src = models + ((int64_t)p * maxSlots + slot) * 9;
#pragma unroll
        for (int j = 0; j < 9; ++j) m[j] = src[j];
    } else {
#pragma unroll
        for (int j = 0; j < 9; ++j) m[j] = 0;
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) Hf[j] = (float)m[j];
    FFilter ffil;
    const float thrM = thr2 * FILTER_MARGIN;
    FFilter2 ffil2;
    if (MODEL != RM_HOMOGRAPHY) { f_filter_init(m, ffil); filter2_init(ffil, ffil2, thrM); }

    int cnt = 0;
    bool skipFilter = false;
    for (int s0 = i0; s0 < i1; s0 += TP) {
        const int rows = min(TP, i1 - s0);
        if (MODEL == RM_ESSENTIAL) {
            // fp64 tile for the exact test, its fp32 rounding for the filter, and the tile's coordinate maxima
            if (threadIdx.x < 4) smax[threadIdx.x] = 0;
            __syncthreads();
            double4* sp = reinterpret_cast<double4*>(sraw);
            const double4* g = reinterpret_cast<const double4*>(pd) + base + s0;
            int m0 = 0, m1 = 0, m2 = 0, m3 = 0;
            for (int j = threadIdx.x; j < rows; j += 128) {
                const double4 q = g[j];
                sp[j] = q;
                const float4 qf = make_float4((float)q.x, (float)q.y, (float)q.z, (float)q.w);
                store_pair_slot(spair, j, qf);
                if (j == rows - 1 && !(j & 1)) store_pair_slot(spair, j + 1, qf);     // odd tail: duplicate, masked off later
                m0 = max(m0, __float_as_int(fabsf(qf.x))); m1 = max(m1, __float_as_int(fabsf(qf.y)));
                m2 = max(m2, __float_as_int(fabsf(qf.z))); m3 = max(m3, __float_as_int(fabsf(qf.w)));
            }
            m0 = __reduce_max_sync(0xFFFFFFFFu, m0); m1 = __reduce_max_sync(0xFFFFFFFFu, m1);
            m2 = __reduce_max_sync(0xFFFFFFFFu, m2); m3 = __reduce_max_sync(0xFFFFFFFFu, m3);
            if ((threadIdx.x & 31) == 0) {
                atomicMax(&smax[0], m0); atomicMax(&smax[1], m1); atomicMax(&smax[2], m2); atomicMax(&smax[3], m3);
            }
        } else if (MODEL == RM_FUNDAMENTAL) {
            // also the largest |x1|, |y1|, |x2|, |y2| of the tile (bit patterns of non-negative floats order like
            // integers, and a NaN sorts above everything, which switches the filter off for the tile)
            if (threadIdx.x < 4) smax[threadIdx.x] = 0;
            __syncthreads();
            float* spair = reinterpret_cast<float*>(sraw);
            const float4* g = pf + base + s0;
            int m0 = 0, m1 = 0, m2 = 0, m3 = 0;
            for (int j = threadIdx.x; j < rows; j += 128) {
                const float4 q = g[j];
                store_pair_slot(spair, j, q);
                if (j == rows - 1 && !(j & 1)) store_pair_slot(spair, j + 1, q);       // odd tail: duplicate, masked off later
                m0 = max(m0, __float_as_int(fabsf(q.x))); m1 = max(m1, __float_as_int(fabsf(q.y)));
                m2 = max(m2, __float_as_int(fabsf(q.z))); m3 = max(m3, __float_as_int(fabsf(q.w)));
            }
            m0 = __reduce_max_sync(0xFFFFFFFFu, m0); m1 = __reduce_max_sync(0xFFFFFFFFu, m1);
            m2 = __reduce_max_sync(0xFFFFFFFFu, m2); m3 = __reduce_max_sync(0xFFFFFFFFu, m3);
            if ((threadIdx.x & 31) == 0) {
                atomicMax(&smax[0], m0); atomicMax(&smax[1], m1); atomicMax(&smax[2], m2); atomicMax(&smax[3], m3);
            }
        } else {
            if (threadIdx.x < 4) smax[threadIdx.x] = 0;
            __syncthreads();
            float4* sp = reinterpret_cast<float4*>(sraw);
            const float4* g = pf + base + s0;
            int m0 = 0, m1 = 0, m2 = 0, m3 = 0;
            for (int j = threadIdx.x; j < rows; j += 128) {
                const float4 q = g[j];
                sp[j] = q;
                m0 = max(m0, __float_as_int(fabsf(q.x))); m1 = max(m1, __float_as_int(fabsf(q.y)));
                m2 = max(m2, __float_as_int(fabsf(q.z))); m3 = max(m3, __float_as_int(fabsf(q.w)));
            }
            m0 = __reduce_max_sync(0xFFFFFFFFu, m0); m1 = __reduce_max_sync(0xFFFFFFFFu, m1);
            m2 = __reduce_max_sync(0xFFFFFFFFu, m2); m3 = __reduce_max_sync(0xFFFFFFFFu, m3);
            if ((threadIdx.x & 31) == 0) {
                atomicMax(&smax[0], m0); atomicMax(&smax[1], m1); atomicMax(&smax[2], m2); atomicMax(&smax[3], m3);
            }
        }
        __syncthreads();
        // ---- filtered exact scoring (filters: cvmath.cuh; resolution of the survivors: filter_pass.cuh) ----
        // 64 correspondences at a time go through the fp32 filter; what it cannot reject is given the reference's
        // exact test.  skipFilter (sticky for the launch): every model of the warp fits (see filter_pass.cuh) -> plain loop.
        if (MODEL == RM_HOMOGRAPHY) {
            const float4* sp = reinterpret_cast<const float4*>(sraw);
            HFilterTile ht;
            h_filter_tile(Hf, __int_as_float(smax[0]), __int_as_float(smax[1]), __int_as_float(smax[2]),
                          __int_as_float(smax[3]), thr2, ht);
            struct MH { float h[8]; };
            MH mine;
#pragma unroll
            for (int j = 0; j < 8; ++j) mine.h[j] = Hf[j];
            auto shfl = [](const MH& a, int src) {
                MH r;
#pragma unroll
                for (int j = 0; j < 8; ++j) r.h[j] = __shfl_sync(0xFFFFFFFFu, a.h[j], src);
                return r;
            };
            for (int j0 = 0; j0 < rows; j0 += 64) {
                const int jn = min(64, rows - j0);
                auto test = [&](const MH& mm, int j) {
                    const float4 q = sp[j0 + j];
                    return h_error(mm.h, q.x, q.y, q.z, q.w) <= thr2;
                };
                unsigned long long pending = 0;
                if (skipFilter) {
                } else if (jn == 64) {
#pragma unroll 1
                    for (int o = 0; o < 4; ++o) {
                        unsigned w = 0;
#pragma unroll
                        for (int j = 0; j < 16; ++j) {
                            const float4 q = sp[j0 + 16 * o + j];
                            if (!h_certain_outlier(Hf, ht, q.x, q.y, q.z, q.w)) w |= 1u << j;
                        }
                        pending |= (unsigned long long)w << (16 * o);
                    }
                } else {
                    for (int j = 0; j < jn; ++j) {
                        const float4 q = sp[j0 + j];
                        if (!h_certain_outlier(Hf, ht, q.x, q.y, q.z, q.w)) pending |= 1ull << j;
                    }
                }
                if (!active) pending = 0;
                if (!skipFilter) skipFilter = warp_too_dense(pending);
                if (skipFilter) {
#pragma unroll 4
                    for (int j = 0; j < jn; ++j) cnt += test(mine, j) ? 1 : 0;
                } else {
                    cnt += resolve_pending(pending, mine, shfl, test);
                }
            }
        } else {
            struct MD { double m[9]; };
            MD mine;
#pragma unroll
            for (int j = 0; j < 9; ++j) mine.m[j] = m[j];
            auto shfl = [](const MD& a, int src) {
                MD r;
#pragma unroll
                for (int j = 0; j < 9; ++j) r.m[j] = __shfl_sync(0xFFFFFFFFu, a.m[j], src);
                return r;
            };
            const float* sfilt = (MODEL == RM_FUNDAMENTAL) ? reinterpret_cast<const float*>(sraw) : spair;
            const double4* spd = reinterpret_cast<const double4*>(sraw);
            FFilterTile ft;
            if (MODEL == RM_FUNDAMENTAL)
                f_filter_tile(ffil, __int_as_float(smax[0]), __int_as_float(smax[1]), __int_as_float(smax[2]),
                              __int_as_float(smax[3]), ft);
            else
                e_filter_tile(ffil, __int_as_float(smax[0]), __int_as_float(smax[1]), __int_as_float(smax[2]),
                              __int_as_float(smax[3]), ft);
            filter2_tile(ffil2, ft);
            for (int j0 = 0; j0 < rows; j0 += 64) {
                const int jn = min(64, rows - j0);
                auto test = [&](const MD& mm, int j) {
                    if (MODEL == RM_FUNDAMENTAL) {
                        const float* q = sfilt + 8 * ((j0 + j) >> 1) + ((j0 + j) & 1);
                        return f_inlier(mm.m, (double)q[0], (double)q[2], (double)q[4], (double)q[6], thr2);
                    } else {
                        const double4 q = spd[j0 + j];
                        return e_inlier(mm.m, q.x, q.y, q.z, q.w, thr2);
                    }
                };
                unsigned long long pending = 0;
                if (!skipFilter) {
                    pending = filter_block64<MODEL>(ffil2, sfilt + 4 * j0, jn);
                    if (!active) pending = 0;
                    skipFilter = warp_too_dense(pending);
                }
                if (skipFilter) {
#pragma unroll 2
                    for (int j = 0; j < jn; ++j) cnt += test(mine, j) ? 1 : 0;
                } else {
                    cnt += resolve_pending(pending, mine, shfl, test);
                }
            }
        }
        __syncthreads();
    }
    if (active) {
        int* c = counts + (int64_t)p * maxSlots + k;
        if (gridDim.z == 1) *c = cnt; else atomicAdd(c, cnt);
    }
}

// ---------------------------------------------------------------------------------------------
// select: OpenCV's sequential accept rule + adaptive iteration count, replayed over the counts
// ---------------------------------------------------------------------------------------------
// One block per pair.  Only the strict records of the count sequence (counts above everything before them and above
// modelPoints - 1) can be accepted by OpenCV's loop, and there are few of them (O(log n) for unordered counts), so the
// scan is split from the replay: every warp takes a contiguous quarter of the compacted model list -- pass 1 finds the
// quarter maxima, pass 2 re-reads the quarter knowing the maximum of everything before it and appends its records, in
// order, to a shared list (four independent loads in flight per lane: as one warp per pair walking 32 models at a time
// this kernel was a chain of dependent DRAM round trips) -- and thread 0 then replays the records in order, shrinking
// the iteration horizon with RANSACUpdateNumIters exactly as the sequential loop does.
constexpr int SEL_CAP = 384;      // records kept per warp (a quarter); beyond that thread 0 falls back to the plain scan

__global__ void __launch_bounds__(128)
select_kernel(const int64_t* __restrict__ off, int nPairs, int maxIters, int maxm, int mp,
              int adaptive, double confidence, const int* __restrict__ nIter,
              const int* __restrict__ flags, const int* __restrict__ iterStart,
              const int* __restrict__ slotList, const int* __restrict__ counts,
              int* __restrict__ best /*[nPairs][4]*/) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int p = blockIdx.x;
    const int n = (int)(off[p + 1] - off[p]);
    int* o = best + 4 * p;   // compacted index, count, iterations run, iteration*maxm+model
    const int fl = flags[p];
    const int* is = iterStart + (int64_t)p * (maxIters + 1);
    const int* c = counts + (int64_t)p * maxIters * maxm;
    const int* sl = slotList + (int64_t)p * maxIters * maxm;
    if (fl & (PAIR_FAIL | PAIR_DIRECT)) {
        if (threadIdx.x == 0) {
            o[0] = -1; o[1] = 0; o[2] = 0; o[3] = -1;
            if ((fl & PAIR_DIRECT) && !(fl & PAIR_FAIL) && is[1] - is[0] > 0) { o[0] = 0; o[1] = n; o[2] = 1; o[3] = 0; }
        }
        return;
    }
    __shared__ int sQMax[4], sCnt[4];
    __shared__ int sRecK[4][SEL_CAP], sRecX[4][SEL_CAP];
    const int total = is[min(nIter[p], maxIters)];
    const int Q = ((total + 3) / 4 + 31) & ~31;
    const int k0 = warp * Q, k1 = min(total, k0 + Q);
    // pass 1: maximum of the quarter
    int m = -1;
    for (int k = k0 + lane; k < k1; k += 32) m = max(m, c[k]);
    m = __reduce_max_sync(0xFFFFFFFFu, m);
    if (lane == 0) sQMax[warp] = m;
    __syncthreads();
    int carry = mp - 1;
    for (int w = 0; w < warp; ++w) carry = max(carry, sQMax[w]);
    // pass 2: records of the quarter, in order
    int cnt = 0;
    if (sQMax[warp] > carry) {               // otherwise nothing in this quarter can be a record
        for (int w0 = k0; w0 < k1; w0 += 128) {
            int x[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) { const int k = w0 + 32 * u + lane; x[u] = k < k1 ? c[k] : -1; }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                int pm = x[u];
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) {
                    const int t = __shfl_up_sync(0xFFFFFFFFu, pm, d);
                    if (lane >= d) pm = max(pm, t);
                }
                int excl = __shfl_up_sync(0xFFFFFFFFu, pm, 1);
                if (lane == 0) excl = -1;
                const bool rec = x[u] > max(excl, carry);
                const unsigned bal = __ballot_sync(0xFFFFFFFFu, rec);
                if (rec) {
                    const int at = cnt + __popc(bal & ((1u << lane) - 1u));
                    if (at < SEL_CAP) { sRecK[warp][at] = w0 + 32 * u + lane; sRecX[warp][at] = x[u]; }
                }
                cnt += __popc(bal);
                carry = max(carry, __shfl_sync(0xFFFFFFFFu, pm, 31));
            }
        }
    }
    if (lane == 0) sCnt[warp] = cnt;
    __syncthreads();
    if (threadIdx.x != 0) return;
    // OpenCV checks `iter < niters` once per iteration: every model of an iteration that was
    // started is evaluated even if an earlier model of the same iteration shrank niters.
    int niters = max(maxIters, 1);
    int maxGood = 0, bestK = -1, bestSlot = -1, curIt = -1;
    const bool overflow = sCnt[0] > SEL_CAP || sCnt[1] > SEL_CAP || sCnt[2] > SEL_CAP || sCnt[3] > SEL_CAP;
    auto accept = [&](int k, int x) -> bool {        // false: the horizon shrank past this record's iteration
        const int slot = sl[k];
        const int itb = slot / maxm;
        if (itb != curIt && itb >= niters) return false;
        curIt = itb; maxGood = x; bestK = k; bestSlot = slot;
        if (adaptive) niters = min(niters, update_num_iters(confidence, (double)(n - x) / n, mp, niters));
        return true;
    };
    if (!overflow) {
        bool go = true;
        for (int w = 0; w < 4 && go; ++w)
            for (int r = 0; r < sCnt[w] && go; ++r) go = accept(sRecK[w][r], sRecX[w][r]);
    } else {
        int run = mp - 1;
        for (int k = 0; k < total; ++k) {
            const int x = c[k];
            if (x > run) { run = x; if (!accept(k, x)) break; }
        }
    }
    // iterations OpenCV's loop would have executed
    o[0] = bestK; o[1] = maxGood; o[2] = min(niters, nIter[p]); o[3] = bestSlot;
}

// ---------------------------------------------------------------------------------------------
// LMedS (OpenCV's LMeDSPointSetRegistrator::run, calib3d/src/ptsetreg.cpp): what cv::findFundamentalMat runs
// instead of RANSAC below 15 points.  The iteration count is fixed up front (outlier ratio 0.45), a model's score
// is the median of its float32 errors (element n/2 in sorted order), the first strictly smaller median wins, and
// the mask comes from sigma = 2.5 * 1.4826 * (1 + 5 / (n - mp)) * sqrt(median), floored at 0.001.
// Small problems only (n <= LMEDS_MAXN): one thread per model holds the errors in registers / local memory.
// ---------------------------------------------------------------------------------------------
constexpr int LMEDS_MAXN = 32;

__global__ void __launch_bounds__(128)
lmeds_select_kernel(const float4* __restrict__ pf, const int64_t* __restrict__ off, const double* __restrict__ models,
                    const int* __restrict__ slotList, const int* __restrict__ nTotal, const int* __restrict__ nIter,
                    const int* __restrict__ flags, int maxSlots, int* __restrict__ best /*[np][4]*/,
                    float* __restrict__ bestMedian /*[np]*/) {
    const int p = blockIdx.x;
    const int64_t base = off[p];
    const int n = (int)(off[p + 1] - base);
    int* o = best + 4 * p;
    const int fl = flags[p];
    if (fl & (PAIR_FAIL | PAIR_DIRECT)) {
        if (threadIdx.x == 0) {
            o[0] = -1; o[1] = 0; o[2] = 0; o[3] = -1;
            if ((fl & PAIR_DIRECT) && !(fl & PAIR_FAIL) && nTotal[p] > 0) { o[0] = 0; o[1] = n; o[2] = 1; o[3] = 0; }
            bestMedian[p] = 0.f;
        }
        return;
    }
    __shared__ float4 sp[LMEDS_MAXN];
    __shared__ unsigned long long sKey[4];
    if (threadIdx.x < n) sp[threadIdx.x] = pf[base + threadIdx.x];
    __syncthreads();
    const int total = nTotal[p];
    unsigned long long key = ~0ull;
    for (int k = threadIdx.x; k < total; k += blockDim.x) {
        const int slot = slotList[(int64_t)p * maxSlots + k];
        const double* src = models + ((int64_t)p * maxSlots + slot) * 9;
        double m[9];
#pragma unroll
        for (int j = 0; j < 9; ++j) m[j] = src[j];
        float e[LMEDS_MAXN];
        for (int i = 0; i < n; ++i) e[i] = f_error(m, (double)sp[i].x, (double)sp[i].y, (double)sp[i].z, (double)sp[i].w);
        // element n/2 of the sorted errors (rank by value, index breaks ties)
        float med = __int_as_float(0x7FC00000);
        bool anyNan = false;
        for (int i = 0; i < n; ++i) anyNan |= !(e[i] == e[i]);
        if (!anyNan) {
            for (int i = 0; i < n; ++i) {
                int rank = 0;
                for (int j = 0; j < n; ++j) rank += (e[j] < e[i] || (e[j] == e[i] && j < i)) ? 1 : 0;
                if (rank == n / 2) med = e[i];
            }
        }
        // `median < minMedian` with minMedian starting at DBL_MAX: only finite medians compete; ties keep the earlier
        if (med == med && med < __int_as_float(0x7F800000))
            key = min(key, ((unsigned long long)__float_as_uint(med) << 32) | (unsigned)k);   // errors are >= 0: bit order = value order
    }
    for (int d = 16; d > 0; d >>= 1) key = min(key, __shfl_xor_sync(0xFFFFFFFFu, key, d));
    if ((threadIdx.x & 31) == 0) sKey[threadIdx.x >> 5] = key;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < (int)blockDim.x / 32; ++w) key = min(key, sKey[w]);
        if (key == ~0ull) { o[0] = -1; o[1] = 0; o[2] = nIter[p]; o[3] = -1; bestMedian[p] = 0.f; }
        else {
            const int k = (int)(unsigned)key;
            o[0] = k; o[1] = 0; o[2] = nIter[p]; o[3] = slotList[(int64_t)p * maxSlots + k];
            bestMedian[p] = __uint_as_float((unsigned)(key >> 32));
        }
    }
}

__global__ void __launch_bounds__(64)
finalize_lmeds_kernel(const float4* __restrict__ pf, const int64_t* __restrict__ off, const double* __restrict__ models,
                      int maxSlots, int maxm, int mp, const int* __restrict__ best, const int* __restrict__ flags,
                      const float* __restrict__ bestMedian, double* __restrict__ outModels, uint8_t* __restrict__ masks,
                      int* __restrict__ inlierCounts, uint8_t* __restrict__ ok, int* __restrict__ itersRun,
                      int* __restrict__ bestOut) {
    const int p = blockIdx.x;
    const int64_t base = off[p];
    const int n = (int)(off[p + 1] - base);
    const int* b = best + 4 * p;
    const int slot = b[3];
    if (threadIdx.x == 0 && itersRun) itersRun[p] = b[2];
    if (slot < 0) {
        for (int i = threadIdx.x; i < n; i += blockDim.x) masks[base + i] = 0;
        if (threadIdx.x == 0) {
            ok[p] = 0;
            if (inlierCounts) inlierCounts[p] = 0;
            if (bestOut) { bestOut[2 * p] = -1; bestOut[2 * p + 1] = -1; }
        }
        return;
    }
    const double* src = models + ((int64_t)p * maxSlots + slot) * 9;
    double m[9];
#pragma unroll
    for (int j = 0; j < 9; ++j) m[j] = src[j];
    const bool direct = (flags[p] & PAIR_DIRECT) != 0;
    double sigma = 2.5 * 1.4826 * (1 + 5. / (n - mp)) * sqrt((double)bestMedian[p]);
    sigma = fmax(sigma, 0.001);
    const float t = (float)(sigma * sigma);
    int cnt = 0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const float4 q = pf[base + i];
        const bool in = direct || f_error(m, (double)q.x, (double)q.y, (double)q.z, (double)q.w) <= t;
        masks[base + i] = in ? 1 : 0;
        cnt += in ? 1 : 0;
    }
    __shared__ int sc[2];
    for (int d = 16; d > 0; d >>= 1) cnt += __shfl_xor_sync(0xFFFFFFFFu, cnt, d);
    if ((threadIdx.x & 31) == 0) sc[threadIdx.x >> 5] = cnt;
    __syncthreads();
    if (threadIdx.x == 0) {
        const int c = sc[0] + sc[1];
        if (inlierCounts) inlierCounts[p] = c;
        ok[p] = c >= mp ? 1 : 0;
        for (int j = 0; j < 9; ++j) outModels[9 * p + j] = m[j];
        if (bestOut) { bestOut[2 * p] = slot / maxm; bestOut[2 * p + 1] = slot % maxm; }
    }
}

// ---------------------------------------------------------------------------------------------
// finalize
// ---------------------------------------------------------------------------------------------
constexpr int FB = 128;   // finalize block size (several blocks per SM overlap the serial thread-0 phases)
// The homography refit also runs with 64-thread blocks (large batches: all pairs resident at once, 7 blocks per SM), so its
// loops and reductions use blockDim.x; shared arrays are sized for FB.

template <int K>
__device__ __forceinline__ void block_sum(double (&v)[K], double* sred /*[K * 8]*/, double* sout /*[K]*/) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < K; ++k) {
        double x = v[k];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xFFFFFFFFu, x, o);
        if (lane == 0) sred[warp * K + k] = x;
    }
    __syncthreads();
    if (threadIdx.x < K) {
        double s = 0;
        for (int w = 0; w < (int)blockDim.x / 32; ++w) s += sred[w * K + threadIdx.x];
        sout[threadIdx.x] = s;
    }
    __syncthreads();
}

__device__ __forceinline__ double block_max(double x, double* sred) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) x = fmax(x, __shfl_xor_sync(0xFFFFFFFFu, x, o));
    if (lane == 0) sred[warp] = x;
    __syncthreads();
    double r = sred[0];
    for (int w = 1; w < (int)blockDim.x / 32; ++w) r = fmax(r, sred[w]);
    __syncthreads();
    return r;
}

// F / E (and H without refinement): mask of the winner, model copied out.
template <int MODEL>
__global__ void __launch_bounds__(FB)
finalize_plain_kernel(const float4* __restrict__ pf, const double* __restrict__ pd,
                      const int64_t* __restrict__ off, const double* __restrict__ models,
                      int maxSlots, const int* __restrict__ best, const int* __restrict__ flags,
                      const int* __restrict__ nTotal, float thr2, double* __restrict__ outModels,
                      uint8_t* __restrict__ masks,
                      int* __restrict__ inlierCounts, uint8_t* __restrict__ ok,
                      int* __restrict__ itersRun, int* __restrict__ bestOut) {
    const int p = blockIdx.x;
    const int64_t base = off[p];
    const int n = (int)(off[p + 1] - base);
    const int* b = best + 4 * p;
    const int slot = b[3];
    __shared__ int sCnt[FB / 32];
    if (threadIdx.x == 0 && itersRun) itersRun[p] = b[2];
    if (slot < 0) {
        for (int i = threadIdx.x; i < n; i += FB) masks[base + i] = 0;
        if (threadIdx.x == 0) {
            ok[p] = 0;
            if (inlierCounts) inlierCounts[p] = 0;
            if (bestOut) { bestOut[2 * p] = -1; bestOut[2 * p + 1] = -1; }
        }
        return;
    }
    const double* src = models + ((int64_t)p * maxSlots + slot) * 9;
    double m[9];
    float Hf[8];
#pragma unroll
    for (int j = 0; j < 9; ++j) m[j] = src[j];
#pragma unroll
    for (int j = 0; j < 8; ++j) Hf[j] = (float)m[j];
    const bool direct = (flags[p] & PAIR_DIRECT) != 0;
    int cnt = 0;
    for (int i = threadIdx.x; i < n; i += FB) {
        bool in;
        if (direct) in = true;
        else if (MODEL == RM_HOMOGRAPHY) { const float4 q = pf[base + i]; in = h_error(Hf, q.x, q.y, q.z, q.w) <= thr2; }
        else if (MODEL == RM_FUNDAMENTAL) { const float4 q = pf[base + i]; in = f_inlier(m, (double)q.x, (double)q.y, (double)q.z, (double)q.w, thr2); }
        else { const double* q = pd + 4 * (base + i); in = e_inlier(m, q[0], q[1], q[2], q[3], thr2); }
        masks[base + i] = in ? 1 : 0;
        cnt += in ? 1 : 0;
    }
    for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xFFFFFFFFu, cnt, o);
    if ((threadIdx.x & 31) == 0) sCnt[threadIdx.x >> 5] = cnt;
    __syncthreads();
    if (threadIdx.x == 0) {
        int t = 0;
        for (int w = 0; w < FB / 32; ++w) t += sCnt[w];
        if (inlierCounts) inlierCounts[p] = t;
        // E with N == 5 and several solutions: OpenCV returns them stacked and the reference
        // then reports failure (E not 3x3, MiniCVNative.cpp:184-193)
        ok[p] = (MODEL == RM_ESSENTIAL && direct && nTotal[p] != 1) ? 0 : 1;
        for (int j = 0; j < 9; ++j) outModels[9 * p + j] = m[j];
        if (bestOut) { bestOut[2 * p] = slot / MT<MODEL>::MAXM; bestOut[2 * p + 1] = slot % MT<MODEL>::MAXM; }
    }
}

// Thread-0 helpers of the homography refit, kept out of line so the kernel's register and
// stack footprint stay small.  -DMCV_DEBUG_INLINE_REFIT builds the fully inlined form that round 1 saw mis-execute
// (eigenvectors returned as the identity) -- kept buildable for the sanitizer runs recorded in profiles/.
#ifdef MCV_DEBUG_INLINE_REFIT
#define MCV_REFIT_NOINLINE __forceinline__
#else
#define MCV_REFIT_NOINLINE __noinline__
#endif
// H from the 45 upper-triangle sums of L'L (runKernel of the refit), executed by one warp: the symmetric 9 x 9
// eigen-decomposition (cvmath.cuh, jacobi_eigen_sym: same operations) is the longest serial stretch of the refit,
// and inside a Jacobi rotation every element update is independent, so lanes 0-8 take the column / row updates of L
// and lanes 16-24 the eigenvector rows.  No sum changes its order (the rotation angle and the convergence sums are
// computed redundantly by every lane, in the serial order), so H is bit-identical to the single-thread form's.
__device__ MCV_REFIT_NOINLINE void h_from_ltl_warp(const double* ltl45, double cmx, double cmy, double cMx, double cMy,
                                             double smx, double smy, double sMx, double sMy, double* H,
                                             double* L /*smem [81]*/, double* V /*smem [81]*/) {
    const int lane = threadIdx.x & 31;
    if (lane == 0) {
        int k = 0;
        for (int r = 0; r < 9; ++r)
            for (int c = r; c < 9; ++c) { L[r * 9 + c] = ltl45[k]; L[c * 9 + r] = ltl45[k]; ++k; }
    }
    for (int idx = lane; idx < 81; idx += 32) V[idx] = (idx / 9 == idx % 9) ? 1.0 : 0.0;
    __syncwarp();
    for (int sweep = 0; sweep < 60; ++sweep) {
        double off = 0, diag = 0;
        for (int i = 0; i < 9; ++i) {
            diag += L[i * 9 + i] * L[i * 9 + i];
            for (int j = i + 1; j < 9; ++j) off += L[i * 9 + j] * L[i * 9 + j];
        }
        if (off <= 1e-32 * diag || off == 0) break;
        for (int p = 0; p < 8; ++p)
            for (int q = p + 1; q < 9; ++q) {
                const double apq = L[p * 9 + q];
                if (apq == 0) continue;
                const double app = L[p * 9 + p], aqq = L[q * 9 + q];
                const double theta = (aqq - app) / (2 * apq);
                const double t = (theta >= 0 ? 1. : -1.) / (fabs(theta) + sqrt(theta * theta + 1));
                const double c = 1 / sqrt(t * t + 1), sn = t * c;
                __syncwarp();                                   // everyone has read app, aqq, apq
                if (lane < 9) {
                    const int k = lane;
                    const double akp = L[k * 9 + p], akq = L[k * 9 + q];
                    L[k * 9 + p] = c * akp - sn * akq;
                    L[k * 9 + q] = sn * akp + c * akq;
                } else if (lane >= 16 && lane < 25) {
                    const int k = lane - 16;
                    const double vp = V[p * 9 + k], vq = V[q * 9 + k];
                    V[p * 9 + k] = c * vp - sn * vq;
                    V[q * 9 + k] = sn * vp + c * vq;
                }
                __syncwarp();
                if (lane < 9) {
                    const int k = lane;
                    const double apk = L[p * 9 + k], aqk = L[q * 9 + k];
                    L[p * 9 + k] = c * apk - sn * aqk;
                    L[q * 9 + k] = sn * apk + c * aqk;
                }
                __syncwarp();
            }
    }
    if (lane == 0) {
        int mi = 0;
        for (int i = 1; i < 9; ++i) if (L[i * 9 + i] < L[mi * 9 + mi]) mi = i;
        const double* h = V + 9 * mi;
        const double iH[9] = {1. / smx, 0, cmx, 0, 1. / smy, cmy, 0, 0, 1};
        const double N2[9] = {sMx, 0, -cMx * sMx, 0, sMy, -cMy * sMy, 0, 0, 1};
        double T[9], Hn[9];
        mat3mul(iH, h, T);
        mat3mul(T, N2, Hn);
        if (Hn[8] != 0) {
            const double sc = 1. / Hn[8];
            for (int j = 0; j < 9; ++j) H[j] = Hn[j] * sc;
        }
    }
}

// Symmetric N x N eigen-decomposition by one warp, in shared memory: the operations of cvmath.cuh jacobi_eigen_sym in the same
// order (rotation angle and convergence sums computed redundantly by every lane, element updates of a rotation spread over
// the lanes), so w and V are bit-identical to the single-thread form's.  A is destroyed (diagonal = eigenvalues), V rows =
// eigenvectors.  All 32 lanes of the warp must call.
template <int N>
__device__ __forceinline__ void jacobi_eigen_sym_warp(double* A, double* V) {
    const int lane = threadIdx.x & 31;
    for (int idx = lane; idx < N * N; idx += 32) V[idx] = (idx / N == idx % N) ? 1.0 : 0.0;
    __syncwarp();
    for (int sweep = 0; sweep < 60; ++sweep) {
        double off = 0, diag = 0;
        for (int i = 0; i < N; ++i) {
            diag += A[i * N + i] * A[i * N + i];
            for (int j = i + 1; j < N; ++j) off += A[i * N + j] * A[i * N + j];
        }
        if (off <= 1e-32 * diag || off == 0) break;
        for (int p = 0; p < N - 1; ++p)
            for (int q = p + 1; q < N; ++q) {
                const double apq = A[p * N + q];
                if (apq == 0) continue;
                const double app = A[p * N + p], aqq = A[q * N + q];
                const double theta = (aqq - app) / (2 * apq);
                const double t = (theta >= 0 ? 1. : -1.) / (fabs(theta) + sqrt(theta * theta + 1));
                const double c = 1 / sqrt(t * t + 1), sn = t * c;
                __syncwarp();                                   // everyone has read app, aqq, apq
                if (lane < N) {
                    const int k = lane;
                    const double akp = A[k * N + p], akq = A[k * N + q];
                    A[k * N + p] = c * akp - sn * akq;
                    A[k * N + q] = sn * akp + c * akq;
                } else if (lane >= 16 && lane < 16 + N) {
                    const int k = lane - 16;
                    const double vp = V[p * N + k], vq = V[q * N + k];
                    V[p * N + k] = c * vp - sn * vq;
                    V[q * N + k] = sn * vp + c * vq;
                }
                __syncwarp();
                if (lane < N) {
                    const int k = lane;
                    const double apk = A[p * N + k], aqk = A[q * N + k];
                    A[p * N + k] = c * apk - sn * aqk;
                    A[q * N + k] = sn * apk + c * aqk;
                }
                __syncwarp();
            }
    }
    __syncwarp();
}

// One LM trial step by warp 0: d = solve(A + lambda diag(D), v), xd = x - d.  Lane 0 tries the Cholesky fast path of
// cvmath.cuh solve_sym_chol_or_eig; homography normal equations are badly scaled (the two projective parameters), so the
// thresholded eigen-solve (cv::solve DECOMP_EIG) is the usual path, and its Jacobi sweeps run on the whole warp.
// sM, sVe: shared scratch, 64 doubles each.
__device__ MCV_REFIT_NOINLINE void lm_trial_warp(const double* A, const double* v, const double* D, double lam,
                                                 const double* x, double* d, double* xd, double* sM, double* sVe) {
    const int lane = threadIdx.x & 31;
    int ok = 0;
    if (lane == 0) {
        double Ap[64], vv[8], dd[8];
        for (int k = 0; k < 64; ++k) Ap[k] = A[k];
        for (int r = 0; r < 8; ++r) { Ap[r * 8 + r] += lam * D[r]; vv[r] = v[r]; }
        ok = cvm::solve_sym_chol<8>(Ap, vv, dd) ? 1 : 0;
        if (ok) for (int r = 0; r < 8; ++r) { d[r] = dd[r]; xd[r] = x[r] - dd[r]; }
        else for (int k = 0; k < 64; ++k) sM[k] = Ap[k];
    }
    ok = __shfl_sync(0xFFFFFFFFu, ok, 0);
    if (ok) return;
    __syncwarp();
    jacobi_eigen_sym_warp<8>(sM, sVe);
    if (lane == 0) {                                     // cvmath.cuh solve_sym_eig, after the decomposition
        double w[8], dd[8];
        double sum = 0;
        for (int i = 0; i < 8; ++i) { w[i] = sM[i * 8 + i]; sum += fabs(w[i]); }
        const double thr = DBL_EPSILON * 2 * sum;
        for (int i = 0; i < 8; ++i) dd[i] = 0;
        for (int i = 0; i < 8; ++i) {
            if (fabs(w[i]) <= thr) continue;
            double sdot = 0;
            for (int k = 0; k < 8; ++k) sdot += sVe[i * 8 + k] * v[k];
            sdot /= w[i];
            for (int k = 0; k < 8; ++k) dd[k] += sdot * sVe[i * 8 + k];
        }
        for (int r = 0; r < 8; ++r) { d[r] = dd[r]; xd[r] = x[r] - dd[r]; }
    }
    __syncwarp();
}

// cvmath.cuh max_diag_pinv_sym<8> by warp 0 (valid in lane 0)
__device__ MCV_REFIT_NOINLINE double lm_reseed_warp(const double* A, double* sM, double* sVe) {
    const int lane = threadIdx.x & 31;
    for (int k = lane; k < 64; k += 32) sM[k] = A[k];
    __syncwarp();
    jacobi_eigen_sym_warp<8>(sM, sVe);
    double maxval = DBL_EPSILON;
    if (lane == 0) {
        double w[8];
        double sum = 0;
        for (int i = 0; i < 8; ++i) { w[i] = sM[i * 8 + i]; sum += fabs(w[i]); }
        const double thr = DBL_EPSILON * 2 * sum;
        for (int k = 0; k < 8; ++k) {
            double dsum = 0;
            for (int i = 0; i < 8; ++i)
                if (fabs(w[i]) > thr) dsum += sVe[i * 8 + k] * sVe[i * 8 + k] / w[i];
            maxval = fmax(maxval, fabs(dsum));
        }
    }
    __syncwarp();
    return maxval;
}

// Homography with OpenCV's post-processing: runKernel on the inliers, LMSolver(10), mask
// recomputed from the refined H, H /= h33 (cv::findHomography, calib3d/src/fundam.cpp).
__global__ void __launch_bounds__(FB, 4)
finalize_homography_kernel(const float4* __restrict__ pf, const int64_t* __restrict__ off,
                           const double* __restrict__ models, int maxSlots,
                           const int* __restrict__ best, const int* __restrict__ flags, float thr2,
                           double* __restrict__ outModels, uint8_t* __restrict__ masks,
                           int* __restrict__ inlierCounts, uint8_t* __restrict__ ok,
                           int* __restrict__ itersRun, int* __restrict__ bestOut) {
    const int p = blockIdx.x;
    const int tid = threadIdx.x;
    const int64_t base = off[p];
    const int n = (int)(off[p + 1] - base);
    const int* b = best + 4 * p;
    const int slot = b[3];
    __shared__ double sred[45 * (FB / 32)];
    __shared__ double sout[48];
    __shared__ double sH[9];      // current parameters x (8) + 1
    __shared__ double sXd[9];     // trial parameters
    __shared__ double sCtl[8];    // [0] continue flag
    if (tid == 0) {
        if (itersRun) itersRun[p] = b[2];
        if (bestOut) { bestOut[2 * p] = slot; bestOut[2 * p + 1] = slot < 0 ? -1 : 0; }
    }
    if (slot < 0) {
        for (int i = tid; i < n; i += (int)blockDim.x) masks[base + i] = 0;
        if (tid == 0) { ok[p] = 0; if (inlierCounts) inlierCounts[p] = 0; }
        return;
    }
    const double* src = models + ((int64_t)p * maxSlots + slot) * 9;
    if (tid < 9) sH[tid] = src[tid];
    __syncthreads();
    const bool direct = (flags[p] & PAIR_DIRECT) != 0;
    float Hf[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) Hf[j] = (float)sH[j];

    // RANSAC mask of the winner
    for (int i = tid; i < n; i += (int)blockDim.x) {
        const float4 q = pf[base + i];
        masks[base + i] = (direct || h_error(Hf, q.x, q.y, q.z, q.w) <= thr2) ? 1 : 0;
    }
    __syncthreads();
    if (direct) {   // 4 points: kernel only, no refinement
        if (tid == 0) {
            ok[p] = 1;
            if (inlierCounts) inlierCounts[p] = n;
            for (int j = 0; j < 9; ++j) outModels[9 * p + j] = sH[j];
        }
        return;
    }

    // ---- runKernel on the inliers: centroids, mean absolute deviations, LtL ----
    {
        double v[5] = {0, 0, 0, 0, 0};
        for (int i = tid; i < n; i += (int)blockDim.x)
            if (masks[base + i]) {
                const float4 q = pf[base + i];
                v[0] += q.z; v[1] += q.w; v[2] += q.x; v[3] += q.y; v[4] += 1;
            }
        block_sum<5>(v, sred, sout);
    }
    const double cnt = sout[4];
    const double cmx = sout[0] / cnt, cmy = sout[1] / cnt, cMx = sout[2] / cnt, cMy = sout[3] / cnt;
    __syncthreads();
    {
        double v[4] = {0, 0, 0, 0};
        for (int i = tid; i < n; i += (int)blockDim.x)
            if (masks[base + i]) {
                const float4 q = pf[base + i];
                v[0] += fabs(q.z - cmx); v[1] += fabs(q.w - cmy);
                v[2] += fabs(q.x - cMx); v[3] += fabs(q.y - cMy);
            }
        block_sum<4>(v, sred, sout);
    }
    double smx = sout[0], smy = sout[1], sMx = sout[2], sMy = sout[3];
    __syncthreads();
    const bool kernel_ok = !(fabs(smx) < DBL_EPSILON || fabs(smy) < DBL_EPSILON ||
                             fabs(sMx) < DBL_EPSILON || fabs(sMy) < DBL_EPSILON);
    if (kernel_ok) {
        smx = cnt / smx; smy = cnt / smy; sMx = cnt / sMx; sMy = cnt / sMy;
        double v[45];
#pragma unroll
        for (int k = 0; k < 45; ++k) v[k] = 0;
        for (int i = tid; i < n; i += (int)blockDim.x)
            if (masks[base + i]) {
                const float4 q = pf[base + i];
                const double x = (q.z - cmx) * smx, y = (q.w - cmy) * smy;
                const double X = (q.x - cMx) * sMx, Y = (q.y - cMy) * sMy;
                const double Lx[9] = {X, Y, 1, 0, 0, 0, -x * X, -x * Y, -x};
                const double Ly[9] = {0, 0, 0, X, Y, 1, -y * X, -y * Y, -y};
                int k = 0;
#pragma unroll
                for (int r = 0; r < 9; ++r)
#pragma unroll
                    for (int c = r; c < 9; ++c) { v[k] += Lx[r] * Lx[c] + Ly[r] * Ly[c]; ++k; }
            }
        block_sum<45>(v, sred, sout);
        __shared__ double sL[81], sV[81];
        if (tid < 32) h_from_ltl_warp(sout, cmx, cmy, cMx, cMy, smx, smy, sMx, sMy, sH, sL, sV);
        __syncthreads();
    }

    // ---- LMSolver, 10 iterations, on the 8 free parameters ----
    // state held by thread 0: A (8x8), v (8), D (8), S, lambda, lc; everyone helps with sums
    __shared__ double sA[64], sV[8], sD[8], sS, sLam, sLc, sRmax;
    auto accumulate = [&](const double* x) {   // A, v, S, max|r| at parameters x
        double v[45];
#pragma unroll
        for (int k = 0; k < 45; ++k) v[k] = 0;
        double rmax = 0;
        for (int i = tid; i < n; i += (int)blockDim.x)
            if (masks[base + i]) {
                const float4 q = pf[base + i];
                const double Mx = q.x, My = q.y;
                double ww = x[6] * Mx + x[7] * My + 1.;
                ww = fabs(ww) > DBL_EPSILON ? 1. / ww : 0;
                const double xi = (x[0] * Mx + x[1] * My + x[2]) * ww;
                const double yi = (x[3] * Mx + x[4] * My + x[5]) * ww;
                const double r0 = xi - q.z, r1 = yi - q.w;
                const double J0[8] = {Mx * ww, My * ww, ww, 0, 0, 0, -Mx * ww * xi, -My * ww * xi};
                const double J1[8] = {0, 0, 0, Mx * ww, My * ww, ww, -Mx * ww * yi, -My * ww * yi};
                int k = 0;
#pragma unroll
                for (int r = 0; r < 8; ++r)
#pragma unroll
                    for (int c = r; c < 8; ++c) { v[k] += J0[r] * J0[c] + J1[r] * J1[c]; ++k; }
#pragma unroll
                for (int r = 0; r < 8; ++r) v[36 + r] += J0[r] * r0 + J1[r] * r1;
                v[44] += r0 * r0 + r1 * r1;
                rmax = fmax(rmax, fmax(fabs(r0), fabs(r1)));
            }
        block_sum<45>(v, sred, sout);
        const double rm = block_max(rmax, sred);
        if (tid == 0) {
            int k = 0;
            for (int r = 0; r < 8; ++r)
                for (int c = r; c < 8; ++c) { sA[r * 8 + c] = sout[k]; sA[c * 8 + r] = sout[k]; ++k; }
            for (int r = 0; r < 8; ++r) sV[r] = sout[36 + r];
            sS = sout[44];
            sRmax = rm;
        }
        __syncthreads();
    };
    auto residual_sq = [&](const double* x) -> double {   // Sd at trial parameters
        double v[1] = {0};
        for (int i = tid; i < n; i += (int)blockDim.x)
            if (masks[base + i]) {
                const float4 q = pf[base + i];
                const double Mx = q.x, My = q.y;
                double ww = x[6] * Mx + x[7] * My + 1.;
                ww = fabs(ww) > DBL_EPSILON ? 1. / ww : 0;
                const double r0 = (x[0] * Mx + x[1] * My + x[2]) * ww - q.z;
                const double r1 = (x[3] * Mx + x[4] * My + x[5]) * ww - q.w;
                v[0] += r0 * r0 + r1 * r1;
            }
        block_sum<1>(v, sred, sout);
        const double r = sout[0];
        __syncthreads();
        return r;
    };

    accumulate(sH);
    if (tid == 0) {
        for (int r = 0; r < 8; ++r) sD[r] = sA[r * 8 + r];
        sLam = 1; sLc = 0.75;
    }
    __syncthreads();
    __shared__ double sDvec[8];
    __shared__ double sEigM[64], sEigV[64];
    for (int iter = 0; iter < 10; ++iter) {
        if (tid < 32) lm_trial_warp(sA, sV, sD, sLam, sH, sDvec, sXd, sEigM, sEigV);
        __syncthreads();
        const double Sd = residual_sq(sXd);
        if (tid < 32) {
            const double S = sS;
            int reseed = 0;
            double nu = 0;
            if (tid == 0) {
                double dS = 0, tdv = 0;
                for (int r = 0; r < 8; ++r) {
                    double Ad = 0;
                    for (int c = 0; c < 8; ++c) Ad += sA[r * 8 + c] * sDvec[c];
                    dS += sDvec[r] * (-Ad + 2 * sV[r]);
                    tdv += sDvec[r] * sV[r];
                }
                const double R = (S - Sd) / (fabs(dS) > DBL_EPSILON ? dS : 1);
                if (R > 0.75) {
                    sLam *= 0.5;
                    if (sLam < sLc) sLam = 0;
                } else if (R < 0.25) {
                    nu = (Sd - S) / (fabs(tdv) > DBL_EPSILON ? tdv : 1) + 2;
                    nu = fmin(fmax(nu, 2.), 10.);
                    if (sLam == 0) reseed = 1;
                    else sLam *= nu;
                }
#ifdef MCV_DEBUG_REFIT
                printf(" lm iter=%d S=%g Sd=%g dS=%g R=%g lam=%g d0=%g\n", iter, S, Sd, dS, R, sLam, sDvec[0]);
#endif
            }
            reseed = __shfl_sync(0xFFFFFFFFu, reseed, 0);
            if (reseed) {                                 // lambda was switched off: restart it from the pseudo-inverse's diagonal
                const double maxval = lm_reseed_warp(sA, sEigM, sEigV);
                if (tid == 0) {
                    sLam = sLc = 1. / maxval;
                    nu *= 0.5;
                    sLam *= nu;
                }
            }
            if (tid == 0) {
                sCtl[0] = (Sd < S) ? 1 : 0;
                if (Sd < S) for (int r = 0; r < 8; ++r) sH[r] = sXd[r];
            }
        }
        __syncthreads();
        if (sCtl[0] != 0) accumulate(sH);
        if (tid == 0) {
            double dinf = 0;
            for (int r = 0; r < 8; ++r) dinf = fmax(dinf, fabs(sDvec[r]));
            const bool proceed = (iter + 1 < 10) && dinf >= (double)FLT_EPSILON && sRmax >= (double)FLT_EPSILON;
            sCtl[1] = proceed ? 1 : 0;
        }
        __syncthreads();
        if (sCtl[1] == 0) break;
    }
    if (tid == 0) sH[8] = 1.0;
    __syncthreads();

    // ---- mask recomputed from the refined H ----
#pragma unroll
    for (int j = 0; j < 8; ++j) Hf[j] = (float)sH[j];
    int c2 = 0;
    for (int i = tid; i < n; i += (int)blockDim.x) {
        const float4 q = pf[base + i];
        const bool in = h_error(Hf, q.x, q.y, q.z, q.w) <= thr2;
        masks[base + i] = in ? 1 : 0;
        c2 += in ? 1 : 0;
    }
    {
        double v[1] = {(double)c2};
        block_sum<1>(v, sred, sout);
    }
    if (tid == 0) {
        ok[p] = 1;
        if (inlierCounts) inlierCounts[p] = (int)sout[0];
        for (int j = 0; j < 9; ++j) outModels[9 * p + j] = sH[j];
    }
}

// debug taps: expand compacted counts back to [maxIters][MAXM] (-1 = no model)
__global__ void __launch_bounds__(256)
dbg_counts_kernel(const int* __restrict__ iterStart, const int* __restrict__ counts, int maxIters,
                  int maxm, int* __restrict__ out) {
    const int p = blockIdx.y;
    const int it = blockIdx.x * blockDim.x + threadIdx.x;
    if (it >= maxIters) return;
    const int* is = iterStart + (int64_t)p * (maxIters + 1);
    const int k0 = is[it], k1 = is[it + 1];
    for (int m = 0; m < maxm; ++m)
        out[((int64_t)p * maxIters + it) * maxm + m] = (m < k1 - k0) ? counts[(int64_t)p * maxIters * maxm + k0 + m] : -1;
}

// ---------------------------------------------------------------------------------------------
// Adaptive stop in two phases.  OpenCV's loop shrinks its iteration horizon every time a better model turns up and
// never grows it again, so once the first K iterations are scored, nothing at or beyond the horizon those K
// iterations leave can ever be evaluated.  Phase 1 scores iterations [0, K), horizon_kernel replays the accept /
// RANSACUpdateNumIters rule over them (the rule select_kernel applies to the whole list afterwards) and cuts the
// pair's iteration count to that horizon, phase 2 scores what is left below it.  On ratio-tested matches the horizon
// after 128 iterations is a few dozen: 2000 hypotheses per pair are solved, ~128 are scored.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
phase_limit_kernel(int np, int maxIters, int K, const int* __restrict__ nIter, const int* __restrict__ iterStart,
                   const int* __restrict__ nTotal, const int* __restrict__ gate, int* __restrict__ phaseEnd) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= np) return;
    // the pilot saw (almost) no fitting model: the horizon will not shrink, phase 1 is the whole list (no phase 2)
    if (gate && !score_gate_split(gate)) { phaseEnd[p] = nTotal[p]; return; }
    phaseEnd[p] = iterStart[(int64_t)p * (maxIters + 1) + min(min(K, nIter[p]), maxIters)];
}
__global__ void __launch_bounds__(128)
horizon_kernel(const int64_t* __restrict__ off, int np, int maxIters, int maxm, int mp, int K, double confidence,
               int* __restrict__ nIter, const int* __restrict__ flags, const int* __restrict__ iterStart,
               const int* __restrict__ slotList, const int* __restrict__ counts, const int* __restrict__ gate,
               int* __restrict__ phaseBegin, int* __restrict__ phaseEnd) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= np) return;
    const int* is = iterStart + (int64_t)p * (maxIters + 1);
    const int scored = phaseEnd[p];                      // models of the first K iterations (or all of them)
    phaseBegin[p] = scored;
    if ((gate && !score_gate_split(gate)) || (flags[p] & (PAIR_FAIL | PAIR_DIRECT))) { phaseEnd[p] = scored; return; }
    const int n = (int)(off[p + 1] - off[p]);
    const int* c = counts + (int64_t)p * maxIters * maxm;
    const int* sl = slotList + (int64_t)p * maxIters * maxm;
    int niters = max(maxIters, 1), run = mp - 1, curIt = -1;
    for (int k = 0; k < scored; ++k) {
        const int x = c[k];
        if (x > run) {
            const int itb = sl[k] / maxm;
            if (itb != curIt && itb >= niters) break;    // the loop had already stopped
            curIt = itb; run = x;
            niters = min(niters, update_num_iters(confidence, (double)(n - x) / n, mp, niters));
        }
    }
    const int lim = min(nIter[p], max(niters, min(K, nIter[p])));
    nIter[p] = lim;
    phaseEnd[p] = is[min(lim, maxIters)];
}

template <int MODEL>
int run_model(Context* ctx, cudaStream_t st, const RansacParams& prm, const RansacIO& io) {
    constexpr int MP = MT<MODEL>::MP, MAXM = MT<MODEL>::MAXM;
    const int nPairs = io.nPairs;
    const int maxIters = std::max(prm.maxIters, 1);
    const int64_t total = io.hOff[nPairs];
    const float thr2 = MODEL == RM_ESSENTIAL
                           ? (float)((prm.threshold / ((prm.focal + prm.focal) / 2)) * (prm.threshold / ((prm.focal + prm.focal) / 2)))
                           : (float)(prm.threshold * prm.threshold);

    // device offsets (whole batch), staged through pinned memory
    const size_t n1 = (size_t)nPairs + 1;
    if (ctx->pin[1].reserve(n1 * 8) || ctx->ws[10].reserve(n1 * 8)) return -1;
    MCV_CUDA_TRY(cudaEventSynchronize(ctx->ev[1]));
    memcpy(ctx->pin[1].p, io.hOff, n1 * 8);
    MCV_CUDA_TRY(cudaMemcpyAsync(ctx->ws[10].p, ctx->pin[1].p, n1 * 8, cudaMemcpyHostToDevice, st));
    MCV_CUDA_TRY(cudaEventRecord(ctx->ev[1], st));
    const int64_t* dOff = ctx->ws[10].as<int64_t>();

    // working point layout
    float4* pf = nullptr;
    double* pd = nullptr;
    if (MODEL == RM_ESSENTIAL) {
        if (ctx->ws[11].reserve((size_t)std::max<int64_t>(total, 1) * 32)) return -1;
        pd = ctx->ws[11].as<double>();
    } else {
        if (ctx->ws[11].reserve((size_t)std::max<int64_t>(total, 1) * 16)) return -1;
        pf = ctx->ws[11].as<float4>();
    }
    if (total > 0) {
        const double invf = 1.0 / prm.focal;
        kt_begin(ctx, st, KT_PREP);
        prep_points_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(
            reinterpret_cast<const double2*>(io.dA), reinterpret_cast<const double2*>(io.dB), total,
            MODEL == RM_ESSENTIAL ? 1 : 0, invf, -prm.ppx * invf, -prm.ppy * invf, pf, pd);
        kt_end(ctx, st);
        MCV_LAUNCH_CHECK();
    }

    // pair chunks bound the per-hypothesis workspaces (~1 GB of models at most)
    const size_t perPair = (size_t)maxIters * MAXM * 72;
    int chunk = (int)std::max<size_t>(1, std::min<size_t>((size_t)nPairs, ((size_t)1 << 30) / std::max<size_t>(perPair, 1)));
    chunk = std::min(chunk, 65535);      // the pair index is gridDim.y of several stage kernels
    if (const char* e = getenv("MINICV_RANSAC_CHUNK_PAIRS")) {           // tests: force chunk seams on small batches
        const int c = atoi(e);
        if (c > 0) chunk = std::min(chunk, c);
    }
    const size_t slots = (size_t)maxIters * MAXM;
    if (ctx->ws[12].reserve((size_t)chunk * maxIters * MP * 4) ||       // samples
        ctx->ws[13].reserve((size_t)chunk * 8 + 64) ||                  // nIter, flags
        ctx->ws[14].reserve((size_t)chunk * slots * 72) ||              // models
        ctx->ws[15].reserve((size_t)chunk * maxIters) ||                // nModels
        ctx->ws[16].reserve((size_t)chunk * (maxIters + 1) * 4) ||      // iterStart
        ctx->ws[17].reserve((size_t)chunk * slots * 4) ||               // slotList
        ctx->ws[18].reserve((size_t)chunk * slots * 4) ||               // counts
        ctx->ws[19].reserve((size_t)chunk * 4 + 64) ||                  // nTotal
        ctx->ws[20].reserve((size_t)chunk * 16))                        // best
        return -1;
    int* dSamples = ctx->ws[12].as<int>();
    int* dNIter = ctx->ws[13].as<int>();
    int* dFlags = dNIter + chunk;
    double* dModels = ctx->ws[14].as<double>();
    uint8_t* dNModels = ctx->ws[15].as<uint8_t>();
    int* dIterStart = ctx->ws[16].as<int>();
    int* dSlotList = ctx->ws[17].as<int>();
    int* dCounts = ctx->ws[18].as<int>();
    int* dNTotal = ctx->ws[19].as<int>();
    int* dBest = ctx->ws[20].as<int>();
    double* dE5Scr = nullptr;
    int* dE5Roots = nullptr;
    if (MODEL == RM_ESSENTIAL && !io.dInjModels) {
        if (ctx->ws[68].reserve((size_t)chunk * maxIters * E5_SCR * 8) || ctx->ws[69].reserve((size_t)chunk * maxIters * 4)) return -1;
        dE5Scr = ctx->ws[68].as<double>();
        dE5Roots = ctx->ws[69].as<int>();
    }

    for (int p0 = 0; p0 < nPairs; p0 += chunk) {
        const int np = std::min(chunk, nPairs - p0);
        const int64_t* off = dOff + p0;
        int64_t maxN = 0;
        for (int p = p0; p < p0 + np; ++p) maxN = std::max(maxN, io.hOff[p + 1] - io.hOff[p]);

        // ---- sample
        kt_begin(ctx, st, KT_SAMPLE);
        if (io.dInjModels) {
            injected_setup_kernel<<<(np + 255) / 256, 256, 0, st>>>(off, np, maxIters, MP, dNIter, dFlags);
            MCV_LAUNCH_CHECK();
        } else if (prm.sampleMode == MCV_SAMPLE_INJECTED) {
            if (!io.dInjSamples) { set_error("MCV_SAMPLE_INJECTED without sample indices"); return -1; }
            MCV_CUDA_TRY(cudaMemcpyAsync(dSamples, io.dInjSamples + (size_t)p0 * maxIters * MP,
                                         (size_t)np * maxIters * MP * 4, cudaMemcpyDeviceToDevice, st));
            injected_setup_kernel<<<(np + 255) / 256, 256, 0, st>>>(off, np, maxIters, MP, dNIter, dFlags);
            MCV_LAUNCH_CHECK();
        } else if (prm.sampleMode == MCV_SAMPLE_COUNTER) {
            dim3 g((maxIters + SC_POOL - 1) / SC_POOL, np);
            sample_counter_kernel<MODEL><<<g, 128, 0, st>>>(pf, off, np, maxIters, prm.seed, prm.pairIndexBase + p0,
                                                            dSamples, dNIter, dFlags);
            MCV_LAUNCH_CHECK();
        } else {
            // OpenCV's stream, replayed in parallel from the precomputed raw draws ...
            if (ctx->ws[28].reserve((size_t)MWC_TABLE * 4) || ctx->ws[29].reserve((size_t)np)) return -1;
            uint32_t* dRaw = ctx->ws[28].as<uint32_t>();
            uint8_t* dNeedSeq = ctx->ws[29].as<uint8_t>();
            if (!ctx->mwc_ready) {
                mwc_table_kernel<<<1, 32, 0, st>>>(dRaw);
                MCV_LAUNCH_CHECK();
                ctx->mwc_ready = true;
            }
            {
                const int smemPts = MODEL == RM_ESSENTIAL ? 0 : (int)std::min<int64_t>(maxN, 8192);   // 128 KB of float4 points
                auto kern = sample_mwc_par_kernel<MODEL>;
                const size_t smemBytes = (size_t)smemPts * 16 + (MWC_BLOCK + MWC_MAXDRAW) * 4 + (size_t)MWC_BLOCK * MP * 2 +
                                         (MWC_BLOCK / MP + 8) * 2 + MWC_BLOCK + 64;
                // the attribute is shared by every host thread that launches this kernel: always the largest size
                const size_t smemMax = (size_t)8192 * 16 + (MWC_BLOCK + MWC_MAXDRAW) * 4 + (size_t)MWC_BLOCK * MP * 2 +
                                       (MWC_BLOCK / MP + 8) * 2 + MWC_BLOCK + 64;
                MCV_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smemMax));
                kern<<<np, 256, smemBytes, st>>>(pf, off, np, maxIters, smemPts, dRaw, dSamples, dNIter, dFlags, dNeedSeq);
                MCV_LAUNCH_CHECK();
            }
            // ... and sequentially for the pairs the table cannot serve
            const int smemPts = (int)std::min<int64_t>(maxN, 12288);       // up to 192 KB of float4 points
            auto kern = sample_mwc_kernel<MODEL>;
            const size_t smemBytes = MODEL == RM_ESSENTIAL ? 0 : (size_t)smemPts * 16;
            if (MODEL != RM_ESSENTIAL)
                MCV_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 12288 * 16));
            kern<<<np, 32, smemBytes, st>>>(pf, off, np, maxIters, smemPts, dSamples, dNIter, dFlags, dNeedSeq);
            MCV_LAUNCH_CHECK();
        }
        kt_end(ctx, st);
        // ---- solve
        kt_begin(ctx, st, KT_SOLVE);
        if (io.dInjModels) {
            const int64_t nh = (int64_t)np * maxIters;
            inject_models_kernel<<<(unsigned)((nh + 255) / 256), 256, 0, st>>>(
                io.dInjModels + (size_t)p0 * maxIters * 9, nh, MAXM, dModels, dNModels);
            MCV_LAUNCH_CHECK();
        } else if constexpr (MODEL == RM_ESSENTIAL) {
            const int64_t nh = (int64_t)np * maxIters;
            dim3 g((maxIters + 63) / 64, np);
            solve_e_null_kernel<<<g, 64, 0, st>>>(pd, off, maxIters, dSamples, dNIter, nullptr, nullptr, 0, dE5Scr, dE5Roots);
            MCV_LAUNCH_CHECK();
            solve_e_elim_kernel<<<(unsigned)((nh + E5_HPB - 1) / E5_HPB), 128, 0, st>>>(nh, dE5Scr, dE5Roots);
            MCV_LAUNCH_CHECK();
            solve_e_roots_kernel<<<(unsigned)((nh + 127) / 128), 128, 0, st>>>(nh, dE5Scr, dE5Roots);
            MCV_LAUNCH_CHECK();
            solve_e_models_kernel<<<(unsigned)((nh * 16 + 127) / 128), 128, 0, st>>>(nh, dE5Scr, dE5Roots, dModels, dNModels);
            MCV_LAUNCH_CHECK();
        } else {
            dim3 g((maxIters + 63) / 64, np);
            solve_kernel<MODEL><<<g, 64, 0, st>>>(pf, pd, off, maxIters, dSamples, dNIter, dModels, dNModels);
            MCV_LAUNCH_CHECK();
        }
        kt_end(ctx, st);
        // ---- compact
        kt_begin(ctx, st, KT_COMPACT);
        compact_kernel<<<np, 256, 0, st>>>(dNModels, dNIter, maxIters, MAXM, dIterStart, dSlotList, dNTotal);
        kt_end(ctx, st);
        MCV_LAUNCH_CHECK();
        if (prm.stopMode == MCV_STOP_LMEDS) {
            // ---- LMedS: median score per model, first minimum wins, sigma-rule mask (fundamental matrices only)
            if constexpr (MODEL == RM_FUNDAMENTAL) {
                if (maxN > LMEDS_MAXN) { set_error("LMedS supports at most %d correspondences per pair", LMEDS_MAXN); return -1; }
                if (ctx->ws[74].reserve((size_t)chunk * 4 + 64)) return -1;
                float* dMedian = ctx->ws[74].as<float>();
                lmeds_select_kernel<<<np, 128, 0, st>>>(pf, off, dModels, dSlotList, dNTotal, dNIter, dFlags, (int)slots,
                                                        dBest, dMedian);
                MCV_LAUNCH_CHECK();
                finalize_lmeds_kernel<<<np, 64, 0, st>>>(pf, off, dModels, (int)slots, MAXM, MP, dBest, dFlags, dMedian,
                                                         io.dModels + (size_t)p0 * 9, io.dMasks,
                                                         io.dInlierCounts ? io.dInlierCounts + p0 : nullptr, io.dOk + p0,
                                                         io.dItersRun ? io.dItersRun + p0 : nullptr,
                                                         io.dBest ? io.dBest + 2 * (size_t)p0 : nullptr);
                MCV_LAUNCH_CHECK();
            }
        } else {
        // ---- score: models [dBegin[p], dEnd[p]) of every pair's compacted list; slotsEff bounds dEnd (grid size)
        // engine: tensor cores (1) / fp32-filter kernel (0) / both launched behind the device-side gate (2) -- score_tc_mode
        const int tcMode = score_tc_mode(MODEL, maxN, (int)slots, np, ctx->sm_count);
        // models [dBegin, dEnd) of every pair; engineGate: the pilot's gate when the device picks the engine (else null)
        auto score_range = [&](const int* dBegin, const int* dEnd, const int* engineGate, bool zeroed, bool useTc,
                               bool useClassic) -> int {
            if (useTc) {
                const int rc = score_tc_run(ctx, st, MODEL, pf, pd, off, io.hOff + p0, p0, np, dModels, dSlotList, dEnd,
                                            (int)slots, thr2, dCounts, engineGate, dBegin, false);
                if (rc) return rc;
            }
            if (useClassic) {
                const int tiles = (int)((slots + 127) / 128);
                int zs = 1;
                const int64_t blocksXY = (int64_t)tiles * np;
                const int64_t fill = 4ll * ctx->sm_count;
                const int TPmin = (MODEL == RM_ESSENTIAL) ? 512 : 1024;
                if (blocksXY < fill) zs = (int)std::min<int64_t>((fill + blocksXY - 1) / blocksXY, std::max<int64_t>(1, (maxN + TPmin - 1) / TPmin));
                int chunkPts = (int)((std::max<int64_t>(maxN, 1) + zs - 1) / zs);
                zs = (int)((std::max<int64_t>(maxN, 1) + chunkPts - 1) / chunkPts);
                if (useTc) { zs = 1; chunkPts = (int)std::max<int64_t>(maxN, 1); }    // a chunk big enough for the tensor engine fills the machine
                if (zs > 1 && !zeroed) MCV_CUDA_TRY(cudaMemsetAsync(dCounts, 0, (size_t)np * slots * 4, st));
                dim3 g(tiles, np, zs);
                score_kernel<MODEL><<<g, 128, 0, st>>>(pf, pd, off, dModels, dSlotList, dEnd, (int)slots,
                                                       thr2, chunkPts, dCounts, engineGate, dBegin);
                MCV_LAUNCH_CHECK();
            }
            return 0;
        };
        {
            constexpr int K1 = 128;                       // iterations of phase 1
            const bool twoPhase = prm.stopMode == MCV_STOP_ADAPTIVE && !io.dDbgCounts && !io.dInjModels && maxIters >= 4 * K1;
            kt_begin(ctx, st, KT_SCORE);
            const int* gate = nullptr;                    // {models that fit, models looked at} of the pilot
            if (tcMode == 2 || (tcMode == 1 && twoPhase)) {
                gate = score_tc_pilot(ctx, st, MODEL, pf, pd, off, np, dModels, dSlotList, dNTotal, (int)slots, thr2);
                if (!gate) return -1;
            }
            const int* engineGate = tcMode == 2 ? gate : nullptr;
            if (!twoPhase) {
                if (score_range(nullptr, dNTotal, engineGate, false, tcMode != 0, tcMode != 1)) return -1;
            } else {
                if (ctx->ws[86].reserve((size_t)chunk * 8 + 64)) return -1;
                int* dPhaseEnd = ctx->ws[86].as<int>();
                int* dPhaseBegin = dPhaseEnd + chunk;
                // counts zeroed once: small calls split a pair's correspondences over blocks that add their partial counts,
                // and phase 2 must not wipe what phase 1 counted
                MCV_CUDA_TRY(cudaMemsetAsync(dCounts, 0, (size_t)np * slots * 4, st));
                phase_limit_kernel<<<(np + 127) / 128, 128, 0, st>>>(np, maxIters, K1, dNIter, dIterStart, dNTotal, gate, dPhaseEnd);
                MCV_LAUNCH_CHECK();
                if (score_range(nullptr, dPhaseEnd, engineGate, true, tcMode != 0, tcMode != 1)) return -1;
                horizon_kernel<<<(np + 127) / 128, 128, 0, st>>>(off, np, maxIters, MAXM, MP, K1, prm.confidence, dNIter, dFlags,
                                                                 dIterStart, dSlotList, dCounts, gate, dPhaseBegin, dPhaseEnd);
                MCV_LAUNCH_CHECK();
                // phase 2 is a few iterations per pair when there is any at all: always the fp32-filter kernel (thousands of
                // tensor-engine CTAs returning at once cost 0.45 ms per chunk when tried)
                if (score_range(dPhaseBegin, dPhaseEnd, nullptr, true, false, true)) return -1;
            }
            kt_end(ctx, st);
        }
        // ---- select
        kt_begin(ctx, st, KT_SELECT);
        select_kernel<<<np, 128, 0, st>>>(off, np, maxIters, MAXM, MP,
                                                    prm.stopMode == MCV_STOP_ADAPTIVE ? 1 : 0,
                                                    prm.confidence, dNIter, dFlags, dIterStart, dSlotList,
                                                    dCounts, dBest);
        kt_end(ctx, st);
        MCV_LAUNCH_CHECK();
        // ---- finalize
        double* oM = io.dModels + (size_t)p0 * 9;
        uint8_t* oOk = io.dOk + p0;
        int* oCnt = io.dInlierCounts ? io.dInlierCounts + p0 : nullptr;
        int* oIt = io.dItersRun ? io.dItersRun + p0 : nullptr;
        int* oBest = io.dBest ? io.dBest + 2 * (size_t)p0 : nullptr;
        kt_begin(ctx, st, KT_FINALIZE);
        if (MODEL == RM_HOMOGRAPHY && prm.refine) {
            // small blocks once there are enough pairs to fill the machine with them: 7 x 64 threads per SM keep every
            // pair of a 1024-pair batch resident, so the serial warp-0 phases of all refits overlap (one wave instead of three)
            const int fb = np >= 4 * ctx->sm_count ? 64 : FB;
            finalize_homography_kernel<<<np, fb, 0, st>>>(pf, off, dModels, (int)slots, dBest, dFlags, thr2,
                                                          oM, io.dMasks, oCnt, oOk, oIt, oBest);
        } else {
            finalize_plain_kernel<MODEL><<<np, FB, 0, st>>>(pf, pd, off, dModels, (int)slots, dBest, dFlags,
                                                            dNTotal, thr2, oM, io.dMasks, oCnt, oOk, oIt, oBest);
        }
        kt_end(ctx, st);
        MCV_LAUNCH_CHECK();
        }
        // ---- debug taps
        if (io.dDbgCounts && prm.stopMode != MCV_STOP_LMEDS) {
            dim3 g((maxIters + 255) / 256, np);
            dbg_counts_kernel<<<g, 256, 0, st>>>(dIterStart, dCounts, maxIters, MAXM,
                                                 io.dDbgCounts + (size_t)p0 * slots);
            MCV_LAUNCH_CHECK();
        }
        if (io.dDbgModels)
            MCV_CUDA_TRY(cudaMemcpyAsync(io.dDbgModels + (size_t)p0 * slots * 9, dModels,
                                         (size_t)np * slots * 72, cudaMemcpyDeviceToDevice, st));
        if (io.dDbgSamples && !io.dInjModels)
            MCV_CUDA_TRY(cudaMemcpyAsync(io.dDbgSamples + (size_t)p0 * maxIters * MP, dSamples,
                                         (size_t)np * maxIters * MP * 4, cudaMemcpyDeviceToDevice, st));
    }
    return 0;
}

}  // namespace

// The model-independent stages, shared with the PnP engine (pnp.cu).
int ransac_compact(cudaStream_t st, int np, const uint8_t* dNModels, const int* dNIter, int maxIters, int maxm,
                   int* dIterStart, int* dSlotList, int* dNTotal) {
    compact_kernel<<<np, 256, 0, st>>>(dNModels, dNIter, maxIters, maxm, dIterStart, dSlotList, dNTotal);
    MCV_LAUNCH_CHECK();
    return 0;
}

int ransac_select(cudaStream_t st, const int64_t* dOff, int np, int maxIters, int maxm, int mp, bool adaptive,
                  double confidence, const int* dNIter, const int* dFlags, const int* dIterStart,
                  const int* dSlotList, const int* dCounts, int* dBest) {
    select_kernel<<<np, 128, 0, st>>>(dOff, np, maxIters, maxm, mp, adaptive ? 1 : 0, confidence, dNIter,
                                                dFlags, dIterStart, dSlotList, dCounts, dBest);
    MCV_LAUNCH_CHECK();
    return 0;
}

int five_point_device(Context* ctx, cudaStream_t st, const double* dA, const double* dB, int nSets, double* dEs,
                      int* dCounts) {
    if (nSets <= 0) return 0;
#ifdef MCV_DEBUG_COMPOSED_E5
    five_point_composed_kernel<<<(nSets + 63) / 64, 64, 0, st>>>(dA, dB, nSets, dEs, dCounts);
    MCV_LAUNCH_CHECK();
    return 0;
#endif
    const size_t n = (size_t)nSets;
    if (ctx->ws[68].reserve(n * E5_SCR * 8) || ctx->ws[69].reserve(n * 4 + n + 64)) return -1;
    double* scr = ctx->ws[68].as<double>();
    int* nRoots = ctx->ws[69].as<int>();
    uint8_t* nModels = reinterpret_cast<uint8_t*>(nRoots + n);
    solve_e_null_kernel<<<(nSets + 63) / 64, 64, 0, st>>>(nullptr, nullptr, 0, nullptr, nullptr, dA, dB, nSets, scr, nRoots);
    MCV_LAUNCH_CHECK();
    solve_e_elim_kernel<<<(unsigned)((n + E5_HPB - 1) / E5_HPB), 128, 0, st>>>((int64_t)n, scr, nRoots);
    MCV_LAUNCH_CHECK();
    solve_e_roots_kernel<<<(unsigned)((n + 127) / 128), 128, 0, st>>>((int64_t)n, scr, nRoots);
    MCV_LAUNCH_CHECK();
    solve_e_models_kernel<<<(unsigned)((n * 16 + 127) / 128), 128, 0, st>>>((int64_t)n, scr, nRoots, dEs, nModels);
    MCV_LAUNCH_CHECK();
    widen_counts_kernel<<<(nSets + 255) / 256, 256, 0, st>>>(nModels, nSets, dCounts);
    MCV_LAUNCH_CHECK();
    return 0;
}

int ransac_batch_device(Context* ctx, cudaStream_t stream, const RansacParams& prm, const RansacIO& io) {
    if (io.nPairs <= 0) return 0;
    if (!io.hOff || io.hOff[0] != 0) { set_error("ransac: offsets must start at 0"); return -1; }
    if (!io.dModels || !io.dMasks || !io.dOk) { set_error("ransac: null outputs"); return -1; }
    if (prm.stopMode < 0 || prm.stopMode > MCV_STOP_LMEDS) { set_error("ransac: unknown stopMode %d", prm.stopMode); return -1; }
    if (!(prm.confidence > 0 && prm.confidence < 1) && prm.stopMode == MCV_STOP_ADAPTIVE) {
        set_error("ransac: confidence must be in (0,1) (OpenCV asserts the same)");
        return -1;
    }
    if (prm.maxIters > (1 << 20)) { set_error("ransac: maxIters too large"); return -1; }
    if (prm.stopMode == MCV_STOP_LMEDS) {
        if (prm.model != RM_FUNDAMENTAL) { set_error("ransac: LMedS is implemented for fundamental matrices only"); return -1; }
        if (!(prm.confidence > 0 && prm.confidence < 1)) { set_error("ransac: confidence must be in (0,1)"); return -1; }
        RansacParams q = prm;
        // niters = RANSACUpdateNumIters(confidence, outlierRatio 0.45, modelPoints, maxIters), fixed before the loop
        q.maxIters = cvm::update_num_iters(prm.confidence, 0.45, 7, std::max(prm.maxIters, 1));
        return run_model<RM_FUNDAMENTAL>(ctx, stream, q, io);
    }
    switch (prm.model) {
        case RM_HOMOGRAPHY: return run_model<RM_HOMOGRAPHY>(ctx, stream, prm, io);
        case RM_FUNDAMENTAL: return run_model<RM_FUNDAMENTAL>(ctx, stream, prm, io);
        case RM_ESSENTIAL: return run_model<RM_ESSENTIAL>(ctx, stream, prm, io);
    }
    set_error("ransac: unknown model %d", prm.model);
    return -1;
}

}  // namespace mcv
