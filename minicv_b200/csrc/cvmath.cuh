// cvmath.cuh -- register/local-memory resident numerical building blocks of the two-view
// estimators, written as __host__ __device__ templates so the SAME source is unit-tested on
// the host (tests/hostmath, a test-only library) and runs inside the CUDA kernels.
//
// What each block mirrors (OpenCV 4.13.0, the un-vendored dependency the reference shim
// forwards to at src/MiniCVNative/MiniCVNative.cpp:177,186,204,212 and re-uses inside its
// vendored solver at src/MiniCVNative/fivepoint.cpp:252,297,323):
//   * Mwc                 cv::RNG (multiply-with-carry), seeded (uint64)-1 per RANSAC call
//   * jacobi_svd_rows     JacobiSVDImpl_ (one-sided Jacobi + pseudo-random completion); the
//                         basis/sign conventions of every null space used below come from it
//   * jacobi_eigen_sym    symmetric eigen-decomposition (cyclic Jacobi)
//   * scoring functions   bit-exact restatements (explicit _rn intrinsics: no FMA contraction)
//   * minimal solvers     4-point homography, 7-point fundamental, 5-point essential
#pragma once

#include <stdint.h>
#include <math.h>
#include <float.h>

#if defined(__CUDACC__)
#define MCV_HD __host__ __device__ __forceinline__
#define MCV_HDN __host__ __device__
#ifdef __CUDA_ARCH__
#define MCV_UNROLL _Pragma("unroll")
#else
#define MCV_UNROLL
#endif
#else
#define MCV_HD inline
#define MCV_HDN
#define MCV_UNROLL
#endif

namespace mcv {
namespace cvm {

// ---- rounding-exact arithmetic (no contraction into FMA on either side) -------------------
#if defined(__CUDA_ARCH__)
MCV_HD float  fmul(float a, float b)   { return __fmul_rn(a, b); }
MCV_HD float  fadd(float a, float b)   { return __fadd_rn(a, b); }
MCV_HD float  fsub(float a, float b)   { return __fsub_rn(a, b); }
MCV_HD float  fdiv(float a, float b)   { return __fdiv_rn(a, b); }
MCV_HD double dmul(double a, double b) { return __dmul_rn(a, b); }
MCV_HD double dadd(double a, double b) { return __dadd_rn(a, b); }
MCV_HD double dsub(double a, double b) { return __dsub_rn(a, b); }
MCV_HD double ddiv(double a, double b) { return __ddiv_rn(a, b); }
#else
// host build: compiled without -mfma / with -ffp-contract=off, plain operators are exact
MCV_HD float  fmul(float a, float b)   { volatile float r = a * b; return r; }
MCV_HD float  fadd(float a, float b)   { volatile float r = a + b; return r; }
MCV_HD float  fsub(float a, float b)   { volatile float r = a - b; return r; }
MCV_HD float  fdiv(float a, float b)   { volatile float r = a / b; return r; }
MCV_HD double dmul(double a, double b) { volatile double r = a * b; return r; }
MCV_HD double dadd(double a, double b) { volatile double r = a + b; return r; }
MCV_HD double dsub(double a, double b) { volatile double r = a - b; return r; }
MCV_HD double ddiv(double a, double b) { volatile double r = a / b; return r; }
#endif

// ---- cv::RNG --------------------------------------------------------------------------------
struct Mwc {
    uint64_t state;
    MCV_HD Mwc() : state(0xFFFFFFFFFFFFFFFFull) {}
    MCV_HD explicit Mwc(uint64_t s) : state(s) {}
    MCV_HD uint32_t next() {
        state = (uint64_t)(uint32_t)state * 4164903690ull + (state >> 32);
        return (uint32_t)state;
    }
    MCV_HD int uniform(int n) { return (int)(next() % (uint32_t)n); }
};

// counter-based generator for the native (non-OpenCV) sampling mode: splitmix64 of the key
MCV_HD uint64_t mix64(uint64_t x) {
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}

// RANSACUpdateNumIters (calib3d/src/ptsetreg.cpp)
MCV_HDN inline int update_num_iters(double p, double ep, int modelPoints, int maxIters) {
    p = fmax(p, 0.); p = fmin(p, 1.);
    ep = fmax(ep, 0.); ep = fmin(ep, 1.);
    double num = fmax(1. - p, DBL_MIN);
    double denom = 1. - pow(1. - ep, (double)modelPoints);
    if (denom < DBL_MIN) return 0;
    num = log(num);
    denom = log(denom);
    if (denom >= 0 || -num >= maxIters * (-denom)) return maxIters;
    return (int)rint(num / denom);   // cvRound: round half to even
}

// ---- one-sided Jacobi SVD, OpenCV's JacobiSVDImpl_ -----------------------------------------
// At: N1 rows of length M (row-major, leading dimension M); the first N rows hold the vectors
// to orthogonalise.  On return rows 0..N1-1 are unit "left" vectors (rows >= N fabricated by
// OpenCV's fixed pseudo-random completion), W[N] the singular values (descending), Vt (N x N)
// the accumulated rotations.  WANT_V = false skips Vt.
// S = element stride of At (1 = a plain array; blockDim for a thread-interleaved shared-memory matrix, see solve_kernel).
// EXT: the fabricated rows N .. N1-1 live in Ext ((N1 - N) x M, stride 1) instead of At, so that a strided (shared-memory)
// At only has to hold the N data rows.
template <int M, int N, int N1, bool WANT_V, int S = 1, bool EXT = false>
MCV_HDN inline void jacobi_svd_rows(double* At, double* W, double* Vt, double* Ext = nullptr) {
    const double eps = DBL_EPSILON * 10, minval = DBL_MIN;
    for (int i = 0; i < N; ++i) {
        double sd = 0;
        for (int k = 0; k < M; ++k) sd += At[(i * M + k) * S] * At[(i * M + k) * S];
        W[i] = sd;
        if (WANT_V) {
            for (int k = 0; k < N; ++k) Vt[(i * N + k) * S] = 0;
            Vt[(i * N + i) * S] = 1;
        }
    }
    const int max_iter = M > 30 ? M : 30;
    for (int iter = 0; iter < max_iter; ++iter) {
        bool changed = false;
        for (int i = 0; i < N - 1; ++i)
            for (int j = i + 1; j < N; ++j) {
                double* Ai = At + i * M * S;
                double* Aj = At + j * M * S;
                double a = W[i], p = 0, b = W[j];
                for (int k = 0; k < M; ++k) p += Ai[k * S] * Aj[k * S];
                if (fabs(p) <= eps * sqrt(a * b)) continue;
                p *= 2;
                double beta = a - b, gamma = hypot(p, beta), c, s;
                if (beta < 0) {
                    double delta = (gamma - beta) * 0.5;
                    s = sqrt(delta / gamma);
                    c = p / (gamma * s * 2);
                } else {
                    c = sqrt((gamma + beta) / (gamma * 2));
                    s = p / (gamma * c * 2);
                }
                a = b = 0;
                for (int k = 0; k < M; ++k) {
                    double t0 = c * Ai[k * S] + s * Aj[k * S];
                    double t1 = -s * Ai[k * S] + c * Aj[k * S];
                    Ai[k * S] = t0; Aj[k * S] = t1;
                    a += t0 * t0; b += t1 * t1;
                }
                W[i] = a; W[j] = b;
                changed = true;
                if (WANT_V) {
                    double* Vi = Vt + i * N * S;
                    double* Vj = Vt + j * N * S;
                    for (int k = 0; k < N; ++k) {
                        double t0 = c * Vi[k * S] + s * Vj[k * S];
                        double t1 = -s * Vi[k * S] + c * Vj[k * S];
                        Vi[k * S] = t0; Vj[k * S] = t1;
                    }
                }
            }
        if (!changed) break;
    }
    for (int i = 0; i < N; ++i) {
        double sd = 0;
        for (int k = 0; k < M; ++k) sd += At[(i * M + k) * S] * At[(i * M + k) * S];
        W[i] = sqrt(sd);
    }
    for (int i = 0; i < N - 1; ++i) {
        int j = i;
        for (int k = i + 1; k < N; ++k)
            if (W[j] < W[k]) j = k;
        if (i != j) {
            double t = W[i]; W[i] = W[j]; W[j] = t;
            for (int k = 0; k < M; ++k) { t = At[(i * M + k) * S]; At[(i * M + k) * S] = At[(j * M + k) * S]; At[(j * M + k) * S] = t; }
            if (WANT_V)
                for (int k = 0; k < N; ++k) { t = Vt[(i * N + k) * S]; Vt[(i * N + k) * S] = Vt[(j * N + k) * S]; Vt[(j * N + k) * S] = t; }
        }
    }
    Mwc rng(0x12345678ull);
    for (int i = 0; i < N1; ++i) {
        double* Ri = (EXT && i >= N) ? Ext + (i - N) * M : At + i * M * S;
        const int si = (EXT && i >= N) ? 1 : S;
        double sd = i < N ? W[i] : 0;
        for (int ii = 0; ii < 100 && sd <= minval; ++ii) {
            const double val0 = 1. / M;
            for (int k = 0; k < M; ++k) Ri[k * si] = (rng.next() & 256) != 0 ? val0 : -val0;
            for (int it = 0; it < 2; ++it)
                for (int j = 0; j < i; ++j) {
                    const double* Rj = (EXT && j >= N) ? Ext + (j - N) * M : At + j * M * S;
                    const int sj = (EXT && j >= N) ? 1 : S;
                    sd = 0;
                    for (int k = 0; k < M; ++k) sd += Ri[k * si] * Rj[k * sj];
                    double asum = 0;
                    for (int k = 0; k < M; ++k) {
                        double t = Ri[k * si] - sd * Rj[k * sj];
                        Ri[k * si] = t;
                        asum += fabs(t);
                    }
                    asum = asum > eps * 100 ? 1 / asum : 0;
                    for (int k = 0; k < M; ++k) Ri[k * si] *= asum;
                }
            sd = 0;
            for (int k = 0; k < M; ++k) sd += Ri[k * si] * Ri[k * si];
            sd = sqrt(sd);
        }
        const double s = sd > minval ? 1 / sd : 0.;
        for (int k = 0; k < M; ++k) Ri[k * si] *= s;
    }
}

// ---- null-space basis in OpenCV's convention, without the SVD ---------------------------------------
// cv::SVD::compute(A, w, u, vt, FULL_UV) of an N x M matrix with N < M returns, as rows N .. M-1 of vt, the vectors
// JacobiSVDImpl_ FABRICATES to complete the basis: fixed pseudo-random +-1/M vectors (RNG(0x12345678)), each
// orthogonalised against all earlier rows (two Gram-Schmidt passes) and normalised.  Those rows therefore depend only
// on the SPAN of the data rows, not on which orthonormal basis of it the Jacobi sweeps produce -- so the run7Point /
// five-point null spaces (signs included) can be had from a modified Gram-Schmidt of the data rows (with
// re-orthogonalisation) followed by the same completion, at a fraction of the Jacobi SVD's arithmetic and without its
// chains of dependent sqrt / hypot / divide.  Checked against cv2.SVDecomp(FULL_UV): 3000 random 7 x 9 (pixel
// coordinates) and 5 x 9 systems agree to 1e-15 median, 2e-13 worst (oracle/linalg_cv.py keeps the Jacobi restatement).
// At: N rows of length M with element stride S (overwritten by an orthonormal basis of their span); Ext: the N1 - N
// completed rows, stride 1.
template <int M, int N, int N1, int S = 1>
MCV_HDN inline void nullspace_rows(double* At, double* Ext) {
    const double eps = DBL_EPSILON * 10, minval = DBL_MIN;
    for (int i = 0; i < N; ++i) {
        double v[M];
MCV_UNROLL
        for (int k = 0; k < M; ++k) v[k] = At[(i * M + k) * S];
        for (int pass = 0; pass < 2; ++pass)
            for (int j = 0; j < i; ++j) {
                double q[M], d = 0;
MCV_UNROLL
                for (int k = 0; k < M; ++k) { q[k] = At[(j * M + k) * S]; d += v[k] * q[k]; }
MCV_UNROLL
                for (int k = 0; k < M; ++k) v[k] -= d * q[k];
            }
        double sd = 0;
MCV_UNROLL
        for (int k = 0; k < M; ++k) sd += v[k] * v[k];
        sd = sqrt(sd);
        const double s = sd > minval ? 1 / sd : 0.;
MCV_UNROLL
        for (int k = 0; k < M; ++k) At[(i * M + k) * S] = v[k] * s;
    }
    Mwc rng(0x12345678ull);
    for (int i = N; i < N1; ++i) {
        double v[M], sd = 0;
        for (int ii = 0; ii < 100 && sd <= minval; ++ii) {
            const double val0 = 1. / M;
MCV_UNROLL
            for (int k = 0; k < M; ++k) v[k] = (rng.next() & 256) != 0 ? val0 : -val0;
            for (int pass = 0; pass < 2; ++pass) {
                for (int j = 0; j < i; ++j) {
                    double q[M], d = 0;
MCV_UNROLL
                    for (int k = 0; k < M; ++k) { q[k] = j < N ? At[(j * M + k) * S] : Ext[(j - N) * M + k]; d += v[k] * q[k]; }
MCV_UNROLL
                    for (int k = 0; k < M; ++k) v[k] -= d * q[k];
                }
                double asum = 0;
MCV_UNROLL
                for (int k = 0; k < M; ++k) asum += fabs(v[k]);
                asum = asum > eps * 100 ? 1 / asum : 0;
MCV_UNROLL
                for (int k = 0; k < M; ++k) v[k] *= asum;
            }
            sd = 0;
MCV_UNROLL
            for (int k = 0; k < M; ++k) sd += v[k] * v[k];
            sd = sqrt(sd);
        }
        const double s = sd > minval ? 1 / sd : 0.;
MCV_UNROLL
        for (int k = 0; k < M; ++k) Ext[(i - N) * M + k] = v[k] * s;
    }
}

// cv::SVD::compute(A, w, u, vt) for a square K x K matrix (row-major).  u and vt row-major.
template <int K>
MCV_HDN inline void svd_square(const double* A, double* w, double* u, double* vt) {
    double At[K * K];
    for (int i = 0; i < K; ++i)
        for (int j = 0; j < K; ++j) At[i * K + j] = A[j * K + i];   // rows of At = columns of A
    jacobi_svd_rows<K, K, K, true>(At, w, vt);
    for (int i = 0; i < K; ++i)
        for (int j = 0; j < K; ++j) u[j * K + i] = At[i * K + j];   // u = At^T
}

// symmetric eigen-decomposition by cyclic Jacobi.  A (N x N, row-major, destroyed) -> eigen
// values w (unsorted) and eigenvectors as ROWS of V.
template <int N>
MCV_HDN inline void jacobi_eigen_sym(double* A, double* w, double* V) {
    for (int i = 0; i < N; ++i) {
        for (int j = 0; j < N; ++j) V[i * N + j] = 0;
        V[i * N + i] = 1;
    }
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
                const double c = 1 / sqrt(t * t + 1), s = t * c;
                for (int k = 0; k < N; ++k) {
                    const double akp = A[k * N + p], akq = A[k * N + q];
                    A[k * N + p] = c * akp - s * akq;
                    A[k * N + q] = s * akp + c * akq;
                }
                for (int k = 0; k < N; ++k) {
                    const double apk = A[p * N + k], aqk = A[q * N + k];
                    A[p * N + k] = c * apk - s * aqk;
                    A[q * N + k] = s * apk + c * aqk;
                }
                for (int k = 0; k < N; ++k) {
                    const double vp = V[p * N + k], vq = V[q * N + k];
                    V[p * N + k] = c * vp - s * vq;
                    V[q * N + k] = s * vp + c * vq;
                }
            }
    }
    for (int i = 0; i < N; ++i) w[i] = A[i * N + i];
}

// cv::solve(A, b, x, DECOMP_EIG) for symmetric A: x = sum_i (v_i . b / w_i) v_i over
// eigenvalues above eps*2*sum|w| (SVBkSb's threshold).
template <int N>
MCV_HDN inline void solve_sym_eig(const double* A, const double* b, double* x) {
    double M[N * N], w[N], V[N * N];
    for (int i = 0; i < N * N; ++i) M[i] = A[i];
    jacobi_eigen_sym<N>(M, w, V);
    double sum = 0;
    for (int i = 0; i < N; ++i) sum += fabs(w[i]);
    const double thr = DBL_EPSILON * 2 * sum;
    for (int i = 0; i < N; ++i) x[i] = 0;
    for (int i = 0; i < N; ++i) {
        if (fabs(w[i]) <= thr) continue;
        double s = 0;
        for (int k = 0; k < N; ++k) s += V[i * N + k] * b[k];
        s /= w[i];
        for (int k = 0; k < N; ++k) x[k] += s * V[i * N + k];
    }
}

// Same solve, fast path: Cholesky when A is comfortably positive definite (J^T J + lambda diag with a healthy
// lambda), where it equals the eigen-solve to rounding; returns false (x untouched) when a pivot falls below
// 1e-12 of the largest diagonal entry -- the thresholded eigen-solve, whose pseudo-inverse semantics matter for
// ill-conditioned systems, takes over.
template <int N>
MCV_HDN inline bool solve_sym_chol(const double* A, const double* b, double* x) {
    double L[N * N];
    double dmax = 0;
    for (int i = 0; i < N; ++i) dmax = fmax(dmax, fabs(A[i * N + i]));
    bool ok = dmax > 0;
    for (int j = 0; j < N && ok; ++j) {
        double d = A[j * N + j];
        for (int k = 0; k < j; ++k) d -= L[j * N + k] * L[j * N + k];
        if (!(d > 1e-12 * dmax)) { ok = false; break; }
        const double ljj = sqrt(d);
        L[j * N + j] = ljj;
        for (int i = j + 1; i < N; ++i) {
            double s = A[i * N + j];
            for (int k = 0; k < j; ++k) s -= L[i * N + k] * L[j * N + k];
            L[i * N + j] = s / ljj;
        }
    }
    if (!ok) return false;
    double y[N];
    for (int i = 0; i < N; ++i) {
        double s = b[i];
        for (int k = 0; k < i; ++k) s -= L[i * N + k] * y[k];
        y[i] = s / L[i * N + i];
    }
    for (int i = N - 1; i >= 0; --i) {
        double s = y[i];
        for (int k = i + 1; k < N; ++k) s -= L[k * N + i] * x[k];
        x[i] = s / L[i * N + i];
    }
    return true;
}
template <int N>
MCV_HDN inline void solve_sym_chol_or_eig(const double* A, const double* b, double* x) {
    if (!solve_sym_chol<N>(A, b, x)) solve_sym_eig<N>(A, b, x);
}

template <int N>
MCV_HDN inline double max_diag_pinv_sym(const double* A) {
    double M[N * N], w[N], V[N * N];
    for (int i = 0; i < N * N; ++i) M[i] = A[i];
    jacobi_eigen_sym<N>(M, w, V);
    double sum = 0;
    for (int i = 0; i < N; ++i) sum += fabs(w[i]);
    const double thr = DBL_EPSILON * 2 * sum;
    double maxval = DBL_EPSILON;
    for (int k = 0; k < N; ++k) {
        double d = 0;
        for (int i = 0; i < N; ++i)
            if (fabs(w[i]) > thr) d += V[i * N + k] * V[i * N + k] / w[i];
        maxval = fmax(maxval, fabs(d));
    }
    return maxval;
}

// ---- scoring (bit-exact) --------------------------------------------------------------------
// Homography reprojection error, float32 (HomographyEstimatorCallback::computeError).
// Hf = first 8 entries of H cast to float (h33 taken as 1).
MCV_HD float h_error(const float* Hf, float Mx, float My, float mx, float my) {
    const float ww = fdiv(1.f, fadd(fadd(fmul(Hf[6], Mx), fmul(Hf[7], My)), 1.f));
    const float dx = fsub(fmul(fadd(fadd(fmul(Hf[0], Mx), fmul(Hf[1], My)), Hf[2]), ww), mx);
    const float dy = fsub(fmul(fadd(fadd(fmul(Hf[3], Mx), fmul(Hf[4], My)), Hf[5]), ww), my);
    return fadd(fmul(dx, dx), fmul(dy, dy));
}

// Fundamental matrix: max of the two squared point-line distances, float64 -> float32
// (FMEstimatorCallback::computeError).  Points are float32 promoted to double.
MCV_HD float f_error(const double* F, double x1, double y1, double x2, double y2) {
    double a = dadd(dadd(dmul(F[0], x1), dmul(F[1], y1)), F[2]);
    double b = dadd(dadd(dmul(F[3], x1), dmul(F[4], y1)), F[5]);
    double c = dadd(dadd(dmul(F[6], x1), dmul(F[7], y1)), F[8]);
    const double s2 = ddiv(1., dadd(dmul(a, a), dmul(b, b)));
    const double d2 = dadd(dadd(dmul(x2, a), dmul(y2, b)), c);
    a = dadd(dadd(dmul(F[0], x2), dmul(F[3], y2)), F[6]);
    b = dadd(dadd(dmul(F[1], x2), dmul(F[4], y2)), F[7]);
    c = dadd(dadd(dmul(F[2], x2), dmul(F[5], y2)), F[8]);
    const double s1 = ddiv(1., dadd(dmul(a, a), dmul(b, b)));
    const double d1 = dadd(dadd(dmul(x1, a), dmul(y1, b)), c);
    const double e1 = dmul(dmul(d1, d1), s1), e2 = dmul(dmul(d2, d2), s2);
    return (float)((e1 < e2) ? e2 : e1);   // std::max(e1, e2)
}

// Essential matrix Sampson error, float64 -> float32 (EMEstimatorCallback::computeError).
MCV_HD float e_error(const double* E, double x1, double y1, double x2, double y2) {
    const double Ex0 = dadd(dadd(dmul(E[0], x1), dmul(E[1], y1)), E[2]);
    const double Ex1 = dadd(dadd(dmul(E[3], x1), dmul(E[4], y1)), E[5]);
    const double Ex2 = dadd(dadd(dmul(E[6], x1), dmul(E[7], y1)), E[8]);
    const double Et0 = dadd(dadd(dmul(E[0], x2), dmul(E[3], y2)), E[6]);
    const double Et1 = dadd(dadd(dmul(E[1], x2), dmul(E[4], y2)), E[7]);
    const double xe = dadd(dadd(dmul(x2, Ex0), dmul(y2, Ex1)), Ex2);
    const double a = dmul(Ex0, Ex0), b = dmul(Ex1, Ex1), c = dmul(Et0, Et0), d = dmul(Et1, Et1);
    return (float)ddiv(dmul(xe, xe), dadd(dadd(dadd(a, b), c), d));
}

// Division-free inlier tests, decision-identical to `x_error(...) <= thr`.
// err = (float)(q * (1/n)) differs from the real quotient q/n by < 4 ulp(double) plus the final
// rounding to float (2^-24 relative), so q <= thr*n*(1 - 2e-7) proves "inlier" and
// q > thr*n*(1 + 2e-7) proves "outlier"; only the sliver in between (and non-finite or zero
// denominators) takes the exact path with the division.  thr is the float threshold (float)(t*t).
MCV_HD bool f_inlier(const double* F, double x1, double y1, double x2, double y2, float thr) {
    const double t = (double)thr, tl = t * (1.0 - 2e-7), th = t * (1.0 + 2e-7);
    double a = dadd(dadd(dmul(F[0], x1), dmul(F[1], y1)), F[2]);
    double b = dadd(dadd(dmul(F[3], x1), dmul(F[4], y1)), F[5]);
    double c = dadd(dadd(dmul(F[6], x1), dmul(F[7], y1)), F[8]);
    const double n2 = dadd(dmul(a, a), dmul(b, b));
    const double d2 = dadd(dadd(dmul(x2, a), dmul(y2, b)), c);
    a = dadd(dadd(dmul(F[0], x2), dmul(F[3], y2)), F[6]);
    b = dadd(dadd(dmul(F[1], x2), dmul(F[4], y2)), F[7]);
    c = dadd(dadd(dmul(F[2], x2), dmul(F[5], y2)), F[8]);
    const double n1 = dadd(dmul(a, a), dmul(b, b));
    const double d1 = dadd(dadd(dmul(x1, a), dmul(y1, b)), c);
    const double q1 = dmul(d1, d1), q2 = dmul(d2, d2);
    if (n1 > 0 && n2 > 0 && n1 <= DBL_MAX && n2 <= DBL_MAX) {   // finite: inf <= inf would pass where inf / inf = NaN fails
        if (q1 <= tl * n1 && q2 <= tl * n2) return true;
        if (q1 > th * n1 || q2 > th * n2) return false;
    }
    const double e1 = dmul(q1, ddiv(1., n1)), e2 = dmul(q2, ddiv(1., n2));
    return (float)((e1 < e2) ? e2 : e1) <= thr;
}

// ---- fp32 filter in front of f_inlier ------------------------------------------------------------
// Almost every (model, correspondence) pair of a RANSAC batch is an outlier by a wide margin, and the exact test
// above costs ~52 FP64-pipe instructions.  The filter evaluates ONE of the two point-line distances (the line F x1
// in image 2) in fp32 with FMAs, carries a bound of its own rounding error, and answers "certainly an outlier" only
// when the distance exceeds the threshold by more than every error that could have been made; the caller runs the
// exact test on everything else, so the result is f_inlier's, bit for bit.
//
// Error budget (u = 2^-24; points are exact fp32, F_i are converted with relative error u; the filter is switched
// off for a model with a non-finite entry or a non-zero entry below 1e-30, so conversions never underflow):
//   a~ = fma(F0,x1, fma(F1,y1,F2)):  |a~ - a| <= 4u A,  A = |F0||x1| + |F1||y1| + |F2|   (same for b, c)
//   d~ = fma(x2,a~, fma(y2,b~,c~)):  |d~ - d| <= 7u D,  D = |x2| A + |y2| B + C
// The bounds are taken per TILE of correspondences: with X1 >= |x1|, Y1 >= |y1|, X2 >= |x2|, Y2 >= |y2| over the tile,
// A <= Amax = |F0| X1 + |F1| Y1 + |F2| etc., so eA = 8u Amax >= |a~ - a|, eB, eD = 16u Dmax >= |d~ - d| (factor two
// of slack; Amax .. Dmax carry an absolute 1e-30 that dominates any underflow) are constants of (model, tile).
// To need neither |.| nor a square root per evaluation, both bounds are applied through (p + q)^2 <= (1 + mu) p^2 +
// (1 + 1/mu) q^2 with mu = 2^-12:
//   a^2 + b^2 <= (|a~| + eA)^2 + (|b~| + eB)^2 <= (1 + mu)(a~^2 + b~^2) + K,   K  = (1 + 1/mu)(eA^2 + eB^2) + 1e-30
//   (sqrt(T) + eD)^2 <= (1 + mu) T + KD,                                       KD = (1 + 1/mu) eD^2 + 1e-30
// so  d~^2 > thr (1 + 2^-9)(1 + mu) [(1 + mu)(a~^2 + b~^2) + K] + KD  implies |d| >= |d~| - eD > sqrt(thr (1 + 2^-9)(a^2 + b^2)):
// the true ratio exceeds thr by 2^-9 relative -- far more than the rounding of this test (a few u) and of the exact
// fp64 path -- so (float)max(e1, e2) > thr and f_inlier returns false on every one of its paths (a^2 + b^2 = 0 and
// NaN included).  The 1e-30 terms also cover the squares' own underflow.  Any NaN / inf (NaN coordinate in the tile,
// huge model, filter off: KD = inf) makes the single comparison fail, i.e. sends the evaluation to the exact test.
struct FFilter {
    float f[9];
    bool on;
};
constexpr float FILTER_TINY = 1e-30f;
constexpr float FILTER_EPS_A = 4.76837158203125e-07f;     // 8 * 2^-24
constexpr float FILTER_EPS_D = 9.5367431640625e-07f;      // 16 * 2^-24
constexpr float FILTER_MU1 = 1.000244140625f;             // 1 + mu, mu = 2^-12
constexpr float FILTER_SQ = 64.02f;                       // sqrt(1 + 1/mu) = 64.0078.., rounded up
constexpr float FILTER_MARGIN = 1.001953125f * 1.000244140625f;   // (1 + 2^-9)(1 + mu)

MCV_HD void f_filter_init(const double* F, FFilter& ff) {
    bool on = true;
    for (int i = 0; i < 9; ++i) {
        const double v = fabs(F[i]);
        if (!(v <= 1e30) || (v != 0.0 && v < 1e-30)) on = false;     // also catches NaN
        ff.f[i] = (float)F[i];
    }
    ff.on = on;
}

struct FFilterTile {
    float K, KD;
};
MCV_HD void f_filter_tile(const FFilter& ff, float X1, float Y1, float X2, float Y2, FFilterTile& t) {
    const float A = fmaf(fabsf(ff.f[0]), X1, fmaf(fabsf(ff.f[1]), Y1, fabsf(ff.f[2]) + FILTER_TINY));
    const float B = fmaf(fabsf(ff.f[3]), X1, fmaf(fabsf(ff.f[4]), Y1, fabsf(ff.f[5]) + FILTER_TINY));
    const float Cc = fmaf(fabsf(ff.f[6]), X1, fmaf(fabsf(ff.f[7]), Y1, fabsf(ff.f[8]) + FILTER_TINY));
    const float D = fmaf(X2, A, fmaf(Y2, B, Cc));
    const float sA = FILTER_SQ * FILTER_EPS_A * A, sB = FILTER_SQ * FILTER_EPS_A * B, sD = FILTER_SQ * FILTER_EPS_D * D;
    t.K = fmaf(sA, sA, fmaf(sB, sB, 1e-30f));
    t.KD = ff.on ? fmaf(sD, sD, 1e-30f) : INFINITY;
}

// thrM = thr * FILTER_MARGIN, precomputed by the caller
MCV_HD bool f_certain_outlier(const FFilter& ff, const FFilterTile& t, float x1, float y1, float x2, float y2, float thrM) {
    const float a = fmaf(ff.f[0], x1, fmaf(ff.f[1], y1, ff.f[2]));
    const float b = fmaf(ff.f[3], x1, fmaf(ff.f[4], y1, ff.f[5]));
    const float c = fmaf(ff.f[6], x1, fmaf(ff.f[7], y1, ff.f[8]));
    const float d = fmaf(x2, a, fmaf(y2, b, c));
    const float N = fmaf(fmaf(a, a, b * b), FILTER_MU1, t.K);
    return d * d > fmaf(thrM, N, t.KD);
}

MCV_HD bool e_inlier(const double* E, double x1, double y1, double x2, double y2, float thr) {
    const double t = (double)thr, tl = t * (1.0 - 2e-7), th = t * (1.0 + 2e-7);
    const double Ex0 = dadd(dadd(dmul(E[0], x1), dmul(E[1], y1)), E[2]);
    const double Ex1 = dadd(dadd(dmul(E[3], x1), dmul(E[4], y1)), E[5]);
    const double Ex2 = dadd(dadd(dmul(E[6], x1), dmul(E[7], y1)), E[8]);
    const double Et0 = dadd(dadd(dmul(E[0], x2), dmul(E[3], y2)), E[6]);
    const double Et1 = dadd(dadd(dmul(E[1], x2), dmul(E[4], y2)), E[7]);
    const double xe = dadd(dadd(dmul(x2, Ex0), dmul(y2, Ex1)), Ex2);
    const double den = dadd(dadd(dadd(dmul(Ex0, Ex0), dmul(Ex1, Ex1)), dmul(Et0, Et0)), dmul(Et1, Et1));
    const double q = dmul(xe, xe);
    if (den > 0 && den <= DBL_MAX) {                            // finite: inf <= inf would pass where inf / inf = NaN fails
        if (q <= tl * den) return true;
        if (q > th * den) return false;
    }
    return (float)ddiv(q, den) <= thr;
}

// ---- fp32 filter in front of e_inlier ------------------------------------------------------------
// Same construction for the Sampson error q / den with q = (x2' E x1)^2 and den = (E x1)_0^2 + (E x1)_1^2 +
// (E' x2)_0^2 + (E' x2)_1^2.  The correspondences are fp64 here (normalised coordinates), so the filter works on
// their fp32 roundings (one more relative error u per coordinate; the budget above -- 4u per row, 8u for the
// residual -- still holds and EPS_A = 8u, EPS_D = 16u keep their slack).  The rejection implies q / den >
// thr (1 + 2^-9), and e_inlier returns false on each of its paths (den = 0 included: q / 0 = inf).
MCV_HD void e_filter_tile(const FFilter& ff, float X1, float Y1, float X2, float Y2, FFilterTile& t) {
    const float A = fmaf(fabsf(ff.f[0]), X1, fmaf(fabsf(ff.f[1]), Y1, fabsf(ff.f[2]) + FILTER_TINY));
    const float B = fmaf(fabsf(ff.f[3]), X1, fmaf(fabsf(ff.f[4]), Y1, fabsf(ff.f[5]) + FILTER_TINY));
    const float Cc = fmaf(fabsf(ff.f[6]), X1, fmaf(fabsf(ff.f[7]), Y1, fabsf(ff.f[8]) + FILTER_TINY));
    const float D = fmaf(X2, A, fmaf(Y2, B, Cc));
    const float T0 = fmaf(fabsf(ff.f[0]), X2, fmaf(fabsf(ff.f[3]), Y2, fabsf(ff.f[6]) + FILTER_TINY));
    const float T1 = fmaf(fabsf(ff.f[1]), X2, fmaf(fabsf(ff.f[4]), Y2, fabsf(ff.f[7]) + FILTER_TINY));
    const float sA = FILTER_SQ * FILTER_EPS_A * A, sB = FILTER_SQ * FILTER_EPS_A * B, sD = FILTER_SQ * FILTER_EPS_D * D;
    const float s0 = FILTER_SQ * FILTER_EPS_A * T0, s1 = FILTER_SQ * FILTER_EPS_A * T1;
    t.K = fmaf(sA, sA, fmaf(sB, sB, fmaf(s0, s0, fmaf(s1, s1, 1e-30f))));
    t.KD = ff.on ? fmaf(sD, sD, 1e-30f) : INFINITY;
}
MCV_HD bool e_certain_outlier(const FFilter& ff, const FFilterTile& t, float x1, float y1, float x2, float y2, float thrM) {
    const float a = fmaf(ff.f[0], x1, fmaf(ff.f[1], y1, ff.f[2]));
    const float b = fmaf(ff.f[3], x1, fmaf(ff.f[4], y1, ff.f[5]));
    const float c = fmaf(ff.f[6], x1, fmaf(ff.f[7], y1, ff.f[8]));
    const float d = fmaf(x2, a, fmaf(y2, b, c));
    const float t0 = fmaf(ff.f[0], x2, fmaf(ff.f[3], y2, ff.f[6]));
    const float t1 = fmaf(ff.f[1], x2, fmaf(ff.f[4], y2, ff.f[7]));
    const float N = fmaf(fmaf(a, a, fmaf(b, b, fmaf(t0, t0, t1 * t1))), FILTER_MU1, t.K);
    return d * d > fmaf(thrM, N, t.KD);
}

// ---- fp32 filter in front of h_error <= thr ------------------------------------------------------
// The exact homography test is already float32, but it must round every operation separately and divide (about 28
// instructions); the filter is the division-free FMA form:  with w = h6 x + h7 y + 1, X = h0 x + h1 y + h2, Y likewise and
// g = (X - u w, Y - v w), the squared error is |g|^2 / w^2.  Both paths start from the same float32 coefficients.
// Budget (u = 2^-24, per tile maxima Xm >= |x| etc.): the exact path's X, w and the filter's X~, w~ each differ from the
// real values by <= 3u Xb resp. 3u Wb (Xb = |h0| Xm + |h1| Ym + |h2|, Wb = |h6| Xm + |h7| Ym + 1), so with g_e = X - u w
// formed from the exact path's values |g_e - g~| <= 6u (Xb + Um Wb), and the exact path's own division / product /
// subtraction roundings give |dx_f| >= |g_e / w| (1 - 4.1u) - 3.1u Um.  With eW = 8u Wb >= |w - w~| and ex = 24u (Xb + Um Wb)
// >= 6u (Xb + Um Wb) + 3.2u Um Wb (1 + 5u), the relative factors inside the 2^-9 threshold margin, the comparison
//     g~x^2 + g~y^2 > thr (1 + 2^-9)(1 + mu)^2 w~^2 + K,  K = thr 1.003 (64.02 eW)^2 + (64.02 ex)^2 + (64.02 ey)^2 + 1e-30
// together with |w~| > eW implies that the error h_error computes exceeds thr (an exact-path overflow or w = 0 gives
// inf / NaN there, which is an outlier as well).
struct HFilterTile {
    float K, eW, thrMM;
};
MCV_HD void h_filter_tile(const float* Hf, float Xm, float Ym, float Um, float Vm, float thr, HFilterTile& t) {
    const float Xb = fmaf(fabsf(Hf[0]), Xm, fmaf(fabsf(Hf[1]), Ym, fabsf(Hf[2]) + FILTER_TINY));
    const float Yb = fmaf(fabsf(Hf[3]), Xm, fmaf(fabsf(Hf[4]), Ym, fabsf(Hf[5]) + FILTER_TINY));
    const float Wb = fmaf(fabsf(Hf[6]), Xm, fmaf(fabsf(Hf[7]), Ym, 1.0f));
    const float u24 = 1.430511474609375e-06f;                    // 24 * 2^-24
    const float eW = FILTER_EPS_A * Wb;
    const float ex = u24 * fmaf(Um, Wb, Xb), ey = u24 * fmaf(Vm, Wb, Yb);
    const float sW = FILTER_SQ * eW, sx = FILTER_SQ * ex, sy = FILTER_SQ * ey;
    bool on = true;
    for (int i = 0; i < 8; ++i) {
        const float v = fabsf(Hf[i]);
        if (!(v <= 1e30f) || (v != 0.0f && v < 1e-30f)) on = false;
    }
    t.eW = eW;
    t.thrMM = thr * (1.001953125f * 1.000244140625f * 1.000244140625f * 1.0001f);
    t.K = on ? fmaf(thr * 1.003f, sW * sW, fmaf(sx, sx, fmaf(sy, sy, 1e-30f))) : INFINITY;
}
MCV_HD bool h_certain_outlier(const float* Hf, const HFilterTile& t, float x, float y, float u, float v) {
    const float w = fmaf(Hf[6], x, fmaf(Hf[7], y, 1.0f));
    const float X = fmaf(Hf[0], x, fmaf(Hf[1], y, Hf[2]));
    const float Y = fmaf(Hf[3], x, fmaf(Hf[4], y, Hf[5]));
    const float gx = fmaf(-u, w, X), gy = fmaf(-v, w, Y);
    const float lhs = fmaf(gx, gx, gy * gy);
    const float rhs = fmaf(t.thrMM, w * w, t.K);
    return lhs > rhs && fabsf(w) > t.eW;
}

#if defined(__CUDACC__)
// ---- the two filters on Blackwell's packed fp32 pipe: two correspondences per instruction ------------------------
// fma.rn.f32x2 / mul.rn.f32x2 (sm_100+, SASS FFMA2) operate on a register pair; the filter is bound by instruction
// issue, so evaluating two correspondences per instruction halves its cost.  Same formulas, same constants as the
// scalar forms above (every packed operation rounds exactly like its scalar counterpart).
struct F2 { unsigned long long v; };
__device__ __forceinline__ F2 f2_make(float lo, float hi) {
    F2 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r.v) : "f"(lo), "f"(hi)); return r;
}
__device__ __forceinline__ F2 f2_bcast(float x) { return f2_make(x, x); }
__device__ __forceinline__ void f2_split(F2 a, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(a.v)); }
__device__ __forceinline__ F2 f2_fma(F2 a, F2 b, F2 c) {
    F2 r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r.v) : "l"(a.v), "l"(b.v), "l"(c.v)); return r;
}
__device__ __forceinline__ F2 f2_mul(F2 a, F2 b) {
    F2 r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v)); return r;
}
struct FFilter2 {          // coefficients and per-tile constants broadcast to both halves
    F2 f[9], K, KD, mu1, thrM;
};
__device__ __forceinline__ void filter2_init(const FFilter& ff, FFilter2& p, float thrM) {
#pragma unroll
    for (int i = 0; i < 9; ++i) p.f[i] = f2_bcast(ff.f[i]);
    p.mu1 = f2_bcast(FILTER_MU1);
    p.thrM = f2_bcast(thrM);
}
__device__ __forceinline__ void filter2_tile(FFilter2& p, const FFilterTile& t) { p.K = f2_bcast(t.K); p.KD = f2_bcast(t.KD); }
// X1 = (x1 of point 0, x1 of point 1) etc.; r0 / r1 = "certainly an outlier" for the two points
__device__ __forceinline__ void f_certain_outlier2(const FFilter2& p, F2 X1, F2 Y1, F2 X2, F2 Y2, bool& r0, bool& r1) {
    const F2 a = f2_fma(p.f[0], X1, f2_fma(p.f[1], Y1, p.f[2]));
    const F2 b = f2_fma(p.f[3], X1, f2_fma(p.f[4], Y1, p.f[5]));
    const F2 c = f2_fma(p.f[6], X1, f2_fma(p.f[7], Y1, p.f[8]));
    const F2 d = f2_fma(X2, a, f2_fma(Y2, b, c));
    const F2 N = f2_fma(f2_fma(a, a, f2_mul(b, b)), p.mu1, p.K);
    const F2 rhs = f2_fma(p.thrM, N, p.KD);
    const F2 lhs = f2_mul(d, d);
    float l0, l1, h0, h1;
    f2_split(lhs, l0, l1);
    f2_split(rhs, h0, h1);
    r0 = l0 > h0;
    r1 = l1 > h1;
}
__device__ __forceinline__ void e_certain_outlier2(const FFilter2& p, F2 X1, F2 Y1, F2 X2, F2 Y2, bool& r0, bool& r1) {
    const F2 a = f2_fma(p.f[0], X1, f2_fma(p.f[1], Y1, p.f[2]));
    const F2 b = f2_fma(p.f[3], X1, f2_fma(p.f[4], Y1, p.f[5]));
    const F2 c = f2_fma(p.f[6], X1, f2_fma(p.f[7], Y1, p.f[8]));
    const F2 d = f2_fma(X2, a, f2_fma(Y2, b, c));
    const F2 t0 = f2_fma(p.f[0], X2, f2_fma(p.f[3], Y2, p.f[6]));
    const F2 t1 = f2_fma(p.f[1], X2, f2_fma(p.f[4], Y2, p.f[7]));
    const F2 N = f2_fma(f2_fma(a, a, f2_fma(b, b, f2_fma(t0, t0, f2_mul(t1, t1)))), p.mu1, p.K);
    const F2 rhs = f2_fma(p.thrM, N, p.KD);
    const F2 lhs = f2_mul(d, d);
    float l0, l1, h0, h1;
    f2_split(lhs, l0, l1);
    f2_split(rhs, h0, h1);
    r0 = l0 > h0;
    r1 = l1 > h1;
}
#endif

// ---- subset checks ----------------------------------------------------------------------------
// haveCollinearPoints: last point against every earlier pair; pts as float (x,y) pairs.
MCV_HDN inline bool have_collinear(const float* pts, int count) {
    const int i = count - 1;
    for (int j = 0; j < i; ++j) {
        const double dx1 = (double)pts[2 * j] - (double)pts[2 * i];
        const double dy1 = (double)pts[2 * j + 1] - (double)pts[2 * i + 1];
        for (int k = 0; k < j; ++k) {
            const double dx2 = (double)pts[2 * k] - (double)pts[2 * i];
            const double dy2 = (double)pts[2 * k + 1] - (double)pts[2 * i + 1];
            if (fabs(dx2 * dy1 - dy2 * dx1) <=
                (double)FLT_EPSILON * (fabs(dx1) + fabs(dy1) + fabs(dx2) + fabs(dy2)))
                return true;
        }
    }
    return false;
}

MCV_HD double det3pts(const float* p, int i0, int i1, int i2) {
    const double a00 = p[2 * i0], a01 = p[2 * i0 + 1], a10 = p[2 * i1], a11 = p[2 * i1 + 1],
                 a20 = p[2 * i2], a21 = p[2 * i2 + 1];
    return a00 * (a11 * 1. - a21 * 1.) - a01 * (a10 * 1. - a20 * 1.) + 1. * (a10 * a21 - a20 * a11);
}

// HomographyEstimatorCallback::checkSubset for count == 4
MCV_HDN inline bool h_check_subset(const float* src, const float* dst) {
    if (have_collinear(src, 4) || have_collinear(dst, 4)) return false;
    const int tt[4][3] = {{0, 1, 2}, {1, 2, 3}, {0, 2, 3}, {0, 1, 3}};
    int negative = 0;
    for (int i = 0; i < 4; ++i)
        negative += det3pts(src, tt[i][0], tt[i][1], tt[i][2]) * det3pts(dst, tt[i][0], tt[i][1], tt[i][2]) < 0;
    return negative == 0 || negative == 4;
}

// ---- homography from 4 correspondences ------------------------------------------------------
// src/dst: 4 (x,y) float pairs.  H: 9 doubles, scaled so H[8] = 1.  Returns false if degenerate.
// OpenCV's kernel (HomographyEstimatorCallback::runKernel) normalises the points per axis and takes the eigenvector of the
// smallest eigenvalue of the 9 x 9 L'L; with four points that vector is the unique null vector of the 8 x 9 system, so any
// exact method returns the same matrix up to rounding.  Here, in registers: the projective-basis form.  With P_i = (X_i, Y_i,
// 1) (normalised source), M = [P1 P2 P3], lambda = adj(M) P4, the map A = M diag(lambda) takes the canonical basis e1, e2,
// e3, e1 + e2 + e3 to P1..P4; likewise B = N diag(mu) for the targets.  H_n = B A^-1 ~ N diag(c) adj(M) with c_i = mu_i
// lambda_j lambda_k, and the rows of adj(M) are the cross products P2 x P3, P3 x P1, P1 x P2.  ~130 flops instead of a
// full-pivot 8 x 9 elimination in shared memory (round 1/2: 2.0 ms of a C3-H step).  A vanishing lambda_i / mu_i means
// three of the four points are collinear: no model (the sampler's degeneracy test rejects those subsets beforehand).
MCV_HDN inline bool h_solve4(const float* src, const float* dst, double* H) {
    double cMx = 0, cMy = 0, cmx = 0, cmy = 0;
    for (int i = 0; i < 4; ++i) { cmx += dst[2 * i]; cmy += dst[2 * i + 1]; cMx += src[2 * i]; cMy += src[2 * i + 1]; }
    cmx /= 4; cmy /= 4; cMx /= 4; cMy /= 4;
    double smx = 0, smy = 0, sMx = 0, sMy = 0;
    for (int i = 0; i < 4; ++i) {
        smx += fabs(dst[2 * i] - cmx); smy += fabs(dst[2 * i + 1] - cmy);
        sMx += fabs(src[2 * i] - cMx); sMy += fabs(src[2 * i + 1] - cMy);
    }
    if (fabs(smx) < DBL_EPSILON || fabs(smy) < DBL_EPSILON || fabs(sMx) < DBL_EPSILON || fabs(sMy) < DBL_EPSILON)
        return false;
    smx = 4 / smx; smy = 4 / smy; sMx = 4 / sMx; sMy = 4 / sMy;
    double X[4], Y[4], x[4], y[4];
    for (int i = 0; i < 4; ++i) {
        x[i] = (dst[2 * i] - cmx) * smx; y[i] = (dst[2 * i + 1] - cmy) * smy;
        X[i] = (src[2 * i] - cMx) * sMx; Y[i] = (src[2 * i + 1] - cMy) * sMy;
    }
    // rows of adj([P1 P2 P3]) for the sources, lambda = adj . P4
    const double r[3][3] = {{Y[1] - Y[2], X[2] - X[1], X[1] * Y[2] - X[2] * Y[1]},
                            {Y[2] - Y[0], X[0] - X[2], X[2] * Y[0] - X[0] * Y[2]},
                            {Y[0] - Y[1], X[1] - X[0], X[0] * Y[1] - X[1] * Y[0]}};
    double lam[3], mu[3];
    for (int i = 0; i < 3; ++i) lam[i] = r[i][0] * X[3] + r[i][1] * Y[3] + r[i][2];
    // the same for the targets: mu = adj([p1 p2 p3]) p4
    mu[0] = (y[1] - y[2]) * x[3] + (x[2] - x[1]) * y[3] + (x[1] * y[2] - x[2] * y[1]);
    mu[1] = (y[2] - y[0]) * x[3] + (x[0] - x[2]) * y[3] + (x[2] * y[0] - x[0] * y[2]);
    mu[2] = (y[0] - y[1]) * x[3] + (x[1] - x[0]) * y[3] + (x[0] * y[1] - x[1] * y[0]);
    for (int i = 0; i < 3; ++i)
        if (!(fabs(lam[i]) > 1e-13) || !(fabs(mu[i]) > 1e-13)) return false;
    const double c[3] = {mu[0] * lam[1] * lam[2], mu[1] * lam[0] * lam[2], mu[2] * lam[0] * lam[1]};
    double h[9];
    for (int j = 0; j < 3; ++j) {
        const double t0 = c[0] * r[0][j], t1 = c[1] * r[1][j], t2 = c[2] * r[2][j];
        h[j] = x[0] * t0 + x[1] * t1 + x[2] * t2;
        h[3 + j] = y[0] * t0 + y[1] * t1 + y[2] * t2;
        h[6 + j] = t0 + t1 + t2;
    }
    // H = invHnorm * h * Hnorm2
    const double iH[9] = {1. / smx, 0, cmx, 0, 1. / smy, cmy, 0, 0, 1};
    const double N2[9] = {sMx, 0, -cMx * sMx, 0, sMy, -cMy * sMy, 0, 0, 1};
    double T[9];
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) T[i * 3 + j] = iH[i * 3] * h[j] + iH[i * 3 + 1] * h[3 + j] + iH[i * 3 + 2] * h[6 + j];
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) H[i * 3 + j] = T[i * 3] * N2[j] + T[i * 3 + 1] * N2[3 + j] + T[i * 3 + 2] * N2[6 + j];
    if (H[8] == 0 || !(fabs(H[8]) > 0)) return false;
    const double sc = 1. / H[8];
    for (int k = 0; k < 9; ++k) H[k] *= sc;
    for (int k = 0; k < 9; ++k) if (!(fabs(H[k]) < 1e300)) return false;
    return true;
}

// ---- cv::solveCubic (closed form); coeffs c[0] x^3 + c[1] x^2 + c[2] x + c[3] -----------------
MCV_HDN inline int solve_cubic(const double* cf, double* roots) {
    double a0 = cf[0], a1 = cf[1], a2 = cf[2], a3 = cf[3];
    double x0 = 0, x1 = 0, x2 = 0;
    int n = 0;
    if (a0 == 0) {
        if (a1 == 0) {
            if (a2 == 0) n = a3 == 0 ? -1 : 0;
            else { x0 = -a3 / a2; n = 1; }
        } else {
            double d = a2 * a2 - 4 * a1 * a3;
            if (d >= 0) {
                d = sqrt(d);
                const double q1 = (-a2 + d) * 0.5, q2 = (a2 + d) * -0.5;
                if (fabs(q1) > fabs(q2)) { x0 = q1 / a1; x1 = a3 / q1; }
                else { x0 = q2 / a1; x1 = a3 / q2; }
                n = d > 0 ? 2 : 1;
            }
        }
    } else {
        a0 = 1. / a0; a1 *= a0; a2 *= a0; a3 *= a0;
        const double Q = (a1 * a1 - 3 * a2) * (1. / 9);
        const double R = (2 * a1 * a1 * a1 - 9 * a1 * a2 + 27 * a3) * (1. / 54);
        const double Qcubed = Q * Q * Q;
        double d = Qcubed - R * R;
        const double PI = 3.14159265358979323846;
        if (d > 0) {
            const double theta = acos(R / sqrt(Qcubed));
            const double sqrtQ = sqrt(Q);
            const double t0 = -2 * sqrtQ, t1 = theta * (1. / 3), t2 = a1 * (1. / 3);
            x0 = t0 * cos(t1) - t2;
            x1 = t0 * cos(t1 + (2. * PI / 3)) - t2;
            x2 = t0 * cos(t1 + (4. * PI / 3)) - t2;
            n = 3;
        } else if (d == 0) {
            if (R >= 0) { x0 = -2 * pow(R, 1. / 3) - a1 / 3; x1 = pow(R, 1. / 3) - a1 / 3; }
            else { x0 = 2 * pow(-R, 1. / 3) - a1 / 3; x1 = -pow(-R, 1. / 3) - a1 / 3; }
            x2 = 0;
            n = x0 == x1 ? 1 : 2;
            x1 = x0 == x1 ? 0 : x1;
        } else {
            d = sqrt(-d);
            double e = pow(d + fabs(R), 1. / 3);
            if (R > 0) e = -e;
            x0 = (e + Q / e) - a1 * (1. / 3);
            n = 1;
        }
    }
    roots[0] = x0; roots[1] = x1; roots[2] = x2;
    return n;
}

// ---- 7-point fundamental matrix (run7Point) ----------------------------------------------------
// m1/m2: 7 (x,y) float pairs.  F: up to 3 matrices (9 doubles each).  Returns the count.
// At: 63 doubles of scratch with element stride S (the 7 x 9 design matrix)
template <int S>
MCV_HDN inline int f_solve7_at(const float* m1, const float* m2, double* Fout, double* At) {
    for (int i = 0; i < 7; ++i) {
        const double x0 = m1[2 * i], y0 = m1[2 * i + 1], x1 = m2[2 * i], y1 = m2[2 * i + 1];
        double* a = At + i * 9 * S;
        a[0] = x1 * x0; a[1 * S] = x1 * y0; a[2 * S] = x1; a[3 * S] = y1 * x0; a[4 * S] = y1 * y0; a[5 * S] = y1;
        a[6 * S] = x0; a[7 * S] = y0; a[8 * S] = 1;
    }
    double ext[18];
    nullspace_rows<9, 7, 9, S>(At, ext);   // ext = rows 7, 8 of OpenCV's vt = the null-space basis run7Point uses
    double f1[9], f2[9];
    for (int i = 0; i < 9; ++i) { f2[i] = ext[9 + i]; f1[i] = ext[i] - f2[i]; }
    double c[4], r[3] = {0, 0, 0};
    const double t0 = f2[4] * f2[8] - f2[5] * f2[7];
    const double t1 = f2[3] * f2[8] - f2[5] * f2[6];
    const double t2 = f2[3] * f2[7] - f2[4] * f2[6];
    c[3] = f2[0] * t0 - f2[1] * t1 + f2[2] * t2;
    c[2] = f1[0] * t0 - f1[1] * t1 + f1[2] * t2 -
           f1[3] * (f2[1] * f2[8] - f2[2] * f2[7]) + f1[4] * (f2[0] * f2[8] - f2[2] * f2[6]) -
           f1[5] * (f2[0] * f2[7] - f2[1] * f2[6]) + f1[6] * (f2[1] * f2[5] - f2[2] * f2[4]) -
           f1[7] * (f2[0] * f2[5] - f2[2] * f2[3]) + f1[8] * (f2[0] * f2[4] - f2[1] * f2[3]);
    const double u0 = f1[4] * f1[8] - f1[5] * f1[7];
    const double u1 = f1[3] * f1[8] - f1[5] * f1[6];
    const double u2 = f1[3] * f1[7] - f1[4] * f1[6];
    c[1] = f2[0] * u0 - f2[1] * u1 + f2[2] * u2 -
           f2[3] * (f1[1] * f1[8] - f1[2] * f1[7]) + f2[4] * (f1[0] * f1[8] - f1[2] * f1[6]) -
           f2[5] * (f1[0] * f1[7] - f1[1] * f1[6]) + f2[6] * (f1[1] * f1[5] - f1[2] * f1[4]) -
           f2[7] * (f1[0] * f1[5] - f1[2] * f1[3]) + f2[8] * (f1[0] * f1[4] - f1[1] * f1[3]);
    c[0] = f1[0] * u0 - f1[1] * u1 + f1[2] * u2;
    const int n = solve_cubic(c, r);
    if (n < 1 || n > 3) return n < 0 ? 0 : (n > 3 ? 0 : n);
    for (int k = 0; k < n; ++k) {
        double* F = Fout + 9 * k;
        double lambda = r[k], mu = 1.;
        const double s = f1[8] * r[k] + f2[8];
        if (fabs(s) > DBL_EPSILON) { mu = 1. / s; lambda *= mu; F[8] = 1.; }
        else F[8] = 0.;
        for (int i = 0; i < 8; ++i) F[i] = f1[i] * lambda + f2[i] * mu;
    }
    return n;
}
MCV_HDN inline int f_solve7(const float* m1, const float* m2, double* Fout) {
    double At[7 * 9];
    return f_solve7_at<1>(m1, m2, Fout, At);
}

// ---- 5-point essential matrix ------------------------------------------------------------------
// Small dense polynomials in (x, y, z): deg1 = [x, y, z, 1]; deg2 (10) and deg3 (20) indexed by
// the tables below.  deg3 order = the elimination order used by the solver:
//   x3 y3 x2y xy2 x2z x2 y2z y2 xyz xy | xz2 xz x yz2 yz y z3 z2 z 1
namespace p5 {
// deg2 order: x2 y2 z2 xy xz yz x y z 1
MCV_HD void mul11(const double* a, const double* b, double* r) {
    r[0] = a[0] * b[0]; r[1] = a[1] * b[1]; r[2] = a[2] * b[2];
    r[3] = a[0] * b[1] + a[1] * b[0];
    r[4] = a[0] * b[2] + a[2] * b[0];
    r[5] = a[1] * b[2] + a[2] * b[1];
    r[6] = a[0] * b[3] + a[3] * b[0];
    r[7] = a[1] * b[3] + a[3] * b[1];
    r[8] = a[2] * b[3] + a[3] * b[2];
    r[9] = a[3] * b[3];
}
// deg2 (a) x deg1 (b) accumulated into deg3 r (+= s * a*b)
MCV_HD void mul21_acc(const double* a, const double* b, double s, double* r) {
    const double bx = b[0] * s, by = b[1] * s, bz = b[2] * s, b1 = b[3] * s;
    // a: 0 x2, 1 y2, 2 z2, 3 xy, 4 xz, 5 yz, 6 x, 7 y, 8 z, 9 one
    r[0] += a[0] * bx;                              // x3
    r[1] += a[1] * by;                              // y3
    r[2] += a[0] * by + a[3] * bx;                  // x2y
    r[3] += a[1] * bx + a[3] * by;                  // xy2
    r[4] += a[0] * bz + a[4] * bx;                  // x2z
    r[5] += a[0] * b1 + a[6] * bx;                  // x2
    r[6] += a[1] * bz + a[5] * by;                  // y2z
    r[7] += a[1] * b1 + a[7] * by;                  // y2
    r[8] += a[3] * bz + a[4] * by + a[5] * bx;      // xyz
    r[9] += a[3] * b1 + a[6] * by + a[7] * bx;      // xy
    r[10] += a[2] * bx + a[4] * bz;                 // xz2
    r[11] += a[4] * b1 + a[6] * bz + a[8] * bx;     // xz
    r[12] += a[6] * b1 + a[9] * bx;                 // x
    r[13] += a[2] * by + a[5] * bz;                 // yz2
    r[14] += a[5] * b1 + a[7] * bz + a[8] * by;     // yz
    r[15] += a[7] * b1 + a[9] * by;                 // y
    r[16] += a[2] * bz;                             // z3
    r[17] += a[2] * b1 + a[8] * bz;                 // z2
    r[18] += a[8] * b1 + a[9] * bz;                 // z
    r[19] += a[9] * b1;                             // 1
}
}  // namespace p5

// univariate polynomial helpers (coefficients in DESCENDING powers, like the reference's b[])
MCV_HD void upoly_mul(const double* a, int na, const double* b, int nb, double* r) {
    for (int i = 0; i < na + nb - 1; ++i) r[i] = 0;
    for (int i = 0; i < na; ++i)
        for (int j = 0; j < nb; ++j) r[i + j] += a[i] * b[j];
}

// Real roots of a degree-10 polynomial c[0] z^10 + ... + c[10] (real coefficients).
// The reference calls cv::solvePoly (Durand-Kerner on all complex roots) and keeps roots with
// |Im| <= 1e-10 (fivepoint.cpp:296-303).  Here: Laguerre iterations with deflation by real
// linear / real quadratic factors (coefficients stay real), every root then polished on the
// ORIGINAL polynomial before the same |Im| <= 1e-10 filter.  ~10x fewer flops than a
// simultaneous all-roots iteration, and all state fits in registers.
MCV_HDN inline void laguerre_step_poly(const double* a, int deg, double& xr, double& xi, int maxIter) {
    // a[0] z^deg + ... + a[deg]; iterates (xr, xi) towards a root of this polynomial
    for (int iter = 0; iter < maxIter; ++iter) {
        double br = a[0], bi = 0, dr = 0, di = 0, fr = 0, fi = 0;
        double err = fabs(br);
        const double ax = sqrt(xr * xr + xi * xi);
        for (int j = 1; j <= deg; ++j) {
            // f = x*f + d; d = x*d + b; b = x*b + a[j]
            const double nfr = xr * fr - xi * fi + dr, nfi = xr * fi + xi * fr + di;
            const double ndr = xr * dr - xi * di + br, ndi = xr * di + xi * dr + bi;
            const double nbr = xr * br - xi * bi + a[j], nbi = xr * bi + xi * br;
            fr = nfr; fi = nfi; dr = ndr; di = ndi; br = nbr; bi = nbi;
            err = (fabs(br) + fabs(bi)) + ax * err;           // |b| <= |Re| + |Im|: a bound, no square root per term
        }
        const double b2 = br * br + bi * bi;
        const double floor_ = err * 2e-16;
        if (b2 <= floor_ * floor_) return;                    // at the rounding floor of p(x)
        // G = d/b, G2 = G^2, H = G2 - 2 f / b
        const double ib2 = 1 / b2;
        const double gr = (dr * br + di * bi) * ib2, gi = (di * br - dr * bi) * ib2;
        const double g2r = gr * gr - gi * gi, g2i = 2 * gr * gi;
        const double qr = (fr * br + fi * bi) * ib2, qi = (fi * br - fr * bi) * ib2;
        const double hr = g2r - 2 * qr, hi = g2i - 2 * qi;
        // sq = sqrt((deg-1) * (deg*H - G2))
        const double tr = (deg - 1) * (deg * hr - g2r), ti = (deg - 1) * (deg * hi - g2i);
        const double tm = sqrt(tr * tr + ti * ti);
        double sr = sqrt(fmax(0.0, 0.5 * (tm + tr))), si = sqrt(fmax(0.0, 0.5 * (tm - tr)));
        if (ti < 0) si = -si;
        double p1r = gr + sr, p1i = gi + si, p2r = gr - sr, p2i = gi - si;
        const double m1 = p1r * p1r + p1i * p1i, m2 = p2r * p2r + p2i * p2i;
        if (m2 > m1) { p1r = p2r; p1i = p2i; }
        const double pm = fmax(m1, m2);
        double dxr, dxi;
        if (pm > 0) { const double ipm = deg / pm; dxr = p1r * ipm; dxi = -p1i * ipm; }
        else { dxr = (1 + ax) * cos((double)iter); dxi = (1 + ax) * sin((double)iter); }
        const double nxr = xr - dxr, nxi = xi - dxi;
        if (nxr == xr && nxi == xi) return;
        xr = nxr; xi = nxi;
        if (fabs(dxr) + fabs(dxi) <= 1e-16 * (fabs(xr) + fabs(xi))) return;
    }
}

MCV_HDN inline int real_roots10(const double* c, double* out) {
    int deg = 10, lead = 0;
    while (deg > 0 && c[lead] == 0) { ++lead; --deg; }
    if (deg == 0) return 0;
    double a0[11], w[11];
    for (int i = 0; i <= deg; ++i) { a0[i] = c[lead + i]; w[i] = a0[i]; }
    const int deg0 = deg;
    int n = 0;
    while (deg >= 1) {
        double xr = 0, xi = 0;
        if (deg == 1) { xr = -w[1] / w[0]; xi = 0; }
        else laguerre_step_poly(w, deg, xr, xi, 80);
        // polish on the original polynomial: complex Newton steps (local, so the polished value
        // stays with the root the deflated estimate belongs to)
        double pr = xr, pi = xi;
        // a clearly complex estimate cannot polish into a real root (|Im| <= 1e-10): it is only deflated away
        const int polishSteps = fabs(xi) > 1e-4 * (1 + fabs(xr)) ? 0 : 4;
        for (int it = 0; it < polishSteps; ++it) {
            double br = a0[0], bi = 0, dr = 0, di = 0;
            for (int j = 1; j <= deg0; ++j) {
                const double ndr = pr * dr - pi * di + br, ndi = pr * di + pi * dr + bi;
                const double nbr = pr * br - pi * bi + a0[j], nbi = pr * bi + pi * br;
                dr = ndr; di = ndi; br = nbr; bi = nbi;
            }
            const double d2 = dr * dr + di * di;
            if (d2 == 0) break;
            const double id2 = 1 / d2;
            const double sr = (br * dr + bi * di) * id2, si = (bi * dr - br * di) * id2;
            // accept the step only while it stays a small correction of the estimate
            if (fabs(sr) + fabs(si) > 1e-3 * (fabs(pr) + fabs(pi)) + 1e-300) break;
            pr -= sr; pi -= si;
            if (fabs(sr) + fabs(si) <= 1e-16 * (fabs(pr) + fabs(pi))) break;
        }
        const bool real_root = fabs(pi) <= 1e-10;
        if (real_root) out[n++] = pr;
        // deflate the working polynomial by a real factor
        if (fabs(xi) <= 1e-12 * (1 + fabs(xr)) || deg == 1) {
            double rem = w[0];
            for (int j = 1; j < deg; ++j) { const double t = w[j] + rem * xr; w[j] = t; rem = t; }
            // w[0..deg-1] now holds the quotient (w[0] unchanged)
            deg -= 1;
        } else {
            // quadratic z^2 - 2 Re(x) z + |x|^2 ; the conjugate is a root too
            const double pcoef = 2 * xr, qcoef = -(xr * xr + xi * xi);
            double b0 = w[0], b1 = w[1] + pcoef * b0;
            w[1] = b1;
            for (int j = 2; j <= deg - 2; ++j) { const double t = w[j] + pcoef * b1 + qcoef * b0; b0 = b1; b1 = t; w[j] = t; }
            deg -= 2;
        }
    }
    for (int i = 1; i < n; ++i) {           // ascending order: deterministic model order
        const double v = out[i];
        int j = i - 1;
        while (j >= 0 && out[j] > v) { out[j + 1] = out[j]; --j; }
        out[j + 1] = v;
    }
    // two deflated estimates that polish onto the same root would double-count it
    int m = 0;
    for (int i = 0; i < n; ++i)
        if (m == 0 || fabs(out[i] - out[m - 1]) > 1e-13 * (fabs(out[i]) + fabs(out[m - 1]))) out[m++] = out[i];
    return m;
}

// pa/pb: 5 points each as (x,y) doubles (already normalised by the intrinsics).
// Es: up to 10 matrices, 9 doubles each, row-major, unit Frobenius norm.  Returns the count.
// Follows the reference's runFivepoint (src/MiniCVNative/fivepoint.cpp:233-339); the
// generated coefficient table of getCoeffMat (:10-231) is replaced by polynomial arithmetic.
// The solver in three stages, so that the GPU can run them as separate kernels (one thread per hypothesis for the
// first two, one thread per root for the last); e_solve5 below is their composition.
// Stage 1: null-space basis EE (4 x 9), the three rows of B(z) (b: 3 x 13) and det B(z) (coef: degree 10, descending).
MCV_HDN inline bool e5_stage1(const double* pa, const double* pb, double* EEout /*[36]*/, double* bout /*[39]*/,
                              double* coef /*[11]*/) {
    // 1. null space of the 5x9 system: rows 5..8 of Vt (OpenCV basis, incl. its completion)
    double At[9 * 9];
    for (int i = 0; i < 5; ++i) {
        const double x1 = pa[2 * i], y1 = pa[2 * i + 1], x2 = pb[2 * i], y2 = pb[2 * i + 1];
        double* q = At + 9 * i;
        q[0] = x1 * x2; q[1] = y1 * x2; q[2] = x2; q[3] = x1 * y2; q[4] = y1 * y2; q[5] = y2;
        q[6] = x1; q[7] = y1; q[8] = 1.0;
    }
    nullspace_rows<9, 5, 9, 1>(At, At + 45);
    const double* EE = At + 5 * 9;   // 4 x 9: X, Y, Z, W

    // 2. cubic constraints.  E[i][j] as deg1 polys: coefficient c of basis b at e1[i*3+j][b]
    double e1[9][4];
    for (int k = 0; k < 9; ++k)
        for (int b = 0; b < 4; ++b) e1[k][b] = EE[b * 9 + k];
    double A[10][20];
    for (int r = 0; r < 10; ++r)
        for (int k = 0; k < 20; ++k) A[r][k] = 0;
    {
        // det(E)
        double m[10];
        p5::mul11(e1[4], e1[8], m); p5::mul21_acc(m, e1[0], 1., A[0]);
        p5::mul11(e1[5], e1[7], m); p5::mul21_acc(m, e1[0], -1., A[0]);
        p5::mul11(e1[3], e1[8], m); p5::mul21_acc(m, e1[1], -1., A[0]);
        p5::mul11(e1[5], e1[6], m); p5::mul21_acc(m, e1[1], 1., A[0]);
        p5::mul11(e1[3], e1[7], m); p5::mul21_acc(m, e1[2], 1., A[0]);
        p5::mul11(e1[4], e1[6], m); p5::mul21_acc(m, e1[2], -1., A[0]);
        // EEt (symmetric), trace
        double eet[3][3][10], tr[10];
        for (int i = 0; i < 3; ++i)
            for (int j = i; j < 3; ++j) {
                double t0[10], t1[10], t2[10];
                p5::mul11(e1[i * 3], e1[j * 3], t0);
                p5::mul11(e1[i * 3 + 1], e1[j * 3 + 1], t1);
                p5::mul11(e1[i * 3 + 2], e1[j * 3 + 2], t2);
                for (int k = 0; k < 10; ++k) { eet[i][j][k] = t0[k] + t1[k] + t2[k]; eet[j][i][k] = eet[i][j][k]; }
            }
        for (int k = 0; k < 10; ++k) tr[k] = eet[0][0][k] + eet[1][1][k] + eet[2][2][k];
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j) {
                double* row = A[1 + i * 3 + j];
                for (int k = 0; k < 3; ++k) p5::mul21_acc(eet[i][k], e1[k * 3 + j], 2., row);
                p5::mul21_acc(tr, e1[i * 3 + j], -1., row);
            }
    }
    // 3. Gauss-Jordan on the ten leading monomials (partial pivoting)
    for (int c = 0; c < 10; ++c) {
        int pr = c;
        double best = fabs(A[c][c]);
        for (int r = c + 1; r < 10; ++r)
            if (fabs(A[r][c]) > best) { best = fabs(A[r][c]); pr = r; }
        if (!(best > 0)) return false;
        if (pr != c) for (int k = 0; k < 20; ++k) { double t = A[c][k]; A[c][k] = A[pr][k]; A[pr][k] = t; }
        const double inv = 1 / A[c][c];
        for (int k = c; k < 20; ++k) A[c][k] *= inv;
        for (int r = 0; r < 10; ++r) {
            if (r == c) continue;
            const double f = A[r][c];
            if (f != 0) for (int k = c; k < 20; ++k) A[r][k] -= f * A[c][k];
        }
    }
    // 4. B(z) from rows (4,5), (6,7), (8,9): [x-part (4) | y-part (4) | 1-part (5)], descending
    double b[3][13];
    for (int i = 0; i < 3; ++i) {
        const double* r1 = A[2 * i + 4] + 10;
        const double* r2 = A[2 * i + 5] + 10;
        double row1[13], row2[13];
        for (int k = 0; k < 13; ++k) { row1[k] = 0; row2[k] = 0; }
        row1[1] = r1[0]; row1[2] = r1[1]; row1[3] = r1[2];
        row1[5] = r1[3]; row1[6] = r1[4]; row1[7] = r1[5];
        row1[9] = r1[6]; row1[10] = r1[7]; row1[11] = r1[8]; row1[12] = r1[9];
        row2[0] = r2[0]; row2[1] = r2[1]; row2[2] = r2[2];
        row2[4] = r2[3]; row2[5] = r2[4]; row2[6] = r2[5];
        row2[8] = r2[6]; row2[9] = r2[7]; row2[10] = r2[8]; row2[11] = r2[9];
        for (int k = 0; k < 13; ++k) b[i][k] = row1[k] - row2[k];
    }
    // det B(z): degree 10, descending coefficients
    for (int k = 0; k < 11; ++k) coef[k] = 0;
    {
        const int perm[6][3] = {{0, 1, 2}, {1, 2, 0}, {2, 0, 1}, {0, 2, 1}, {1, 0, 2}, {2, 1, 0}};
        for (int t = 0; t < 6; ++t) {
            const double sgn = t < 3 ? 1. : -1.;
            // columns: 0 -> x-part (deg 3), 1 -> y-part (deg 3), 2 -> 1-part (deg 4)
            double xy[7], full[11];
            upoly_mul(b[perm[t][0]] + 0, 4, b[perm[t][1]] + 4, 4, xy);
            upoly_mul(xy, 7, b[perm[t][2]] + 8, 5, full);
            for (int k = 0; k < 11; ++k) coef[k] += sgn * full[k];
        }
    }
    for (int k = 0; k < 36; ++k) EEout[k] = EE[k];
    for (int i = 0; i < 3; ++i)
        for (int k = 0; k < 13; ++k) bout[13 * i + k] = b[i][k];
    return true;
}

// Stage 3: one real root z of det B(z) -> the essential matrix (unit Frobenius norm); false when the root is rejected.
MCV_HDN inline bool e5_root_to_E(const double* EE /*[36]*/, const double* b /*[39]*/, double z1, double* E) {
    const double z2 = z1 * z1, z3 = z2 * z1, z4 = z3 * z1;
    double bz[9];
    for (int j = 0; j < 3; ++j) {
        const double* br = b + 13 * j;
        bz[j * 3 + 0] = br[0] * z3 + br[1] * z2 + br[2] * z1 + br[3];
        bz[j * 3 + 1] = br[4] * z3 + br[5] * z2 + br[6] * z1 + br[7];
        bz[j * 3 + 2] = br[8] * z4 + br[9] * z3 + br[10] * z2 + br[11] * z1 + br[12];
    }
    // SVD::solveZ(Bz) in the reference: the right singular vector of the smallest singular value, used only through
    // the ratios x = v0 / v2, y = v1 / v2 (its sign cancels).  At a root of det B(z) the matrix has rank two and that
    // vector is the cross product of any two independent rows; the best conditioned of the three cross products is
    // taken and scaled to unit length, so the |v2| < 1e-10 rejection keeps the reference's meaning.  (A 3 x 3 Jacobi
    // SVD per root was most of this stage's arithmetic.)
    double c0[3], c1[3], c2[3];
    c0[0] = bz[1] * bz[5] - bz[2] * bz[4]; c0[1] = bz[2] * bz[3] - bz[0] * bz[5]; c0[2] = bz[0] * bz[4] - bz[1] * bz[3];   // r0 x r1
    c1[0] = bz[1] * bz[8] - bz[2] * bz[7]; c1[1] = bz[2] * bz[6] - bz[0] * bz[8]; c1[2] = bz[0] * bz[7] - bz[1] * bz[6];   // r0 x r2
    c2[0] = bz[4] * bz[8] - bz[5] * bz[7]; c2[1] = bz[5] * bz[6] - bz[3] * bz[8]; c2[2] = bz[3] * bz[7] - bz[4] * bz[6];   // r1 x r2
    const double n0 = c0[0] * c0[0] + c0[1] * c0[1] + c0[2] * c0[2];
    const double n1 = c1[0] * c1[0] + c1[1] * c1[1] + c1[2] * c1[2];
    const double n2 = c2[0] * c2[0] + c2[1] * c2[1] + c2[2] * c2[2];
    const double* cb = c0;
    double nb = n0;
    if (n1 > nb) { cb = c1; nb = n1; }
    if (n2 > nb) { cb = c2; nb = n2; }
    if (!(nb > 0)) return false;
    const double inb = 1 / sqrt(nb);
    const double vx = cb[0] * inb, vy = cb[1] * inb, vw = cb[2] * inb;
    if (fabs(vw) < 1e-10) return false;
    const double xs = vx / vw, ys = vy / vw;
    double nrm = 0;
    for (int k = 0; k < 9; ++k) {
        E[k] = EE[k] * xs + EE[9 + k] * ys + EE[18 + k] * z1 + EE[27 + k];
        nrm += E[k] * E[k];
    }
    nrm = sqrt(nrm);
    if (!(nrm > 0)) return false;
    for (int k = 0; k < 9; ++k) E[k] /= nrm;
    return true;
}

MCV_HDN inline int e_solve5(const double* pa, const double* pb, double* Es) {
    double EE[36], b[39], coef[11];
    if (!e5_stage1(pa, pb, EE, b, coef)) return 0;
    double roots[10];
    const int nr = real_roots10(coef, roots);
    int count = 0;
    for (int ri = 0; ri < nr; ++ri)
        if (e5_root_to_E(EE, b, roots[ri], Es + 9 * count)) ++count;
    return count;
}

// ---- decomposeEssentialMat ----------------------------------------------------------------------
MCV_HD double det3(const double* m) {
    return m[0] * (m[4] * m[8] - m[5] * m[7]) - m[1] * (m[3] * m[8] - m[5] * m[6]) + m[2] * (m[3] * m[7] - m[4] * m[6]);
}
MCV_HD void mat3mul(const double* a, const double* b, double* r) {
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) r[i * 3 + j] = a[i * 3] * b[j] + a[i * 3 + 1] * b[3 + j] + a[i * 3 + 2] * b[6 + j];
}
MCV_HDN inline void decompose_essential(const double* E, double* R1, double* R2, double* t) {
    double w[3], U[9], Vt[9];
    svd_square<3>(E, w, U, Vt);
    if (det3(U) < 0) for (int k = 0; k < 9; ++k) U[k] *= -1.;
    if (det3(Vt) < 0) for (int k = 0; k < 9; ++k) Vt[k] *= -1.;
    const double Wm[9] = {0, 1, 0, -1, 0, 0, 0, 0, 1};
    const double Wt[9] = {0, -1, 0, 1, 0, 0, 0, 0, 1};
    double T[9];
    mat3mul(U, Wm, T); mat3mul(T, Vt, R1);
    mat3mul(U, Wt, T); mat3mul(T, Vt, R2);
    t[0] = U[2]; t[1] = U[5]; t[2] = U[8];
}

// ---- recoverPose cheirality test for one correspondence and one candidate (R, t) ---------------
// DLT triangulation (cv::triangulatePoints): 4x4 system, right singular vector of the smallest
// singular value; then z*w > 0, z/w < dist, 0 < z2 < dist in the second camera.
// scr: 32 doubles of scratch with element stride S (the transposed 4 x 4 system and the accumulated rotations of its
// Jacobi SVD; 1 = plain array, blockDim = thread-interleaved shared memory).  Only the last right singular vector is used.
template <int S>
MCV_HDN inline bool cheirality_ok_at(const double* R, const double* t, double x1, double y1, double x2, double y2,
                                     double dist, double* scr) {
    double A[16];
    // P0 = [I | 0], P1 = [R | t]
    A[0] = x1 * 0 - 1; A[1] = x1 * 0 - 0; A[2] = x1 * 1 - 0; A[3] = x1 * 0 - 0;
    A[4] = y1 * 0 - 0; A[5] = y1 * 0 - 1; A[6] = y1 * 1 - 0; A[7] = y1 * 0 - 0;
    for (int k = 0; k < 3; ++k) {
        A[8 + k] = x2 * R[6 + k] - R[k];
        A[12 + k] = y2 * R[6 + k] - R[3 + k];
    }
    A[11] = x2 * t[2] - t[0];
    A[15] = y2 * t[2] - t[1];
    double* At = scr;
    double* Vt = scr + 16 * S;
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) At[(i * 4 + j) * S] = A[j * 4 + i];   // rows of At = columns of A (svd_square)
    double w[4];
    jacobi_svd_rows<4, 4, 4, true, S>(At, w, Vt);
    double Q0 = Vt[12 * S], Q1 = Vt[13 * S], Q2 = Vt[14 * S], Q3 = Vt[15 * S];
    bool ok = (Q2 * Q3) > 0;
    Q0 /= Q3; Q1 /= Q3; Q2 /= Q3; Q3 /= Q3;
    ok = ok && (Q2 < dist);
    const double z2 = R[6] * Q0 + R[7] * Q1 + R[8] * Q2 + t[2] * Q3;
    ok = ok && (z2 > 0) && (z2 < dist);
    return ok;
}
MCV_HDN inline bool cheirality_ok(const double* R, const double* t, double x1, double y1, double x2,
                                  double y2, double dist) {
    double scr[32];
    return cheirality_ok_at<1>(R, t, x1, y1, x2, y2, dist, scr);
}

}  // namespace cvm
}  // namespace mcv
