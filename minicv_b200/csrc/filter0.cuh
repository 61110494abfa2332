// filter0.cuh -- the arithmetic of the tensor-core first-level filter of the H / F / E scoring stage (score_tc.cu), as
// __host__ __device__ functions so that the host build (tests/hostmath) can hammer the bound in the CPU suite.
//
// The residual of the epipolar constraint, r = x2' F x1, is bilinear: r = f . psi with the nine model entries f and the
// nine monomials psi = (x2 x1, x2 y1, x2, y2 x1, y2 y1, y2, x1, y1, 1) of a correspondence.  For all (model,
// correspondence) pairs of an image pair that is a K = 9 GEMM, and the tensor cores evaluate it at a fraction of what the
// fp32 filter of cvmath.cuh costs per evaluation.  What the epilogue has to do per evaluation is ONE comparison
// |r~| > T against a threshold that is constant per (model, stage of 64 correspondences), because
//   * coordinates are recentred and rescaled per image pair (u = (x - c) k, both images, |u| <= 1) and the model is
//     carried into those coordinates in fp64 (f^ below: exactly the same bilinear form, r = f^ . psi(u));
//   * both operands are split into two TF32 pieces (x = hi + lo) and the GEMM accumulates hi.hi + hi.lo + lo.hi
//     (K = 27, padded to 32): r~ reproduces r to ~2^-20 of |f^| |psi| -- a plain TF32 product would be off by as much
//     as the inlier threshold itself (2^-10 x image extent ~ 1.5 px);
//   * the correspondences of a pair are sorted by their image-1 cell (counting sort over a 32 x 32 Morton grid), so the
//     64 correspondences of a stage sit in a small box (centre cu, cv, half extents pu, pv) of image 1 and the norm of
//     the epipolar line (a, b) = first two rows of F applied to x1 is bounded over the box:
//         sqrt(a^2 + b^2) <= k B,   B = hypot(A0 . c, A1 . c) + hypot(|f^0| pu + |f^1| pv, |f^3| pu + |f^4| pv)
//     (A0, A1 = first two rows of f^, c = (cu, cv, 1)); for the Sampson denominator of E the image-2 half is bounded by
//     Cauchy-Schwarz over the whole pair: a'^2 + b'^2 <= k^2 G' V2max, G' = squares of the first two columns of f^.
// "Certainly an outlier" (sufficient, in real arithmetic): |r| > sqrt(thr) k B  [E: sqrt(thr) k sqrt(B^2 + G' V2max)].
// Everything r~ and the reference's own fp64 evaluation can get wrong is added to the right-hand side:
//   e_m  = 2^-16 |f^| PsiMax + 2^-48 S PsiMax          split-TF32 representation (|x - hi - lo| <= 2^-22 |x|, the dropped
//                                                      lo.lo term 2^-22, fp32 view of the operands 2^-23), the tensor
//                                                      core's accumulation (budgeted at 2^-18 of the sum of absolute terms:
//                                                      32 truncated fp32 additions), all together < 2^-17, doubled; the
//                                                      fp64 rounding of f^ (S = sum of the absolute terms it is built
//                                                      from); PsiMax >= |psi| over the pair, sum |f^_k psi_k| <= |f^| |psi|
//   n_m  = 8.1 u D + sqrt(thr) 3 u hypot(A, B [, A', B'])   the reference's fp64 roundings (u = 2^-53) of the residual and
//                                                      of the line coefficients, from the per-pair coordinate maxima
//                                                      in ORIGINAL coordinates (A = |F0| X1 + |F1| Y1 + |F2| ...), so
//                                                      that a far-off-centre image pair, whose exact test is itself
//                                                      noisy, widens the threshold instead of breaking the proof
//   T    = stk B~ + base,  stk = sqrt(thr (1 + 2^-19)) k (1 + 2^-17),  base = [stk slack + n_m + e_m] (1 + 2^-20) + 1e-30
// (B~ = B evaluated in fp32 from the absolute fp32 roundings of f^; slack = 2^-20 (|f^0| + .. + |f^5|) covers those
// roundings) and the proof obligation is:  |r~| > T  implies  f_inlier / e_inlier (cvmath.cuh) return false on every
// path (the fast path's q > th n, and the division path: the computed ratio exceeds thr by > 2^-20 relative, n = 0 gives
// inf / NaN, overflow is excluded by the guards below).  A NaN anywhere makes the comparison fail, i.e. the evaluation
// goes on to the fp32 classifier and the exact test like everything level 0 could not reject.  Level 0 is switched off
// (T = inf) for models with non-finite, huge (> 1e30) or tiny (< 1e-30, non-zero) entries and for pairs without finite
// extent or with a non-finite coordinate.
#pragma once
#include "cvmath.cuh"

namespace mcv {
namespace cvm {

constexpr int F0_K = 32;                  // GEMM K: 3 x 9 split terms, zero padded to four TF32 MMAs

struct PairNorm {                         // per image pair, written by the feature kernel
    double c1x, c1y, c2x, c2y, k;         // u1 = (x1 - c1x) k, v1 = (y1 - c1y) k, u2 = (x2 - c2x) k, v2 = (y2 - c2y) k
    float  psiMax;                        // >= |psi| of every correspondence of the pair
    float  v2Max;                         // >= u2^2 + v2^2 + 1
    float  X1, Y1, X2, Y2;                // max |coordinate| in original coordinates (fp32 view; a NaN sticks)
    int    on;                            // 0: no usable extent, level 0 off for the pair
};

struct StageBox { float cu, cv, pu, pv; };

// The classifier uses its own mu = 2^-6 in (p + q)^2 <= (1 + mu) p^2 + (1 + 1/mu) q^2: the additive constants shrink by
// 64 against the outlier filter's mu = 2^-12 (near an epipole, where the line normal is small against its error bound,
// they decide whether anything can be classified at all) and the price is a relative band of ~3 % around the threshold
// that goes to the exact test.
constexpr float CLASS_MU1 = 1.015625f;            // 1 + mu
constexpr float CLASS_MUM = 0.984375f;            // 1 - mu
constexpr float CLASS_SQ = 8.07f;                 // sqrt(1 + 1/mu) = 8.062.., rounded up
constexpr float CLASS_MARGIN = 1.001953125f * 1.015625f;     // (1 + 2^-9)(1 + mu)
constexpr float CLASS_LOW = 0.998046875f;         // 1 - 2^-9
   // image-1 box of one stage in the pair's coordinates

// per-model constants: the model in the pair's coordinates (fp32 roundings of f^), the stage threshold T = stk * bound +
// base, and the additive constants of the two-sided classifier in those coordinates (hat_classify below)
struct Model0 {
    float g[9];
    float stk, base;
    float g2;         // E: G' V2max (0 for F)
    float K2, K1, KD;
};

// round-to-nearest (ties away) to TF32's 10 mantissa bits -- what cvt.rna.tf32.f32 does; integer form so that the host
// build rounds identically
MCV_HD float tf32_round(float x) {
    union { float f; uint32_t u; } v;
    v.f = x;
    if ((v.u & 0x7F800000u) == 0x7F800000u) return x;
    v.u = (v.u + 0x1000u) & 0xFFFFE000u;
    return v.f;
}
MCV_HD void tf32_split(float x, float& hi, float& lo) {
    hi = tf32_round(x);
    lo = tf32_round(x - hi);              // x - hi is exact in fp32
}

// centre / scale from the finite coordinate ranges (lo / hi per axis; lo > hi: no finite value seen)
MCV_HD void pair_norm_init(const double* lo, const double* hi, PairNorm& pn) {
    double ext = 0;
    bool any = true;
    double c[4];
    for (int i = 0; i < 4; ++i) {
        if (!(lo[i] <= hi[i])) { any = false; c[i] = 0; continue; }
        c[i] = 0.5 * lo[i] + 0.5 * hi[i];
        const double e = fmax(hi[i] - c[i], c[i] - lo[i]);
        ext = fmax(ext, e);
    }
    pn.c1x = c[0]; pn.c1y = c[1]; pn.c2x = c[2]; pn.c2y = c[3];
    const bool ok = any && ext > 1e-150 && ext < 1e150;
    pn.k = ok ? 1.0 / ext : 1.0;
    pn.on = ok ? 1 : 0;
}

struct Hat { double u1, v1, u2, v2; };
MCV_HD Hat hat_coords(const PairNorm& pn, double x1, double y1, double x2, double y2) {
    Hat h;
    h.u1 = (x1 - pn.c1x) * pn.k; h.v1 = (y1 - pn.c1y) * pn.k;
    h.u2 = (x2 - pn.c2x) * pn.k; h.v2 = (y2 - pn.c2y) * pn.k;
    return h;
}
// 32 x 32 Morton cell of the image-1 position (|u| <= 1 for finite coordinates; anything else lands in a border cell)
MCV_HD int hat_cell(const Hat& h) {
    const double fx = (h.u1 + 1.0) * 16.0, fy = (h.v1 + 1.0) * 16.0;
    const int ix = (fx >= 0.0 && fx < 32.0) ? (int)fx : (fx >= 32.0 ? 31 : 0);       // NaN -> 0
    const int iy = (fy >= 0.0 && fy < 32.0) ? (int)fy : (fy >= 32.0 ? 31 : 0);
    int c = 0;
    for (int b = 0; b < 5; ++b) c |= (((ix >> b) & 1) << (2 * b)) | (((iy >> b) & 1) << (2 * b + 1));
    return c;
}

// one correspondence -> its operand row [psi_hi (9) | psi_lo (9) | psi_hi (9) | 0 (5)]; returns |psi|^2; v2 = V2
MCV_HD double filter0_features(const Hat& h, float* row, double* v2) {
    const double psi[9] = {h.u2 * h.u1, h.u2 * h.v1, h.u2, h.v2 * h.u1, h.v2 * h.v1, h.v2, h.u1, h.v1, 1.0};
    for (int i = 0; i < 9; ++i) {
        float hi, lo;
        tf32_split((float)psi[i], hi, lo);
        row[i] = hi; row[9 + i] = lo; row[18 + i] = hi;
    }
    for (int i = 27; i < F0_K; ++i) row[i] = 0.f;
    const double V1 = h.u1 * h.u1 + h.v1 * h.v1 + 1.0, V2 = h.u2 * h.u2 + h.v2 * h.v2 + 1.0;
    *v2 = V2;
    return V1 * V2;
}

// one model -> its operand row [f^_hi | f^_hi | f^_lo | 0] in the pair's coordinates and the constants of the stage
// threshold (m0.base = inf: level 0 off for the model).  fhOut (9, optional): f^ in fp64; emOut (optional): the
// arithmetic part e_m of the threshold (audit builds)
MCV_HD void filter0_model(int isE, const PairNorm& pn, const double* F, float thr, float* row, Model0& m0,
                          double* fhOut = nullptr, double* emOut = nullptr) {
    const double k = pn.k, ik = 1.0 / k, ik2 = ik * ik;
    const double c1x = pn.c1x, c1y = pn.c1y, c2x = pn.c2x, c2y = pn.c2y;
    bool on = pn.on != 0 && thr >= 0.f && thr <= 1e30f;
    for (int i = 0; i < 9; ++i) {
        const double v = fabs(F[i]);
        if (!(v <= 1e30) || (v != 0.0 && v < 1e-30)) on = false;
    }
    // rows of F applied to the image-1 centre, columns applied to the image-2 centre (with the sums of absolute terms)
    const double r0 = F[0] * c1x + F[1] * c1y + F[2], r1 = F[3] * c1x + F[4] * c1y + F[5], r2 = F[6] * c1x + F[7] * c1y + F[8];
    const double s0 = fabs(F[0] * c1x) + fabs(F[1] * c1y) + fabs(F[2]), s1 = fabs(F[3] * c1x) + fabs(F[4] * c1y) + fabs(F[5]);
    const double s2 = fabs(F[6] * c1x) + fabs(F[7] * c1y) + fabs(F[8]);
    const double q0 = F[0] * c2x + F[3] * c2y + F[6], q1 = F[1] * c2x + F[4] * c2y + F[7];
    const double t0 = fabs(F[0] * c2x) + fabs(F[3] * c2y) + fabs(F[6]), t1 = fabs(F[1] * c2x) + fabs(F[4] * c2y) + fabs(F[7]);
    double fh[9];
    fh[0] = F[0] * ik2; fh[1] = F[1] * ik2; fh[2] = r0 * ik;
    fh[3] = F[3] * ik2; fh[4] = F[4] * ik2; fh[5] = r1 * ik;
    fh[6] = q0 * ik;    fh[7] = q1 * ik;    fh[8] = c2x * r0 + c2y * r1 + r2;
    const double S = (fabs(F[0]) + fabs(F[1]) + fabs(F[3]) + fabs(F[4])) * ik2 + (s0 + s1 + t0 + t1) * ik +
                     fabs(c2x) * s0 + fabs(c2y) * s1 + s2;
    double nf = 0;
    for (int i = 0; i < 9; ++i) {
        nf += fh[i] * fh[i];
        float hi, lo;
        tf32_split((float)fh[i], hi, lo);
        row[i] = hi; row[9 + i] = hi; row[18 + i] = lo;
    }
    for (int i = 27; i < F0_K; ++i) row[i] = 0.f;
    nf = sqrt(nf);
    const double Gt = fh[0] * fh[0] + fh[3] * fh[3] + fh[6] * fh[6] + fh[1] * fh[1] + fh[4] * fh[4] + fh[7] * fh[7];
    // the reference's own fp64 roundings, from the coordinate maxima in original coordinates
    const double up = 1.0 + 1e-6;                   // the maxima are taken over the fp32 view of the coordinates
    const double X1 = pn.X1 * up, Y1 = pn.Y1 * up, X2 = pn.X2 * up, Y2 = pn.Y2 * up;
    const double A = fabs(F[0]) * X1 + fabs(F[1]) * Y1 + fabs(F[2]), B = fabs(F[3]) * X1 + fabs(F[4]) * Y1 + fabs(F[5]);
    const double Cc = fabs(F[6]) * X1 + fabs(F[7]) * Y1 + fabs(F[8]);
    const double D = X2 * A + Y2 * B + Cc;
    const double At = fabs(F[0]) * X2 + fabs(F[3]) * Y2 + fabs(F[6]), Bt = fabs(F[1]) * X2 + fabs(F[4]) * Y2 + fabs(F[7]);
    const double u = 1.1102230246251565e-16;                                   // 2^-53
    const double hyp = isE ? sqrt(A * A + B * B + At * At + Bt * Bt) : sqrt(A * A + B * B);
    const double st = sqrt((double)thr * (1.0 + 1.9073486328125e-06));         // thr (1 + 2^-19)
    const double psiMax = pn.psiMax;
    const double em = 1.52587890625e-05 * nf * psiMax + 3.5527136788005009e-15 * S * psiMax;    // 2^-16, 2^-48
    const double stk = st * k * (1.0 + 7.62939453125e-06);                     // 1 + 2^-17
    const double slack = 9.5367431640625e-07 * (fabs(fh[0]) + fabs(fh[1]) + fabs(fh[2]) + fabs(fh[3]) + fabs(fh[4]) + fabs(fh[5]));
    const double base = (stk * slack + st * 3.0 * u * hyp + 8.1 * u * D + em) * (1.0 + 9.5367431640625e-07) + 1e-30;
    const double g2 = isE ? Gt * (double)pn.v2Max * (1.0 + 1e-6) : 0.0;
    if (fhOut) for (int i = 0; i < 9; ++i) fhOut[i] = fh[i];
    if (emOut) *emOut = em;
    if (!(D < 1e100) || !(hyp < 1e100) || !(S < 1e100) || !(base < 1e30) || !(stk < 1e30) || !(g2 < 1e30) || !(nf < 1e30)) on = false;   // NaN too
    for (int i = 0; i < 9; ++i) m0.g[i] = (float)fh[i];
    // classifier constants in the pair's coordinates (|u|, |v| <= UM): the fp32 evaluation (8u / 16u of the absolute sums,
    // as in cvmath.cuh), the fp64 rounding of f^ (2^-48 S), and the reference's own fp64 roundings carried over from
    // original coordinates (a = k a^: 3 u64 A / k; the residual is the same number in both: 8.1 u64 D)
    {
        const double UM = 1.000001, u32 = 5.9604644775390625e-08, s48 = 3.5527136788005009e-15 * S;
        const double Ah = (fabs(fh[0]) + fabs(fh[1])) * UM + fabs(fh[2]), Bh = (fabs(fh[3]) + fabs(fh[4])) * UM + fabs(fh[5]);
        const double Ch = (fabs(fh[6]) + fabs(fh[7])) * UM + fabs(fh[8]), Dh = UM * (Ah + Bh) + Ch;
        const double T0h = (fabs(fh[0]) + fabs(fh[3])) * UM + fabs(fh[6]), T1h = (fabs(fh[1]) + fabs(fh[4])) * UM + fabs(fh[7]);
        const double eA = 8.0 * u32 * Ah + s48 + 3.0 * u * A * ik, eB = 8.0 * u32 * Bh + s48 + 3.0 * u * B * ik;
        const double e0 = 8.0 * u32 * T0h + s48 + 3.0 * u * At * ik, e1 = 8.0 * u32 * T1h + s48 + 3.0 * u * Bt * ik;
        const double eD = 16.0 * u32 * Dh + 3.0 * s48 + 8.1 * u * D;
        const double sq = (double)CLASS_SQ * (double)CLASS_SQ * (1.0 + 1e-6);
        const double K2 = sq * (eA * eA + eB * eB) + 1e-30, K1 = sq * (e0 * e0 + e1 * e1) + 1e-30, KD = sq * eD * eD + 1e-30;
        const double thrk = (double)thr * k * k;
        // nothing the classifier computes may leave fp32's range (an infinite line norm would pass "q <= thr n"): every
        // fp32 quantity is bounded by these sums
        const double nmax = Ah * Ah + Bh * Bh + T0h * T0h + T1h * T1h, qmax = Dh * Dh;
        const bool con = on && K2 < 1e30 && K1 < 1e30 && KD < 1e30 && thrk > 1e-30 && thrk < 1e30 && nmax < 1e30 && qmax < 1e30 &&
                         thrk * 1.1 * (nmax + K2 + K1) < 1e37;
        // off: all three infinite, so that both tests fail whatever the evaluation overflows to
        m0.K2 = con ? (float)(K2 * (1.0 + 1.2e-7)) : INFINITY;
        m0.K1 = con ? (float)(K1 * (1.0 + 1.2e-7)) : INFINITY;
        m0.KD = con ? (float)(KD * (1.0 + 1.2e-7)) : INFINITY;
    }
    m0.stk = (float)(stk * (1.0 + 1.2e-7));
    m0.base = on ? (float)(base * (1.0 + 1.2e-7)) : INFINITY;
    m0.g2 = (float)(g2 * (1.0 + 1.2e-7));
}

// threshold of one (model, stage): fp32, every rounding inside the 2^-17 / slack margins of stk and base (the device
// build takes sqrt.approx, 2^-22 relative)
MCV_HD float f0_sqrt(float x) {
#if defined(__CUDA_ARCH__)
    float r;
    asm("sqrt.approx.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
#else
    return sqrtf(x);
#endif
}
MCV_HD float filter0_threshold(int isE, const Model0& m0, const StageBox& b) {
    const float a0 = fmaf(m0.g[0], b.cu, fmaf(m0.g[1], b.cv, m0.g[2]));
    const float a1 = fmaf(m0.g[3], b.cu, fmaf(m0.g[4], b.cv, m0.g[5]));
    const float e0 = fmaf(fabsf(m0.g[0]), b.pu, fabsf(m0.g[1]) * b.pv);
    const float e1 = fmaf(fabsf(m0.g[3]), b.pu, fabsf(m0.g[4]) * b.pv);
    const float bound = f0_sqrt(fmaf(a0, a0, a1 * a1)) + f0_sqrt(fmaf(e0, e0, e1 * e1));
    const float tot = isE ? f0_sqrt(fmaf(bound, bound, m0.g2)) : bound;
    return fmaf(m0.stk, tot, m0.base);
}

// ---- homography ------------------------------------------------------------------------------------------------------
// The reprojection residual g = (X - u w, Y - v w), (X, Y, w) = H (x, y, 1), is bilinear in (H, correspondence) as well, and
// the reference's error is |g|^2 / w^2 (h_error, float32).  One accumulator per evaluation carries the component of g along
// the diagonal, s = k (gx + gy) = f^ . psi with psi = (u^, v^, 1, -t u^, -t v^, -t), t = p^ + q^ (recentred source (u^, v^) and
// target (p^, q^)), f^ = (h^0 + h^3, h^1 + h^4, h^2 + h^5, h^6, h^7, h^8), h^ = T2 H T1^-1 (so g^ = k g, w^ = w).  |s| <=
// sqrt(2) k |g|, hence  |s| > sqrt(2) k R  proves |g| > R.  An outlier whose residual happens to be perpendicular to the
// diagonal is not caught (a few per cent); like every survivor it gets the exact test.  R comes from the float32 path's own
// roundings (cvmath.cuh, h_certain_outlier: with Xb = |h0| Xm + |h1| Ym + |h2|, Wb = |h6| Xm + |h7| Ym + 1 over the pair's
// coordinate maxima, the float32 X, Y, w are within 3u Xb, 3u Yb, 3u Wb of the real values and
// |d_f| >= |g_e / w_f| (1 - 4.1u) - 3.1u (Um, Vm) componentwise), which give, with a = 4u (Xb + Um Wb, Yb + Vm Wb), aw = 4u Wb,
//     err_f > thr   if   |g| > [sqrt(thr) (1 + 2u) + 3.1u |(Um, Vm)|] (|w| + aw) (1 + 5u) + |a|
// (a zero, infinite or NaN float32 w makes h_error inf / NaN, an outlier as well).  |w| <= |w^ . c| + |h^6| pu + |h^7| pv over
// the stage's source box.  H here is the float32 matrix the exact test uses (entries 0..7 rounded, h33 = 1).
MCV_HD double filter0_features_h(const Hat& h, float* row) {
    const double t = h.u2 + h.v2;
    const double psi[6] = {h.u1, h.v1, 1.0, -t * h.u1, -t * h.v1, -t};
    double n2 = 0;
    for (int i = 0; i < 6; ++i) {
        float hi, lo;
        tf32_split((float)psi[i], hi, lo);
        row[i] = hi; row[6 + i] = lo; row[12 + i] = hi;
        n2 += psi[i] * psi[i];
    }
    for (int i = 18; i < F0_K; ++i) row[i] = 0.f;
    return n2;
}
// m0.g[0..5] = f^ (audit), g[6..8] = h^6, h^7, h^8 (stage threshold), stk, base; K2, K1, KD, g2 unused
MCV_HD void filter0_model_h(const PairNorm& pn, const float* Hf, float thr, float* row, Model0& m0, double* fhOut = nullptr,
                            double* emOut = nullptr) {
    const double k = pn.k, ik = 1.0 / k;
    const double c1x = pn.c1x, c1y = pn.c1y, c2x = pn.c2x, c2y = pn.c2y;
    bool on = pn.on != 0 && thr >= 0.f && thr <= 1e30f;
    double h[9];
    for (int i = 0; i < 8; ++i) {
        h[i] = (double)Hf[i];
        const double v = fabs(h[i]);
        if (!(v <= 1e30) || (v != 0.0 && v < 1e-30)) on = false;
    }
    h[8] = 1.0;
    const double r0 = h[0] * c1x + h[1] * c1y + h[2], r1 = h[3] * c1x + h[4] * c1y + h[5], rw = h[6] * c1x + h[7] * c1y + 1.0;
    const double s0 = fabs(h[0] * c1x) + fabs(h[1] * c1y) + fabs(h[2]), s1 = fabs(h[3] * c1x) + fabs(h[4] * c1y) + fabs(h[5]);
    const double sw = fabs(h[6] * c1x) + fabs(h[7] * c1y) + 1.0;
    double hh[9];
    hh[0] = h[0] - c2x * h[6]; hh[1] = h[1] - c2x * h[7]; hh[2] = k * (r0 - c2x * rw);
    hh[3] = h[3] - c2y * h[6]; hh[4] = h[4] - c2y * h[7]; hh[5] = k * (r1 - c2y * rw);
    hh[6] = h[6] * ik;         hh[7] = h[7] * ik;         hh[8] = rw;
    // sum of the absolute terms every entry is built from (fp64 rounding of the transform)
    const double S = fabs(h[0]) + fabs(h[1]) + fabs(h[3]) + fabs(h[4]) + (fabs(c2x) + fabs(c2y)) * (fabs(h[6]) + fabs(h[7])) +
                     k * (s0 + s1 + (fabs(c2x) + fabs(c2y)) * sw) + (fabs(h[6]) + fabs(h[7])) * ik + sw;
    const double fh[6] = {hh[0] + hh[3], hh[1] + hh[4], hh[2] + hh[5], hh[6], hh[7], hh[8]};
    double nf = 0;
    for (int i = 0; i < 6; ++i) {
        nf += fh[i] * fh[i];
        float hi, lo;
        tf32_split((float)fh[i], hi, lo);
        row[i] = hi; row[6 + i] = hi; row[12 + i] = lo;
    }
    for (int i = 18; i < F0_K; ++i) row[i] = 0.f;
    nf = sqrt(nf);
    const double up = 1.0 + 1e-6, u32 = 5.9604644775390625e-08;
    const double Xm = pn.X1 * up, Ym = pn.Y1 * up, Um = pn.X2 * up, Vm = pn.Y2 * up;
    const double Xb = fabs(h[0]) * Xm + fabs(h[1]) * Ym + fabs(h[2]), Yb = fabs(h[3]) * Xm + fabs(h[4]) * Ym + fabs(h[5]);
    const double Wb = fabs(h[6]) * Xm + fabs(h[7]) * Ym + 1.0;
    const double ax = 4.0 * u32 * (Xb + Um * Wb), ay = 4.0 * u32 * (Yb + Vm * Wb), aw = 4.0 * u32 * Wb;
    const double st = sqrt((double)thr) * (1.0 + 2.0 * u32) + 3.1 * u32 * sqrt(Um * Um + Vm * Vm);
    const double psiMax = pn.psiMax;
    const double em = 1.52587890625e-05 * nf * psiMax + 3.5527136788005009e-15 * S * psiMax;    // 2^-16, 2^-48
    const double r2k = 1.4142135623730951 * k;
    const double stk = r2k * st * (1.0 + 5.0 * u32) * (1.0 + 7.62939453125e-06);                // (1 + 2^-17)
    const double slack = 9.5367431640625e-07 * (fabs(hh[6]) + fabs(hh[7]) + fabs(hh[8])) + 3.5527136788005009e-15 * S;
    const double base = (stk * (aw + slack) + r2k * sqrt(ax * ax + ay * ay) + em) * (1.0 + 9.5367431640625e-07) + 1e-30;
    if (fhOut) { for (int i = 0; i < 6; ++i) fhOut[i] = fh[i]; fhOut[6] = fhOut[7] = fhOut[8] = 0; }
    if (emOut) *emOut = em;
    if (!(Xb < 1e100) || !(Yb < 1e100) || !(Wb < 1e100) || !(S < 1e100) || !(base < 1e30) || !(stk < 1e30) || !(nf < 1e30) ||
        !(Um < 1e100) || !(Vm < 1e100))
        on = false;                                                                              // NaN too
    for (int i = 0; i < 6; ++i) m0.g[i] = (float)fh[i];
    m0.g[6] = (float)hh[6]; m0.g[7] = (float)hh[7]; m0.g[8] = (float)hh[8];
    m0.stk = (float)(stk * (1.0 + 1.2e-7));
    m0.base = on ? (float)(base * (1.0 + 1.2e-7)) : INFINITY;
    m0.g2 = 0.f; m0.K2 = m0.K1 = m0.KD = INFINITY;
}
MCV_HD float filter0_threshold_h(const Model0& m0, const StageBox& b) {
    const float w0 = fabsf(fmaf(m0.g[6], b.cu, fmaf(m0.g[7], b.cv, m0.g[8])));
    const float e0 = fmaf(fabsf(m0.g[6]), b.pu, fabsf(m0.g[7]) * b.pv);
    return fmaf(m0.stk, w0 + e0, m0.base);
}

// ---- two-sided fp32 classifier in front of f_inlier / e_inlier (survivors of level 0) ----------------------------------
// The quantities of f_certain_outlier / e_certain_outlier (cvmath.cuh) -- residual d, line norms n2 = a^2 + b^2 (line F x1
// in image 2) and n1 (line F' x2 in image 1) -- evaluated in fp32 for BOTH epipolar lines, in the pair's coordinates, and
// used in both directions: with |d~ - d| <= eD, |a~ - a| <= eA ... (the fp32 evaluation, the fp64 rounding of f^, AND the
// reference's own fp64 path are all inside these bounds, see filter0_model),
//   d^2 <= (1 + mu) d~^2 + KD   and   n >= (1 - mu) n~ - K      [ (p - q)^2 >= (1 - mu) p^2 - (1/mu - 1) q^2 ]
// so  (1 + mu) d~^2 + KD <= thr (1 - 2^-9) [(1 - mu) n~ - K]  for both lines proves that the true ratios are below thr
// by 2^-9 relative: f_inlier's fast path (q <= tl n, n > 0 finite) returns true.  Outlier side, as in cvmath.cuh:
//   d~^2 > thr (1 + 2^-9)(1 + mu) [(1 + mu) n~ + K] + KD  for either line (E: for the sum of both).
// 0 = certainly an outlier, 1 = certainly an inlier, 2 = undecided (NaN, near the threshold, near an epipole): exact test.
// g = fp32 f^, (u1, v1, u2, v2) = fp32 roundings of the
// recentred coordinates, thresholds thr k^2 (a = k a^ while the residual is the same number in both coordinate systems).
// thrkM = thr k^2 CLASS_MARGIN, thrkL = thr k^2 CLASS_LOW.
MCV_HD int hat_classify(int isE, const Model0& m, float thrkM, float thrkL, float u1, float v1, float u2, float v2) {
    const float a = fmaf(m.g[0], u1, fmaf(m.g[1], v1, m.g[2]));
    const float b = fmaf(m.g[3], u1, fmaf(m.g[4], v1, m.g[5]));
    const float c = fmaf(m.g[6], u1, fmaf(m.g[7], v1, m.g[8]));
    const float d = fmaf(u2, a, fmaf(v2, b, c));
    const float t0 = fmaf(m.g[0], u2, fmaf(m.g[3], v2, m.g[6]));
    const float t1 = fmaf(m.g[1], u2, fmaf(m.g[4], v2, m.g[7]));
    const float n2 = fmaf(a, a, b * b), n1 = fmaf(t0, t0, t1 * t1);
    const float q = d * d;
    const float qu = fmaf(q, CLASS_MU1, m.KD);
    bool out, in;                                   // no short-circuit: straight-line code, one select at the end
    if (isE) {
        const float n = n2 + n1, K = m.K2 + m.K1;
        out = q > fmaf(thrkM, fmaf(n, CLASS_MU1, K), m.KD);
        in = qu <= thrkL * fmaf(n, CLASS_MUM, -K);
    } else {
        out = (q > fmaf(thrkM, fmaf(n2, CLASS_MU1, m.K2), m.KD)) | (q > fmaf(thrkM, fmaf(n1, CLASS_MU1, m.K1), m.KD));
        in = (qu <= thrkL * fmaf(n2, CLASS_MUM, -m.K2)) & (qu <= thrkL * fmaf(n1, CLASS_MUM, -m.K1));
    }
    return out ? 0 : (in ? 1 : 2);
}

}  // namespace cvm
}  // namespace mcv
