// pnpmath.cuh -- building blocks of the PnP-RANSAC row, written as
// __host__ __device__ code like cvmath.cuh so that the same source is unit-tested on the host
// (tests/hostmath) and runs inside the CUDA kernels of pnp.cu.
//
// What each block follows:
//   * ap3p_*            the reference's vendored AP3P solver, src/MiniCVNative/ap3p.cpp:9-280 (Ke &
//                       Roumeliotis 2017); `solveAp3p` (ap3p.cpp:282-317) is ap3p_bearings + ap3p_poses.
//                       The 4-point form (fourth-point disambiguation, complex root pairs dropped) is what
//                       cv::solvePnP(SOLVEPNP_AP3P) of cv2 4.13.0 does inside cv::solvePnPRansac, the call
//                       behind cvSolvePnPRansac (src/MiniCVNative/MiniCVNative.cpp:93-139).
//   * rodrigues_*       cv::Rodrigues in both directions
//   * undistort_point / project_error   cv::undistortPoints / cv::projectPoints + PnPRansacCallback::computeError
//   * epnp_*            OpenCV's EPnP (calib3d epnp.cpp), split into the per-point pieces (parallel over
//                       points in the kernel, plain loops on the host) and the small dense-algebra stages
#pragma once
#include "cvmath.cuh"

namespace mcv {
namespace cvm {

// ---- AP3P ------------------------------------------------------------------------------------
// Re(sqrt(x + i y))
MCV_HD double csqrt_re(double x, double y) {
    if (y == 0) return x >= 0 ? sqrt(x) : 0.0;
    const double h = hypot(x, y);
    if (x >= 0) return sqrt((h + x) * 0.5);
    return fabs(y) / (2 * sqrt((h - x) * 0.5));
}

// ap3p.cpp:9-62: closed-form real parts of the four roots; returns false when the arithmetic broke down
MCV_HDN inline bool ap3p_quartic(const double* f, double* roots) {
    const double a4 = f[0], a3 = f[1], a2 = f[2], a1 = f[3], a0 = f[4];
    const double a4_2 = a4 * a4, a3_2 = a3 * a3, a4_3 = a4_2 * a4, a2a4 = a2 * a4;
    const double p4 = (8 * a2a4 - 3 * a3_2) / (8 * a4_2);
    const double q4 = (a3_2 * a3 - 4 * a2a4 * a3 + 8 * a1 * a4_2) / (8 * a4_3);
    const double r4 = (256 * a0 * a4_3 - 3 * (a3_2 * a3_2) - 64 * a1 * a3 * a4_2 + 16 * a2a4 * a3_2) / (256 * (a4_3 * a4));
    const double p3 = ((p4 * p4) / 12 + r4) / 3;
    const double q3 = (72 * r4 * p4 - 2 * p4 * p4 * p4 - 27 * q4 * q4) / 432;
    const double disc = q3 * q3 - p3 * p3 * p3;
    double t;
    if (disc >= 0) {
        // w is real
        const double sd = sqrt(disc);
        double wr = (q3 >= 0) ? (-sd - q3) : (sd - q3);
        wr = cbrt(wr);
        t = 2.0 * (wr + p3 / wr);
    } else {
        // w = -+ i sqrt(-disc) - q3, principal cube root
        const double si = sqrt(-disc);
        const double wi = (q3 >= 0) ? -si : si;
        const double wr = -q3;
        const double mag = exp(log(hypot(wr, wi)) * (1.0 / 3));
        t = 4.0 * mag * cos(atan2(wi, wr) * (1.0 / 3));
    }
    const double m2 = -2 * p4 / 3 + t;           // sqrt_2m = sqrt(complex(m2, 0))
    const double B_4A = -a3 / (4 * a4);
    const double complex1 = 4 * p4 / 3 + t;
    double sqrt_2m_rh, sqrt1, sqrt2;
    if (m2 >= 0) {
        const double s2m = sqrt(m2);
        const double c2 = 2 * q4 / s2m;          // real
        sqrt_2m_rh = s2m / 2;
        sqrt1 = csqrt_re(-(complex1 + c2), 0.0) / 2;
        sqrt2 = csqrt_re(-(complex1 - c2), 0.0) / 2;
    } else {
        const double s2m = sqrt(-m2);            // sqrt_2m = i s2m, 2 q4 / (i s2m) = -i 2 q4 / s2m
        const double c2i = -2 * q4 / s2m;
        sqrt_2m_rh = 0;
        sqrt1 = csqrt_re(-complex1, -c2i) / 2;
        sqrt2 = csqrt_re(-complex1, c2i) / 2;
    }
    roots[0] = B_4A + sqrt_2m_rh + sqrt1;
    roots[1] = B_4A + sqrt_2m_rh - sqrt1;
    roots[2] = B_4A - sqrt_2m_rh + sqrt2;
    roots[3] = B_4A - sqrt_2m_rh - sqrt2;
    return true;
}

MCV_HD void v3cross(const double* a, const double* b, double* r) {
    r[0] = a[1] * b[2] - a[2] * b[1];
    r[1] = -(a[0] * b[2] - a[2] * b[0]);
    r[2] = a[0] * b[1] - a[1] * b[0];
}
MCV_HD double v3dot(const double* a, const double* b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
MCV_HD double v3norm(const double* a) { return sqrt(a[0] * a[0] + a[1] * a[1] + a[2] * a[2]); }
MCV_HD void m3mul(const double* a, const double* b, double* r) {
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) r[3 * i + j] = a[3 * i] * b[j] + a[3 * i + 1] * b[3 + j] + a[3 * i + 2] * b[6 + j];
}

// pixel -> unit bearing (ap3p.cpp:288-309); returns (mu, mv) before normalisation in muv
MCV_HD void ap3p_bearing(double u, double v, double inv_fx, double inv_fy, double cx_fx, double cy_fy, double* b,
                         double* muv) {
    const double mu = inv_fx * u - cx_fx, mv = inv_fy * v - cy_fy;
    const double mk = 1. / sqrt(mu * mu + mv * mv + 1);
    b[0] = mu * mk; b[1] = mv * mk; b[2] = mk;
    if (muv) { muv[0] = mu; muv[1] = mv; }
}

// computePoses (ap3p.cpp:131-280).  b1..b3: bearings, w1..w3: world points.
//   cvMode = false: the reference's export -- every root with |root| <= 1 yields a pose (a complex root pair
//     appears as two identical poses), R stored as the un-transposed product (camera = R^T X + t).
//   cvMode = true: cv2 4.13's solvePnP(AP3P) on four points -- complex pairs dropped, R transposed
//     (camera = R X + t), poses insertion-sorted by the squared normalised re-projection error of the fourth
//     point (mu3, mv3, X3).
// Rs [4][9], ts [4][3]; returns the number of poses.
MCV_HDN inline int ap3p_poses(const double* b1, const double* b2, const double* b3, const double* w1, const double* w2,
                              const double* w3, bool cvMode, const double* muv3, const double* X3, double* Rs,
                              double* ts) {
    double u0[3] = {w1[0] - w2[0], w1[1] - w2[1], w1[2] - w2[2]};
    const double nu0 = v3norm(u0);
    double k1[3] = {u0[0] / nu0, u0[1] / nu0, u0[2] / nu0};
    double k3[3];
    v3cross(b1, b2, k3);
    const double nk3 = v3norm(k3);
    k3[0] /= nk3; k3[1] /= nk3; k3[2] /= nk3;
    double tz[3], v1[3], v2[3];
    v3cross(b1, k3, tz);
    v3cross(b1, b3, v1);
    v3cross(b2, b3, v2);
    double u1[3] = {w1[0] - w3[0], w1[1] - w3[1], w1[2] - w3[2]};
    const double u1k1 = v3dot(u1, k1);
    const double k3b3 = v3dot(k3, b3);
    double f11 = k3b3;
    double f13 = v3dot(k3, v1);
    const double f15 = -u1k1 * f11;
    double nl[3];
    v3cross(u1, k1, nl);
    const double delta = v3norm(nl);
    nl[0] /= delta; nl[1] /= delta; nl[2] /= delta;
    f11 *= delta;
    f13 *= delta;
    const double u2k1 = u1k1 - nu0;
    double f21 = v3dot(tz, v2);
    double f22 = nk3 * k3b3;
    double f23 = v3dot(k3, v2);
    const double f24 = u2k1 * f22;
    const double f25 = -u2k1 * f21;
    f21 *= delta;
    f22 *= delta;
    f23 *= delta;
    const double g1 = f13 * f22;
    const double g2 = f13 * f25 - f15 * f23;
    const double g3 = f11 * f23 - f13 * f21;
    const double g4 = -f13 * f24;
    const double g5 = f11 * f22;
    const double g6 = f11 * f25 - f15 * f21;
    const double g7 = -f15 * f24;
    const double coeffs[5] = {g5 * g5 + g1 * g1 + g3 * g3,
                              2 * (g5 * g6 + g1 * g2 + g3 * g4),
                              g6 * g6 + 2 * g5 * g7 + g2 * g2 + g4 * g4 - g1 * g1 - g3 * g3,
                              2 * (g6 * g7 - g1 * g2 - g3 * g4),
                              g7 * g7 - g2 * g2 - g4 * g4};
    double s[4];
    ap3p_quartic(coeffs, s);
    bool keep[4] = {true, true, true, true};
    if (cvMode) {
        if (s[0] == s[1]) keep[0] = keep[1] = false;
        if (s[2] == s[3]) keep[2] = keep[3] = false;
    }
    // polishQuarticRoots (ap3p.cpp:64-78)
    for (int it = 0; it < 2; ++it)
        for (int j = 0; j < 4; ++j) {
            const double x = s[j];
            const double err = (((coeffs[0] * x + coeffs[1]) * x + coeffs[2]) * x + coeffs[3]) * x + coeffs[4];
            const double der = ((4 * coeffs[0] * x + 3 * coeffs[1]) * x + 2 * coeffs[2]) * x + coeffs[3];
            s[j] -= err / der;
        }
    double temp[3];
    v3cross(k1, nl, temp);
    const double Ck1nl[9] = {k1[0], nl[0], temp[0], k1[1], nl[1], temp[1], k1[2], nl[2], temp[2]};
    const double Cb1k3tzT[9] = {b1[0], b1[1], b1[2], k3[0], k3[1], k3[2], tz[0], tz[1], tz[2]};
    const double sc = delta / k3b3;
    const double b3p[3] = {b3[0] * sc, b3[1] * sc, b3[2] * sc};
    int nb = 0;
    double errs[4];
    for (int i = 0; i < 4; ++i) {
        const double ct1 = s[i];
        if (!keep[i]) continue;
        if (!(fabs(ct1) <= 1)) continue;         // also rejects NaN roots
        double st1 = sqrt(1 - ct1 * ct1);
        st1 = (k3b3 > 0) ? st1 : -st1;
        double ct3 = g1 * ct1 + g2;
        double st3 = g3 * ct1 + g4;
        const double nt3 = st1 / ((g5 * ct1 + g6) * ct1 + g7);
        ct3 *= nt3;
        st3 *= nt3;
        const double C13[9] = {ct3, 0, -st3, st1 * st3, ct1, st1 * ct3, ct1 * st3, -st1, ct1 * ct3};
        double tm[9], R[9];
        m3mul(Ck1nl, C13, tm);
        m3mul(tm, Cb1k3tzT, R);
        const double rp3[3] = {w3[0] * R[0] + w3[1] * R[3] + w3[2] * R[6], w3[0] * R[1] + w3[1] * R[4] + w3[2] * R[7],
                               w3[0] * R[2] + w3[1] * R[5] + w3[2] * R[8]};
        double* t = ts + 3 * nb;
        t[0] = b3p[0] * st1 - rp3[0];
        t[1] = b3p[1] * st1 - rp3[1];
        t[2] = b3p[2] * st1 - rp3[2];
        double* Ro = Rs + 9 * nb;
        if (!cvMode) {
            for (int k = 0; k < 9; ++k) Ro[k] = R[k];
        } else {
            for (int r = 0; r < 3; ++r)
                for (int c = 0; c < 3; ++c) Ro[3 * r + c] = R[3 * c + r];
            const double X3p = Ro[0] * X3[0] + Ro[1] * X3[1] + Ro[2] * X3[2] + t[0];
            const double Y3p = Ro[3] * X3[0] + Ro[4] * X3[1] + Ro[5] * X3[2] + t[1];
            const double Z3p = Ro[6] * X3[0] + Ro[7] * X3[1] + Ro[8] * X3[2] + t[2];
            const double mu3p = X3p / Z3p, mv3p = Y3p / Z3p;
            errs[nb] = (mu3p - muv3[0]) * (mu3p - muv3[0]) + (mv3p - muv3[1]) * (mv3p - muv3[1]);
        }
        ++nb;
    }
    if (cvMode)
        for (int i = 1; i < nb; ++i)
            for (int j = i; j > 0 && errs[j - 1] > errs[j]; --j) {
                double e = errs[j - 1]; errs[j - 1] = errs[j]; errs[j] = e;
                for (int k = 0; k < 9; ++k) { e = Rs[9 * (j - 1) + k]; Rs[9 * (j - 1) + k] = Rs[9 * j + k]; Rs[9 * j + k] = e; }
                for (int k = 0; k < 3; ++k) { e = ts[3 * (j - 1) + k]; ts[3 * (j - 1) + k] = ts[3 * j + k]; ts[3 * j + k] = e; }
            }
    return nb;
}

// ---- camera model ---------------------------------------------------------------------------------
struct PnpCamera {
    double fx, fy, cx, cy;
    double k1, k2, p1, p2;      // the four coefficients cvSolvePnPRansac passes (MiniCVNative.cpp:121)
    MCV_HD bool distorted() const { return k1 != 0 || k2 != 0 || p1 != 0 || p2 != 0; }
};

// cv::undistortPoints without R / P: five fixed-point iterations
MCV_HDN inline void undistort_point(const PnpCamera& c, double u, double v, double* x, double* y) {
    const double ifx = 1. / c.fx, ify = 1. / c.fy;
    double X = dmul(dsub(u, c.cx), ifx), Y = dmul(dsub(v, c.cy), ify);
    if (c.distorted()) {
        const double x0 = X, y0 = Y;
        for (int j = 0; j < 5; ++j) {
            const double r2 = dadd(dmul(X, X), dmul(Y, Y));
            // (1 + ((k7 r2 + k6) r2 + k5) r2) / (1 + ((k3 r2 + k2) r2 + k1) r2) with only k1, k2 non-zero
            const double icdist = ddiv(1., dadd(1., dmul(dadd(dmul(c.k2, r2), c.k1), r2)));
            if (icdist < 0) { X = dmul(dsub(u, c.cx), ifx); Y = dmul(dsub(v, c.cy), ify); break; }
            const double dX = dadd(dmul(dmul(dmul(2, c.p1), X), Y), dmul(c.p2, dadd(r2, dmul(dmul(2, X), X))));
            const double dY = dadd(dmul(c.p1, dadd(r2, dmul(dmul(2, Y), Y))), dmul(dmul(dmul(2, c.p2), X), Y));
            X = dmul(dsub(x0, dX), icdist);
            Y = dmul(dsub(y0, dY), icdist);
        }
    }
    *x = X; *y = Y;
}

// cv::Rodrigues, vector -> matrix
MCV_HDN inline void rodrigues_v2m(const double* r, double* R) {
    const double theta = sqrt(dadd(dadd(dmul(r[0], r[0]), dmul(r[1], r[1])), dmul(r[2], r[2])));
    if (theta < DBL_EPSILON) {
        for (int k = 0; k < 9; ++k) R[k] = (k % 4 == 0) ? 1.0 : 0.0;
        return;
    }
    const double c = cos(theta), s = sin(theta), c1 = 1. - c, it = 1. / theta;
    const double x = r[0] * it, y = r[1] * it, z = r[2] * it;
    const double rrt[9] = {x * x, x * y, x * z, x * y, y * y, y * z, x * z, y * z, z * z};
    const double rx[9] = {0, -z, y, z, 0, -x, -y, x, 0};
    for (int k = 0; k < 9; ++k) {
        const double e = (k % 4 == 0) ? 1.0 : 0.0;
        R[k] = dadd(dadd(dmul(c, e), dmul(c1, rrt[k])), dmul(s, rx[k]));
    }
}

// cv::Rodrigues, matrix -> vector (R re-orthonormalised as U V^T first)
MCV_HDN inline void rodrigues_m2v(const double* Rin, double* r) {
    for (int k = 0; k < 9; ++k)
        if (!(fabs(Rin[k]) < 100)) { r[0] = r[1] = r[2] = 0; return; }     // checkRange(-100, 100), NaN included
    double w[3], u[9], vt[9], R[9];
    svd_square<3>(Rin, w, u, vt);
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) R[3 * i + j] = u[3 * i] * vt[j] + u[3 * i + 1] * vt[3 + j] + u[3 * i + 2] * vt[6 + j];
    double x = R[7] - R[5], y = R[2] - R[6], z = R[3] - R[1];
    const double s = sqrt((x * x + y * y + z * z) * 0.25);
    double c = (R[0] + R[4] + R[8] - 1) * 0.5;
    c = c > 1. ? 1. : c < -1. ? -1. : c;
    double theta = acos(c);
    if (s < 1e-5) {
        if (c > 0) { r[0] = r[1] = r[2] = 0; return; }
        double t = (R[0] + 1) * 0.5;
        x = sqrt(t > 0 ? t : 0.);
        t = (R[4] + 1) * 0.5;
        y = sqrt(t > 0 ? t : 0.) * (R[1] < 0 ? -1. : 1.);
        t = (R[8] + 1) * 0.5;
        z = sqrt(t > 0 ? t : 0.) * (R[2] < 0 ? -1. : 1.);
        if (fabs(x) < fabs(y) && fabs(x) < fabs(z) && (R[5] > 0) != (y * z > 0)) z = -z;
        theta /= sqrt(x * x + y * y + z * z);
        r[0] = x * theta; r[1] = y * theta; r[2] = z * theta;
        return;
    }
    double vth = 1 / (2 * s);
    vth *= theta;
    r[0] = x * vth; r[1] = y * vth; r[2] = z * vth;
}

// cv::projectPoints for one point (R from rodrigues_v2m) followed by PnPRansacCallback::computeError:
// the projection is rounded to float32, the residual and its squared norm are float32 arithmetic.
// Every operation is rounded separately (no FMA contraction), in OpenCV's order.
MCV_HD float project_error(const double* R, const double* t, const PnpCamera& c, bool dist, float Xf, float Yf, float Zf,
                           float uf, float vf) {
    const double X = Xf, Y = Yf, Z = Zf;
    double x = dadd(dadd(dadd(dmul(R[0], X), dmul(R[1], Y)), dmul(R[2], Z)), t[0]);
    double y = dadd(dadd(dadd(dmul(R[3], X), dmul(R[4], Y)), dmul(R[5], Z)), t[1]);
    double z = dadd(dadd(dadd(dmul(R[6], X), dmul(R[7], Y)), dmul(R[8], Z)), t[2]);
    z = z != 0 ? ddiv(1., z) : 1.;
    x = dmul(x, z);
    y = dmul(y, z);
    if (dist) {
        const double r2 = dadd(dmul(x, x), dmul(y, y));
        const double r4 = dmul(r2, r2);
        const double a1 = dmul(dmul(2, x), y);
        const double a2 = dadd(r2, dmul(dmul(2, x), x));
        const double a3 = dadd(r2, dmul(dmul(2, y), y));
        // 1 + k1 r2 + k2 r4 + k3 r6 with k3 = 0 (adding +0 changes nothing); icdist2 = 1
        const double cdist = dadd(dadd(1, dmul(c.k1, r2)), dmul(c.k2, r4));
        const double xd = dadd(dadd(dmul(x, cdist), dmul(c.p1, a1)), dmul(c.p2, a2));
        const double yd = dadd(dadd(dmul(y, cdist), dmul(c.p1, a3)), dmul(c.p2, a1));
        x = xd; y = yd;
    }
    const float pu = (float)dadd(dmul(x, c.fx), c.cx);
    const float pv = (float)dadd(dmul(y, c.fy), c.cy);
    const float dx = fsub(uf, pu), dy = fsub(vf, pv);
    return fadd(fmul(dx, dx), fmul(dy, dy));
}

// ---- fp32 filter in front of project_error <= thr (cameras without distortion) -----------------------------------
// Same idea as the fundamental / essential filters of cvmath.cuh: most (pose, correspondence) pairs of a RANSAC batch
// are outliers by a wide margin, and the exact test above costs ~30 FP64-pipe instructions, a division and five
// conversions.  In fp32, with rows 0 / 1 of [R | t] pre-multiplied by fx / fy:
//     x' = fx (R0 . X + t0),  y' = fy (R1 . X + t1),  z = R2 . X + t2,   u - pu = (uc z - x') / z,  uc = u - cx
// so with g = (uc z - x', vc z - y') the squared reprojection error is |g|^2 / z^2 -- no division needed to compare it
// with the threshold.  Error budget per tile of correspondences (u = 2^-24; Xm >= |X| etc. over the tile):
//     |x~' - x'| <= 5u Xb,  Xb = |r0| Xm + |r1| Ym + |r2| Zm + |t0'|   (same for y', z);   eZ = 8u Zb >= |z~ - z|
//     |g~x - gx| <= 8u (UCm Zb + Xb)                                  -> 16u (..) used
//   and the exact path's own float32 roundings (pu = (float)(..), dx = u - pu) move its |dx| by at most
//   u (|u| + 2 |dx|): a further 2.1u Um Zb on the g scale; the relative part goes into the threshold margin.
// With ex, ey those bounds, zh = |z~| + eZ >= |z| and (p + q)^2 <= (1 + mu) p^2 + (1 + 1/mu) q^2 (mu = 2^-12) applied to
// both sides, the single comparison
//     g~x^2 + g~y^2 > thr (1 + 2^-9)(1 + mu)^2 z~^2 + K,   K = thr 1.003 (64.02 eZ)^2 + (64.02 ex)^2 + (64.02 ey)^2 + 1e-30
// together with |z~| > eZ (z certainly non-zero: project_error special-cases z = 0) implies that the float32 error the
// exact path computes exceeds thr.  Filter off (K = inf) for distorted cameras and non-finite / extreme poses.
struct PnpFilter {
    float r[9], t[3];
    bool on;
};
struct PnpFilterTile {
    float K, eZ, thrMM;
};
MCV_HD void pnp_filter_init(const double* R, const double* t, const PnpCamera& c, bool dist, PnpFilter& pf) {
    const double sc[3] = {c.fx, c.fy, 1.0};
    bool on = !dist;
    for (int i = 0; i < 3; ++i) {
        for (int j = 0; j < 4; ++j) {
            const double v = sc[i] * (j < 3 ? R[3 * i + j] : t[i]);
            const double av = fabs(v);
            if (!(av <= 1e30) || (av != 0.0 && av < 1e-30)) on = false;
            if (j < 3) pf.r[3 * i + j] = (float)v; else pf.t[i] = (float)v;
        }
    }
    pf.on = on;
}
// maxima over the tile: |X|, |Y|, |Z| of the world points, |u - cx|, |v - cy|, |u|, |v| of the image points
MCV_HD void pnp_filter_tile(const PnpFilter& pf, float Xm, float Ym, float Zm, float UCm, float VCm, float Um, float Vm,
                            float thr, PnpFilterTile& ft) {
    using cvm::FILTER_EPS_A; using cvm::FILTER_EPS_D; using cvm::FILTER_SQ; using cvm::FILTER_TINY;
    const float Xb = fmaf(fabsf(pf.r[0]), Xm, fmaf(fabsf(pf.r[1]), Ym, fmaf(fabsf(pf.r[2]), Zm, fabsf(pf.t[0]) + FILTER_TINY)));
    const float Yb = fmaf(fabsf(pf.r[3]), Xm, fmaf(fabsf(pf.r[4]), Ym, fmaf(fabsf(pf.r[5]), Zm, fabsf(pf.t[1]) + FILTER_TINY)));
    const float Zb = fmaf(fabsf(pf.r[6]), Xm, fmaf(fabsf(pf.r[7]), Ym, fmaf(fabsf(pf.r[8]), Zm, fabsf(pf.t[2]) + FILTER_TINY)));
    const float u21 = 1.2516975402832031e-07f;                   // 2.1 * 2^-24
    const float eZ = FILTER_EPS_A * Zb;
    const float ex = fmaf(FILTER_EPS_D, fmaf(UCm, Zb, Xb), u21 * Um * Zb);
    const float ey = fmaf(FILTER_EPS_D, fmaf(VCm, Zb, Yb), u21 * Vm * Zb);
    const float sZ = FILTER_SQ * eZ, sx = FILTER_SQ * ex, sy = FILTER_SQ * ey;
    ft.eZ = eZ;
    ft.thrMM = thr * (1.001953125f * 1.000244140625f * 1.000244140625f * 1.0001f);
    ft.K = pf.on ? fmaf(thr * 1.003f, sZ * sZ, fmaf(sx, sx, fmaf(sy, sy, 1e-30f))) : INFINITY;
}
// uc = (float)(u - cx), vc = (float)(v - cy)
MCV_HD bool pnp_certain_outlier(const PnpFilter& pf, const PnpFilterTile& ft, float X, float Y, float Z, float uc, float vc) {
    const float x = fmaf(pf.r[0], X, fmaf(pf.r[1], Y, fmaf(pf.r[2], Z, pf.t[0])));
    const float y = fmaf(pf.r[3], X, fmaf(pf.r[4], Y, fmaf(pf.r[5], Z, pf.t[1])));
    const float z = fmaf(pf.r[6], X, fmaf(pf.r[7], Y, fmaf(pf.r[8], Z, pf.t[2])));
    const float gx = fmaf(uc, z, -x), gy = fmaf(vc, z, -y);
    const float lhs = fmaf(gx, gx, gy * gy);
    const float rhs = fmaf(ft.thrMM, z * z, ft.K);
    return lhs > rhs && fabsf(z) > ft.eZ;
}

// cv::solvePnP(SOLVEPNP_AP3P) on four float32 correspondences, as PnPRansacCallback::runKernel calls it:
// undistort (float32 result), AP3P on the first three, fourth point picks the pose, Rodrigues.
// obj [4][3], img [4][2]; model = rvec[3], tvec[3].
MCV_HDN inline bool pnp_solve4(const PnpCamera& c, const float* obj, const float* img, double* rvec, double* tvec) {
    const double inv_fx = 1. / c.fx, inv_fy = 1. / c.fy, cx_fx = c.cx / c.fx, cy_fy = c.cy / c.fy;
    double b[3][3], muv3[2] = {0, 0};
    for (int i = 0; i < 4; ++i) {
        double x, y;
        undistort_point(c, (double)img[2 * i], (double)img[2 * i + 1], &x, &y);
        // float32 undistorted point, K re-applied by ap3p::extract_points and removed again by ap3p::solve
        const double pu = (double)(float)x * c.fx + c.cx;
        const double pv = (double)(float)y * c.fy + c.cy;
        if (i < 3) ap3p_bearing(pu, pv, inv_fx, inv_fy, cx_fx, cy_fy, b[i], nullptr);
        else { muv3[0] = inv_fx * pu - cx_fx; muv3[1] = inv_fy * pv - cy_fy; }
    }
    double w[4][3];
    for (int i = 0; i < 4; ++i)
        for (int k = 0; k < 3; ++k) w[i][k] = (double)obj[3 * i + k];
    double Rs[36], ts[12];
    const int n = ap3p_poses(b[0], b[1], b[2], w[0], w[1], w[2], true, muv3, w[3], Rs, ts);
    if (n == 0) return false;
    rodrigues_m2v(Rs, rvec);
    tvec[0] = ts[0]; tvec[1] = ts[1]; tvec[2] = ts[2];
    return true;
}

// ---- EPnP -------------------------------------------------------------------------------------------
// cv::SVBkSb for one right-hand side.  u: N rows of length M (left vectors), vt: N x N, w[N].
template <int M, int N>
MCV_HDN inline void sv_bksb(const double* w, const double* u, const double* vt, const double* b, double* x) {
    double thr = 0;
    for (int i = 0; i < N; ++i) thr += w[i];
    thr *= DBL_EPSILON * 2;
    for (int j = 0; j < N; ++j) x[j] = 0;
    for (int i = 0; i < N; ++i) {
        double wi = w[i];
        if (fabs(wi) <= thr) continue;
        wi = 1 / wi;
        double s = 0;
        for (int j = 0; j < M; ++j) s += u[i * M + j] * b[j];
        s *= wi;
        for (int j = 0; j < N; ++j) x[j] = x[j] + s * vt[i * N + j];
    }
}

// cv::solve(A, b, x, DECOMP_SVD) for a 6 x N system whose columns are cols[] of the 6 x 10 matrix L
template <int N>
MCV_HDN inline void epnp_solve6(const double* L, const int* cols, const double* rho, double* x) {
    double At[N * 6], W[N], Vt[N * N];
    for (int i = 0; i < N; ++i)
        for (int k = 0; k < 6; ++k) At[i * 6 + k] = L[10 * k + cols[i]];
    jacobi_svd_rows<6, N, N, true>(At, W, Vt);
    sv_bksb<6, N>(W, At, Vt, rho, x);
}

// epnp::qr_solve for the 6 x 4 Gauss-Newton system (its pivot-scale loop scans rows k .. nr-2, kept as is)
MCV_HDN inline bool epnp_qr_solve(double* A, double* b, double* X) {
    constexpr int nr = 6, nc = 4;
    double A1[nc], A2[nc];
    for (int k = 0; k < nc; ++k) {
        double eta = fabs(A[k * nc + k]);
        for (int i = k + 1; i < nr; ++i) {
            const double elt = fabs(A[(i - 1) * nc + k]);
            if (eta < elt) eta = elt;
        }
        if (eta == 0) return false;
        const double inv_eta = 1. / eta;
        double sum2 = 0;
        for (int i = k; i < nr; ++i) { A[i * nc + k] *= inv_eta; sum2 += A[i * nc + k] * A[i * nc + k]; }
        double sigma = sqrt(sum2);
        if (A[k * nc + k] < 0) sigma = -sigma;
        A[k * nc + k] += sigma;
        A1[k] = sigma * A[k * nc + k];
        A2[k] = -eta * sigma;
        for (int j = k + 1; j < nc; ++j) {
            double sum = 0;
            for (int i = k; i < nr; ++i) sum += A[i * nc + k] * A[i * nc + j];
            const double tau = sum / A1[k];
            for (int i = k; i < nr; ++i) A[i * nc + j] -= tau * A[i * nc + k];
        }
    }
    for (int j = 0; j < nc; ++j) {
        double tau = 0;
        for (int i = j; i < nr; ++i) tau += A[i * nc + j] * b[i];
        tau /= A1[j];
        for (int i = j; i < nr; ++i) b[i] -= tau * A[i * nc + j];
    }
    X[nc - 1] = b[nc - 1] / A2[nc - 1];
    for (int i = nc - 2; i >= 0; --i) {
        double sum = 0;
        for (int j = i + 1; j < nc; ++j) sum += A[i * nc + j] * X[j];
        X[i] = (b[i] - sum) / A2[i];
    }
    return true;
}

// Stage 1 (after the centroid and the 3x3 scatter of the world points): control points + inverse basis.
//   cws [4][3] with cws[0] = centroid on entry; cov = sum (p - c)(p - c)^T; ccInv [9] out.
MCV_HDN inline void epnp_control_points(int n, const double* cov, double* cws, double* ccInv) {
    double dc[3], u[9], vt[9];
    svd_square<3>(cov, dc, u, vt);
    for (int i = 1; i < 4; ++i) {
        const double k = sqrt(dc[i - 1] / n);
        for (int j = 0; j < 3; ++j) cws[3 * i + j] = cws[j] + k * u[3 * j + (i - 1)];      // uct[3 (i-1) + j]
    }
    double cc[9];
    for (int i = 0; i < 3; ++i)
        for (int j = 1; j < 4; ++j) cc[3 * i + j - 1] = cws[3 * j + i] - cws[i];
    // cv::invert(DECOMP_SVD)
    double w[3];
    svd_square<3>(cc, w, u, vt);
    const double thr = (w[0] + w[1] + w[2]) * DBL_EPSILON * 2;
    for (int k = 0; k < 9; ++k) ccInv[k] = 0;
    for (int i = 0; i < 3; ++i) {
        double wi = w[i];
        if (fabs(wi) <= thr) continue;
        wi = 1 / wi;
        for (int j = 0; j < 3; ++j)
            for (int k = 0; k < 3; ++k) ccInv[3 * j + k] = ccInv[3 * j + k] + vt[3 * i + j] * (u[3 * k + i] * wi);
    }
}

// per point: barycentric coordinates
MCV_HD void epnp_alphas(const double* cws, const double* ci, const double* p, double* a) {
    const double d0 = p[0] - cws[0], d1 = p[1] - cws[1], d2 = p[2] - cws[2];
    for (int j = 0; j < 3; ++j) a[1 + j] = ci[3 * j] * d0 + ci[3 * j + 1] * d1 + ci[3 * j + 2] * d2;
    a[0] = 1.0 - a[1] - a[2] - a[3];
}

// per point: contribution to M^T M.  M's two rows for a point are [a_j fu, 0, a_j du] and [0, a_j fv, a_j dv]
// (du = uc - u, dv = vc - v), so M^T M = sum kron(a a^T, S(du, dv)); acc holds, for the ten pairs i <= j,
// sum a_i a_j * {1, du, dv, du^2 + dv^2}.
MCV_HD void epnp_accum(const double* a, double du, double dv, double* acc /*[40]*/) {
    const double q = du * du + dv * dv;
    int k = 0;
    for (int i = 0; i < 4; ++i)
        for (int j = i; j < 4; ++j, ++k) {
            const double aij = a[i] * a[j];
            acc[k] += aij;
            acc[10 + k] += aij * du;
            acc[20 + k] += aij * dv;
            acc[30 + k] += aij * q;
        }
}

MCV_HDN inline void epnp_mtm(const double* acc, double fu, double fv, double* MtM /*[144]*/) {
    int k = 0;
    for (int i = 0; i < 4; ++i)
        for (int j = i; j < 4; ++j, ++k) {
            double blk[9] = {fu * fu * acc[k], 0, fu * acc[10 + k], 0, fv * fv * acc[k], fv * acc[20 + k],
                             fu * acc[10 + k], fv * acc[20 + k], acc[30 + k]};
            for (int p = 0; p < 3; ++p)
                for (int q = 0; q < 3; ++q) {
                    MtM[(3 * i + p) * 12 + 3 * j + q] = blk[3 * p + q];
                    MtM[(3 * j + q) * 12 + 3 * i + p] = blk[3 * p + q];
                }
        }
}

// Stage 2: null space of M^T M, L (6x10), rho, the three beta initialisations and their Gauss-Newton
// refinements.  ut4 [4][12]: rows 11, 10, 9, 8 of U^T; betas [3][4].
MCV_HDN inline void epnp_betas(const double* MtM, const double* cws, double* ut4, double* betas) {
    double At[144], W[12];
    for (int i = 0; i < 12; ++i)
        for (int j = 0; j < 12; ++j) At[i * 12 + j] = MtM[j * 12 + i];
    jacobi_svd_rows<12, 12, 12, false>(At, W, nullptr);
    for (int i = 0; i < 4; ++i)
        for (int k = 0; k < 12; ++k) ut4[12 * i + k] = At[(11 - i) * 12 + k];
    double L[60], rho[6];
    {
        double dv[4][6][3];
        for (int i = 0; i < 4; ++i) {
            const double* v = ut4 + 12 * i;
            int a = 0, b = 1;
            for (int j = 0; j < 6; ++j) {
                for (int k = 0; k < 3; ++k) dv[i][j][k] = v[3 * a + k] - v[3 * b + k];
                if (++b > 3) { ++a; b = a + 1; }
            }
        }
        for (int i = 0; i < 6; ++i) {
            double* row = L + 10 * i;
            row[0] = v3dot(dv[0][i], dv[0][i]);
            row[1] = 2.0 * v3dot(dv[0][i], dv[1][i]);
            row[2] = v3dot(dv[1][i], dv[1][i]);
            row[3] = 2.0 * v3dot(dv[0][i], dv[2][i]);
            row[4] = 2.0 * v3dot(dv[1][i], dv[2][i]);
            row[5] = v3dot(dv[2][i], dv[2][i]);
            row[6] = 2.0 * v3dot(dv[0][i], dv[3][i]);
            row[7] = 2.0 * v3dot(dv[1][i], dv[3][i]);
            row[8] = 2.0 * v3dot(dv[2][i], dv[3][i]);
            row[9] = v3dot(dv[3][i], dv[3][i]);
        }
        int k = 0;
        for (int a = 0; a < 4; ++a)
            for (int b = a + 1; b < 4; ++b, ++k) {
                const double d0 = cws[3 * a] - cws[3 * b], d1 = cws[3 * a + 1] - cws[3 * b + 1], d2 = cws[3 * a + 2] - cws[3 * b + 2];
                rho[k] = d0 * d0 + d1 * d1 + d2 * d2;
            }
    }
    {   // approx 1: [B11 B12 B13 B14]
        const int cols[4] = {0, 1, 3, 6};
        double b4[4];
        epnp_solve6<4>(L, cols, rho, b4);
        double* bt = betas;
        if (b4[0] < 0) { bt[0] = sqrt(-b4[0]); bt[1] = -b4[1] / bt[0]; bt[2] = -b4[2] / bt[0]; bt[3] = -b4[3] / bt[0]; }
        else { bt[0] = sqrt(b4[0]); bt[1] = b4[1] / bt[0]; bt[2] = b4[2] / bt[0]; bt[3] = b4[3] / bt[0]; }
    }
    {   // approx 2: [B11 B12 B22]
        const int cols[3] = {0, 1, 2};
        double b3[3];
        epnp_solve6<3>(L, cols, rho, b3);
        double* bt = betas + 4;
        if (b3[0] < 0) { bt[0] = sqrt(-b3[0]); bt[1] = (b3[2] < 0) ? sqrt(-b3[2]) : 0.0; }
        else { bt[0] = sqrt(b3[0]); bt[1] = (b3[2] > 0) ? sqrt(b3[2]) : 0.0; }
        if (b3[1] < 0) bt[0] = -bt[0];
        bt[2] = 0.0; bt[3] = 0.0;
    }
    {   // approx 3: [B11 B12 B22 B13 B23]
        const int cols[5] = {0, 1, 2, 3, 4};
        double b5[5];
        epnp_solve6<5>(L, cols, rho, b5);
        double* bt = betas + 8;
        if (b5[0] < 0) { bt[0] = sqrt(-b5[0]); bt[1] = (b5[2] < 0) ? sqrt(-b5[2]) : 0.0; }
        else { bt[0] = sqrt(b5[0]); bt[1] = (b5[2] > 0) ? sqrt(b5[2]) : 0.0; }
        if (b5[1] < 0) bt[0] = -bt[0];
        bt[2] = b5[3] / bt[0]; bt[3] = 0.0;
    }
    for (int c = 0; c < 3; ++c) {
        double* bt = betas + 4 * c;
        for (int it = 0; it < 5; ++it) {
            double A[24], B[6], x[4];
            for (int i = 0; i < 6; ++i) {
                const double* r = L + 10 * i;
                A[4 * i + 0] = 2 * r[0] * bt[0] + r[1] * bt[1] + r[3] * bt[2] + r[6] * bt[3];
                A[4 * i + 1] = r[1] * bt[0] + 2 * r[2] * bt[1] + r[4] * bt[2] + r[7] * bt[3];
                A[4 * i + 2] = r[3] * bt[0] + r[4] * bt[1] + 2 * r[5] * bt[2] + r[8] * bt[3];
                A[4 * i + 3] = r[6] * bt[0] + r[7] * bt[1] + r[8] * bt[2] + 2 * r[9] * bt[3];
                B[i] = rho[i] - (r[0] * bt[0] * bt[0] + r[1] * bt[0] * bt[1] + r[2] * bt[1] * bt[1] + r[3] * bt[0] * bt[2] +
                                 r[4] * bt[1] * bt[2] + r[5] * bt[2] * bt[2] + r[6] * bt[0] * bt[3] + r[7] * bt[1] * bt[3] +
                                 r[8] * bt[2] * bt[3] + r[9] * bt[3] * bt[3]);
            }
            if (!epnp_qr_solve(A, B, x)) break;
            for (int i = 0; i < 4; ++i) bt[i] += x[i];
        }
    }
}

// control points in the camera frame for one beta set
MCV_HD void epnp_ccs(const double* ut4, const double* bt, double* ccs /*[4][3]*/) {
    for (int k = 0; k < 12; ++k) ccs[k] = 0;
    for (int i = 0; i < 4; ++i)
        for (int k = 0; k < 12; ++k) ccs[k] += bt[i] * ut4[12 * i + k];
}
MCV_HD void epnp_pc(const double* a, const double* ccs, double* pc) {
    for (int j = 0; j < 3; ++j) pc[j] = a[0] * ccs[j] + a[1] * ccs[3 + j] + a[2] * ccs[6 + j] + a[3] * ccs[9 + j];
}

// Stage 3: rotation and translation from the 3x3 correlation of centred camera / world points
MCV_HDN inline void epnp_rt(const double* abt, const double* pc0, const double* pw0, double* R, double* t) {
    double d[3], u[9], vt[9];
    svd_square<3>(abt, d, u, vt);
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) R[3 * i + j] = u[3 * i] * vt[j] + u[3 * i + 1] * vt[3 + j] + u[3 * i + 2] * vt[6 + j];
    const double det = R[0] * R[4] * R[8] + R[1] * R[5] * R[6] + R[2] * R[3] * R[7] - R[2] * R[4] * R[6] - R[1] * R[3] * R[8] -
                       R[0] * R[5] * R[7];
    if (det < 0) { R[6] = -R[6]; R[7] = -R[7]; R[8] = -R[8]; }
    t[0] = pc0[0] - v3dot(R, pw0);
    t[1] = pc0[1] - v3dot(R + 3, pw0);
    t[2] = pc0[2] - v3dot(R + 6, pw0);
}

// per point: epnp::reprojection_error term
MCV_HD double epnp_reproj(const double* R, const double* t, const double* pw, double fu, double fv, double uc, double vc,
                          double u, double v) {
    const double Xc = v3dot(R, pw) + t[0], Yc = v3dot(R + 3, pw) + t[1], inv_Zc = 1.0 / (v3dot(R + 6, pw) + t[2]);
    const double ue = uc + fu * Xc * inv_Zc, ve = vc + fv * Yc * inv_Zc;
    return sqrt((u - ue) * (u - ue) + (v - ve) * (v - ve));
}

// Whole EPnP, sequential: the host-side composition of the stages above (unit tests, and the shape the
// finalize kernel of pnp.cu follows with block-wide reductions in place of the loops over points).
//   pw [n][3] world points, und [n][2] undistorted (normalised) image points.
MCV_HDN inline void epnp_sequential(int n, const double* pw, const double* und, double fu, double fv, double uc, double vc,
                                    double* R, double* t) {
    double cws[12] = {0};
    for (int i = 0; i < n; ++i)
        for (int j = 0; j < 3; ++j) cws[j] += pw[3 * i + j];
    for (int j = 0; j < 3; ++j) cws[j] /= n;
    double cov[9] = {0};
    for (int i = 0; i < n; ++i) {
        const double d[3] = {pw[3 * i] - cws[0], pw[3 * i + 1] - cws[1], pw[3 * i + 2] - cws[2]};
        for (int p = 0; p < 3; ++p)
            for (int q = 0; q < 3; ++q) cov[3 * p + q] += d[p] * d[q];
    }
    double ci[9];
    epnp_control_points(n, cov, cws, ci);
    double acc[40] = {0};
    for (int i = 0; i < n; ++i) {
        double a[4];
        epnp_alphas(cws, ci, pw + 3 * i, a);
        const double u = und[2 * i] * fu + uc, v = und[2 * i + 1] * fv + vc;
        epnp_accum(a, uc - u, vc - v, acc);
    }
    double MtM[144], ut4[48], betas[12];
    epnp_mtm(acc, fu, fv, MtM);
    epnp_betas(MtM, cws, ut4, betas);
    double bestErr = 0;
    for (int c = 0; c < 3; ++c) {
        double ccs[12], a[4], pc[3];
        epnp_ccs(ut4, betas + 4 * c, ccs);
        epnp_alphas(cws, ci, pw, a);
        epnp_pc(a, ccs, pc);
        if (pc[2] < 0.0)
            for (int k = 0; k < 12; ++k) ccs[k] = -ccs[k];
        double pc0[3] = {0, 0, 0}, pw0[3] = {0, 0, 0};
        for (int i = 0; i < n; ++i) {
            epnp_alphas(cws, ci, pw + 3 * i, a);
            epnp_pc(a, ccs, pc);
            for (int j = 0; j < 3; ++j) { pc0[j] += pc[j]; pw0[j] += pw[3 * i + j]; }
        }
        for (int j = 0; j < 3; ++j) { pc0[j] /= n; pw0[j] /= n; }
        double abt[9] = {0};
        for (int i = 0; i < n; ++i) {
            epnp_alphas(cws, ci, pw + 3 * i, a);
            epnp_pc(a, ccs, pc);
            for (int j = 0; j < 3; ++j)
                for (int k = 0; k < 3; ++k) abt[3 * j + k] += (pc[j] - pc0[j]) * (pw[3 * i + k] - pw0[k]);
        }
        double Rc[9], tc[3];
        epnp_rt(abt, pc0, pw0, Rc, tc);
        double sum2 = 0;
        for (int i = 0; i < n; ++i)
            sum2 += epnp_reproj(Rc, tc, pw + 3 * i, fu, fv, uc, vc, und[2 * i] * fu + uc, und[2 * i + 1] * fv + vc);
        const double err = sum2 / n;
        // N = 1; if (err[2] < err[1]) N = 2; if (err[3] < err[N]) N = 3
        if (c == 0 || err < bestErr) {
            bestErr = err;
            for (int k = 0; k < 9; ++k) R[k] = Rc[k];
            for (int k = 0; k < 3; ++k) t[k] = tc[k];
        }
    }
}

}  // namespace cvm
}  // namespace mcv
