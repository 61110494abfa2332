"""CPU oracle for the PnP-RANSAC row.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).

Reference call sites:
  * ``solveAp3p``            ``src/MiniCVNative/ap3p.cpp:282-317``  -> :func:`solve_ap3p`
      - ``computePoses``     ``ap3p.cpp:131-280``                   -> :func:`ap3p_compute_poses`
      - ``solveQuartic``     ``ap3p.cpp:9-62``                      -> :func:`solve_quartic`
      - ``polishQuarticRoots`` ``ap3p.cpp:64-78``                   -> :func:`polish_quartic_roots`
  * ``cvSolvePnPRansac``     ``src/MiniCVNative/MiniCVNative.cpp:93-139`` forwards to
    ``cv::solvePnPRansac`` (OpenCV, un-vendored; pinned here to cv2 4.13.0)   -> :func:`solve_pnp_ransac`
  * ``cvSolvePnP``           ``MiniCVNative.cpp:48-91`` -> ``cv::solvePnP``   -> :func:`solve_pnp_ap3p4`,
    :func:`epnp`

Pinning: :func:`solve_ap3p` is checked against the reference's own ``ap3p.cpp`` compiled in place
(``oracle/ref_build/Makefile`` -> ``oracle/_ref/libap3p_ref.so``) and against ``cv2.solveP3P``;
everything that follows OpenCV is checked against cv2 4.13.0 outputs frozen in
``tests/golden/pnp.npz`` by ``tests/golden/make_golden_pnp.py``.
"""
from __future__ import annotations

import cmath
import math

import numpy as np

from . import linalg_cv
from .ransac import F32, ransac_run

F64 = np.float64


# ------------------------------------------------------------------------------------------
# AP3P (Ke & Roumeliotis, CVPR 2017) as vendored by the reference
# ------------------------------------------------------------------------------------------
def solve_quartic(f):
    """Closed-form real parts of the four roots of f0 x^4 + ... + f4 (ap3p.cpp:9-62)."""
    a4, a3, a2, a1, a0 = (float(v) for v in f)
    a4_2 = a4 * a4
    a3_2 = a3 * a3
    a4_3 = a4_2 * a4
    a2a4 = a2 * a4
    p4 = (8 * a2a4 - 3 * a3_2) / (8 * a4_2)
    q4 = (a3_2 * a3 - 4 * a2a4 * a3 + 8 * a1 * a4_2) / (8 * a4_3)
    r4 = (256 * a0 * a4_3 - 3 * (a3_2 * a3_2) - 64 * a1 * a3 * a4_2 + 16 * a2a4 * a3_2) / (256 * (a4_3 * a4))
    p3 = ((p4 * p4) / 12 + r4) / 3
    q3 = (72 * r4 * p4 - 2 * p4 * p4 * p4 - 27 * q4 * q4) / 432
    disc = cmath.sqrt(complex(q3 * q3 - p3 * p3 * p3, 0.0))
    w = (-disc - q3) if q3 >= 0 else (disc - q3)
    if w.imag == 0.0:
        wr = math.copysign(abs(w.real) ** (1.0 / 3.0), w.real) if w.real != 0 else 0.0
        # cbrt: refine the power so that it agrees with libm's correctly-rounded cbrt on exact cubes
        if wr != 0.0:
            wr = wr - (wr * wr * wr - w.real) / (3 * wr * wr)
        t = 2.0 * (wr + p3 / wr)
    else:
        w = w ** (1.0 / 3)
        t = 4.0 * w.real
    sqrt_2m = cmath.sqrt(complex(-2 * p4 / 3 + t, 0.0))
    b_4a = -a3 / (4 * a4)
    complex1 = 4 * p4 / 3 + t
    complex2 = 2 * q4 / sqrt_2m
    sqrt_2m_rh = sqrt_2m.real / 2
    sqrt1 = cmath.sqrt(-(complex1 + complex2)).real / 2
    sqrt2 = cmath.sqrt(-(complex1 - complex2)).real / 2
    return [b_4a + sqrt_2m_rh + sqrt1, b_4a + sqrt_2m_rh - sqrt1,
            b_4a - sqrt_2m_rh + sqrt2, b_4a - sqrt_2m_rh - sqrt2]


def polish_quartic_roots(c, roots, iterations=2):
    """Two Newton steps per root (ap3p.cpp:64-78)."""
    r = list(roots)
    for _ in range(iterations):
        for j in range(4):
            x = r[j]
            err = (((c[0] * x + c[1]) * x + c[2]) * x + c[3]) * x + c[4]
            der = ((4 * c[0] * x + 3 * c[1]) * x + 2 * c[2]) * x + c[3]
            r[j] = x - err / der
    return r


def _cross(a, b):
    return np.array([a[1] * b[2] - a[2] * b[1], -(a[0] * b[2] - a[2] * b[0]), a[0] * b[1] - a[1] * b[0]], F64)


def _dot(a, b):
    return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]


def _norm(a):
    return math.sqrt(a[0] * a[0] + a[1] * a[1] + a[2] * a[2])


def _matmul3(a, b):
    r = np.empty((3, 3), F64)
    for i in range(3):
        for j in range(3):
            r[i, j] = a[i][0] * b[0][j] + a[i][1] * b[1][j] + a[i][2] * b[2][j]
    return r


def ap3p_compute_poses(fv, wp, fourth=None, drop_complex=False):
    """``computePoses`` (ap3p.cpp:131-280).  ``fv``/``wp`` hold the bearing / world vectors as columns.

    Returns a list of (R, t).  ``fourth = (mu3, mv3, X3, Y3, Z3)`` adds what OpenCV's 4-point form does on
    top of the reference's 3-point code: the squared reprojection error of the fourth point per solution and
    a stable insertion sort of the solutions by it.
    """
    fv = np.asarray(fv, F64)
    wp = np.asarray(wp, F64)
    w1, w2, w3 = wp[:, 0].copy(), wp[:, 1].copy(), wp[:, 2].copy()
    u0 = w1 - w2
    nu0 = _norm(u0)
    k1 = u0 / nu0
    b1, b2, b3 = fv[:, 0].copy(), fv[:, 1].copy(), fv[:, 2].copy()
    k3 = _cross(b1, b2)
    nk3 = _norm(k3)
    k3 = k3 / nk3
    tz = _cross(b1, k3)
    v1 = _cross(b1, b3)
    v2 = _cross(b2, b3)
    u1 = w1 - w3
    u1k1 = _dot(u1, k1)
    k3b3 = _dot(k3, b3)
    f11 = k3b3
    f13 = _dot(k3, v1)
    f15 = -u1k1 * f11
    nl = _cross(u1, k1)
    delta = _norm(nl)
    nl = nl / delta
    f11 *= delta
    f13 *= delta
    u2k1 = u1k1 - nu0
    f21 = _dot(tz, v2)
    f22 = nk3 * k3b3
    f23 = _dot(k3, v2)
    f24 = u2k1 * f22
    f25 = -u2k1 * f21
    f21 *= delta
    f22 *= delta
    f23 *= delta
    g1 = f13 * f22
    g2 = f13 * f25 - f15 * f23
    g3 = f11 * f23 - f13 * f21
    g4 = -f13 * f24
    g5 = f11 * f22
    g6 = f11 * f25 - f15 * f21
    g7 = -f15 * f24
    coeffs = [g5 * g5 + g1 * g1 + g3 * g3,
              2 * (g5 * g6 + g1 * g2 + g3 * g4),
              g6 * g6 + 2 * g5 * g7 + g2 * g2 + g4 * g4 - g1 * g1 - g3 * g3,
              2 * (g6 * g7 - g1 * g2 - g3 * g4),
              g7 * g7 - g2 * g2 - g4 * g4]
    with np.errstate(all="ignore"):
        try:
            s0 = solve_quartic(coeffs)
            s = polish_quartic_roots(coeffs, s0)
        except (ZeroDivisionError, OverflowError, ValueError):
            return []
    # cv2 4.13 returns no pose for a complex-conjugate root pair (observed on cv2.solveP3P: the reference's
    # older copy turns the shared real part into two identical poses); the closed form yields such a pair
    # as two exactly equal values
    keep = [True] * 4
    if drop_complex:
        if s0[0] == s0[1]:
            keep[0] = keep[1] = False
        if s0[2] == s0[3]:
            keep[2] = keep[3] = False
    temp = _cross(k1, nl)
    ck1nl = np.array([[k1[0], nl[0], temp[0]], [k1[1], nl[1], temp[1]], [k1[2], nl[2], temp[2]]], F64)
    cb1k3tzT = np.array([b1, k3, tz], F64)
    b3p = b3 * (delta / k3b3)
    out, errs = [], []
    for i in range(4):
        c1 = s[i]
        if not keep[i]:
            continue
        if abs(c1) > 1:                 # NaN roots fall through, as in the C++ comparison
            continue
        st1 = math.sqrt(1 - c1 * c1) if c1 == c1 else float("nan")
        st1 = st1 if k3b3 > 0 else -st1
        ct3 = g1 * c1 + g2
        st3 = g3 * c1 + g4
        with np.errstate(all="ignore"):
            n3 = float(F64(st1) / F64((g5 * c1 + g6) * c1 + g7))
        ct3 *= n3
        st3 *= n3
        c13 = np.array([[ct3, 0, -st3], [st1 * st3, c1, st1 * ct3], [c1 * st3, -st1, c1 * ct3]], F64)
        R = _matmul3(_matmul3(ck1nl, c13), cb1k3tzT)
        rp3 = np.array([w3[0] * R[0, 0] + w3[1] * R[1, 0] + w3[2] * R[2, 0],
                        w3[0] * R[0, 1] + w3[1] * R[1, 1] + w3[2] * R[2, 1],
                        w3[0] * R[0, 2] + w3[1] * R[1, 2] + w3[2] * R[2, 2]], F64)
        t = b3p * st1 - rp3
        # reference stores R as is (row-major R[i][j]); OpenCV's copy stores the transpose -- see solve_ap3p
        out.append((R, t))
        if fourth is not None:
            mu3, mv3, X3, Y3, Z3 = fourth
            Rt = R.T
            X3p = Rt[0, 0] * X3 + Rt[0, 1] * Y3 + Rt[0, 2] * Z3 + t[0]
            Y3p = Rt[1, 0] * X3 + Rt[1, 1] * Y3 + Rt[1, 2] * Z3 + t[1]
            Z3p = Rt[2, 0] * X3 + Rt[2, 1] * Y3 + Rt[2, 2] * Z3 + t[2]
            mu3p = X3p / Z3p
            mv3p = Y3p / Z3p
            errs.append((mu3p - mu3) * (mu3p - mu3) + (mv3p - mv3) * (mv3p - mv3))
    if fourth is not None:
        for i in range(1, len(out)):
            j = i
            while j > 0 and errs[j - 1] > errs[j]:
                errs[j - 1], errs[j] = errs[j], errs[j - 1]
                out[j - 1], out[j] = out[j], out[j - 1]
                j -= 1
    return out


def solve_ap3p(img3, obj3, inv_fx, inv_fy, cx_fx, cy_fy):
    """``solveAp3p`` (ap3p.cpp:282-317): pixel -> unit bearing vectors, then :func:`ap3p_compute_poses`.

    Returns (Rs [n,3,3], ts [n,3]) in the reference's convention: camera point = R^T X + t
    (the reference stores the un-transposed product, ap3p.cpp:259-264).
    """
    fv = np.empty((3, 3), F64)
    for i in range(3):
        mu = inv_fx * float(img3[i][0]) - cx_fx
        mv = inv_fy * float(img3[i][1]) - cy_fy
        nrm = math.sqrt(mu * mu + mv * mv + 1)
        mk = 1.0 / nrm
        fv[0, i], fv[1, i], fv[2, i] = mu * mk, mv * mk, mk
    wp = np.asarray(obj3, F64).T.copy()
    sols = ap3p_compute_poses(fv, wp)
    if not sols:
        return np.zeros((0, 3, 3)), np.zeros((0, 3))
    return np.stack([r for r, _ in sols]), np.stack([t for _, t in sols])


# ------------------------------------------------------------------------------------------
# OpenCV pieces behind cv::solvePnP / cv::solvePnPRansac (cv2 4.13.0)
# ------------------------------------------------------------------------------------------
def rodrigues_to_matrix(rvec):
    """cv::Rodrigues, vector -> matrix: R = c I + (1 - c) r r^T + s [r]x, evaluated term by term."""
    rx, ry, rz = (float(v) for v in np.asarray(rvec, F64).ravel())
    theta = math.sqrt(rx * rx + ry * ry + rz * rz)
    if theta < np.finfo(F64).eps:
        return np.eye(3)
    c, s = math.cos(theta), math.sin(theta)
    c1 = 1.0 - c
    it = 1.0 / theta
    rx, ry, rz = rx * it, ry * it, rz * it
    rrt = [rx * rx, rx * ry, rx * rz, rx * ry, ry * ry, ry * rz, rx * rz, ry * rz, rz * rz]
    r_x = [0.0, -rz, ry, rz, 0.0, -rx, -ry, rx, 0.0]
    eye = [1.0, 0.0, 0.0, 0.0, 1.0, 0.0, 0.0, 0.0, 1.0]
    return np.array([c * eye[k] + c1 * rrt[k] + s * r_x[k] for k in range(9)], F64).reshape(3, 3)


def rodrigues_to_vector(R):
    """cv::Rodrigues, matrix -> vector: R is first re-orthonormalised as U V^T by OpenCV's Jacobi SVD."""
    R = np.asarray(R, F64).reshape(3, 3)
    if not np.all(np.isfinite(R)) or np.any(np.abs(R) >= 100):
        return np.zeros(3)
    w, u, vt = linalg_cv.svd_full(R)
    R = np.empty((3, 3), F64)
    for i in range(3):
        for j in range(3):
            R[i, j] = u[i, 0] * vt[0, j] + u[i, 1] * vt[1, j] + u[i, 2] * vt[2, j]
    r = np.array([R[2, 1] - R[1, 2], R[0, 2] - R[2, 0], R[1, 0] - R[0, 1]], F64)
    s = math.sqrt((r[0] * r[0] + r[1] * r[1] + r[2] * r[2]) * 0.25)
    c = (R[0, 0] + R[1, 1] + R[2, 2] - 1) * 0.5
    c = 1.0 if c > 1.0 else (-1.0 if c < -1.0 else c)
    theta = math.acos(c)
    if s < 1e-5:
        if c > 0:
            return np.zeros(3)
        t = (R[0, 0] + 1) * 0.5
        x = math.sqrt(max(t, 0.0))
        t = (R[1, 1] + 1) * 0.5
        y = math.sqrt(max(t, 0.0)) * (-1.0 if R[0, 1] < 0 else 1.0)
        t = (R[2, 2] + 1) * 0.5
        z = math.sqrt(max(t, 0.0)) * (-1.0 if R[0, 2] < 0 else 1.0)
        if abs(x) < abs(y) and abs(x) < abs(z) and (R[1, 2] > 0) != (y * z > 0):
            z = -z
        theta /= math.sqrt(x * x + y * y + z * z)
        return np.array([x * theta, y * theta, z * theta], F64)
    vth = 1 / (2 * s)
    vth *= theta
    return r * vth


def undistort_points(uv, K, dist, out_f32):
    """cv::undistortPoints without R/P: pixel -> normalised, 5 fixed-point iterations of the
    (k1, k2, p1, p2) model; the result takes the input's depth (float32 inside solvePnPRansac)."""
    fx, fy, cx, cy = float(K[0, 0]), float(K[1, 1]), float(K[0, 2]), float(K[1, 2])
    k = np.zeros(14)
    d = np.asarray(dist, F64).ravel()
    k[:len(d)] = d
    ifx, ify = 1.0 / fx, 1.0 / fy
    out = np.empty((len(uv), 2), F64)
    for i, (u, v) in enumerate(np.asarray(uv)):
        x = (float(u) - cx) * ifx
        y = (float(v) - cy) * ify
        if d.size and np.any(d != 0):
            x0, y0 = x, y
            for _ in range(5):
                r2 = x * x + y * y
                icdist = (1 + ((k[7] * r2 + k[6]) * r2 + k[5]) * r2) / (1 + ((k[4] * r2 + k[1]) * r2 + k[0]) * r2)
                if icdist < 0:
                    x, y = (float(u) - cx) * ifx, (float(v) - cy) * ify
                    break
                dx = 2 * k[2] * x * y + k[3] * (r2 + 2 * x * x) + k[8] * r2 + k[9] * r2 * r2
                dy = k[2] * (r2 + 2 * y * y) + 2 * k[3] * x * y + k[10] * r2 + k[11] * r2 * r2
                x = (x0 - dx) * icdist
                y = (y0 - dy) * icdist
        out[i] = (x, y)
    return out.astype(np.float32) if out_f32 else out


def project_points(obj, rvec, tvec, K, dist):
    """cv::projectPoints (float64 result; callers round to float32 where OpenCV stores float)."""
    R = rodrigues_to_matrix(rvec)
    t = np.asarray(tvec, F64).ravel()
    fx, fy, cx, cy = float(K[0, 0]), float(K[1, 1]), float(K[0, 2]), float(K[1, 2])
    k = np.zeros(14)
    d = np.asarray(dist, F64).ravel()
    k[:len(d)] = d
    obj = np.asarray(obj).astype(F64)
    X, Y, Z = obj[:, 0], obj[:, 1], obj[:, 2]
    # numpy evaluates these element-wise with one rounding per operation, left to right, exactly like the
    # scalar C++ expressions
    with np.errstate(all="ignore"):
        x = R[0, 0] * X + R[0, 1] * Y + R[0, 2] * Z + t[0]
        y = R[1, 0] * X + R[1, 1] * Y + R[1, 2] * Z + t[1]
        z = R[2, 0] * X + R[2, 1] * Y + R[2, 2] * Z + t[2]
        z = np.where(z != 0, 1.0 / np.where(z != 0, z, 1.0), 1.0)
        x = x * z
        y = y * z
        r2 = x * x + y * y
        r4 = r2 * r2
        r6 = r4 * r2
        a1 = 2 * x * y
        a2 = r2 + 2 * x * x
        a3 = r2 + 2 * y * y
        cdist = 1 + k[0] * r2 + k[1] * r4 + k[4] * r6
        icdist2 = 1.0 / (1 + k[5] * r2 + k[6] * r4 + k[7] * r6)
        xd = x * cdist * icdist2 + k[2] * a1 + k[3] * a2 + k[8] * r2 + k[9] * r4
        yd = y * cdist * icdist2 + k[2] * a3 + k[3] * a1 + k[10] * r2 + k[11] * r4
        out = np.stack([xd * fx + cx, yd * fy + cy], axis=1)
    return out


def solve_pnp_ap3p4(obj4, img4, K, dist):
    """cv::solvePnP(..., SOLVEPNP_AP3P) on exactly four points: AP3P on the first three, the pose whose
    re-projection of the fourth is closest wins.  Returns (ok, rvec, tvec)."""
    fx, fy, cx, cy = float(K[0, 0]), float(K[1, 1]), float(K[0, 2]), float(K[1, 2])
    f32 = np.asarray(img4).dtype == np.float32
    und = undistort_points(img4, K, dist, out_f32=f32)
    inv_fx, inv_fy = 1.0 / fx, 1.0 / fy
    cx_fx, cy_fy = cx / fx, cy / fy
    fv = np.empty((3, 3), F64)
    mus = []
    for i in range(4):
        # ap3p::extract_points re-applies K to the normalised point, ap3p::solve removes it again
        pu = float(und[i, 0]) * fx + cx
        pv = float(und[i, 1]) * fy + cy
        mus.append((inv_fx * pu - cx_fx, inv_fy * pv - cy_fy))
    for i in range(3):
        mu, mv = mus[i]
        nrm = math.sqrt(mu * mu + mv * mv + 1)
        mk = 1.0 / nrm
        fv[:, i] = (mu * mk, mv * mk, mk)
    o = np.asarray(obj4)
    wp = np.array([[float(o[j, i]) for j in range(3)] for i in range(3)], F64)
    sols = ap3p_compute_poses(fv, wp, fourth=(mus[3][0], mus[3][1], float(o[3, 0]), float(o[3, 1]), float(o[3, 2])),
                              drop_complex=True)
    if not sols:
        return False, None, None
    R, t = sols[0]
    return True, rodrigues_to_vector(R.T), t.copy()


class PnPCB:
    """PnPRansacCallback with the AP3P minimal solver (model = [rvec | tvec], 3x2)."""
    model_points = 4

    def __init__(self, K, dist):
        self.K = np.asarray(K, F64)
        self.dist = np.asarray(dist, F64).ravel()

    @staticmethod
    def check_subset(ms1, ms2, count):
        return True

    def run_kernel(self, m1, m2):
        ok, rvec, tvec = solve_pnp_ap3p4(m1, m2, self.K, self.dist)
        if not ok:
            return []
        return [np.stack([rvec, tvec], axis=1)]

    def compute_error(self, m1, m2, model):
        model = np.asarray(model, F64).reshape(3, 2)
        proj = project_points(m1, model[:, 0], model[:, 1], self.K, self.dist).astype(np.float32)
        d = m2.astype(np.float32) - proj
        with np.errstate(all="ignore"):
            return (d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1]).astype(np.float32)


# ------------------------------------------------------------------------------------------
# EPnP (Lepetit, Moreno-Noguer, Fua 2009) as OpenCV runs it (calib3d epnp.cpp, cv2 4.13.0)
# ------------------------------------------------------------------------------------------
def _sv_bksb(w, u_cols, v_cols, b):
    """cv::SVBkSb for one right-hand side: x = sum_i v_i (u_i . b) / w_i over w_i > 2 eps sum(w)."""
    nm = len(w)
    thr = 0.0
    for i in range(nm):
        thr += w[i]
    thr *= linalg_cv.DBL_EPSILON * 2
    n = len(v_cols[0])
    x = [0.0] * n
    for i in range(nm):
        wi = w[i]
        if abs(wi) <= thr:
            continue
        wi = 1 / wi
        s = 0.0
        for j in range(len(b)):
            s += u_cols[i][j] * b[j]
        s *= wi
        for j in range(n):
            x[j] = x[j] + s * v_cols[i][j]
    return x


def _solve_svd(A, b):
    """cv::solve(A, b, DECOMP_SVD) for a tall float64 system."""
    w, u, vt = linalg_cv.svd_full(np.asarray(A, F64))
    n = A.shape[1]
    return _sv_bksb([float(v) for v in w], [u[:, i] for i in range(n)], [vt[i] for i in range(n)], [float(v) for v in b])


def _invert_svd3(A):
    """cv::invert(A, DECOMP_SVD) for 3x3: sum_i v_i (u_i / w_i)^T accumulated in singular-value order."""
    w, u, vt = linalg_cv.svd_full(np.asarray(A, F64))
    thr = (w[0] + w[1] + w[2]) * linalg_cv.DBL_EPSILON * 2
    x = np.zeros((3, 3))
    for i in range(3):
        wi = float(w[i])
        if abs(wi) <= thr:
            continue
        wi = 1 / wi
        buf = [u[k, i] * wi for k in range(3)]
        for j in range(3):
            for k in range(3):
                x[j, k] = x[j, k] + vt[i, j] * buf[k]
    return x


def _mul_transposed(M):
    """cv::mulTransposed(M, ., aTa=true) for a small result: every entry is a sequential sum over rows."""
    n = M.shape[1]
    out = np.zeros((n, n))
    for i in range(n):
        for j in range(i, n):
            s = 0.0
            for k in range(M.shape[0]):
                s += M[k, i] * M[k, j]
            out[i, j] = out[j, i] = s
    return out


def _epnp_qr_solve(A, b):
    """epnp::qr_solve: Householder QR with epnp's own pivot-scale loop (which scans rows k..nr-2)."""
    A = np.array(A, F64)
    b = np.array(b, F64)
    nr, nc = A.shape
    A1 = np.zeros(nc)
    A2 = np.zeros(nc)
    for k in range(nc):
        eta = abs(A[k, k])
        for i in range(k + 1, nr):
            elt = abs(A[i - 1, k])
            if eta < elt:
                eta = elt
        if eta == 0:
            return None
        inv_eta = 1.0 / eta
        sum2 = 0.0
        for i in range(k, nr):
            A[i, k] *= inv_eta
            sum2 += A[i, k] * A[i, k]
        sigma = math.sqrt(sum2)
        if A[k, k] < 0:
            sigma = -sigma
        A[k, k] += sigma
        A1[k] = sigma * A[k, k]
        A2[k] = -eta * sigma
        for j in range(k + 1, nc):
            s = 0.0
            for i in range(k, nr):
                s += A[i, k] * A[i, j]
            tau = s / A1[k]
            for i in range(k, nr):
                A[i, j] -= tau * A[i, k]
    for j in range(nc):
        tau = 0.0
        for i in range(j, nr):
            tau += A[i, j] * b[i]
        tau /= A1[j]
        for i in range(j, nr):
            b[i] -= tau * A[i, j]
    x = np.zeros(nc)
    x[nc - 1] = b[nc - 1] / A2[nc - 1]
    for i in range(nc - 2, -1, -1):
        s = 0.0
        for j in range(i + 1, nc):
            s += A[i, j] * x[j]
        x[i] = (b[i] - s) / A2[i]
    return x


def epnp(obj, und, fu=1.0, fv=1.0, uc=0.0, vc=0.0):
    """epnp::compute_pose on normalised image points ``und`` (solvePnP hands EPnP the undistorted points
    together with the real camera matrix, so us = und * f + c).  Returns (R, t)."""
    pws = np.asarray(obj, F64)
    n = len(pws)
    us = np.empty((n, 2))
    for i in range(n):
        us[i, 0] = float(und[i, 0]) * fu + uc
        us[i, 1] = float(und[i, 1]) * fv + vc
    # control points: centroid + principal axes
    cws = np.zeros((4, 3))
    for i in range(n):
        for j in range(3):
            cws[0, j] += pws[i, j]
    for j in range(3):
        cws[0, j] /= n
    pw0 = pws - cws[0]
    w, u, vt = linalg_cv.svd_full(_mul_transposed(pw0))
    uct = u.T
    for i in range(1, 4):
        k = math.sqrt(w[i - 1] / n)
        for j in range(3):
            cws[i, j] = cws[0, j] + k * uct[i - 1, j]
    # barycentric coordinates
    cc = np.zeros((3, 3))
    for i in range(3):
        for j in range(1, 4):
            cc[i, j - 1] = cws[j, i] - cws[0, i]
    ci = _invert_svd3(cc)
    alphas = np.zeros((n, 4))
    for i in range(n):
        p = pws[i]
        for j in range(3):
            alphas[i, 1 + j] = (ci[j, 0] * (p[0] - cws[0, 0]) + ci[j, 1] * (p[1] - cws[0, 1])
                                + ci[j, 2] * (p[2] - cws[0, 2]))
        alphas[i, 0] = 1.0 - alphas[i, 1] - alphas[i, 2] - alphas[i, 3]
    # M and the null space of M^T M
    M = np.zeros((2 * n, 12))
    for i in range(n):
        a = alphas[i]
        for j in range(4):
            M[2 * i, 3 * j] = a[j] * fu
            M[2 * i, 3 * j + 2] = a[j] * (uc - us[i, 0])
            M[2 * i + 1, 3 * j + 1] = a[j] * fv
            M[2 * i + 1, 3 * j + 2] = a[j] * (vc - us[i, 1])
    _, U, _ = linalg_cv.svd_full(_mul_transposed(M))
    ut = U.T
    v = [ut[11], ut[10], ut[9], ut[8]]
    dv = np.zeros((4, 6, 3))
    for i in range(4):
        a, b = 0, 1
        for j in range(6):
            for k in range(3):
                dv[i, j, k] = v[i][3 * a + k] - v[i][3 * b + k]
            b += 1
            if b > 3:
                a += 1
                b = a + 1
    L = np.zeros((6, 10))
    d = lambda p, q: p[0] * q[0] + p[1] * q[1] + p[2] * q[2]
    for i in range(6):
        L[i] = [d(dv[0, i], dv[0, i]), 2.0 * d(dv[0, i], dv[1, i]), d(dv[1, i], dv[1, i]),
                2.0 * d(dv[0, i], dv[2, i]), 2.0 * d(dv[1, i], dv[2, i]), d(dv[2, i], dv[2, i]),
                2.0 * d(dv[0, i], dv[3, i]), 2.0 * d(dv[1, i], dv[3, i]), 2.0 * d(dv[2, i], dv[3, i]),
                d(dv[3, i], dv[3, i])]
    d2 = lambda p, q: (p[0] - q[0]) * (p[0] - q[0]) + (p[1] - q[1]) * (p[1] - q[1]) + (p[2] - q[2]) * (p[2] - q[2])
    rho = [d2(cws[0], cws[1]), d2(cws[0], cws[2]), d2(cws[0], cws[3]),
           d2(cws[1], cws[2]), d2(cws[1], cws[3]), d2(cws[2], cws[3])]

    def betas1():
        b4 = _solve_svd(L[:, [0, 1, 3, 6]], rho)
        if b4[0] < 0:
            b0 = math.sqrt(-b4[0])
            return [b0, -b4[1] / b0, -b4[2] / b0, -b4[3] / b0]
        b0 = math.sqrt(b4[0])
        return [b0, b4[1] / b0, b4[2] / b0, b4[3] / b0]

    def betas2():
        b3 = _solve_svd(L[:, [0, 1, 2]], rho)
        if b3[0] < 0:
            b0 = math.sqrt(-b3[0])
            b1 = math.sqrt(-b3[2]) if b3[2] < 0 else 0.0
        else:
            b0 = math.sqrt(b3[0])
            b1 = math.sqrt(b3[2]) if b3[2] > 0 else 0.0
        if b3[1] < 0:
            b0 = -b0
        return [b0, b1, 0.0, 0.0]

    def betas3():
        b5 = _solve_svd(L[:, [0, 1, 2, 3, 4]], rho)
        if b5[0] < 0:
            b0 = math.sqrt(-b5[0])
            b1 = math.sqrt(-b5[2]) if b5[2] < 0 else 0.0
        else:
            b0 = math.sqrt(b5[0])
            b1 = math.sqrt(b5[2]) if b5[2] > 0 else 0.0
        if b5[1] < 0:
            b0 = -b0
        return [b0, b1, b5[3] / b0, 0.0]

    def gauss_newton(betas):
        bt = list(betas)
        for _ in range(5):
            A = np.zeros((6, 4))
            B = np.zeros(6)
            for i in range(6):
                r = L[i]
                A[i, 0] = 2 * r[0] * bt[0] + r[1] * bt[1] + r[3] * bt[2] + r[6] * bt[3]
                A[i, 1] = r[1] * bt[0] + 2 * r[2] * bt[1] + r[4] * bt[2] + r[7] * bt[3]
                A[i, 2] = r[3] * bt[0] + r[4] * bt[1] + 2 * r[5] * bt[2] + r[8] * bt[3]
                A[i, 3] = r[6] * bt[0] + r[7] * bt[1] + r[8] * bt[2] + 2 * r[9] * bt[3]
                B[i] = rho[i] - (r[0] * bt[0] * bt[0] + r[1] * bt[0] * bt[1] + r[2] * bt[1] * bt[1]
                                 + r[3] * bt[0] * bt[2] + r[4] * bt[1] * bt[2] + r[5] * bt[2] * bt[2]
                                 + r[6] * bt[0] * bt[3] + r[7] * bt[1] * bt[3] + r[8] * bt[2] * bt[3]
                                 + r[9] * bt[3] * bt[3])
            x = _epnp_qr_solve(A, B)
            if x is None:
                break
            for i in range(4):
                bt[i] += x[i]
        return bt

    def r_and_t(bt):
        ccs = np.zeros((4, 3))
        for i in range(4):
            vv = ut[11 - i]
            for j in range(4):
                for k in range(3):
                    ccs[j, k] += bt[i] * vv[3 * j + k]
        pcs = np.zeros((n, 3))
        for i in range(n):
            a = alphas[i]
            for j in range(3):
                pcs[i, j] = a[0] * ccs[0, j] + a[1] * ccs[1, j] + a[2] * ccs[2, j] + a[3] * ccs[3, j]
        if pcs[0, 2] < 0.0:
            ccs = -ccs
            pcs = -pcs
        pc0 = np.zeros(3)
        pw0m = np.zeros(3)
        for i in range(n):
            for j in range(3):
                pc0[j] += pcs[i, j]
                pw0m[j] += pws[i, j]
        for j in range(3):
            pc0[j] /= n
            pw0m[j] /= n
        abt = np.zeros((3, 3))
        for i in range(n):
            for j in range(3):
                for k in range(3):
                    abt[j, k] += (pcs[i, j] - pc0[j]) * (pws[i, k] - pw0m[k])
        _, uu, vvt = linalg_cv.svd_full(abt)
        vv = vvt.T
        R = np.zeros((3, 3))
        for i in range(3):
            for j in range(3):
                R[i, j] = d(uu[i], vv[j])
        det = (R[0, 0] * R[1, 1] * R[2, 2] + R[0, 1] * R[1, 2] * R[2, 0] + R[0, 2] * R[1, 0] * R[2, 1]
               - R[0, 2] * R[1, 1] * R[2, 0] - R[0, 1] * R[1, 0] * R[2, 2] - R[0, 0] * R[1, 2] * R[2, 1])
        if det < 0:
            R[2] = -R[2]
        t = np.array([pc0[0] - d(R[0], pw0m), pc0[1] - d(R[1], pw0m), pc0[2] - d(R[2], pw0m)])
        sum2 = 0.0
        for i in range(n):
            pw = pws[i]
            Xc = d(R[0], pw) + t[0]
            Yc = d(R[1], pw) + t[1]
            inv_zc = 1.0 / (d(R[2], pw) + t[2])
            ue = uc + fu * Xc * inv_zc
            ve = vc + fv * Yc * inv_zc
            sum2 += math.sqrt((us[i, 0] - ue) * (us[i, 0] - ue) + (us[i, 1] - ve) * (us[i, 1] - ve))
        return sum2 / n, R, t

    with np.errstate(all="ignore"):
        cands = [r_and_t(gauss_newton(f())) for f in (betas1, betas2, betas3)]
    N = 0
    if cands[1][0] < cands[0][0]:
        N = 1
    if cands[2][0] < cands[N][0]:
        N = 2
    return cands[N][1], cands[N][2]


def solve_pnp_epnp(obj, img, K, dist):
    """cv::solvePnP(..., SOLVEPNP_EPNP): undistort, EPnP with the camera matrix, Rodrigues."""
    f32 = np.asarray(img).dtype == np.float32
    und = undistort_points(img, K, dist, out_f32=f32)
    R, t = epnp(obj, und, float(K[0, 0]), float(K[1, 1]), float(K[0, 2]), float(K[1, 2]))
    return rodrigues_to_vector(R), t


def solve_pnp_ransac(obj, img, K, dist, iterations=100, reproj_err=8.0, confidence=0.99,
                     samples=None, models=None, trace=None):
    """cv::solvePnPRansac(..., flags = SOLVEPNP_AP3P) as ``cvSolvePnPRansac`` calls it with solverKind 5
    (``MiniCVNative.cpp:93-139``).  Returns (ok, rvec, tvec, inlier indices).

    Points are converted to float32 for the RANSAC loop (OpenCV does so for float64 input); the final pose
    is EPnP over the inliers (float32 values promoted to float64); the inlier list is the RANSAC mask of
    the best minimal model, not re-evaluated under the final pose.
    """
    o32 = np.ascontiguousarray(np.asarray(obj, F64).astype(np.float32))
    i32 = np.ascontiguousarray(np.asarray(img, F64).astype(np.float32))
    n = len(o32)
    cb = PnPCB(K, dist)
    if n < 4:
        return False, None, None, np.zeros(0, np.int32)
    if n == 4:
        ok, rvec, tvec = solve_pnp_ap3p4(o32, i32, cb.K, cb.dist)
        if not ok:
            return False, None, None, np.zeros(0, np.int32)
        return True, rvec, tvec, np.arange(4, dtype=np.int32)
    ok, model, mask = ransac_run(cb, o32, i32, reproj_err, confidence, iterations, samples=samples,
                                 models=models, trace=trace)
    if not ok:
        return False, None, None, np.zeros(0, np.int32)
    idx = np.nonzero(mask)[0].astype(np.int32)
    rvec, tvec = solve_pnp_epnp(o32[idx].astype(F64), i32[idx].astype(F64), cb.K, cb.dist)
    return True, rvec, tvec, idx
