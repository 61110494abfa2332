"""RANSAC driver + homography / fundamental callbacks -- CPU oracle (numpy / pure Python).

TEST INFRASTRUCTURE (see oracle/__init__.py).  The reference exports neither
findHomography nor findFundamentalMat; the essential-matrix call it
does make (``src/MiniCVNative/MiniCVNative.cpp:177,204``: ``findEssentialMat(..., RANSAC,
prob, thr, mask)``) runs through the same OpenCV point-set registrator restated here.

What is restated, per OpenCV 4.13.0 (the pinned third-party dependency; un-vendored, so the
published algorithm is written out below and checked against cv2 outputs frozen in
tests/golden/ransac_*.npz -- the reference itself pins nothing at this boundary):
  * RNG: 64-bit multiply-with-carry, state 0xFFFF_FFFF_FFFF_FFFF at the start of every call;
    ``next(): state = (u32)state * 4164903690 + (state >> 32)``; ``uniform(0,n) = next() % n``.
  * getSubset: draw modelPoints indices sequentially, redraw an index while it duplicates
    an earlier one, then checkSubset; on failure redraw the whole subset (<= 10000 attempts).
  * run(): per accepted subset one iteration; for each returned model count
    ``err <= (float)(thr*thr)``; accept iff ``good > max(best, modelPoints-1)``; on accept
    ``niters = RANSACUpdateNumIters(conf, (N-good)/N, modelPoints, niters)``.
  * ``samples=`` injects the sample index sets instead of the RNG (the parity contract of
    BASELINE.json's north_star), ``fixed_iters`` disables the adaptive stop, ``models=``
    injects model matrices per iteration instead of running the minimal solver.
Scoring functions are bit-exact restatements (float32 for H, float64->float32 for F/E, no
FMA contraction: numpy never fuses).  Minimal solvers use LAPACK and therefore agree with
OpenCV's own (Jacobi) solvers to rounding, not bitwise.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field

import numpy as np

F32 = np.float32
FLT_EPSILON = float(np.finfo(np.float32).eps)
DBL_EPSILON = float(np.finfo(np.float64).eps)
DBL_MIN = float(np.finfo(np.float64).tiny)


class MWC:
    """cv::RNG: multiply-with-carry; ``cv::RNG rng((uint64)-1)`` at the top of every
    RANSACPointSetRegistrator::run call."""
    A = 4164903690

    def __init__(self, state: int = 0xFFFFFFFFFFFFFFFF):
        self.state = state & 0xFFFFFFFFFFFFFFFF

    def next(self) -> int:
        self.state = ((self.state & 0xFFFFFFFF) * self.A + (self.state >> 32)) & 0xFFFFFFFFFFFFFFFF
        return self.state & 0xFFFFFFFF

    def uniform(self, a: int, b: int) -> int:
        return a if a == b else (self.next() % (b - a)) + a


def cv_round(x: float) -> int:
    """cvRound: round half to even (SSE2 cvtsd2si)."""
    return int(np.rint(x))


def update_num_iters(p: float, ep: float, model_points: int, max_iters: int) -> int:
    p = min(max(p, 0.0), 1.0)
    ep = min(max(ep, 0.0), 1.0)
    num = max(1.0 - p, DBL_MIN)
    denom = 1.0 - math.pow(1.0 - ep, model_points)
    if denom < DBL_MIN:
        return 0
    num = math.log(num)
    denom = math.log(denom)
    if denom >= 0 or -num >= max_iters * (-denom):
        return max_iters
    return cv_round(num / denom)


# ------------------------------------------------------------------------------------------
# callbacks
# ------------------------------------------------------------------------------------------
def have_collinear_points(pts: np.ndarray, count: int) -> bool:
    """pts: float32 [count, 2].  Collinearity of the last point vs all earlier pairs."""
    i = count - 1
    p = pts.astype(np.float64)
    for j in range(i):
        dx1 = p[j, 0] - p[i, 0]
        dy1 = p[j, 1] - p[i, 1]
        for k in range(j):
            dx2 = p[k, 0] - p[i, 0]
            dy2 = p[k, 1] - p[i, 1]
            if abs(dx2 * dy1 - dy2 * dx1) <= FLT_EPSILON * (abs(dx1) + abs(dy1) + abs(dx2) + abs(dy2)):
                return True
    return False


def _det3(a):
    return (a[0][0] * (a[1][1] * a[2][2] - a[2][1] * a[1][2])
            - a[0][1] * (a[1][0] * a[2][2] - a[2][0] * a[1][2])
            + a[0][2] * (a[1][0] * a[2][1] - a[2][0] * a[1][1]))


class HomographyCB:
    model_points = 4

    @staticmethod
    def prepare(a, b):
        return np.ascontiguousarray(a, F32).reshape(-1, 2), np.ascontiguousarray(b, F32).reshape(-1, 2)

    @staticmethod
    def check_subset(ms1, ms2, count):
        if have_collinear_points(ms1, count) or have_collinear_points(ms2, count):
            return False
        if count == 4:
            tt = ((0, 1, 2), (1, 2, 3), (0, 2, 3), (0, 1, 3))
            s = ms1.astype(np.float64)
            d = ms2.astype(np.float64)
            negative = 0
            for t in tt:
                A = [[s[t[0], 0], s[t[0], 1], 1.0], [s[t[1], 0], s[t[1], 1], 1.0], [s[t[2], 0], s[t[2], 1], 1.0]]
                B = [[d[t[0], 0], d[t[0], 1], 1.0], [d[t[1], 0], d[t[1], 1], 1.0], [d[t[2], 0], d[t[2], 1], 1.0]]
                negative += 1 if _det3(A) * _det3(B) < 0 else 0
            if negative != 0 and negative != 4:
                return False
        return True

    @staticmethod
    def run_kernel(m1, m2):
        """Normalised DLT: per-axis mean / mean-abs-deviation scaling, 9x9 LtL, eigenvector of
        the smallest eigenvalue, de-normalise, scale so h33 = 1.  Returns list of 3x3."""
        M = m1.astype(np.float64)
        m = m2.astype(np.float64)
        count = M.shape[0]
        cM = M.sum(axis=0) / count
        cm = m.sum(axis=0) / count
        sM = np.abs(M - cM).sum(axis=0)
        sm = np.abs(m - cm).sum(axis=0)
        if (np.abs(sm) < DBL_EPSILON).any() or (np.abs(sM) < DBL_EPSILON).any():
            return []
        sm = count / sm
        sM = count / sM
        inv_hnorm = np.array([[1.0 / sm[0], 0, cm[0]], [0, 1.0 / sm[1], cm[1]], [0, 0, 1]])
        hnorm2 = np.array([[sM[0], 0, -cM[0] * sM[0]], [0, sM[1], -cM[1] * sM[1]], [0, 0, 1]])
        x = (m[:, 0] - cm[0]) * sm[0]
        y = (m[:, 1] - cm[1]) * sm[1]
        X = (M[:, 0] - cM[0]) * sM[0]
        Y = (M[:, 1] - cM[1]) * sM[1]
        one = np.ones(count)
        zero = np.zeros(count)
        Lx = np.stack([X, Y, one, zero, zero, zero, -x * X, -x * Y, -x], axis=1)
        Ly = np.stack([zero, zero, zero, X, Y, one, -y * X, -y * Y, -y], axis=1)
        LtL = Lx.T @ Lx + Ly.T @ Ly
        w, v = np.linalg.eigh(LtL)
        h0 = v[:, 0].reshape(3, 3)
        H = inv_hnorm @ h0 @ hnorm2
        if H[2, 2] == 0:
            return []
        return [H * (1.0 / H[2, 2])]

    @staticmethod
    def compute_error(m1, m2, H):
        """float32 reprojection error, separate multiplies and adds, in the order
        ((h0*x + h1*y) + h2) etc. (no FMA)."""
        Hf = np.asarray(H, np.float64).reshape(-1)[:8].astype(F32)
        Mx, My = m1[:, 0], m1[:, 1]
        one = F32(1.0)
        ww = one / ((Hf[6] * Mx + Hf[7] * My) + one)
        dx = ((Hf[0] * Mx + Hf[1] * My) + Hf[2]) * ww - m2[:, 0]
        dy = ((Hf[3] * Mx + Hf[4] * My) + Hf[5]) * ww - m2[:, 1]
        return (dx * dx + dy * dy).astype(F32)


class FundamentalCB:
    """7-point fundamental matrix callback (FM_RANSAC)."""
    model_points = 7

    @staticmethod
    def prepare(a, b):
        return np.ascontiguousarray(a, F32).reshape(-1, 2), np.ascontiguousarray(b, F32).reshape(-1, 2)

    @staticmethod
    def check_subset(ms1, ms2, count):
        return not (have_collinear_points(ms1, count) or have_collinear_points(ms2, count))

    @staticmethod
    def run_kernel(m1, m2):
        """7-point: null space of the 7x9 system (raw pixel coordinates, no normalisation),
        det(lambda*f1 + (1-lambda)*f2) = 0 cubic, each real root scaled so F33 = 1."""
        a = np.zeros((7, 9))
        x0, y0 = m1[:, 0].astype(np.float64), m1[:, 1].astype(np.float64)
        x1, y1 = m2[:, 0].astype(np.float64), m2[:, 1].astype(np.float64)
        a[:, 0] = x1 * x0; a[:, 1] = x1 * y0; a[:, 2] = x1
        a[:, 3] = y1 * x0; a[:, 4] = y1 * y0; a[:, 5] = y1
        a[:, 6] = x0; a[:, 7] = y0; a[:, 8] = 1.0
        _, _, vt = np.linalg.svd(a, full_matrices=True)
        f1 = vt[7].copy()
        f2 = vt[8].copy()
        f1 -= f2
        # det(f2 + l*f1) = c3 + c2 l + c1 l^2 + c0 l^3  (OpenCV stores c[0] as the cubic term)
        F1 = f1.reshape(3, 3)
        F2 = f2.reshape(3, 3)

        def det(m):
            return _det3(m.tolist())
        c3 = det(F2)
        c0 = det(F1)
        # exact polynomial identification from two more evaluations
        p1 = det(F2 + F1)
        m1_ = det(F2 - F1)
        c1 = (p1 + m1_) / 2.0 - c3          # l^2 coefficient
        c2 = (p1 - m1_) / 2.0 - c0          # l^1 coefficient
        roots = solve_cubic(c0, c1, c2, c3)
        out = []
        for r in roots:
            lam = r
            mu = 1.0
            s = f1[8] * r + f2[8]
            F = np.zeros(9)
            if abs(s) > DBL_EPSILON:
                mu = 1.0 / s
                lam *= mu
                F[8] = 1.0
            else:
                F[8] = 0.0
            F[:8] = f1[:8] * lam + f2[:8] * mu
            out.append(F.reshape(3, 3))
        return out

    @staticmethod
    def compute_error(m1, m2, F):
        F = np.asarray(F, np.float64).reshape(-1)
        x1 = m1[:, 0].astype(np.float64); y1 = m1[:, 1].astype(np.float64)
        x2 = m2[:, 0].astype(np.float64); y2 = m2[:, 1].astype(np.float64)
        a = (F[0] * x1 + F[1] * y1) + F[2]
        b = (F[3] * x1 + F[4] * y1) + F[5]
        c = (F[6] * x1 + F[7] * y1) + F[8]
        s2 = 1.0 / (a * a + b * b)
        d2 = (x2 * a + y2 * b) + c
        a = (F[0] * x2 + F[3] * y2) + F[6]
        b = (F[1] * x2 + F[4] * y2) + F[7]
        c = (F[2] * x2 + F[5] * y2) + F[8]
        s1 = 1.0 / (a * a + b * b)
        d1 = (x1 * a + y1 * b) + c
        return np.maximum(d1 * d1 * s1, d2 * d2 * s2).astype(F32)


def solve_cubic(a, b, c, d):
    """Real roots of a x^3 + b x^2 + c x + d (cv::solveCubic's closed form)."""
    if a == 0:
        if b == 0:
            if c == 0:
                return []
            return [-d / c]
        disc = c * c - 4 * b * d
        if disc < 0:
            return []
        disc = math.sqrt(disc)
        q = 1.0 / (2 * b)
        return [(-c + disc) * q, (-c - disc) * q]
    a1 = b / a
    a2 = c / a
    a3 = d / a
    Q = (a1 * a1 - 3 * a2) * (1.0 / 9)
    R = (2 * a1 * a1 * a1 - 9 * a1 * a2 + 27 * a3) * (1.0 / 54)
    Qcubed = Q * Q * Q
    dd = Qcubed - R * R
    if dd > 0:
        theta = math.acos(max(-1.0, min(1.0, R / math.sqrt(Qcubed))))
        sq = math.sqrt(Q)
        t0 = -2 * sq
        t1 = theta * (1.0 / 3)
        t2 = a1 * (1.0 / 3)
        return [t0 * math.cos(t1) - t2, t0 * math.cos(t1 + 2.0 * math.pi / 3) - t2,
                t0 * math.cos(t1 + 4.0 * math.pi / 3) - t2]
    if dd == 0:
        if R >= 0:
            sq = -math.pow(R, 1.0 / 3)
        else:
            sq = math.pow(-R, 1.0 / 3)
        return [2 * sq - a1 / 3, -sq - a1 / 3]
    e = math.pow(math.sqrt(-dd) + abs(R), 1.0 / 3)
    if R > 0:
        e = -e
    return [(e + Q / e) - a1 * (1.0 / 3)]


# ------------------------------------------------------------------------------------------
# driver
# ------------------------------------------------------------------------------------------
@dataclass
class RansacTrace:
    samples: list = field(default_factory=list)      # per iteration: index array [modelPoints]
    models: list = field(default_factory=list)       # per iteration: list of 3x3
    counts: list = field(default_factory=list)       # per iteration: list of inlier counts
    best_iter: int = -1
    best_model_idx: int = -1
    niters_final: int = 0


def find_inliers(cb, m1, m2, model, thr):
    err = cb.compute_error(m1, m2, model)
    t = F32(thr * thr)
    mask = (err <= t)
    return int(mask.sum()), mask.astype(np.uint8)


def get_subset(cb, m1, m2, rng: MWC, max_attempts=10000):
    count = m1.shape[0]
    mp = cb.model_points
    for _ in range(max_attempts):
        idx = []
        for i in range(mp):
            v = rng.uniform(0, count)
            while v in idx:
                v = rng.uniform(0, count)
            idx.append(v)
        if cb.check_subset(m1[idx], m2[idx], mp):
            return np.array(idx, dtype=np.int32)
    return None


def ransac_run(cb, m1, m2, thr, conf, max_iters, samples=None, models=None, fixed_iters=False,
               trace: RansacTrace | None = None):
    """RANSACPointSetRegistrator::run.  Returns (ok, model, mask[N] uint8)."""
    count = m1.shape[0]
    mp = cb.model_points
    niters = max(max_iters, 1)
    if count < mp:
        return False, None, None
    best_mask = np.zeros(count, np.uint8)
    if count == mp:
        ms = cb.run_kernel(m1, m2)
        if len(ms) <= 0:
            return False, None, None
        return True, np.vstack(ms), np.ones(count, np.uint8)
    rng = MWC()
    max_good = 0
    best_model = None
    it = 0
    while it < niters:
        if samples is not None:
            if it >= len(samples):
                break
            idx = np.asarray(samples[it], dtype=np.int64)
        else:
            idx = get_subset(cb, m1, m2, rng)
            if idx is None:
                if it == 0:
                    return False, None, None
                break
        ms = models[it] if models is not None else cb.run_kernel(m1[idx], m2[idx])
        if trace is not None:
            trace.samples.append(np.asarray(idx, np.int32))
            trace.models.append([np.array(m, np.float64) for m in ms])
            trace.counts.append([])
        for mi, model in enumerate(ms):
            good, mask = find_inliers(cb, m1, m2, model, thr)
            if trace is not None:
                trace.counts[-1].append(good)
            if good > max(max_good, mp - 1):
                best_mask = mask
                best_model = np.array(model, np.float64)
                max_good = good
                if trace is not None:
                    trace.best_iter, trace.best_model_idx = it, mi
                if not fixed_iters:
                    niters = update_num_iters(conf, (count - good) / count, mp, niters)
        it += 1
    if trace is not None:
        trace.niters_final = it
    if max_good > 0:
        return True, best_model, best_mask
    return False, None, None


# ------------------------------------------------------------------------------------------
# homography: RANSAC + refit (DLT on inliers + 10 LM iterations) + mask from the refined H
# ------------------------------------------------------------------------------------------
def _h_refine_compute(h, M, m, want_j):
    Mx = M[:, 0].astype(np.float64); My = M[:, 1].astype(np.float64)
    ww = h[6] * Mx + h[7] * My + 1.0
    ww = np.where(np.abs(ww) > DBL_EPSILON, 1.0 / ww, 0.0)
    xi = (h[0] * Mx + h[1] * My + h[2]) * ww
    yi = (h[3] * Mx + h[4] * My + h[5]) * ww
    n = M.shape[0]
    err = np.empty(2 * n)
    err[0::2] = xi - m[:, 0]
    err[1::2] = yi - m[:, 1]
    if not want_j:
        return err, None
    J = np.zeros((2 * n, 8))
    J[0::2, 0] = Mx * ww; J[0::2, 1] = My * ww; J[0::2, 2] = ww
    J[0::2, 6] = -Mx * ww * xi; J[0::2, 7] = -My * ww * xi
    J[1::2, 3] = Mx * ww; J[1::2, 4] = My * ww; J[1::2, 5] = ww
    J[1::2, 6] = -Mx * ww * yi; J[1::2, 7] = -My * ww * yi
    return err, J


def _solve_eig(A, b):
    """cv::solve(A, b, x, DECOMP_EIG): symmetric eigen-decomposition pseudo-solve."""
    w, v = np.linalg.eigh(A)
    thr = DBL_EPSILON * 2 * np.abs(w).sum()  # mirrors SVBkSb's threshold on the spectrum
    inv = np.where(np.abs(w) > thr, 1.0 / np.where(w == 0, 1.0, w), 0.0)
    return v @ (inv * (v.T @ b))


def lm_refine_homography(H, M, m, max_iters=10, eps=FLT_EPSILON):
    """cv::LMSolver (calib3d levmarq.cpp) on the 8 free parameters of H (h33 fixed at 1)."""
    x = np.asarray(H, np.float64).reshape(-1)[:8].copy()
    r, J = _h_refine_compute(x, M, m, True)
    S = float(r @ r)
    A = J.T @ J
    v = J.T @ r
    D = np.diag(A).copy()
    Rlo, Rhi = 0.25, 0.75
    lam, lc = 1.0, 0.75
    it = 0
    while True:
        Ap = A + np.diag(lam * D)
        d = _solve_eig(Ap, v)
        xd = x - d
        rd, _ = _h_refine_compute(xd, M, m, False)
        Sd = float(rd @ rd)
        temp_d = -(A @ d) + 2.0 * v
        dS = float(d @ temp_d)
        R = (S - Sd) / (dS if abs(dS) > DBL_EPSILON else 1.0)
        if R > Rhi:
            lam *= 0.5
            if lam < lc:
                lam = 0.0
        elif R < Rlo:
            t = float(d @ v)
            nu = (Sd - S) / (t if abs(t) > DBL_EPSILON else 1.0) + 2.0
            nu = min(max(nu, 2.0), 10.0)
            if lam == 0:
                Ainv = np.linalg.pinv(A)
                maxval = max(DBL_EPSILON, float(np.abs(np.diag(Ainv)).max()))
                lam = lc = 1.0 / maxval
                nu *= 0.5
            lam *= nu
        if Sd < S:
            S = Sd
            x = xd
            r, J = _h_refine_compute(x, M, m, True)
            A = J.T @ J
            v = J.T @ r
        it += 1
        proceed = it < max_iters and np.abs(d).max() >= eps and np.abs(r).max() >= eps
        if not proceed:
            break
    return np.concatenate([x, [1.0]]).reshape(3, 3)


def find_homography(a, b, thr=3.0, max_iters=2000, conf=0.995, samples=None, models=None,
                    fixed_iters=False, trace=None, remask=True):
    """cv2.findHomography(a, b, RANSAC, thr, maxIters=max_iters, confidence=conf).
    Returns (H 3x3 float64 or None, mask[N] uint8)."""
    m1, m2 = HomographyCB.prepare(a, b)
    n = m1.shape[0]
    if n < 4:
        return None, None
    if n == 4:
        ms = HomographyCB.run_kernel(m1, m2)
        if not ms:
            return None, None
        return ms[0], np.ones(n, np.uint8)
    ok, H, mask = ransac_run(HomographyCB, m1, m2, thr, conf, max_iters, samples, models,
                             fixed_iters, trace)
    if not ok:
        return None, np.zeros(n, np.uint8)
    inl = mask.astype(bool)
    s, d = m1[inl], m2[inl]
    if s.shape[0] > 0:
        ms = HomographyCB.run_kernel(s, d)
        if ms:
            H = ms[0]
        H = lm_refine_homography(H, s, d, 10)
        if remask:
            _, mask = find_inliers(HomographyCB, m1, m2, H, thr)
    if H[2, 2] != 0:
        H = H * (1.0 / H[2, 2])
    return H, mask


def find_fundamental(a, b, thr=3.0, conf=0.99, max_iters=1000, samples=None, models=None,
                     fixed_iters=False, trace=None):
    """cv2.findFundamentalMat(a, b, FM_RANSAC, thr, conf, max_iters) for N >= 8: best 7-point
    model, no refit, no mask recompute.  Returns (F or None, mask)."""
    m1, m2 = FundamentalCB.prepare(a, b)
    n = m1.shape[0]
    if n < 7:
        return None, None
    ok, F, mask = ransac_run(FundamentalCB, m1, m2, thr, conf, max_iters, samples, models,
                             fixed_iters, trace)
    if not ok:
        return None, np.zeros(n, np.uint8)
    return F, mask


# ------------------------------------------------------------------------------------------
# LMedS (LMeDSPointSetRegistrator::run) and cv::findFundamentalMat's method dispatch
# ------------------------------------------------------------------------------------------
def lmeds_run(cb, m1, m2, conf, max_iters, trace: RansacTrace | None = None):
    """OpenCV's LMeDSPointSetRegistrator::run (calib3d/src/ptsetreg.cpp).  The iteration count is fixed up front
    from an assumed outlier ratio of 0.45; every model's score is the MEDIAN (element count/2 of the sorted float32
    errors); the first strictly smaller median wins; the mask comes from sigma = 2.5 * 1.4826 * (1 + 5 / (N - mp)) *
    sqrt(median), floored at 0.001.  Returns (ok, model, mask)."""
    count = m1.shape[0]
    mp = cb.model_points
    if count < mp:
        return False, None, None
    if count == mp:
        ms = cb.run_kernel(m1, m2)
        if len(ms) <= 0:
            return False, None, None
        return True, np.vstack(ms), np.ones(count, np.uint8)
    niters = update_num_iters(conf, 0.45, mp, max_iters)
    rng = MWC()
    min_median = np.inf
    best_model = None
    for it in range(niters):
        idx = get_subset(cb, m1, m2, rng)
        if idx is None:
            if it == 0:
                return False, None, None
            break
        ms = cb.run_kernel(m1[idx], m2[idx])
        if trace is not None:
            trace.samples.append(np.asarray(idx, np.int32))
            trace.models.append([np.array(m, np.float64) for m in ms])
        for mi, model in enumerate(ms):
            err = np.asarray(cb.compute_error(m1, m2, model), np.float32)
            median = float(np.sort(err)[count // 2])
            if median < min_median:
                min_median = median
                best_model = np.array(model, np.float64)
                if trace is not None:
                    trace.best_iter, trace.best_model_idx = it, mi
    if best_model is None:
        return False, None, None
    sigma = 2.5 * 1.4826 * (1 + 5.0 / (count - mp)) * np.sqrt(min_median)
    sigma = max(sigma, 0.001)
    good, mask = find_inliers(cb, m1, m2, best_model, sigma)
    return good >= mp, best_model, mask


def cv_find_fundamental(a, b, thr=3.0, conf=0.99, max_iters=1000):
    """cv2.findFundamentalMat(a, b, cv2.FM_RANSAC, thr, conf, max_iters) with the dispatch of calib3d/src/fundam.cpp:
    N < 7 -> nothing; N == 7 -> the 7-point kernel directly (up to three stacked models, mask all ones);
    8 <= N < 15 -> LMedS; N >= 15 -> RANSAC; thr <= 0 -> 3; confidence outside (eps, 1 - eps) -> 0.99.
    Returns (F [3k x 3] or None, mask or None)."""
    m1, m2 = FundamentalCB.prepare(a, b)
    n = m1.shape[0]
    if n < 7:
        return None, None
    if thr <= 0:
        thr = 3.0
    eps = np.finfo(np.float64).eps
    if conf < eps or conf > 1 - eps:
        conf = 0.99
    if n == 7:
        ms = FundamentalCB.run_kernel(m1, m2)
        if len(ms) <= 0:
            return None, None
        return np.vstack(ms), np.ones(n, np.uint8)
    if n >= 15:
        ok, F, mask = ransac_run(FundamentalCB, m1, m2, thr, conf, max_iters)
    else:
        ok, F, mask = lmeds_run(FundamentalCB, m1, m2, conf, max_iters)
    if not ok:
        return None, (np.zeros(n, np.uint8) if mask is None else mask)
    return F, mask
