"""OpenCV's small-matrix SVD (one-sided Jacobi, ``JacobiSVDImpl_`` in core/src/lapack.cpp),
restated in pure Python for the tiny matrices of this path (5x9, 3x3, 4x4).

TEST INFRASTRUCTURE (see oracle/__init__.py).  Why restate instead of calling LAPACK: the
reference's outputs depend on the *sign and basis conventions* of this SVD --
``cvFivePoint`` takes rows 5..8 of Vt of a 5x9 matrix (``fivepoint.cpp:251-254``), which
OpenCV fabricates from a fixed pseudo-random sequence (RNG 0x12345678) orthogonalised
against the row space, and ``decomposeEssentialMat`` (``MiniCVNative.cpp:186``) returns
R1/R2/t whose order and sign follow U and Vt of the 3x3 SVD.  Pinned against
cv2.SVDecomp in tests/test_oracle_linalg.py.
"""
from __future__ import annotations

import math

import numpy as np

DBL_EPSILON = float(np.finfo(np.float64).eps)
DBL_MIN = float(np.finfo(np.float64).tiny)


class _RNG:
    def __init__(self, state):
        self.state = state

    def next(self):
        self.state = ((self.state & 0xFFFFFFFF) * 4164903690 + (self.state >> 32)) & 0xFFFFFFFFFFFFFFFF
        return self.state & 0xFFFFFFFF


def _jacobi_svd(At: np.ndarray, m: int, n: int, n1: int):
    """At: n1 x m array whose first n rows hold the n vectors (length m) to orthogonalise.
    Returns (W[n], At (rows = left vectors, n1 of them), Vt[n x n])."""
    At = At.copy()
    eps = DBL_EPSILON * 10
    minval = DBL_MIN
    W = np.zeros(n)
    Vt = np.eye(n)
    for i in range(n):
        W[i] = float(np.dot(At[i], At[i]))
    max_iter = max(m, 30)
    for _ in range(max_iter):
        changed = False
        for i in range(n - 1):
            for j in range(i + 1, n):
                Ai, Aj = At[i], At[j]
                a, b = W[i], W[j]
                p = 0.0
                for k in range(m):
                    p += Ai[k] * Aj[k]
                if abs(p) <= eps * math.sqrt(a * b):
                    continue
                p *= 2
                beta = a - b
                gamma = math.hypot(p, beta)
                if beta < 0:
                    delta = (gamma - beta) * 0.5
                    s = math.sqrt(delta / gamma)
                    c = p / (gamma * s * 2)
                else:
                    c = math.sqrt((gamma + beta) / (gamma * 2))
                    s = p / (gamma * c * 2)
                a = b = 0.0
                for k in range(m):
                    t0 = c * Ai[k] + s * Aj[k]
                    t1 = -s * Ai[k] + c * Aj[k]
                    Ai[k] = t0
                    Aj[k] = t1
                    a += t0 * t0
                    b += t1 * t1
                W[i] = a
                W[j] = b
                changed = True
                Vi, Vj = Vt[i], Vt[j]
                for k in range(n):
                    t0 = c * Vi[k] + s * Vj[k]
                    t1 = -s * Vi[k] + c * Vj[k]
                    Vi[k] = t0
                    Vj[k] = t1
        if not changed:
            break
    for i in range(n):
        sd = 0.0
        for k in range(m):
            sd += At[i, k] * At[i, k]
        W[i] = math.sqrt(sd)
    for i in range(n - 1):
        j = i
        for k in range(i + 1, n):
            if W[j] < W[k]:
                j = k
        if i != j:
            W[i], W[j] = W[j], W[i]
            At[[i, j]] = At[[j, i]]
            Vt[[i, j]] = Vt[[j, i]]
    rng = _RNG(0x12345678)
    for i in range(n1):
        sd = W[i] if i < n else 0.0
        ii = 0
        while ii < 100 and sd <= minval:
            val0 = 1.0 / m
            for k in range(m):
                At[i, k] = val0 if (rng.next() & 256) != 0 else -val0
            for _it in range(2):
                for j in range(i):
                    sd = 0.0
                    for k in range(m):
                        sd += At[i, k] * At[j, k]
                    asum = 0.0
                    for k in range(m):
                        t = At[i, k] - sd * At[j, k]
                        At[i, k] = t
                        asum += abs(t)
                    asum = 1.0 / asum if asum > eps * 100 else 0.0
                    for k in range(m):
                        At[i, k] *= asum
            sd = 0.0
            for k in range(m):
                sd += At[i, k] * At[i, k]
            sd = math.sqrt(sd)
            ii += 1
        s = 1.0 / sd if sd > minval else 0.0
        for k in range(m):
            At[i, k] *= s
    return W, At, Vt


def svd_full(A: np.ndarray):
    """cv::SVD::compute(A, w, u, vt, FULL_UV) for a small float64 matrix.
    Returns (w[min(m,n)], u[m x m], vt[n x n])."""
    A = np.asarray(A, np.float64)
    m, n = A.shape
    if m >= n:
        # At = A^T (n rows of length m); left vectors -> U^T rows, Vt accumulates rotations
        At = np.zeros((m, m))
        At[:n, :] = A.T
        W, Ut, Vt = _jacobi_svd(At, m, n, m)
        return W, Ut.T.copy(), Vt
    # m < n: decompose A^T: the n-vectors are the m rows of A
    At = np.zeros((n, n))
    At[:m, :] = A
    W, Vt_full, Ut_small = _jacobi_svd(At, n, m, n)
    # roles swap: "left vectors" of A^T are A's right singular vectors (rows of Vt);
    # the accumulated rotation matrix is U^T of A, returned transposed
    return W, Ut_small.T.copy(), Vt_full


def solve_z(A: np.ndarray) -> np.ndarray:
    """cv::SVD::solveZ: right singular vector of the smallest singular value (last row of Vt)."""
    _, _, vt = svd_full(A)
    return vt[-1].copy()
