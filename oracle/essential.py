"""Essential-matrix RANSAC, decomposeEssentialMat and recoverPose -- CPU oracle.

TEST INFRASTRUCTURE (see oracle/__init__.py).  Restates what the reference's exports do:
  * ``cvRecoverPoses`` (``src/MiniCVNative/MiniCVNative.cpp:165-194``):
    N < 5 -> false; findEssentialMat(pa, pb, focal, pp, RANSAC, prob, thr, mask); copy the
    RANSAC mask to ``ms``; decomposeEssentialMat -> R1, R2, t.
  * ``cvRecoverPose`` (``:197-215``): same RANSAC, ``ms`` = RANSAC mask (copied BEFORE
    recoverPose), then recoverPose(E, a, b, R, t, focal, pp, mask) -> cheirality count.
and OpenCV 4.13.0's algorithms behind those calls (pinned by tests/golden/essential_*.npz):
  * points -> float64, normalised (x - pp) / f as x*(1/f) + (-pp*(1/f)); threshold / f;
    maxIters = 1000 for this overload; 5-point models (oracle/fivepoint.py), Sampson error
    in float64 cast to float32, no checkSubset, no refit.
  * decomposeEssentialMat: SVD (OpenCV Jacobi, oracle/linalg_cv.py), U/Vt flipped to
    det > 0, R1 = U W Vt, R2 = U W^T Vt, t = U[:, 2].  For an exactly rank-2 E the sign of
    U[:, 2] is decided by rounding noise inside OpenCV's SVD, so (R1, R2, t) is only defined
    up to {R1 <-> R2, t -> -t}; tests compare modulo that.
  * recoverPose: 4 candidates (R1,t), (R2,t), (R1,-t), (R2,-t); DLT triangulation of every
    point; keep z*w > 0, z < 50 and, in camera 2, 0 < z < 50; AND with the input mask;
    choose by the >= cascade in that order.
"""
from __future__ import annotations

import numpy as np

from . import fivepoint, linalg_cv
from .ransac import F32, RansacTrace, ransac_run


def normalize_points(p, focal, pp):
    p = np.asarray(p, np.float64).reshape(-1, 2)
    inv = 1.0 / focal
    out = np.empty_like(p)
    out[:, 0] = p[:, 0] * inv + (-pp[0] * inv)
    out[:, 1] = p[:, 1] * inv + (-pp[1] * inv)
    return out


class EssentialCB:
    model_points = 5

    @staticmethod
    def check_subset(ms1, ms2, count):
        return True

    @staticmethod
    def run_kernel(m1, m2):
        return list(fivepoint.five_point(m1, m2))

    @staticmethod
    def compute_error(m1, m2, E):
        """Sampson distance, float64 with OpenCV's operation order, cast to float32."""
        E = np.asarray(E, np.float64).reshape(3, 3)
        x1, y1 = m1[:, 0], m1[:, 1]
        x2, y2 = m2[:, 0], m2[:, 1]
        Ex1_0 = (E[0, 0] * x1 + E[0, 1] * y1) + E[0, 2] * 1.0
        Ex1_1 = (E[1, 0] * x1 + E[1, 1] * y1) + E[1, 2] * 1.0
        Ex1_2 = (E[2, 0] * x1 + E[2, 1] * y1) + E[2, 2] * 1.0
        Etx2_0 = (E[0, 0] * x2 + E[1, 0] * y2) + E[2, 0] * 1.0
        Etx2_1 = (E[0, 1] * x2 + E[1, 1] * y2) + E[2, 1] * 1.0
        x2tEx1 = (x2 * Ex1_0 + y2 * Ex1_1) + 1.0 * Ex1_2
        a = Ex1_0 * Ex1_0
        b = Ex1_1 * Ex1_1
        c = Etx2_0 * Etx2_0
        d = Etx2_1 * Etx2_1
        with np.errstate(divide="ignore", invalid="ignore"):
            return (x2tEx1 * x2tEx1 / (((a + b) + c) + d)).astype(F32)


def find_essential(a, b, focal=1.0, pp=(0.0, 0.0), prob=0.999, thr=1.0, max_iters=1000,
                   samples=None, models=None, fixed_iters=False, trace: RansacTrace | None = None):
    """cv2.findEssentialMat(a, b, focal, pp, RANSAC, prob, thr).  Returns (E or stacked Es, mask)."""
    m1 = normalize_points(a, focal, pp)
    m2 = normalize_points(b, focal, pp)
    n = m1.shape[0]
    if n < 5:
        return None, None
    t = thr / ((focal + focal) / 2)
    ok, E, mask = ransac_run(EssentialCB, m1, m2, t, prob, max_iters, samples, models, fixed_iters, trace)
    if not ok:
        return None, np.zeros(n, np.uint8)
    return E, mask


def _det3(m):
    return float(np.linalg.det(m))


def decompose_essential(E):
    E = np.asarray(E, np.float64).reshape(3, 3)
    _, U, Vt = linalg_cv.svd_full(E)
    if _det3(U) < 0:
        U = U * -1.0
    if _det3(Vt) < 0:
        Vt = Vt * -1.0
    W = np.array([[0.0, 1.0, 0.0], [-1.0, 0.0, 0.0], [0.0, 0.0, 1.0]])
    R1 = U @ W @ Vt
    R2 = U @ W.T @ Vt
    t = U[:, 2].copy()
    return R1, R2, t


def triangulate(P0, P1, x1, x2):
    """DLT triangulation (cv::triangulatePoints): null vector of the 4x4 system per point.
    x1, x2: [N, 2].  Returns Q [4, N]."""
    n = x1.shape[0]
    A = np.empty((n, 4, 4))
    for j, (P, x) in enumerate(((P0, x1), (P1, x2))):
        A[:, 2 * j + 0, :] = x[:, 0:1] * P[2][None, :] - P[0][None, :]
        A[:, 2 * j + 1, :] = x[:, 1:2] * P[2][None, :] - P[1][None, :]
    _, _, vt = np.linalg.svd(A)
    return vt[:, 3, :].T.copy()


def cheirality_mask(R, t, x1, x2, dist=50.0):
    P0 = np.hstack([np.eye(3), np.zeros((3, 1))])
    P1 = np.hstack([R, t.reshape(3, 1)])
    Q = triangulate(P0, P1, x1, x2)
    mask = (Q[2] * Q[3]) > 0
    with np.errstate(divide="ignore", invalid="ignore"):
        Q = Q / Q[3]
        mask &= Q[2] < dist
        Q2 = P1 @ Q
        mask &= Q2[2] > 0
        mask &= Q2[2] < dist
    return mask


def recover_pose(E, a, b, focal=1.0, pp=(0.0, 0.0), mask=None, dist=50.0):
    """cv2.recoverPose(E, a, b, focal=, pp=, mask=).  Returns (count, R, t, mask_out uint8,
    candidate counts [4])."""
    x1 = normalize_points(a, focal, pp)
    x2 = normalize_points(b, focal, pp)
    R1, R2, t = decompose_essential(E)
    cands = [(R1, t), (R2, t), (R1, -t), (R2, -t)]
    masks = [cheirality_mask(R, tt, x1, x2, dist) for R, tt in cands]
    if mask is not None:
        mm = np.asarray(mask).reshape(-1) != 0
        masks = [m & mm for m in masks]
    good = [int(m.sum()) for m in masks]
    g1, g2, g3, g4 = good
    if g1 >= g2 and g1 >= g3 and g1 >= g4:
        k = 0
    elif g2 >= g1 and g2 >= g3 and g2 >= g4:
        k = 1
    elif g3 >= g1 and g3 >= g2 and g3 >= g4:
        k = 2
    else:
        k = 3
    R, tt = cands[k]
    return good[k], R.copy(), tt.copy(), masks[k].astype(np.uint8), good


def cv_recover_pose(cfg, a, b):
    """Reference export ``cvRecoverPose`` (MiniCVNative.cpp:197-215).
    cfg = (focal, (ppx, ppy), prob, thr).  Returns (count, R, t, ms)."""
    focal, pp, prob, thr = cfg
    E, mask = find_essential(a, b, focal, pp, prob, thr)
    n = np.asarray(a).reshape(-1, 2).shape[0]
    if E is None:
        return 0, None, None, np.zeros(n, np.uint8)
    ms = mask.copy()
    cnt, R, t, _, _ = recover_pose(E, a, b, focal, pp, mask)
    return cnt, R, t, ms


def cv_recover_poses(cfg, a, b):
    """Reference export ``cvRecoverPoses`` (MiniCVNative.cpp:165-194).
    Returns (ok, R1, R2, t, ms)."""
    focal, pp, prob, thr = cfg
    n = np.asarray(a).reshape(-1, 2).shape[0]
    if n < 5:
        return False, None, None, None, np.zeros(n, np.uint8)
    E, mask = find_essential(a, b, focal, pp, prob, thr)
    if E is None or E.shape != (3, 3):
        return False, None, None, None, (mask if mask is not None else np.zeros(n, np.uint8))
    R1, R2, t = decompose_essential(E)
    return True, R1, R2, t, mask
