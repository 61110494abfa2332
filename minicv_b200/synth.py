"""Deterministic synthetic workloads for tests and bench.py (definitions: SURVEY.md section 8d).

numpy only; no GPU, no oracle.  Every generator is a pure function of its seed."""
from __future__ import annotations

import numpy as np


def descriptor_pair(seed: int, n_query: int = 8192, n_train: int = 8192, desc_bytes: int = 32,
                    flip_p: float = 0.10, true_frac: float = 0.5):
    """C2 pair: random train set; ``true_frac`` of the queries are copies of distinct train
    rows with every bit flipped w.p. ``flip_p`` (true matches), the rest are fresh random rows
    (nearest neighbours ~95-105 bits away with frequent d1 == d2 ties -> exercises the
    lowest-index tie-break).  Returns (query, train, gt) with gt[i] = source train row or -1."""
    rng = np.random.default_rng(1000 + seed)
    train = rng.integers(0, 256, (n_train, desc_bytes), dtype=np.uint8)
    query = rng.integers(0, 256, (n_query, desc_bytes), dtype=np.uint8)
    gt = np.full(n_query, -1, dtype=np.int64)
    n_true = min(int(n_query * true_frac), n_train)
    if n_true > 0:
        src = rng.permutation(n_train)[:n_true]
        flips = rng.random((n_true, desc_bytes * 8), dtype=np.float32) < flip_p
        noise = np.packbits(flips, axis=1)
        pos = rng.permutation(n_query)[:n_true]
        query[pos] = train[src] ^ noise
        gt[pos] = src
    return query, train, gt


def descriptor_batch(seed0: int, n_pairs: int, n_query: int = 8192, n_train: int = 8192,
                     desc_bytes: int = 32, out_query=None, out_train=None):
    """``n_pairs`` C2 pairs packed back to back.  Returns (query, q_off, train, t_off)."""
    q = out_query if out_query is not None else np.empty((n_pairs * n_query, desc_bytes), np.uint8)
    t = out_train if out_train is not None else np.empty((n_pairs * n_train, desc_bytes), np.uint8)
    for p in range(n_pairs):
        qq, tt, _ = descriptor_pair(seed0 + p, n_query, n_train, desc_bytes)
        q[p * n_query:(p + 1) * n_query] = qq
        t[p * n_train:(p + 1) * n_train] = tt
    q_off = np.arange(n_pairs + 1, dtype=np.int64) * n_query
    t_off = np.arange(n_pairs + 1, dtype=np.int64) * n_train
    return q, q_off, t, t_off


def ragged_descriptor_batch(seed: int, n_pairs: int, desc_bytes: int = 32, max_rows: int = 700):
    """Ragged batch with empty and tiny pairs for edge-case tests."""
    rng = np.random.default_rng(seed)
    nq = rng.integers(0, max_rows, n_pairs)
    nt = rng.integers(0, max_rows, n_pairs)
    nq[0] = 0
    if n_pairs > 1:
        nt[1] = 0
    if n_pairs > 2:
        nt[2] = 1
    if n_pairs > 3:
        nq[3] = 1
    qs, ts = [], []
    for p in range(n_pairs):
        qq, tt, _ = descriptor_pair(seed * 131 + p, int(nq[p]), max(int(nt[p]), 0), desc_bytes)
        qs.append(qq)
        ts.append(tt[:nt[p]])
    q = np.concatenate(qs) if qs else np.empty((0, desc_bytes), np.uint8)
    t = np.concatenate(ts) if ts else np.empty((0, desc_bytes), np.uint8)
    q_off = np.concatenate([[0], np.cumsum(nq)]).astype(np.int64)
    t_off = np.concatenate([[0], np.cumsum(nt)]).astype(np.int64)
    return q, q_off, t, t_off


# ---------------------------------------------------------------------------------------------
# two-view correspondences (C3: homography / fundamental; C4: essential + recoverPose)
# ---------------------------------------------------------------------------------------------
H_GT = np.array([[1.02, 0.08, 12.0], [-0.09, 1.03, -7.0], [-1.6e-4, -0.7e-4, 1.0]])


def homography_pair(seed: int, n: int = 2000, inlier_frac: float = 0.30, noise: float = 0.5,
                    jitter: bool = True):
    """C3 homography pair: src uniform in [0,1000)^2, inliers follow a (per-seed jittered)
    ground-truth H with N(0, noise px) noise, outliers uniform.  Returns (a, b, gt_inlier, H)."""
    rng = np.random.default_rng(2000 + seed)
    H = H_GT.copy()
    if jitter:
        H[:2, :2] += rng.normal(0, 0.01, (2, 2))
        H[:2, 2] += rng.normal(0, 3.0, 2)
        H[2, :2] += rng.normal(0, 2e-5, 2)
    a = rng.uniform(0, 1000, (n, 2))
    ah = np.concatenate([a, np.ones((n, 1))], axis=1) @ H.T
    b = ah[:, :2] / ah[:, 2:3] + rng.normal(0, noise, (n, 2))
    n_in = int(round(n * inlier_frac))
    gt = np.zeros(n, dtype=bool)
    gt[rng.permutation(n)[:n_in]] = True
    b[~gt] = rng.uniform(0, 1000, (int((~gt).sum()), 2))
    return np.ascontiguousarray(a), np.ascontiguousarray(b), gt, H


def rodrigues(axis, angle):
    axis = np.asarray(axis, float)
    axis = axis / np.linalg.norm(axis)
    K = np.array([[0, -axis[2], axis[1]], [axis[2], 0, -axis[0]], [-axis[1], axis[0], 0]])
    return np.eye(3) + np.sin(angle) * K + (1 - np.cos(angle)) * (K @ K)


def two_view_pair(seed: int, n: int = 2000, inlier_frac: float = 0.30, noise: float = 0.3,
                  focal: float = 800.0, pp=(320.0, 240.0), normalized: bool = False):
    """C3(F)/C4(E) pair: 3-D points in [-2,2]^2 x [4,8], second camera rotated 0.15 rad about
    (0.1,1,0.2) and translated along unit(0.8,0.1,0.2); pixel noise sigma; outliers uniform in
    the 640x480 image.  normalized=True returns coordinates already divided by the intrinsics
    (MiniCV's default RecoverPoseConfig{1,(0,0),0.999,0.005}).  Returns (a, b, gt, R, t)."""
    rng = np.random.default_rng(3000 + seed)
    R = rodrigues((0.1, 1.0, 0.2), 0.15 + 0.02 * rng.standard_normal())
    t = np.array([0.8, 0.1, 0.2]) + 0.05 * rng.standard_normal(3)
    t /= np.linalg.norm(t)
    X = np.stack([rng.uniform(-2, 2, n), rng.uniform(-2, 2, n), rng.uniform(4, 8, n)], axis=1)
    X2 = X @ R.T + t
    a = X[:, :2] / X[:, 2:3] * focal + np.asarray(pp)
    b = X2[:, :2] / X2[:, 2:3] * focal + np.asarray(pp)
    a = a + rng.normal(0, noise, (n, 2))
    b = b + rng.normal(0, noise, (n, 2))
    n_in = int(round(n * inlier_frac))
    gt = np.zeros(n, dtype=bool)
    gt[rng.permutation(n)[:n_in]] = True
    n_out = int((~gt).sum())
    b[~gt] = np.stack([rng.uniform(0, 640, n_out), rng.uniform(0, 480, n_out)], axis=1)
    if normalized:
        a = (a - np.asarray(pp)) / focal
        b = (b - np.asarray(pp)) / focal
    return np.ascontiguousarray(a), np.ascontiguousarray(b), gt, R, t


def descriptor_pair_geo(seed: int, n_query: int = 8192, n_train: int = 8192, desc_bytes: int = 32,
                        noise: float = 0.5):
    """C2 pair with geometry for the match -> homography pipeline: descriptors as descriptor_pair;
    every true-match query sits at a position whose image under a (per-seed) homography is its train
    keypoint (+ N(0, noise) px); all other keypoints are uniform in [0, 1000)^2.
    Returns (query, train, gt, kp_query [nq,2], kp_train [nt,2], H)."""
    q, t, gt = descriptor_pair(seed, n_query, n_train, desc_bytes)
    rng = np.random.default_rng(5000 + seed)
    H = H_GT.copy()
    H[:2, :2] += rng.normal(0, 0.01, (2, 2))
    H[:2, 2] += rng.normal(0, 3.0, 2)
    kq = rng.uniform(0, 1000, (n_query, 2))
    kt = rng.uniform(0, 1000, (n_train, 2))
    m = gt >= 0
    ph = np.concatenate([kq[m], np.ones((int(m.sum()), 1))], axis=1) @ H.T
    kt[gt[m]] = ph[:, :2] / ph[:, 2:3] + rng.normal(0, noise, (int(m.sum()), 2))
    return q, t, gt, np.ascontiguousarray(kq), np.ascontiguousarray(kt), H


def pnp_scene(rng, n: int, inlier_frac: float = 0.4, noise: float = 0.5, K=None):
    """PnP problem: 3-D points in [-2,2]^2 x [4,8] seen by a camera rotated by a random Rodrigues
    vector (sigma 0.3 rad) and shifted by N(0, 0.5); pixel noise sigma; the first n - round(n*inlier_frac)
    observations are replaced by uniform outliers in the 640x480 image.  ``rng`` is a numpy Generator
    (or a seed).  Returns (world [n,3], image [n,2], rvec, tvec)."""
    if not isinstance(rng, np.random.Generator):
        rng = np.random.default_rng(5000 + int(rng))
    if K is None:
        K = np.array([[800.0, 0, 320.0], [0, 800.0, 240.0], [0, 0, 1.0]])
    X = np.stack([rng.uniform(-2, 2, n), rng.uniform(-2, 2, n), rng.uniform(4, 8, n)], axis=1)
    rvec = rng.normal(size=3) * 0.3
    tvec = rng.normal(size=3) * 0.5
    ang = float(np.linalg.norm(rvec))
    R = rodrigues(rvec, ang) if ang > 0 else np.eye(3)
    Xc = X @ R.T + tvec
    uv = Xc[:, :2] / Xc[:, 2:3] * np.array([K[0, 0], K[1, 1]]) + np.array([K[0, 2], K[1, 2]])
    uv = uv + rng.normal(0, noise, (n, 2)) if noise > 0 else uv
    n_out = n - int(round(n * inlier_frac))
    if n_out > 0:
        uv[:n_out] = np.stack([rng.uniform(0, 640, n_out), rng.uniform(0, 480, n_out)], axis=1)
    return np.ascontiguousarray(X), np.ascontiguousarray(uv), rvec, tvec


def scale_scene(seed: int, n: int = 2000, noise: float = 1e-3, outlier_frac: float = 0.0):
    """Scale-search problem (``CameraPose.findScaled``): a source camera, a unit-translation
    pose (R, T) as ``recoverPoses2`` returns it, a planted scale, and ``n`` (world point, observation) pairs seen by
    the camera that pose and scale produce; ``outlier_frac`` of the observations are uniform in the image.
    Observations are in the reference's normalised image units (|x|,|y| <= 1) with focal = (-1,-1): the candidate
    equations of ``findScaled`` (``CameraPose.fs:104-113``) agree with ``Camera.project1`` only for that sign.
    Returns (cam_row [14], R [3,3], T [3], world [n,3], obs [n,2], planted_scale)."""
    rng = np.random.default_rng(5000 + seed)
    eye = np.array([0.3, -4.0, 1.2]) + rng.normal(0, 0.2, 3)
    fw = np.array([0.0, 0.0, 0.5]) - eye
    fw /= np.linalg.norm(fw)
    right = np.cross(fw, [0.0, 0.0, 1.0]); right /= np.linalg.norm(right)
    up = np.cross(right, fw); up /= np.linalg.norm(up)
    focal = np.array([-1.0, -1.0])
    axis = rng.normal(0, 1, 3)
    R = rodrigues(axis, rng.uniform(0.05, 0.25))
    T = rng.normal(0, 1, 3); T /= np.linalg.norm(T)
    s = float(rng.uniform(0.2, 0.8))
    basis = np.stack([right, up, -fw], axis=1)                    # columns: camera axes in the world
    loc2 = eye + basis @ (R @ (s * T))
    axes2 = basis @ R                                             # columns: right, up, -forward of the moved camera
    # points in front of the moved camera, inside its field of view
    depth = rng.uniform(2.5, 6.0, n)
    xy = rng.uniform(-0.9, 0.9, (n, 2))
    pc = np.stack([xy[:, 0] * depth / focal[0], xy[:, 1] * depth / focal[1], depth], axis=1)   # (right, up, forward)
    world = loc2 + pc[:, 0:1] * axes2[:, 0] + pc[:, 1:2] * axes2[:, 1] - pc[:, 2:3] * axes2[:, 2]
    obs = xy + rng.normal(0, noise, (n, 2))
    n_out = int(n * outlier_frac)
    if n_out:
        idx = rng.permutation(n)[:n_out]
        obs[idx] = rng.uniform(-1, 1, (n_out, 2))
    cam_row = np.concatenate([eye, fw, up, right, focal])
    return cam_row, R, T, world, obs, s
