"""CPU oracle for ``CameraPose.findScaled``.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).

Restates, line by line, the managed code of the reference:
  * ``CameraPose.findScaled``        ``src/MiniCV/CameraPose.fs:39-134``
  * ``CameraPose.scale`` / ``transformation``   ``CameraPose.fs:30-37``
  * ``Camera.viewTrafo`` / ``transformedView`` / ``project1``   ``src/MiniCV/Camera.fs:36-41, 61-84``
with the Aardvark.Base value types they use written out (``Trafo3d.FromBasis``, ``M44d`` product,
``TransformPos`` / ``TransformDir``, ``Vec.normalize``, ``Fun.IsTiny``).

PARITY UNPINNED: the reference holds no test or fixture for this function and F# cannot run in this image
(no ``dotnet``), so this restatement is checked only for internal consistency (tests/test_oracle_camera_pose.py:
the recovered scale reproduces a planted one, the rigid-basis inverse agrees with ``numpy.linalg.inv``).  The one
numerical liberty: ``Trafo3d.Inverse`` is Aardvark.Base's LU inverse (un-vendored NuGet dependency,
``paket.dependencies``: Aardvark.Base, no pinned source here); it is restated as plain partial-pivot Gaussian
elimination + one forward/back substitution per unit column (``lu_inverse4``), so the view-space coordinates that
generate the candidate scales may differ from .NET's in the last bits.  Everything downstream of the candidate
scales (the cost of a given scale) involves no inverse.
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import numpy as np

F64 = np.float64


@dataclass
class Camera:
    """``Camera`` record (Camera.fs:7-14)."""
    location: np.ndarray
    forward: np.ndarray
    up: np.ndarray
    right: np.ndarray
    focal: np.ndarray            # V2d

    @staticmethod
    def look_at(eye, center, sky, f):
        """``Camera.lookAt`` (Camera.fs:94-105)."""
        eye, center, sky = (np.asarray(v, F64) for v in (eye, center, sky))
        fw = _normalize(center - eye)
        r = _normalize(np.cross(fw, sky))
        u = _normalize(np.cross(r, fw))
        return Camera(eye, fw, u, r, np.asarray(f, F64))


@dataclass
class CameraPose:
    """``CameraPose`` struct (CameraPose.fs:7-16)."""
    RotationIndex: int = 0
    ScaleSign: int = 0
    Rotation: np.ndarray = None
    Translation: np.ndarray = None
    IsInverse: bool = False


def _normalize(v):
    v = np.asarray(v, F64)
    return v / math.sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2])


def _from_basis(x, y, z, p):
    """Forward matrix of ``Trafo3d.FromBasis`` / ``M44d.FromBasis``: columns x, y, z, origin p."""
    m = np.eye(4)
    m[:3, 0], m[:3, 1], m[:3, 2], m[:3, 3] = x, y, z, p
    return m


def _mul44(a, b):
    """``M44d * M44d``: each entry a_i0 b_0j + a_i1 b_1j + a_i2 b_2j + a_i3 b_3j, left to right."""
    r = np.empty((4, 4))
    for i in range(4):
        for j in range(4):
            r[i, j] = a[i, 0] * b[0, j] + a[i, 1] * b[1, j] + a[i, 2] * b[2, j] + a[i, 3] * b[3, j]
    return r


def _transform_pos(m, p):
    return np.array([m[i, 0] * p[0] + m[i, 1] * p[1] + m[i, 2] * p[2] + m[i, 3] for i in range(3)])


def _transform_dir(m, v):
    return np.array([m[i, 0] * v[0] + m[i, 1] * v[1] + m[i, 2] * v[2] for i in range(3)])


def _mat_vec(m, v):
    return np.array([m[i, 0] * v[0] + m[i, 1] * v[1] + m[i, 2] * v[2] for i in range(3)])


def scale(f: float, pose: CameraPose) -> CameraPose:
    """``CameraPose.scale`` (CameraPose.fs:35-37)."""
    f = float(f)
    s = (1 if f > 0 else 0) - (1 if f < 0 else 0)
    return CameraPose(pose.RotationIndex, s * pose.ScaleSign, pose.Rotation, f * np.asarray(pose.Translation, F64),
                      pose.IsInverse)


def transformation(pose: CameraPose) -> np.ndarray:
    """``CameraPose.transformation`` (CameraPose.fs:30-34): forward matrix [R | R t]."""
    m = np.asarray(pose.Rotation, F64)
    return _from_basis(m[:, 0], m[:, 1], m[:, 2], _mat_vec(m, np.asarray(pose.Translation, F64)))


def lu_inverse4(a: np.ndarray) -> np.ndarray:
    """4x4 inverse by LU factorisation with partial pivoting (largest |pivot| in the column, first wins), L stored
    as multipliers, then L y = P e_c and U x = y for each unit column c.  Singular (pivot exactly 0) -> zeros, the
    value Aardvark's ``M44d.Inverse`` documents for a non-invertible matrix."""
    lu = np.array(a, F64).reshape(4, 4).copy()
    perm = [0, 1, 2, 3]
    for k in range(4):
        pi, best = k, abs(lu[k, k])
        for i in range(k + 1, 4):
            if abs(lu[i, k]) > best:
                pi, best = i, abs(lu[i, k])
        if best == 0.0:
            return np.zeros((4, 4))
        if pi != k:
            lu[[k, pi]] = lu[[pi, k]]
            perm[k], perm[pi] = perm[pi], perm[k]
        for i in range(k + 1, 4):
            lu[i, k] = lu[i, k] / lu[k, k]
            for j in range(k + 1, 4):
                lu[i, j] = lu[i, j] - lu[i, k] * lu[k, j]
    inv = np.zeros((4, 4))
    for c in range(4):
        y = [0.0] * 4
        for i in range(4):
            v = 1.0 if perm[i] == c else 0.0
            for j in range(i):
                v = v - lu[i, j] * y[j]
            y[i] = v
        for i in range(3, -1, -1):
            v = y[i]
            for j in range(i + 1, 4):
                v = v - lu[i, j] * inv[j, c]
            inv[i, c] = v / lu[i, i]
    return inv


def view_forward(c: Camera) -> np.ndarray:
    """Forward matrix of ``Camera.viewTrafo`` (Camera.fs:36-41): inverse of the camera basis."""
    return lu_inverse4(_from_basis(c.right, c.up, -np.asarray(c.forward, F64), c.location))


def transformed_view(t_forward: np.ndarray, c: Camera) -> Camera:
    """``Camera.transformedView`` (Camera.fs:61-71)."""
    to_world = _from_basis(c.right, c.up, -np.asarray(c.forward, F64), c.location)
    m = _mul44(to_world, t_forward)
    return Camera(_transform_pos(m, np.zeros(3)), -_normalize(_transform_dir(m, (0, 0, 1))),
                  _normalize(_transform_dir(m, (0, 1, 0))), _normalize(_transform_dir(m, (1, 0, 0))), c.focal)


def project_all(c: Camera, pts: np.ndarray):
    """``Camera.project1`` (Camera.fs:73-84) for every point: (projection [n,2], valid [n])."""
    o = pts - c.location
    pcx = o[:, 0] * c.right[0] + o[:, 1] * c.right[1] + o[:, 2] * c.right[2]
    pcy = o[:, 0] * c.up[0] + o[:, 1] * c.up[1] + o[:, 2] * c.up[2]
    pcz = o[:, 0] * c.forward[0] + o[:, 1] * c.forward[1] + o[:, 2] * c.forward[2]
    with np.errstate(all="ignore"):
        cx = c.focal[0] * pcx / pcz
        cy = c.focal[1] * pcy / pcz
    valid = (pcz >= 0.0) & (cx >= -1.0) & (cy >= -1.0) & (cx <= 1.0) & (cy <= 1.0)
    return np.stack([cx, cy], axis=1), valid


def avg_reprojection_error(s: float, src: Camera, world: np.ndarray, obs: np.ndarray, pose: CameraPose) -> float:
    """``avgReprojectionError`` (CameraPose.fs:75-90): mean squared distance over the observations that project."""
    dst = transformed_view(transformation(scale(s, pose)), src)
    c, valid = project_all(dst, world)
    d = c - obs
    terms = (d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1])[valid]
    if terms.size == 0:
        return math.inf
    return float(np.cumsum(terms)[-1]) / float(terms.size)          # cumsum: strictly sequential, like the F# loop


def candidate_scales(src: Camera, world: np.ndarray, obs: np.ndarray, pose: CameraPose):
    """The two candidate scales every observation proposes (CameraPose.fs:97-113): [n,2] and a valid flag [n]."""
    dst0 = transformed_view(transformation(scale(0.0, pose)), src)
    src_basis = _from_basis(src.right, src.up, -np.asarray(src.forward, F64), src.location)   # srcView.Backward
    dst0_view = view_forward(dst0)
    R = np.asarray(pose.Rotation, F64)
    t = _normalize(_transform_dir(dst0_view, _transform_dir(src_basis, _mat_vec(R, np.asarray(pose.Translation, F64)))))
    n = len(world)
    out = np.zeros((n, 2))
    ok = np.zeros(n, bool)
    for i in range(n):
        p = _transform_pos(dst0_view, world[i])
        z = obs[i] * p[2] - p[:2]
        nn = t[:2] - obs[i] * t[2]
        if abs(nn[0]) < 1e-5 or abs(nn[1]) < 1e-5:           # Fun.IsTiny
            continue
        out[i] = -z / nn
        ok[i] = True
    return out, ok


def find_scaled(inlier_threshold: float, src: Camera, world, obs, pose: CameraPose):
    """``CameraPose.findScaled`` (CameraPose.fs:39-134) -> (cost, scaled pose).  ``inlier_threshold`` is unused by
    the reference too (``countInliers`` is never called)."""
    world = np.asarray(world, F64).reshape(-1, 3)
    obs = np.asarray(obs, F64).reshape(-1, 2)
    if len(world) == 0:
        return math.inf, CameraPose()
    cand, ok = candidate_scales(src, world, obs, pose)
    best_scale, best_cost = 0.0, math.inf
    for i in range(len(world)):
        if not ok[i]:
            continue
        cx = avg_reprojection_error(cand[i, 0], src, world, obs, pose)
        cy = avg_reprojection_error(cand[i, 1], src, world, obs, pose)
        if cx < best_cost:
            best_cost, best_scale = cx, cand[i, 0]
        if cy < best_cost:
            best_cost, best_scale = cy, cand[i, 1]
    return best_cost, scale(best_scale, pose)


def avg_reprojection_errors(scales, src: Camera, world, obs, pose: CameraPose) -> np.ndarray:
    """``avgReprojectionError`` for many scales at once: the same operations in the same order as
    :func:`avg_reprojection_error`, vectorised over the scales (rows) -- sums run left to right over the
    observations (``cumsum``), so every cost is bit-identical to the scalar loop."""
    scales = np.asarray(scales, F64).ravel()
    world = np.asarray(world, F64).reshape(-1, 3)
    obs = np.asarray(obs, F64).reshape(-1, 2)
    cams = [transformed_view(transformation(scale(s, pose)), src) for s in scales]
    loc = np.array([c.location for c in cams]).reshape(-1, 3)
    axes = {k: np.array([getattr(c, k) for c in cams]).reshape(-1, 3) for k in ("right", "up", "forward")}
    o = [world[None, :, k] - loc[:, k, None] for k in range(3)]
    pc = {k: o[0] * a[:, 0, None] + o[1] * a[:, 1, None] + o[2] * a[:, 2, None] for k, a in axes.items()}
    with np.errstate(all="ignore"):
        cx = src.focal[0] * pc["right"] / pc["forward"]
        cy = src.focal[1] * pc["up"] / pc["forward"]
        valid = (pc["forward"] >= 0.0) & (cx >= -1.0) & (cy >= -1.0) & (cx <= 1.0) & (cy <= 1.0)
        dx, dy = cx - obs[None, :, 0], cy - obs[None, :, 1]
        terms = np.where(valid, dx * dx + dy * dy, 0.0)          # adding +0.0 never changes the running sum
    count = valid.sum(axis=1)
    total = np.cumsum(terms, axis=1)[:, -1] if terms.shape[1] else np.zeros(len(scales))
    with np.errstate(all="ignore"):
        return np.where(count > 0, total / np.maximum(count, 1), math.inf)


def find_scaled_fast(src: Camera, world, obs, pose: CameraPose, chunk: int = 512):
    """:func:`find_scaled` with the costs evaluated by :func:`avg_reprojection_errors` (same values, same
    first-strictly-smaller selection order) -> (cost, best scale, candidates [n,2], ok [n], costs [n,2])."""
    world = np.asarray(world, F64).reshape(-1, 3)
    obs = np.asarray(obs, F64).reshape(-1, 2)
    cand, ok = candidate_scales(src, world, obs, pose)
    costs = np.full(cand.shape, math.inf)
    flat = cand[ok].ravel()
    vals = np.concatenate([avg_reprojection_errors(flat[i:i + chunk], src, world, obs, pose)
                           for i in range(0, len(flat), chunk)]) if len(flat) else np.zeros(0)
    costs[ok] = vals.reshape(-1, 2)
    best_scale, best_cost = 0.0, math.inf
    for c, s in zip(costs[ok].ravel(), flat):
        if c < best_cost:
            best_cost, best_scale = float(c), float(s)
    return best_cost, best_scale, cand, ok, costs
