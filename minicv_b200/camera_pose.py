"""Scale search -- host-side mirror of ``MiniCV.CameraPose`` / ``MiniCV.Camera`` (reference
``src/MiniCV/CameraPose.fs:7-134``, ``src/MiniCV/Camera.fs:7-14, 94-105``): the ``Camera`` record, the
``CameraPose`` struct and ``findScaled`` with the reference's argument order and return shape (cost, pose).
The O(n^2) search itself runs in the CUDA library (``cvFindScaled*``); nothing is computed here beyond
``Camera.lookAt`` (three cross products that build an input) and struct marshalling."""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field

import numpy as np

from . import native


def _v(x, n):
    a = np.asarray(x, dtype=np.float64).reshape(-1)
    assert a.size == n
    return a


@dataclass
class Camera:
    """``Camera`` record (``Camera.fs:7-14``)."""
    location: np.ndarray
    forward: np.ndarray
    up: np.ndarray
    right: np.ndarray
    focal: np.ndarray

    @staticmethod
    def lookAt(eye, center, sky, f) -> "Camera":
        """``Camera.lookAt`` (``Camera.fs:94-105``)."""
        eye, center, sky = _v(eye, 3), _v(center, 3), _v(sky, 3)
        unit = lambda v: v / np.sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2])
        fw = unit(center - eye)
        r = unit(np.cross(fw, sky))
        u = unit(np.cross(r, fw))
        return Camera(eye, fw, u, r, _v(f, 2))

    def row(self) -> np.ndarray:
        """The 14 doubles of ``McvCamera``."""
        return np.concatenate([_v(self.location, 3), _v(self.forward, 3), _v(self.up, 3), _v(self.right, 3),
                               _v(self.focal, 2)])


@dataclass
class CameraPose:
    """``CameraPose`` struct (``CameraPose.fs:7-16``); the default is the all-zero ``CameraPose()``."""
    RotationIndex: int = 0
    ScaleSign: int = 0
    Rotation: np.ndarray = field(default_factory=lambda: np.zeros((3, 3)))
    Translation: np.ndarray = field(default_factory=lambda: np.zeros(3))
    IsInverse: bool = False

    def native(self) -> native.CameraPose:
        p = native.CameraPose()
        p.RotationIndex, p.ScaleSign, p.IsInverse = int(self.RotationIndex), int(self.ScaleSign), int(bool(self.IsInverse))
        p.Rotation[:] = _v(self.Rotation, 9)
        p.Translation[:] = _v(self.Translation, 3)
        return p

    @staticmethod
    def from_native(p: native.CameraPose) -> "CameraPose":
        return CameraPose(p.RotationIndex, p.ScaleSign, np.array(p.Rotation[:]).reshape(3, 3),
                          np.array(p.Translation[:]), bool(p.IsInverse))


def name(pose: CameraPose) -> str:
    """``CameraPose.name`` (``CameraPose.fs:20-28``)."""
    trans = {1: "t", -1: "-t"}.get(pose.ScaleSign, "0")
    return "r%d|%s" % (pose.RotationIndex, trans)


def _native_cam(cam: Camera) -> native.Camera:
    c = native.Camera()
    c.location[:], c.forward[:], c.up[:], c.right[:] = _v(cam.location, 3), _v(cam.forward, 3), _v(cam.up, 3), _v(cam.right, 3)
    c.focal[:] = _v(cam.focal, 2)
    return c


def _split(worldObservations):
    """``list<V3d * V2d>`` -> packed world [n,3] and observation [n,2] arrays."""
    if isinstance(worldObservations, tuple) and len(worldObservations) == 2 and (
            np.ndim(worldObservations[0]) == 2 or np.size(worldObservations[0]) == 0):
        w, o = worldObservations
    else:
        w = [wo[0] for wo in worldObservations]
        o = [wo[1] for wo in worldObservations]
    w = np.ascontiguousarray(w, dtype=np.float64).reshape(-1, 3)
    o = np.ascontiguousarray(o, dtype=np.float64).reshape(-1, 2)
    assert len(w) == len(o)
    return w, o


def _p(a):
    return a.ctypes.data if a is not None and a.size else 0


def findScaled(inlierThreshold: float, srcCam: Camera, worldObservations, pose: CameraPose):
    """``CameraPose.findScaled`` (``CameraPose.fs:39-134``) -> (cost, scaled pose).  ``worldObservations`` is a
    list of (V3d, V2d) pairs like the reference's, or a (world [n,3], obs [n,2]) tuple of arrays."""
    lib = native.load()
    w, o = _split(worldObservations)
    cam, p, out = _native_cam(srcCam), pose.native(), native.CameraPose()
    cost = lib.cvFindScaled(float(inlierThreshold), C.addressof(cam), _p(w), _p(o), len(w), C.addressof(p),
                            C.addressof(out))
    if cost != cost and native.last_error():
        raise native.NativeError("cvFindScaled failed: " + native.last_error())
    return cost, CameraPose.from_native(out)


@dataclass
class FindScaledBatchResult:
    costs: np.ndarray            # [nProblems]
    scales: np.ndarray           # [nProblems] best scale
    poses: list                  # CameraPose per problem
    candidates: np.ndarray = None    # [total, 2]   (taps=True)
    cand_ok: np.ndarray = None       # [total] bool
    cand_costs: np.ndarray = None    # [total, 2]


def find_scaled_batch(srcCams, poses, world, obs, offsets, taps: bool = False, device: int = -1) -> FindScaledBatchResult:
    """Batched ``findScaled``: problem p owns rows [offsets[p], offsets[p+1]) of ``world`` / ``obs``."""
    lib = native.load()
    off = np.ascontiguousarray(offsets, dtype=np.int64)
    n_prob = len(off) - 1
    w = np.ascontiguousarray(world, dtype=np.float64).reshape(-1, 3)
    o = np.ascontiguousarray(obs, dtype=np.float64).reshape(-1, 2)
    cams = (native.Camera * max(n_prob, 1))(*[_native_cam(c) for c in srcCams])
    ps = (native.CameraPose * max(n_prob, 1))(*[p.native() for p in poses])
    out = (native.CameraPose * max(n_prob, 1))()
    costs, scales = np.zeros(n_prob), np.zeros(n_prob)
    total = int(off[-1]) if n_prob else 0
    cand = np.zeros((total, 2)) if taps else None
    cok = np.zeros((total, 2), np.uint8) if taps else None
    ccost = np.zeros((total, 2)) if taps else None
    native.check(lib.cvFindScaledBatch(C.addressof(cams), C.addressof(ps), _p(w), _p(o), _p(off), n_prob, _p(costs),
                                       C.addressof(out), _p(scales), _p(cand), _p(cok), _p(ccost), device),
                 "cvFindScaledBatch")
    return FindScaledBatchResult(costs, scales, [CameraPose.from_native(out[i]) for i in range(n_prob)], cand,
                                 cok[:, 0].astype(bool) if taps else None, ccost)


def avgReprojectionError(scales, srcCam: Camera, worldObservations, pose: CameraPose) -> np.ndarray:
    """The closure ``avgReprojectionError`` of ``findScaled`` (``CameraPose.fs:75-90``) for given scales."""
    lib = native.load()
    w, o = _split(worldObservations)
    s = np.ascontiguousarray(scales, dtype=np.float64).reshape(-1)
    cam, p = _native_cam(srcCam), pose.native()
    costs = np.zeros(len(s))
    native.check(lib.cvScaleCosts(C.addressof(cam), C.addressof(p), _p(w), _p(o), len(w), _p(s), len(s), _p(costs)),
                 "cvScaleCosts")
    return costs


def _dp(x):
    return 0 if x is None else int(x.data_ptr())


def find_scaled_batch_device(d_cams, d_rot, d_trans, d_world, d_obs, offsets, d_costs, d_scales, d_scaled_t, d_signs,
                             device: int = -1, stream=None) -> None:
    """Device-resident batched ``findScaled`` (torch tensors only carry the device memory)."""
    lib = native.load()
    off = np.ascontiguousarray(offsets, dtype=np.int64)
    native.check(lib.cvFindScaledBatchDevice(_dp(d_cams), _dp(d_rot), _dp(d_trans), _dp(d_world), _dp(d_obs), _p(off),
                                             len(off) - 1, _dp(d_costs), _dp(d_scales), _dp(d_scaled_t), _dp(d_signs),
                                             device, 0 if stream is None else int(stream)),
                 "cvFindScaledBatchDevice")
