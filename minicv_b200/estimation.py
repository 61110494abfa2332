"""Two-view estimation -- host-side mirror of ``MiniCV.OpenCV`` (reference
``src/MiniCV/OpenCV.fs:855-920``): ``recoverPose``, ``recoverPoses``, ``recoverPoses2``,
``fivepoint`` with the reference's argument meaning, sentinels and return shapes, plus the
new ``findHomography`` / ``findFundamentalMat`` and the batched / injected forms.
Everything calls through the C ABI into the CUDA library; nothing is computed here."""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import numpy as np

from . import native

MODEL_H, MODEL_F, MODEL_E = 0, 1, 2
SAMPLE_MWC, SAMPLE_COUNTER, SAMPLE_INJECTED = 0, 1, 2
STOP_ADAPTIVE, STOP_FIXED, STOP_LMEDS = 0, 1, 2
MODEL_POINTS = {MODEL_H: 4, MODEL_F: 7, MODEL_E: 5}
MAX_MODELS = {MODEL_H: 1, MODEL_F: 3, MODEL_E: 10}
SCORE_AUTO, SCORE_FP32_FILTER, SCORE_TENSOR = 0, 1, 2


def set_score_engine(engine: int) -> None:
    """Scoring engine of the F / E RANSAC stage (the counts are the exact test's with either): 0 auto, 1 the fp32-filter
    kernel (one model per thread), 2 tcgen05 (residuals as a TF32 GEMM, one comparison per evaluation)."""
    if native.load().mcvSetScoreEngine(engine) != 0:
        raise ValueError("engine must be 0..2")


def score_audit(on: bool) -> float:
    """Switch the audit mode of the tcgen05 scoring engine; returns the largest |level-0 error| / budget recorded
    since it was last switched on."""
    lib = native.load()
    worst = float(lib.mcvScoreAuditMax())
    lib.mcvSetScoreAudit(1 if on else 0)
    return worst


@dataclass
class RecoverPoseConfig:
    """Reference ``RecoverPoseConfig`` (``OpenCV.fs:16-36``); ``Default`` = (1, (0,0), .999, .005)."""
    FocalLength: float = 1.0
    PrincipalPoint: tuple = (0.0, 0.0)
    Probability: float = 0.999
    InlierThreshold: float = 0.005

    def native(self) -> native.RecoverPoseConfig:
        c = native.RecoverPoseConfig()
        c.FocalLength = self.FocalLength
        c.PrincipalPoint.X, c.PrincipalPoint.Y = self.PrincipalPoint
        c.Probability = self.Probability
        c.InlierThreshold = self.InlierThreshold
        return c


def _pts(a) -> np.ndarray:
    a = np.ascontiguousarray(a, dtype=np.float64)
    return a.reshape(-1, 2)


def _p(a):
    return a.ctypes.data if a is not None and a.size else 0


def recoverPose(cfg: RecoverPoseConfig, a, b):
    """``OpenCV.recoverPose`` (``OpenCV.fs:855-861``): returns (res, m, t, ms) with the
    reference's sentinels m = identity, t = (100,123,432) left untouched on failure."""
    lib = native.load()
    a, b = _pts(a), _pts(b)
    m = np.eye(3)
    t = np.array([100.0, 123.0, 432.0])
    ms = np.zeros(a.shape[0], np.uint8)
    c = cfg.native()
    res = lib.cvRecoverPose(C.addressof(c), a.shape[0], _p(a), _p(b), _p(m), _p(t), _p(ms))
    return res, m, t, ms


def recoverPoses(cfg: RecoverPoseConfig, a, b):
    """``OpenCV.recoverPoses`` (``OpenCV.fs:863-870``): (m1, m2, t, ms)."""
    lib = native.load()
    a, b = _pts(a), _pts(b)
    m1, m2 = np.eye(3), np.eye(3)
    t = np.array([100.0, 123.0, 432.0])
    ms = np.zeros(a.shape[0], np.uint8)
    c = cfg.native()
    lib.cvRecoverPoses(C.addressof(c), a.shape[0], _p(a), _p(b), _p(m1), _p(m2), _p(t), _p(ms))
    return m1, m2, t, ms


def recoverPoses2(cfg: RecoverPoseConfig, a, b):
    """``OpenCV.recoverPoses2`` (``OpenCV.fs:872-909``): flips y, rescales the config, calls
    cvRecoverPoses, transposes R and conjugates by C = diag(1,-1,-1).  Returns
    (list of (R, t) candidate poses, boolean mask)."""
    scale = 2.0
    off = np.array([0.0, 0.0])
    a, b = _pts(a), _pts(b)

    def mp(v):
        return np.stack([0.5 * v[:, 0] + off[0], off[1] - 0.5 * v[:, 1]], axis=1) * scale

    cfg2 = RecoverPoseConfig(scale * cfg.FocalLength / 2.0, tuple(scale * off), cfg.Probability,
                             scale * 0.5 * cfg.InlierThreshold)
    m1, m2, t, ms = recoverPoses(cfg2, mp(a), mp(b))
    Cm = np.diag([1.0, -1.0, -1.0])
    m1 = Cm @ m1.T @ Cm
    m2 = Cm @ m2.T @ Cm
    t = Cm @ t
    poses = [(m1, t)] if np.array_equal(m1, m2) else [(m1, t), (m2, t)]
    return poses, ms != 0


def fivepoint(a, b) -> np.ndarray:
    """``OpenCV.fivepoint`` (``OpenCV.fs:912-920``): all essential matrices through 5 points."""
    lib = native.load()
    a, b = _pts(a), _pts(b)
    if a.shape[0] != 5 or b.shape[0] != 5:
        raise ValueError("fivepoint takes exactly 5 correspondences")
    res = np.zeros((10, 3, 3))
    cnt = lib.cvFivePoint(_p(a), _p(b), _p(res))
    return res[:cnt].copy() if cnt > 0 else np.zeros((0, 3, 3))


def fivepoint_batch(a, b):
    lib = native.load()
    a = np.ascontiguousarray(a, np.float64).reshape(-1, 5, 2)
    b = np.ascontiguousarray(b, np.float64).reshape(-1, 5, 2)
    n = a.shape[0]
    res = np.zeros((n, 10, 3, 3))
    cnt = np.zeros(n, np.int32)
    native.check(lib.cvFivePointBatch(_p(a), _p(b), n, _p(res), _p(cnt)), "cvFivePointBatch")
    return res, cnt


def findHomography(a, b, thr: float = 3.0, maxIters: int = 2000, confidence: float = 0.995):
    """== ``cv2.findHomography(a, b, cv2.RANSAC, thr, maxIters=, confidence=)`` -> (H | None, mask)."""
    lib = native.load()
    a, b = _pts(a), _pts(b)
    H = np.zeros((3, 3))
    mask = np.zeros(a.shape[0], np.uint8)
    ok = lib.cvFindHomography(_p(a), _p(b), a.shape[0], thr, maxIters, confidence, _p(H), _p(mask))
    return (H if ok else None), mask


def findFundamentalMat(a, b, thr: float = 3.0, confidence: float = 0.99, maxIters: int = 1000):
    """== ``cv2.findFundamentalMat(a, b, cv2.FM_RANSAC, thr, confidence, maxIters)`` with OpenCV's dispatch:
    7 points -> every solution of the 7-point kernel stacked ([3k, 3], mask all ones), 8..14 -> LMedS,
    15 and more -> RANSAC; returns (F | None, mask)."""
    lib = native.load()
    a, b = _pts(a), _pts(b)
    F = np.zeros((9, 3))
    mask = np.zeros(a.shape[0], np.uint8)
    k = lib.cvFindFundamentalMat(_p(a), _p(b), a.shape[0], thr, confidence, maxIters, _p(F), _p(mask))
    return (F[:3 * k].copy() if k > 0 else None), mask


@dataclass
class RansacBatchResult:
    models: np.ndarray        # [nPairs, 3, 3]
    masks: np.ndarray         # [total] uint8
    inliers: np.ndarray       # [nPairs]
    ok: np.ndarray            # [nPairs] uint8
    iters_run: np.ndarray     # [nPairs]
    best: np.ndarray          # [nPairs, 2] (iteration, model index)
    dbg_counts: np.ndarray | None = None   # [nPairs, maxIters, maxModels]
    dbg_models: np.ndarray | None = None   # [nPairs, maxIters, maxModels, 3, 3]
    dbg_samples: np.ndarray | None = None  # [nPairs, maxIters, modelPoints]


def ransac_batch(model: int, a, b, offsets, thr: float, maxIters: int, confidence: float = 0.995,
                 sample_mode: int = SAMPLE_MWC, stop_mode: int = STOP_ADAPTIVE, seed: int = 0,
                 focal: float = 1.0, pp=(0.0, 0.0), refine: bool = True, samples=None, models=None,
                 debug: bool = False, device: int = -1) -> RansacBatchResult:
    """Batched RANSAC (``cvRansacBatch``).  ``samples`` injects index sets
    [nPairs, maxIters, modelPoints]; ``models`` injects matrices [nPairs, maxIters, 3, 3]."""
    lib = native.load()
    a, b = _pts(a), _pts(b)
    off = np.ascontiguousarray(offsets, np.int64)
    n_pairs = off.size - 1
    prm = native.RansacParams(model, sample_mode, stop_mode, maxIters, thr, confidence, seed, focal,
                              pp[0], pp[1], 1 if refine else 0, 0)
    mp, mm = MODEL_POINTS[model], MAX_MODELS[model]
    inj_s = inj_m = None
    if samples is not None:
        inj_s = np.ascontiguousarray(samples, np.int32).reshape(n_pairs, maxIters, mp)
        prm.sampleMode = SAMPLE_INJECTED
    if models is not None:
        inj_m = np.ascontiguousarray(models, np.float64).reshape(n_pairs, maxIters, 9)
    out = RansacBatchResult(np.zeros((n_pairs, 3, 3)), np.zeros(a.shape[0], np.uint8),
                            np.zeros(n_pairs, np.int32), np.zeros(n_pairs, np.uint8),
                            np.zeros(n_pairs, np.int32), np.zeros((n_pairs, 2), np.int32))
    if debug:
        out.dbg_counts = np.zeros((n_pairs, maxIters, mm), np.int32)
        out.dbg_models = np.zeros((n_pairs, maxIters, mm, 3, 3))
        out.dbg_samples = np.zeros((n_pairs, maxIters, mp), np.int32)
    native.check(lib.cvRansacBatch(C.addressof(prm), _p(a), _p(b), _p(off), n_pairs, _p(inj_s), _p(inj_m),
                                   _p(out.models), _p(out.masks), _p(out.inliers), _p(out.ok),
                                   _p(out.iters_run), _p(out.best), _p(out.dbg_counts),
                                   _p(out.dbg_models), _p(out.dbg_samples), device), "cvRansacBatch")
    return out


def recover_pose_batch(cfg: RecoverPoseConfig, a, b, offsets, sample_mode: int = SAMPLE_MWC,
                       maxIters: int = 1000, device: int = -1):
    """Batched ``cvRecoverPose``: returns (R [n,3,3], t [n,3], ms [total], good [n], ok [n])."""
    lib = native.load()
    a, b = _pts(a), _pts(b)
    off = np.ascontiguousarray(offsets, np.int64)
    n = off.size - 1
    R = np.zeros((n, 3, 3)); t = np.zeros((n, 3)); ms = np.zeros(a.shape[0], np.uint8)
    good = np.zeros(n, np.int32); ok = np.zeros(n, np.uint8)
    c = cfg.native()
    native.check(lib.cvRecoverPoseBatch(C.addressof(c), _p(a), _p(b), _p(off), n, sample_mode, maxIters,
                                        _p(R), _p(t), _p(ms), _p(good), _p(ok), device), "cvRecoverPoseBatch")
    return R, t, ms, good, ok


def ransac_batch_multi(model: int, a, b, offsets, thr: float, maxIters: int, confidence: float = 0.995,
                       sample_mode: int = SAMPLE_MWC, stop_mode: int = STOP_ADAPTIVE, seed: int = 0,
                       focal: float = 1.0, pp=(0.0, 0.0), refine: bool = True,
                       device_mask: int = 0) -> RansacBatchResult:
    """``cvRansacBatchMulti``: pairs block-partitioned over the GPUs of ``device_mask`` (0 = all)."""
    lib = native.load()
    a, b = _pts(a), _pts(b)
    off = np.ascontiguousarray(offsets, np.int64)
    n_pairs = off.size - 1
    prm = native.RansacParams(model, sample_mode, stop_mode, maxIters, thr, confidence, seed, focal,
                              pp[0], pp[1], 1 if refine else 0, 0)
    out = RansacBatchResult(np.zeros((n_pairs, 3, 3)), np.zeros(a.shape[0], np.uint8),
                            np.zeros(n_pairs, np.int32), np.zeros(n_pairs, np.uint8),
                            np.zeros(n_pairs, np.int32), np.zeros((n_pairs, 2), np.int32))
    native.check(lib.cvRansacBatchMulti(C.addressof(prm), _p(a), _p(b), _p(off), n_pairs, _p(out.models),
                                        _p(out.masks), _p(out.inliers), _p(out.ok), _p(out.iters_run),
                                        device_mask), "cvRansacBatchMulti")
    return out


def recover_pose_batch_multi(cfg: RecoverPoseConfig, a, b, offsets, sample_mode: int = SAMPLE_MWC,
                             maxIters: int = 1000, device_mask: int = 0):
    """``cvRecoverPoseBatchMulti``: returns (R [n,3,3], t [n,3], ms [total], good [n], ok [n])."""
    lib = native.load()
    a, b = _pts(a), _pts(b)
    off = np.ascontiguousarray(offsets, np.int64)
    n = off.size - 1
    R = np.zeros((n, 3, 3)); t = np.zeros((n, 3)); ms = np.zeros(a.shape[0], np.uint8)
    good = np.zeros(n, np.int32); ok = np.zeros(n, np.uint8)
    c = cfg.native()
    native.check(lib.cvRecoverPoseBatchMulti(C.addressof(c), _p(a), _p(b), _p(off), n, sample_mode, maxIters,
                                             _p(R), _p(t), _p(ms), _p(good), _p(ok), device_mask),
                 "cvRecoverPoseBatchMulti")
    return R, t, ms, good, ok


@dataclass
class MatchHomographyResult:
    H: np.ndarray            # [nPairs, 3, 3]
    ok: np.ndarray           # [nPairs] uint8
    n_matches: np.ndarray    # [nPairs]
    n_inliers: np.ndarray    # [nPairs]
    match_query: np.ndarray | None = None   # packed per pair (offsets = cumsum of n_matches), local indices
    match_train: np.ndarray | None = None
    inlier_mask: np.ndarray | None = None


def match_homography_batch(query, query_offsets, train, train_offsets, kp_query, kp_train, ratio: float = 0.8,
                           thr: float = 3.0, maxIters: int = 2000, confidence: float = 0.995, want_matches: bool = False,
                           device: int = -1, device_mask: int | None = None) -> MatchHomographyResult:
    """``cvMatchHomographyBatch`` (one device) / ``cvMatchHomographyBatchMulti`` (``device_mask`` given): Hamming
    kNN-2 + ratio test -> matched keypoints -> ``findHomography`` semantics, one call, hand-off on the device."""
    lib = native.load()
    q = np.ascontiguousarray(query, np.uint8); t = np.ascontiguousarray(train, np.uint8)
    kq, kt = _pts(kp_query), _pts(kp_train)
    qo = np.ascontiguousarray(query_offsets, np.int64); to = np.ascontiguousarray(train_offsets, np.int64)
    n = qo.size - 1
    r = MatchHomographyResult(np.zeros((n, 3, 3)), np.zeros(n, np.uint8), np.zeros(n, np.int32), np.zeros(n, np.int32))
    if device_mask is not None:
        native.check(lib.cvMatchHomographyBatchMulti(_p(q), _p(qo), _p(t), _p(to), _p(kq), _p(kt), n, q.shape[1], ratio,
                                                     thr, maxIters, confidence, _p(r.H), _p(r.ok), _p(r.n_matches),
                                                     _p(r.n_inliers), device_mask), "cvMatchHomographyBatchMulti")
        return r
    if want_matches:
        r.match_query = np.zeros(q.shape[0], np.int32); r.match_train = np.zeros(q.shape[0], np.int32)
        r.inlier_mask = np.zeros(q.shape[0], np.uint8)
    native.check(lib.cvMatchHomographyBatch(_p(q), _p(qo), _p(t), _p(to), _p(kq), _p(kt), n, q.shape[1], ratio, thr,
                                            maxIters, confidence, _p(r.H), _p(r.ok), _p(r.n_matches), _p(r.n_inliers),
                                            _p(r.match_query), _p(r.match_train), _p(r.inlier_mask), device),
                 "cvMatchHomographyBatch")
    if want_matches:
        tot = int(r.n_matches.sum())
        r.match_query, r.match_train, r.inlier_mask = r.match_query[:tot], r.match_train[:tot], r.inlier_mask[:tot]
    return r


def recover_pose_from_E(E, a, b, focal=1.0, pp=(0.0, 0.0), mask=None):
    """== ``cv2.recoverPose(E, a, b, focal=, pp=, mask=)`` -> (count, R, t, mask_out, good4)."""
    lib = native.load()
    a, b = _pts(a), _pts(b)
    E = np.ascontiguousarray(E, np.float64).reshape(3, 3)
    R = np.zeros((3, 3)); t = np.zeros(3); g4 = np.zeros(4, np.int32)
    m = None if mask is None else np.ascontiguousarray(mask, np.uint8).reshape(-1).copy()
    if m is None:
        m = np.ones(a.shape[0], np.uint8)
    cnt = native.check(lib.cvRecoverPoseFromE(_p(E), _p(a), _p(b), a.shape[0], focal, pp[0], pp[1], _p(m),
                                              _p(R), _p(t), _p(g4)), "cvRecoverPoseFromE")
    return cnt, R, t, m, g4


def decompose_essential(E):
    lib = native.load()
    E = np.ascontiguousarray(E, np.float64).reshape(3, 3)
    R1 = np.zeros((3, 3)); R2 = np.zeros((3, 3)); t = np.zeros(3)
    if not lib.cvDecomposeEssentialMat(_p(E), _p(R1), _p(R2), _p(t)):
        raise native.NativeError("cvDecomposeEssentialMat failed: " + native.last_error())
    return R1, R2, t


def _dp(x):
    return x if isinstance(x, int) else (x.data_ptr() if x is not None else 0)


def ransac_batch_device(model: int, d_a, d_b, offsets, thr: float, maxIters: int, d_models, d_masks,
                        d_inliers, d_ok, confidence: float = 0.995, sample_mode: int = SAMPLE_COUNTER,
                        stop_mode: int = STOP_FIXED, seed: int = 0, focal: float = 1.0, pp=(0.0, 0.0),
                        refine: bool = True, device: int = -1, stream: int = 0) -> None:
    """Device-resident batched RANSAC (``cvRansacBatchDevice``): torch CUDA tensors (float64
    [total, 2] points) or raw device pointers; asynchronous on ``stream``."""
    lib = native.load()
    off = np.ascontiguousarray(offsets, np.int64)
    prm = native.RansacParams(model, sample_mode, stop_mode, maxIters, thr, confidence, seed, focal,
                              pp[0], pp[1], 1 if refine else 0, 0)
    native.check(lib.cvRansacBatchDevice(C.addressof(prm), _dp(d_a), _dp(d_b), _p(off), off.size - 1,
                                         _dp(d_models), _dp(d_masks), _dp(d_inliers), _dp(d_ok), device,
                                         stream), "cvRansacBatchDevice")


def recover_pose_batch_device(cfg: RecoverPoseConfig, d_a, d_b, offsets, d_E, d_R, d_t, d_ms, d_good, d_ok,
                              sample_mode: int = SAMPLE_MWC, maxIters: int = 1000, device: int = -1,
                              stream: int = 0) -> None:
    """Device-resident batched ``cvRecoverPose`` (``cvRecoverPoseBatchDevice``)."""
    lib = native.load()
    off = np.ascontiguousarray(offsets, np.int64)
    c = cfg.native()
    native.check(lib.cvRecoverPoseBatchDevice(C.addressof(c), _dp(d_a), _dp(d_b), _p(off), off.size - 1,
                                              sample_mode, maxIters, _dp(d_E), _dp(d_R), _dp(d_t), _dp(d_ms),
                                              _dp(d_good), _dp(d_ok), device, stream),
                 "cvRecoverPoseBatchDevice")


# ------------------------------------------------------------------------------------------
# PnP RANSAC -- mirror of ``OpenCV.solvePnPRansac`` / ``OpenCV.solveAp3p`` (reference
# ``src/MiniCV/OpenCV.fs:383-432, 956-1052``); SURVEY.md section 8f rank 3
# ------------------------------------------------------------------------------------------
class SolverKind:
    """Reference ``SolverKind`` (``OpenCV.fs:956-961``) with the native codes of ``solvePnPInternal``."""
    Iterative, EPNP, P3P, AP3P, SQPNP = 0, 1, 2, 5, 6


@dataclass
class PnPRansacParams:
    """Reference ``RansacParams`` (``OpenCV.fs:962-974``)."""
    reprojectionError: float = 8.0
    iterationsCount: int = 100
    confidence: float = 0.99

    @staticmethod
    def MaxReprojError(e: float) -> "PnPRansacParams":
        return PnPRansacParams(float(np.float32(e)), 100, 0.99)


def solvePnPRansac(solver: int, cfg: PnPRansacParams, imgPoints, worldPoints, intern, distortionCoeffs):
    """``OpenCV.solvePnPRansac`` up to the native call (``OpenCV.fs:976-1000``): returns
    ``(rvec, tvec, inliers)`` or ``None`` when the lengths differ or the estimator fails.
    The Rot3d / Euclidean3d conversion that follows in the reference is Aardvark-side."""
    lib = native.load()
    img = _pts(imgPoints)
    world = np.ascontiguousarray(worldPoints, dtype=np.float64).reshape(-1, 3)
    if img.shape[0] != world.shape[0]:
        return None
    K = native.Matx33d()
    Kf = np.asarray(intern, np.float64).reshape(9)
    for i in range(9):
        K.M[i] = Kf[i]
    dist = np.zeros(4)
    d = np.asarray(distortionCoeffs, np.float64).ravel()
    dist[:min(4, d.size)] = d[:4]
    t = np.zeros(3)
    r = np.zeros(3)
    icnt = C.c_int(0)
    inliers = np.zeros(max(img.shape[0], 1), np.int32)
    ok = lib.cvSolvePnPRansac(_p(img), _p(world), img.shape[0], K, _p(dist), int(solver), int(cfg.iterationsCount),
                              float(cfg.reprojectionError), float(cfg.confidence), _p(t), _p(r), C.addressof(icnt),
                              _p(inliers))
    if not ok:
        return None
    return r, t, inliers[:icnt.value].copy()


def solveAp3p(img, world, inv_fx, inv_fy, cx_fx, cy_fy):
    """``Native.solveAp3p`` (``OpenCV.fs:373-374``, reference ``ap3p.cpp:282``): up to four (R, t) through
    three correspondences, R in the reference's layout (camera point = R^T X + t)."""
    lib = native.load()
    img = _pts(img)
    world = np.ascontiguousarray(world, dtype=np.float64).reshape(-1, 3)
    Rs = np.zeros((4, 9))
    ts = np.zeros((4, 3))
    args = []
    for i in range(3):
        args += [img[i, 0], img[i, 1], world[i, 0], world[i, 1], world[i, 2]]
    n = lib.solveAp3p(_p(Rs), _p(ts), *[float(v) for v in args], float(inv_fx), float(inv_fy), float(cx_fx), float(cy_fy))
    return Rs[:n].reshape(n, 3, 3).copy(), ts[:n].copy()


def solve_ap3p_batch(img, world, cam4):
    """Batched ``solveAp3p``: img [n,3,2], world [n,3,3], cam4 [n,4] -> (Rs [n,4,3,3], ts [n,4,3], counts)."""
    lib = native.load()
    img = np.ascontiguousarray(img, np.float64).reshape(-1, 3, 2)
    world = np.ascontiguousarray(world, np.float64).reshape(-1, 3, 3)
    cam4 = np.ascontiguousarray(cam4, np.float64).reshape(-1, 4)
    n = img.shape[0]
    Rs = np.zeros((n, 4, 9))
    ts = np.zeros((n, 4, 3))
    counts = np.zeros(n, np.int32)
    native.check(lib.cvSolveAp3pBatch(_p(img), _p(world), _p(cam4), n, _p(Rs), _p(ts), _p(counts)), "cvSolveAp3pBatch")
    return Rs.reshape(n, 4, 3, 3), ts, counts


@dataclass
class PnpBatchResult:
    rvecs: np.ndarray
    tvecs: np.ndarray
    masks: np.ndarray
    inlier_counts: np.ndarray
    ok: np.ndarray
    iters_run: np.ndarray
    best_iter: np.ndarray
    minimal: np.ndarray
    dbg_counts: np.ndarray | None = None
    dbg_models: np.ndarray | None = None
    dbg_samples: np.ndarray | None = None
    n_ok: int = 0


def camera_rows(K, dist=None, n: int = 1) -> np.ndarray:
    """[n, 8] camera rows (fx, fy, cx, cy, k1, k2, p1, p2) from a 3x3 K and up to four coefficients."""
    K = np.asarray(K, np.float64).reshape(3, 3)
    row = np.zeros(8)
    row[:4] = (K[0, 0], K[1, 1], K[0, 2], K[1, 2])
    if dist is not None:
        d = np.asarray(dist, np.float64).ravel()
        row[4:4 + min(4, d.size)] = d[:4]
    return np.tile(row, (n, 1))


def solve_pnp_ransac_batch(world, img, offsets, cams, thr: float = 8.0, maxIters: int = 100, confidence: float = 0.99,
                           sample_mode: int = SAMPLE_MWC, stop_mode: int = STOP_ADAPTIVE, seed: int = 0,
                           refine: bool = True, inj_samples=None, inj_models=None, debug: bool = False,
                           device: int = -1) -> PnpBatchResult:
    """``cvSolvePnPRansacBatch``: problem p owns rows [offsets[p], offsets[p+1])."""
    lib = native.load()
    world = np.ascontiguousarray(world, np.float64).reshape(-1, 3)
    img = _pts(img)
    off = np.ascontiguousarray(offsets, np.int64)
    n = off.size - 1
    cams = np.ascontiguousarray(cams, np.float64).reshape(-1, 8)
    assert cams.shape[0] == n
    prm = native.PnpParams(sample_mode if inj_samples is None else SAMPLE_INJECTED, stop_mode, maxIters, 1 if refine else 0,
                           thr, confidence, seed)
    total = int(off[-1])
    rv = np.zeros((n, 3)); tv = np.zeros((n, 3))
    masks = np.zeros(total, np.uint8)
    cnt = np.zeros(n, np.int32); ok = np.zeros(n, np.uint8)
    iters = np.zeros(n, np.int32); best = np.zeros(n, np.int32); minimal = np.zeros((n, 6))
    js = np.ascontiguousarray(inj_samples, np.int32) if inj_samples is not None else None
    jm = np.ascontiguousarray(inj_models, np.float64) if inj_models is not None else None
    mi = max(maxIters, 1)
    dc = np.zeros((n, mi), np.int32) if debug else None
    dm = np.zeros((n, mi, 6)) if debug else None
    ds = np.zeros((n, mi, 4), np.int32) if debug else None
    rc = lib.cvSolvePnPRansacBatch(C.addressof(prm), _p(world), _p(img), _p(off), n, _p(cams), _p(js), _p(jm), _p(rv), _p(tv),
                                   _p(masks), _p(cnt), _p(ok), _p(iters), _p(best), _p(minimal), _p(dc), _p(dm), _p(ds),
                                   device)
    native.check(rc, "cvSolvePnPRansacBatch")
    return PnpBatchResult(rv, tv, masks, cnt, ok, iters, best, minimal, dc, dm, ds, rc)


def solve_pnp_ransac_batch_device(d_world, d_img, offsets, d_cams, thr, maxIters, confidence, d_rvecs, d_tvecs, d_masks,
                                  d_counts, d_ok, sample_mode: int = SAMPLE_MWC, stop_mode: int = STOP_ADAPTIVE,
                                  seed: int = 0, refine: bool = True, device: int = -1, stream: int = 0):
    """``cvSolvePnPRansacBatchDevice``: torch tensors (or raw device pointers), asynchronous."""
    lib = native.load()
    off = np.ascontiguousarray(offsets, np.int64)
    prm = native.PnpParams(sample_mode, stop_mode, maxIters, 1 if refine else 0, thr, confidence, seed)
    rc = lib.cvSolvePnPRansacBatchDevice(C.addressof(prm), _dp(d_world), _dp(d_img), _p(off), off.size - 1, _dp(d_cams),
                                         _dp(d_rvecs), _dp(d_tvecs), _dp(d_masks), _dp(d_counts), _dp(d_ok), device, stream)
    native.check(rc, "cvSolvePnPRansacBatchDevice")
