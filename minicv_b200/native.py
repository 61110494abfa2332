"""ctypes binding of libMiniCVNative.so -- the stand-in for MiniCV's F# P/Invoke layer.

Mirrors ``module Native`` of the reference (``src/MiniCV/OpenCV.fs:339-382``): the same
library name, the same entry-point names, blittable structs passed by pointer.  There is no
``dotnet`` in this image, so this module plays the role of the managed caller.

The library is built in-tree by ``make`` (see ``__graft_entry__.build``).  Loading fails
loudly when it is missing; nothing here falls back to a CPU implementation.
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

_HERE = Path(__file__).resolve().parent
LIB_PATH = _HERE / "lib" / "libMiniCVNative.so"


class NativeLibraryMissing(RuntimeError):
    pass


class NativeError(RuntimeError):
    pass


class Point2d(C.Structure):
    _fields_ = [("X", C.c_double), ("Y", C.c_double)]


class Matx33d(C.Structure):
    _fields_ = [("M", C.c_double * 9)]


class Vec3d(C.Structure):
    _fields_ = [("V", C.c_double * 3)]


class RecoverPoseConfig(C.Structure):
    """Reference: ``src/MiniCVNative/MiniCVNative.cpp:39-46`` / ``OpenCV.fs:16-36`` (40 bytes)."""
    _fields_ = [("FocalLength", C.c_double), ("PrincipalPoint", Point2d),
                ("Probability", C.c_double), ("InlierThreshold", C.c_double)]


class KeyPoint2d(C.Structure):
    _fields_ = [("ptX", C.c_float), ("ptY", C.c_float), ("size", C.c_float), ("angle", C.c_float),
                ("response", C.c_float), ("octave", C.c_int), ("class_id", C.c_int)]


class DetectorResult(C.Structure):
    """Reference: ``src/MiniCVNative/MiniCVNative.h:23-29``."""
    _fields_ = [("PointCount", C.c_int), ("DescriptorEntries", C.c_int),
                ("DescriptorElementType", C.c_int), ("Points", C.POINTER(KeyPoint2d)),
                ("Descriptors", C.POINTER(C.c_uint8))]


assert C.sizeof(RecoverPoseConfig) == 40 and C.sizeof(Point2d) == 16
assert C.sizeof(Matx33d) == 72 and C.sizeof(Vec3d) == 24

_vp, _i, _f, _d, _i64 = C.c_void_p, C.c_int, C.c_float, C.c_double, C.c_int64

# name -> (restype, argtypes).  Every symbol include/MiniCVNative.h declares is listed here;
# tests/test_abi.py checks the two stay in step.
SIGNATURES = {
    "mcvLastError": (C.c_char_p, []),
    "mcvDeviceCount": (_i, []),
    "mcvKernelLaunchCount": (_i64, []),
    "mcvBuildInfo": (C.c_char_p, []),
    "mcvSetHammingEngine": (_i, [_i]),
    "mcvSetL2Engine": (_i, [_i]),
    "mcvSetScoreEngine": (_i, [_i]),
    "mcvSetScoreAudit": (_i, [_i]),
    "mcvScoreAuditMax": (_f, []),
    "mcvTrim": (_i, []),
    "mcvWorkspaceBytes": (_i64, [_i]),
    "mcvSetWorkspaceLimit": (_i, [_i64]),
    "mcvContextCount": (_i, [_i]),
    "mcvAllocPinned": (_vp, [_i64]),
    "mcvFreePinned": (None, [_vp]),
    "mcvHostRegister": (_i, [_vp, _i64]),
    "mcvHostUnregister": (_i, [_vp]),
    "mcvSetStagingMode": (_i, [_i]),
    "mcvCopyThreads": (_i, []),
    "mcvMeasureCopyBandwidth": (_d, [_i, _i64, _i, _i, _i]),
    "mcvCopyProbeCreate": (_vp, [_i, _i64]),
    "mcvCopyProbeRun": (_i, [_vp, _d, _i, _vp]),
    "mcvCopyProbeDestroy": (None, [_vp]),
    "mcvNcclVersion": (_i, []),
    "mcvShardCommUniqueId": (_i, [_vp]),
    "mcvShardCommInit": (_vp, [_vp, _i, _i, _i]),
    "mcvShardCommDestroy": (None, [_vp]),
    "cvMatchKnn2ShardedDevice": (_i, [_vp, _vp, _i, _vp, _i64, _i64, _i, _f, _vp, _vp, _vp, _vp]),
    "cvLargeDbCreate": (_vp, [_vp, _i64, _i, C.c_uint32]),
    "cvLargeDbMatch": (_i, [_vp, _vp, _i, _f, _vp, _vp, _vp]),
    "cvLargeDbInfo": (_i, [_vp, _vp, _vp, _vp]),
    "cvLargeDbDestroy": (None, [_vp]),
    "cvMatchHomographyBatch": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _f, _d, _i, _d, _vp, _vp, _vp, _vp, _vp, _vp,
                                    _vp, _i]),
    "cvMatchHomographyBatchMulti": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _f, _d, _i, _d, _vp, _vp, _vp, _vp,
                                         C.c_uint32]),
    "cvMatchKnn2BatchMulti": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _f, _vp, _vp, _vp, C.c_uint32]),
    "cvRansacBatchMulti": (_i, [_vp, _vp, _vp, _vp, _i, _vp, _vp, _vp, _vp, _vp, C.c_uint32]),
    "cvRecoverPoseBatchMulti": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, C.c_uint32]),
    "cvSolvePnPRansacBatchMulti": (_i, [_vp, _vp, _vp, _vp, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, C.c_uint32]),
    "mcvKernelTimingStart": (_i, [_i, _i]),
    "mcvKernelTimingStop": (_i, [_i, _vp, _i]),
    "mcvKernelTimingStopTagged": (_i, [_i, _vp, _vp, _i]),
    "cvMatchKnn2": (_i, [_vp, _i, _vp, _i, _i, _i, _f, _vp, _vp, _vp]),
    "cvMatchKnn2Batch": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _f, _vp, _vp, _vp, _i]),
    "cvMatchKnn2BatchDevice": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _f, _vp, _vp, _vp, _i, _vp]),
    "cvMatchFeatures": (_i, [C.POINTER(DetectorResult), C.POINTER(DetectorResult), _f, _vp, _vp, _vp]),
    "cvGatherMatchesDevice": (_i, [_vp, _vp, _vp, _vp, _i, _vp, _vp, _vp, _vp, _vp, _i, _vp]),
    "cvMatchKnn2ShardKeysDevice": (_i, [_vp, _i, _vp, _i, _i64, _i, _vp, _i, _vp]),
    "cvMergeShardKeysDevice": (_i, [_vp, _i, _i, _f, _vp, _vp, _vp, _i, _vp]),
    "cvRecoverPose": (_i, [_vp, _i, _vp, _vp, _vp, _vp, _vp]),
    "cvRecoverPoses": (C.c_int32, [_vp, _i, _vp, _vp, _vp, _vp, _vp, _vp]),
    "cvFivePoint": (_i, [_vp, _vp, _vp]),
    "cvFindHomography": (_i, [_vp, _vp, _i, _d, _i, _d, _vp, _vp]),
    "cvFindFundamentalMat": (_i, [_vp, _vp, _i, _d, _d, _i, _vp, _vp]),
    "cvRansacBatch": (_i, [_vp, _vp, _vp, _vp, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i]),
    "cvRansacBatchDevice": (_i, [_vp, _vp, _vp, _vp, _i, _vp, _vp, _vp, _vp, _i, _vp]),
    "cvRecoverPoseBatch": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _i]),
    "cvRecoverPoseBatchDevice": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _i, _vp]),
    "cvRecoverPoseFromE": (_i, [_vp, _vp, _vp, _i, _d, _d, _d, _vp, _vp, _vp, _vp]),
    "cvDecomposeEssentialMat": (_i, [_vp, _vp, _vp, _vp]),
    "cvFivePointBatch": (_i, [_vp, _vp, _i, _vp, _vp]),
    # group (3): forwarded to the original shim when present
    "cvSolvePnP": (C.c_bool, [_vp, _vp, _i, Matx33d, _vp, _i, _vp, _vp]),
    "cvRefinePnPLM": (None, [_vp, _vp, _i, Matx33d, _vp, _vp, _vp]),
    "cvRefinePnPVVS": (None, [_vp, _vp, _i, Matx33d, _vp, _vp, _vp]),
    "cvDetectFeatures": (_vp, [_vp, _i, _i, _i, _i, _vp]),
    "cvFreeFeatures": (None, [_vp]),
    "cvDetectQRCode": (C.c_bool, [_vp, _i, _i, _i, _vp, _vp]),
    "cvDetectArucoMarkers": (C.c_bool, [_vp, _i, _i, _i, _vp, _vp]),
    "cvTest": (None, []),
}
# group (2b): PnP RANSAC (AP3P solver kind native; other kinds of cvSolvePnPRansac are forwarded)
SIGNATURES.update({
    "cvSolvePnPRansac": (C.c_bool, [_vp, _vp, _i, Matx33d, _vp, _i, _i, _f, _d, _vp, _vp, _vp, _vp]),
    "solveAp3p": (_i, [_vp, _vp] + [_d] * 19),
    "cvSolveAp3pBatch": (_i, [_vp, _vp, _vp, _i, _vp, _vp, _vp]),
    "cvSolvePnPRansacBatch": (_i, [_vp, _vp, _vp, _vp, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp,
                                   _vp, _vp, _vp, _i]),
    "cvSolvePnPRansacBatchDevice": (_i, [_vp, _vp, _vp, _vp, _i, _vp, _vp, _vp, _vp, _vp, _vp, _i, _vp]),
})
# group (2c): scale search (CameraPose.findScaled moved behind the native boundary)
SIGNATURES.update({
    "cvFindScaled": (_d, [_d, _vp, _vp, _vp, _i, _vp, _vp]),
    "cvFindScaledBatch": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _vp, _vp, _vp, _vp, _vp, _vp, _i]),
    "cvFindScaledBatchDevice": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _i, _vp, _vp, _vp, _vp, _i, _vp]),
    "cvScaleCosts": (_i, [_vp, _vp, _vp, _vp, _i, _vp, _i, _vp]),
})


class Camera(C.Structure):
    """``McvCamera``: blittable mirror of the F# record ``Camera`` (``src/MiniCV/Camera.fs:7-14``)."""
    _fields_ = [("location", C.c_double * 3), ("forward", C.c_double * 3), ("up", C.c_double * 3),
                ("right", C.c_double * 3), ("focal", C.c_double * 2)]


class CameraPose(C.Structure):
    """``McvCameraPose``: the F# struct ``CameraPose`` (``src/MiniCV/CameraPose.fs:7-16``)."""
    _fields_ = [("RotationIndex", C.c_int32), ("ScaleSign", C.c_int32), ("Rotation", C.c_double * 9),
                ("Translation", C.c_double * 3), ("IsInverse", C.c_int32), ("_pad", C.c_int32)]


assert C.sizeof(Camera) == 112 and C.sizeof(CameraPose) == 112


class RansacParams(C.Structure):
    """``McvRansacParams`` of include/MiniCVNative.h."""
    _fields_ = [("model", C.c_int32), ("sampleMode", C.c_int32), ("stopMode", C.c_int32),
                ("maxIters", C.c_int32), ("threshold", C.c_double), ("confidence", C.c_double),
                ("seed", C.c_uint64), ("focal", C.c_double), ("ppx", C.c_double), ("ppy", C.c_double),
                ("refine", C.c_int32), ("pairIndexBase", C.c_int32)]



class PnpParams(C.Structure):
    """``McvPnpParams`` of include/MiniCVNative.h."""
    _fields_ = [("sampleMode", C.c_int32), ("stopMode", C.c_int32), ("maxIters", C.c_int32), ("refine", C.c_int32),
                ("threshold", C.c_double), ("confidence", C.c_double), ("seed", C.c_uint64),
                ("problemIndexBase", C.c_int64)]


_lib = None


def _prefer_bundled_nccl() -> None:
    """The library resolves NCCL at run time (dlopen).  In a Python process that may import torch later, it must pick the
    SAME libnccl torch was built against: the loader de-duplicates by soname, so an older system libnccl.so.2 loaded
    first would later be handed to libtorch_cuda.so (undefined symbol at `import torch`).  Point MINICV_NCCL_LIB at the
    wheel-bundled copy when there is one and the caller has not chosen."""
    if os.environ.get("MINICV_NCCL_LIB"):
        return
    try:
        import importlib.util
        spec = importlib.util.find_spec("nvidia.nccl")
        for root in (spec.submodule_search_locations or []) if spec else []:
            cand = Path(root) / "lib" / "libnccl.so.2"
            if cand.exists():
                os.environ["MINICV_NCCL_LIB"] = str(cand)
                return
    except Exception:
        pass


def load(path: os.PathLike | str | None = None) -> C.CDLL:
    """Load libMiniCVNative.so and attach signatures.  Raises if the library is absent."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    p = Path(path) if path else Path(os.environ.get("MINICV_NATIVE_LIB", LIB_PATH))
    if not p.exists():
        raise NativeLibraryMissing(
            f"{p} not found: build it with `make` (or __graft_entry__.build()). "
            "minicv_b200 has no CPU fallback for the match/estimate path.")
    _prefer_bundled_nccl()
    lib = C.CDLL(str(p))
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError here means header/library drift
        fn.restype = res
        fn.argtypes = args
    if path is None:
        _lib = lib
    return lib


def last_error() -> str:
    return load().mcvLastError().decode("utf-8", "replace")


def check(rc: int, what: str) -> int:
    if rc < 0:
        raise NativeError(f"{what} failed: {last_error()}")
    return rc
