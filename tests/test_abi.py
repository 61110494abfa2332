"""The C-ABI library loads and exports every symbol include/MiniCVNative.h declares
(no compute calls: runs without a GPU)."""
import ctypes
import re
from pathlib import Path

from minicv_b200 import native

ROOT = Path(__file__).resolve().parent.parent


def declared_symbols():
    text = (ROOT / "include" / "MiniCVNative.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return re.findall(r"MCV_API\s+[\w\s\*]+?\b(\w+)\s*\(", text)


def test_header_declares_symbols():
    names = declared_symbols()
    assert "cvMatchKnn2" in names and "mcvLastError" in names
    assert len(names) == len(set(names))


def test_library_exports_every_declared_symbol():
    lib = ctypes.CDLL(str(native.LIB_PATH))
    missing = [n for n in declared_symbols() if not hasattr(lib, n)]
    assert not missing, f"declared in the header but not exported: {missing}"


def test_python_signatures_cover_header():
    names = set(declared_symbols())
    assert names == set(native.SIGNATURES), names ^ set(native.SIGNATURES)


def test_struct_sizes_match_reference_abi():
    # reference: MiniCVNative.cpp:39-46 (RecoverPoseConfig 40 B), Point2d 16 B, Matx33d 72 B, Vec3d 24 B
    assert ctypes.sizeof(native.RecoverPoseConfig) == 40
    assert ctypes.sizeof(native.Point2d) == 16
    assert ctypes.sizeof(native.Matx33d) == 72
    assert ctypes.sizeof(native.Vec3d) == 24
    assert ctypes.sizeof(native.KeyPoint2d) == 28


def test_no_device_means_loud_failure_not_fallback():
    lib = native.load()
    if lib.mcvDeviceCount() > 0:
        return
    import numpy as np
    q = np.zeros((4, 32), np.uint8)
    idx = np.zeros((4, 2), np.int32)
    dist = np.zeros((4, 2), np.float32)
    rc = lib.cvMatchKnn2(q.ctypes.data, 4, q.ctypes.data, 4, 32, 6, 0.8, idx.ctypes.data,
                         dist.ctypes.data, 0)
    assert rc == -1
    assert b"no usable CUDA device" in lib.mcvLastError()


def test_out_of_scope_exports_resolve_and_fail_cleanly():
    """The ten non-path reference exports exist (DllImports keep resolving) and, with no original
    shim loaded, return the reference's failure values without computing anything."""
    lib = native.load()
    import numpy as np
    t = np.zeros(3); r = np.zeros(3)
    K = native.Matx33d()
    assert lib.cvSolvePnP(0, 0, 0, K, 0, 0, t.ctypes.data, r.ctypes.data) is False
    assert b"not part of the GPU path" in lib.mcvLastError()
    assert lib.cvDetectFeatures(0, 0, 0, 0, 1, 0) is None
    assert lib.solveAp3p(0, 0, *([0.0] * 19)) == 0
    lib.cvTest()
    lib.cvFreeFeatures(0)
