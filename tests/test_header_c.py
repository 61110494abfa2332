"""The public header is valid C (the reference ABI is plain C) and the struct sizes the F# side
relies on hold when compiled by gcc."""
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent

SRC = r'''
#include "MiniCVNative.h"
#include <stdio.h>
int main(void) {
    _Static_assert(sizeof(RecoverPoseConfig) == 40, "RecoverPoseConfig");
    _Static_assert(sizeof(McvPoint2d) == 16, "Point2d");
    _Static_assert(sizeof(McvMatx33d) == 72, "Matx33d");
    _Static_assert(sizeof(McvVec3d) == 24, "Vec3d");
    _Static_assert(sizeof(McvKeyPoint2d) == 28, "KeyPoint2d");
    _Static_assert(sizeof(McvRansacParams) == 72, "McvRansacParams");
    printf("%d %d\n", MCV_NORM_HAMMING, MCV_SAMPLE_INJECTED);
    return 0;
}
'''


def test_header_compiles_as_c(tmp_path):
    c = tmp_path / "t.c"
    c.write_text(SRC)
    exe = tmp_path / "t"
    subprocess.run(["gcc", "-std=c11", "-Wall", "-Werror", "-I", str(ROOT / "include"), str(c), "-o", str(exe)], check=True)
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.split()
    assert out == ["6", "2"]


def test_python_struct_matches_c_struct():
    import ctypes
    from minicv_b200 import native
    assert ctypes.sizeof(native.RansacParams) == 72
