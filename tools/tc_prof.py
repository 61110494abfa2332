"""Per-CTA cycle counter of hamming_tc_kernel against the train-set size: the slope is the steady-state
cost of one 128-row train stage, the intercept the fixed cost of a CTA (launch, tile lookup, query
staging, pipeline fill and drain).  Needs a library built with -DMCV_TC_PROF=1 (2 adds wait counters):
   make LIB=build/libprof.so OBJDIR=build/objprof EXTRA=-DMCV_TC_PROF=1
   MINICV_NATIVE_LIB=build/libprof.so python tools/tc_prof.py"""
import ctypes as C
import sys
from pathlib import Path
import numpy as np
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import minicv_b200 as mcv  # noqa: E402
from minicv_b200 import native  # noqa: E402

pairs = int(sys.argv[1]) if len(sys.argv) > 1 else 74      # 74 pairs x 32 tiles = 16 CTAs per SM
rng = np.random.default_rng(1)
mcv.set_hamming_engine(mcv.ENGINE_TENSOR)
lib = native.load()
out = (C.c_ulonglong * 16)()
pts = []
for nt in (1024, 2048, 4096, 8192, 16384, 32768):
    q = rng.integers(0, 256, (pairs * 8192, 32), dtype=np.uint8)
    t = rng.integers(0, 256, (pairs * nt, 32), dtype=np.uint8)
    qo = np.arange(pairs + 1, dtype=np.int64) * 8192
    to = np.arange(pairs + 1, dtype=np.int64) * nt
    mcv.match_knn2_batch(q, qo, t, to, 0.8)
    lib.mcvDebugTcProf(out, 1)
    mcv.match_knn2_batch(q, qo, t, to, 0.8)
    lib.mcvDebugTcProf(out, 1)
    v = list(out)
    ctas = max(v[6], 1)
    stages = nt // 128
    pts.append((stages, v[5] / ctas))
    print(f"train rows {nt:6d} ({stages:4d} stages): {ctas} CTAs, {v[5] / ctas:9.0f} cycles per CTA, {v[5] / ctas / stages:7.0f} per stage")
    if v[7]:
        units = ctas * stages * 2
        print(f"    MMA issuer: wait b_full {v[0] / (ctas * stages):.0f} per stage, wait acc_empty {v[1] / units:.0f} per unit, "
              f"issue 8 MMAs + commit {v[7] / units:.0f} per unit")
        print(f"    epilogue warp 2 per own unit: wait acc_full {3 * v[2] / units:.0f}, tcgen05.ld+wait {3 * v[3] / units:.0f}, "
              f"ld+top-2 {3 * v[4] / units:.0f}")
(s0, c0), (s1, c1) = pts[-3], pts[-1]
slope = (c1 - c0) / (s1 - s0)
print(f"steady state {slope:.0f} cycles per stage (ideal 1024); fixed cost per CTA {c1 - slope * s1:.0f} cycles")
