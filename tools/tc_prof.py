"""Per-CTA cycle counter of hamming_tc_kernel against the train-set size: the slope is the steady-state
cost of one 128-row train stage, the intercept the fixed cost of a work item (query staging, pipeline hand-over between items).  Needs a library built with -DMCV_TC_PROF=1 (2 adds wait counters):
   make LIB=build/libprof.so OBJDIR=build/objprof EXTRA=-DMCV_TC_PROF=1
   MINICV_NATIVE_LIB=build/libprof.so python tools/tc_prof.py"""
import ctypes as C
import sys
from pathlib import Path
import numpy as np
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import minicv_b200 as mcv  # noqa: E402
from minicv_b200 import native  # noqa: E402

pairs = int(sys.argv[1]) if len(sys.argv) > 1 else 74      # 74 pairs x 32 tiles = 16 work items per SM
rng = np.random.default_rng(1)
mcv.set_hamming_engine(mcv.ENGINE_TENSOR)
lib = native.load()
out = (C.c_ulonglong * 16)()
pts = []
for nt in (1024, 2048, 4096, 8192, 16384, 32768):
    if pairs * nt > 16 * 1024 * 1024:
        break
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
    items = pairs * 32                     # 256-row query tiles per launch (persistent CTAs walk them)
    stages = nt // 128
    per_item = v[5] / items                # CTA-cycles per work item
    pts.append((stages, per_item))
    print(f"train rows {nt:6d} ({stages:4d} stages): {ctas} CTAs, {items} items, {per_item:9.0f} cycles per item, "
          f"{per_item / stages:7.0f} per stage")
if len(pts) >= 3:
    (s0, c0), (s1, c1) = pts[-3], pts[-1]
    slope = (c1 - c0) / (s1 - s0)
    print(f"steady state {slope:.0f} cycles per stage (ideal 1024); fixed cost per work item {c1 - slope * s1:.0f} cycles")
