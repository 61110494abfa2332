"""A few calls of the tcgen05 L2 matcher on the bench's M2 shape (256 pairs x 2048 x 2048 x 128 f32) --
for an ncu launch list / capture of l2_tc_kernel and its helpers."""
import sys
from pathlib import Path
import numpy as np
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import minicv_b200 as mcv  # noqa: E402

P, n, d = 256, 2048, 128
rng = np.random.default_rng(3)
t = rng.integers(0, 256, (P * n, d)).astype(np.float32)
q = t.copy()
off = np.arange(P + 1, dtype=np.int64) * n
mcv.set_hamming_engine(mcv.ENGINE_TENSOR)
for _ in range(3):
    r = mcv.match_knn2_batch(q, off, t, off, 0.8, norm=mcv.NORM_L2)
print("ok", int((r.train_idx[:, 0] == np.arange(P * n) % n).sum()), "of", P * n)
