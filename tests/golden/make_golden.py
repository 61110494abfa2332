"""Freeze golden vectors from cv2 4.13.0 (the OpenCV the reference shim forwards to; the
reference itself carries no fixtures for this path -- SURVEY.md section 8c).

Run in the build container:  python tests/golden/make_golden.py
Writes small .npz files next to this script.  The GPU box never runs this script and
never reads /root/reference; tests only load the committed .npz files."""
import sys
from pathlib import Path

import cv2
import numpy as np

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE.parent.parent))

from minicv_b200 import synth  # noqa: E402


def cv_knn2(q, t, norm):
    m = cv2.BFMatcher(norm).knnMatch(q, t, 2)
    idx = np.full((len(q), 2), -1, np.int32)
    dist = np.full((len(q), 2), np.finfo(np.float32).max, np.float32)
    for i, row in enumerate(m):
        for k, mm in enumerate(row):
            assert mm.queryIdx == i and mm.imgIdx == 0
            idx[i, k] = mm.trainIdx
            dist[i, k] = mm.distance
    return idx, dist


def golden_knn():
    out = {}
    for name, nb in (("b256", 32), ("b488", 61), ("b512", 64)):
        q, t, _ = synth.descriptor_pair(7, 600, 900, nb)
        # engineered ties: duplicated train rows (3-way), exact-copy queries
        t[5] = t[3]; t[9] = t[3]; t[700] = t[3]
        q[0] = t[3]
        q[1] = t[9]; q[1, 0] ^= 0x10
        t[899] = t[0]            # tie between first and last train row
        q[2] = t[0]
        idx, dist = cv_knn2(q, t, cv2.NORM_HAMMING)
        out[f"{name}_q"] = q; out[f"{name}_t"] = t
        out[f"{name}_idx"] = idx; out[f"{name}_dist"] = dist
    # tiny train sets: 1 row -> only one neighbour exists
    q, t, _ = synth.descriptor_pair(11, 17, 1, 32)
    idx, dist = cv_knn2(q, t, cv2.NORM_HAMMING)
    out["one_q"] = q; out["one_t"] = t; out["one_idx"] = idx; out["one_dist"] = dist
    np.savez_compressed(HERE / "knn_hamming.npz", **out)
    print("knn_hamming.npz", {k: v.shape for k, v in out.items() if k.endswith("_idx")})


if __name__ == "__main__":
    print("cv2", cv2.__version__)
    assert cv2.__version__.startswith("4.13"), "goldens are pinned to cv2 4.13.0"
    which = sys.argv[1:] or ["knn"]
    if "knn" in which:
        golden_knn()
