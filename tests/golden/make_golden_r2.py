"""Round-2 goldens, frozen from cv2 4.13.0 (run in the build container: `python tests/golden/make_golden_r2.py`):

* C4-size cvRecoverPose: findEssentialMat(RANSAC) + recoverPose + decomposeEssentialMat at N = 5000 for both
  configurations of SURVEY.md section 8d (pixel units f = 800 / thr 1.0, and MiniCV's default normalised form
  f = 1 / thr 0.005, OpenCV.fs:33-34);
* cv2.findFundamentalMat(FM_RANSAC)'s small-N dispatch: N == 7 (7-point kernel, all solutions stacked, mask of ones),
  8 <= N < 15 (LMedS), and the parameter clamps (thr <= 0 -> 3, confidence outside (0, 1) -> 0.99);
* L2 kNN on general (non-integer) float descriptors.
"""
import sys
from pathlib import Path

import cv2
import numpy as np

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE.parent.parent))
from minicv_b200 import synth  # noqa: E402


def golden_e5k():
    out = {}
    cfgs = [(800.0, (320.0, 240.0), 0.999, 1.0, False), (1.0, (0.0, 0.0), 0.999, 0.005, True)]
    for k, (f, pp, prob, thr, norm) in enumerate(cfgs):
        a, b, gt, R, t = synth.two_view_pair(7, 5000, normalized=norm)
        Ec, mc = cv2.findEssentialMat(a, b, f, pp, cv2.RANSAC, prob, thr)
        cnt, Rc, tc, mk = cv2.recoverPose(Ec, a, b, focal=f, pp=pp, mask=mc.copy())
        R1, R2, tt = cv2.decomposeEssentialMat(Ec)
        out[f"e{k}_a"] = a; out[f"e{k}_b"] = b; out[f"e{k}_cfg"] = np.array([f, pp[0], pp[1], prob, thr])
        out[f"e{k}_E"] = Ec; out[f"e{k}_mask"] = mc.ravel(); out[f"e{k}_cnt"] = np.array(cnt)
        out[f"e{k}_R"] = Rc; out[f"e{k}_t"] = tc.ravel(); out[f"e{k}_posemask"] = mk.ravel()
        out[f"e{k}_R1"] = R1; out[f"e{k}_R2"] = R2; out[f"e{k}_tdec"] = tt.ravel()
        print("e5k", k, "inliers", int(mc.sum()), "cheirality", cnt)
    np.savez_compressed(HERE / "ransac_e5k.npz", **out)


def golden_f_small():
    out = {}
    k = 0
    for n in range(7, 16):
        for seed in range(3):
            a, b, gt, R, t = synth.two_view_pair(1000 + seed * 17 + n, n, inlier_frac=0.8)
            F, mask = cv2.findFundamentalMat(a, b, cv2.FM_RANSAC, 3.0, 0.99, 1000)
            if F is None:
                continue
            pad = np.zeros((9, 3)); pad[:F.shape[0]] = F
            out[f"s{k}_a"] = a; out[f"s{k}_b"] = b; out[f"s{k}_F"] = pad; out[f"s{k}_nF"] = np.array(F.shape[0] // 3)
            out[f"s{k}_mask"] = mask.ravel()
            k += 1
    out["s_n"] = np.array(k)
    a, b, *_ = synth.two_view_pair(5, 300)
    out["c_a"] = a; out["c_b"] = b
    for j, (thr, conf) in enumerate([(0.0, 0.99), (-1.0, 0.99), (3.0, 1.5), (3.0, 0.0), (3.0, 1.0)]):
        F, mask = cv2.findFundamentalMat(a, b, cv2.FM_RANSAC, thr, conf, 500)
        out[f"c{j}_prm"] = np.array([thr, conf]); out[f"c{j}_F"] = F; out[f"c{j}_mask"] = mask.ravel()
    out["c_n"] = np.array(5)
    np.savez_compressed(HERE / "ransac_f_small.npz", **out)
    print("f_small", k, "cases")


def golden_l2_general():
    """General float descriptors (not integer valued): distances of BFMatcher(NORM_L2) as this cv2 build computes
    them (its normL2Sqr accumulation order), k = 2."""
    out = {}
    rng = np.random.default_rng(77)
    for name, (nq, nt, d) in {"g128": (300, 700, 128), "g64": (200, 500, 64), "g61": (120, 333, 61), "g7": (50, 90, 7)}.items():
        t = rng.normal(size=(nt, d)).astype(np.float32)
        q = rng.normal(size=(nq, d)).astype(np.float32)
        q[::2] = t[rng.integers(0, nt, (nq + 1) // 2)] + rng.normal(scale=0.05, size=((nq + 1) // 2, d)).astype(np.float32)
        m = cv2.BFMatcher(cv2.NORM_L2).knnMatch(q, t, k=2)
        out[f"{name}_q"] = q; out[f"{name}_t"] = t
        out[f"{name}_idx"] = np.array([[x.trainIdx for x in r] for r in m], np.int32)
        out[f"{name}_dist"] = np.array([[x.distance for x in r] for r in m], np.float32)
    np.savez_compressed(HERE / "knn_l2_general.npz", **out)
    print("l2 general", len(out))


if __name__ == "__main__":
    print("cv2", cv2.__version__)
    assert cv2.__version__.startswith("4.13"), "goldens are pinned to cv2 4.13.0"
    which = sys.argv[1:] or ["e5k", "fsmall", "l2"]
    if "e5k" in which:
        golden_e5k()
    if "fsmall" in which:
        golden_f_small()
    if "l2" in which:
        golden_l2_general()
