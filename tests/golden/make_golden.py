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


def golden_l2():
    out = {}
    rng = np.random.default_rng(5)
    # SIFT-like: integer-valued float32 in [0, 255] -> every partial sum is exact in float32
    t = rng.integers(0, 120, (700, 128)).astype(np.float32)
    q = rng.integers(0, 120, (300, 128)).astype(np.float32)
    q[:100] = t[rng.permutation(700)[:100]] + rng.integers(-3, 4, (100, 128)).astype(np.float32)
    t[9] = t[3]; q[0] = t[3]                       # exact tie -> lowest index
    idx, dist = cv_knn2(q, t, cv2.NORM_L2)
    out["sift_q"] = q; out["sift_t"] = t; out["sift_idx"] = idx; out["sift_dist"] = dist
    # general floats
    t = rng.normal(size=(500, 64)).astype(np.float32)
    q = rng.normal(size=(200, 64)).astype(np.float32)
    idx, dist = cv_knn2(q, t, cv2.NORM_L2)
    out["gen_q"] = q; out["gen_t"] = t; out["gen_idx"] = idx; out["gen_dist"] = dist
    np.savez_compressed(HERE / "knn_l2.npz", **out)
    print("knn_l2.npz")


def golden_c1():
    """Config C1 (BASELINE.json configs[0]): procedurally textured plane seen under a known
    homography; ORB and AKAZE detect+describe on the CPU (cv2 -- detection is out of scope of the
    path), then BFMatcher k=2 + ratio 0.8 and findHomography RANSAC as the reference results."""
    rng = np.random.default_rng(42)
    h, w = 480, 640
    img = np.zeros((h, w), np.float32)
    for _ in range(900):                                  # random blobs, rectangles and strokes
        x, y = int(rng.integers(0, w)), int(rng.integers(0, h))
        r = int(rng.integers(3, 22)); v = float(rng.uniform(40, 255))
        kind = rng.integers(0, 3)
        if kind == 0:
            cv2.circle(img, (x, y), r, v, -1)
        elif kind == 1:
            cv2.rectangle(img, (x, y), (x + r * 2, y + r), v, -1)
        else:
            cv2.line(img, (x, y), (int(x + rng.integers(-60, 60)), int(y + rng.integers(-60, 60))), v, int(rng.integers(1, 4)))
    img = cv2.GaussianBlur(img, (0, 0), 1.0)
    img1 = np.clip(img + rng.normal(0, 2, img.shape), 0, 255).astype(np.uint8)
    Hgt = np.array([[0.96, 0.07, 18.0], [-0.06, 1.02, 9.0], [-6e-5, 4e-5, 1.0]])
    img2 = cv2.warpPerspective(img1, Hgt, (w, h), flags=cv2.INTER_LINEAR, borderValue=0)
    img2 = np.clip(img2.astype(np.float32) + rng.normal(0, 2, img2.shape), 0, 255).astype(np.uint8)
    out = {"Hgt": Hgt}
    for name, det in (("orb", cv2.ORB_create(nfeatures=3000)), ("akaze", cv2.AKAZE_create())):
        k1, d1 = det.detectAndCompute(img1, None)
        k2, d2 = det.detectAndCompute(img2, None)
        p1 = np.array([k.pt for k in k1], np.float64); p2 = np.array([k.pt for k in k2], np.float64)
        idx, dist = cv_knn2(d1, d2, cv2.NORM_HAMMING)
        ok = (idx[:, 1] >= 0) & (dist[:, 0] < np.float32(0.8) * dist[:, 1])
        a = p1[ok]; b = p2[idx[ok, 0]]
        Hc, mc = cv2.findHomography(a, b, cv2.RANSAC, 3.0, maxIters=2000, confidence=0.995)
        out[f"{name}_d1"] = d1; out[f"{name}_d2"] = d2; out[f"{name}_p1"] = p1; out[f"{name}_p2"] = p2
        out[f"{name}_idx"] = idx; out[f"{name}_dist"] = dist; out[f"{name}_H"] = Hc; out[f"{name}_mask"] = mc.ravel()
        print(name, d1.shape, d2.shape, "ratio-pass", int(ok.sum()), "inliers", int(mc.sum()),
              "H err", np.abs(Hc - Hgt).max())
    np.savez_compressed(HERE / "c1_planar.npz", **out)


def golden_ransac():
    out = {}
    # homography: cv2.findHomography(RANSAC, 3.0, maxIters 2000, confidence .995)
    for k, (seed, n) in enumerate([(0, 400), (1, 700), (4, 2000), (5, 333)]):
        a, b, gt, H = synth.homography_pair(seed, n)
        Hc, mc = cv2.findHomography(a, b, cv2.RANSAC, 3.0, maxIters=2000, confidence=0.995)
        out[f"h{k}_a"] = a; out[f"h{k}_b"] = b; out[f"h{k}_H"] = Hc; out[f"h{k}_mask"] = mc.ravel()
    # 4-point kernel (method 0 on 4 points == kernel only)
    A, B, HH = [], [], []
    for seed in range(24):
        a, b, gt, H = synth.homography_pair(100 + seed, 4, inlier_frac=1.0)
        Hc, _ = cv2.findHomography(a.astype(np.float32), b.astype(np.float32), 0)
        if Hc is None:
            continue
        A.append(a.astype(np.float32)); B.append(b.astype(np.float32)); HH.append(Hc)
    out["h4_a"] = np.array(A); out["h4_b"] = np.array(B); out["h4_H"] = np.array(HH)
    # fundamental: FM_RANSAC 3.0 / .99 / 1000
    for k, (seed, n) in enumerate([(0, 300), (1, 500), (3, 2000)]):
        a, b, gt, R, t = synth.two_view_pair(seed, n)
        Fc, mc = cv2.findFundamentalMat(a, b, cv2.FM_RANSAC, 3.0, 0.99, 1000)
        out[f"f{k}_a"] = a; out[f"f{k}_b"] = b; out[f"f{k}_F"] = Fc; out[f"f{k}_mask"] = mc.ravel()
    # 7-point minimal sets (stacked models; order is OpenCV's)
    A, B, FF, NN = [], [], [], []
    for seed in range(24):
        a, b, gt, R, t = synth.two_view_pair(200 + seed, 7, inlier_frac=1.0)
        Fc, _ = cv2.findFundamentalMat(a, b, cv2.FM_7POINT)
        if Fc is None:
            continue
        pad = np.zeros((9, 3)); pad[:Fc.shape[0]] = Fc
        A.append(a.astype(np.float32)); B.append(b.astype(np.float32)); FF.append(pad); NN.append(Fc.shape[0] // 3)
    out["f7_a"] = np.array(A); out["f7_b"] = np.array(B); out["f7_F"] = np.array(FF); out["f7_n"] = np.array(NN)
    np.savez_compressed(HERE / "ransac_hf.npz", **out)
    print("ransac_hf.npz", len(out), "arrays")

    out = {}
    cfgs = [(800.0, (320.0, 240.0), 0.999, 1.0, False), (1.0, (0.0, 0.0), 0.999, 0.005, True)]
    k = 0
    for seed, n in [(0, 300), (1, 600), (2, 1500)]:
        for (f, pp, prob, thr, norm) in cfgs:
            a, b, gt, R, t = synth.two_view_pair(seed, n, normalized=norm)
            Ec, mc = cv2.findEssentialMat(a, b, f, pp, cv2.RANSAC, prob, thr)
            cnt, Rc, tc, mk = cv2.recoverPose(Ec, a, b, focal=f, pp=pp, mask=mc.copy())
            R1, R2, tt = cv2.decomposeEssentialMat(Ec)
            out[f"e{k}_a"] = a; out[f"e{k}_b"] = b; out[f"e{k}_cfg"] = np.array([f, pp[0], pp[1], prob, thr])
            out[f"e{k}_E"] = Ec; out[f"e{k}_mask"] = mc.ravel(); out[f"e{k}_cnt"] = np.array(cnt)
            out[f"e{k}_R"] = Rc; out[f"e{k}_t"] = tc.ravel(); out[f"e{k}_posemask"] = mk.ravel()
            out[f"e{k}_R1"] = R1; out[f"e{k}_R2"] = R2; out[f"e{k}_tdec"] = tt.ravel()
            k += 1
    out["e_n"] = np.array(k)
    # 5-point minimal sets: findEssentialMat on exactly 5 points returns every solution stacked
    A, B, EE, NN = [], [], [], []
    for seed in range(40):
        a, b, gt, R, t = synth.two_view_pair(300 + seed, 5, inlier_frac=1.0,
                                             noise=0.0 if seed % 2 else 0.3, normalized=True)
        Ec, _ = cv2.findEssentialMat(a, b, 1.0, (0.0, 0.0), cv2.RANSAC, 0.999, 1.0)
        n = 0 if Ec is None else Ec.shape[0] // 3
        pad = np.zeros((30, 3))
        if n:
            pad[:3 * n] = Ec
        A.append(a); B.append(b); EE.append(pad); NN.append(n)
    out["min5_a"] = np.array(A); out["min5_b"] = np.array(B); out["min5_E"] = np.array(EE); out["min5_n"] = np.array(NN)
    np.savez_compressed(HERE / "ransac_e.npz", **out)
    print("ransac_e.npz", len(out), "arrays")


if __name__ == "__main__":
    print("cv2", cv2.__version__)
    assert cv2.__version__.startswith("4.13"), "goldens are pinned to cv2 4.13.0"
    which = sys.argv[1:] or ["knn"]
    if "knn" in which:
        golden_knn()
    if "ransac" in which:
        golden_ransac()
    if "l2" in which:
        golden_l2()
    if "c1" in which:
        golden_c1()
