"""Freeze golden vectors for the PnP row.

Two sources:
  * cv2 4.13.0 (the OpenCV ``cvSolvePnPRansac`` / ``cvSolvePnP`` forward to, MiniCVNative.cpp:48-139):
    solveP3P / solvePnP(AP3P, 4 points) / solvePnP(EPNP) / projectPoints / Rodrigues / solvePnPRansac(AP3P);
  * the reference's own ``solveAp3p`` (src/MiniCVNative/ap3p.cpp:282) compiled in place by
    ``oracle/ref_build/Makefile`` into ``oracle/_ref/libap3p_ref.so``.

Run in the build container:  make -C oracle/ref_build && python tests/golden/make_golden_pnp.py
The GPU box never runs this script and never reads /root/reference; tests only load pnp.npz."""
import ctypes as C
import sys
from pathlib import Path

import cv2
import numpy as np

HERE = Path(__file__).resolve().parent
ROOT = HERE.parent.parent
sys.path.insert(0, str(ROOT))

from minicv_b200 import synth  # noqa: E402

K = np.array([[800.0, 0, 320.0], [0, 810.0, 240.0], [0, 0, 1.0]])
DISTS = {"d0": np.zeros(4), "d1": np.array([0.05, -0.02, 0.001, -0.002])}


def main():
    out = {"K": K}
    for k, v in DISTS.items():
        out[f"dist_{k}"] = v
    rng = np.random.default_rng(2026)

    # --- reference solveAp3p (compiled from /root/reference) -------------------------------------
    lib = C.CDLL(str(ROOT / "oracle" / "_ref" / "libap3p_ref.so"))
    lib.solveAp3p.restype = C.c_int
    lib.solveAp3p.argtypes = [C.c_void_p, C.c_void_p] + [C.c_double] * 19
    n_ap = 256
    ap_img = np.zeros((n_ap, 3, 2)); ap_obj = np.zeros((n_ap, 3, 3))
    ap_n = np.zeros(n_ap, np.int32); ap_R = np.zeros((n_ap, 4, 9)); ap_t = np.zeros((n_ap, 4, 3))
    for i in range(n_ap):
        X, uv, _, _ = synth.pnp_scene(rng, 3, 1.0, noise=0.0, K=K)
        ap_img[i], ap_obj[i] = uv, X
        args = []
        for j in range(3):
            args += [uv[j, 0], uv[j, 1], X[j, 0], X[j, 1], X[j, 2]]
        ap_n[i] = lib.solveAp3p(ap_R[i].ctypes.data, ap_t[i].ctypes.data, *args,
                                1 / K[0, 0], 1 / K[1, 1], K[0, 2] / K[0, 0], K[1, 2] / K[1, 1])
    out.update(ap3p_img=ap_img, ap3p_obj=ap_obj, ap3p_n=ap_n, ap3p_R=ap_R, ap3p_t=ap_t)

    # --- cv2 building blocks --------------------------------------------------------------------------
    rv = rng.normal(size=(64, 3)) * rng.uniform(0, 2, (64, 1))
    out["rod_rvec"] = rv
    out["rod_R"] = np.stack([cv2.Rodrigues(r)[0] for r in rv])
    Rn = out["rod_R"] + rng.normal(size=(64, 3, 3)) * 1e-9
    out["rod_Rin"] = Rn
    out["rod_rvec_back"] = np.stack([cv2.Rodrigues(r)[0].ravel() for r in Rn])
    for dk, dist in DISTS.items():
        X, uv, rvec, tvec = synth.pnp_scene(rng, 400, 1.0, noise=0.5, K=K)
        X32 = X.astype(np.float32)
        out[f"proj_{dk}_obj"] = X32; out[f"proj_{dk}_rvec"] = rvec; out[f"proj_{dk}_tvec"] = tvec
        out[f"proj_{dk}_f32"] = cv2.projectPoints(X32, rvec, tvec, K, dist)[0].reshape(-1, 2)
        out[f"proj_{dk}_f64"] = cv2.projectPoints(X32.astype(np.float64), rvec, tvec, K, dist)[0].reshape(-1, 2)
        u32 = uv.astype(np.float32)
        out[f"und_{dk}_in"] = u32
        out[f"und_{dk}_out"] = cv2.undistortPoints(u32.reshape(-1, 1, 2), K, dist).reshape(-1, 2)
        # 4-point AP3P (float32 inputs, as inside the RANSAC loop)
        n4 = 96
        o4 = np.zeros((n4, 4, 3), np.float32); i4 = np.zeros((n4, 4, 2), np.float32)
        ok4 = np.zeros(n4, np.uint8); r4 = np.zeros((n4, 3)); t4 = np.zeros((n4, 3))
        for i in range(n4):
            X, uv, _, _ = synth.pnp_scene(rng, 4, 1.0, noise=0.5, K=K)
            o4[i], i4[i] = X, uv
            ok, rvec, tvec = cv2.solvePnP(o4[i], i4[i], K, dist, flags=cv2.SOLVEPNP_AP3P)
            ok4[i] = ok
            if ok:
                r4[i], t4[i] = rvec.ravel(), tvec.ravel()
        out.update({f"p4_{dk}_obj": o4, f"p4_{dk}_img": i4, f"p4_{dk}_ok": ok4, f"p4_{dk}_rvec": r4, f"p4_{dk}_tvec": t4})
        # EPnP
        for n in (6, 40, 300):
            X, uv, _, _ = synth.pnp_scene(rng, n, 1.0, noise=0.5, K=K)
            ok, rvec, tvec = cv2.solvePnP(X, uv, K, dist, flags=cv2.SOLVEPNP_EPNP)
            assert ok
            out.update({f"epnp_{dk}_{n}_obj": X, f"epnp_{dk}_{n}_img": uv,
                        f"epnp_{dk}_{n}_rvec": rvec.ravel(), f"epnp_{dk}_{n}_tvec": tvec.ravel()})
        # full solvePnPRansac(AP3P)
        for ci, (n, inl, iters, thr, conf) in enumerate(((300, 0.5, 300, 3.0, 0.99), (2000, 0.3, 1000, 2.0, 0.999),
                                                         (60, 0.6, 100, 8.0, 0.99), (5, 1.0, 100, 8.0, 0.99),
                                                         (4, 1.0, 100, 8.0, 0.99))):
            X, uv, _, _ = synth.pnp_scene(rng, n, inl, noise=0.5, K=K)
            ok, rvec, tvec, inliers = cv2.solvePnPRansac(X, uv, K, dist, iterationsCount=iters, reprojectionError=thr,
                                                         confidence=conf, flags=cv2.SOLVEPNP_AP3P)
            key = f"ransac_{dk}_{ci}"
            out.update({f"{key}_obj": X, f"{key}_img": uv, f"{key}_prm": np.array([iters, thr, conf]),
                        f"{key}_ok": np.array([int(ok)]),
                        f"{key}_rvec": rvec.ravel() if ok else np.zeros(3),
                        f"{key}_tvec": tvec.ravel() if ok else np.zeros(3),
                        f"{key}_inliers": inliers.ravel().astype(np.int32) if inliers is not None else np.zeros(0, np.int32)})
            print(key, n, ok, 0 if inliers is None else len(inliers))
    np.savez_compressed(HERE / "pnp.npz", **out)
    print("pnp.npz", len(out), "arrays", (HERE / "pnp.npz").stat().st_size, "bytes")


if __name__ == "__main__":
    main()
