"""CPU: the device math header (minicv_b200/csrc/cvmath.cuh) compiled for the host and checked
against the oracle -- same source the CUDA kernels include.  Test-only build, not a fallback."""
import ctypes as C
import subprocess
from pathlib import Path

import numpy as np
import pytest

from minicv_b200 import synth
from oracle import essential, fivepoint, ransac
from conftest import GOLDEN, ROOT


@pytest.fixture(scope="module")
def hm():
    out = ROOT / "build" / "libmcv_hostmath_test.so"
    out.parent.mkdir(exist_ok=True)
    src = ROOT / "tests" / "hostmath" / "hostmath.cpp"
    subprocess.run(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-shared", "-fPIC", "-x", "c++",
                    "-o", str(out), str(src)], check=True)
    lib = C.CDLL(str(out))
    lib.hm_update_num_iters.argtypes = [C.c_double, C.c_double, C.c_int, C.c_int]
    return lib


def p(a):
    return a.ctypes.data_as(C.c_void_p)


def test_scoring_functions_bit_exact(hm):
    a, b, gt, H = synth.homography_pair(1, 2000)
    af, bf = a.astype(np.float32), b.astype(np.float32)
    Hn = (H / H[2, 2]).ravel().copy()
    err = np.zeros(2000, np.float32)
    hm.hm_h_error(p(Hn), p(af), p(bf), 2000, p(err))
    np.testing.assert_array_equal(err, ransac.HomographyCB.compute_error(af, bf, Hn))
    g = np.load(GOLDEN / "ransac_hf.npz")
    af, bf = g["f2_a"].astype(np.float32), g["f2_b"].astype(np.float32)
    F = g["f2_F"].ravel().copy()
    err = np.zeros(len(af), np.float32)
    hm.hm_f_error(p(F), p(af), p(bf), len(af), p(err))
    np.testing.assert_array_equal(err, ransac.FundamentalCB.compute_error(af, bf, F))
    ge = np.load(GOLDEN / "ransac_e.npz")
    x1 = essential.normalize_points(ge["e4_a"], 800.0, (320.0, 240.0))
    x2 = essential.normalize_points(ge["e4_b"], 800.0, (320.0, 240.0))
    E = ge["e4_E"].ravel().copy()
    err = np.zeros(len(x1), np.float32)
    hm.hm_e_error(p(E), p(x1), p(x2), len(x1), p(err))
    np.testing.assert_array_equal(err, essential.EssentialCB.compute_error(x1, x2, E))


def test_h_solve4_and_check_subset(hm):
    rng = np.random.default_rng(0)
    for seed in range(60):
        a, b, gt, H = synth.homography_pair(seed, 50, inlier_frac=1.0)
        idx = rng.permutation(50)[:4]
        s, d = a[idx].astype(np.float32), b[idx].astype(np.float32)
        assert bool(hm.hm_h_check_subset(p(s), p(d))) == ransac.HomographyCB.check_subset(s, d, 4)
        Hh = np.zeros(9)
        ok = hm.hm_h_solve4(p(s), p(d), p(Hh))
        ms = ransac.HomographyCB.run_kernel(s, d)
        if ok and ms:
            assert np.abs(Hh.reshape(3, 3) - ms[0]).max() / np.abs(ms[0]).max() < 1e-6


def test_f_solve7_sets(hm):
    g = np.load(GOLDEN / "ransac_hf.npz")
    for a, b, Fc, n in zip(g["f7_a"], g["f7_b"], g["f7_F"], g["f7_n"]):
        F = np.zeros(27)
        cnt = hm.hm_f_solve7(p(np.ascontiguousarray(a)), p(np.ascontiguousarray(b)), p(F))
        assert cnt == n
        for i in range(n):
            ref = Fc[3 * i:3 * i + 3]
            assert min(np.abs(F[9 * k:9 * k + 9].reshape(3, 3) - ref).max() for k in range(cnt)) / np.abs(ref).max() < 1e-6


def test_e_solve5_golden(hm):
    g = np.load(GOLDEN / "ransac_e.npz")
    tot = bad = 0
    for a, b, Ec, n in zip(g["min5_a"], g["min5_b"], g["min5_E"], g["min5_n"]):
        Es = np.zeros(90)
        cnt = hm.hm_e_solve5(p(np.ascontiguousarray(a)), p(np.ascontiguousarray(b)), p(Es))
        assert cnt == n
        for i in range(n):
            tot += 1
            bad += min(np.abs(Es[9 * k:9 * k + 9].reshape(3, 3) - Ec[3 * i:3 * i + 3]).max() for k in range(cnt)) > 1e-6
    assert tot > 100 and bad / tot < 0.02


def test_pose_math(hm):
    g = np.load(GOLDEN / "ransac_e.npz")
    E = g["e0_E"].ravel().copy()
    R1, R2, t = np.zeros(9), np.zeros(9), np.zeros(3)
    hm.hm_decompose_essential(p(E), p(R1), p(R2), p(t))
    Ro1, Ro2, to = essential.decompose_essential(g["e0_E"])
    pairs = [(R1.reshape(3, 3), R2.reshape(3, 3)), (R2.reshape(3, 3), R1.reshape(3, 3))]
    assert min(np.abs(x - Ro1).max() + np.abs(y - Ro2).max() for x, y in pairs) < 1e-9
    x1 = essential.normalize_points(g["e0_a"], 800.0, (320.0, 240.0))
    x2 = essential.normalize_points(g["e0_b"], 800.0, (320.0, 240.0))
    for Rr, ts in [(Ro1, to), (Ro2, to), (Ro1, -to), (Ro2, -to)]:
        ok = np.zeros(len(x1), np.uint8)
        hm.hm_cheirality(p(Rr.ravel().copy()), p(ts.copy()), p(x1), p(x2), len(x1), C.c_double(50.0), p(ok))
        np.testing.assert_array_equal(ok.astype(bool), essential.cheirality_mask(Rr, ts, x1, x2))


def test_rng_and_iteration_update(hm):
    out = np.zeros(64, np.uint32)
    hm.hm_mwc(64, p(out))
    r = ransac.MWC()
    assert out.tolist() == [r.next() for _ in range(64)]
    for ep in np.linspace(0, 1, 401):
        for mp, mi in ((4, 2000), (7, 1000), (5, 1000)):
            assert hm.hm_update_num_iters(0.995, ep, mp, mi) == ransac.update_num_iters(0.995, ep, mp, mi)


def test_division_free_inlier_tests_match_exact_errors(hm):
    """f_inlier / e_inlier == (error <= thr) for every point, including thresholds placed exactly
    on, one ulp below and one ulp above individual error values (the borderline sliver)."""
    g = np.load(GOLDEN / "ransac_hf.npz")
    af, bf = g["f2_a"].astype(np.float32), g["f2_b"].astype(np.float32)
    F = g["f2_F"].ravel().copy()
    err = ransac.FundamentalCB.compute_error(af, bf, F)
    n = len(af)
    picks = np.concatenate([[np.float32(9.0)], err[:40], np.nextafter(err[:40], np.float32(0)),
                            np.nextafter(err[:40], np.float32(np.inf))]).astype(np.float32)
    for thr in picks:
        out = np.zeros(n, np.uint8)
        hm.hm_f_inlier(p(F), p(af), p(bf), n, C.c_float(float(thr)), p(out))
        np.testing.assert_array_equal(out.astype(bool), err <= thr)
    ge = np.load(GOLDEN / "ransac_e.npz")
    x1 = essential.normalize_points(ge["e4_a"], 800.0, (320.0, 240.0))
    x2 = essential.normalize_points(ge["e4_b"], 800.0, (320.0, 240.0))
    E = ge["e4_E"].ravel().copy()
    err = essential.EssentialCB.compute_error(x1, x2, E)
    n = len(x1)
    picks = np.concatenate([[np.float32((1.0 / 800) ** 2)], err[:40], np.nextafter(err[:40], np.float32(0)),
                            np.nextafter(err[:40], np.float32(np.inf))]).astype(np.float32)
    for thr in picks:
        out = np.zeros(n, np.uint8)
        hm.hm_e_inlier(p(E), p(x1), p(x2), n, C.c_float(float(thr)), p(out))
        np.testing.assert_array_equal(out.astype(bool), err <= thr)


def test_symmetric_solvers(hm):
    rng = np.random.default_rng(2)
    for cond in (1.0, 1e2, 1e13):
        B = rng.normal(size=(30, 8))
        B[:, 7] = B[:, 6] + B[:, 7] / cond          # near-dependent columns for the large cond cases
        A = B.T @ B
        b = rng.normal(size=8)
        x1 = np.zeros(8); x2 = np.zeros(8)
        hm.hm_solve_sym_eig8(p(A.copy()), p(b), p(x1))
        hm.hm_solve_sym_chol8(p(A.copy()), p(b), p(x2))
        if cond < 1e10:
            ref = np.linalg.solve(A, b)
            assert np.abs(x1 - ref).max() / np.abs(ref).max() < 1e-6
            assert np.abs(x2 - ref).max() / np.abs(ref).max() < 1e-6
        else:   # ill-conditioned: the Cholesky guard must hand over to the eigen-solve
            np.testing.assert_array_equal(x1, x2)


# ---- PnP row: pnpmath.cuh against the oracle and the cv2 / compiled-reference goldens ----------------
def _cam(K, dist):
    return np.array([K[0, 0], K[1, 1], K[0, 2], K[1, 2], dist[0], dist[1], dist[2], dist[3]], np.float64)


def test_pnp_ap3p_reference_export(hm):
    """ap3p_poses (reference mode) == the reference's solveAp3p compiled from ap3p.cpp (frozen golden)."""
    g = np.load(GOLDEN / "pnp.npz")
    K = g["K"]
    hm.hm_ap3p.argtypes = [C.c_void_p, C.c_void_p, C.c_double, C.c_double, C.c_double, C.c_double, C.c_void_p, C.c_void_p]
    for i in range(len(g["ap3p_n"])):
        Rs = np.zeros((4, 9)); ts = np.zeros((4, 3))
        img = np.ascontiguousarray(g["ap3p_img"][i]); obj = np.ascontiguousarray(g["ap3p_obj"][i])
        n = hm.hm_ap3p(p(img), p(obj), 1 / K[0, 0], 1 / K[1, 1], K[0, 2] / K[0, 0], K[1, 2] / K[1, 1], p(Rs), p(ts))
        assert n == int(g["ap3p_n"][i])
        np.testing.assert_allclose(Rs[:n], g["ap3p_R"][i, :n], rtol=0, atol=1e-8)
        np.testing.assert_allclose(ts[:n], g["ap3p_t"][i, :n], rtol=0, atol=1e-8)


@pytest.mark.parametrize("dk", ["d0", "d1"])
def test_pnp_building_blocks(hm, dk):
    g = np.load(GOLDEN / "pnp.npz")
    cam = _cam(g["K"], g[f"dist_{dk}"])
    # Rodrigues both ways
    for rv, R in zip(g["rod_rvec"], g["rod_R"]):
        out = np.zeros(9)
        hm.hm_rodrigues_v2m(p(np.ascontiguousarray(rv)), p(out))
        assert np.array_equal(out.reshape(3, 3), R)
    for Rin, rv in zip(g["rod_Rin"], g["rod_rvec_back"]):
        out = np.zeros(3)
        hm.hm_rodrigues_m2v(p(np.ascontiguousarray(Rin)), p(out))
        np.testing.assert_allclose(out, rv, rtol=0, atol=1e-12)
    # undistortPoints, float32 in / out: bit-exact
    uv = np.ascontiguousarray(g[f"und_{dk}_in"]); out = np.zeros_like(uv)
    hm.hm_undistort(p(cam), p(uv), len(uv), p(out))
    assert np.array_equal(out, g[f"und_{dk}_out"])
    # projectPoints + squared float32 residual: bit-exact vs cv2's float32 projection
    obj = np.ascontiguousarray(g[f"proj_{dk}_obj"]); rvec = g[f"proj_{dk}_rvec"].copy(); tvec = g[f"proj_{dk}_tvec"].copy()
    img = np.ascontiguousarray((g[f"proj_{dk}_f32"] + np.float32(1.25)).astype(np.float32))
    err = np.zeros(len(obj), np.float32)
    hm.hm_project_error(p(cam), p(rvec), p(tvec), p(obj), p(img), len(obj), p(err))
    d = img - g[f"proj_{dk}_f32"]
    assert np.array_equal(err, d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1])
    # solvePnP(AP3P) on four float32 points
    close = 0
    n4 = len(g[f"p4_{dk}_ok"])
    for i in range(n4):
        rv = np.zeros(3); tv = np.zeros(3)
        ok = hm.hm_pnp_solve4(p(cam), p(np.ascontiguousarray(g[f"p4_{dk}_obj"][i])), p(np.ascontiguousarray(g[f"p4_{dk}_img"][i])), p(rv), p(tv))
        assert bool(ok) == bool(g[f"p4_{dk}_ok"][i])
        if ok:
            close += max(np.abs(rv - g[f"p4_{dk}_rvec"][i]).max(), np.abs(tv - g[f"p4_{dk}_tvec"][i]).max()) < 1e-6
    assert close >= n4 - 2
    # EPnP
    for n in (6, 40, 300):
        rv = np.zeros(3); tv = np.zeros(3)
        hm.hm_epnp(p(cam), p(np.ascontiguousarray(g[f"epnp_{dk}_{n}_obj"])), p(np.ascontiguousarray(g[f"epnp_{dk}_{n}_img"])), n, p(rv), p(tv))
        np.testing.assert_allclose(rv, g[f"epnp_{dk}_{n}_rvec"], rtol=0, atol=1e-8)
        np.testing.assert_allclose(tv, g[f"epnp_{dk}_{n}_tvec"], rtol=0, atol=1e-8)


# ---- fp32 filter in front of the exact fundamental-matrix inlier test --------------------------------
def test_f_filter_is_conservative(hm):
    """f_certain_outlier may only fire where f_inlier is false -- on realistic models, on models that almost fit
    (where the decision is within rounding of the threshold), and on badly scaled / degenerate / non-finite ones --
    and it must remove most of the work on a wrong model."""
    rng = np.random.default_rng(7)
    a, b, gt, R, t = synth.two_view_pair(2, 4000)
    af, bf = np.ascontiguousarray(a, np.float32), np.ascontiguousarray(b, np.float32)
    g = np.load(GOLDEN / "ransac_hf.npz")
    Fgood = g["f2_F"].ravel()
    models = []
    # 7-point models from random samples of this very pair (what RANSAC scores)
    F7 = np.zeros(27)
    for _ in range(300):
        idx = rng.permutation(len(af))[:7]
        n = hm.hm_f_solve7(p(np.ascontiguousarray(af[idx])), p(np.ascontiguousarray(bf[idx])), p(F7))
        for k in range(n):
            models.append(F7[9 * k:9 * k + 9].copy())
    # the same with inlier-only samples (near-perfect models: many decisions close to the threshold)
    inl = np.flatnonzero(gt)
    for _ in range(100):
        idx = rng.choice(inl, 7, replace=False)
        n = hm.hm_f_solve7(p(np.ascontiguousarray(af[idx])), p(np.ascontiguousarray(bf[idx])), p(F7))
        for k in range(n):
            models.append(F7[9 * k:9 * k + 9].copy())
    base = models[0]
    # scalings over the whole exponent range, degenerate rows, non-finite entries
    for sc in (1e-40, 1e-33, 1e-20, 1e-6, 1e6, 1e20, 1e29, 1e35, 1e300):
        models.append(base * sc)
    z = base.copy(); z[:6] = 0; models.append(z)
    z = base.copy(); z[:6] *= 1e-25; models.append(z)
    z = base.copy(); z[6:] = 0; models.append(z)
    z = base.copy(); z[4] = np.nan; models.append(z)
    z = base.copy(); z[8] = np.inf; models.append(z)
    models.append(np.zeros(9))
    models.append(rng.normal(0, 1, 9))
    M = np.ascontiguousarray(np.stack(models))
    n = len(af)
    for thr in (9.0, 1.0, 1e-4):
        filt = np.zeros((len(M), n), np.uint8); exact = np.zeros((len(M), n), np.uint8)
        hm.hm_f_filter(p(M), len(M), p(af), p(bf), n, C.c_float(thr), p(filt), p(exact))
        assert not (filt & exact).any(), "filter rejected an inlier"
        if thr == 9.0:
            wrong = exact[:300].sum(axis=1) < 0.05 * n          # wrong models among the first random-sample ones
            assert filt[:300][wrong].mean() > 0.95              # >95 % of their evaluations never reach fp64
    # points far outside the image, tiny and huge coordinates
    af2 = (af * rng.choice([1e-20, 1e-3, 1.0, 1e3, 1e15], size=(n, 1))).astype(np.float32)
    bf2 = (bf * rng.choice([1e-20, 1e-3, 1.0, 1e3, 1e15], size=(n, 1))).astype(np.float32)
    filt = np.zeros((len(M), n), np.uint8); exact = np.zeros((len(M), n), np.uint8)
    hm.hm_f_filter(p(M), len(M), p(af2), p(bf2), n, C.c_float(9.0), p(filt), p(exact))
    assert not (filt & exact).any()


def test_e_filter_is_conservative(hm):
    """e_certain_outlier may only fire where e_inlier is false (5-point models of random and of inlier-only samples,
    rescaled / degenerate / non-finite models, tiny and huge coordinates) and removes most of the work."""
    rng = np.random.default_rng(11)
    a, b, gt, R, t = synth.two_view_pair(3, 5000)
    x1 = np.ascontiguousarray(essential.normalize_points(a, 800.0, (320.0, 240.0)))
    x2 = np.ascontiguousarray(essential.normalize_points(b, 800.0, (320.0, 240.0)))
    models = []
    Es = np.zeros(90)
    inl = np.flatnonzero(gt)
    for it in range(160):
        idx = rng.permutation(len(x1))[:5] if it < 120 else rng.choice(inl, 5, replace=False)
        n = hm.hm_e_solve5(p(np.ascontiguousarray(x1[idx])), p(np.ascontiguousarray(x2[idx])), p(Es))
        for k in range(n):
            models.append(Es[9 * k:9 * k + 9].copy())
    n_random = len(models)
    base = models[0]
    for sc in (1e-40, 1e-33, 1e-20, 1e-6, 1e6, 1e20, 1e29, 1e35, 1e300):
        models.append(base * sc)
    z = base.copy(); z[:6] = 0; models.append(z)
    z = base.copy(); z[[0, 1, 3, 4]] = 0; models.append(z)
    z = base.copy(); z[4] = np.nan; models.append(z)
    z = base.copy(); z[8] = -np.inf; models.append(z)
    models.append(np.zeros(9))
    M = np.ascontiguousarray(np.stack(models))
    n = len(x1)
    for thr in ((1.0 / 800) ** 2, 0.005 ** 2, 1e-12):
        filt = np.zeros((len(M), n), np.uint8); exact = np.zeros((len(M), n), np.uint8)
        hm.hm_e_filter(p(M), len(M), p(x1), p(x2), n, C.c_float(thr), p(filt), p(exact))
        assert not (filt & exact).any(), "filter rejected an inlier"
        if thr == (1.0 / 800) ** 2:
            wrong = exact[:n_random].sum(axis=1) < 0.05 * n
            assert filt[:n_random][wrong].mean() > 0.9
    sc = rng.choice([1e-20, 1e-3, 1.0, 1e3, 1e15], size=(n, 1))
    y1, y2 = np.ascontiguousarray(x1 * sc), np.ascontiguousarray(x2 * sc[::-1])
    filt = np.zeros((len(M), n), np.uint8); exact = np.zeros((len(M), n), np.uint8)
    hm.hm_e_filter(p(M), len(M), p(y1), p(y2), n, C.c_float((1.0 / 800) ** 2), p(filt), p(exact))
    assert not (filt & exact).any()


def test_pnp_filter_is_conservative(hm):
    """pnp_certain_outlier may only fire where the exact reprojection test says outlier: true pose, perturbed poses
    (decisions near the threshold), random poses, points behind / on the camera plane, extreme scalings."""
    rng = np.random.default_rng(5)
    K = np.array([[800.0, 0, 320.0], [0, 810.0, 240.0], [0, 0, 1.0]])
    X, uv, rv_gt, tv_gt = synth.pnp_scene(4, 3000, 0.5, noise=0.7, K=K)
    Xf, uf = np.ascontiguousarray(X, np.float32), np.ascontiguousarray(uv, np.float32)
    cam4 = np.array([K[0, 0], K[1, 1], K[0, 2], K[1, 2]])
    models = [np.concatenate([rv_gt, tv_gt])]
    for s in (1e-6, 1e-4, 1e-3, 1e-2, 0.1):
        for _ in range(20):
            models.append(models[0] + rng.normal(0, s, 6))
    for _ in range(100):
        models.append(np.concatenate([rng.normal(0, 1.0, 3), rng.normal(0, 3.0, 3)]))
    z = models[0].copy(); z[3:] *= 1e20; models.append(z)
    z = models[0].copy(); z[3:] *= 1e-20; models.append(z)
    z = models[0].copy(); z[3:] = 0; models.append(z)
    z = models[0].copy(); z[5] = -float(np.mean(X[:, 2])); models.append(z)     # camera plane through the scene
    z = models[0].copy(); z[2] = np.nan; models.append(z)
    z = models[0].copy(); z[4] = np.inf; models.append(z)
    M = np.ascontiguousarray(np.stack(models))
    n = len(Xf)
    for thr in (4.0, 64.0, 1e-3):
        filt = np.zeros((len(M), n), np.uint8); exact = np.zeros((len(M), n), np.uint8)
        hm.hm_pnp_filter(p(M), len(M), p(Xf), p(uf), n, p(cam4), C.c_float(thr), p(filt), p(exact))
        assert not (filt & exact).any(), "filter rejected an inlier"
        if thr == 4.0:
            assert exact[0].mean() > 0.3                                        # the true pose has its inliers
            assert filt[101:201].mean() > 0.95                                  # random poses: almost nothing survives
    # a point exactly on the camera plane (z = 0 is special-cased by the exact path) and huge coordinates
    Xb = Xf.copy(); Xb[::7] *= 1e12; Xb[1::7] *= 1e-12
    filt = np.zeros((len(M), n), np.uint8); exact = np.zeros((len(M), n), np.uint8)
    hm.hm_pnp_filter(p(M), len(M), p(Xb), p(uf), n, p(cam4), C.c_float(4.0), p(filt), p(exact))
    assert not (filt & exact).any()


def test_h_filter_is_conservative(hm):
    """h_certain_outlier may only fire where the exact float32 homography test says outlier."""
    rng = np.random.default_rng(3)
    a, b, gt, Hgt = synth.homography_pair(2, 4000)
    af, bf = np.ascontiguousarray(a, np.float32), np.ascontiguousarray(b, np.float32)
    models = [(Hgt / Hgt[2, 2]).ravel()]
    Hh = np.zeros(9)
    inl = np.flatnonzero(gt)
    for it in range(300):
        idx = rng.permutation(len(af))[:4] if it < 200 else rng.choice(inl, 4, replace=False)
        if hm.hm_h_solve4(p(np.ascontiguousarray(af[idx])), p(np.ascontiguousarray(bf[idx])), p(Hh)):
            models.append(Hh.copy())
    base = models[0]
    for s in (1e-7, 1e-5, 1e-3):
        for _ in range(20):
            z = base * (1 + rng.normal(0, s, 9)); z[8] = 1; models.append(z)
    for sc in (1e-35, 1e-20, 1e-6, 1e6, 1e20, 1e35):
        z = base * sc; z[8] = 1; models.append(z)
    z = base.copy(); z[6] = -1.0 / 500; z[7] = 0; models.append(z)      # line at infinity through the image (w = 0 at x = 500)
    z = base.copy(); z[3] = np.nan; models.append(z)
    z = base.copy(); z[7] = np.inf; models.append(z)
    models.append(np.array([0, 0, 0, 0, 0, 0, 0, 0, 1.0]))
    M = np.ascontiguousarray(np.stack(models))
    n = len(af)
    for thr in (9.0, 0.25, 1e-4):
        filt = np.zeros((len(M), n), np.uint8); exact = np.zeros((len(M), n), np.uint8)
        hm.hm_h_filter(p(M), len(M), p(af), p(bf), n, C.c_float(thr), p(filt), p(exact))
        assert not (filt & exact).any(), "filter rejected an inlier"
        if thr == 9.0:
            assert exact[0].mean() > 0.25
            wrong = exact[1:201].sum(axis=1) < 0.05 * n
            assert filt[1:201][wrong].mean() > 0.98
    af2 = np.ascontiguousarray(af.copy()); af2[::5, 0] = 500.0               # exactly on the w = 0 line of that model
    af3 = (af * rng.choice([1e-20, 1.0, 1e12], size=(n, 1))).astype(np.float32)
    for A_ in (af2, af3):
        filt = np.zeros((len(M), n), np.uint8); exact = np.zeros((len(M), n), np.uint8)
        hm.hm_h_filter(p(M), len(M), p(A_), p(bf), n, C.c_float(9.0), p(filt), p(exact))
        assert not (filt & exact).any()


# ---- level 0 of the tensor-core scoring engine (filter0.cuh): host emulation of score_tc.cu -------------------------------
def _filter0(hm, is_e, M, a, b, thr, mode=0, sorted_=1):
    M = np.ascontiguousarray(M, np.float64)
    a = np.ascontiguousarray(a, np.float64); b = np.ascontiguousarray(b, np.float64)
    n = len(a)
    filt = np.zeros((len(M), n), np.uint8); exact = np.zeros((len(M), n), np.uint8); cls = np.zeros((len(M), n), np.uint8)
    hm.hm_filter0(int(is_e), p(M), len(M), p(a), p(b), n, C.c_float(thr), mode, sorted_, p(filt), p(exact), p(cls))
    return filt, exact, cls


def _check_filter0(filt, exact, cls, msg=""):
    assert not (filt & exact).any(), f"level 0 rejected an inlier {msg}"
    assert not ((cls == 0) & (exact == 1)).any(), f"classifier called an inlier an outlier {msg}"
    assert not ((cls == 1) & (exact == 0)).any(), f"classifier called an outlier an inlier {msg}"


def _adversarial_models(base, rng):
    out = [base.copy()]
    for s in (1e-12, 1e-9, 1e-7, 1e-5, 1e-3, 1e-1):
        for _ in range(4):
            out.append(base * (1 + rng.normal(0, s, 9)))
    for sc in (1e-40, 1e-33, 1e-20, 1e-6, 1e6, 1e15, 1e20, 1e29, 1e35, 1e300):
        out.append(base * sc)
    z = base.copy(); z[:6] = 0; out.append(z)
    z = base.copy(); z[:6] *= 1e-25; out.append(z)
    z = base.copy(); z[6:] = 0; out.append(z)
    z = base.copy(); z[[0, 1, 3, 4]] = 0; out.append(z)
    z = base.copy(); z[4] = np.nan; out.append(z)
    z = base.copy(); z[8] = np.inf; out.append(z)
    out.append(np.zeros(9))
    out.append(rng.normal(0, 1, 9))
    return out


def test_f_level0_and_classifier_are_conservative(hm):
    """Level 0 (split-TF32 residual against a per-stage threshold) may only fire where f_inlier is false, and the two-sided
    fp32 classifier may only decide what f_inlier decides the same way -- random 7-point models, near-perfect ones
    (decisions at the threshold), rescaled / degenerate / non-finite ones, far-off-centre and badly scaled coordinates,
    NaN coordinates, three accumulation orders, sorted and unsorted stages -- and level 0 must remove most of the work."""
    rng = np.random.default_rng(17)
    a, b, gt, R, t = synth.two_view_pair(2, 3000)
    af, bf = np.ascontiguousarray(a, np.float32), np.ascontiguousarray(b, np.float32)
    models = []
    F7 = np.zeros(27)
    inl = np.flatnonzero(gt)
    for it in range(260):
        idx = rng.permutation(len(af))[:7] if it < 200 else rng.choice(inl, 7, replace=False)
        k = hm.hm_f_solve7(p(np.ascontiguousarray(af[idx])), p(np.ascontiguousarray(bf[idx])), p(F7))
        for j in range(k):
            models.append(F7[9 * j:9 * j + 9].copy())
    n_random = len(models)
    g = np.load(GOLDEN / "ransac_hf.npz")
    models += _adversarial_models(g["f2_F"].ravel().copy(), rng)
    M = np.stack(models)
    for thr in (9.0, 1.0, 1e-4, 400.0):
        for mode in (0, 1, 2):
            filt, exact, cls = _filter0(hm, 0, M, af, bf, thr, mode)
            _check_filter0(filt, exact, cls, f"thr {thr} mode {mode}")
        if thr == 9.0:
            wrong = exact[:n_random].sum(axis=1) < 0.05 * len(af)
            assert filt[:n_random][wrong].mean() > 0.94                       # ~3 % survive
            assert (cls[:n_random] == 2).mean() < 2e-3                        # the exact test is rare
            assert ((cls[:n_random] == 1) == (exact[:n_random] == 1)).mean() > 0.998
    filt, exact, cls = _filter0(hm, 0, M, af, bf, 9.0, 0, sorted_=0)          # boxes as large as the image
    _check_filter0(filt, exact, cls, "unsorted")
    # far-off-centre pairs (the exact test's own fp64 noise grows), badly scaled rows, a NaN coordinate
    for shift in (1e4, 1e7, 1e10, 1e14):
        filt, exact, cls = _filter0(hm, 0, M, af.astype(np.float64) + shift, bf.astype(np.float64) - 0.5 * shift, 9.0)
        _check_filter0(filt, exact, cls, f"shift {shift}")
    sc = rng.choice([1e-20, 1e-3, 1.0, 1e3, 1e15], size=(len(af), 1))
    filt, exact, cls = _filter0(hm, 0, M, af * sc, bf * sc[::-1], 9.0)
    _check_filter0(filt, exact, cls, "scaled rows")
    a3 = af.astype(np.float64).copy(); a3[5, 0] = np.nan
    filt, exact, cls = _filter0(hm, 0, M, a3, bf, 9.0)
    _check_filter0(filt, exact, cls, "nan")
    assert not filt.any()                                                      # level 0 is off for the pair
    filt, exact, cls = _filter0(hm, 0, M, af[:9], bf[:9], 9.0)
    _check_filter0(filt, exact, cls, "tiny")


def test_e_level0_and_classifier_are_conservative(hm):
    """The same for e_inlier (Sampson error, fp64 normalised coordinates)."""
    rng = np.random.default_rng(19)
    a, b, gt, R, t = synth.two_view_pair(3, 5000)
    x1 = np.ascontiguousarray(essential.normalize_points(a, 800.0, (320.0, 240.0)))
    x2 = np.ascontiguousarray(essential.normalize_points(b, 800.0, (320.0, 240.0)))
    models = []
    Es = np.zeros(90)
    inl = np.flatnonzero(gt)
    for it in range(160):
        idx = rng.permutation(len(x1))[:5] if it < 120 else rng.choice(inl, 5, replace=False)
        k = hm.hm_e_solve5(p(np.ascontiguousarray(x1[idx])), p(np.ascontiguousarray(x2[idx])), p(Es))
        for j in range(k):
            models.append(Es[9 * j:9 * j + 9].copy())
    n_random = len(models)
    models += _adversarial_models(models[0], rng)
    M = np.stack(models)
    for thr in ((1.0 / 800) ** 2, 0.005 ** 2, 1e-12):
        for mode in (0, 1, 2):
            filt, exact, cls = _filter0(hm, 1, M, x1, x2, thr, mode)
            _check_filter0(filt, exact, cls, f"thr {thr} mode {mode}")
        if thr == (1.0 / 800) ** 2:
            wrong = exact[:n_random].sum(axis=1) < 0.05 * len(x1)
            assert filt[:n_random][wrong].mean() > 0.95
            assert (cls[:n_random] == 2).mean() < 2e-3
    filt, exact, cls = _filter0(hm, 1, M, x1, x2, (1.0 / 800) ** 2, 0, sorted_=0)
    _check_filter0(filt, exact, cls, "unsorted")
    for shift in (1e3, 1e6, 1e9):
        filt, exact, cls = _filter0(hm, 1, M, x1 + shift, x2 - 0.5 * shift, (1.0 / 800) ** 2)
        _check_filter0(filt, exact, cls, f"shift {shift}")
    sc = rng.choice([1e-20, 1e-3, 1.0, 1e3, 1e15], size=(len(x1), 1))
    filt, exact, cls = _filter0(hm, 1, M, x1 * sc, x2 * sc[::-1], (1.0 / 800) ** 2)
    _check_filter0(filt, exact, cls, "scaled rows")


def test_tf32_split_is_what_the_kernel_assumes(hm):
    """hi + lo reproduces a float to 2^-22 relative and both pieces are TF32 values (13 low mantissa bits clear)."""
    rng = np.random.default_rng(23)
    x = np.concatenate([rng.normal(0, 1, 4000) * 10.0 ** rng.integers(-30, 30, 4000), [0.0, 1.0, -1.0, 3e38, 1e-38]]).astype(np.float32)
    hi = np.zeros_like(x); lo = np.zeros_like(x)
    hm.hm_tf32_split(p(x), len(x), p(hi), p(lo))
    assert not (hi.view(np.uint32) & 0x1FFF).any() and not (lo.view(np.uint32) & 0x1FFF).any()
    err = np.abs(x.astype(np.float64) - hi.astype(np.float64) - lo.astype(np.float64))
    assert (err <= 2.0 ** -22 * np.abs(x.astype(np.float64)) + 2e-41).all()      # + the subnormal grid


def test_h_level0_is_conservative(hm):
    """Homography: one accumulator carries the diagonal component of the reprojection residual; level 0 may only fire where
    the float32 exact test says outlier (random 4-point models, near-perfect ones, rescaled / degenerate / non-finite ones,
    off-centre and badly scaled coordinates) and must remove most of the work."""
    rng = np.random.default_rng(29)
    a, b, gt, H = synth.homography_pair(4, 3000)
    af, bf = np.ascontiguousarray(a, np.float32), np.ascontiguousarray(b, np.float32)
    models = []
    Hh = np.zeros(9)
    inl = np.flatnonzero(gt)
    for it in range(400):
        idx = rng.permutation(len(af))[:4] if it < 300 else rng.choice(inl, 4, replace=False)
        if hm.hm_h_solve4(p(np.ascontiguousarray(af[idx])), p(np.ascontiguousarray(bf[idx])), p(Hh)):
            models.append(Hh.copy())
    n_random = len(models)
    base = (H / H[2, 2]).ravel()
    for m in _adversarial_models(base, rng):
        m = m.copy(); m[8] = 1.0
        models.append(m)
    M = np.stack(models)
    for thr in (9.0, 1.0, 1e-4, 400.0):
        for mode in (0, 1, 2):
            filt, exact, cls = _filter0(hm, 2, M, af, bf, thr, mode)
            assert not (filt & exact).any(), f"level 0 rejected an inlier, thr {thr} mode {mode}"
        if thr == 9.0:
            wrong = exact[:n_random].sum(axis=1) < 0.05 * len(af)
            assert wrong.sum() > 100 and filt[:n_random][wrong].mean() > 0.95
    filt, exact, cls = _filter0(hm, 2, M, af, bf, 9.0, 0, sorted_=0)
    assert not (filt & exact).any()
    for shift in (1e4, 1e6, 1e9):
        filt, exact, cls = _filter0(hm, 2, M, af.astype(np.float64) + shift, bf.astype(np.float64) - 0.5 * shift, 9.0)
        assert not (filt & exact).any(), f"shift {shift}"
    sc = rng.choice([1e-20, 1e-3, 1.0, 1e3, 1e15], size=(len(af), 1))
    filt, exact, cls = _filter0(hm, 2, M, af * sc, bf * sc[::-1], 9.0)
    assert not (filt & exact).any()
    a3 = af.astype(np.float64).copy(); a3[7, 1] = np.nan
    filt, exact, cls = _filter0(hm, 2, M, a3, bf, 9.0)
    assert not filt.any()
