"""GPU: the error-bounded fp32 filters in front of the exact inlier tests (csrc/cvmath.cuh, pnpmath.cuh, filter_pass.cuh)
must not change a single count.  The host suite checks the scalar filter formulas for conservativeness
(tests/test_hostmath.py::test_*_filter_is_conservative); here the DEVICE paths -- packed f32x2 filters, per-tile bounds from
the kernels' own maxima, cooperative / plain-loop resolution of the survivors -- are driven with adversarial injected
models and coordinates, and every inlier count is compared with the oracle's exact test."""
import numpy as np
import pytest

from minicv_b200 import estimation as est
from minicv_b200 import synth
from oracle import essential, pnp, ransac

pytestmark = pytest.mark.gpu


def adversarial(base, rng, n_random=40, scales=(1e-40, 1e-33, 1e-20, 1e-6, 1e6, 1e20, 1e29, 1e35), fix33=False):
    """The true model, perturbations down to the last digits (decisions at the threshold), random models, rescalings
    over the exponent range, degenerate and non-finite models."""
    b = np.asarray(base, float).ravel()
    out = [b.copy()]
    for s in (1e-12, 1e-9, 1e-7, 1e-5, 1e-3, 1e-1):
        for _ in range(6):
            out.append(b * (1 + rng.normal(0, s, b.size)))
    for _ in range(n_random):
        out.append(rng.normal(0, 1, b.size) * np.abs(b).max())
    for sc in scales:
        out.append(b * sc)
    z = b.copy(); z[:6] = 0; out.append(z)
    z = b.copy(); z[6:] = 0; out.append(z)
    z = b.copy(); z[[0, 1, 3, 4]] = 0; out.append(z)
    z = b.copy(); z[4] = np.nan; out.append(z)
    z = b.copy(); z[2] = np.inf; out.append(z)
    z = b.copy(); z[7] = -np.inf; out.append(z)
    out.append(np.zeros(b.size))
    M = np.array(out)
    if fix33:
        M[:, 8] = 1.0
    return M


def point_variants(a, b, rng):
    """The pair as it is; all inliers of the true model only (every model near it fits: plain-loop mode); coordinates
    scaled far out of range on a few rows; a NaN coordinate in one tile."""
    yield "plain", a, b
    yield "odd count (pair-interleaved tile with a half-filled last pair)", a[:-1], b[:-1]
    yield "tiny", a[:9], b[:9]
    a2, b2 = a.copy(), b.copy()
    rows = rng.permutation(len(a))[:40]
    a2[rows[:20]] *= 1e7; b2[rows[20:]] *= 1e-7
    yield "scaled rows", a2, b2
    a3 = a.copy(); a3[5, 0] = np.nan
    yield "nan", a3, b


K_TWO_VIEW = np.array([[800.0, 0, 320.0], [0, 800.0, 240.0], [0, 0, 1.0]])


def gt_essential(R, t):
    """x2' E x1 = 0 for x2 = R x1 + t (normalised coordinates): E = [t]x R."""
    tx = np.array([[0, -t[2], t[1]], [t[2], 0, -t[0]], [-t[1], t[0], 0]])
    return tx @ R


def counts_device(model, a, b, M, thr, **kw):
    r = est.ransac_batch(model, a, b, [0, len(a)], thr, len(M), 0.99, stop_mode=est.STOP_FIXED,
                         models=M.reshape(1, len(M), 3, 3), refine=False, debug=True, **kw)
    return r.dbg_counts[0, :, 0]


def test_homography_filter_never_changes_a_count():
    rng = np.random.default_rng(1)
    a, b, gt, H = synth.homography_pair(21, 3000)
    M = adversarial(H / H[2, 2], rng, fix33=True)
    with np.errstate(all="ignore"):
        for name, aa, bb in point_variants(a, b, rng):
            m1, m2 = ransac.HomographyCB.prepare(aa, bb)
            want = np.array([ransac.find_inliers(ransac.HomographyCB, m1, m2, m.reshape(3, 3), 3.0)[0] for m in M])
            np.testing.assert_array_equal(counts_device(est.MODEL_H, aa, bb, M, 3.0), want, err_msg=name)
        # all-inlier input: every near-true model keeps everything (dense lanes -> plain-loop mode)
        inl = gt.astype(bool)
        m1, m2 = ransac.HomographyCB.prepare(a[inl], b[inl])
        want = np.array([ransac.find_inliers(ransac.HomographyCB, m1, m2, m.reshape(3, 3), 3.0)[0] for m in M])
        np.testing.assert_array_equal(counts_device(est.MODEL_H, a[inl], b[inl], M, 3.0), want)
        assert want[0] == inl.sum()


def test_fundamental_filter_never_changes_a_count():
    rng = np.random.default_rng(2)
    a, b, gt, R, t = synth.two_view_pair(22, 3000)
    Ki = np.linalg.inv(K_TWO_VIEW)
    F = Ki.T @ gt_essential(R, t) @ Ki
    M = adversarial(F / F[2, 2], rng)
    with np.errstate(all="ignore"):
        for name, aa, bb in point_variants(a, b, rng):
            m1, m2 = ransac.FundamentalCB.prepare(aa, bb)
            want = np.array([ransac.find_inliers(ransac.FundamentalCB, m1, m2, m.reshape(3, 3), 3.0)[0] for m in M])
            np.testing.assert_array_equal(counts_device(est.MODEL_F, aa, bb, M, 3.0), want, err_msg=name)
        inl = gt.astype(bool)
        m1, m2 = ransac.FundamentalCB.prepare(a[inl], b[inl])
        want = np.array([ransac.find_inliers(ransac.FundamentalCB, m1, m2, m.reshape(3, 3), 3.0)[0] for m in M])
        np.testing.assert_array_equal(counts_device(est.MODEL_F, a[inl], b[inl], M, 3.0), want)
        assert want[0] > 0.9 * inl.sum()


def test_essential_filter_never_changes_a_count():
    rng = np.random.default_rng(3)
    a, b, gt, R, t = synth.two_view_pair(23, 3000)
    E = gt_essential(R, t)
    M = adversarial(E / np.linalg.norm(E), rng)
    thr_n = 1.0 / 800.0
    with np.errstate(all="ignore"):
        for name, aa, bb in point_variants(a, b, rng):
            m1 = essential.normalize_points(aa, 800.0, (320.0, 240.0))
            m2 = essential.normalize_points(bb, 800.0, (320.0, 240.0))
            want = np.array([ransac.find_inliers(essential.EssentialCB, m1, m2, m.reshape(3, 3), thr_n)[0] for m in M])
            got = counts_device(est.MODEL_E, aa, bb, M, 1.0, focal=800.0, pp=(320.0, 240.0))
            np.testing.assert_array_equal(got, want, err_msg=name)
        inl = gt.astype(bool)
        m1 = essential.normalize_points(a[inl], 800.0, (320.0, 240.0))
        m2 = essential.normalize_points(b[inl], 800.0, (320.0, 240.0))
        want = np.array([ransac.find_inliers(essential.EssentialCB, m1, m2, m.reshape(3, 3), thr_n)[0] for m in M])
        got = counts_device(est.MODEL_E, a[inl], b[inl], M, 1.0, focal=800.0, pp=(320.0, 240.0))
        np.testing.assert_array_equal(got, want)
        assert want[0] > 0.9 * inl.sum()


def test_pnp_filter_never_changes_a_count():
    rng = np.random.default_rng(4)
    K = np.array([[800.0, 0, 320.0], [0, 810.0, 240.0], [0, 0, 1.0]])
    X, uv, rvec, tvec = synth.pnp_scene(24, 3000, 0.5, noise=0.7, K=K)
    base = np.concatenate([rvec, tvec])
    out = [base.copy()]
    for s in (1e-12, 1e-9, 1e-7, 1e-5, 1e-3, 1e-1):
        for _ in range(8):
            out.append(base + rng.normal(0, s, 6))
    for _ in range(60):
        out.append(np.concatenate([rng.normal(0, 1.0, 3), rng.normal(0, 3.0, 3)]))
    z = base.copy(); z[3:] *= 1e20; out.append(z)
    z = base.copy(); z[3:] *= 1e-20; out.append(z)
    z = base.copy(); z[3:] = 0; out.append(z)
    z = base.copy(); z[5] = -float(np.mean(X[:, 2])); out.append(z)      # camera plane through the scene
    z = base.copy(); z[1] = np.nan; out.append(z)
    z = base.copy(); z[4] = np.inf; out.append(z)
    M = np.array(out)
    cb = pnp.PnPCB(K, np.zeros(4))
    thr2 = np.float32(4.0)
    X2 = X.copy(); X2[::9] *= 1e9; X2[1::9] *= 1e-9
    with np.errstate(all="ignore"):
        for name, XX, sel in (("plain", X, slice(None)), ("scaled rows", X2, slice(None)),
                              ("inliers only", X, slice(len(X) // 2, None))):
            o32, i32 = XX[sel].astype(np.float32), uv[sel].astype(np.float32)
            want = np.array([int((cb.compute_error(o32, i32, m.reshape(2, 3).T) <= thr2).sum()) for m in M])
            r = est.solve_pnp_ransac_batch(XX[sel], uv[sel], [0, len(o32)], est.camera_rows(K, np.zeros(4)), thr=2.0,
                                           maxIters=len(M), stop_mode=est.STOP_FIXED, inj_models=M[None], debug=True)
            np.testing.assert_array_equal(r.dbg_counts[0], want, err_msg=name)
    assert want[0] > 0.9 * (len(X) - len(X) // 2)
