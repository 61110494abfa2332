"""GPU parity for the round-2 holes: cvRecoverPose at C4 size against cv2 goldens, and
cvFindFundamentalMat's small-N dispatch (7-point direct, LMedS for 8..14 points, parameter clamps) against cv2
goldens and the oracle (tests/golden/make_golden_r2.py, oracle/ransac.py cv_find_fundamental / lmeds_run)."""
import numpy as np
import pytest

from minicv_b200 import estimation as est, native
from oracle import essential, ransac
from conftest import GOLDEN

pytestmark = pytest.mark.gpu

REL = 1e-6


def rel(A, B):
    return np.abs(np.asarray(A, float) - np.asarray(B, float)).max() / max(np.abs(B).max(), 1e-300)


def rel_upto_sign(A, B):
    return min(rel(A, B), rel(-np.asarray(A), B))


@pytest.mark.parametrize("k", [0, 1])
def test_recover_pose_c4_size_vs_cv2_golden(k):
    """N = 5000 (config C4) in pixel units (f 800, thr 1.0) and in MiniCV's default normalised form (f 1, thr 0.005):
    RANSAC mask >= 99.9 % identical (it is identical), cheirality count within 0.1 %, R / t within 1e-6."""
    g = np.load(GOLDEN / "ransac_e5k.npz")
    a, b = g[f"e{k}_a"], g[f"e{k}_b"]
    f, px, py, prob, thr = g[f"e{k}_cfg"]
    cfg = est.RecoverPoseConfig(f, (px, py), prob, thr)
    res, R, t, ms = est.recoverPose(cfg, a, b)
    assert (ms == g[f"e{k}_mask"]).mean() >= 0.999
    assert abs(res - int(g[f"e{k}_cnt"])) <= 5
    assert rel(R, g[f"e{k}_R"]) < REL and rel(t, g[f"e{k}_t"]) < REL
    m1, m2, tt, ms2 = est.recoverPoses(cfg, a, b)
    np.testing.assert_array_equal(ms, ms2)
    R1, R2, td = g[f"e{k}_R1"], g[f"e{k}_R2"], g[f"e{k}_tdec"]
    assert min(rel(m1, R1) + rel(m2, R2), rel(m1, R2) + rel(m2, R1)) < 2 * REL
    assert rel_upto_sign(tt, td) < REL
    # the batched export on the same pair twice (two pairs, one launch) gives the single-call answer
    A2 = np.concatenate([a, a]); B2 = np.concatenate([b, b])
    Rb, tb, msb, good, ok = est.recover_pose_batch(cfg, A2, B2, [0, len(a), 2 * len(a)])
    assert ok.tolist() == [1, 1] and good[0] == res and good[1] == res
    np.testing.assert_array_equal(msb[:len(a)], ms)
    np.testing.assert_array_equal(msb[len(a):], ms)
    # cheirality stage on cv2's own E: exact
    cnt, R3, t3, m3, g4 = est.recover_pose_from_E(g[f"e{k}_E"], a, b, f, (px, py), g[f"e{k}_mask"])
    assert cnt == int(g[f"e{k}_cnt"])
    np.testing.assert_array_equal(m3, g[f"e{k}_posemask"])


@pytest.fixture(scope="module")
def gs():
    return np.load(GOLDEN / "ransac_f_small.npz")


def test_fundamental_small_n_dispatch_vs_cv2(gs):
    seen = set()
    for k in range(int(gs["s_n"])):
        a, b = gs[f"s{k}_a"], gs[f"s{k}_b"]
        n = a.shape[0]
        nF = int(gs[f"s{k}_nF"])
        Fg, mg = gs[f"s{k}_F"][:3 * nF], gs[f"s{k}_mask"]
        F, mask = est.findFundamentalMat(a, b, 3.0, 0.99, 1000)
        assert F is not None, native.last_error()
        seen.add(n)
        if n == 7:
            # every solution of the 7-point kernel, as a set (the order follows a noise-decided null-space basis)
            assert F.shape == Fg.shape
            for i in range(nF):
                assert min(rel(F[3 * j:3 * j + 3], Fg[3 * i:3 * i + 3]) for j in range(nF)) < REL
            assert mask.tolist() == [1] * 7
        elif n >= 14:
            # LMedS with a genuine median (N = 14) and RANSAC (N = 15): model and mask equal cv2's
            assert F.shape == (3, 3) and rel(F, Fg) < REL
            np.testing.assert_array_equal(mask, mg)
        else:
            # 8 <= N <= 13: the LMedS median is a sample point's residual = rounding noise (see tests/test_oracle_r2.py);
            # the winner must be an exact-fit hypothesis of OpenCV's sampling stream and the mask the sigma rule's
            assert F.shape == (3, 3)
            m1, m2 = ransac.FundamentalCB.prepare(a, b)
            tr = ransac.RansacTrace()
            ransac.lmeds_run(ransac.FundamentalCB, m1, m2, 0.99, 1000, trace=tr)
            allm = [m for ms in tr.models for m in ms]
            assert min(rel(m, F) for m in allm) < REL
            err = np.asarray(ransac.FundamentalCB.compute_error(m1, m2, F), np.float32)
            med = np.sort(err)[n // 2]
            assert med < 1e-18
            sigma = max(2.5 * 1.4826 * (1 + 5.0 / (n - 7)) * np.sqrt(float(med)), 0.001)
            want = (err <= np.float32(sigma * sigma)).astype(np.uint8)
            np.testing.assert_array_equal(mask, want)
    assert seen == set(range(7, 16))


def test_fundamental_lmeds_batch_matches_oracle_at_14(gs):
    """MCV_STOP_LMEDS through the batched export: iterations = RANSACUpdateNumIters(conf, 0.45, 7, maxIters) = 300."""
    ks = [k for k in range(int(gs["s_n"])) if gs[f"s{k}_a"].shape[0] == 14]
    A = np.concatenate([gs[f"s{k}_a"] for k in ks]); B = np.concatenate([gs[f"s{k}_b"] for k in ks])
    off = np.arange(len(ks) + 1) * 14
    r = est.ransac_batch(est.MODEL_F, A, B, off, 3.0, 1000, 0.99, stop_mode=est.STOP_LMEDS)
    assert r.ok.tolist() == [1] * len(ks) and r.iters_run.tolist() == [300] * len(ks)
    for i, k in enumerate(ks):
        Fo, mo = ransac.cv_find_fundamental(gs[f"s{k}_a"], gs[f"s{k}_b"], 3.0, 0.99, 1000)
        assert rel(r.models[i], Fo) < REL
        np.testing.assert_array_equal(r.masks[off[i]:off[i + 1]], mo)
        assert r.inliers[i] == int(mo.sum())
    with pytest.raises(native.NativeError):
        est.ransac_batch(est.MODEL_H, A, B, off, 3.0, 1000, 0.99, stop_mode=est.STOP_LMEDS)


def test_fundamental_parameter_clamps_vs_cv2(gs):
    a, b = gs["c_a"], gs["c_b"]
    for j in range(int(gs["c_n"])):
        thr, conf = gs[f"c{j}_prm"]
        F, mask = est.findFundamentalMat(a, b, float(thr), float(conf), 500)
        assert F is not None and rel(F, gs[f"c{j}_F"]) < REL
        assert (mask == gs[f"c{j}_mask"]).mean() >= 0.999
    F, mask = est.findFundamentalMat(a[:6], b[:6], 3.0, 0.99, 500)
    assert F is None
