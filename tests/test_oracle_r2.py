"""CPU: round-2 oracle pieces pinned to cv2 4.13.0 goldens (tests/golden/make_golden_r2.py):
cv::findFundamentalMat's small-N dispatch (7-point direct, LMedS below 15 points, parameter clamps) and the
C4-size essential-matrix path (N = 5000, both configurations of SURVEY.md section 8d)."""
import numpy as np
import pytest

from oracle import essential, ransac
from conftest import GOLDEN


def rel(A, B):
    return np.abs(np.asarray(A) - np.asarray(B)).max() / np.abs(B).max()


@pytest.fixture(scope="module")
def gs():
    return np.load(GOLDEN / "ransac_f_small.npz")


def _cases(gs):
    return [(k, gs[f"s{k}_a"].shape[0]) for k in range(int(gs["s_n"]))]


def test_fundamental_small_n_dispatch(gs):
    """N == 7: every 7-point solution, stacked, mask of ones.  N >= 14: the median sits on a genuine residual and the
    oracle equals cv2 exactly.  8 <= N <= 13: the median (element N/2 of the sorted errors) is one of the seven sample
    points, i.e. pure rounding noise (~1e-27) -- cv2's pick among the exact-fit hypotheses is not reproducible by any
    other implementation of the same arithmetic; what is checked is that cv2's F is one of OUR hypotheses and that our
    own pick obeys the LMedS rules (noise-level median, sigma-rule mask)."""
    seen = set()
    for k, n in _cases(gs):
        a, b = gs[f"s{k}_a"], gs[f"s{k}_b"]
        nF = int(gs[f"s{k}_nF"])
        Fg, mg = gs[f"s{k}_F"][:3 * nF], gs[f"s{k}_mask"]
        F, mask = ransac.cv_find_fundamental(a, b, 3.0, 0.99, 1000)
        seen.add(n)
        if n == 7:
            # the solutions as a set: their order follows the (noise-decided) basis OpenCV's SVD picks for the exact
            # two-dimensional null space of the 7 x 9 system
            assert F.shape == Fg.shape
            for i in range(nF):
                assert min(rel(F[3 * j:3 * j + 3], Fg[3 * i:3 * i + 3]) for j in range(nF)) < 1e-7
            assert mask.tolist() == [1] * 7 and mg.tolist() == [1] * 7
        elif n >= 14:
            assert rel(F, Fg) < 1e-9
            np.testing.assert_array_equal(mask, mg)
        else:
            m1, m2 = ransac.FundamentalCB.prepare(a, b)
            tr = ransac.RansacTrace()
            ok, Fo, mo = ransac.lmeds_run(ransac.FundamentalCB, m1, m2, 0.99, 1000, trace=tr)
            assert ok and len(tr.samples) == 300          # niters = update(0.99, 0.45, 7, 1000)
            allm = [m for ms in tr.models for m in ms]
            assert min(rel(m, Fg) for m in allm) < 1e-7   # cv2's winner is among the hypotheses of the same stream
            err = np.sort(np.asarray(ransac.FundamentalCB.compute_error(m1, m2, Fo), np.float32))
            assert err[n // 2] < 1e-18                    # an exact-fit hypothesis
            errg = np.sort(np.asarray(ransac.FundamentalCB.compute_error(m1, m2, Fg), np.float32))
            assert errg[n // 2] < 1e-18
    assert seen == set(range(7, 16))


def test_fundamental_parameter_clamps(gs):
    a, b = gs["c_a"], gs["c_b"]
    for j in range(int(gs["c_n"])):
        thr, conf = gs[f"c{j}_prm"]
        F, mask = ransac.cv_find_fundamental(a, b, float(thr), float(conf), 500)
        assert rel(F, gs[f"c{j}_F"]) < 1e-9
        np.testing.assert_array_equal(mask, gs[f"c{j}_mask"])


@pytest.mark.parametrize("k", [0, 1])
def test_c4_size_recover_pose_oracle(k):
    """5000 correspondences (config C4): E-RANSAC mask, cheirality count, R, t vs cv2."""
    g = np.load(GOLDEN / "ransac_e5k.npz")
    a, b = g[f"e{k}_a"], g[f"e{k}_b"]
    f, px, py, prob, thr = g[f"e{k}_cfg"]
    cnt, R, t, mask = essential.cv_recover_pose((f, (px, py), prob, thr), a, b)
    np.testing.assert_array_equal(mask, g[f"e{k}_mask"])
    assert cnt == int(g[f"e{k}_cnt"])
    assert rel(R, g[f"e{k}_R"]) < 1e-6 and rel(t, g[f"e{k}_t"]) < 1e-6
