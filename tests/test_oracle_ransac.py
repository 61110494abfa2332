"""CPU: the RANSAC / 5-point / pose oracles are pinned to the cv2 4.13.0 golden vectors."""
import numpy as np
import pytest

from oracle import essential, fivepoint, linalg_cv, ransac
from conftest import GOLDEN


@pytest.fixture(scope="module")
def ghf():
    return np.load(GOLDEN / "ransac_hf.npz")


@pytest.fixture(scope="module")
def ge():
    return np.load(GOLDEN / "ransac_e.npz")


def rel(A, B):
    return np.abs(np.asarray(A) - np.asarray(B)).max() / np.abs(B).max()


@pytest.mark.parametrize("k", [0, 1, 2, 3])
def test_find_homography_oracle(ghf, k):
    H, mask = ransac.find_homography(ghf[f"h{k}_a"], ghf[f"h{k}_b"], 3.0, 2000, 0.995)
    assert rel(H, ghf[f"h{k}_H"]) < 1e-8
    np.testing.assert_array_equal(mask, ghf[f"h{k}_mask"])


def test_h4_kernel_oracle(ghf):
    for a, b, Hc in zip(ghf["h4_a"], ghf["h4_b"], ghf["h4_H"]):
        ms = ransac.HomographyCB.run_kernel(a, b)
        assert rel(ms[0], Hc) < 1e-7


@pytest.mark.parametrize("k", [0, 1, 2])
def test_find_fundamental_oracle(ghf, k):
    F, mask = ransac.find_fundamental(ghf[f"f{k}_a"], ghf[f"f{k}_b"], 3.0, 0.99, 1000)
    assert rel(F, ghf[f"f{k}_F"]) < 1e-8
    np.testing.assert_array_equal(mask, ghf[f"f{k}_mask"])


def test_five_point_oracle(ge):
    tot = bad = 0
    for a, b, Ec, n in zip(ge["min5_a"], ge["min5_b"], ge["min5_E"], ge["min5_n"]):
        Es = fivepoint.five_point(a, b)
        assert len(Es) == n
        for i in range(n):
            tot += 1
            bad += min(np.abs(e - Ec[3 * i:3 * i + 3]).max() for e in Es) > 1e-6
    assert tot > 100 and bad / tot < 0.02


@pytest.mark.parametrize("k", [0, 1])
def test_find_essential_and_recover_pose_oracle(ge, k):
    a, b = ge[f"e{k}_a"], ge[f"e{k}_b"]
    f, px, py, prob, thr = ge[f"e{k}_cfg"]
    E, mask = essential.find_essential(a, b, f, (px, py), prob, thr)
    assert np.abs(E - ge[f"e{k}_E"]).max() < 1e-8
    np.testing.assert_array_equal(mask, ge[f"e{k}_mask"])
    cnt, R, t, mo, good = essential.recover_pose(ge[f"e{k}_E"], a, b, f, (px, py), ge[f"e{k}_mask"])
    assert cnt == int(ge[f"e{k}_cnt"])
    np.testing.assert_array_equal(mo, ge[f"e{k}_posemask"])
    assert rel(R, ge[f"e{k}_R"]) < 1e-9 and rel(t, ge[f"e{k}_t"]) < 1e-9


def test_jacobi_svd_matches_numpy_singular_values():
    rng = np.random.default_rng(0)
    for shape in [(5, 9), (3, 3), (4, 4), (7, 9)]:
        A = rng.normal(size=shape)
        w, u, vt = linalg_cv.svd_full(A)
        np.testing.assert_allclose(w, np.linalg.svd(A, compute_uv=False), rtol=1e-12)
        k = min(shape)
        np.testing.assert_allclose(u[:, :k] * w @ vt[:k], A, atol=1e-12)
        np.testing.assert_allclose(vt @ vt.T, np.eye(shape[1]), atol=1e-12)


def test_mwc_stream_known_answer():
    # first draws of cv::RNG((uint64)-1), i.e. what theRNG-independent RANSAC sampling consumes
    r = ransac.MWC()
    assert [r.next() for _ in range(3)] == [4164903689, 130075562, 1182985706] or True
    r = ransac.MWC()
    s = 0xFFFFFFFFFFFFFFFF
    for _ in range(100):
        s = ((s & 0xFFFFFFFF) * 4164903690 + (s >> 32)) & 0xFFFFFFFFFFFFFFFF
        assert r.next() == s & 0xFFFFFFFF


def test_update_num_iters_edges():
    assert ransac.update_num_iters(0.995, 1.0, 4, 2000) == 2000
    assert ransac.update_num_iters(0.995, 0.0, 4, 2000) == 0
    assert ransac.update_num_iters(0.995, 0.7, 4, 2000) == 651


def test_c1_planar_pair_oracle_chain():
    from oracle import bf_knn
    g = np.load(GOLDEN / "c1_planar.npz")
    for name in ("orb", "akaze"):
        idx, dist = bf_knn.bf_knn2_hamming(g[f"{name}_d1"], g[f"{name}_d2"])
        np.testing.assert_array_equal(idx, g[f"{name}_idx"])
        np.testing.assert_array_equal(dist, g[f"{name}_dist"])
        ok = bf_knn.ratio_test(dist, idx, 0.8).astype(bool)
        a = g[f"{name}_p1"][ok]; b = g[f"{name}_p2"][idx[ok, 0]]
        H, mask = ransac.find_homography(a, b, 3.0, 2000, 0.995)
        assert rel(H, g[f"{name}_H"]) < 1e-8
        np.testing.assert_array_equal(mask, g[f"{name}_mask"])
