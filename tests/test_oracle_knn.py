"""CPU: the kNN oracle is pinned to the cv2 4.13.0 golden vectors (incl. engineered ties)."""
import numpy as np
import pytest

from oracle import bf_knn
from conftest import GOLDEN


@pytest.fixture(scope="module")
def g():
    return np.load(GOLDEN / "knn_hamming.npz")


@pytest.mark.parametrize("name", ["b256", "b488", "b512", "one"])
def test_oracle_matches_cv2_golden(g, name):
    idx, dist = bf_knn.bf_knn2_hamming(g[f"{name}_q"], g[f"{name}_t"])
    np.testing.assert_array_equal(idx, g[f"{name}_idx"])
    np.testing.assert_array_equal(dist, g[f"{name}_dist"])


def test_golden_contains_ties(g):
    d = g["b256_dist"]
    assert (d[:, 0] == d[:, 1]).sum() > 10          # d1 == d2 ties are exercised
    assert tuple(g["b256_idx"][0]) == (3, 5)        # 4-way duplicate -> two lowest indices


def test_ratio_is_float32():
    dist = np.array([[80.0, 100.0], [79.0, 100.0], [4.0, 5.0]], np.float32)
    idx = np.array([[0, 1], [0, 1], [0, 1]], np.int32)
    # 0.8f * 100 = 80.00000119..., rounded to fp32 -> 80.0 exactly, so 80 < 80 fails
    assert bf_knn.ratio_test(dist, idx, 0.8).tolist() == [0, 1, 0]


def test_merge_keys_prefers_low_global_index():
    k = np.array([[[(5 << 32) | 900, (7 << 32) | 901]], [[(5 << 32) | 12, (6 << 32) | 13]]], np.uint64)
    m = bf_knn.merge_top2_keys(k)
    assert m[0, 0] == (5 << 32) | 12 and m[0, 1] == (5 << 32) | 900


def test_empty_inputs():
    idx, dist = bf_knn.bf_knn2_hamming(np.zeros((3, 32), np.uint8), np.zeros((0, 32), np.uint8))
    assert (idx == -1).all() and (dist == bf_knn.FLT_MAX).all()
    idx, dist = bf_knn.bf_knn2_hamming(np.zeros((0, 32), np.uint8), np.zeros((5, 32), np.uint8))
    assert idx.shape == (0, 2)


def test_l2_oracle_vs_cv2_golden():
    g = np.load(GOLDEN / "knn_l2.npz")
    idx, dist = bf_knn.bf_knn2_l2(g["sift_q"], g["sift_t"])          # integer-valued: bit-exact
    np.testing.assert_array_equal(idx, g["sift_idx"])
    np.testing.assert_array_equal(dist, g["sift_dist"])
    assert tuple(idx[0]) == (3, 9)
    idx, dist = bf_knn.bf_knn2_l2(g["gen_q"], g["gen_t"])            # general floats: also bit-exact (cv2's order)
    np.testing.assert_array_equal(idx, g["gen_idx"])
    np.testing.assert_array_equal(dist, g["gen_dist"])


@pytest.mark.parametrize("name", ["g128", "g64", "g61", "g7"])
def test_l2_oracle_follows_cv2_accumulation_order(name):
    """hal::normL2Sqr_'s 16 strided partial sums + sequential tail, restated in numpy, equals cv2 4.13.0 bit for bit
    on non-integer descriptors; the plain index-order sum does not."""
    g = np.load(GOLDEN / "knn_l2_general.npz")
    q, t = g[f"{name}_q"], g[f"{name}_t"]
    idx, dist = bf_knn.bf_knn2_l2(q, t)
    np.testing.assert_array_equal(idx, g[f"{name}_idx"])
    np.testing.assert_array_equal(dist, g[f"{name}_dist"])
    if q.shape[1] >= 64:
        seq = bf_knn.l2_matrix_seq(q, t)
        got = np.take_along_axis(seq, g[f"{name}_idx"].astype(np.int64), axis=1)
        assert (got != g[f"{name}_dist"]).any()
