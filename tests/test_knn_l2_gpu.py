"""GPU parity of the L2 matcher: both engines return cv2.BFMatcher(NORM_L2)'s indices AND float32 distances bit
for bit -- on integer-valued (SIFT-like) and on general float descriptors -- because the kernels follow OpenCV's
accumulation order (oracle/bf_knn.py l2_matrix_cv, pinned to cv2 4.13.0 by tests/golden/knn_l2*.npz)."""
import numpy as np
import pytest

import minicv_b200 as mcv
from oracle import bf_knn
from conftest import GOLDEN

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True, params=[mcv.ENGINE_POPC, mcv.ENGINE_TENSOR], ids=["fp32-cuda-core", "tcgen05-bf16x2+rescore"])
def l2_engine(request):
    # mcvSetL2Engine: 1 = CUDA cores, 2 = tensor cores
    mcv.set_l2_engine(request.param)
    yield
    mcv.set_l2_engine(0)


def test_l2_golden_cv2():
    g = np.load(GOLDEN / "knn_l2.npz")
    r = mcv.match_knn2(g["sift_q"], g["sift_t"], 0.8, norm=mcv.NORM_L2)
    np.testing.assert_array_equal(r.train_idx, g["sift_idx"])
    np.testing.assert_array_equal(r.distance, g["sift_dist"])
    r = mcv.match_knn2(g["gen_q"], g["gen_t"], 0.8, norm=mcv.NORM_L2)
    np.testing.assert_array_equal(r.train_idx, g["gen_idx"])
    np.testing.assert_array_equal(r.distance, g["gen_dist"])


@pytest.mark.parametrize("name", ["g128", "g64", "g61", "g7"])
def test_l2_general_floats_bit_exact_vs_cv2(name):
    """Non-integer descriptors, dimensions with and without a 16-element tail: distances are cv2's float32 values."""
    g = np.load(GOLDEN / "knn_l2_general.npz")
    r = mcv.match_knn2(g[f"{name}_q"], g[f"{name}_t"], 0.8, norm=mcv.NORM_L2)
    np.testing.assert_array_equal(r.train_idx, g[f"{name}_idx"])
    np.testing.assert_array_equal(r.distance, g[f"{name}_dist"])


def test_l2_crowded_runner_up_is_certified_or_recomputed():
    """the tensor engine keeps 4 candidates per slot, ranked by an approximate score.  Here dozens of train
    rows sit within the approximation's noise of the runner-up (tiny perturbations of one row), so the candidate list
    cannot be trusted; the engine must notice and fall back to the exact kernel for those queries.  Both engines must
    still equal the oracle exactly, indices included (lowest index on exact ties)."""
    rng = np.random.default_rng(21)
    nt, nq, dim = 3000, 400, 128
    t = rng.normal(size=(nt, dim)).astype(np.float32)
    q = rng.normal(size=(nq, dim)).astype(np.float32)
    base = t[7].copy()
    for j in range(60):                                  # 60 rows within ~1e-6 relative of each other
        t[100 + 13 * j] = base + rng.normal(scale=2e-7, size=dim).astype(np.float32) * np.abs(base)
    t[2000] = t[100]; t[2500] = t[100]                   # and exact duplicates (index tie-break)
    q[:100] = base + rng.normal(scale=0.01, size=(100, dim)).astype(np.float32)
    qo = np.array([0, nq]); to = np.array([0, nt])
    r = mcv.match_knn2_batch(np.tile(q, (1, 1)), qo, t, to, 0.8, norm=mcv.NORM_L2)
    idx, dist = bf_knn.bf_knn2_l2(q, t)
    np.testing.assert_array_equal(r.train_idx, idx)
    np.testing.assert_array_equal(r.distance, dist)
    np.testing.assert_array_equal(r.ratio_pass, bf_knn.ratio_test(dist, idx, 0.8))
    assert r.n_pass == int(r.ratio_pass.sum())


@pytest.mark.parametrize("nq,nt,dim", [(1, 1, 128), (5, 3, 64), (130, 257, 128), (64, 64, 32), (300, 2000, 128), (33, 70, 61)])
def test_l2_vs_oracle_bit_exact(nq, nt, dim):
    rng = np.random.default_rng(nq * 31 + nt)
    q = rng.normal(size=(nq, dim)).astype(np.float32)
    t = rng.normal(size=(nt, dim)).astype(np.float32)
    r = mcv.match_knn2(q, t, 0.8, norm=mcv.NORM_L2)
    idx, dist = bf_knn.bf_knn2_l2(q, t)
    np.testing.assert_array_equal(r.train_idx, idx)
    np.testing.assert_array_equal(r.distance, dist)
    np.testing.assert_array_equal(r.ratio_pass, bf_knn.ratio_test(dist, idx, 0.8))


def test_l2_sift_like_batch_full_width():
    """SIFT-shaped batch (integer-valued 128-d, 3 pairs of 2000 x 3000): exact on both engines."""
    rng = np.random.default_rng(9)
    P, nq, nt = 3, 2000, 3000
    q = rng.integers(0, 160, (P * nq, 128)).astype(np.float32)
    t = rng.integers(0, 160, (P * nt, 128)).astype(np.float32)
    for p in range(P):   # plant near-duplicates
        src = rng.permutation(nt)[:500]
        q[p * nq:p * nq + 500] = t[p * nt + src] + rng.integers(-2, 3, (500, 128)).astype(np.float32)
    qo = np.arange(P + 1) * nq; to = np.arange(P + 1) * nt
    r = mcv.match_knn2_batch(q, qo, t, to, 0.8, norm=mcv.NORM_L2)
    for p in range(P):
        idx, dist = bf_knn.bf_knn2_l2(q[qo[p]:qo[p + 1]], t[to[p]:to[p + 1]])
        np.testing.assert_array_equal(r.train_idx[qo[p]:qo[p + 1]], idx)
        np.testing.assert_array_equal(r.distance[qo[p]:qo[p + 1]], dist)


def test_l2_features_container_and_empty():
    rng = np.random.default_rng(1)
    q = rng.integers(0, 255, (50, 128)).astype(np.float32)
    t = rng.integers(0, 255, (80, 128)).astype(np.float32)
    r = mcv.match_features(q, t, 0.8)
    idx, dist = bf_knn.bf_knn2_l2(q, t)
    np.testing.assert_array_equal(r.train_idx, idx)
    r = mcv.match_knn2(q, np.zeros((0, 128), np.float32), 0.8, norm=mcv.NORM_L2)
    assert (r.train_idx == -1).all()
