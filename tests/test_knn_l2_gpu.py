"""GPU parity of the L2 matcher: bit-exact vs the oracle (same accumulation order), bit-exact vs
cv2 on integer-valued (SIFT-like) descriptors, within 2e-6 relative on general floats."""
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
    np.testing.assert_allclose(r.distance, g["gen_dist"], rtol=2e-6)


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
