"""GPU parity of the tcgen05 (tensor-core) Hamming engine: bit-identical to the oracle, the cv2
goldens and the integer-pipe engine, including ties, ragged batches and train splits."""
import numpy as np
import pytest

import minicv_b200 as mcv
from minicv_b200 import synth
from oracle import bf_knn
from conftest import GOLDEN

pytestmark = pytest.mark.gpu


# engine 3 (fp8 packed accumulators) is a build option since round 2 (make EXTRA=-DMCV_WITH_TC2); in the default library the
# switch value falls back to the int8 engine, which the second parameter checks
@pytest.fixture(autouse=True, params=[mcv.ENGINE_TENSOR, mcv.ENGINE_TENSOR_PACKED], ids=["int8", "engine3-switch"])
def tensor_engine(request):
    mcv.set_hamming_engine(request.param)
    yield
    mcv.set_hamming_engine(mcv.ENGINE_AUTO)


def check(q, t, ratio=0.8):
    r = mcv.match_knn2(q, t, ratio)
    idx, dist = bf_knn.bf_knn2_hamming(q, t)
    np.testing.assert_array_equal(r.train_idx, idx)
    np.testing.assert_array_equal(r.distance, dist)
    np.testing.assert_array_equal(r.ratio_pass, bf_knn.ratio_test(dist, idx, ratio))
    return r


@pytest.mark.parametrize("name", ["b256", "b488", "b512", "one"])
def test_golden_cv2(name):
    g = np.load(GOLDEN / "knn_hamming.npz")
    r = mcv.match_knn2(g[f"{name}_q"], g[f"{name}_t"], 0.8)
    np.testing.assert_array_equal(r.train_idx, g[f"{name}_idx"])
    np.testing.assert_array_equal(r.distance, g[f"{name}_dist"])


@pytest.mark.parametrize("nb", [16, 32, 61, 64])
@pytest.mark.parametrize("nq,nt", [(1, 1), (3, 2), (255, 257), (256, 128), (1000, 513), (37, 5000)])
def test_vs_oracle_shapes(nb, nq, nt):
    q, t, _ = synth.descriptor_pair(nq * 7 + nt, nq, nt, nb)
    check(q, t)


def test_all_equal_rows_tie_break():
    t = np.tile(np.arange(32, dtype=np.uint8), (3000, 1))
    q = t[:300].copy()
    r = check(q, t)
    assert (r.train_idx[:, 0] == 0).all() and (r.train_idx[:, 1] == 1).all()


def test_full_size_pair_c2():
    q, t, gt = synth.descriptor_pair(0)
    r = check(q, t)
    true = gt >= 0
    assert (r.train_idx[true, 0] == gt[true]).mean() > 0.999


def test_ragged_batch_vs_oracle():
    q, qo, t, to = synth.ragged_descriptor_batch(3, 24, 32)
    r = mcv.match_knn2_batch(q, qo, t, to, 0.8)
    for p in range(len(qo) - 1):
        idx, dist = bf_knn.bf_knn2_hamming(q[qo[p]:qo[p + 1]], t[to[p]:to[p + 1]])
        np.testing.assert_array_equal(r.train_idx[qo[p]:qo[p + 1]], idx)
        np.testing.assert_array_equal(r.distance[qo[p]:qo[p + 1]], dist)


def test_split_tail_and_hi_lo_boundaries():
    # train counts around the 128/256-row stage and half-split boundaries, few queries (forces splits)
    for nt in (127, 128, 129, 255, 256, 257, 383, 1023, 1025, 4097):
        q, t, _ = synth.descriptor_pair(900 + nt, 70, nt, 32)
        check(q, t)


def test_engines_agree_on_batch():
    q, qo, t, to = synth.descriptor_batch(200, 8)
    r_tc = mcv.match_knn2_batch(q, qo, t, to, 0.8)
    mcv.set_hamming_engine(mcv.ENGINE_POPC)
    r_pc = mcv.match_knn2_batch(q, qo, t, to, 0.8)
    np.testing.assert_array_equal(r_tc.train_idx, r_pc.train_idx)
    np.testing.assert_array_equal(r_tc.distance, r_pc.distance)
    np.testing.assert_array_equal(r_tc.ratio_pass, r_pc.ratio_pass)
    assert r_tc.n_pass == r_pc.n_pass
