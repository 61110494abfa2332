"""GPU parity: Hamming kNN-2 + ratio through the C ABI vs the oracle and the cv2 goldens.
Bit-exact: indices (lowest-index tie-break), distances, ratio flags."""
import numpy as np
import pytest

import minicv_b200 as mcv
from minicv_b200 import synth
from oracle import bf_knn
from conftest import GOLDEN

pytestmark = pytest.mark.gpu


def check(q, t, ratio=0.8):
    r = mcv.match_knn2(q, t, ratio)
    idx, dist = bf_knn.bf_knn2_hamming(q, t)
    np.testing.assert_array_equal(r.train_idx, idx)
    np.testing.assert_array_equal(r.distance, dist)
    ok = bf_knn.ratio_test(dist, idx, ratio)
    np.testing.assert_array_equal(r.ratio_pass, ok)
    assert r.n_pass == int(ok.sum())
    return r


@pytest.mark.parametrize("name", ["b256", "b488", "b512", "one"])
def test_golden_cv2(name):
    g = np.load(GOLDEN / "knn_hamming.npz")
    r = mcv.match_knn2(g[f"{name}_q"], g[f"{name}_t"], 0.8)
    np.testing.assert_array_equal(r.train_idx, g[f"{name}_idx"])
    np.testing.assert_array_equal(r.distance, g[f"{name}_dist"])


@pytest.mark.parametrize("nb", [16, 32, 61, 64, 128])
@pytest.mark.parametrize("nq,nt", [(1, 1), (3, 2), (255, 257), (256, 256), (1000, 513), (37, 5000)])
def test_vs_oracle_shapes(nb, nq, nt):
    q, t, _ = synth.descriptor_pair(nq * 7 + nt, nq, nt, nb)
    check(q, t)


def test_empty_train_and_query():
    q = np.zeros((5, 32), np.uint8)
    r = mcv.match_knn2(q, np.zeros((0, 32), np.uint8))
    assert (r.train_idx == -1).all() and r.n_pass == 0
    r = mcv.match_knn2(np.zeros((0, 32), np.uint8), q)
    assert r.train_idx.shape == (0, 2)


def test_all_equal_rows_tie_break():
    t = np.tile(np.arange(32, dtype=np.uint8), (3000, 1))
    q = t[:300].copy()
    r = check(q, t)
    assert (r.train_idx[:, 0] == 0).all() and (r.train_idx[:, 1] == 1).all()


def test_full_size_pair_c2():
    q, t, gt = synth.descriptor_pair(0)          # 8192 x 8192 x 256 bit
    r = check(q, t)
    true = gt >= 0
    assert (r.train_idx[true, 0] == gt[true]).mean() > 0.999
    assert r.ratio_pass[true].mean() > 0.99


def test_single_pair_split_path_large_train():
    # few queries, many train rows -> the train range is split over blocks and merged
    q, t, _ = synth.descriptor_pair(5, 300, 70000, 32)
    check(q, t)


def test_ragged_batch_vs_oracle():
    q, qo, t, to = synth.ragged_descriptor_batch(3, 24, 32)
    r = mcv.match_knn2_batch(q, qo, t, to, 0.8)
    for p in range(len(qo) - 1):
        qq, tt = q[qo[p]:qo[p + 1]], t[to[p]:to[p + 1]]
        idx, dist = bf_knn.bf_knn2_hamming(qq, tt)
        np.testing.assert_array_equal(r.train_idx[qo[p]:qo[p + 1]], idx)
        np.testing.assert_array_equal(r.distance[qo[p]:qo[p + 1]], dist)
        np.testing.assert_array_equal(r.ratio_pass[qo[p]:qo[p + 1]], bf_knn.ratio_test(dist, idx, 0.8))


def test_batch_properties_full_size():
    # size-independent properties at the bench shape (16 pairs of 8192 x 8192):
    # every true match is recovered and batch == per-pair calls
    n = 16
    q, qo, t, to = synth.descriptor_batch(100, n)
    r = mcv.match_knn2_batch(q, qo, t, to, 0.8)
    one = mcv.match_knn2(q[qo[5]:qo[6]], t[to[5]:to[6]], 0.8)
    np.testing.assert_array_equal(r.train_idx[qo[5]:qo[6]], one.train_idx)
    np.testing.assert_array_equal(r.distance[qo[5]:qo[6]], one.distance)
    assert (r.distance[:, 0] <= r.distance[:, 1]).all()
    assert (r.train_idx >= 0).all() and (r.train_idx < 8192).all()
    assert (r.train_idx[:, 0] != r.train_idx[:, 1]).all()


def test_match_features_container():
    q, t, _ = synth.descriptor_pair(9, 200, 300, 61)
    r = mcv.match_features(q, t, 0.8)
    idx, dist = bf_knn.bf_knn2_hamming(q, t)
    np.testing.assert_array_equal(r.train_idx, idx)


def test_device_resident_batch_and_shard_merge():
    import torch
    from minicv_b200 import native
    lib = native.load()
    dev = torch.device("cuda:0")
    q, t, _ = synth.descriptor_pair(21, 500, 4096, 32)
    dq = torch.from_numpy(q).to(dev)
    dt = torch.from_numpy(t).to(dev)
    didx = torch.empty((500, 2), dtype=torch.int32, device=dev)
    ddist = torch.empty((500, 2), dtype=torch.float32, device=dev)
    dok = torch.empty(500, dtype=torch.uint8, device=dev)
    s = torch.cuda.current_stream().cuda_stream
    mcv.match_knn2_batch_device(dq, [0, 500], dt, [0, 4096], 32, 0.8, didx, ddist, dok, stream=s)
    torch.cuda.synchronize()
    idx, dist = bf_knn.bf_knn2_hamming(q, t)
    np.testing.assert_array_equal(didx.cpu().numpy(), idx)
    np.testing.assert_array_equal(ddist.cpu().numpy(), dist)
    # sharded train set: 4 shards of 1024 rows, keys merged == unsharded result
    keys = torch.empty((4, 500, 2), dtype=torch.int64, device=dev)
    for sh in range(4):
        rc = lib.cvMatchKnn2ShardKeysDevice(dq.data_ptr(), 500, dt[sh * 1024:(sh + 1) * 1024].data_ptr(),
                                            1024, sh * 1024, 32, keys[sh].data_ptr(), 0, s)
        assert rc == 0, native.last_error()
    idx64 = torch.empty((500, 2), dtype=torch.int64, device=dev)
    rc = lib.cvMergeShardKeysDevice(keys.data_ptr(), 4, 500, 0.8, idx64.data_ptr(), ddist.data_ptr(),
                                    dok.data_ptr(), 0, s)
    assert rc == 0, native.last_error()
    torch.cuda.synchronize()
    np.testing.assert_array_equal(idx64.cpu().numpy(), idx.astype(np.int64))
    np.testing.assert_array_equal(ddist.cpu().numpy(), dist)
    np.testing.assert_array_equal(dok.cpu().numpy(), bf_knn.ratio_test(dist, idx, 0.8))
    ok_keys = bf_knn.merge_top2_keys(keys.cpu().numpy().view(np.uint64))
    assert ((ok_keys & 0xFFFFFFFF) == idx.astype(np.uint64)).all()
