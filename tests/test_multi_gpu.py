"""GPU: the multi-GPU forms behind the C ABI (SURVEY.md section 8b deviceMask, 8e).  One ctypes call drives every
device of the mask -- no torch, no torchrun in the data path.  On a one-GPU box the same exports run with a one-bit
mask (and a repeated-device partition is exercised through the block partitioner); with >= 2 GPUs the results must
equal the single-device ones bit for bit."""
import numpy as np
import pytest

import minicv_b200 as mcv
from minicv_b200 import estimation as est, native, synth
from oracle import bf_knn

pytestmark = pytest.mark.gpu


def _mask(lib):
    n = lib.mcvDeviceCount()
    return n, (1 << n) - 1


def test_match_multi_equals_single(lib):
    n, mask = _mask(lib)
    q, qo, t, to = synth.ragged_descriptor_batch(11, 40, 32, 1500)
    ref = mcv.match_knn2_batch(q, qo, t, to, 0.8)
    for m in {1, mask, 0}:
        r = mcv.match_knn2_batch_multi(q, qo, t, to, 0.8, device_mask=m)
        np.testing.assert_array_equal(r.train_idx, ref.train_idx)
        np.testing.assert_array_equal(r.distance, ref.distance)
        np.testing.assert_array_equal(r.ratio_pass, ref.ratio_pass)
        assert r.n_pass == ref.n_pass
    p = 7
    idx, dist = bf_knn.bf_knn2_hamming(q[qo[p]:qo[p + 1]], t[to[p]:to[p + 1]])
    np.testing.assert_array_equal(ref.train_idx[qo[p]:qo[p + 1]], idx)
    np.testing.assert_array_equal(ref.distance[qo[p]:qo[p + 1]], dist)


def test_match_multi_bad_mask(lib):
    n, _ = _mask(lib)
    q, qo, t, to = synth.descriptor_batch(1, 2, 64, 64)
    with pytest.raises(native.NativeError):
        mcv.match_knn2_batch_multi(q, qo, t, to, 0.8, device_mask=1 << n)


@pytest.mark.parametrize("model", [est.MODEL_H, est.MODEL_F])
def test_ransac_multi_equals_single(lib, model):
    n, mask = _mask(lib)
    P = 13
    A, B, off = [], [], [0]
    for p in range(P):
        if model == est.MODEL_H:
            a, b, _, _ = synth.homography_pair(200 + p, 300 + 40 * p)
        else:
            a, b, _, _, _ = synth.two_view_pair(200 + p, 300 + 40 * p)
        A.append(a); B.append(b); off.append(off[-1] + a.shape[0])
    A, B, off = np.concatenate(A), np.concatenate(B), np.array(off, np.int64)
    for mode, stop in ((est.SAMPLE_MWC, est.STOP_ADAPTIVE), (est.SAMPLE_COUNTER, est.STOP_FIXED)):
        ref = est.ransac_batch(model, A, B, off, 3.0, 600, 0.99, sample_mode=mode, stop_mode=stop, seed=5)
        r = est.ransac_batch_multi(model, A, B, off, 3.0, 600, 0.99, sample_mode=mode, stop_mode=stop, seed=5,
                                   device_mask=mask)
        np.testing.assert_array_equal(r.ok, ref.ok)
        np.testing.assert_array_equal(r.inliers, ref.inliers)
        np.testing.assert_array_equal(r.masks, ref.masks)
        np.testing.assert_array_equal(r.models, ref.models)
        np.testing.assert_array_equal(r.iters_run, ref.iters_run)


def test_recover_pose_multi_equals_single(lib):
    n, mask = _mask(lib)
    P = 9
    A, B, off = [], [], [0]
    for p in range(P):
        a, b, _, _, _ = synth.two_view_pair(300 + p, 500 + 30 * p)
        A.append(a); B.append(b); off.append(off[-1] + a.shape[0])
    A, B, off = np.concatenate(A), np.concatenate(B), np.array(off, np.int64)
    cfg = est.RecoverPoseConfig(800.0, (320.0, 240.0), 0.999, 1.0)
    ref = est.recover_pose_batch(cfg, A, B, off)
    got = est.recover_pose_batch_multi(cfg, A, B, off, device_mask=mask)
    for g, w in zip(got, ref):
        np.testing.assert_array_equal(g, w)


@pytest.mark.parametrize("engine", [mcv.ENGINE_POPC, mcv.ENGINE_TENSOR])
def test_large_db_in_process(lib, engine):
    """cvLargeDb*: shards resident on every visible GPU, keys all-gathered inside the library (NCCL when there is
    more than one shard), merged result == unsharded oracle including a cross-shard tie."""
    n, mask = _mask(lib)
    mcv.set_hamming_engine(engine)
    try:
        q, t, gt = synth.descriptor_pair(31, 300, 150001, 32)
        t[140000] = t[17]
        q[5] = t[17]
        ridx, rdist = bf_knn.bf_knn2_hamming(q, t)
        for m in {1, mask}:
            with mcv.LargeDb(t, m) as db:
                assert db.n_train == t.shape[0] and db.n_shards == bin(m).count("1")
                if db.n_shards > 1 and lib.mcvNcclVersion() > 0:
                    assert db.uses_nccl
                for _ in range(2):
                    idx, dist, ok, npass = db.match(q, 0.8)
                    np.testing.assert_array_equal(idx, ridx.astype(np.int64))
                    np.testing.assert_array_equal(dist, rdist)
                    np.testing.assert_array_equal(ok, bf_knn.ratio_test(rdist, ridx, 0.8))
                    assert npass == int(ok.sum())
                assert tuple(idx[5]) == (17, 140000)
    finally:
        mcv.set_hamming_engine(mcv.ENGINE_AUTO)


def test_large_db_p2p_transport(lib, monkeypatch):
    n, mask = _mask(lib)
    if n < 2:
        pytest.skip("needs 2 GPUs")
    monkeypatch.setenv("MINICV_SHARD_TRANSPORT", "p2p")
    q, t, _ = synth.descriptor_pair(32, 200, 50000, 32)
    ridx, rdist = bf_knn.bf_knn2_hamming(q, t)
    with mcv.LargeDb(t, mask) as db:
        assert not db.uses_nccl
        idx, dist, ok, _ = db.match(q, 0.8)
    np.testing.assert_array_equal(idx, ridx.astype(np.int64))
    np.testing.assert_array_equal(dist, rdist)


def test_shard_comm_single_rank(lib):
    """One-process-per-GPU form with a world of one: id -> communicator -> sharded match on a torch stream."""
    import torch
    if lib.mcvNcclVersion() == 0:
        pytest.skip("NCCL not loadable")
    import ctypes as C
    idb = (C.c_uint8 * 128)()
    assert lib.mcvShardCommUniqueId(idb) == 0, native.last_error()
    comm = lib.mcvShardCommInit(idb, 1, 0, 0)
    assert comm, native.last_error()
    try:
        dev = torch.device("cuda:0")
        q, t, _ = synth.descriptor_pair(33, 257, 30000, 32)
        dq, dt = torch.from_numpy(q).to(dev), torch.from_numpy(t).to(dev)
        idx = torch.empty((257, 2), dtype=torch.int64, device=dev)
        dd = torch.empty((257, 2), dtype=torch.float32, device=dev)
        ok = torch.empty(257, dtype=torch.uint8, device=dev)
        s = torch.cuda.current_stream().cuda_stream
        rc = lib.cvMatchKnn2ShardedDevice(comm, dq.data_ptr(), 257, dt.data_ptr(), 30000, 0, 32, 0.8, idx.data_ptr(),
                                          dd.data_ptr(), ok.data_ptr(), s)
        assert rc == 0, native.last_error()
        torch.cuda.synchronize()
        ridx, rdist = bf_knn.bf_knn2_hamming(q, t)
        np.testing.assert_array_equal(idx.cpu().numpy(), ridx.astype(np.int64))
        np.testing.assert_array_equal(dd.cpu().numpy(), rdist)
        # an index base that would carry into the distance word is refused
        rc = lib.cvMatchKnn2ShardedDevice(comm, dq.data_ptr(), 257, dt.data_ptr(), 30000, (1 << 32) - 100, 32, 0.8,
                                          idx.data_ptr(), dd.data_ptr(), ok.data_ptr(), s)
        assert rc == -1 and "32-bit" in native.last_error()
    finally:
        lib.mcvShardCommDestroy(comm)
