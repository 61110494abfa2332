"""GPU: sharded large-database matching (C5 shape, single process): shard keys for several
contiguous shards merged on the device == unsharded oracle, on both engines."""
import numpy as np
import pytest

import minicv_b200 as mcv
from minicv_b200 import native, sharding, synth
from oracle import bf_knn

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("engine", [mcv.ENGINE_POPC, mcv.ENGINE_TENSOR])
def test_sharded_keys_merge_equals_unsharded(engine):
    import torch
    mcv.set_hamming_engine(engine)
    try:
        lib = native.load()
        dev = torch.device("cuda:0")
        q, t, gt = synth.descriptor_pair(31, 400, 150001, 32)
        t[140000] = t[17]                       # cross-shard tie
        q[5] = t[17]
        dq = torch.from_numpy(q).to(dev)
        dt = torch.from_numpy(t).to(dev)
        world = 4
        s = torch.cuda.current_stream().cuda_stream
        keys = torch.empty((world, 400, 2), dtype=torch.int64, device=dev)
        for r in range(world):
            lo, hi = sharding.block_range(t.shape[0], world, r)
            shard = dt[lo:hi].contiguous()
            rc = lib.cvMatchKnn2ShardKeysDevice(dq.data_ptr(), 400, shard.data_ptr(), hi - lo, lo, 32,
                                                keys[r].data_ptr(), 0, s)
            assert rc == 0, native.last_error()
            torch.cuda.synchronize()
        idx64 = torch.empty((400, 2), dtype=torch.int64, device=dev)
        dd = torch.empty((400, 2), dtype=torch.float32, device=dev)
        ok = torch.empty(400, dtype=torch.uint8, device=dev)
        rc = lib.cvMergeShardKeysDevice(keys.data_ptr(), world, 400, 0.8, idx64.data_ptr(), dd.data_ptr(),
                                        ok.data_ptr(), 0, s)
        assert rc == 0, native.last_error()
        torch.cuda.synchronize()
        ridx, rdist = bf_knn.bf_knn2_hamming(q, t)
        np.testing.assert_array_equal(idx64.cpu().numpy(), ridx.astype(np.int64))
        np.testing.assert_array_equal(dd.cpu().numpy(), rdist)
        np.testing.assert_array_equal(ok.cpu().numpy(), bf_knn.ratio_test(rdist, ridx, 0.8))
        assert tuple(idx64[5].tolist()) == (17, 140000)
        # the single-process helper (world 1) gives the same answer
        i2, d2, o2 = sharding.match_large_db(dq, dt, 0, 32, 0.8)
        np.testing.assert_array_equal(i2.cpu().numpy(), ridx.astype(np.int64))
    finally:
        mcv.set_hamming_engine(mcv.ENGINE_AUTO)


@pytest.mark.parametrize("engine", [mcv.ENGINE_POPC, mcv.ENGINE_TENSOR])
def test_train_set_beyond_packed_key_range(engine):
    """More than 2^22 train rows: the integer-pipe engine must split (22-bit packed index), the
    tensor engine sweeps >32k stages; both must keep exact indices and the lowest-index tie-break."""
    rng = np.random.default_rng(3)
    nt = (1 << 22) + 1000
    t = rng.integers(0, 256, (nt, 32), dtype=np.uint8)
    q = rng.integers(0, 256, (12, 32), dtype=np.uint8)
    q[0] = t[nt - 1]; q[1] = t[5]; t[(1 << 22) + 7] = t[5]      # tie across the 2^22 boundary
    mcv.set_hamming_engine(engine)
    try:
        r = mcv.match_knn2(q, t, 0.8)
    finally:
        mcv.set_hamming_engine(mcv.ENGINE_AUTO)
    idx, dist = bf_knn.bf_knn2_hamming(q, t)
    np.testing.assert_array_equal(r.train_idx, idx)
    np.testing.assert_array_equal(r.distance, dist)
    assert r.train_idx[0, 0] == nt - 1 and tuple(r.train_idx[1]) == (5, (1 << 22) + 7)
