"""CPU, world_size 2 over gloo: the host-side logic of the multi-GPU path -- pair partitioning
without a collective, and the sharded large-database merge (keys -> all_gather -> top-2)."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from minicv_b200 import sharding, synth
from oracle import bf_knn


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # ---- C5: train set sharded by contiguous range, queries replicated
        q, t, _ = synth.descriptor_pair(77, 300, 2001, 32)
        t[1500] = t[10]                       # cross-shard tie: lowest global index must win
        q[0] = t[10]
        lo, hi = sharding.block_range(t.shape[0], world, rank)
        idx, d = bf_knn.bf_knn2_hamming(q, t[lo:hi])       # stands in for the device kernel
        keys = torch.from_numpy(sharding.pack_keys(d, idx, lo).view(np.int64))
        gathered = [torch.empty_like(keys) for _ in range(world)]
        dist.all_gather(gathered, keys)
        allk = torch.stack(gathered).numpy().view(np.uint64)
        merged = bf_knn.merge_top2_keys(allk)
        midx, mdist = sharding.unpack_keys(merged)
        ref_idx, ref_dist = bf_knn.bf_knn2_hamming(q, t)
        assert (midx == ref_idx).all() and (mdist == ref_dist).all()
        assert tuple(midx[0]) == (10, 1500)
        # ---- C2-C4: pairs partitioned, no collective; results concatenate to the unsharded answer
        n_pairs = 5
        p0, p1 = sharding.block_range(n_pairs, world, rank)
        mine = []
        for p in range(p0, p1):
            qq, tt, _ = synth.descriptor_pair(200 + p, 40, 60, 32)
            mine.append(bf_knn.bf_knn2_hamming(qq, tt)[0])
        np.save(os.path.join(out_dir, f"r{rank}.npy"), np.concatenate(mine) if mine else np.zeros((0, 2), np.int32))
        dist.barrier()
    finally:
        dist.destroy_process_group()


def test_block_range_covers_everything():
    for n in (0, 1, 5, 1024, 1025):
        for world in (1, 2, 3, 8):
            r = [sharding.block_range(n, world, k) for k in range(world)]
            assert r[0][0] == 0 and r[-1][1] == n
            assert all(r[k][1] == r[k + 1][0] for k in range(world - 1))
            sizes = [b - a for a, b in r]
            assert max(sizes) - min(sizes) <= 1


def test_key_pack_roundtrip():
    idx = np.array([[3, -1], [7, 9]], np.int32)
    d = np.array([[5, 0], [0, 256]], np.float32)
    k = sharding.pack_keys(d, idx, 1000)
    i2, d2 = sharding.unpack_keys(k)
    assert i2.tolist() == [[1003, -1], [1007, 1009]]
    assert d2[0, 0] == 5 and d2[1, 1] == 256 and d2[0, 1] == np.finfo(np.float32).max


def test_world2_gloo_shard_merge_and_pair_partition(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    got = np.concatenate([np.load(tmp_path / f"r{r}.npy") for r in range(world)])
    ref = np.concatenate([bf_knn.bf_knn2_hamming(*synth.descriptor_pair(200 + p, 40, 60, 32)[:2])[0] for p in range(5)])
    assert (got == ref).all()
