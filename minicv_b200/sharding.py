"""Multi-GPU partitioning of the path.

* Pair batches (C2-C4): image pairs are independent -> static block partition of the pair list
  over ranks, no data-path collective.
* Large database (C5): the TRAIN set is sharded by contiguous index range; every rank matches
  the replicated queries against its shard, emits local top-2 as packed keys
  ``(dist << 32) | globalIdx``, one all_gather (NCCL over NVLink; gloo in the CPU tests) collects
  ``world x Q x 2`` keys and the two smallest keys per query are the global result -- the
  lowest-global-index tie-break of BFMatcher survives because the index is the low word of the key.
"""
from __future__ import annotations

import numpy as np

from . import native

NO_KEY = np.uint64(0xFFFFFFFFFFFFFFFF)


def block_range(n: int, world: int, rank: int) -> tuple[int, int]:
    """Contiguous block [lo, hi) of ``n`` units owned by ``rank`` (sizes differ by at most 1)."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def pack_keys(dist: np.ndarray, idx: np.ndarray, global_base: int) -> np.ndarray:
    """Host-side key packing (tests / reference for the device format)."""
    d = dist.astype(np.uint64)
    i = (idx.astype(np.int64) + global_base).astype(np.uint64)
    keys = (d << np.uint64(32)) | i
    keys[idx < 0] = NO_KEY
    return keys


def unpack_keys(keys: np.ndarray):
    idx = (keys & np.uint64(0xFFFFFFFF)).astype(np.int64)
    dist = (keys >> np.uint64(32)).astype(np.float32)
    none = keys == NO_KEY
    idx[none] = -1
    dist[none] = np.finfo(np.float32).max
    return idx, dist


def match_large_db(d_query, d_train_shard, global_base: int, desc_bytes: int, ratio: float,
                   group=None, device: int = -1, stream=None):
    """C5 on the GPU: local shard keys -> all_gather -> merge.  ``d_query`` [Q, descBytes] and
    ``d_train_shard`` [T_local, descBytes] are uint8 CUDA tensors.  Returns
    (train_idx int64 [Q,2] global indices, dist fp32 [Q,2], ratio_pass u8 [Q])."""
    import torch
    import torch.distributed as dist
    lib = native.load()
    nq = d_query.shape[0]
    dev = d_query.device
    s = (stream or torch.cuda.current_stream(dev)).cuda_stream
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    keys = torch.empty((nq, 2), dtype=torch.int64, device=dev)
    native.check(lib.cvMatchKnn2ShardKeysDevice(d_query.data_ptr(), nq, d_train_shard.data_ptr(),
                                                d_train_shard.shape[0], int(global_base), desc_bytes,
                                                keys.data_ptr(), device, s), "cvMatchKnn2ShardKeysDevice")
    if world > 1:
        gathered = torch.empty((world, nq, 2), dtype=torch.int64, device=dev)
        dist.all_gather_into_tensor(gathered, keys, group=group)
    else:
        gathered = keys.view(1, nq, 2)
    idx64 = torch.empty((nq, 2), dtype=torch.int64, device=dev)
    dd = torch.empty((nq, 2), dtype=torch.float32, device=dev)
    ok = torch.empty(nq, dtype=torch.uint8, device=dev)
    native.check(lib.cvMergeShardKeysDevice(gathered.data_ptr(), world, nq, ratio, idx64.data_ptr(),
                                            dd.data_ptr(), ok.data_ptr(), device, s), "cvMergeShardKeysDevice")
    return idx64, dd, ok
