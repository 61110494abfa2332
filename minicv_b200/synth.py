"""Deterministic synthetic workloads for tests and bench.py (definitions: SURVEY.md section 8d).

numpy only; no GPU, no oracle.  Every generator is a pure function of its seed."""
from __future__ import annotations

import numpy as np


def descriptor_pair(seed: int, n_query: int = 8192, n_train: int = 8192, desc_bytes: int = 32,
                    flip_p: float = 0.10, true_frac: float = 0.5):
    """C2 pair: random train set; ``true_frac`` of the queries are copies of distinct train
    rows with every bit flipped w.p. ``flip_p`` (true matches), the rest are fresh random rows
    (nearest neighbours ~95-105 bits away with frequent d1 == d2 ties -> exercises the
    lowest-index tie-break).  Returns (query, train, gt) with gt[i] = source train row or -1."""
    rng = np.random.default_rng(1000 + seed)
    train = rng.integers(0, 256, (n_train, desc_bytes), dtype=np.uint8)
    query = rng.integers(0, 256, (n_query, desc_bytes), dtype=np.uint8)
    gt = np.full(n_query, -1, dtype=np.int64)
    n_true = min(int(n_query * true_frac), n_train)
    if n_true > 0:
        src = rng.permutation(n_train)[:n_true]
        flips = rng.random((n_true, desc_bytes * 8), dtype=np.float32) < flip_p
        noise = np.packbits(flips, axis=1)
        pos = rng.permutation(n_query)[:n_true]
        query[pos] = train[src] ^ noise
        gt[pos] = src
    return query, train, gt


def descriptor_batch(seed0: int, n_pairs: int, n_query: int = 8192, n_train: int = 8192,
                     desc_bytes: int = 32, out_query=None, out_train=None):
    """``n_pairs`` C2 pairs packed back to back.  Returns (query, q_off, train, t_off)."""
    q = out_query if out_query is not None else np.empty((n_pairs * n_query, desc_bytes), np.uint8)
    t = out_train if out_train is not None else np.empty((n_pairs * n_train, desc_bytes), np.uint8)
    for p in range(n_pairs):
        qq, tt, _ = descriptor_pair(seed0 + p, n_query, n_train, desc_bytes)
        q[p * n_query:(p + 1) * n_query] = qq
        t[p * n_train:(p + 1) * n_train] = tt
    q_off = np.arange(n_pairs + 1, dtype=np.int64) * n_query
    t_off = np.arange(n_pairs + 1, dtype=np.int64) * n_train
    return q, q_off, t, t_off


def ragged_descriptor_batch(seed: int, n_pairs: int, desc_bytes: int = 32, max_rows: int = 700):
    """Ragged batch with empty and tiny pairs for edge-case tests."""
    rng = np.random.default_rng(seed)
    nq = rng.integers(0, max_rows, n_pairs)
    nt = rng.integers(0, max_rows, n_pairs)
    nq[0] = 0
    if n_pairs > 1:
        nt[1] = 0
    if n_pairs > 2:
        nt[2] = 1
    if n_pairs > 3:
        nq[3] = 1
    qs, ts = [], []
    for p in range(n_pairs):
        qq, tt, _ = descriptor_pair(seed * 131 + p, int(nq[p]), max(int(nt[p]), 0), desc_bytes)
        qs.append(qq)
        ts.append(tt[:nt[p]])
    q = np.concatenate(qs) if qs else np.empty((0, desc_bytes), np.uint8)
    t = np.concatenate(ts) if ts else np.empty((0, desc_bytes), np.uint8)
    q_off = np.concatenate([[0], np.cumsum(nq)]).astype(np.int64)
    t_off = np.concatenate([[0], np.cumsum(nt)]).astype(np.int64)
    return q, q_off, t, t_off
