"""Descriptor matching -- host-side mirror of the (new) matcher surface of ``MiniCV.OpenCV``.

The reference has no matcher; these wrappers follow the style of its
F# wrappers (``src/MiniCV/OpenCV.fs:855-920``): allocate outputs, pin arrays, call native,
return plain arrays.  Inputs are the descriptor blobs ``detectFeatures`` yields
(``OpenCV.fs:440-535``: row-major ``points x descriptorDimension``).
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import numpy as np

from . import native

NORM_L2 = 4
NORM_HAMMING = 6


@dataclass
class KnnResult:
    train_idx: np.ndarray   # [nq, 2] int32, -1 where the neighbour does not exist
    distance: np.ndarray    # [nq, 2] float32 (Hamming distances stored as float, like DMatch)
    ratio_pass: np.ndarray  # [nq] uint8
    n_pass: int


def _rows(a: np.ndarray, norm: int) -> tuple[np.ndarray, int]:
    if norm == NORM_HAMMING:
        a = np.ascontiguousarray(a, dtype=np.uint8)
        return a, a.shape[1] if a.ndim == 2 else 0
    a = np.ascontiguousarray(a, dtype=np.float32)
    return a, (a.shape[1] * 4 if a.ndim == 2 else 0)


def _ptr(a: np.ndarray) -> int:
    return a.ctypes.data if a.size else 0


def match_knn2(query: np.ndarray, train: np.ndarray, ratio: float = 0.8,
               norm: int = NORM_HAMMING) -> KnnResult:
    """``BFMatcher(norm).knnMatch(query, train, 2)`` + Lowe ratio test, on the GPU."""
    lib = native.load()
    q, db = _rows(query, norm)
    t, dbt = _rows(train, norm)
    if t.shape[0] and q.shape[0] and db != dbt:
        raise ValueError("query/train descriptor sizes differ")
    nq, nt = q.shape[0], t.shape[0]
    idx = np.full((nq, 2), -1, dtype=np.int32)
    dist = np.full((nq, 2), np.finfo(np.float32).max, dtype=np.float32)
    ok = np.zeros(nq, dtype=np.uint8)
    n = native.check(lib.cvMatchKnn2(_ptr(q), nq, _ptr(t), nt, db or dbt, norm, ratio,
                                     _ptr(idx), _ptr(dist), _ptr(ok)), "cvMatchKnn2")
    return KnnResult(idx, dist, ok, n)


def match_knn2_batch(query: np.ndarray, query_offsets, train: np.ndarray, train_offsets,
                     ratio: float = 0.8, norm: int = NORM_HAMMING, device: int = -1,
                     out: KnnResult | None = None) -> KnnResult:
    """Batched matcher over host buffers (pinned buffers make the copies asynchronous).

    ``query_offsets`` / ``train_offsets``: ``nPairs+1`` row offsets.  ``train_idx`` is relative
    to each pair's own train block."""
    lib = native.load()
    q, db = _rows(query, norm)
    t, dbt = _rows(train, norm)
    qo = np.ascontiguousarray(query_offsets, dtype=np.int64)
    to = np.ascontiguousarray(train_offsets, dtype=np.int64)
    if qo.shape != to.shape or qo.ndim != 1 or qo.size < 1:
        raise ValueError("offset arrays must both have nPairs+1 entries")
    n_pairs = qo.size - 1
    nq = q.shape[0]
    if out is None:
        out = KnnResult(np.full((nq, 2), -1, dtype=np.int32),
                        np.full((nq, 2), np.finfo(np.float32).max, dtype=np.float32),
                        np.zeros(nq, dtype=np.uint8), 0)
    out.n_pass = native.check(
        lib.cvMatchKnn2Batch(_ptr(q), _ptr(qo), _ptr(t), _ptr(to), n_pairs, db or dbt, norm, ratio,
                             _ptr(out.train_idx), _ptr(out.distance), _ptr(out.ratio_pass), device),
        "cvMatchKnn2Batch")
    return out


def match_knn2_batch_device(d_query, query_offsets, d_train, train_offsets, desc_bytes: int,
                            ratio: float, d_train_idx, d_dist, d_ratio_pass, norm: int = NORM_HAMMING,
                            device: int = -1, stream: int = 0) -> None:
    """Device-resident batched matcher: arguments are torch CUDA tensors (or raw device
    pointers as ints); work is enqueued on ``stream`` (a ``cudaStream_t`` as int)."""
    lib = native.load()

    def dp(x):
        return x if isinstance(x, int) else (x.data_ptr() if x is not None else 0)

    qo = np.ascontiguousarray(query_offsets, dtype=np.int64)
    to = np.ascontiguousarray(train_offsets, dtype=np.int64)
    native.check(
        lib.cvMatchKnn2BatchDevice(dp(d_query), _ptr(qo), dp(d_train), _ptr(to), qo.size - 1,
                                   desc_bytes, norm, ratio, dp(d_train_idx), dp(d_dist),
                                   dp(d_ratio_pass), device, stream),
        "cvMatchKnn2BatchDevice")


def gather_matches_device(d_kp_query, d_kp_train, query_offsets, train_offsets, d_train_idx, d_ratio_pass,
                          d_a, d_b, device: int = -1, stream: int = 0):
    """Device-resident hand-off matcher -> estimators (``cvGatherMatchesDevice``): returns the host
    array of nPairs+1 correspondence offsets; ``d_a`` / ``d_b`` receive the packed (x, y) doubles."""
    lib = native.load()

    def dp(x):
        return x if isinstance(x, int) else x.data_ptr()

    qo = np.ascontiguousarray(query_offsets, dtype=np.int64)
    to = np.ascontiguousarray(train_offsets, dtype=np.int64)
    out = np.zeros(qo.size, dtype=np.int64)
    native.check(lib.cvGatherMatchesDevice(dp(d_kp_query), dp(d_kp_train), _ptr(qo), _ptr(to), qo.size - 1,
                                           dp(d_train_idx), dp(d_ratio_pass), dp(d_a), dp(d_b), _ptr(out),
                                           device, stream), "cvGatherMatchesDevice")
    return out


def match_features(query_desc: np.ndarray, train_desc: np.ndarray, ratio: float = 0.8) -> KnnResult:
    """Matcher over two ``DetectorResult`` containers (reference layout
    ``src/MiniCVNative/MiniCVNative.h:23-29``): u8 blobs match with Hamming, f32 with L2."""
    lib = native.load()

    def make(a: np.ndarray):
        if a.dtype == np.uint8:
            ty = 0
        elif a.dtype == np.float32:
            ty = 5
        else:
            raise ValueError("descriptors must be uint8 or float32")
        a = np.ascontiguousarray(a)
        r = native.DetectorResult()
        r.PointCount = a.shape[0]
        r.DescriptorEntries = a.size
        r.DescriptorElementType = ty
        r.Points = None
        r.Descriptors = C.cast(a.ctypes.data, C.POINTER(C.c_uint8)) if a.size else None
        return r, a

    rq, keep_q = make(query_desc)
    rt, keep_t = make(train_desc)
    nq = keep_q.shape[0]
    idx = np.full((nq, 2), -1, dtype=np.int32)
    dist = np.full((nq, 2), np.finfo(np.float32).max, dtype=np.float32)
    ok = np.zeros(nq, dtype=np.uint8)
    n = native.check(lib.cvMatchFeatures(C.byref(rq), C.byref(rt), ratio, _ptr(idx), _ptr(dist),
                                         _ptr(ok)), "cvMatchFeatures")
    return KnnResult(idx, dist, ok, n)


ENGINE_AUTO, ENGINE_POPC, ENGINE_TENSOR, ENGINE_TENSOR_PACKED = 0, 1, 2, 3


def set_hamming_engine(engine: int) -> None:
    """Choose the Hamming matcher engine (results are bit-identical): 0 auto, 1 integer pipe
    (XOR + POPC), 2 tcgen05 int8 tensor cores, 3 tcgen05 fp8 with packed accumulators."""
    if native.load().mcvSetHammingEngine(engine) != 0:
        raise ValueError("engine must be 0..3")


def set_l2_engine(engine: int) -> None:
    """Choose the L2 matcher engine: 0 auto, 1 fp32 CUDA cores, 2 tcgen05 bf16 split + exact re-scoring."""
    if native.load().mcvSetL2Engine(engine) != 0:
        raise ValueError("engine must be 0..2")


def match_knn2_batch_multi(query: np.ndarray, query_offsets, train: np.ndarray, train_offsets,
                           ratio: float = 0.8, norm: int = NORM_HAMMING, device_mask: int = 0,
                           out: KnnResult | None = None) -> KnnResult:
    """``cvMatchKnn2BatchMulti``: one call, pairs block-partitioned over the GPUs of ``device_mask``
    (bit d = CUDA ordinal d, 0 = all visible devices); same results as ``match_knn2_batch``."""
    lib = native.load()
    q, db = _rows(query, norm)
    t, dbt = _rows(train, norm)
    qo = np.ascontiguousarray(query_offsets, dtype=np.int64)
    to = np.ascontiguousarray(train_offsets, dtype=np.int64)
    if qo.shape != to.shape or qo.ndim != 1 or qo.size < 1:
        raise ValueError("offset arrays must both have nPairs+1 entries")
    nq = q.shape[0]
    if out is None:
        out = KnnResult(np.full((nq, 2), -1, dtype=np.int32),
                        np.full((nq, 2), np.finfo(np.float32).max, dtype=np.float32),
                        np.zeros(nq, dtype=np.uint8), 0)
    out.n_pass = native.check(
        lib.cvMatchKnn2BatchMulti(_ptr(q), _ptr(qo), _ptr(t), _ptr(to), qo.size - 1, db or dbt, norm, ratio,
                                  _ptr(out.train_idx), _ptr(out.distance), _ptr(out.ratio_pass), device_mask),
        "cvMatchKnn2BatchMulti")
    return out


class LargeDb:
    """Train set sharded over the GPUs of ``device_mask`` and kept resident (config C5): every
    ``match`` replicates the queries, all-gathers the per-shard top-2 keys with NCCL inside the
    library and merges them.  ``train_idx`` are global indices (int64)."""

    def __init__(self, train: np.ndarray, device_mask: int = 0):
        lib = native.load()
        t = np.ascontiguousarray(train, dtype=np.uint8)
        self.desc_bytes = t.shape[1]
        self._h = lib.cvLargeDbCreate(_ptr(t), t.shape[0], t.shape[1], device_mask)
        if not self._h:
            raise native.NativeError("cvLargeDbCreate failed: " + native.last_error())
        ns, nccl, nt = C.c_int(0), C.c_int(0), C.c_int64(0)
        lib.cvLargeDbInfo(self._h, C.byref(ns), C.byref(nccl), C.byref(nt))
        self.n_shards, self.uses_nccl, self.n_train = ns.value, bool(nccl.value), nt.value

    def match(self, query: np.ndarray, ratio: float = 0.8):
        lib = native.load()
        q = np.ascontiguousarray(query, dtype=np.uint8)
        nq = q.shape[0]
        idx = np.full((nq, 2), -1, dtype=np.int64)
        dist = np.full((nq, 2), np.finfo(np.float32).max, dtype=np.float32)
        ok = np.zeros(nq, dtype=np.uint8)
        n = native.check(lib.cvLargeDbMatch(self._h, _ptr(q), nq, ratio, _ptr(idx), _ptr(dist), _ptr(ok)),
                         "cvLargeDbMatch")
        return idx, dist, ok, n

    def close(self):
        if self._h:
            native.load().cvLargeDbDestroy(self._h)
            self._h = None

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class PinnedArray:
    """numpy view over ``mcvAllocPinned`` memory (CUDA-pinned, so the exports DMA it directly)."""

    def __init__(self, shape, dtype):
        lib = native.load()
        self.nbytes = int(np.prod(shape)) * np.dtype(dtype).itemsize
        self._p = lib.mcvAllocPinned(max(self.nbytes, 1))
        if not self._p:
            raise native.NativeError("mcvAllocPinned failed: " + native.last_error())
        buf = (C.c_uint8 * max(self.nbytes, 1)).from_address(self._p)
        self.array = np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape))).reshape(shape)

    def free(self):
        if self._p:
            self.array = None
            native.load().mcvFreePinned(self._p)
            self._p = None
