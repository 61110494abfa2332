"""GPU: the host side of the C ABI -- re-entrancy from many threads (threading row), the
pooled contexts (bounded, trimmable), pinned vs pageable caller buffers, stream-safety of the asynchronous
*Device exports, and the multi-chunk tensor-engine path bench.py runs (28 launches per step at C2)."""
import threading

import numpy as np
import pytest

import minicv_b200 as mcv
from minicv_b200 import estimation as est, native, synth
from oracle import bf_knn

pytestmark = pytest.mark.gpu


def test_sixteen_threads_identical_to_serial(lib):
    """16 host threads hammer cvMatchKnn2 + cvFindHomography concurrently: every result equals the serial one,
    and the context pool stays within its cap."""
    jobs = []
    for k in range(16):
        q, t, _ = synth.descriptor_pair(500 + k, 700 + 13 * k, 900 + 7 * k, 32)
        a, b, _, _ = synth.homography_pair(40 + k, 400 + 10 * k)
        jobs.append((q, t, a, b))
    serial = []
    for q, t, a, b in jobs:
        r = mcv.match_knn2(q, t, 0.8)
        H, m = est.findHomography(a, b, 3.0, 500, 0.995)
        serial.append((r.train_idx.copy(), r.distance.copy(), r.ratio_pass.copy(), H.copy(), m.copy()))
    out = [None] * 16
    errs = []

    def work(k):
        try:
            q, t, a, b = jobs[k]
            res = []
            for _ in range(6):
                r = mcv.match_knn2(q, t, 0.8)
                H, m = est.findHomography(a, b, 3.0, 500, 0.995)
                res.append((r.train_idx, r.distance, r.ratio_pass, H, m))
            out[k] = res
        except Exception as e:  # noqa: BLE001
            errs.append(repr(e))

    th = [threading.Thread(target=work, args=(k,)) for k in range(16)]
    for x in th:
        x.start()
    for x in th:
        x.join()
    assert not errs, errs
    for k in range(16):
        for res in out[k]:
            for got, want in zip(res, serial[k]):
                np.testing.assert_array_equal(got, want)
    assert 1 <= lib.mcvContextCount(0) <= 8


def test_threads_come_and_go_without_growing_memory(lib):
    """Contexts belong to the process-wide pool, not to threads: 40 short-lived threads in sequence reuse the
    same context (no growth), and mcvTrim returns the workspaces."""
    q, t, _ = synth.descriptor_pair(9, 2048, 2048, 32)
    mcv.match_knn2(q, t, 0.8)
    n0 = lib.mcvContextCount(0)
    b0 = lib.mcvWorkspaceBytes(0)
    for _ in range(40):
        th = threading.Thread(target=lambda: mcv.match_knn2(q, t, 0.8))
        th.start()
        th.join()
    assert lib.mcvContextCount(0) == n0
    assert lib.mcvWorkspaceBytes(0) <= b0
    assert lib.mcvWorkspaceBytes(0) > 0
    assert lib.mcvTrim() >= 1
    assert lib.mcvWorkspaceBytes(0) == 0 and lib.mcvWorkspaceBytes(1) == 0
    r = mcv.match_knn2(q, t, 0.8)          # and the library keeps working after a trim
    idx, dist = bf_knn.bf_knn2_hamming(q, t)
    np.testing.assert_array_equal(r.train_idx, idx)


def test_workspace_limit_frees_on_return(lib):
    q, t, _ = synth.descriptor_pair(10, 4096, 4096, 32)
    lib.mcvTrim()
    assert lib.mcvSetWorkspaceLimit(1 << 16) == 0
    try:
        mcv.match_knn2(q, t, 0.8)
        assert lib.mcvWorkspaceBytes(0) == 0
    finally:
        lib.mcvSetWorkspaceLimit(-1)


def test_caller_device_is_restored():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    torch.cuda.set_device(0)
    q, qo, t, to = synth.descriptor_batch(77, 2, 512, 512)
    mcv.match_knn2_batch(q, qo, t, to, 0.8, device=1)
    assert torch.cuda.current_device() == 0
    x = torch.zeros(1, device="cuda")
    assert x.device.index == 0


@pytest.mark.parametrize("engine", [mcv.ENGINE_POPC, mcv.ENGINE_TENSOR])
def test_pinned_and_pageable_buffers_agree(lib, engine):
    """numpy arrays are pageable (staged through the pinned ring, > 256 KiB); mcvAllocPinned buffers take the
    direct DMA path; mcvSetStagingMode(1) forces plain cudaMemcpyAsync.  All three give identical results."""
    mcv.set_hamming_engine(engine)
    try:
        P, n = 12, 4096
        q, qo, t, to = synth.descriptor_batch(300, P, n, n)
        ref = mcv.match_knn2_batch(q, qo, t, to, 0.8)
        p = 3
        idx, dist = bf_knn.bf_knn2_hamming(q[qo[p]:qo[p + 1]], t[to[p]:to[p + 1]])
        np.testing.assert_array_equal(ref.train_idx[qo[p]:qo[p + 1]], idx)
        np.testing.assert_array_equal(ref.distance[qo[p]:qo[p + 1]], dist)
        pq, pt = mcv.PinnedArray(q.shape, np.uint8), mcv.PinnedArray(t.shape, np.uint8)
        pi, pd, po = mcv.PinnedArray((P * n, 2), np.int32), mcv.PinnedArray((P * n, 2), np.float32), mcv.PinnedArray((P * n,), np.uint8)
        try:
            pq.array[:] = q
            pt.array[:] = t
            out = mcv.KnnResult(pi.array, pd.array, po.array, 0)
            r = mcv.match_knn2_batch(pq.array, qo, pt.array, to, 0.8, out=out)
            np.testing.assert_array_equal(r.train_idx, ref.train_idx)
            np.testing.assert_array_equal(r.distance, ref.distance)
            np.testing.assert_array_equal(r.ratio_pass, ref.ratio_pass)
            assert r.n_pass == ref.n_pass
        finally:
            for x in (pq, pt, pi, pd, po):
                x.free()
        assert lib.mcvSetStagingMode(1) == 0
        try:
            r = mcv.match_knn2_batch(q, qo, t, to, 0.8)
        finally:
            lib.mcvSetStagingMode(0)
        np.testing.assert_array_equal(r.train_idx, ref.train_idx)
        np.testing.assert_array_equal(r.distance, ref.distance)
    finally:
        mcv.set_hamming_engine(mcv.ENGINE_AUTO)


def test_host_register_roundtrip(lib):
    a = np.zeros(1 << 20, np.uint8)
    assert lib.mcvHostRegister(a.ctypes.data, a.nbytes) == 0, native.last_error()
    assert lib.mcvHostUnregister(a.ctypes.data) == 0
    assert lib.mcvHostRegister(0, 16) == -1


def test_copy_bandwidth_probe(lib):
    pinned = lib.mcvMeasureCopyBandwidth(0, 64 << 20, 4, 0, 0)
    ring = lib.mcvMeasureCopyBandwidth(0, 64 << 20, 4, 0, 1)
    plain = lib.mcvMeasureCopyBandwidth(0, 64 << 20, 4, 0, 2)
    d2h = lib.mcvMeasureCopyBandwidth(0, 64 << 20, 4, 1, 0)
    assert pinned > 5 and d2h > 5 and ring > 1 and plain > 0.5
    assert lib.mcvMeasureCopyBandwidth(0, 0, 1, 0, 0) < 0
    assert lib.mcvCopyThreads() >= 1


def test_device_exports_on_two_streams_do_not_corrupt_each_other():
    """two un-synchronised asynchronous calls from one thread on different streams share scratch.
    The lease orders the second call behind the first on the device: both results equal the serial ones."""
    import torch
    dev = torch.device("cuda:0")
    s1, s2 = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
    prm = []
    for k in range(2):
        a, b, _, _ = synth.homography_pair(90 + k, 1500)
        P = 24
        A = np.tile(a, (P, 1)); B = np.tile(b, (P, 1))
        off = np.arange(P + 1, dtype=np.int64) * a.shape[0]
        ref = est.ransac_batch(est.MODEL_H, A, B, off, 3.0, 1500, 0.995)
        prm.append((A, B, off, ref))
    outs = []
    for k, s in enumerate((s1, s2)):
        A, B, off, _ = prm[k]
        with torch.cuda.stream(s):
            dA = torch.from_numpy(A).to(dev, non_blocking=False); dB = torch.from_numpy(B).to(dev)
            P = off.size - 1
            dM = torch.zeros((P, 9), dtype=torch.float64, device=dev)
            dK = torch.zeros(A.shape[0], dtype=torch.uint8, device=dev)
            dC = torch.zeros(P, dtype=torch.int32, device=dev)
            dO = torch.zeros(P, dtype=torch.uint8, device=dev)
        outs.append((dA, dB, dM, dK, dC, dO))
    torch.cuda.synchronize()
    for rep in range(3):
        for k, s in enumerate((s1, s2)):          # back to back, no synchronisation in between
            A, B, off, _ = prm[k]
            dA, dB, dM, dK, dC, dO = outs[k]
            est.ransac_batch_device(est.MODEL_H, dA, dB, off, 3.0, 1500, dM, dK, dC, dO, confidence=0.995,
                                    sample_mode=est.SAMPLE_MWC, stop_mode=est.STOP_ADAPTIVE, stream=s.cuda_stream)
    torch.cuda.synchronize()
    for k in range(2):
        ref = prm[k][3]
        dA, dB, dM, dK, dC, dO = outs[k]
        np.testing.assert_array_equal(dC.cpu().numpy(), ref.inliers)
        np.testing.assert_array_equal(dK.cpu().numpy(), ref.masks)
        np.testing.assert_array_equal(dM.cpu().numpy().reshape(-1, 3, 3), ref.models)


def test_multi_chunk_tensor_engine_vs_popc_and_oracle():
    """>= 3 chunks of the persistent tcgen05 matcher (96 MB of expanded train rows per chunk = 384 C2 pairs):
    the whole batch equals the integer-pipe engine, and sampled pairs (first / last of every chunk, the
    chunk seams) equal the oracle."""
    P, n = 800, 8192
    rng = np.random.default_rng(5)
    base_q, base_t, _ = synth.descriptor_pair(123, n, n, 32)
    q = np.empty((P * n, 32), np.uint8)
    t = np.empty((P * n, 32), np.uint8)
    for p in range(P):                     # cheap distinct pairs: per-pair byte permutation + xor mask of one base pair
        perm = rng.permutation(32)
        mask = rng.integers(0, 256, 32, dtype=np.uint8)
        q[p * n:(p + 1) * n] = base_q[:, perm] ^ mask
        t[p * n:(p + 1) * n] = np.roll(base_t[:, perm] ^ mask, p * 37, axis=0)
    off = np.arange(P + 1, dtype=np.int64) * n
    mcv.set_hamming_engine(mcv.ENGINE_TENSOR)
    try:
        r_tc = mcv.match_knn2_batch(q, off, t, off, 0.8)
        mcv.set_hamming_engine(mcv.ENGINE_POPC)
        r_pc = mcv.match_knn2_batch(q, off, t, off, 0.8)
    finally:
        mcv.set_hamming_engine(mcv.ENGINE_AUTO)
    np.testing.assert_array_equal(r_tc.train_idx, r_pc.train_idx)
    np.testing.assert_array_equal(r_tc.distance, r_pc.distance)
    np.testing.assert_array_equal(r_tc.ratio_pass, r_pc.ratio_pass)
    assert r_tc.n_pass == r_pc.n_pass
    for p in (0, 1, 383, 384, 385, 767, 768, 799):
        idx, dist = bf_knn.bf_knn2_hamming(q[off[p]:off[p + 1]], t[off[p]:off[p + 1]])
        np.testing.assert_array_equal(r_tc.train_idx[off[p]:off[p + 1]], idx)
        np.testing.assert_array_equal(r_tc.distance[off[p]:off[p + 1]], dist)


@pytest.mark.parametrize("sample_mode", ["mwc", "counter"])
def test_host_buffer_batches_are_pipelined_in_runs_of_pairs(sample_mode):
    """A host-buffer batch in pageable memory above 8 MB of points is estimated in up to four runs of whole pairs, run k + 1 travelling while
    run k is estimated.  Pairs are independent (OpenCV's stream restarts per pair, the counter-based streams are keyed by
    the global pair index), so the result must be the device-resident, single-run one -- ragged pair sizes included."""
    import torch
    from minicv_b200 import estimation as est, synth
    rng = np.random.default_rng(5)
    sizes = rng.integers(900, 2600, 170).tolist()
    offs = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)
    assert offs[-1] * 32 > (8 << 20)
    A = np.empty((offs[-1], 2)); B = np.empty((offs[-1], 2))
    for p, n in enumerate(sizes):
        r = synth.homography_pair(600 + p, int(n))
        A[offs[p]:offs[p + 1]] = r[0]; B[offs[p]:offs[p + 1]] = r[1]
    sm = est.SAMPLE_MWC if sample_mode == "mwc" else est.SAMPLE_COUNTER
    host = est.ransac_batch(est.MODEL_H, A, B, offs, 3.0, 256, 0.995, sample_mode=sm, stop_mode=est.STOP_ADAPTIVE, seed=11)
    dev = torch.device("cuda:0")
    dA, dB = torch.from_numpy(A).to(dev), torch.from_numpy(B).to(dev)
    dM = torch.empty((len(sizes), 9), dtype=torch.float64, device=dev)
    dMask = torch.empty(int(offs[-1]), dtype=torch.uint8, device=dev)
    dCnt = torch.empty(len(sizes), dtype=torch.int32, device=dev); dOk = torch.empty(len(sizes), dtype=torch.uint8, device=dev)
    est.ransac_batch_device(est.MODEL_H, dA, dB, offs, 3.0, 256, dM, dMask, dCnt, dOk, confidence=0.995, sample_mode=sm,
                            stop_mode=est.STOP_ADAPTIVE, seed=11, stream=torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    np.testing.assert_array_equal(host.models.reshape(-1, 9), dM.cpu().numpy())
    np.testing.assert_array_equal(host.masks, dMask.cpu().numpy())
    np.testing.assert_array_equal(host.inliers, dCnt.cpu().numpy())
    np.testing.assert_array_equal(host.ok, dOk.cpu().numpy())
    assert host.ok.all()


def test_recover_pose_host_batch_equals_device_resident():
    import torch
    from minicv_b200 import estimation as est, synth
    n_pairs, n = 96, 3000
    A = np.empty((n_pairs * n, 2)); B = np.empty((n_pairs * n, 2))
    for p in range(n_pairs):
        r = synth.two_view_pair(700 + p, n)
        A[p * n:(p + 1) * n] = r[0]; B[p * n:(p + 1) * n] = r[1]
    off = np.arange(n_pairs + 1, dtype=np.int64) * n
    assert off[-1] * 32 > (8 << 20)
    cfg = est.RecoverPoseConfig(800.0, (320.0, 240.0), 0.999, 1.0)
    R, t, ms, good, ok = est.recover_pose_batch(cfg, A, B, off, maxIters=300)
    dev = torch.device("cuda:0")
    dA, dB = torch.from_numpy(A).to(dev), torch.from_numpy(B).to(dev)
    dE = torch.empty((n_pairs, 9), dtype=torch.float64, device=dev); dR = torch.empty((n_pairs, 9), dtype=torch.float64, device=dev)
    dT = torch.empty((n_pairs, 3), dtype=torch.float64, device=dev); dMs = torch.empty(n_pairs * n, dtype=torch.uint8, device=dev)
    dGood = torch.empty(n_pairs, dtype=torch.int32, device=dev); dOk = torch.empty(n_pairs, dtype=torch.uint8, device=dev)
    est.recover_pose_batch_device(cfg, dA, dB, off, dE, dR, dT, dMs, dGood, dOk, maxIters=300,
                                  stream=torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    np.testing.assert_array_equal(R.reshape(-1, 9), dR.cpu().numpy())
    np.testing.assert_array_equal(t, dT.cpu().numpy())
    np.testing.assert_array_equal(ms, dMs.cpu().numpy())
    np.testing.assert_array_equal(good, dGood.cpu().numpy())
    assert ok.all()
