#!/usr/bin/env python
"""bench.py -- headline benchmark of the match + two-view-estimation hot path.

  python bench.py --gpus N --steps K --warmup W            (ours; torchrun for N > 1)
  python bench.py --impl reference --gpus N --steps K ...  (OpenCV CPU path on the host cores)

Workload (BASELINE.json configs[1], "C2"): brute-force Hamming kNN k=2 + ratio 0.8 over
256-bit descriptors, 8192 x 8192 per image pair, 1024 pairs per GPU (weak scaling: image
pairs are independent units, sharded over ranks with no data-path collective).  One step =
one pass of the matcher over the rank's whole batch.  `value` is device-resident throughput
(inputs already in HBM); `e2e` goes through the host-buffer C-ABI call with pinned host
buffers, host<->device copies inside the timed region.  Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

METRIC = "image pairs/sec (kNN match + H/F/E RANSAC); Hamming Gdist/s vs roofline"
UNIT = "pairs/s"
NQ = NT = 8192
DESC_BYTES = 32
RATIO = 0.8


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--pairs", type=int, default=1024, help="image pairs per GPU per step")
    ap.add_argument("--cpu-sample-pairs", type=int, default=96)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extra", action="store_true", help="skip the C3/C4 RANSAC side measurements")
    ap.add_argument("--engine", default="auto", choices=["auto", "popc", "tensor", "packed"],
                    help="Hamming engine: auto (tensor cores for large batches), popc, tensor")
    return ap.parse_args()


# ---------------------------------------------------------------------------------------------
# clocks: sample nvidia-smi DURING the timed region (B200_PROFILING.md recipe)
# ---------------------------------------------------------------------------------------------
class ClockSampler:
    """Polls NVML (SM clock, max SM clock, power, throttle reasons) every few milliseconds from a
    thread; start() before the timed region, mark_begin()/mark_end() bracket it, and only samples
    inside the bracket are reported.  Falls back to `nvidia-smi -lms` when pynvml is missing."""
    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown",
               0x4: "sw_power_cap", 0x80: "hw_power_brake_slowdown"}

    def __init__(self, gpu_index: int, uuid: str | None = None):
        self.gpu = gpu_index
        self.uuid = uuid
        self.samples: list[tuple] = []       # (t, sm_mhz, power_w, reasons_mask)
        self.max_mhz = None
        self.t0 = self.t1 = None
        self._stop = threading.Event()
        self.thread = None
        self.mode = None
        self.proc = None
        self.lines: list[str] = []

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            h = None
            if self.uuid:
                try:
                    h = pynvml.nvmlDeviceGetHandleByUUID(self.uuid if self.uuid.startswith("GPU-") else "GPU-" + self.uuid)
                except Exception:
                    h = None
            if h is None:
                h = pynvml.nvmlDeviceGetHandleByIndex(self.gpu)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM))
            self.mode = "nvml"

            def poll():
                while not self._stop.is_set():
                    try:
                        mhz = pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)
                        pw = pynvml.nvmlDeviceGetPowerUsage(h) / 1000.0
                        rs = pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                        self.samples.append((time.perf_counter(), float(mhz), pw, int(rs)))
                    except Exception:
                        pass
                    time.sleep(0.004)
            self.thread = threading.Thread(target=poll, daemon=True)
            self.thread.start()
            return
        except Exception:
            self.mode = None
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "--query-gpu=clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
                 "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
                 "clocks_event_reasons.sw_power_cap", "--format=csv,noheader,nounits", "-lms", "50",
                 "-i", str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.mode = "smi"

            def pump():
                for line in self.proc.stdout:
                    self.lines.append((time.perf_counter(), line.strip()))
            self.thread = threading.Thread(target=pump, daemon=True)
            self.thread.start()
        except Exception:
            self.mode = None

    def mark_begin(self):
        self.t0 = time.perf_counter()

    def mark_end(self):
        self.t1 = time.perf_counter()

    def stop(self) -> dict:
        self._stop.set()
        if self.proc is not None:
            self.proc.terminate()
        if self.mode is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no clock source (pynvml / nvidia-smi unavailable)"]}
        t0 = self.t0 if self.t0 is not None else 0.0
        t1 = self.t1 if self.t1 is not None else float("inf")
        sm, pw, reasons = [], [], set()
        if self.mode == "nvml":
            for (t, mhz, p, rs) in self.samples:
                if t0 <= t <= t1:
                    sm.append(mhz); pw.append(p)
                    for bit, name in self.REASONS.items():
                        if rs & bit:
                            reasons.add(name)
            mx = self.max_mhz
        else:
            mxs = []
            names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
            for (t, ln) in self.lines:
                f = [x.strip() for x in ln.split(",")]
                if len(f) < 7 or not (t0 <= t <= t1):
                    continue
                try:
                    sm.append(float(f[0])); mxs.append(float(f[1])); pw.append(float(f[2]))
                except ValueError:
                    continue
                for name, v in zip(names, f[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            mx = max(mxs) if mxs else None
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": mx, "reasons": ["no samples inside the timed region"],
                    "source": self.mode}
        return {"sm_mhz": float(np.median(sm)), "sm_min_mhz": float(min(sm)), "sm_max_mhz": mx,
                "power_w_max": float(max(pw)), "samples": len(sm), "reasons": sorted(reasons), "source": self.mode}


# ---------------------------------------------------------------------------------------------
# CPU arm: the OpenCV calls the reference shim forwards to, via cv2 (same OpenCV, the only one
# in the image); falls back to the numpy oracle port when cv2 is not importable.
# ---------------------------------------------------------------------------------------------
def cpu_match_pairs(pairs, threads: int):
    """Time the CPU matcher over a list of (query, train) pairs.  Returns (seconds, kind)."""
    try:
        import cv2
        cv2.setNumThreads(threads)
        bf = cv2.BFMatcher(cv2.NORM_HAMMING)
        bf.knnMatch(pairs[0][0][:256], pairs[0][1][:256], 2)
        t0 = time.perf_counter()
        for q, t in pairs:
            m = bf.knnMatch(q, t, 2)
            _ = [1 for a in m if len(a) == 2 and a[0].distance < RATIO * a[1].distance]
        return time.perf_counter() - t0, "reference"
    except ImportError:
        from oracle import bf_knn
        t0 = time.perf_counter()
        for q, t in pairs:
            idx, dist = bf_knn.bf_knn2_hamming(q, t)
            bf_knn.ratio_test(dist, idx, RATIO)
        return time.perf_counter() - t0, "port"


# ---------------------------------------------------------------------------------------------
# host-buffer path (what a caller on the reference's side of the C ABI gets): every config of
# BASELINE.json through the host-pointer exports, copies inside the timed region, on every rank
# (weak scaling), with CUDA-pinned and with ordinary pageable caller buffers.
# ---------------------------------------------------------------------------------------------
def _pinned_like(arrs):
    """Copies of numpy arrays in CUDA-pinned host memory (torch as the allocator; plumbing only)."""
    import torch
    out = []
    for a in arrs:
        t = torch.empty(a.shape, dtype=torch.from_numpy(a[:0]).dtype).pin_memory()
        t.numpy()[...] = a
        out.append(t)
    return out


def _timed_host(fn, barrier, reduce_max, reps=3):
    """Seconds per synchronous host-API call: one warm-up, then `reps` calls timed together on the host clock
    (the exports return when their results are in the caller's buffers); max over ranks."""
    fn()
    barrier()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    return reduce_max((time.perf_counter() - t0) / reps)


def host_path_configs(lib, local, rank, world, barrier, reduce_max, with_cpu):
    """Returns (e2e others, cpu others, device-resident figures [rank 0], stage rows [rank 0])."""
    import torch
    import minicv_b200 as mcv
    from minicv_b200 import estimation as est, synth
    cores = os.cpu_count() or 1
    out, cpu, devres, kernels = {}, {}, {}, []
    dev = torch.device("cuda", local)
    stream = torch.cuda.current_stream()

    def timed_dev(fn, reps=3):
        fn(); fn()
        torch.cuda.synchronize()
        ts = []
        for _ in range(max(reps, 3)):
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record(stream); fn(); e1.record(stream)
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        return float(np.median(ts)) / 1e3

    def both(name, arrays, call, pairs, extra):
        """call(arrays) with pageable numpy arrays, then with pinned copies."""
        t_page = _timed_host(lambda: call(arrays), barrier, reduce_max)
        pins = _pinned_like(arrays)
        parr = [t.numpy() for t in pins]
        t_pin = _timed_host(lambda: call(parr), barrier, reduce_max)
        del pins, parr
        h2d = int(sum(a.nbytes for a in arrays))
        out[name] = dict({"value": world * pairs / t_pin, "unit": UNIT, "pageable": world * pairs / t_page,
                          "ms_per_batch": t_pin * 1e3, "pairs_per_gpu": pairs, "h2d_bytes_per_step": h2d}, **extra)

    # ---- C3: homography 4-pt and fundamental 7-pt, 2000 correspondences, 4096 hypotheses, 1024 pairs
    P, N, HYP = 1024, 2000, 4096
    off = np.arange(P + 1, dtype=np.int64) * N
    for kind, model, gen, thr, conf in (("H", est.MODEL_H, synth.homography_pair, 3.0, 0.995),
                                        ("F", est.MODEL_F, synth.two_view_pair, 3.0, 0.99)):
        A = np.empty((P * N, 2)); B = np.empty((P * N, 2))
        for p in range(P):
            r = gen(rank * P + p, N)
            A[p * N:(p + 1) * N] = r[0]; B[p * N:(p + 1) * N] = r[1]
        res = {}

        def call(arr, model=model, thr=thr, conf=conf):
            res["r"] = est.ransac_batch(model, arr[0], arr[1], off, thr, HYP, conf, sample_mode=est.SAMPLE_COUNTER,
                                        stop_mode=est.STOP_FIXED, seed=1, device=local)
        both(f"C3_{kind}", [A, B], call, P,
             {"api": "cvRansacBatch (host buffers): points up, models + masks + counts back",
              "correspondences": N, "hypotheses_per_pair": HYP, "mode": "counter-based sampling, fixed hypothesis count"})
        out[f"C3_{kind}"]["ok_pairs"] = int(res["r"].ok.sum())
        out[f"C3_{kind}"]["d2h_bytes_per_step"] = int(res["r"].models.nbytes + res["r"].masks.nbytes + res["r"].inliers.nbytes + P * 13)
        if rank == 0:
            dA = torch.from_numpy(A).to(dev); dB = torch.from_numpy(B).to(dev)
            dM = torch.empty((P, 9), dtype=torch.float64, device=dev); dMask = torch.empty(P * N, dtype=torch.uint8, device=dev)
            dCnt = torch.empty(P, dtype=torch.int32, device=dev); dOk = torch.empty(P, dtype=torch.uint8, device=dev)

            def run(model=model, thr=thr, conf=conf):
                est.ransac_batch_device(model, dA, dB, off, thr, HYP, dM, dMask, dCnt, dOk, confidence=conf,
                                        sample_mode=est.SAMPLE_COUNTER, stop_mode=est.STOP_FIXED, seed=1,
                                        device=local, stream=stream.cuda_stream)
            t = timed_dev(run)
            devres[f"C3_{kind}"] = {"pairs_per_s": P / t, "ms_per_batch": t * 1e3, "pairs": P,
                                    "ghyp_x_corr_per_s": P * N * HYP / t / 1e9, "ok_pairs": int(dOk.sum().item())}
            rows, _ = stage_breakdown(lib, local, f"C3_{kind}", run,
                                      {14: {"gevals_per_s": P * N * HYP * (1 if kind == "H" else 2.7) / 1e9},
                                       12: {"ghyp_per_s": P * HYP / 1e9}})
            kernels += rows
            del dA, dB, dM, dMask
        if with_cpu and rank == 0:
            k = max(cores, 8) if kind == "H" else cores
            cpu[f"C3_{kind}"] = cpu_ransac_pairs_per_s(kind, [(A[p * N:(p + 1) * N], B[p * N:(p + 1) * N]) for p in range(k)], cores)
        barrier()
        del A, B
    # ---- C4: essential 5-pt RANSAC + recoverPose, 5000 correspondences, 512 pairs (cvRecoverPose semantics)
    P, N = 512, 5000
    off4 = np.arange(P + 1, dtype=np.int64) * N
    A = np.empty((P * N, 2)); B = np.empty((P * N, 2))
    for p in range(P):
        r = synth.two_view_pair(rank * P + p, N)
        A[p * N:(p + 1) * N] = r[0]; B[p * N:(p + 1) * N] = r[1]
    cfg = est.RecoverPoseConfig(800.0, (320.0, 240.0), 0.999, 1.0)
    res = {}

    def call4(arr):
        res["r"] = est.recover_pose_batch(cfg, arr[0], arr[1], off4, device=local)
    both("C4_E_pose", [A, B], call4, P,
         {"api": "cvRecoverPoseBatch (host buffers): points up, R, t, RANSAC masks, cheirality counts back",
          "correspondences": N, "max_iters": 1000, "mode": "OpenCV MWC sampling + adaptive stop (cvRecoverPose semantics)"})
    out["C4_E_pose"]["ok_pairs"] = int(res["r"][4].sum())
    out["C4_E_pose"]["mean_cheirality_inliers"] = float(res["r"][3].mean())
    if rank == 0:
        dA = torch.from_numpy(A).to(dev); dB = torch.from_numpy(B).to(dev)
        dE = torch.empty((P, 9), dtype=torch.float64, device=dev); dR = torch.empty((P, 9), dtype=torch.float64, device=dev)
        dT = torch.empty((P, 3), dtype=torch.float64, device=dev); dMs = torch.empty(P * N, dtype=torch.uint8, device=dev)
        dGood = torch.empty(P, dtype=torch.int32, device=dev); dOk = torch.empty(P, dtype=torch.uint8, device=dev)

        def run_e():
            est.recover_pose_batch_device(cfg, dA, dB, off4, dE, dR, dT, dMs, dGood, dOk, sample_mode=est.SAMPLE_MWC,
                                          maxIters=1000, device=local, stream=stream.cuda_stream)
        t = timed_dev(run_e)
        devres["C4_E_pose"] = {"pairs_per_s": P / t, "ms_per_batch": t * 1e3, "pairs": P, "ok_pairs": int(dOk.sum().item())}
        rows, _ = stage_breakdown(lib, local, "C4", run_e, {12: {"ghyp_per_s": P * 1000 / 1e9}})
        kernels += rows
        del dA, dB, dMs
    if with_cpu and rank == 0:
        cpu["C4_E_pose"] = cpu_ransac_pairs_per_s("E", [(A[p * N:(p + 1) * N], B[p * N:(p + 1) * N]) for p in range(cores)], cores)
    barrier()
    del A, B
    # ---- match + homography RANSAC (the north-star pipeline), one call, 256 pairs of 8192 x 8192
    Pp = 256
    Q = np.empty((Pp * NQ, DESC_BYTES), np.uint8); T = np.empty((Pp * NT, DESC_BYTES), np.uint8)
    KQ = np.empty((Pp * NQ, 2)); KT = np.empty((Pp * NT, 2))
    for p in range(Pp):
        q, t, gt, kq, kt, H = synth.descriptor_pair_geo(rank * Pp + p)
        Q[p * NQ:(p + 1) * NQ] = q; T[p * NT:(p + 1) * NT] = t; KQ[p * NQ:(p + 1) * NQ] = kq; KT[p * NT:(p + 1) * NT] = kt
    qo = np.arange(Pp + 1, dtype=np.int64) * NQ; to = np.arange(Pp + 1, dtype=np.int64) * NT
    res = {}

    def callp(arr):
        res["r"] = est.match_homography_batch(arr[0], qo, arr[1], to, arr[2], arr[3], RATIO, 3.0, 2000, 0.995, device=local)
    both("match_plus_H", [Q, T, KQ, KT], callp, Pp,
         {"api": "cvMatchHomographyBatch (host buffers): descriptors + keypoints up, H + counts back; kNN-2 + ratio -> "
                 "device hand-off -> findHomography semantics (OpenCV MWC sampling, maxIters 2000, conf .995, refit + LM)",
          "n_query": NQ, "n_train": NT})
    out["match_plus_H"]["ok_pairs"] = int(res["r"].ok.sum())
    out["match_plus_H"]["mean_matches"] = float(res["r"].n_matches.mean())
    out["match_plus_H"]["mean_inliers"] = float(res["r"].n_inliers.mean())
    if rank == 0:
        dq = torch.from_numpy(Q).to(dev); dt = torch.from_numpy(T).to(dev)
        dkq = torch.from_numpy(KQ).to(dev); dkt = torch.from_numpy(KT).to(dev)
        didx = torch.empty((Pp * NQ, 2), dtype=torch.int32, device=dev); ddist = torch.empty((Pp * NQ, 2), dtype=torch.float32, device=dev)
        dok = torch.empty(Pp * NQ, dtype=torch.uint8, device=dev)
        da = torch.empty((Pp * NQ, 2), dtype=torch.float64, device=dev); db = torch.empty((Pp * NQ, 2), dtype=torch.float64, device=dev)
        dM = torch.empty((Pp, 9), dtype=torch.float64, device=dev); dMask = torch.empty(Pp * NQ, dtype=torch.uint8, device=dev)
        dCnt = torch.empty(Pp, dtype=torch.int32, device=dev); dOkp = torch.empty(Pp, dtype=torch.uint8, device=dev)

        def run_pipe():
            mcv.match_knn2_batch_device(dq, qo, dt, to, DESC_BYTES, RATIO, didx, ddist, dok, device=local, stream=stream.cuda_stream)
            off2 = mcv.gather_matches_device(dkq, dkt, qo, to, didx, dok, da, db, device=local, stream=stream.cuda_stream)
            est.ransac_batch_device(est.MODEL_H, da, db, off2, 3.0, 2000, dM, dMask, dCnt, dOkp, confidence=0.995,
                                    sample_mode=est.SAMPLE_MWC, stop_mode=est.STOP_ADAPTIVE, device=local, stream=stream.cuda_stream)
        t = timed_dev(run_pipe)
        devres["match_plus_H"] = {"pairs_per_s": Pp / t, "ms_per_batch": t * 1e3, "pairs": Pp, "ok_pairs": int(dOkp.sum().item())}
        rows, _ = stage_breakdown(lib, local, "match_plus_H", run_pipe)
        kernels += rows
        del dq, dt, dkq, dkt, didx, ddist, dok, da, db, dMask
    if with_cpu and rank == 0:
        try:
            import cv2
            cv2.setNumThreads(cores)
            bf = cv2.BFMatcher(cv2.NORM_HAMMING)
            k = 12
            t0 = time.perf_counter()
            for p in range(k):
                m = bf.knnMatch(Q[p * NQ:(p + 1) * NQ], T[p * NT:(p + 1) * NT], 2)
                good = [(a.queryIdx, a.trainIdx) for a, b_ in m if a.distance < RATIO * b_.distance]
                gi = np.array(good)
                cv2.findHomography(KQ[p * NQ:(p + 1) * NQ][gi[:, 0]], KT[p * NT:(p + 1) * NT][gi[:, 1]], cv2.RANSAC, 3.0,
                                   maxIters=2000, confidence=0.995)
            dtc = time.perf_counter() - t0
            cpu["match_plus_H"] = {"pairs_per_s": k / dtc, "cores": cores, "sample_pairs": k,
                                   "note": "cv2 knnMatch (all cores) + ratio + findHomography per pair"}
        except ImportError:
            pass
    barrier()
    for name, c in cpu.items():
        if c:
            out[name]["speedup_vs_cpu"] = out[name]["value"] / world / c["pairs_per_s"] if world == 1 else None
            out[name]["speedup_vs_cpu_per_gpu"] = out[name]["value"] / world / c["pairs_per_s"]
    torch.cuda.empty_cache()
    return out, cpu, devres, kernels


def h2d_ceiling(lib, local, world, barrier, all_gather_floats):
    """The box's host<->device ceiling, measured with every rank copying AT THE SAME TIME: each rank allocates 64 MiB
    pinned + device blocks (mcvCopyProbeCreate), all ranks meet at a barrier, then every rank keeps its copy engines busy
    for 0.6 s of wall time (one cudaMemcpyAsync per copy, CUDA events around the whole window) -- host->device alone,
    and in the 4 : 1 up : down byte mix of the C2 host-buffer path."""
    import ctypes as C
    from minicv_b200 import native
    probe = lib.mcvCopyProbeCreate(local, 64 << 20)
    if not probe:
        raise RuntimeError("mcvCopyProbeCreate failed: " + native.last_error())
    out = (C.c_double * 2)()
    res = {}
    for name, direction in (("h2d_only", 0), ("c2_mix", 3), ("both_1to1", 2)):
        barrier()
        native.check(lib.mcvCopyProbeRun(probe, 0.6, direction, out), "mcvCopyProbeRun")
        res[name] = (all_gather_floats(out[0]), all_gather_floats(out[1]))
    lib.mcvCopyProbeDestroy(probe)
    barrier()
    page = lib.mcvMeasureCopyBandwidth(local, 256 << 20, 4, 0, 1)
    barrier()
    plain = lib.mcvMeasureCopyBandwidth(local, 256 << 20, 2, 0, 2)
    pages = all_gather_floats(page); plains = all_gather_floats(plain)
    return {"h2d_pinned_gbs_per_rank": res["h2d_only"][0], "h2d_pinned_gbs_aggregate": float(sum(res["h2d_only"][0])),
            "c2_mix_h2d_gbs_aggregate": float(sum(res["c2_mix"][0])), "c2_mix_d2h_gbs_aggregate": float(sum(res["c2_mix"][1])),
            "both_1to1_h2d_gbs_aggregate": float(sum(res["both_1to1"][0])), "both_1to1_d2h_gbs_aggregate": float(sum(res["both_1to1"][1])),
            "h2d_pageable_ring_gbs_aggregate": float(sum(pages)), "h2d_pageable_plain_gbs_aggregate": float(sum(plains)),
            "copy_threads": int(lib.mcvCopyThreads()),
            "how": "all ranks at once after a barrier: 0.6 s of back-to-back 64 MiB cudaMemcpyAsync from cudaMallocHost memory "
                   "per rank, CUDA events around the window (mcvCopyProbeRun); c2_mix = one block device->host for every four "
                   "host->device (the byte mix of the timed C2 call), both_1to1 = equal bytes each way"}


STAGE_NAMES = {0: "matcher", 1: "expand_bits_kernel", 2: "knn2_finalize_kernel", 3: "l2", 10: "prep_points_kernel",
               11: "sample", 12: "solve", 13: "compact_kernel", 14: "score_kernel", 15: "select_kernel", 16: "finalize",
               17: "pose (decompose + cheirality + select)", 18: "gather_kernel", 19: "scale_search"}
KERNEL_OF = {
    ("C2", 0): "hamming_tc_kernel<2,2,6>",
    ("C3_H", 11): "sample_counter_kernel<0>", ("C3_H", 12): "solve_kernel<0>", ("C3_H", 14): "score_tc_kernel<0>",
    ("C3_H", 16): "finalize_homography_kernel",
    ("C3_F", 11): "sample_counter_kernel<1>", ("C3_F", 12): "solve_kernel<1>", ("C3_F", 14): "score_tc_kernel<1>",
    ("C3_F", 16): "finalize_plain_kernel<1>",
    ("C4", 11): "sample_mwc_par_kernel<2>", ("C4", 12): "solve_e_stage1 + solve_e_roots + solve_e_models",
    ("C4", 14): "score_tc_kernel<2>", ("C4", 16): "finalize_plain_kernel<2>", ("C4", 17): "decompose + cheirality + pose_select",
}


def stage_breakdown(lib, local, config, run, units=None):
    """CUDA-event time of every instrumented stage of one device-resident step of `config` (events recorded by
    the library on the launching stream around each kernel / kernel group)."""
    import torch
    run(); torch.cuda.synchronize()
    lib.mcvKernelTimingStart(local, 4096)
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(torch.cuda.current_stream())
    run()
    e1.record(torch.cuda.current_stream())
    torch.cuda.synchronize()
    ms = np.zeros(4096, np.float32); tags = np.zeros(4096, np.int32)
    n = lib.mcvKernelTimingStopTagged(local, ms.ctypes.data, tags.ctypes.data, 4096)
    step = e0.elapsed_time(e1)
    rows = []
    for tag in sorted(set(tags[:n].tolist())):
        sel = tags[:n] == tag
        t = float(ms[:n][sel].sum())
        row = {"config": config, "stage": STAGE_NAMES.get(tag, str(tag)), "kernel": KERNEL_OF.get((config, tag), STAGE_NAMES.get(tag, str(tag))),
               "launch_groups": int(sel.sum()), "ms_per_step": t, "share_of_step": t / step if step > 0 else None}
        if units and tag in units:
            row.update({k: (v / (t / 1e3) if t > 0 else None) for k, v in units[tag].items()})
        rows.append(row)
    return rows, step



# ---------------------------------------------------------------------------------------------
# CPU arms of the RANSAC configs: the same OpenCV calls on the host cores (one pair per task,
# cv2.setNumThreads(1), a pool of `cores` workers -- OpenCV's RANSAC loops do not multi-thread); and the
# side measurements that stay in `extra` (488-bit Hamming, L2, PnP, scale search: device-resident)
# ---------------------------------------------------------------------------------------------
def _cpu_ransac_init():
    import cv2
    cv2.setNumThreads(1)


def _cpu_ransac_task(task):
    import cv2
    kind, a, b = task
    t0 = time.perf_counter()
    if kind == "H":
        cv2.findHomography(a, b, cv2.RANSAC, 3.0, maxIters=4096, confidence=0.995)
    elif kind == "F":
        cv2.findFundamentalMat(a, b, cv2.FM_RANSAC, 3.0, 0.99, 4096)
    elif kind == "P":          # a = world points, b = image points
        K = np.array([[800.0, 0, 320.0], [0, 800.0, 240.0], [0, 0, 1.0]])
        cv2.solvePnPRansac(a, b, K, np.zeros(4), iterationsCount=1000, reprojectionError=2.0, confidence=0.999,
                           flags=cv2.SOLVEPNP_AP3P)
    else:
        E, m = cv2.findEssentialMat(a, b, 800.0, (320.0, 240.0), cv2.RANSAC, 0.999, 1.0)
        cv2.recoverPose(E, a, b, focal=800.0, pp=(320.0, 240.0), mask=m)
    return time.perf_counter() - t0


def cpu_ransac_pairs_per_s(kind, pairs, cores):
    """pairs/s of the OpenCV CPU path over `pairs` with a pool of `cores` single-thread workers."""
    try:
        import cv2  # noqa: F401
    except ImportError:
        return None
    import multiprocessing as mp
    ctx = mp.get_context("spawn")
    with ctx.Pool(cores, initializer=_cpu_ransac_init) as pool:
        pool.map(_cpu_ransac_task, [(kind, a, b) for a, b in pairs[:cores]])     # warm-up / page-in
        t0 = time.perf_counter()
        per = pool.map(_cpu_ransac_task, [(kind, a, b) for a, b in pairs])
        wall = time.perf_counter() - t0
    return {"pairs_per_s": len(pairs) / wall, "ms_per_pair_one_core": 1e3 * float(np.mean(per)),
            "cores": cores, "sample_pairs": len(pairs)}


def extra_ransac(local, with_cpu=True):
    import torch
    from minicv_b200 import estimation as est, synth
    dev = torch.device("cuda", local)
    stream = torch.cuda.current_stream()
    cores = os.cpu_count() or 1
    out = {}

    def timed(fn, reps=3):
        """Seconds per call: two warm-up calls, then the median of `reps` (>= 3) individually event-timed calls -- the
        extras with host synchronisation inside a call (gather -> offsets) showed 4x outliers on a freshly leased box."""
        fn(); fn()
        torch.cuda.synchronize()
        reps = max(reps, 3)
        ts = []
        for _ in range(reps):
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            fn()
            e1.record(stream)
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        return float(np.median(ts)) / 1e3

    import minicv_b200 as mcv
    # ---- M1 at 488 bit (AKAZE-MLDB, 61-byte rows): both Hamming engines, 128 pairs of 8192 x 8192
    Pa = 128
    qa = np.empty((Pa * NQ, 61), np.uint8); ta = np.empty((Pa * NT, 61), np.uint8)
    for p in range(Pa):
        qq_, tt_, _ = synth.descriptor_pair(3000 + p, NQ, NT, 61)
        qa[p * NQ:(p + 1) * NQ] = qq_; ta[p * NT:(p + 1) * NT] = tt_
    oa = np.arange(Pa + 1, dtype=np.int64) * NQ
    dqa = torch.from_numpy(qa).to(dev); dta = torch.from_numpy(ta).to(dev)
    dia = torch.empty((Pa * NQ, 2), dtype=torch.int32, device=dev); dda = torch.empty((Pa * NQ, 2), dtype=torch.float32, device=dev)
    doa = torch.empty(Pa * NQ, dtype=torch.uint8, device=dev)
    ak = {}
    ref_idx = None
    for ename, eng in (("popc", 1), ("tcgen05_int8", 2)):
        mcv.set_hamming_engine(eng)
        ta_ = timed(lambda: mcv.match_knn2_batch_device(dqa, oa, dta, oa, 61, RATIO, dia, dda, doa, device=local,
                                                        stream=stream.cuda_stream))
        ak[ename] = {"pairs_per_s": Pa / ta_, "ms_per_batch": ta_ * 1e3, "gdist_per_s": Pa * NQ * NT / ta_ / 1e9}
        if ref_idx is None:
            ref_idx = dia.clone()
        else:
            ak["engines_identical"] = bool(torch.equal(ref_idx, dia))
    mcv.set_hamming_engine(0)
    ak["pairs"] = Pa; ak["shape"] = f"{NQ} x {NT} x 488 bit (61-byte rows)"
    ak["pairs_per_s"] = max(ak["popc"]["pairs_per_s"], ak["tcgen05_int8"]["pairs_per_s"])
    ak["ms_per_batch"] = min(ak["popc"]["ms_per_batch"], ak["tcgen05_int8"]["ms_per_batch"])
    out["M1_hamming_488bit"] = ak
    del dqa, dta, dia, dda, doa, ref_idx
    # ---- M2: L2 kNN for float descriptors (SIFT-shaped: 128-d, 2048 x 2048 per pair, 256 pairs)
    Pl, nl, dl = 256, 2048, 128
    rngl = np.random.default_rng(11)
    ql = rngl.integers(0, 160, (Pl * nl, dl)).astype(np.float32)
    tl = rngl.integers(0, 160, (Pl * nl, dl)).astype(np.float32)
    ql[::2] = tl[::2] + rngl.integers(-2, 3, (Pl * nl // 2, dl)).astype(np.float32)
    lo_ = np.arange(Pl + 1, dtype=np.int64) * nl
    dql = torch.from_numpy(ql).to(dev); dtl = torch.from_numpy(tl).to(dev)
    dil = torch.empty((Pl * nl, 2), dtype=torch.int32, device=dev); ddl = torch.empty((Pl * nl, 2), dtype=torch.float32, device=dev)
    dol = torch.empty(Pl * nl, dtype=torch.uint8, device=dev)
    l2 = {}
    for ename, eng in (("fp32_cuda_core", 1), ("tcgen05_bf16_split_rescore", 2)):
        mcv.set_l2_engine(eng)
        tl2 = timed(lambda: mcv.match_knn2_batch_device(dql, lo_, dtl, lo_, dl * 4, RATIO, dil, ddl, dol, norm=mcv.NORM_L2,
                                                        device=local, stream=stream.cuda_stream))
        l2[ename] = {"pairs_per_s": Pl / tl2, "ms_per_batch": tl2 * 1e3, "gdist_per_s": Pl * nl * nl / tl2 / 1e9,
                     "planted_found": int((dil[::2, 0] == torch.arange(0, Pl * nl, 2, device=dev) % nl).sum().item())}
    mcv.set_l2_engine(0)
    l2["pairs"] = Pl; l2["shape"] = f"{nl} x {nl} x {dl} f32"
    l2["pairs_per_s"] = max(l2["fp32_cuda_core"]["pairs_per_s"], l2["tcgen05_bf16_split_rescore"]["pairs_per_s"])
    l2["ms_per_batch"] = min(l2["fp32_cuda_core"]["ms_per_batch"], l2["tcgen05_bf16_split_rescore"]["ms_per_batch"])
    if with_cpu:
        try:
            import cv2
            cv2.setNumThreads(cores)
            bfl = cv2.BFMatcher(cv2.NORM_L2)
            t0 = time.perf_counter()
            for p in range(8):
                bfl.knnMatch(ql[p * nl:(p + 1) * nl], tl[p * nl:(p + 1) * nl], 2)
            l2["cpu"] = {"pairs_per_s": 8 / (time.perf_counter() - t0), "cores": cores, "sample_pairs": 8}
        except ImportError:
            pass
    out["M2_l2_knn"] = l2
    del dql, dtl, dil, ddl, dol
    # ---- single-pair calls through the reference-style exports (what one F# call costs)
    lat = {}
    qq, tt, _ = synth.descriptor_pair(1)
    a_h, b_h, _, _ = synth.homography_pair(1, 2000)
    a_e, b_e, _, _, _ = synth.two_view_pair(1, 5000)
    cfg1 = est.RecoverPoseConfig(800.0, (320.0, 240.0), 0.999, 1.0)
    for name, fn in (("cvMatchKnn2_8192x8192", lambda: mcv.match_knn2(qq, tt, RATIO)),
                     ("cvFindHomography_2000", lambda: est.findHomography(a_h, b_h, 3.0, 2000, 0.995)),
                     ("cvFindFundamentalMat_2000", lambda: est.findFundamentalMat(a_h, b_h, 3.0, 0.99, 1000)),
                     ("cvRecoverPose_5000", lambda: est.recoverPose(cfg1, a_e, b_e))):
        fn(); fn()
        t0 = time.perf_counter()
        for _ in range(5):
            fn()
        lat[name] = (time.perf_counter() - t0) / 5 * 1e3
    out["single_call_ms"] = lat
    # ---- PnP RANSAC: cvSolvePnPRansac(AP3P) semantics, 2000 correspondences, 30 % inliers
    P, N = 1024, 2000
    off = np.arange(P + 1, dtype=np.int64) * N
    X = np.empty((P * N, 3)); U = np.empty((P * N, 2))
    Kc = np.array([[800.0, 0, 320.0], [0, 800.0, 240.0], [0, 0, 1.0]])
    for p in range(P):
        X[p * N:(p + 1) * N], U[p * N:(p + 1) * N], _, _ = synth.pnp_scene(p, N, 0.3, noise=0.5, K=Kc)
    dX = torch.from_numpy(X).to(dev); dU = torch.from_numpy(U).to(dev)
    dC = torch.from_numpy(est.camera_rows(Kc, None, P)).to(dev)
    dRv = torch.empty((P, 3), dtype=torch.float64, device=dev); dTv = torch.empty((P, 3), dtype=torch.float64, device=dev)
    dMk = torch.empty(P * N, dtype=torch.uint8, device=dev)
    dCn = torch.empty(P, dtype=torch.int32, device=dev); dOk2 = torch.empty(P, dtype=torch.uint8, device=dev)
    for name, sm, stp, mode in (("PnP_AP3P_opencv_semantics", est.SAMPLE_MWC, est.STOP_ADAPTIVE,
                                 "OpenCV MWC sampling + adaptive stop + EPnP refit (cvSolvePnPRansac semantics)"),
                                ("PnP_AP3P_fixed_1000", est.SAMPLE_COUNTER, est.STOP_FIXED,
                                 "counter-based sampling, 1000 hypotheses scored per problem, EPnP refit")):
        def run_p():
            est.solve_pnp_ransac_batch_device(dX, dU, off, dC, 2.0, 1000, 0.999, dRv, dTv, dMk, dCn, dOk2, sample_mode=sm,
                                              stop_mode=stp, seed=3, device=local, stream=stream.cuda_stream)
        t = timed(run_p, reps=2)
        out[name] = {"problems_per_s": P / t, "ms_per_batch": t * 1e3, "problems": P, "correspondences": N,
                     "max_iters": 1000, "ok": int(dOk2.sum().item()), "mean_inliers": float(dCn.float().mean().item()),
                     "ghyp_x_corr_per_s": P * N * 1000 / t / 1e9 if stp == est.STOP_FIXED else None, "mode": mode}
    if with_cpu:
        out["PnP_AP3P_opencv_semantics"]["cpu"] = cpu_ransac_pairs_per_s(
            "P", [(X[p * N:(p + 1) * N], U[p * N:(p + 1) * N]) for p in range(cores)], cores)
    del dX, dU
    # ---- scale search: CameraPose.findScaled, 2000 observations -> 4000 candidate scales x
    #      2000 projections per problem
    from minicv_b200 import camera_pose as mcp
    P, N = 256, 2000
    off = np.arange(P + 1, dtype=np.int64) * N
    rows = np.empty((P, 14)); Rs = np.empty((P, 9)); Ts = np.empty((P, 3)); W = np.empty((P * N, 3)); O = np.empty((P * N, 2))
    planted = np.empty(P)
    for p in range(P):
        rows[p], R, Ts[p], W[p * N:(p + 1) * N], O[p * N:(p + 1) * N], planted[p] = synth.scale_scene(p, N)
        Rs[p] = R.ravel()
    dCam = torch.from_numpy(rows).to(dev); dRot = torch.from_numpy(Rs).to(dev); dTr = torch.from_numpy(Ts).to(dev)
    dW = torch.from_numpy(W).to(dev); dO = torch.from_numpy(O).to(dev)
    dCost = torch.empty(P, dtype=torch.float64, device=dev); dSc = torch.empty(P, dtype=torch.float64, device=dev)
    dST = torch.empty((P, 3), dtype=torch.float64, device=dev); dSg = torch.empty(P, dtype=torch.int32, device=dev)

    def run_s():
        mcp.find_scaled_batch_device(dCam, dRot, dTr, dW, dO, off, dCost, dSc, dST, dSg, device=local, stream=stream.cuda_stream)
    t = timed(run_s, reps=3)
    rel = np.abs(dSc.cpu().numpy() - planted) / planted
    out["findScaled"] = {"problems_per_s": P / t, "ms_per_batch": t * 1e3, "problems": P, "observations": N,
                         "candidates_per_problem": 2 * N, "geval_per_s": P * 2 * N * N / t / 1e9,
                         "planted_scale_recovered": int((rel < 0.01).sum()),
                         "mode": "CameraPose.findScaled semantics (every candidate costed against every observation, sums in order)"}
    if with_cpu:
        # no F# runtime and no OpenCV counterpart: the CPU figure is the numpy port (oracle/camera_pose.py), one core, one problem
        from oracle import camera_pose as ocp
        cam = ocp.Camera(rows[0, 0:3], rows[0, 3:6], rows[0, 6:9], rows[0, 9:12], rows[0, 12:14])
        pose = ocp.CameraPose(0, 1, Rs[0].reshape(3, 3), Ts[0], False)
        t0 = time.perf_counter()
        oc, osc = ocp.find_scaled_fast(cam, W[:N], O[:N], pose)[:2]
        dt = time.perf_counter() - t0
        out["findScaled"]["cpu"] = {"problems_per_s": 1.0 / dt, "cores": 1, "kind": "port (numpy restatement, vectorised)",
                                    "sample": "problem 0", "bit_identical_to_gpu": bool(oc == float(dCost[0].item()) and osc == float(dSc[0].item()))}
    return out


def extra_large_db(lib, local, rank, world, n_train=4 * 1024 * 1024, n_query=10000):
    """C5: train set sharded by contiguous range over the ranks, queries replicated, local top-2 keys gathered with
    ncclAllGather INSIDE the library (cvMatchKnn2ShardedDevice; the communicator is built from an id rank 0 obtains
    from the library and torch.distributed only carries those 128 bytes), merged on the device.  Time = max over ranks."""
    import ctypes as C
    import torch
    import torch.distributed as dist
    from minicv_b200 import native, sharding
    dev = torch.device("cuda", local)
    lo, hi = sharding.block_range(n_train, world, rank)
    g = torch.Generator(device=dev)
    g.manual_seed(7 + rank)
    shard = torch.randint(0, 256, (hi - lo, DESC_BYTES), dtype=torch.uint8, device=dev, generator=g)
    gq = torch.Generator(device=dev)
    gq.manual_seed(99)
    q = torch.randint(0, 256, (n_query, DESC_BYTES), dtype=torch.uint8, device=dev, generator=gq)
    if lo == 0:
        q[: n_query // 2] = shard[: n_query // 2] ^ 1       # half the queries have a near-duplicate in shard 0
    idt = torch.zeros(128, dtype=torch.uint8)
    if rank == 0:
        buf = (C.c_uint8 * 128)()
        native.check(lib.mcvShardCommUniqueId(buf), "mcvShardCommUniqueId")
        idt = torch.tensor(list(buf), dtype=torch.uint8)
    if world > 1:
        dist.broadcast(q, 0)
        idd = idt.to(dev)
        dist.broadcast(idd, 0)
        idt = idd.cpu()
    idb = (C.c_uint8 * 128)(*idt.tolist())
    comm = lib.mcvShardCommInit(idb, world, rank, local)
    if not comm:
        raise RuntimeError("mcvShardCommInit failed: " + native.last_error())
    stream = torch.cuda.current_stream()
    idx = torch.empty((n_query, 2), dtype=torch.int64, device=dev)
    dd = torch.empty((n_query, 2), dtype=torch.float32, device=dev)
    ok = torch.empty(n_query, dtype=torch.uint8, device=dev)

    def run():
        native.check(lib.cvMatchKnn2ShardedDevice(comm, q.data_ptr(), n_query, shard.data_ptr(), hi - lo, lo, DESC_BYTES, RATIO,
                                                  idx.data_ptr(), dd.data_ptr(), ok.data_ptr(), stream.cuda_stream),
                     "cvMatchKnn2ShardedDevice")
    run()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    reps = 5
    e0.record(stream)
    for _ in range(reps):
        run()
    e1.record(stream)
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / reps], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    found = int((idx[: n_query // 2, 0] == torch.arange(n_query // 2, device=dev)).sum().item())
    lib.mcvShardCommDestroy(comm)
    return {"ms_per_image": ms, "images_per_s": 1e3 / ms, "n_train_total": n_train, "n_query": n_query,
            "shards": world, "gdist_per_s": n_query * n_train / ms / 1e6,
            "planted_matches_found": found, "planted": n_query // 2, "nccl_version": int(lib.mcvNcclVersion()),
            "api": "cvMatchKnn2ShardedDevice (local keys -> ncclAllGather inside libMiniCVNative -> merge)",
            "collective": "ncclAllGather of Q x 2 x 8 B per rank, issued by the library" if world > 1 else "none (1 rank)"}


def single_call_multi_gpu(lib, n_dev, pairs_per_dev=256):
    """ONE cvMatchKnn2BatchMulti call from this process over every visible device (pinned host buffers): what a
    P/Invoke caller on the reference's side of the boundary gets from an 8-GPU box, no torch / torchrun in the data
    path.  Also the in-process large-database form (cvLargeDb*, ncclCommInitAll + grouped all-gather)."""
    import minicv_b200 as mcv
    from minicv_b200 import synth
    P = pairs_per_dev * n_dev
    pq = mcv.PinnedArray((P * NQ, DESC_BYTES), np.uint8); pt = mcv.PinnedArray((P * NT, DESC_BYTES), np.uint8)
    base_q, _, base_t, _ = synth.descriptor_batch(5000, 16, NQ, NT, DESC_BYTES)
    for k in range(P // 16):
        pq.array[k * 16 * NQ:(k + 1) * 16 * NQ] = np.roll(base_q, k, axis=1)
        pt.array[k * 16 * NT:(k + 1) * 16 * NT] = np.roll(base_t, k, axis=1)
    pi = mcv.PinnedArray((P * NQ, 2), np.int32); pd = mcv.PinnedArray((P * NQ, 2), np.float32); po = mcv.PinnedArray((P * NQ,), np.uint8)
    qo = np.arange(P + 1, dtype=np.int64) * NQ; to = np.arange(P + 1, dtype=np.int64) * NT
    out = mcv.KnnResult(pi.array, pd.array, po.array, 0)
    res = {}
    for nd in sorted({1, n_dev}):
        mask = (1 << nd) - 1
        Pn = pairs_per_dev * nd
        o2 = mcv.KnnResult(pi.array[:Pn * NQ], pd.array[:Pn * NQ], po.array[:Pn * NQ], 0)
        mcv.match_knn2_batch_multi(pq.array[:Pn * NQ], qo[:Pn + 1], pt.array[:Pn * NT], to[:Pn + 1], RATIO, device_mask=mask, out=o2)
        t0 = time.perf_counter()
        reps = 3
        for _ in range(reps):
            mcv.match_knn2_batch_multi(pq.array[:Pn * NQ], qo[:Pn + 1], pt.array[:Pn * NT], to[:Pn + 1], RATIO, device_mask=mask, out=o2)
        dt = (time.perf_counter() - t0) / reps
        res[f"{nd}_gpu"] = {"pairs_per_s": Pn / dt, "pairs": Pn, "ms_per_call": dt * 1e3, "n_pass": int(o2.n_pass)}
    res["api"] = "cvMatchKnn2BatchMulti(deviceMask): one C call, one host thread + stream set per device, pinned host buffers"
    if n_dev > 1:
        res["scaling_1_to_n"] = res[f"{n_dev}_gpu"]["pairs_per_s"] / res["1_gpu"]["pairs_per_s"]
    for x in (pq, pt, pi, pd, po):
        x.free()
    # in-process large database over every visible device
    rng = np.random.default_rng(7)
    n_train, n_query = 4 * 1024 * 1024, 10000
    train = rng.integers(0, 256, (n_train, DESC_BYTES), dtype=np.uint8)
    query = rng.integers(0, 256, (n_query, DESC_BYTES), dtype=np.uint8)
    query[: n_query // 2] = train[: n_query // 2] ^ 1
    db_res = {}
    for nd in sorted({1, n_dev}):
        with mcv.LargeDb(train, (1 << nd) - 1) as db:
            db.match(query, RATIO)
            t0 = time.perf_counter()
            for _ in range(5):
                idx, dist_, ok, npass = db.match(query, RATIO)
            dt = (time.perf_counter() - t0) / 5
            db_res[f"{nd}_gpu"] = {"ms_per_image": dt * 1e3, "shards": db.n_shards, "uses_nccl": db.uses_nccl,
                                   "planted_matches_found": int((idx[: n_query // 2, 0] == np.arange(n_query // 2)).sum())}
    db_res["api"] = "cvLargeDbCreate / cvLargeDbMatch: shards resident, queries from host memory, ncclCommInitAll + grouped ncclAllGather inside the library, result back in host memory"
    res["large_db"] = db_res
    return res


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from minicv_b200 import synth
    cores = os.cpu_count() or 1
    sample = max(1, min(args.cpu_sample_pairs // 4, args.pairs))
    pairs = [synth.descriptor_pair(p, NQ, NT, DESC_BYTES)[:2] for p in range(sample)]
    kind = "reference"
    for _ in range(args.warmup):
        cpu_match_pairs(pairs[:1], cores)
    tot = 0.0
    for _ in range(args.steps):
        dt, kind = cpu_match_pairs(pairs, cores)
        tot += dt
    ms = tot / args.steps * 1e3
    value = sample / (tot / args.steps)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": workload_config(args, sample_pairs=sample),
        "gdist_per_s": value * NQ * NT / 1e9,
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind,
                         "sample": f"{sample} of the {args.pairs} C2 pairs per step, cv2.BFMatcher(NORM_HAMMING)"
                                   f".knnMatch k=2 + ratio {RATIO}, cv2.setNumThreads({cores})"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def bind_to_gpu_numa_node(local: int):
    """One process per GPU: pin the rank's host threads (and therefore its first-touch pinned
    staging buffers) to the CPUs NVML reports as local to the GPU, so the e2e copies of 8 ranks do
    not all cross the socket interconnect.  Best effort; returns a short description or None."""
    try:
        import pynvml
        pynvml.nvmlInit()
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        idx = int(vis.split(",")[local]) if vis and vis.split(",")[local].isdigit() else local
        h = pynvml.nvmlDeviceGetHandleByIndex(idx)
        ncpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (ncpu + 63) // 64)
        cpus = [w * 64 + b for w, m in enumerate(words) for b in range(64) if (m >> b) & 1 and w * 64 + b < ncpu]
        if cpus:
            os.sched_setaffinity(0, cpus)
            return f"{len(cpus)} cpus local to gpu {idx}"
    except Exception:
        return None
    return None


def workload_config(args, sample_pairs=None):
    c = {"workload": "C2: Hamming kNN k=2 + ratio 0.8, 256-bit descriptors, 8192x8192 per pair",
         "pairs_per_gpu": args.pairs, "n_query": NQ, "n_train": NT, "desc_bytes": DESC_BYTES,
         "ratio": RATIO, "parallelism": f"pairs sharded over {args.gpus} rank(s), no collective",
         "cache": "inputs per step (pairs x 512 KiB) exceed the 126 MB L2 at the default batch; no flush"}
    if sample_pairs is not None:
        c["cpu_sample_pairs_per_step"] = sample_pairs
    return c


# ---------------------------------------------------------------------------------------------
# ours
# ---------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist
    import minicv_b200 as mcv
    from minicv_b200 import native, synth

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    lib = native.load()
    if lib.mcvDeviceCount() < 1:
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback): " + native.last_error())
    all_cpus = os.sched_getaffinity(0)
    numa = bind_to_gpu_numa_node(local)
    engine = {"auto": 0, "popc": 1, "tensor": 2, "packed": 3}[args.engine]
    mcv.set_hamming_engine(engine)
    use_tc = engine >= 2 or (engine == 0 and args.pairs * NQ * NT >= 64 * 1024 * 1024)
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    P = args.pairs
    # --- synthetic batch, unique pairs per rank, generated on the host then made resident
    hq = torch.empty((P * NQ, DESC_BYTES), dtype=torch.uint8).pin_memory()
    ht = torch.empty((P * NT, DESC_BYTES), dtype=torch.uint8).pin_memory()
    q_np, qo, t_np, to = synth.descriptor_batch(rank * P, P, NQ, NT, DESC_BYTES,
                                                out_query=hq.numpy(), out_train=ht.numpy())
    dq = hq.to(dev, non_blocking=True)
    dt = ht.to(dev, non_blocking=True)
    didx = torch.empty((P * NQ, 2), dtype=torch.int32, device=dev)
    ddist = torch.empty((P * NQ, 2), dtype=torch.float32, device=dev)
    dok = torch.empty(P * NQ, dtype=torch.uint8, device=dev)
    stream = torch.cuda.current_stream()

    def step_device():
        mcv.match_knn2_batch_device(dq, qo, dt, to, DESC_BYTES, RATIO, didx, ddist, dok,
                                    device=local, stream=stream.cuda_stream)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    try:
        uuid = str(torch.cuda.get_device_properties(dev).uuid)
    except Exception:
        uuid = None
    sampler = ClockSampler(local, uuid)
    if rank == 0:
        sampler.start()
    for _ in range(max(args.warmup, 3)):
        step_device()
    barrier()
    launches0 = lib.mcvKernelLaunchCount()
    # CUDA events around the dominant kernel alone, recorded by the library on the launching stream
    lib.mcvKernelTimingStart(local, (args.steps + 1) * 256)        # capacity in event pairs; ~57 per step at C2
    e0 = torch.cuda.Event(enable_timing=True)
    e1 = torch.cuda.Event(enable_timing=True)
    barrier()
    sampler.mark_begin()
    e0.record(stream)
    for _ in range(args.steps):
        step_device()
    e1.record(stream)
    barrier()
    sampler.mark_end()
    launches = lib.mcvKernelLaunchCount() - launches0
    ms_total = e0.elapsed_time(e1)
    kbuf = np.zeros((args.steps + 1) * 256, dtype=np.float32)
    ktag = np.zeros((args.steps + 1) * 256, dtype=np.int32)
    nk_all = lib.mcvKernelTimingStopTagged(local, kbuf.ctypes.data, ktag.ctypes.data, len(kbuf))
    # the matcher launches hamming_tc_kernel once per L2-sized chunk of pairs: sum per step
    sel0 = ktag[:nk_all] == 0
    nk = int(sel0.sum())
    kern_ms = float(kbuf[:nk_all][sel0].sum()) / args.steps if nk > 0 else None
    kern_launches_per_step = nk / args.steps if nk > 0 else None
    c2_rows = []
    for tag in sorted(set(ktag[:nk_all].tolist())):
        tms = float(kbuf[:nk_all][ktag[:nk_all] == tag].sum()) / args.steps
        c2_rows.append({"config": "C2", "stage": STAGE_NAMES.get(tag, str(tag)),
                        "kernel": KERNEL_OF.get(("C2", tag), STAGE_NAMES.get(tag, str(tag))),
                        "launch_groups": int((ktag[:nk_all] == tag).sum() / args.steps), "ms_per_step": tms})
    clocks = sampler.stop() if rank == 0 else None
    tmax = torch.tensor([ms_total], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    ms_step = float(tmax.item()) / args.steps
    value = world * P / (ms_step / 1e3)
    n_pass_dev = int(dok.sum().item())

    # --- e2e: the host-buffer C-ABI call, host inputs and outputs, copies inside the timing.  Twice: with CUDA-pinned
    #     buffers (direct DMA) and with ordinary pageable arrays (what a P/Invoke caller passes; staged by the library)
    def reduce_max(x):
        te = torch.tensor([x], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
        return float(te.item())

    def all_gather_floats(x):
        te = torch.tensor([x], dtype=torch.float64, device=dev)
        if world > 1:
            outl = [torch.zeros_like(te) for _ in range(world)]
            dist.all_gather(outl, te)
            return [float(o.item()) for o in outl]
        return [float(x)]

    out = mcv.KnnResult(torch.empty((P * NQ, 2), dtype=torch.int32).pin_memory().numpy(),
                        torch.empty((P * NQ, 2), dtype=torch.float32).pin_memory().numpy(),
                        torch.empty(P * NQ, dtype=torch.uint8).pin_memory().numpy(), 0)
    e2e_steps = max(2, min(args.steps, 5))
    e2e_s = _timed_host(lambda: mcv.match_knn2_batch(q_np, qo, t_np, to, RATIO, device=local, out=out), barrier, reduce_max,
                        reps=e2e_steps)
    e2e_value = world * P / e2e_s
    assert out.n_pass == n_pass_dev, "host-path and device-path results differ"
    q_pg, t_pg = np.array(q_np, copy=True), np.array(t_np, copy=True)           # ordinary (pageable) memory
    out_pg = mcv.KnnResult(np.empty((P * NQ, 2), np.int32), np.empty((P * NQ, 2), np.float32), np.empty(P * NQ, np.uint8), 0)
    e2e_pg_s = _timed_host(lambda: mcv.match_knn2_batch(q_pg, qo, t_pg, to, RATIO, device=local, out=out_pg), barrier, reduce_max,
                           reps=e2e_steps)
    assert out_pg.n_pass == n_pass_dev and np.array_equal(out_pg.train_idx, out.train_idx)
    lib.mcvSetStagingMode(1)
    e2e_plain_s = _timed_host(lambda: mcv.match_knn2_batch(q_pg, qo, t_pg, to, RATIO, device=local, out=out_pg), barrier, reduce_max,
                              reps=2)
    lib.mcvSetStagingMode(0)
    del q_pg, t_pg, out_pg
    h2d_bytes = int(q_np.nbytes + t_np.nbytes)
    d2h_bytes = int(out.train_idx.nbytes + out.distance.nbytes + out.ratio_pass.nbytes)
    ceiling = h2d_ceiling(lib, local, world, barrier, all_gather_floats)
    e2e_h2d_gbs = world * h2d_bytes / e2e_s / 1e9

    # --- C5 (all ranks): 4 Mi train descriptors sharded over the ranks, 10 000 queries, keys all-gathered in the library
    c5 = None
    others = cpu_others = devres = krows = None
    if not args.no_extra:
        c5 = extra_large_db(lib, local, rank, world)
        del dq, dt, didx, ddist, dok
        torch.cuda.empty_cache()
        os.sched_setaffinity(0, all_cpus)      # CPU arms may use every core
        others, cpu_others, devres, krows = host_path_configs(lib, local, rank, world, barrier, reduce_max,
                                                              with_cpu=not args.no_cpu_baseline)
    # --- one C call driving every visible device (rank 0, the other ranks wait)
    multi = None
    if not args.no_extra and rank == 0 and world == 1 and lib.mcvDeviceCount() > 1:
        multi = single_call_multi_gpu(lib, lib.mcvDeviceCount())
    barrier()

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # --- roofline of the dominant kernel (hamming_knn2_kernel: 1 launch per step)
    peaks = {}
    pk = ROOT / "MEASURED_PEAKS.json"
    if pk.exists():
        peaks = json.loads(pk.read_text())
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured" if "hbm_gbs" in peaks else "fallback"
    alg_bytes = P * ((NQ + NT) * DESC_BYTES + NQ * 16 + NQ)
    kern_s = ms_step / 1e3
    gdist = world * P * NQ * NT / kern_s / 1e9
    sm_mhz = (clocks or {}).get("sm_mhz") or float(peaks.get("sm_max_mhz", 1965.0))
    sms = torch.cuda.get_device_properties(dev).multi_processor_count
    popc_per_dist = DESC_BYTES // 4
    pipe_peak = sms * 16 * sm_mhz * 1e6 / popc_per_dist / 1e9   # Gdist/s at 16 POPC/clk/SM
    int_pipe = {"achieved_gdist_s": gdist / world, "peak_gdist_s": pipe_peak, "frac": gdist / world / pipe_peak,
                "model": f"{sms} SMs x 16 POPC/clk/SM x {sm_mhz:.0f} MHz / {popc_per_dist} POPC per distance"}
    if use_tc:
        # tcgen05 int8 GEMM: 2*Q*T*K operations per pair and launch, over the duration of the GEMM + top-2
        # kernel alone (library-recorded CUDA events on the launching stream).  Denominator: the int8 rate
        # of tcgen05.mma.kind::i8 measured on this pool's B200 by tools/mma_bench.cu (8185 MAC/clk/SM at the
        # 1965 MHz this kernel runs at: profiles/r01_mma_bench.txt).  MEASURED_PEAKS.json carries no int8
        # figure; its cuBLAS bf16 number (power-capped to ~1.3 GHz) x2 is quoted beside it.
        bf16_peak = float(peaks.get("bf16_tflops", 1590.0))
        int8_peak = 2.0 * 8185.0 * sms * float(peaks.get("sm_max_mhz", 1965.0)) * 1e6 / 1e12
        ops = 2.0 * P * NQ * NT * (DESC_BYTES * 8)
        k_s = (kern_ms if kern_ms else ms_step) / 1e3
        ach = ops / k_s / 1e12
        traffic = None
        tj = ROOT / "profiles" / "r01_tc_traffic.json"
        if tj.exists() and P == 1024:
            traffic = json.loads(tj.read_text()).get("traffic_bytes_per_step")
        roofline = {
            "bound": "tensor", "achieved": ach, "peak": int8_peak, "unit": "TFLOP/s", "frac": ach / int8_peak,
            "traffic": traffic,
            "peak_source": "measured int8 tcgen05.mma issue rate, tools/mma_bench.cu (8185 MAC/clk/SM x SMs x sm_max_mhz)",
            "kernel": "hamming_tc_kernel<2,2,6> (tcgen05.mma.kind::i8, A in TMEM + packed 16-bit top-2 epilogue)",
            "kernel_ms": kern_ms, "kernel_launches_per_step": kern_launches_per_step,
            "kernel_share_of_step": (kern_ms / ms_step) if kern_ms else None,
            "whole_step": {"achieved": ops / kern_s / 1e12, "frac": ops / kern_s / 1e12 / int8_peak},
            "vs_2x_measured_bf16": ach / (2 * bf16_peak),
            "note": "achieved = 2*Q*T*K int8 operations of one step / summed duration of its hamming_tc_kernel launches "
                    "(one per L2-sized chunk of pairs; CUDA events recorded by the library around each launch on the "
                    "launching stream, inside the timed region); whole_step divides by the step time instead (train-bit expansion + this kernel + "
                    "finalize); the kernel's steady state sits at the MMA issue floor (1024 clk per 128x256x256 "
                    "stage), the rest is per-CTA fill/drain; traffic = DRAM bytes of this kernel from the ncu "
                    "capture of this same command",
            "hbm": {"achieved_gbs": alg_bytes / kern_s / 1e9, "peak_gbs": hbm_peak},
            "vs_int_pipe_roof": gdist / world / pipe_peak,
        }
    else:
        roofline = {
            "bound": "hbm", "achieved": alg_bytes / kern_s / 1e9, "peak": hbm_peak, "unit": "GB/s",
            "frac": alg_bytes / kern_s / 1e9 / hbm_peak, "traffic": None, "peak_source": peak_src,
            "kernel": "hamming_knn2_kernel<8>", "kernel_ms": kern_ms,
            "note": "not HBM-bound by construction (~100 distances per byte moved); the binding roof is "
                    "the integer POPC pipe, reported in int_pipe",
            "int_pipe": int_pipe,
        }

    cpu = None
    os.sched_setaffinity(0, all_cpus)      # the CPU baseline and the extras' CPU arms may use every core
    if not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        sample = max(1, min(args.cpu_sample_pairs, P))
        pairs = [(q_np[p * NQ:(p + 1) * NQ], t_np[p * NT:(p + 1) * NT]) for p in range(sample)]
        cpu_match_pairs(pairs[:1], cores)
        dt_cpu, kind = cpu_match_pairs(pairs, cores)
        cpu = {"value": sample / dt_cpu, "unit": UNIT, "cores": cores, "kind": kind,
               "sample": f"first {sample} of the {P} pairs of this step, "
                         f"{'cv2.BFMatcher(NORM_HAMMING).knnMatch k=2' if kind == 'reference' else 'numpy oracle port'}"
                         f" + ratio test, {cores} threads"}

    extra = None
    if not args.no_extra:
        extra = extra_ransac(local, with_cpu=not args.no_cpu_baseline)
        extra["C5_large_db"] = c5
        extra["device_resident"] = devres
        if multi is not None:
            extra["single_call_multi_gpu"] = multi

    # ncu-derived pipe utilisations of the kernels (profiles/r02_kernels.json, captured with the recipe's
    # `ncu --set full` on this same workload) attached to the live per-stage times
    ncu = {}
    nj = ROOT / "profiles" / "r02_kernels.json"
    if nj.exists():
        try:
            ncu = json.loads(nj.read_text())
        except Exception:
            ncu = {}
    kernels = c2_rows + (krows or [])
    for row in kernels:
        if row["config"] == "C2" and ms_step > 0:
            row["share_of_step"] = row["ms_per_step"] / ms_step
        extra_info = ncu.get(row["kernel"])
        if extra_info:
            row["ncu"] = extra_info
    roofline["kernels"] = kernels

    e2e = {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": d2h_bytes,
           "steps": e2e_steps, "api": "cvMatchKnn2Batch (host buffers, CUDA-pinned)",
           "pageable": {"value": world * P / e2e_pg_s, "unit": UNIT,
                        "api": "cvMatchKnn2Batch, ordinary pageable numpy arrays in and out (what a P/Invoke caller passes): "
                               "staged through the library's pinned ring by its copy threads",
                        "plain_cudaMemcpyAsync": world * P / e2e_plain_s, "copy_threads": ceiling["copy_threads"]},
           "h2d": dict(ceiling, e2e_h2d_gbs=e2e_h2d_gbs, e2e_d2h_gbs=world * d2h_bytes / e2e_s / 1e9,
                       frac_of_h2d_ceiling=e2e_h2d_gbs / ceiling["c2_mix_h2d_gbs_aggregate"] if ceiling["c2_mix_h2d_gbs_aggregate"] > 0 else None,
                       frac_of_h2d_only_ceiling=e2e_h2d_gbs / ceiling["h2d_pinned_gbs_aggregate"] if ceiling["h2d_pinned_gbs_aggregate"] > 0 else None,
                       note="e2e_h2d_gbs = bytes the e2e step moves host->device per second over all ranks; frac_of_h2d_ceiling "
                            "divides it by the host->device rate bare pinned copies reach on this box when all ranks copy at "
                            "once WITH the same share of device->host traffic beside them (c2_mix); frac_of_h2d_only_ceiling by "
                            "the rate with no return traffic"),
           "others": others}
    if cpu is not None and cpu_others:
        cpu["others"] = cpu_others
    if cpu is not None:
        e2e["speedup_vs_cpu_per_gpu"] = e2e_value / world / cpu["value"]

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": dict(workload_config(args), engine=("tcgen05 int8 GEMM" if use_tc else "integer pipe XOR+POPC")),
        "gdist_per_s": gdist,
        "e2e": e2e,
        "numa_binding": numa, "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu,
        "ratio_pass_per_pair": n_pass_dev / P, "extra": extra,
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
