"""C3-F and C4 device-resident with both scoring engines: step time (CUDA events, median of 5) and the per-stage
CUDA-event breakdown of the library.   python tools/score_probe.py [pairsHF pairsE] [engines, e.g. 12] [configs, e.g. HFE]"""
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))

if __name__ == "__main__":
    import torch
    from minicv_b200 import estimation as est, synth, native
    lib = native.load()
    PF = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
    PE = int(sys.argv[2]) if len(sys.argv) > 2 else 512
    engines = [int(c) for c in (sys.argv[3] if len(sys.argv) > 3 else "12")]
    dev = torch.device("cuda:0")
    torch.cuda.set_device(0)
    s = torch.cuda.current_stream()

    def timed(run, tag):
        for _ in range(2):
            run()
        torch.cuda.synchronize()
        ts = []
        for _ in range(5):
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record(s); run(); e1.record(s); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        lib.mcvKernelTimingStart(0, 4096)
        run(); torch.cuda.synchronize()
        ms = np.zeros(4096, np.float32); tags = np.zeros(4096, np.int32)
        n = lib.mcvKernelTimingStopTagged(0, ms.ctypes.data, tags.ctypes.data, 4096)
        st = {int(t): round(float(ms[:n][tags[:n] == t].sum()), 3) for t in sorted(set(tags[:n].tolist()))}
        print(tag, "median ms", round(float(np.median(ts)), 3), "stages", st, flush=True)

    N, HYP = 2000, 4096
    if PF > 0 and "H" in (sys.argv[4] if len(sys.argv) > 4 else "HFE"):
        A = np.empty((PF * N, 2)); B = np.empty((PF * N, 2))
        for p in range(PF):
            r = synth.homography_pair(p, N)
            A[p * N:(p + 1) * N] = r[0]; B[p * N:(p + 1) * N] = r[1]
        off = np.arange(PF + 1, dtype=np.int64) * N
        dA = torch.from_numpy(A).to(dev); dB = torch.from_numpy(B).to(dev)
        dM = torch.empty((PF, 9), dtype=torch.float64, device=dev); dMask = torch.empty(PF * N, dtype=torch.uint8, device=dev)
        dCnt = torch.empty(PF, dtype=torch.int32, device=dev); dOk = torch.empty(PF, dtype=torch.uint8, device=dev)
        res = {}
        for eng in engines:
            est.set_score_engine(eng)
            timed(lambda: est.ransac_batch_device(est.MODEL_H, dA, dB, off, 3.0, HYP, dM, dMask, dCnt, dOk, confidence=0.995,
                                                  sample_mode=est.SAMPLE_COUNTER, stop_mode=est.STOP_FIXED, seed=1,
                                                  stream=s.cuda_stream), f"C3-H {PF} pairs engine {eng}")
            res[eng] = (dM.cpu().numpy().copy(), dCnt.cpu().numpy().copy(), dMask.cpu().numpy().copy())
        if len(engines) > 1:
            a, b = res[engines[0]], res[engines[1]]
            print("H identical:", all(np.array_equal(x, y, equal_nan=True) for x, y in zip(a, b)), "mean inliers", a[1].mean())
        del dA, dB, dM, dMask
    if PF > 0 and "F" in (sys.argv[4] if len(sys.argv) > 4 else "HFE"):
        A = np.empty((PF * N, 2)); B = np.empty((PF * N, 2))
        for p in range(PF):
            r = synth.two_view_pair(p, N)
            A[p * N:(p + 1) * N] = r[0]; B[p * N:(p + 1) * N] = r[1]
        off = np.arange(PF + 1, dtype=np.int64) * N
        dA = torch.from_numpy(A).to(dev); dB = torch.from_numpy(B).to(dev)
        dM = torch.empty((PF, 9), dtype=torch.float64, device=dev); dMask = torch.empty(PF * N, dtype=torch.uint8, device=dev)
        dCnt = torch.empty(PF, dtype=torch.int32, device=dev); dOk = torch.empty(PF, dtype=torch.uint8, device=dev)
        res = {}
        for eng in engines:
            est.set_score_engine(eng)
            timed(lambda: est.ransac_batch_device(est.MODEL_F, dA, dB, off, 3.0, HYP, dM, dMask, dCnt, dOk, confidence=0.99,
                                                  sample_mode=est.SAMPLE_COUNTER, stop_mode=est.STOP_FIXED, seed=1,
                                                  stream=s.cuda_stream), f"C3-F {PF} pairs engine {eng}")
            res[eng] = (dM.cpu().numpy().copy(), dCnt.cpu().numpy().copy(), dMask.cpu().numpy().copy())
        if len(engines) > 1:
            a, b = res[engines[0]], res[engines[1]]
            print("F identical:", all(np.array_equal(x, y, equal_nan=True) for x, y in zip(a, b)), "mean inliers", a[1].mean())
        del dA, dB, dM, dMask
    if PE > 0 and "E" in (sys.argv[4] if len(sys.argv) > 4 else "HFE"):
        N = 5000
        A = np.empty((PE * N, 2)); B = np.empty((PE * N, 2))
        for p in range(PE):
            r = synth.two_view_pair(p, N)
            A[p * N:(p + 1) * N] = r[0]; B[p * N:(p + 1) * N] = r[1]
        off = np.arange(PE + 1, dtype=np.int64) * N
        dA = torch.from_numpy(A).to(dev); dB = torch.from_numpy(B).to(dev)
        dE = torch.empty((PE, 9), dtype=torch.float64, device=dev); dR = torch.empty((PE, 9), dtype=torch.float64, device=dev)
        dT = torch.empty((PE, 3), dtype=torch.float64, device=dev); dMs = torch.empty(PE * N, dtype=torch.uint8, device=dev)
        dGood = torch.empty(PE, dtype=torch.int32, device=dev); dOk = torch.empty(PE, dtype=torch.uint8, device=dev)
        cfg = est.RecoverPoseConfig(800.0, (320.0, 240.0), 0.999, 1.0)
        res = {}
        for eng in engines:
            est.set_score_engine(eng)
            timed(lambda: est.recover_pose_batch_device(cfg, dA, dB, off, dE, dR, dT, dMs, dGood, dOk, sample_mode=est.SAMPLE_MWC,
                                                        maxIters=1000, stream=s.cuda_stream), f"C4 {PE} pairs engine {eng}")
            res[eng] = (dE.cpu().numpy().copy(), dGood.cpu().numpy().copy(), dMs.cpu().numpy().copy(), dR.cpu().numpy().copy())
        if len(engines) > 1:
            a, b = res[engines[0]], res[engines[1]]
            print("E identical:", all(np.array_equal(x, y, equal_nan=True) for x, y in zip(a, b)), "mean cheirality", a[1].mean())
    est.set_score_engine(0)
