"""Where the tensor-core scoring engine stops paying: C3-H / C3-F shaped batches at several inlier ratios, scoring-stage
time (library CUDA events) with engine 1 (fp32 filter) and engine 2 (tcgen05), and what auto mode picks.
    python tools/gate_probe.py [pairs]"""
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))

if __name__ == "__main__":
    import torch
    from minicv_b200 import estimation as est, synth, native
    lib = native.load()
    P = int(sys.argv[1]) if len(sys.argv) > 1 else 256
    dev = torch.device("cuda:0")
    s = torch.cuda.current_stream()
    N, HYP = 2000, 4096

    def score_ms(run):
        run(); torch.cuda.synchronize()
        lib.mcvKernelTimingStart(0, 4096)
        run(); torch.cuda.synchronize()
        ms = np.zeros(4096, np.float32); tags = np.zeros(4096, np.int32)
        n = lib.mcvKernelTimingStopTagged(0, ms.ctypes.data, tags.ctypes.data, 4096)
        return float(ms[:n][tags[:n] == 14].sum())

    for model, name in ((est.MODEL_H, "H"), (est.MODEL_F, "F")):
        for frac in (0.3, 0.4, 0.5, 0.6, 0.7, 0.85):
            A = np.empty((P * N, 2)); B = np.empty((P * N, 2))
            for p in range(P):
                r = synth.homography_pair(p, N, inlier_frac=frac) if model == est.MODEL_H else synth.two_view_pair(p, N, inlier_frac=frac)
                A[p * N:(p + 1) * N] = r[0]; B[p * N:(p + 1) * N] = r[1]
            off = np.arange(P + 1, dtype=np.int64) * N
            dA = torch.from_numpy(A).to(dev); dB = torch.from_numpy(B).to(dev)
            dM = torch.empty((P, 9), dtype=torch.float64, device=dev); dMask = torch.empty(P * N, dtype=torch.uint8, device=dev)
            dCnt = torch.empty(P, dtype=torch.int32, device=dev); dOk = torch.empty(P, dtype=torch.uint8, device=dev)
            out = {}
            for eng in (1, 2, 0):
                est.set_score_engine(eng)
                out[eng] = score_ms(lambda: est.ransac_batch_device(model, dA, dB, off, 3.0, HYP, dM, dMask, dCnt, dOk, confidence=0.99,
                                                                    sample_mode=est.SAMPLE_COUNTER, stop_mode=est.STOP_FIXED, seed=1,
                                                                    stream=s.cuda_stream))
            est.set_score_engine(0)
            print(f"{name} inliers {frac:.2f}: fp32-filter {out[1]:.3f} ms, tcgen05 {out[2]:.3f} ms, auto {out[0]:.3f} ms", flush=True)
