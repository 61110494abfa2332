"""One device-resident pass of C3-F (fundamental, counter sampling, 4096 hypotheses) and C4 (cvRecoverPose semantics)
on a reduced pair count -- the command the ncu launch list / full captures of the solver kernels are taken from.
  python tools/solver_probe.py [pairsF pairsE]"""
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))

if __name__ == "__main__":
    import torch
    from minicv_b200 import estimation as est, synth
    PF = int(sys.argv[1]) if len(sys.argv) > 1 else 256
    PE = int(sys.argv[2]) if len(sys.argv) > 2 else 128
    dev = torch.device("cuda:0")
    torch.cuda.set_device(0)
    s = torch.cuda.current_stream()
    N, HYP = 2000, 4096
    for kind, model, gen in (("H", est.MODEL_H, synth.homography_pair), ("F", est.MODEL_F, synth.two_view_pair)):
        A = np.empty((PF * N, 2)); B = np.empty((PF * N, 2))
        for p in range(PF):
            r = gen(p, N)
            A[p * N:(p + 1) * N] = r[0]; B[p * N:(p + 1) * N] = r[1]
        off = np.arange(PF + 1, dtype=np.int64) * N
        dA = torch.from_numpy(A).to(dev); dB = torch.from_numpy(B).to(dev)
        dM = torch.empty((PF, 9), dtype=torch.float64, device=dev); dMask = torch.empty(PF * N, dtype=torch.uint8, device=dev)
        dCnt = torch.empty(PF, dtype=torch.int32, device=dev); dOk = torch.empty(PF, dtype=torch.uint8, device=dev)
        for _ in range(2):
            est.ransac_batch_device(model, dA, dB, off, 3.0, HYP, dM, dMask, dCnt, dOk, confidence=0.99,
                                    sample_mode=est.SAMPLE_COUNTER, stop_mode=est.STOP_FIXED, seed=1, stream=s.cuda_stream)
        torch.cuda.synchronize()
        print(kind, "ok", int(dOk.sum().item()), "mean inliers", float(dCnt.float().mean().item()))
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
    for _ in range(2):
        est.recover_pose_batch_device(cfg, dA, dB, off, dE, dR, dT, dMs, dGood, dOk, sample_mode=est.SAMPLE_MWC, maxIters=1000,
                                      stream=s.cuda_stream)
    torch.cuda.synchronize()
    print("E ok", int(dOk.sum().item()), "mean cheirality", float(dGood.float().mean().item()))
