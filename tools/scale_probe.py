"""Probe for the scale-search row: times cvFindScaledBatchDevice and its dominant kernel (scale_cost_kernel)."""
import sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import ctypes as C
import numpy as np
import torch
from minicv_b200 import camera_pose as mcp, native, synth

P = int(sys.argv[1]) if len(sys.argv) > 1 else 256
N = int(sys.argv[2]) if len(sys.argv) > 2 else 2000
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
lib = native.load()
dev = torch.device("cuda:0")
off = np.arange(P + 1, dtype=np.int64) * N
rows = np.empty((P, 14)); Rs = np.empty((P, 9)); Ts = np.empty((P, 3)); W = np.empty((P * N, 3)); O = np.empty((P * N, 2))
planted = np.empty(P)
for p in range(P):
    rows[p], R, Ts[p], W[p * N:(p + 1) * N], O[p * N:(p + 1) * N], planted[p] = synth.scale_scene(p, N)
    Rs[p] = R.ravel()
t = lambda a: torch.from_numpy(a).to(dev)
dCam, dRot, dTr, dW, dO = t(rows), t(Rs), t(Ts), t(W), t(O)
dCost = torch.empty(P, dtype=torch.float64, device=dev); dSc = torch.empty(P, dtype=torch.float64, device=dev)
dST = torch.empty((P, 3), dtype=torch.float64, device=dev); dSg = torch.empty(P, dtype=torch.int32, device=dev)
st = torch.cuda.current_stream()
run = lambda: mcp.find_scaled_batch_device(dCam, dRot, dTr, dW, dO, off, dCost, dSc, dST, dSg, device=0, stream=st.cuda_stream)
run(); torch.cuda.synchronize()
lib.mcvKernelTimingStart(0, 64)
e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
e0.record(st)
for _ in range(reps):
    run()
e1.record(st); torch.cuda.synchronize()
ms = (C.c_float * 64)()
k = lib.mcvKernelTimingStop(0, ms, 64)
tot = e0.elapsed_time(e1) / reps
kms = sum(ms[i] for i in range(k)) / reps
ev = P * 2 * N * N
rel = np.abs(dSc.cpu().numpy() - planted) / planted
print(f"P={P} N={N}: {tot:.3f} ms per batch, cost kernel {kms:.3f} ms ({k} timed launches), "
      f"{ev / tot / 1e6:.1f} Geval/s whole, {ev / kms / 1e6:.1f} Geval/s kernel, planted recovered {(rel < 0.01).sum()}/{P}")
