"""Runs the PnP-RANSAC bench configuration (1024 problems x 2000 correspondences x 1000 iterations, device
resident) a few times in both sampling modes -- used plain for timing and under ncu for the launch list."""
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from minicv_b200 import estimation as est, synth  # noqa: E402

if __name__ == "__main__":
    import torch
    P = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
    N, ITERS = 2000, 1000
    dev = torch.device("cuda", 0)
    st = torch.cuda.current_stream()
    K = np.array([[800.0, 0, 320.0], [0, 800.0, 240.0], [0, 0, 1.0]])
    X = np.empty((P * N, 3)); U = np.empty((P * N, 2))
    for p in range(P):
        X[p * N:(p + 1) * N], U[p * N:(p + 1) * N], _, _ = synth.pnp_scene(p, N, 0.3, noise=0.5, K=K)
    off = np.arange(P + 1, dtype=np.int64) * N
    dX = torch.from_numpy(X).to(dev); dU = torch.from_numpy(U).to(dev)
    dC = torch.from_numpy(est.camera_rows(K, None, P)).to(dev)
    dR = torch.empty((P, 3), dtype=torch.float64, device=dev); dT = torch.empty_like(dR)
    dM = torch.empty(P * N, dtype=torch.uint8, device=dev)
    dCn = torch.empty(P, dtype=torch.int32, device=dev); dOk = torch.empty(P, dtype=torch.uint8, device=dev)
    for name, sm, stp in (("mwc+adaptive", est.SAMPLE_MWC, est.STOP_ADAPTIVE), ("counter+fixed", est.SAMPLE_COUNTER, est.STOP_FIXED)):
        def run():
            est.solve_pnp_ransac_batch_device(dX, dU, off, dC, 2.0, ITERS, 0.999, dR, dT, dM, dCn, dOk, sample_mode=sm,
                                              stop_mode=stp, seed=3, device=0, stream=st.cuda_stream)
        run(); torch.cuda.synchronize()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(st)
        for _ in range(3):
            run()
        e1.record(st); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 3
        print(f"{name}: {ms:.3f} ms per batch of {P} = {P / ms * 1e3:.0f} problems/s, ok {int(dOk.sum())}, "
              f"mean inliers {float(dCn.float().mean()):.1f}")
