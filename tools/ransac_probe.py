"""Runs the C3 / C4 RANSAC configs once each (device resident) -- used under ncu for the launch list."""
import json
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import bench  # noqa: E402

if __name__ == "__main__":
    import torch
    torch.cuda.set_device(0)
    out = bench.extra_ransac(0, with_cpu=False)
    print(json.dumps(out))
