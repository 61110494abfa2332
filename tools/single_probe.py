"""One cvFindHomography / cvRecoverPose call each (after a warm-up call) -- for an ncu launch list."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from minicv_b200 import estimation as est, synth  # noqa: E402

a, b, _, _ = synth.homography_pair(1, 2000)
est.findHomography(a, b, 3.0, 2000, 0.995)
est.findHomography(a, b, 3.0, 2000, 0.995)
