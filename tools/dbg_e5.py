"""Debug aid: the 5-point goldens through the library given by $MINICV_NATIVE_LIB (e.g. the composed debug variant),
classifying every mismatch against cv2: sign flip, other solution, or garbage."""
import sys
from pathlib import Path
import numpy as np
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from minicv_b200 import estimation as est

g = np.load(Path(__file__).resolve().parent.parent / "tests" / "golden" / "ransac_e.npz")
tot = bad = flips = 0
for si, (a, b, Ec, n) in enumerate(zip(g["min5_a"], g["min5_b"], g["min5_E"], g["min5_n"])):
    Es = est.fivepoint(a, b)
    for i in range(n):
        tot += 1
        ref = Ec[3 * i:3 * i + 3]
        d = min(np.abs(e - ref).max() for e in Es) if len(Es) else 9
        dn = min(np.abs(e + ref).max() for e in Es) if len(Es) else 9
        if d > 1e-6:
            bad += 1
            flips += dn < 1e-6
            if bad <= 12:
                print(f"set {si} root {i}: n={n} got={len(Es)} |e-ref|={d:.2e} |e+ref|={dn:.2e}")
print("tot", tot, "bad", bad, "sign flips", flips)
