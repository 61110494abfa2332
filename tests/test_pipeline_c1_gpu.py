"""GPU, config C1 (BASELINE.json configs[0]): textured planar pair, ORB / AKAZE descriptors from
cv2 (frozen in tests/golden/c1_planar.npz), then the whole path on the GPU through the C ABI:
kNN-2 + ratio (bit-exact vs cv2.BFMatcher) -> findHomography (vs cv2.findHomography)."""
import numpy as np
import pytest

import minicv_b200 as mcv
from minicv_b200 import estimation as est
from oracle import bf_knn, ransac
from conftest import GOLDEN

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", ["orb", "akaze"])
@pytest.mark.parametrize("engine", [mcv.ENGINE_POPC, mcv.ENGINE_TENSOR])
def test_c1_match_then_homography(name, engine):
    g = np.load(GOLDEN / "c1_planar.npz")
    mcv.set_hamming_engine(engine)
    try:
        r = mcv.match_features(g[f"{name}_d1"], g[f"{name}_d2"], 0.8)
    finally:
        mcv.set_hamming_engine(mcv.ENGINE_AUTO)
    np.testing.assert_array_equal(r.train_idx, g[f"{name}_idx"])
    np.testing.assert_array_equal(r.distance, g[f"{name}_dist"])
    ok = r.ratio_pass.astype(bool)
    assert ok.sum() > 1000
    a = g[f"{name}_p1"][ok]
    b = g[f"{name}_p2"][r.train_idx[ok, 0]]
    H, mask = est.findHomography(a, b, 3.0, 2000, 0.995)
    assert H is not None
    Hc = g[f"{name}_H"]
    assert np.abs(H - Hc).max() / np.abs(Hc).max() < 1e-6
    assert (mask == g[f"{name}_mask"]).mean() >= 0.999
    # and the estimate is the true plane-induced homography to sub-pixel accuracy
    Hgt = g["Hgt"]
    pts = np.array([[50.0, 50, 1], [600, 60, 1], [320, 240, 1], [80, 430, 1], [590, 400, 1]]).T
    pa = (H @ pts); pa = pa[:2] / pa[2]
    pb = (Hgt @ pts); pb = pb[:2] / pb[2]
    assert np.abs(pa - pb).max() < 1.0
    # oracle agrees too (CPU restatement of the same chain)
    Ho, mo = ransac.find_homography(a, b, 3.0, 2000, 0.995)
    assert np.abs(H - Ho).max() / np.abs(Ho).max() < 1e-6
