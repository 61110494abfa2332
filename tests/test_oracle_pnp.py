"""CPU suite for the PnP row: the numpy oracle (oracle/pnp.py) against (a) the reference's own AP3P solver
compiled from /root/reference (frozen outputs in tests/golden/pnp.npz, and live when oracle/_ref exists)
and (b) cv2 4.13.0 outputs frozen by tests/golden/make_golden_pnp.py."""
import ctypes as C

import numpy as np
import pytest

from conftest import GOLDEN, ROOT
from oracle import pnp


@pytest.fixture(scope="module")
def g():
    return np.load(GOLDEN / "pnp.npz")


def _ap3p_args(K):
    return 1 / K[0, 0], 1 / K[1, 1], K[0, 2] / K[0, 0], K[1, 2] / K[1, 1]


def test_ap3p_matches_reference_golden(g):
    """solve_ap3p == the reference's solveAp3p (ap3p.cpp:282) on 256 triplets, duplicates included."""
    K = g["K"]
    for i in range(len(g["ap3p_n"])):
        Rs, ts = pnp.solve_ap3p(g["ap3p_img"][i], g["ap3p_obj"][i], *_ap3p_args(K))
        n = int(g["ap3p_n"][i])
        assert len(Rs) == n
        if n:
            np.testing.assert_allclose(Rs.reshape(n, 9), g["ap3p_R"][i, :n], rtol=0, atol=1e-8)
            np.testing.assert_allclose(ts, g["ap3p_t"][i, :n], rtol=0, atol=1e-8)


def test_ap3p_matches_compiled_reference_live(g):
    so = ROOT / "oracle" / "_ref" / "libap3p_ref.so"
    if not so.exists():
        pytest.skip("oracle/_ref not built (needs /root/reference; built by __graft_entry__.build())")
    lib = C.CDLL(str(so))
    lib.solveAp3p.restype = C.c_int
    lib.solveAp3p.argtypes = [C.c_void_p, C.c_void_p] + [C.c_double] * 19
    rng = np.random.default_rng(99)
    K = g["K"]
    from minicv_b200 import synth
    for _ in range(200):
        X, uv, _, _ = synth.pnp_scene(rng, 3, 1.0, noise=0.0, K=K)
        R = np.zeros((4, 9)); t = np.zeros((4, 3))
        args = []
        for j in range(3):
            args += [uv[j, 0], uv[j, 1], X[j, 0], X[j, 1], X[j, 2]]
        n = lib.solveAp3p(R.ctypes.data, t.ctypes.data, *args, *_ap3p_args(K))
        Rs, ts = pnp.solve_ap3p(uv, X, *_ap3p_args(K))
        assert n == len(Rs)
        if n:
            np.testing.assert_allclose(Rs.reshape(n, 9), R[:n], rtol=0, atol=1e-8)
            np.testing.assert_allclose(ts, t[:n], rtol=0, atol=1e-8)


def test_rodrigues(g):
    for rv, R in zip(g["rod_rvec"], g["rod_R"]):
        assert np.array_equal(pnp.rodrigues_to_matrix(rv), R)          # bit-exact
    for Rin, rv in zip(g["rod_Rin"], g["rod_rvec_back"]):
        np.testing.assert_allclose(pnp.rodrigues_to_vector(Rin), rv, rtol=0, atol=1e-12)


@pytest.mark.parametrize("dk", ["d0", "d1"])
def test_project_and_undistort_bit_exact(g, dk):
    K, dist = g["K"], g[f"dist_{dk}"]
    p = pnp.project_points(g[f"proj_{dk}_obj"], g[f"proj_{dk}_rvec"], g[f"proj_{dk}_tvec"], K, dist)
    assert np.array_equal(p, g[f"proj_{dk}_f64"])
    assert np.array_equal(p.astype(np.float32), g[f"proj_{dk}_f32"])
    u = pnp.undistort_points(g[f"und_{dk}_in"], K, dist, out_f32=True)
    assert u.dtype == np.float32 and np.array_equal(u, g[f"und_{dk}_out"])


@pytest.mark.parametrize("dk", ["d0", "d1"])
def test_solvepnp_ap3p_four_points(g, dk):
    K, dist = g["K"], g[f"dist_{dk}"]
    close = 0
    n = len(g[f"p4_{dk}_ok"])
    for i in range(n):
        ok, rvec, tvec = pnp.solve_pnp_ap3p4(g[f"p4_{dk}_obj"][i], g[f"p4_{dk}_img"][i], K, dist)
        assert bool(ok) == bool(g[f"p4_{dk}_ok"][i])
        if ok:
            d = max(np.abs(rvec - g[f"p4_{dk}_rvec"][i]).max(), np.abs(tvec - g[f"p4_{dk}_tvec"][i]).max())
            close += d < 1e-6
    assert close >= n - 2          # ill-conditioned quartics amplify the last-ulp differences of cbrt/pow


@pytest.mark.parametrize("dk", ["d0", "d1"])
@pytest.mark.parametrize("n", [6, 40, 300])
def test_epnp(g, dk, n):
    K, dist = g["K"], g[f"dist_{dk}"]
    rvec, tvec = pnp.solve_pnp_epnp(g[f"epnp_{dk}_{n}_obj"], g[f"epnp_{dk}_{n}_img"], K, dist)
    np.testing.assert_allclose(rvec, g[f"epnp_{dk}_{n}_rvec"], rtol=0, atol=1e-9)
    np.testing.assert_allclose(tvec, g[f"epnp_{dk}_{n}_tvec"], rtol=0, atol=1e-9)


@pytest.mark.parametrize("dk", ["d0", "d1"])
@pytest.mark.parametrize("ci", [0, 2, 3, 4])
def test_solvepnp_ransac_matches_cv2(g, dk, ci):
    """Inlier lists identical to cv2.solvePnPRansac(flags=AP3P); final pose (EPnP on the inliers) to 1e-9."""
    K, dist = g["K"], g[f"dist_{dk}"]
    key = f"ransac_{dk}_{ci}"
    iters, thr, conf = g[f"{key}_prm"]
    ok, rvec, tvec, idx = pnp.solve_pnp_ransac(g[f"{key}_obj"], g[f"{key}_img"], K, dist, int(iters), float(thr), float(conf))
    assert bool(ok) == bool(g[f"{key}_ok"][0])
    assert np.array_equal(idx, g[f"{key}_inliers"])
    # EPnP on 4 or 5 points has a (numerically) 4-dimensional null space whose basis is decided by rounding
    # noise inside the 12x12 Jacobi SVD: the final pose is only comparable from 6 inliers up
    if ok and len(idx) >= 6:
        np.testing.assert_allclose(rvec, g[f"{key}_rvec"], rtol=0, atol=1e-9)
        np.testing.assert_allclose(tvec, g[f"{key}_tvec"], rtol=0, atol=1e-9)
