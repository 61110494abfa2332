"""GPU parity for the PnP-RANSAC row: cvSolvePnPRansac (AP3P kind),
solveAp3p, and the batched / injected forms -- all through the C ABI.

Contract (same as the two-view estimators, BASELINE.json north_star): injected hypotheses -> inlier counts
exactly equal to the oracle's; the device sampling stream reproduces OpenCV's sample sequence; inlier lists
identical to cv2.solvePnPRansac's on the goldens and final poses within 1e-6 relative.  Oracle =
oracle/pnp.py (pinned to the compiled reference AP3P and to cv2 4.13 goldens in tests/golden/pnp.npz)."""
import numpy as np
import pytest

from minicv_b200 import estimation as est
from minicv_b200 import synth
from oracle import pnp
from oracle.ransac import RansacTrace
from conftest import GOLDEN

pytestmark = pytest.mark.gpu

REL = 1e-6


def rel_err(A, B):
    A = np.asarray(A, float); B = np.asarray(B, float)
    return np.abs(A - B).max() / max(np.abs(B).max(), 1e-300)


@pytest.fixture(scope="module")
def g():
    return np.load(GOLDEN / "pnp.npz")


def _cam4(K):
    return 1 / K[0, 0], 1 / K[1, 1], K[0, 2] / K[0, 0], K[1, 2] / K[1, 1]


# ----------------------------------------------------------------------------- solveAp3p
def test_solve_ap3p_vs_compiled_reference(g):
    """The reference's own export (ap3p.cpp:282): counts equal, poses to 1e-7 on (almost) every triplet."""
    K = g["K"]
    n = len(g["ap3p_n"])
    Rs, ts, counts = est.solve_ap3p_batch(g["ap3p_img"], g["ap3p_obj"], np.tile(_cam4(K), (n, 1)))
    np.testing.assert_array_equal(counts, g["ap3p_n"])
    close = 0
    for i in range(n):
        c = counts[i]
        d = 0.0
        if c:
            d = max(np.abs(Rs[i, :c].reshape(c, 9) - g["ap3p_R"][i, :c]).max(), np.abs(ts[i, :c] - g["ap3p_t"][i, :c]).max())
        close += d < 1e-7
    assert close >= n - 2          # an ill-conditioned quartic amplifies last-ulp differences of cbrt / cos
    # the single-call export with the reference's signature
    for i in (0, 5, 17):
        R1, t1 = est.solveAp3p(g["ap3p_img"][i], g["ap3p_obj"][i], *_cam4(K))
        assert len(R1) == counts[i]
        np.testing.assert_array_equal(R1, Rs[i, :counts[i]])
        np.testing.assert_array_equal(t1, ts[i, :counts[i]])


# ----------------------------------------------------------------------------- cvSolvePnPRansac vs cv2
@pytest.mark.parametrize("dk", ["d0", "d1"])
@pytest.mark.parametrize("ci", [0, 1, 2, 3, 4])
def test_solve_pnp_ransac_vs_cv2_golden(g, dk, ci):
    K, dist = g["K"], g[f"dist_{dk}"]
    key = f"ransac_{dk}_{ci}"
    iters, thr, conf = g[f"{key}_prm"]
    res = est.solvePnPRansac(est.SolverKind.AP3P, est.PnPRansacParams(float(thr), int(iters), float(conf)),
                             g[f"{key}_img"], g[f"{key}_obj"], K, dist)
    assert (res is not None) == bool(g[f"{key}_ok"][0])
    if res is None:
        return
    r, t, inl = res
    np.testing.assert_array_equal(inl, g[f"{key}_inliers"])
    if len(inl) >= 6:              # EPnP on 4-5 points: null-space basis decided by rounding noise
        assert rel_err(r, g[f"{key}_rvec"]) < REL
        assert rel_err(t, g[f"{key}_tvec"]) < REL


def test_solve_pnp_ransac_failure_conventions(g):
    K, dist = g["K"], g["dist_d0"]
    X, uv, _, _ = synth.pnp_scene(3, 3, 1.0)
    assert est.solvePnPRansac(est.SolverKind.AP3P, est.PnPRansacParams(), uv, X, K, dist) is None      # N < 4
    assert est.solvePnPRansac(est.SolverKind.AP3P, est.PnPRansacParams(), uv, X[:2], K, dist) is None  # lengths differ
    # a solver kind that is not native is forwarded; without the original shim that is the failure value
    X, uv, _, _ = synth.pnp_scene(4, 50, 0.8)
    assert est.solvePnPRansac(est.SolverKind.EPNP, est.PnPRansacParams(), uv, X, K, dist) is None


# ----------------------------------------------------------------------------- injected hypotheses
@pytest.mark.parametrize("dk", ["d0", "d1"])
def test_injected_models_counts_exact(g, dk):
    """Same (rvec, tvec) hypotheses scored by both sides: inlier counts must be identical."""
    K, dist = g["K"], g[f"dist_{dk}"]
    rng = np.random.default_rng(11)
    X, uv, rvec, tvec = synth.pnp_scene(rng, 1500, 0.4, noise=1.0, K=K)
    H = 256
    models = np.zeros((1, H, 6))
    models[0, :, :3] = rvec + rng.normal(size=(H, 3)) * 0.004
    models[0, :, 3:] = tvec + rng.normal(size=(H, 3)) * 0.01
    models[0, 0] = np.concatenate([rvec, tvec])
    models[0, 1, :3] = 0                                   # theta = 0 branch of Rodrigues
    r = est.solve_pnp_ransac_batch(X, uv, [0, len(X)], est.camera_rows(K, dist), thr=2.0, maxIters=H,
                                   stop_mode=est.STOP_FIXED, inj_models=models, debug=True)
    cb = pnp.PnPCB(K, dist)
    o32, i32 = X.astype(np.float32), uv.astype(np.float32)
    thr2 = np.float32(2.0 * 2.0)
    oc = np.array([int((cb.compute_error(o32, i32, models[0, h].reshape(2, 3).T) <= thr2).sum()) for h in range(H)])
    np.testing.assert_array_equal(r.dbg_counts[0], oc)
    assert oc.max() > 100                                  # the test is not vacuous
    # winner = first maximum (strict improvement rule), mask = that model's inliers
    b = int(np.argmax(oc))
    assert r.best_iter[0] == b and r.inlier_counts[0] == oc[b]
    mask = cb.compute_error(o32, i32, models[0, b].reshape(2, 3).T) <= thr2
    np.testing.assert_array_equal(r.masks.astype(bool), mask)


@pytest.mark.parametrize("dk", ["d0", "d1"])
def test_sampling_stream_models_and_counts_vs_oracle(g, dk):
    K, dist = g["K"], g[f"dist_{dk}"]
    X, uv, _, _ = synth.pnp_scene(21, 400, 0.5, K=K)
    tr = RansacTrace()
    iters = 300
    cb = pnp.PnPCB(K, dist)
    from oracle.ransac import ransac_run
    ransac_run(cb, X.astype(np.float32), uv.astype(np.float32), 3.0, 0.99, iters, fixed_iters=True, trace=tr)
    r = est.solve_pnp_ransac_batch(X, uv, [0, len(X)], est.camera_rows(K, dist), thr=3.0, maxIters=iters,
                                   stop_mode=est.STOP_FIXED, debug=True)
    # (1) OpenCV's sample sequence, exactly
    np.testing.assert_array_equal(r.dbg_samples[0], np.array(tr.samples))
    # (2) minimal models and counts
    oc = np.array([c[0] if c else -1 for c in tr.counts])
    assert ((oc >= 0) == (r.dbg_counts[0] >= 0)).mean() >= 0.99
    assert (oc == r.dbg_counts[0]).mean() >= 0.99
    close = tot = 0
    for it in range(iters):
        if tr.models[it] and r.dbg_counts[0, it] >= 0:
            m = tr.models[it][0]
            tot += 1
            close += max(np.abs(r.dbg_models[0, it, :3] - m[:, 0]).max(), np.abs(r.dbg_models[0, it, 3:] - m[:, 1]).max()) < 1e-6
    assert tot > 200 and close >= 0.98 * tot


# ----------------------------------------------------------------------------- batch: ragged, edge cases
def test_ragged_batch_vs_oracle(g):
    K = g["K"]
    sizes = [0, 3, 4, 5, 40, 333, 1200, 64]
    dists = [g["dist_d0"], g["dist_d1"]]
    Xs, uvs, cams = [], [], []
    for i, n in enumerate(sizes):
        X, uv, _, _ = synth.pnp_scene(100 + i, max(n, 1), 0.5, K=K)
        Xs.append(X[:n]); uvs.append(uv[:n])
        cams.append(est.camera_rows(K, dists[i % 2])[0])
    off = np.concatenate([[0], np.cumsum(sizes)])
    r = est.solve_pnp_ransac_batch(np.concatenate(Xs), np.concatenate(uvs), off, np.array(cams), thr=3.0, maxIters=200,
                                   confidence=0.99)
    for i, n in enumerate(sizes):
        ok, rvec, tvec, idx = pnp.solve_pnp_ransac(Xs[i], uvs[i], K, dists[i % 2], 200, 3.0, 0.99) if n else (False, None, None, [])
        assert bool(r.ok[i]) == bool(ok), i
        m = r.masks[off[i]:off[i + 1]]
        if not ok:
            assert m.sum() == 0 and r.inlier_counts[i] == 0
            continue
        np.testing.assert_array_equal(np.nonzero(m)[0], idx)
        assert r.inlier_counts[i] == len(idx)
        if len(idx) >= 6:
            assert rel_err(r.rvecs[i], rvec) < REL and rel_err(r.tvecs[i], tvec) < REL
    assert r.n_ok == int(r.ok.sum())


def test_counter_sampling_recovers_pose():
    """Production mode (counter RNG, fixed hypothesis count): every problem of a batch finds its pose."""
    K = np.array([[800.0, 0, 320.0], [0, 800.0, 240.0], [0, 0, 1.0]])
    P, n = 32, 1000
    Xs, uvs, gt = [], [], []
    for p in range(P):
        X, uv, rvec, tvec = synth.pnp_scene(500 + p, n, 0.3, K=K)
        Xs.append(X); uvs.append(uv); gt.append((rvec, tvec))
    off = np.arange(P + 1) * n
    r = est.solve_pnp_ransac_batch(np.concatenate(Xs), np.concatenate(uvs), off, est.camera_rows(K, None, P), thr=2.0,
                                   maxIters=512, sample_mode=est.SAMPLE_COUNTER, stop_mode=est.STOP_FIXED, seed=7)
    assert r.n_ok == P
    for p in range(P):
        assert r.inlier_counts[p] >= 0.75 * 0.3 * n       # the best of 512 minimal models, 0.5 px noise, 2 px gate
        assert np.abs(r.rvecs[p] - gt[p][0]).max() < 5e-3 and np.abs(r.tvecs[p] - gt[p][1]).max() < 2e-2
    # deterministic call to call
    r2 = est.solve_pnp_ransac_batch(np.concatenate(Xs), np.concatenate(uvs), off, est.camera_rows(K, None, P), thr=2.0,
                                    maxIters=512, sample_mode=est.SAMPLE_COUNTER, stop_mode=est.STOP_FIXED, seed=7)
    np.testing.assert_array_equal(r.rvecs, r2.rvecs)
    np.testing.assert_array_equal(r.masks, r2.masks)


def test_device_resident_form_matches_host_form():
    import torch
    K = np.array([[800.0, 0, 320.0], [0, 800.0, 240.0], [0, 0, 1.0]])
    P, n = 8, 500
    Xs, uvs = [], []
    for p in range(P):
        X, uv, _, _ = synth.pnp_scene(900 + p, n, 0.4, K=K)
        Xs.append(X); uvs.append(uv)
    X = np.concatenate(Xs); uv = np.concatenate(uvs)
    off = np.arange(P + 1) * n
    cams = est.camera_rows(K, None, P)
    h = est.solve_pnp_ransac_batch(X, uv, off, cams, thr=3.0, maxIters=300, confidence=0.99)
    dev = torch.device("cuda", 0)
    dX = torch.from_numpy(X).to(dev); dU = torch.from_numpy(uv).to(dev); dC = torch.from_numpy(cams).to(dev)
    dR = torch.zeros((P, 3), dtype=torch.float64, device=dev); dT = torch.zeros_like(dR)
    dM = torch.zeros(P * n, dtype=torch.uint8, device=dev)
    dCnt = torch.zeros(P, dtype=torch.int32, device=dev); dOk = torch.zeros(P, dtype=torch.uint8, device=dev)
    st = torch.cuda.current_stream()
    est.solve_pnp_ransac_batch_device(dX, dU, off, dC, 3.0, 300, 0.99, dR, dT, dM, dCnt, dOk, device=0, stream=st.cuda_stream)
    torch.cuda.synchronize()
    np.testing.assert_array_equal(dOk.cpu().numpy(), h.ok)
    np.testing.assert_array_equal(dM.cpu().numpy(), h.masks)
    np.testing.assert_array_equal(dCnt.cpu().numpy(), h.inlier_counts)
    np.testing.assert_array_equal(dR.cpu().numpy(), h.rvecs)
    np.testing.assert_array_equal(dT.cpu().numpy(), h.tvecs)
