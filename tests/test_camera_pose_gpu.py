"""GPU parity for the scale-search row: CameraPose.findScaled
(src/MiniCV/CameraPose.fs:39-134) re-implemented behind the C ABI (cvFindScaled / cvFindScaledBatch /
cvFindScaledBatchDevice / cvScaleCosts).

Contract: everything is separately rounded fp64 arithmetic in the reference's evaluation order, so the bar is
BIT-EXACT against the oracle (oracle/camera_pose.py -- itself "parity unpinned": the reference holds no test
or fixture for this function): candidate scales, the cost of every candidate, the cost of injected scales, the
selected scale and the scaled pose.  At sizes the oracle cannot walk, size-independent properties: a planted
scale is recovered, the best cost equals the cost tap's minimum at its first position, and re-costing the
returned scale through the injected-scale entry point reproduces the returned cost."""
import math

import numpy as np
import pytest

from minicv_b200 import camera_pose as mcp
from minicv_b200 import synth
from oracle import camera_pose as cp

pytestmark = pytest.mark.gpu


def scene(seed, n, **kw):
    row, R, T, world, obs, s = synth.scale_scene(seed, n, **kw)
    cam = mcp.Camera(row[0:3], row[3:6], row[6:9], row[9:12], row[12:14])
    ocam = cp.Camera(row[0:3], row[3:6], row[6:9], row[9:12], row[12:14])
    return cam, ocam, mcp.CameraPose(2, 1, R, T, False), cp.CameraPose(2, 1, R, T, False), world, obs, s


def test_find_scaled_bit_exact_vs_oracle():
    for seed, n, kw in [(0, 1, {}), (1, 7, {}), (2, 130, {}), (3, 257, dict(outlier_frac=0.3)), (4, 600, dict(noise=0.0))]:
        cam, ocam, pose, opose, world, obs, s = scene(seed, n, **kw)
        cost, out = mcp.findScaled(0.01, cam, list(zip(world, obs)), pose)
        ocost, oscale, ocand, ook, ocosts = cp.find_scaled_fast(ocam, world, obs, opose)
        assert cost == ocost, (seed, cost, ocost)
        ref = cp.scale(oscale, opose)
        np.testing.assert_array_equal(out.Translation, ref.Translation)
        assert (out.ScaleSign, out.RotationIndex, out.IsInverse) == (ref.ScaleSign, 2, False)
        np.testing.assert_array_equal(out.Rotation, opose.Rotation)


def test_batch_taps_bit_exact_vs_oracle():
    """Ragged batch incl. an empty problem, a tiny-rejected observation and a problem where nothing projects."""
    sizes = [40, 0, 1, 300, 77, 129]
    scenes = [scene(10 + i, max(n, 1), outlier_frac=0.2 if i % 2 else 0.0) for i, n in enumerate(sizes)]
    worlds, obss = [], []
    for (cam, ocam, pose, opose, world, obs, s), n in zip(scenes, sizes):
        worlds.append(world[:n]); obss.append(obs[:n])
    # problem 4: all points behind every camera -> +inf, scale 0, zero translation
    worlds[4] = worlds[4] - 100.0 * scenes[4][0].forward
    off = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)
    res = mcp.find_scaled_batch([sc[0] for sc in scenes], [sc[2] for sc in scenes], np.concatenate(worlds),
                                np.concatenate(obss), off, taps=True)
    for i, (sc, n) in enumerate(zip(scenes, sizes)):
        cam, ocam, pose, opose = sc[:4]
        lo, hi = off[i], off[i + 1]
        if n == 0:
            assert res.costs[i] == math.inf and res.scales[i] == 0.0
            assert res.poses[i].ScaleSign == 0 and not res.poses[i].Rotation.any() and not res.poses[i].Translation.any()
            continue
        ocost, oscale, ocand, ook, ocosts = cp.find_scaled_fast(ocam, worlds[i], obss[i], opose)
        np.testing.assert_array_equal(res.cand_ok[lo:hi], ook)
        np.testing.assert_array_equal(res.candidates[lo:hi][ook], ocand[ook])
        np.testing.assert_array_equal(res.cand_costs[lo:hi], ocosts)
        assert res.costs[i] == ocost and res.scales[i] == oscale
        np.testing.assert_array_equal(res.poses[i].Translation, cp.scale(oscale, opose).Translation)
    assert res.costs[4] == math.inf and res.scales[4] == 0.0 and res.poses[4].ScaleSign == 0


def test_tiny_rejection_matches():
    """An observation whose x- or y-equation degenerates (|n| < 1e-5, Fun.IsTiny) proposes no candidate."""
    cam, ocam, pose, opose, world, obs, s = scene(5, 50)
    dst0 = cp.transformed_view(cp.transformation(cp.scale(0.0, opose)), ocam)
    t = cp._normalize(cp._transform_dir(cp.view_forward(dst0), cp._transform_dir(
        cp._from_basis(ocam.right, ocam.up, -ocam.forward, ocam.location), cp._mat_vec(opose.Rotation, opose.Translation))))
    obs = obs.copy()
    obs[7, 0] = t[0] / t[2]                       # n.x = t.x - obs.x t.z == 0 (to rounding)
    obs[20, 1] = t[1] / t[2] + 5e-6 / abs(t[2])   # |n.y| just below 1e-5
    res = mcp.find_scaled_batch([cam], [pose], world, obs, [0, 50], taps=True)
    ocost, oscale, ocand, ook, ocosts = cp.find_scaled_fast(ocam, world, obs, opose)
    assert not ook[7] and not ook[20]
    np.testing.assert_array_equal(res.cand_ok, ook)
    np.testing.assert_array_equal(res.cand_costs, ocosts)
    assert res.costs[0] == ocost and res.scales[0] == oscale


def test_injected_scales_bit_exact():
    """cvScaleCosts == avgReprojectionError for arbitrary (injected) scales, incl. 0, negatives, huge, inf, nan."""
    cam, ocam, pose, opose, world, obs, s = scene(6, 333, outlier_frac=0.1)
    rng = np.random.default_rng(1)
    scales = np.concatenate([[0.0, s, -s, 1e9, -1e9, math.inf, math.nan], rng.uniform(-3, 3, 300)])
    got = mcp.avgReprojectionError(scales, cam, (world, obs), pose)
    want = cp.avg_reprojection_errors(scales, ocam, world, obs, opose)
    np.testing.assert_array_equal(got, want)
    assert mcp.avgReprojectionError([], cam, (world, obs), pose).size == 0
    assert (mcp.avgReprojectionError([0.5, 1.0], cam, ([], []), pose) == math.inf).all()


def test_full_size_properties():
    """n = 5000 (the C4 correspondence count): 10 000 candidates x 5 000 observations per problem."""
    n, n_prob = 5000, 6
    scenes = [scene(40 + i, n, noise=2e-4) for i in range(n_prob)]
    off = np.arange(n_prob + 1, dtype=np.int64) * n
    res = mcp.find_scaled_batch([sc[0] for sc in scenes], [sc[2] for sc in scenes], np.concatenate([sc[4] for sc in scenes]),
                                np.concatenate([sc[5] for sc in scenes]), off, taps=True)
    for i, sc in enumerate(scenes):
        cam, ocam, pose, opose, world, obs, s = sc
        assert abs(res.scales[i] - s) < 0.01 * s                       # planted scale recovered
        assert res.poses[i].ScaleSign == 1
        cc = res.cand_costs[off[i]:off[i + 1]].ravel()
        k = int(np.argmin(cc))                                         # numpy argmin = first minimum
        assert res.costs[i] == cc[k] and res.scales[i] == res.candidates[off[i]:off[i + 1]].ravel()[k]
        again = mcp.avgReprojectionError([res.scales[i]], cam, (world, obs), pose)
        assert again[0] == res.costs[i]                                # idempotent through the injected entry point
        some = np.array([0, 1, 4999, 5000, 9999])                      # a few candidates against the oracle
        np.testing.assert_array_equal(
            cc[some], cp.avg_reprojection_errors(res.candidates[off[i]:off[i + 1]].ravel()[some], ocam, world, obs, opose))
    # single call == batch
    cam, ocam, pose, opose, world, obs, s = scenes[0]
    cost, out = mcp.findScaled(0.01, cam, (world, obs), pose)
    assert cost == res.costs[0] and np.array_equal(out.Translation, res.poses[0].Translation)


def test_device_resident_form():
    import torch
    n_prob, n = 5, 800
    scenes = [scene(60 + i, n) for i in range(n_prob)]
    off = np.arange(n_prob + 1, dtype=np.int64) * n
    dev = torch.device("cuda:0")
    cams = torch.tensor(np.stack([sc[0].row() for sc in scenes]), device=dev)
    rot = torch.tensor(np.stack([sc[2].Rotation.ravel() for sc in scenes]), device=dev)
    trans = torch.tensor(np.stack([sc[2].Translation for sc in scenes]), device=dev)
    world = torch.tensor(np.concatenate([sc[4] for sc in scenes]), device=dev)
    obs = torch.tensor(np.concatenate([sc[5] for sc in scenes]), device=dev)
    costs = torch.zeros(n_prob, dtype=torch.float64, device=dev)
    scales = torch.zeros(n_prob, dtype=torch.float64, device=dev)
    st = torch.zeros(n_prob, 3, dtype=torch.float64, device=dev)
    signs = torch.zeros(n_prob, dtype=torch.int32, device=dev)
    torch.cuda.synchronize()
    mcp.find_scaled_batch_device(cams, rot, trans, world, obs, off, costs, scales, st, signs, device=0,
                                 stream=torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    host = mcp.find_scaled_batch([sc[0] for sc in scenes], [sc[2] for sc in scenes], world.cpu().numpy(), obs.cpu().numpy(), off)
    np.testing.assert_array_equal(costs.cpu().numpy(), host.costs)
    np.testing.assert_array_equal(scales.cpu().numpy(), host.scales)
    np.testing.assert_array_equal(st.cpu().numpy(), np.stack([p.Translation for p in host.poses]))
    assert (signs.cpu().numpy() == 1).all()


def test_error_convention():
    cam, ocam, pose, opose, world, obs, s = scene(7, 10)
    from minicv_b200 import native
    lib = native.load()
    assert math.isnan(lib.cvFindScaled(0.01, None, None, None, 3, None, None))
    assert lib.cvFindScaledBatch(None, None, None, None, None, 2, None, None, None, None, None, None, -1) == -1
    assert lib.cvScaleCosts(None, None, None, None, 1, None, 1, None) == -1


def test_non_finite_and_extreme_inputs_bit_exact():
    """NaN / inf / huge / tiny world points and observations: the shared-reciprocal division falls back to the generic one
    exactly where ptxas' own sequence would, so candidates and costs still equal the oracle's bit for bit (variant 0: only
    world-side specials, which project1 rejects -> finite costs; variant 1: observation-side specials as well, which
    poison the sums with inf / NaN exactly as the reference's `sum <- sum + d` would)."""
    cam, ocam, pose, opose, world0, obs0, s = scene(8, 200)
    for variant in (0, 1):
        world, obs = world0.copy(), obs0.copy()
        world[3] = [np.nan, 0.0, 1.0]
        world[10] *= 1e200
        world[11] *= 1e-200
        world[12] = [np.inf, 1.0, -np.inf]
        world[30] = ocam.location            # pc = 0: 0 / 0
        obs[21] = [1e-310, -1e-310]          # denormals
        if variant:
            obs[20] = [np.inf, 0.1]
            obs[22] = [np.nan, np.nan]
        with np.errstate(all="ignore"):
            res = mcp.find_scaled_batch([cam], [pose], world, obs, [0, 200], taps=True)
            ocost, oscale, ocand, ook, ocosts = cp.find_scaled_fast(ocam, world, obs, opose)
            np.testing.assert_array_equal(res.cand_ok, ook)
            np.testing.assert_array_equal(res.candidates[ook], ocand[ook])
            np.testing.assert_array_equal(res.cand_costs, ocosts)
            assert (res.costs[0] == ocost or (np.isnan(res.costs[0]) and np.isnan(ocost))) and res.scales[0] == oscale
            if not variant:
                assert np.isfinite(ocosts[ook]).mean() > 0.5
            scales = np.array([0.0, s, 1e300, -1e300, 1e-300, np.inf, -np.inf, np.nan])
            np.testing.assert_array_equal(mcp.avgReprojectionError(scales, cam, (world, obs), pose),
                                          cp.avg_reprojection_errors(scales, ocam, world, obs, opose))
