"""CPU: the scale-search oracle (oracle/camera_pose.py, a restatement of CameraPose.findScaled,
src/MiniCV/CameraPose.fs:39-134) -- internal consistency (parity is UNPINNED: the reference holds no test or
fixture for it and F# cannot run here) -- and the device math header scalemath.cuh compiled for the host against it."""
import ctypes as C
import math
import subprocess

import numpy as np
import pytest

from minicv_b200 import synth
from oracle import camera_pose as cp
from conftest import ROOT


def scene(seed, n, **kw):
    row, R, T, world, obs, s = synth.scale_scene(seed, n, **kw)
    cam = cp.Camera(row[0:3], row[3:6], row[6:9], row[9:12], row[12:14])
    return row, cam, cp.CameraPose(2, 1, R, T, False), world, obs, s


def test_planted_scale_is_recovered():
    for seed in range(4):
        row, cam, pose, world, obs, s = scene(seed, 60, noise=0.0, outlier_frac=0.0)
        cost, out = cp.find_scaled(0.01, cam, world, obs, pose)
        assert cost < 1e-20
        np.testing.assert_allclose(out.Translation, s * pose.Translation, rtol=1e-9)
        assert out.ScaleSign == 1 and out.RotationIndex == 2
    # with noise the best candidate is near the planted scale (no outliers: the reference's mean over "the points
    # that project" is not robust -- its own "todo: remove outliers", CameraPose.fs:40)
    row, cam, pose, world, obs, s = scene(9, 150, noise=1e-3, outlier_frac=0.0)
    cost, out = cp.find_scaled(0.01, cam, world, obs, pose)
    assert abs(out.Translation @ pose.Translation - s) < 0.02 * s and cost < 1e-5


def test_fast_form_equals_the_loop():
    row, cam, pose, world, obs, s = scene(3, 90)
    cost, out = cp.find_scaled(0.01, cam, world, obs, pose)
    c2, s2, cand, ok, costs = cp.find_scaled_fast(cam, world, obs, pose)
    assert c2 == cost and np.array_equal(s2 * pose.Translation, out.Translation)
    some = cand[ok].ravel()[:25]
    loop = np.array([cp.avg_reprojection_error(x, cam, world, obs, pose) for x in some])
    np.testing.assert_array_equal(cp.avg_reprojection_errors(some, cam, world, obs, pose), loop)


def test_edge_cases():
    row, cam, pose, world, obs, s = scene(1, 20)
    cost, out = cp.find_scaled(0.01, cam, [], [], pose)
    assert cost == math.inf and out.Rotation is None and out.ScaleSign == 0          # CameraPose()
    # nothing projects (points behind every candidate camera) -> +inf, scale 0
    behind = world - 100.0 * cam.forward
    cost, out = cp.find_scaled(0.01, cam, behind, obs, pose)
    assert cost == math.inf and out.ScaleSign == 0 and not out.Translation.any()
    # scale sign follows the sign of the best scale
    neg = cp.CameraPose(1, 1, pose.Rotation, -pose.Translation, False)
    cost, out = cp.find_scaled(0.01, cam, world, obs, neg)
    assert out.ScaleSign == -1
    # rigid-basis inverse agrees with LAPACK
    b = cp._from_basis(cam.right, cam.up, -cam.forward, cam.location)
    np.testing.assert_allclose(cp.lu_inverse4(b), np.linalg.inv(b), atol=1e-14)
    assert not cp.lu_inverse4(np.zeros((4, 4))).any()


@pytest.fixture(scope="module")
def hm():
    out = ROOT / "build" / "libmcv_hostmath_test.so"
    out.parent.mkdir(exist_ok=True)
    src = ROOT / "tests" / "hostmath" / "hostmath.cpp"
    subprocess.run(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-shared", "-fPIC", "-x", "c++",
                    "-o", str(out), str(src)], check=True)
    return C.CDLL(str(out))


def p(a):
    return a.ctypes.data_as(C.c_void_p)


def test_scalemath_header_bit_exact(hm):
    """minicv_b200/csrc/scalemath.cuh (the source the kernels include) == the oracle, bit for bit."""
    rng = np.random.default_rng(0)
    for seed in range(3):
        row, cam, pose, world, obs, s = scene(seed, 70)
        R, T = np.ascontiguousarray(pose.Rotation), np.ascontiguousarray(pose.Translation)
        inv = np.zeros(16)
        b = cp._from_basis(cam.right, cam.up, -cam.forward, cam.location)
        assert hm.hm_lu_inverse4(p(np.ascontiguousarray(b)), p(inv)) == 1
        np.testing.assert_array_equal(inv.reshape(4, 4), cp.lu_inverse4(b))
        for sc in (0.0, s, -1.7, 1e6):
            out = np.zeros(14)
            hm.hm_camera_for_scale(p(row), p(R), p(T), C.c_double(sc), p(out))
            ref = cp.transformed_view(cp.transformation(cp.scale(sc, pose)), cam)
            np.testing.assert_array_equal(out, np.concatenate([ref.location, ref.forward, ref.up, ref.right, ref.focal]))
        cand, ok = np.zeros((len(world), 2)), np.zeros(len(world), np.uint8)
        hm.hm_candidate_scales(p(row), p(R), p(T), p(world), p(obs), len(world), p(cand), p(ok))
        rc, rok = cp.candidate_scales(cam, world, obs, pose)
        np.testing.assert_array_equal(ok.astype(bool), rok)
        np.testing.assert_array_equal(cand[rok], rc[rok])
        scales = np.concatenate([rc[rok].ravel()[:40], rng.uniform(-2, 2, 10)])
        costs = np.zeros(len(scales))
        hm.hm_scale_costs(p(row), p(R), p(T), p(world), p(obs), len(world), p(scales), len(scales), p(costs))
        np.testing.assert_array_equal(costs, cp.avg_reprojection_errors(scales, cam, world, obs, pose))
