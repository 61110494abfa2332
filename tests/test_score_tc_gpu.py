"""GPU: the tensor-core scoring engine of the F / E RANSAC stage (csrc/score_tc.cu, filter0.cuh).  Level 0 -- the
epipolar residuals of a tile as a TF32 GEMM and one comparison per evaluation -- may only reject what the reference's
exact test rejects, so every inlier count must equal the oracle's and the fp32-filter kernel's: adversarial injected
models and coordinates (the cases of test_filter_gpu.py), whole RANSAC batches compared count by count between the two
engines, and an audit pass that recomputes every level-0 value in fp64 and reports the worst |error| / budget."""
import numpy as np
import pytest

from minicv_b200 import estimation as est
from minicv_b200 import synth
from oracle import essential, ransac

from test_filter_gpu import K_TWO_VIEW, adversarial, counts_device, gt_essential, point_variants

pytestmark = pytest.mark.gpu


@pytest.fixture
def tensor_engine():
    est.set_score_engine(est.SCORE_TENSOR)
    yield
    est.set_score_engine(est.SCORE_AUTO)


def test_fundamental_level0_never_changes_a_count(tensor_engine):
    rng = np.random.default_rng(2)
    a, b, gt, R, t = synth.two_view_pair(22, 3000)
    Ki = np.linalg.inv(K_TWO_VIEW)
    F = Ki.T @ gt_essential(R, t) @ Ki
    M = adversarial(F / F[2, 2], rng, n_random=300)
    with np.errstate(all="ignore"):
        for name, aa, bb in point_variants(a, b, rng):
            m1, m2 = ransac.FundamentalCB.prepare(aa, bb)
            want = np.array([ransac.find_inliers(ransac.FundamentalCB, m1, m2, m.reshape(3, 3), 3.0)[0] for m in M])
            np.testing.assert_array_equal(counts_device(est.MODEL_F, aa, bb, M, 3.0), want, err_msg=name)
        # far off-centre pair: the exact test's own fp64 noise must widen the level-0 threshold, not break it
        for shift in (1e4, 1e7, 1e10):
            aa, bb = a + shift, b - 0.5 * shift
            m1, m2 = ransac.FundamentalCB.prepare(aa, bb)
            want = np.array([ransac.find_inliers(ransac.FundamentalCB, m1, m2, m.reshape(3, 3), 3.0)[0] for m in M])
            np.testing.assert_array_equal(counts_device(est.MODEL_F, aa, bb, M, 3.0), want, err_msg=f"shift {shift}")
        inl = gt.astype(bool)
        m1, m2 = ransac.FundamentalCB.prepare(a[inl], b[inl])
        want = np.array([ransac.find_inliers(ransac.FundamentalCB, m1, m2, m.reshape(3, 3), 3.0)[0] for m in M])
        np.testing.assert_array_equal(counts_device(est.MODEL_F, a[inl], b[inl], M, 3.0), want)
        assert want[0] > 0.9 * inl.sum()
        for thr in (0.05, 30.0):
            m1, m2 = ransac.FundamentalCB.prepare(a, b)
            want = np.array([ransac.find_inliers(ransac.FundamentalCB, m1, m2, m.reshape(3, 3), thr)[0] for m in M])
            np.testing.assert_array_equal(counts_device(est.MODEL_F, a, b, M, thr), want, err_msg=f"thr {thr}")


def test_homography_level0_never_changes_a_count(tensor_engine):
    rng = np.random.default_rng(1)
    a, b, gt, H = synth.homography_pair(21, 3000)
    M = adversarial(H / H[2, 2], rng, n_random=300, fix33=True)
    with np.errstate(all="ignore"):
        for name, aa, bb in point_variants(a, b, rng):
            m1, m2 = ransac.HomographyCB.prepare(aa, bb)
            want = np.array([ransac.find_inliers(ransac.HomographyCB, m1, m2, m.reshape(3, 3), 3.0)[0] for m in M])
            np.testing.assert_array_equal(counts_device(est.MODEL_H, aa, bb, M, 3.0), want, err_msg=name)
        for shift in (1e4, 1e6):
            aa, bb = a + shift, b - 0.5 * shift
            m1, m2 = ransac.HomographyCB.prepare(aa, bb)
            want = np.array([ransac.find_inliers(ransac.HomographyCB, m1, m2, m.reshape(3, 3), 3.0)[0] for m in M])
            np.testing.assert_array_equal(counts_device(est.MODEL_H, aa, bb, M, 3.0), want, err_msg=f"shift {shift}")
        inl = gt.astype(bool)
        m1, m2 = ransac.HomographyCB.prepare(a[inl], b[inl])
        want = np.array([ransac.find_inliers(ransac.HomographyCB, m1, m2, m.reshape(3, 3), 3.0)[0] for m in M])
        np.testing.assert_array_equal(counts_device(est.MODEL_H, a[inl], b[inl], M, 3.0), want)
        assert want[0] == inl.sum()
        for thr in (0.05, 30.0):
            m1, m2 = ransac.HomographyCB.prepare(a, b)
            want = np.array([ransac.find_inliers(ransac.HomographyCB, m1, m2, m.reshape(3, 3), thr)[0] for m in M])
            np.testing.assert_array_equal(counts_device(est.MODEL_H, a, b, M, thr), want, err_msg=f"thr {thr}")


def test_essential_level0_never_changes_a_count(tensor_engine):
    rng = np.random.default_rng(3)
    a, b, gt, R, t = synth.two_view_pair(23, 3000)
    E = gt_essential(R, t)
    M = adversarial(E / np.linalg.norm(E), rng, n_random=300)
    thr_n = 1.0 / 800.0
    with np.errstate(all="ignore"):
        for name, aa, bb in point_variants(a, b, rng):
            m1 = essential.normalize_points(aa, 800.0, (320.0, 240.0))
            m2 = essential.normalize_points(bb, 800.0, (320.0, 240.0))
            want = np.array([ransac.find_inliers(essential.EssentialCB, m1, m2, m.reshape(3, 3), thr_n)[0] for m in M])
            got = counts_device(est.MODEL_E, aa, bb, M, 1.0, focal=800.0, pp=(320.0, 240.0))
            np.testing.assert_array_equal(got, want, err_msg=name)
        inl = gt.astype(bool)
        m1 = essential.normalize_points(a[inl], 800.0, (320.0, 240.0))
        m2 = essential.normalize_points(b[inl], 800.0, (320.0, 240.0))
        want = np.array([ransac.find_inliers(essential.EssentialCB, m1, m2, m.reshape(3, 3), thr_n)[0] for m in M])
        got = counts_device(est.MODEL_E, a[inl], b[inl], M, 1.0, focal=800.0, pp=(320.0, 240.0))
        np.testing.assert_array_equal(got, want)
        assert want[0] > 0.9 * inl.sum()


def _batch(model, n_pairs, n, iters, **kw):
    A = np.empty((n_pairs * n, 2)); B = np.empty((n_pairs * n, 2))
    for p in range(n_pairs):
        r = synth.homography_pair(100 + p, n) if model == est.MODEL_H else synth.two_view_pair(100 + p, n)
        A[p * n:(p + 1) * n] = r[0]; B[p * n:(p + 1) * n] = r[1]
    off = np.arange(n_pairs + 1, dtype=np.int64) * n
    return est.ransac_batch(model, A, B, off, kw.pop("thr"), iters, 0.99, sample_mode=est.SAMPLE_COUNTER,
                            stop_mode=est.STOP_FIXED, seed=5, debug=True, **kw)


@pytest.mark.parametrize("model,n,iters,kw", [
    (est.MODEL_H, 2000, 2048, dict(thr=3.0, refine=False)),
    (est.MODEL_H, 501, 300, dict(thr=1.0)),
    (est.MODEL_F, 2000, 1024, dict(thr=3.0)),
    (est.MODEL_F, 777, 300, dict(thr=1.0)),                       # ragged: last stage partly filled, last tile partly filled
    (est.MODEL_E, 5000, 256, dict(thr=1.0, focal=800.0, pp=(320.0, 240.0))),
    (est.MODEL_E, 333, 100, dict(thr=2.0, focal=800.0, pp=(320.0, 240.0))),
])
def test_engines_agree_count_by_count(model, n, iters, kw):
    """Every hypothesis of a RANSAC batch (counter-based sampling, the solver's own models) gets the same inlier count
    from both scoring engines, and the batch returns the same models, masks and counts."""
    est.set_score_engine(est.SCORE_FP32_FILTER)
    try:
        ref = _batch(model, 6, n, iters, **dict(kw))
        est.set_score_engine(est.SCORE_TENSOR)
        est.score_audit(True)
        got = _batch(model, 6, n, iters, **dict(kw))
        worst = est.score_audit(False)
    finally:
        est.set_score_engine(est.SCORE_AUTO)
    np.testing.assert_array_equal(got.dbg_counts, ref.dbg_counts)
    np.testing.assert_array_equal(got.masks, ref.masks)
    np.testing.assert_array_equal(got.inliers, ref.inliers)
    np.testing.assert_array_equal(got.models, ref.models)
    assert (ref.dbg_counts >= 0).sum() > 6 * iters * 0.9
    # the audit recomputed the level-0 values in fp64: the arithmetic error stayed inside its budget with room to spare
    assert 0 < worst < 0.5, worst
    print(f"level-0 audit, model {model}, n {n}: worst |error| / budget = {worst:.3f}")


@pytest.mark.parametrize("model,kw", [
    (est.MODEL_H, dict(thr=2.0)),
    (est.MODEL_F, dict(thr=2.0)),
    (est.MODEL_E, dict(thr=1.5, focal=800.0, pp=(320.0, 240.0))),
])
def test_ragged_batch_and_chunk_seams(model, kw, monkeypatch):
    """Pairs of very different sizes in one batch (empty, below the sample size, below / at / just above one stage of 64,
    several stages), processed in one chunk and in chunks of two pairs (every buffer of the tensor-core engine is addressed
    relative to the chunk's first pair): both engines return the same counts, models and masks, chunked or not."""
    sizes = [2000, 0, 5, 64, 65, 777, 63, 1, 130, 1024, 8]
    offs = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)
    A = np.empty((offs[-1], 2)); B = np.empty((offs[-1], 2))
    for p, n in enumerate(sizes):
        if n:
            r = synth.homography_pair(200 + p, n) if model == est.MODEL_H else synth.two_view_pair(200 + p, n)
            A[offs[p]:offs[p + 1]] = r[0]; B[offs[p]:offs[p + 1]] = r[1]

    def run(engine, chunk):
        if chunk:
            monkeypatch.setenv("MINICV_RANSAC_CHUNK_PAIRS", str(chunk))
        else:
            monkeypatch.delenv("MINICV_RANSAC_CHUNK_PAIRS", raising=False)
        est.set_score_engine(engine)
        try:
            return est.ransac_batch(model, A, B, offs, kw["thr"], 200, 0.99, sample_mode=est.SAMPLE_COUNTER,
                                    stop_mode=est.STOP_FIXED, seed=9, debug=True,
                                    **{k: v for k, v in kw.items() if k != "thr"})
        finally:
            est.set_score_engine(est.SCORE_AUTO)

    ref = run(est.SCORE_FP32_FILTER, 0)
    for engine, chunk in ((est.SCORE_TENSOR, 0), (est.SCORE_TENSOR, 2), (est.SCORE_FP32_FILTER, 3), (est.SCORE_TENSOR, 1)):
        got = run(engine, chunk)
        np.testing.assert_array_equal(got.dbg_counts, ref.dbg_counts, err_msg=f"engine {engine} chunk {chunk}")
        np.testing.assert_array_equal(got.masks, ref.masks)
        np.testing.assert_array_equal(got.ok, ref.ok)
        np.testing.assert_array_equal(got.models, ref.models)
    assert ref.ok[0] and not ref.ok[1] and not ref.ok[7]                  # 2000, 0 and 1 correspondences


@pytest.mark.parametrize("inlier_frac", [0.3, 0.9])
def test_auto_mode_selects_on_the_device(inlier_frac):
    """Auto mode on a batch large enough for the tensor-core engine: a pilot kernel looks at the first models of every pair
    and the gate lets exactly one of the two scoring kernels run (level 0 when almost every model is wrong, the fp32-filter
    kernel when many models fit, as on ratio-tested matches).  Whichever runs, the results are those of either engine."""
    n_pairs, n, iters = 40, 600, 2048
    A = np.empty((n_pairs * n, 2)); B = np.empty((n_pairs * n, 2))
    for p in range(n_pairs):
        r = synth.homography_pair(300 + p, n, inlier_frac=inlier_frac)
        A[p * n:(p + 1) * n] = r[0]; B[p * n:(p + 1) * n] = r[1]
    off = np.arange(n_pairs + 1, dtype=np.int64) * n
    out = {}
    for engine in (est.SCORE_FP32_FILTER, est.SCORE_TENSOR, est.SCORE_AUTO):
        est.set_score_engine(engine)
        try:
            out[engine] = est.ransac_batch(est.MODEL_H, A, B, off, 3.0, iters, 0.995, sample_mode=est.SAMPLE_COUNTER,
                                           stop_mode=est.STOP_FIXED, seed=3, debug=True)
        finally:
            est.set_score_engine(est.SCORE_AUTO)
    ref = out[est.SCORE_FP32_FILTER]
    for engine in (est.SCORE_TENSOR, est.SCORE_AUTO):
        np.testing.assert_array_equal(out[engine].dbg_counts, ref.dbg_counts)
        np.testing.assert_array_equal(out[engine].masks, ref.masks)
        np.testing.assert_array_equal(out[engine].models, ref.models)
    assert ref.ok.all() and abs(ref.inliers.mean() / n - inlier_frac) < 0.1


@pytest.mark.parametrize("model,kw,inlier_frac,n_pairs", [
    (est.MODEL_H, dict(thr=3.0), 0.8, 40),
    (est.MODEL_H, dict(thr=3.0), 0.3, 40),
    (est.MODEL_H, dict(thr=3.0), 0.6, 3),
    (est.MODEL_F, dict(thr=2.0), None, 24),
    (est.MODEL_E, dict(thr=1.0, focal=800.0, pp=(320.0, 240.0)), None, 8),
])
def test_two_phase_adaptive_scoring_equals_one_pass(model, kw, inlier_frac, n_pairs):
    """With OpenCV's adaptive stop the engine scores the first 128 iterations, replays the accept / RANSACUpdateNumIters rule
    over them, cuts every pair's iteration count to the horizon that leaves and scores only what is below it.  The result
    must be what one pass over all hypotheses gives (requesting the debug taps switches the two phases off): models, masks,
    counts, the iteration the winner came from and the number of iterations OpenCV's loop would have run."""
    n, iters = 700, 1024
    A = np.empty((n_pairs * n, 2)); B = np.empty((n_pairs * n, 2))
    for p in range(n_pairs):
        r = synth.homography_pair(400 + p, n, inlier_frac=inlier_frac) if model == est.MODEL_H else synth.two_view_pair(400 + p, n)
        A[p * n:(p + 1) * n] = r[0]; B[p * n:(p + 1) * n] = r[1]
    off = np.arange(n_pairs + 1, dtype=np.int64) * n
    args = dict(sample_mode=est.SAMPLE_MWC, stop_mode=est.STOP_ADAPTIVE, **{k: v for k, v in kw.items() if k != "thr"})
    for engine in (est.SCORE_AUTO, est.SCORE_FP32_FILTER, est.SCORE_TENSOR):
        est.set_score_engine(engine)
        try:
            one = est.ransac_batch(model, A, B, off, kw["thr"], iters, 0.995, debug=True, **args)
            two = est.ransac_batch(model, A, B, off, kw["thr"], iters, 0.995, debug=False, **args)
        finally:
            est.set_score_engine(est.SCORE_AUTO)
        np.testing.assert_array_equal(two.models, one.models)
        np.testing.assert_array_equal(two.masks, one.masks)
        np.testing.assert_array_equal(two.inliers, one.inliers)
        np.testing.assert_array_equal(two.best, one.best)
        np.testing.assert_array_equal(two.iters_run, one.iters_run)
        np.testing.assert_array_equal(two.ok, one.ok)
    if inlier_frac == 0.8:
        assert (one.iters_run < 64).all()              # the horizon really is short here: phase 2 had nothing to do
