"""GPU: device-resident pipeline match -> gather -> homography RANSAC (no host round trip of the
data), checked stage by stage against the oracle."""
import numpy as np
import pytest

import minicv_b200 as mcv
from minicv_b200 import estimation as est
from minicv_b200 import synth
from oracle import bf_knn, ransac

pytestmark = pytest.mark.gpu


def test_device_pipeline_matches_oracle_chain():
    import torch
    dev = torch.device("cuda:0")
    P, nq, nt = 3, 700, 900
    Q, T, KQ, KT = [], [], [], []
    for p in range(P):
        q, t, gt, kq, kt, H = synth.descriptor_pair_geo(70 + p, nq, nt)
        Q.append(q); T.append(t); KQ.append(kq); KT.append(kt)
    qo = np.arange(P + 1) * nq; to = np.arange(P + 1) * nt
    dq = torch.from_numpy(np.concatenate(Q)).to(dev); dt = torch.from_numpy(np.concatenate(T)).to(dev)
    dkq = torch.from_numpy(np.concatenate(KQ)).to(dev); dkt = torch.from_numpy(np.concatenate(KT)).to(dev)
    didx = torch.empty((P * nq, 2), dtype=torch.int32, device=dev)
    ddist = torch.empty((P * nq, 2), dtype=torch.float32, device=dev)
    dok = torch.empty(P * nq, dtype=torch.uint8, device=dev)
    s = torch.cuda.current_stream().cuda_stream
    mcv.match_knn2_batch_device(dq, qo, dt, to, 32, 0.8, didx, ddist, dok, stream=s)
    da = torch.empty((P * nq, 2), dtype=torch.float64, device=dev)
    db = torch.empty((P * nq, 2), dtype=torch.float64, device=dev)
    off = mcv.gather_matches_device(dkq, dkt, qo, to, didx, dok, da, db, stream=s)
    n = int(off[-1])
    dM = torch.empty((P, 9), dtype=torch.float64, device=dev)
    dMask = torch.empty(max(n, 1), dtype=torch.uint8, device=dev)
    dCnt = torch.empty(P, dtype=torch.int32, device=dev); dOk = torch.empty(P, dtype=torch.uint8, device=dev)
    est.ransac_batch_device(est.MODEL_H, da, db, off, 3.0, 2000, dM, dMask, dCnt, dOk, confidence=0.995,
                            sample_mode=est.SAMPLE_MWC, stop_mode=est.STOP_ADAPTIVE, stream=s)
    torch.cuda.synchronize()
    a_all, b_all = da.cpu().numpy(), db.cpu().numpy()
    for p in range(P):
        idx, dist = bf_knn.bf_knn2_hamming(Q[p], T[p])
        ok = bf_knn.ratio_test(dist, idx, 0.8).astype(bool)
        a = KQ[p][ok]; b = KT[p][idx[ok, 0]]
        assert off[p + 1] - off[p] == ok.sum()
        np.testing.assert_array_equal(a_all[off[p]:off[p + 1]], a)        # gather keeps query order
        np.testing.assert_array_equal(b_all[off[p]:off[p + 1]], b)
        Ho, mo = ransac.find_homography(a, b, 3.0, 2000, 0.995)
        Hg = dM[p].cpu().numpy().reshape(3, 3)
        assert dOk[p].item() == 1
        assert np.abs(Hg - Ho).max() / np.abs(Ho).max() < 1e-6
        assert (dMask[off[p]:off[p + 1]].cpu().numpy() == mo).mean() >= 0.999


@pytest.mark.parametrize("engine", [mcv.ENGINE_POPC, mcv.ENGINE_TENSOR])
def test_match_homography_one_call_equals_oracle_chain(engine):
    """cvMatchHomographyBatch: host buffers in, H + counts (+ match lists, inlier masks) out, the hand-off on the
    device; every stage equals the oracle chain cv2-style: knnMatch + ratio -> matched keypoints -> findHomography."""
    P = 5
    sizes = [(700, 900), (300, 250), (0, 40), (900, 700), (64, 3)]
    Q, T, KQ, KT = [], [], [], []
    for p, (nq, nt) in enumerate(sizes):
        q, t, gt, kq, kt, H = synth.descriptor_pair_geo(170 + p, max(nq, 1), max(nt, 1))
        Q.append(q[:nq]); T.append(t[:nt]); KQ.append(kq[:nq]); KT.append(kt[:nt])
    qo = np.concatenate([[0], np.cumsum([s[0] for s in sizes])]); to = np.concatenate([[0], np.cumsum([s[1] for s in sizes])])
    mcv.set_hamming_engine(engine)
    try:
        r = est.match_homography_batch(np.concatenate(Q), qo, np.concatenate(T), to, np.concatenate(KQ), np.concatenate(KT),
                                       0.8, 3.0, 2000, 0.995, want_matches=True)
        rm = est.match_homography_batch(np.concatenate(Q), qo, np.concatenate(T), to, np.concatenate(KQ), np.concatenate(KT),
                                        0.8, 3.0, 2000, 0.995, device_mask=0)
    finally:
        mcv.set_hamming_engine(mcv.ENGINE_AUTO)
    co = np.concatenate([[0], np.cumsum(r.n_matches)])
    for p, (nq, nt) in enumerate(sizes):
        idx, dist = bf_knn.bf_knn2_hamming(Q[p], T[p])
        okq = bf_knn.ratio_test(dist, idx, 0.8).astype(bool)
        assert r.n_matches[p] == okq.sum()
        np.testing.assert_array_equal(r.match_query[co[p]:co[p + 1]], np.nonzero(okq)[0])
        np.testing.assert_array_equal(r.match_train[co[p]:co[p + 1]], idx[okq, 0])
        a = KQ[p][okq]; b = KT[p][idx[okq, 0]]
        if okq.sum() < 4:
            assert r.ok[p] == 0
            continue
        Ho, mo = ransac.find_homography(a, b, 3.0, 2000, 0.995)
        if Ho is None:
            assert r.ok[p] == 0
            continue
        assert r.ok[p] == 1
        assert np.abs(r.H[p] - Ho).max() / np.abs(Ho).max() < 1e-6
        assert (r.inlier_mask[co[p]:co[p + 1]] == mo).mean() >= 0.999
        assert abs(int(r.n_inliers[p]) - int(mo.sum())) <= max(1, int(0.001 * len(mo)))
    np.testing.assert_array_equal(rm.ok, r.ok)
    np.testing.assert_array_equal(rm.H, r.H)
    np.testing.assert_array_equal(rm.n_matches, r.n_matches)
    np.testing.assert_array_equal(rm.n_inliers, r.n_inliers)
