"""Pin the C restatement (oracle/rsdet_oracle.c) to the reference's own source compiled for the host
(oracle/_ref, built by oracle/build_ref.py from the strings inside /root/reference/python/jdet/ops/*.py).
Bit-for-bit: any difference is a failure."""
import numpy as np
import pytest

import workloads as W


def _pairs(seed=0, n=1500, g=192):
    P = W.rotated_boxes(n, seed)
    G = W.jittered_copies(P, g, seed + 1)
    return G, P


@pytest.mark.parametrize("version", [0, 1])
@pytest.mark.parametrize("cudasort", [False, True])
def test_iou_bit_exact(oracle, ref, version, cudasort):
    G, P = _pairs()
    a = oracle.box_iou_rotated(G, P, version, int(cudasort))
    r = ref.box_iou(G, P, version, cudasort)
    assert (a > 0.5).sum() > 50 and (a == 0).mean() > 0.5  # the sample spans both regimes
    assert np.array_equal(a, r)


def test_iou_known_answer(oracle, ref):
    # the reference's own self-test boxes, box_iou_rotated.py:513-514 -> [[1,0.2],[0.2,1]]
    b = np.array([[0, 0, 1, 1, 0], [0.5, 0.5, 1, 2, 0]], np.float32)
    for f in (lambda: oracle.box_iou_rotated(b, b), lambda: ref.box_iou(b, b)):
        np.testing.assert_allclose(f(), [[1, 0.2], [0.2, 1]], rtol=0, atol=1e-7)


def test_iou_degenerate_inputs(oracle, ref):
    b = np.array([[5, 5, 4, 2, 0.3], [5, 5, 4, 2, 0.3], [5, 5, 0, 2, 0], [5, 5, 1e-8, 1e-8, 0], [9, 5, 4, 2, 0.3],
                  [5, 5, 2, 4, 0.3 + np.pi / 2], [7, 5, 4, 2, 0.3], [5, 5, 4, 2, -0.3]], np.float32)
    for v in (0, 1):
        for cs in (0, 1):
            assert np.array_equal(oracle.box_iou_rotated(b, b, v, cs), ref.box_iou(b, b, v, bool(cs)))


@pytest.mark.parametrize("thr", [0.1, 0.5])
@pytest.mark.parametrize("ge", [False, True])
def test_nms_keep_exact(oracle, ref, thr, ge):
    n = 2500
    d = W.rotated_boxes(n, 3, smin=16, smax=128)
    s = W.distinct_scores(n, 3)
    order = np.argsort(-s.astype(np.float64), kind="stable").astype(np.int32)
    k = oracle.nms_rotated_keep(d, order, thr, 5, ge)
    kr = ref.nms_keep(d, order, thr, 5, ge)
    assert 0 < k.sum() < n
    assert np.array_equal(k, kr)


def test_ml_nms_keep_exact(oracle, ref):
    n = 2500
    d = W.rotated_boxes(n, 4, smin=16, smax=128)
    s = W.distinct_scores(n, 4)
    lab = np.random.default_rng(0).integers(0, 15, n)
    d6 = np.concatenate([d, lab[:, None].astype(np.float32)], 1)
    order = np.argsort(-s.astype(np.float64), kind="stable").astype(np.int32)
    assert np.array_equal(oracle.nms_rotated_keep(d6, order, 0.1, 6, False), ref.nms_keep(d6, order, 0.1, 6, False))


def test_nms_known_answer(oracle, ref):
    # nms_rotated.py:598-603 -> keep [2]
    dets = np.array([[0, 0, 1, 1, 0], [0, 0, 0.5, 0.5, 0.3], [0, 0, 0.9, 0.9, 0]], np.float32)
    scores = np.array([0.1, 0.2, 0.3], np.float32)
    assert oracle.nms_rotated(dets, scores, 0.3).tolist() == [2]
    assert oracle.ml_nms_rotated(dets, scores, np.array([1, 1, 1]), 0.3).tolist() == [2]
    order = np.argsort(-scores).astype(np.int32)
    assert np.nonzero(ref.nms_keep(dets, order, 0.3, 5, False))[0].tolist() == [2]


@pytest.mark.parametrize("version", [0, 1])
@pytest.mark.parametrize("sampling_ratio", [2, 0])
def test_roi_align_bit_exact(oracle, ref, version, sampling_ratio):
    rng = np.random.default_rng(0)
    feat = rng.standard_normal((2, 8, 64, 64)).astype(np.float32)
    rois = W.proposals(60, 0, batch=2)
    rois[:4, 3:5] = [[0.5, 0.5], [3000, 20], [20, 3000], [1, 900]]  # clamped / huge / out-of-map RoIs
    rois[4, 1:3] = [-40, 1100]
    a = oracle.roi_align_rotated_fwd(feat, rois, (7, 7), 1 / 16, sampling_ratio, version)
    r = ref.roi_fwd(feat, rois, (7, 7), 1 / 16, sampling_ratio, version)
    assert np.array_equal(a, r)
    g = rng.standard_normal(a.shape).astype(np.float32)
    ab = oracle.roi_align_rotated_bwd(g, rois, feat.shape, 1 / 16, sampling_ratio, version)
    rb = ref.roi_bwd(g, rois, feat.shape, 1 / 16, sampling_ratio, version)
    assert np.array_equal(ab, rb)


def test_poly_iou_bit_exact(oracle, ref):
    pp = oracle.obb2poly(W.rotated_boxes(300, 5, smin=16, smax=128))
    a = oracle.poly_iou_matrix(pp, pp)
    r = ref.poly_iou_matrix(pp, pp)
    assert (a > 0.3).sum() > 300
    assert np.array_equal(a, r)
    # clockwise / counter-clockwise vertex order and far-from-origin quads (multiclass offsets)
    q = pp[:, [0, 1, 6, 7, 4, 5, 2, 3]] + np.float32(7000)
    assert np.array_equal(oracle.poly_iou_matrix(q, pp + np.float32(7000)), ref.poly_iou_matrix(q, pp + np.float32(7000)))


def test_poly_nms_exact(oracle, ref):
    n = 600
    pp = oracle.obb2poly(W.rotated_boxes(n, 6, smin=16, smax=128))
    s = W.distinct_scores(n, 6)
    order = np.argsort(-s.astype(np.float64), kind="stable")
    b9 = np.concatenate([pp, s[:, None]], 1)
    keep = oracle.poly_nms(b9, 0.1)
    kr = ref.poly_nms_sorted_keep(b9[order], 0.1)
    assert np.array_equal(keep, order[kr])


def test_fma_contraction_stays_inside_band(oracle, ref):
    """nvcc contracts a*b+c by default, g++ does not: the reference's CUDA and CPU paths differ by < 1e-6
    (SURVEY appendix A), which is where the 1e-6 exclusion band of the parity contract comes from."""
    G, P = _pairs(7)
    a = ref.box_iou(G, P, 1, False, fma=False)
    b = ref.box_iou(G, P, 1, False, fma=True)
    assert np.abs(a - b).max() < 1e-6
