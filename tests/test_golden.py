"""The oracle against the committed golden vectors (tests/golden/rotated_path_golden.npz, produced from
the reference's own source by tests/golden/make_golden.py).  Runs anywhere (no reference tree needed)."""
import os

import numpy as np
import pytest

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "rotated_path_golden.npz")


@pytest.fixture(scope="module")
def g():
    return dict(np.load(GOLD))


def test_iou(oracle, g):
    for v in (0, 1):
        for sort_kind in (0, 1):
            assert np.array_equal(oracle.box_iou_rotated(g["iou_boxes1"], g["iou_boxes2"], v, sort_kind), g[f"iou_v{v}"])
    np.testing.assert_allclose(g["iou_kat"], [[1, 0.2], [0.2, 1]], atol=1e-7)
    assert np.array_equal(oracle.box_iou_rotated(g["iou_kat_boxes"], g["iou_kat_boxes"]), g["iou_kat"])


def test_nms(oracle, g):
    d, s, lab = g["nms_dets"], g["nms_scores"], g["nms_labels"]
    order = np.argsort(-s.astype(np.float64), kind="stable").astype(np.int32)
    for thr in (0.1, 0.5):
        assert np.array_equal(oracle.nms_rotated_keep(d, order, thr, 5, False), g[f"nms_keep_gt_{thr}"])
        assert np.array_equal(oracle.nms_rotated_keep(d, order, thr, 5, True), g[f"nms_keep_ge_{thr}"])
        assert np.array_equal(oracle.nms_rotated(d, s, thr), np.nonzero(g[f"nms_keep_gt_{thr}"])[0])
    assert np.array_equal(oracle.ml_nms_rotated(d, s, lab, 0.1), np.nonzero(g["mlnms_keep_0.1"])[0])


def test_roi_align(oracle, g):
    feat, rois, grad = g["roi_feat"], g["roi_rois"], g["roi_grad"]
    for v in (0, 1):
        assert np.array_equal(oracle.roi_align_rotated_fwd(feat, rois, (7, 7), 1 / 16, 2, v), g[f"roi_fwd_v{v}"])
        assert np.array_equal(oracle.roi_align_rotated_bwd(grad, rois, feat.shape, 1 / 16, 2, v), g[f"roi_bwd_v{v}"])
    assert np.array_equal(oracle.roi_align_rotated_fwd(feat, rois, (7, 7), 1 / 16, 0, 1), g["roi_fwd_v1_adaptive"])


def test_poly(oracle, g):
    pp, ps = g["poly_polys"], g["poly_scores"]
    assert np.array_equal(oracle.poly_iou_matrix(pp[:40], pp[40:80]), g["poly_iou"])
    assert np.array_equal(oracle.poly_nms(np.concatenate([pp, ps[:, None]], 1), 0.1), g["poly_nms_keep_0.1"])


def test_merge_drift_guard(oracle, g):
    # UNPINNED (no Shapely): guards the restatement against drift only
    for thr in (0.1, 0.3):
        assert oracle.py_cpu_nms_poly_fast(g["merge_dets"], thr) == g[f"merge_keep_{thr}"].tolist()
