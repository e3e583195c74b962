"""The CUDA path against the committed golden vectors (same fixtures as tests/test_golden.py)."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "rotated_path_golden.npz")


@pytest.fixture(scope="module")
def g():
    return dict(np.load(GOLD))


def _t(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def test_iou_and_nms(cuda, g):
    from rs_detection_b200.jdet.ops import box_iou_rotated, box_iou_rotated_v1
    from rs_detection_b200.jdet.ops.nms_rotated import ml_nms_rotated, nms_rotated, nms_rotated_cpu
    assert np.array_equal(box_iou_rotated(_t(g["iou_boxes1"]), _t(g["iou_boxes2"])).cpu().numpy(), g["iou_v0"])
    assert np.array_equal(box_iou_rotated_v1(_t(g["iou_boxes1"]), _t(g["iou_boxes2"])).cpu().numpy(), g["iou_v1"])
    d, s, lab = g["nms_dets"], g["nms_scores"], g["nms_labels"]
    order = np.argsort(-s.astype(np.float64), kind="stable")
    for thr in (0.1, 0.5):
        assert np.array_equal(nms_rotated(_t(d), _t(s), thr).cpu().numpy(), np.nonzero(g[f"nms_keep_gt_{thr}"])[0])
        assert np.array_equal(nms_rotated_cpu(_t(d), _t(order), thr, box_length=5).cpu().numpy(), g[f"nms_keep_ge_{thr}"])
    assert np.array_equal(ml_nms_rotated(_t(d), _t(s), _t(lab), 0.1).cpu().numpy(), np.nonzero(g["mlnms_keep_0.1"])[0])


def test_roi_align(cuda, g):
    from rs_detection_b200.jdet.ops.roi_align_rotated import ROIAlignRotated
    from rs_detection_b200.jdet.ops.roi_align_rotated_v1 import ROIAlignRotated_v1
    for v, cls in ((0, ROIAlignRotated), (1, ROIAlignRotated_v1)):
        x = _t(g["roi_feat"]).requires_grad_(True)
        y = cls(7, 1 / 16, 2)(x, _t(g["roi_rois"]))
        np.testing.assert_allclose(y.detach().cpu().numpy(), g[f"roi_fwd_v{v}"], rtol=1e-5, atol=2e-5)
        y.backward(_t(g["roi_grad"]))
        np.testing.assert_allclose(x.grad.cpu().numpy(), g[f"roi_bwd_v{v}"], rtol=1e-4, atol=2e-4)
    y = ROIAlignRotated_v1(7, 1 / 16, 0)(_t(g["roi_feat"]), _t(g["roi_rois"]))
    np.testing.assert_allclose(y.cpu().numpy(), g["roi_fwd_v1_adaptive"], rtol=1e-5, atol=2e-5)


def test_poly_and_merge(cuda, g):
    from rs_detection_b200.jdet.data.devkits.result_merge import py_cpu_nms_poly_fast
    from rs_detection_b200.jdet.ops.nms_poly import poly_nms
    pp, ps = g["poly_polys"], g["poly_scores"]
    assert np.array_equal(poly_nms(_t(np.concatenate([pp, ps[:, None]], 1)), 0.1).cpu().numpy(), g["poly_nms_keep_0.1"])
    for thr in (0.1, 0.3):
        assert py_cpu_nms_poly_fast(g["merge_dets"], thr) == g[f"merge_keep_{thr}"].tolist()
