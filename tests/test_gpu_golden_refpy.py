"""GPU: the product against tests/golden/refpy_golden.npz -- outputs of the reference's OWN Python
(`ops/bbox_transforms.py`, `models/boxes/assigner.py`, `models/roi_heads/oriented_head.py` + `models/boxes/coder.py`)
executed on the Jittor shim by tests/golden/make_golden_refpy.py.  Pins SURVEY 8 rows a6, a10 and f1 for the
device path: assignment indices / labels and the surviving (roi, class) pairs are exact, coordinates agree to float32
rounding of cos/sin/exp (CUDA vs torch CPU kernels), tolerances below."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "refpy_golden.npz")


@pytest.fixture(scope="module")
def g():
    return dict(np.load(GOLD))


def _t(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def _close(got, want, what, rtol=2e-6):
    scale = max(1.0, float(np.abs(want).max()))
    err = float(np.abs(got.astype(np.float64) - want.astype(np.float64)).max())
    print(f"{what}: max abs err {err:.3g} at scale {scale:.3g}")
    assert got.shape == want.shape and err <= rtol * scale, what


def test_transforms_vs_reference_python(cuda, g):
    from rs_detection_b200.jdet.ops.bbox_transforms import obb2hbb, obb2poly, poly2hbb
    _close(obb2poly(_t(g["tf_obb"])).cpu().numpy(), g["tf_obb2poly"], "obb2poly")
    _close(obb2hbb(_t(g["tf_obb"])).cpu().numpy(), g["tf_obb2hbb"], "obb2hbb")
    assert np.array_equal(poly2hbb(_t(g["tf_obb2poly"])).cpu().numpy(), g["tf_poly2hbb"])


@pytest.mark.parametrize("tag,kw", [("rcnn", dict(pos_iou_thr=0.5, neg_iou_thr=0.5, min_pos_iou=0.5, match_low_quality=False)),
                                    ("rpn", dict(pos_iou_thr=0.7, neg_iou_thr=0.3, min_pos_iou=0.3, match_low_quality=True)),
                                    ("rpn_one", dict(pos_iou_thr=0.7, neg_iou_thr=0.3, min_pos_iou=0.3, match_low_quality=True,
                                                     gt_max_assign_all=False))])
def test_assigner_vs_reference_python(cuda, g, tag, kw):
    """incl. argmax ties over GTs / over proposals, an all-zero column and the class default assigned_labels_filled=0"""
    from rs_detection_b200.jdet.models.boxes.assigner import MaxIoUAssigner
    res = MaxIoUAssigner(**kw).assign_wrt_overlaps(_t(g["as_overlaps"]), _t(g["as_gt_labels"]))
    assert np.array_equal(res.gt_inds.cpu().numpy(), g[f"as_{tag}_gt_inds"])
    assert np.array_equal(res.max_overlaps.cpu().numpy(), g[f"as_{tag}_max_overlaps"])
    assert np.array_equal(res.labels.cpu().numpy(), g[f"as_{tag}_labels"])


@pytest.mark.parametrize("tag,agnostic", [("agn", True), ("cls", False), ("raw", True)])
def test_head_tail_vs_reference_python(cuda, g, tag, agnostic):
    from rs_detection_b200.jdet.models.roi_heads.oriented_head import OrientedHeadTail
    scale = g[f"hd_{tag}_scale"]
    scale = None if scale.size == 1 and scale[0] == 0 else (float(scale[0]) if scale.size == 1 else scale.tolist())
    wd, wl = g[f"hd_{tag}_dets"], g[f"hd_{tag}_labels"]
    head = OrientedHeadTail(10, float(g[f"hd_{tag}_thr"]), reg_class_agnostic=agnostic)
    gd, gl = head.get_bboxes(_t(g["hd_rois"]), _t(g["hd_cls"]), _t(g[f"hd_{tag}_pred"]), (1024, 1024), scale, rescale=scale is not None)
    gd, gl = gd.cpu().numpy(), gl.cpu().numpy()
    assert gd.shape == wd.shape and np.array_equal(gl, wl)
    np.testing.assert_allclose(gd[:, 8], wd[:, 8], rtol=1e-5, atol=1e-7)
    fin = np.isfinite(wd).all(1)
    tol = 2e-5 * max(1.0, float(np.abs(wd[fin, :8]).max()))
    err = np.abs(gd[fin, :8] - wd[fin, :8]).max(1)
    bad = err > tol
    if bad.any():   # angle within an ulp of the +-pi/2 wrap: same rectangle, vertices rotated by two positions
        assert np.abs(np.roll(gd[fin][bad, :8], 4, axis=1) - wd[fin][bad, :8]).max() <= tol
    print(f"{tag}: {wd.shape[0]} detections, max coord err {float(err[~bad].max()):.3g} (tol {tol:.3g}), wrap-ambiguous {int(bad.sum())}")
    assert bad.sum() <= 3
