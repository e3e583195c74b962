"""The oracle's numpy restatements against tests/golden/refpy_golden.npz -- outputs of the reference's OWN Python
(`ops/bbox_transforms.py`, `models/boxes/coder.py`, `models/boxes/assigner.py`, `models/roi_heads/oriented_head.py`),
executed by tests/golden/make_golden_refpy.py on the torch-backed Jittor shim.  Integer results (assignment,
labels, which detections survive) must be identical; float results agree to float32 rounding of cos/sin/exp
(numpy vs torch kernels; tolerances written below)."""
import os

import numpy as np
import pytest

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "refpy_golden.npz")
COORD_RTOL = 2e-6      # relative to the coordinate scale (~1e3 px): a few float32 ulps


@pytest.fixture(scope="module")
def g():
    return dict(np.load(GOLD))


def _close(got, want, what, rtol=COORD_RTOL):
    scale = max(1.0, float(np.abs(want).max()))
    err = float(np.abs(got.astype(np.float64) - want.astype(np.float64)).max())
    print(f"{what}: max abs err {err:.3g} at scale {scale:.3g}")
    assert got.shape == want.shape and err <= rtol * scale, what


def test_transforms(oracle, g):
    _close(oracle.obb2poly(g["tf_obb"]), g["tf_obb2poly"], "obb2poly")
    _close(oracle.obb2hbb(g["tf_obb"]), g["tf_obb2hbb"], "obb2hbb")
    assert np.array_equal(oracle.poly2hbb(g["tf_obb2poly"]), g["tf_poly2hbb"])


def test_regular_theta_obb(oracle, g):
    _close(oracle.regular_theta(g["rt_theta"]), g["rt_180"], "regular_theta 180", 1e-6)
    _close(oracle.regular_theta(g["rt_theta"], mode='360', start=-np.pi), g["rt_360"], "regular_theta 360", 1e-6)
    got, want = oracle.regular_obb(g["ro_in"]), g["ro_out"]
    assert np.array_equal(got[:, :4], want[:, :4])
    _close(got[:, 4], want[:, 4], "regular_obb theta", 1e-6)


def _angle_close(got, want, what, tol=2e-5):
    """angles are compared modulo the pi wrap of regular_theta (a value within an ulp of +-pi/2 may wrap either way)"""
    d = np.abs(got - want)
    d = np.minimum(d, np.abs(d - np.float32(np.pi)))
    print(f"{what}: max angle err {float(d.max()):.3g}")
    assert float(d.max()) <= tol, what


def test_rectpoly2obb_and_midpoint_decode(oracle, g):
    got, want = oracle.rectpoly2obb(g["rp_in"]), g["rp_out"]
    _close(got[:, :4], want[:, :4], "rectpoly2obb xywh", 1e-5)
    _angle_close(got[:, 4], want[:, 4], "rectpoly2obb theta", 1e-4)
    got, want = oracle.midpoint_offset_decode(g["mo_anchors"], g["mo_pred"]), g["mo_decode"]
    assert got.shape == want.shape
    ok = np.isfinite(want).all(1)
    assert ok.sum() >= len(want) - 10 and np.array_equal(np.isfinite(got).all(1), ok)
    # w/h of near-square boxes can swap with a pi/2 turn when w ~ h within rounding: compare area + centre there
    same = np.abs(got[ok, 2] - want[ok, 2]) <= 1e-3 * np.maximum(1.0, np.abs(want[ok, 2]))
    assert same.mean() > 0.99
    _close(got[ok][same][:, :4], want[ok][same][:, :4], "midpoint decode xywh", 2e-5)
    _angle_close(got[ok][same][:, 4], want[ok][same][:, 4], "midpoint decode theta", 2e-4)


@pytest.mark.parametrize("tag", ["agn", "cls"])
def test_delta_xywht_decode(oracle, g, tag):
    got = oracle.delta_xywht_decode(g["od_rois"], g[f"od_{tag}_pred"], (0., 0., 0., 0., 0.), (0.1, 0.1, 0.2, 0.2, 0.1))
    want = g[f"od_{tag}_decode"]
    gg, ww = got.reshape(-1, 5), want.reshape(-1, 5)
    fin = np.isfinite(ww).all(1)
    assert np.array_equal(np.isfinite(gg).all(1), fin)
    big = np.abs(ww[fin, :4]).max() if fin.any() else 1.0
    _close(gg[fin][:, :4], ww[fin][:, :4], f"decode {tag} xywh", 5e-6 if big < 1e6 else 1e-5)
    _angle_close(gg[fin][:, 4], ww[fin][:, 4], f"decode {tag} theta")


@pytest.mark.parametrize("tag,kw", [("rcnn", dict(pos_iou_thr=0.5, neg_iou_thr=0.5, min_pos_iou=0.5, match_low_quality=False)),
                                    ("rpn", dict(pos_iou_thr=0.7, neg_iou_thr=0.3, min_pos_iou=0.3, match_low_quality=True)),
                                    ("rpn_one", dict(pos_iou_thr=0.7, neg_iou_thr=0.3, min_pos_iou=0.3, match_low_quality=True,
                                                     gt_max_assign_all=False))])
def test_assigner(oracle, g, tag, kw):
    """MaxIoUAssigner.assign_wrt_overlaps run by the reference itself, incl. argmax ties over GTs / over proposals."""
    gi, mo, lab = oracle.max_iou_assign(g["as_overlaps"], gt_labels=g["as_gt_labels"], **kw)
    assert np.array_equal(gi, g[f"as_{tag}_gt_inds"])
    assert np.array_equal(mo, g[f"as_{tag}_max_overlaps"])
    assert np.array_equal(lab, g[f"as_{tag}_labels"])


@pytest.mark.parametrize("tag", ["agn", "cls", "raw"])
def test_head_get_bboxes(oracle, g, tag):
    scale = g[f"hd_{tag}_scale"]
    scale = None if scale.size == 1 and scale[0] == 0 else (float(scale[0]) if scale.size == 1 else scale.tolist())
    wd, wl = g[f"hd_{tag}_dets"], g[f"hd_{tag}_labels"]
    gd, gl = oracle.oriented_head_get_bboxes(g["hd_rois"], g["hd_cls"], g[f"hd_{tag}_pred"], scale,
                                             score_thresh=float(g[f"hd_{tag}_thr"]))
    assert gd.shape == wd.shape and np.array_equal(gl, wl)        # the same (roi, class) pairs survive, same order
    np.testing.assert_allclose(gd[:, 8], wd[:, 8], rtol=2e-6, atol=1e-8)
    fin = np.isfinite(wd).all(1)
    err = np.abs(gd[fin, :8] - wd[fin, :8]).max(1)
    tol = 2e-5 * max(1.0, float(np.abs(wd[fin, :8]).max()))
    bad = err > tol
    if bad.any():   # an angle within an ulp of the +-pi/2 wrap: same rectangle, vertices rotated by two positions
        rolled = np.roll(gd[fin][bad, :8], 4, axis=1)
        assert np.abs(rolled - wd[fin][bad, :8]).max() <= tol
    assert bad.sum() <= 3
