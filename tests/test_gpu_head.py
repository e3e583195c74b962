"""fused OrientedHead.get_bboxes/get_results vs the numpy restatement (oracle)."""
import numpy as np
import pytest
import torch

import workloads as W

pytestmark = pytest.mark.gpu


def _t(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


@pytest.mark.parametrize("agnostic", [True, False])
@pytest.mark.parametrize("C,thr,scale", [(10, 0.001, None), (15, 0.05, 1.5), (37, 0.05, (0.5, 0.5, 0.5, 0.5))])
def test_head_tail_vs_oracle(cuda, oracle, agnostic, C, thr, scale):
    from rs_detection_b200.jdet.models.roi_heads.oriented_head import OrientedHeadTail
    K = 3000
    for seed in range(C, C + 20):  # pick a draw with no softmax score inside the 1e-6 band around the threshold
        rng = np.random.default_rng(seed)
        cls = (rng.standard_normal((K, C + 1)) * 2.0).astype(np.float32)
        e0 = np.exp(cls - cls.max(1, keepdims=True))
        if int((np.abs((e0 / e0.sum(1, keepdims=True))[:, :-1] - thr) < 1e-5 * thr).sum()) == 0:
            break
    rois = W.proposals(K, C)
    pred = (rng.standard_normal((K, 5 if agnostic else 5 * C)) * 0.8).astype(np.float32)
    pred[:7, 2::5] = 40.0  # exercises the wh_ratio_clip clamp
    head = OrientedHeadTail(C, thr, reg_class_agnostic=agnostic)
    gd, gl = head.get_bboxes(_t(rois), _t(cls), _t(pred), scale_factor=scale, rescale=scale is not None)
    wd, wl = oracle.oriented_head_get_bboxes(rois, cls, pred, scale, score_thresh=thr)
    # candidates whose softmax score sits within 1e-6 of the threshold may flip with exp() ulps
    e = np.exp(cls - cls.max(1, keepdims=True)); sc = e / e.sum(1, keepdims=True)
    near = int((np.abs(sc[:, :-1] - thr) < 1e-5 * thr).sum())  # expf/exp differ by ulps: band relative to thr
    print(f"C={C} agnostic={agnostic}: {wd.shape[0]} detections, {near} scores within 1e-5*thr of the threshold")
    assert near == 0
    assert gd.shape == wd.shape and wd.shape[0] > 1000
    assert np.array_equal(gl.cpu().numpy(), wl)
    g = gd.cpu().numpy()
    # polygons: coordinates up to ~2e3 -> 1e-5 relative to the coordinate scale (cosf/sinf/expf ulps)
    tol = 1e-5 * max(1.0, float(np.abs(wd[:, :8]).max()))
    err = np.abs(g[:, :8] - wd[:, :8])
    # an angle that lands within an ulp of the +-pi/2 wrap may wrap differently: the polygon is the same
    # rectangle with its vertices rotated by two positions; count those rows, compare the rest
    bad = err.max(1) > 20 * tol
    if bad.any():
        rolled = np.roll(g[bad, :8], 4, axis=1)
        assert np.abs(rolled - wd[bad, :8]).max() <= 20 * tol
        print("wrap-ambiguous rows:", int(bad.sum()))
    assert bad.sum() <= 3
    np.testing.assert_allclose(g[:, 8], wd[:, 8], rtol=1e-5, atol=1e-7)
    assert gl.dtype == torch.int64


def test_head_tail_empty(cuda):
    from rs_detection_b200.jdet.models.roi_heads.oriented_head import OrientedHeadTail
    head = OrientedHeadTail(10, 0.999999)
    gd, gl = head.get_bboxes(_t(W.proposals(50, 1)), torch.zeros((50, 11), device="cuda"), torch.zeros((50, 5), device="cuda"))
    assert tuple(gd.shape) == (0, 9) and gl.numel() == 0
    gd, gl = head.get_bboxes(torch.zeros((0, 6), device="cuda"), torch.zeros((0, 11), device="cuda"), torch.zeros((0, 5), device="cuda"))
    assert tuple(gd.shape) == (0, 9)
