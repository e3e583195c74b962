"""CPU: an independent implementation as a SECOND OPINION on the oracle's geometry (not a pin: OpenCV works in
float32).  The merge-stage polygon IoU is "parity unpinned" because the reference delegates it to Shapely/GEOS
(absent); here it is cross-checked against cv2.intersectConvexConvex, and so is the rotated IoU restatement."""
import numpy as np
import pytest

import workloads as W

cv2 = pytest.importorskip("cv2")


def _cv_iou(p, q, floor):
    p, q = p.reshape(4, 2), q.reshape(4, 2)
    c = p.mean(0)
    p32, q32 = (p - c).astype(np.float32), (q - c).astype(np.float32)
    inter, _ = cv2.intersectConvexConvex(p32, q32)
    a1, a2 = cv2.contourArea(p32), cv2.contourArea(q32)
    return inter / max(a1 + a2 - inter, floor)


def test_merge_iou_poly_vs_opencv(oracle):
    sc = W.merge_scene(num_objects=400, scene=2500, seed=5)
    P = sc["polys"]
    hb = np.concatenate([P[:, 0::2].min(1)[:, None], P[:, 1::2].min(1)[:, None], P[:, 0::2].max(1)[:, None],
                         P[:, 1::2].max(1)[:, None]], 1)
    ii, jj = np.nonzero((hb[:, None, 0] < hb[None, :, 2]) & (hb[None, :, 0] < hb[:, None, 2]) &
                        (hb[:, None, 1] < hb[None, :, 3]) & (hb[None, :, 1] < hb[:, None, 3]))
    sel = np.nonzero(ii < jj)[0][:1500]
    assert sel.size >= 500
    err, pos = 0.0, 0
    for k in sel:
        a, b = int(ii[k]), int(jj[k])
        want = oracle.iou_poly(P[a], P[b])
        err = max(err, abs(_cv_iou(P[a], P[b], 0.01) - want))
        pos += want > 0.05
    assert pos >= 300 and err < 5e-4, (pos, err)


@pytest.mark.parametrize("version", [0, 1])
def test_rotated_iou_vs_opencv(oracle, version):
    boxes = W.rotated_boxes(250, 9, canvas=600, smin=16, smax=160)
    gts = W.jittered_copies(boxes, 60, 4)
    fn = oracle.box_iou_rotated_v1 if version else oracle.box_iou_rotated
    iou = fn(gts, boxes)
    # obb2poly draws the clockwise (v1) convention; v0 is the same box with the angle negated
    sgn = np.array([1, 1, 1, 1, 1 if version else -1], np.float32)
    pg, pb = oracle.obb2poly(gts * sgn).astype(np.float64), oracle.obb2poly(boxes * sgn).astype(np.float64)
    err, pos = 0.0, 0
    for a, b in zip(*np.nonzero(iou > 0.01)):
        err = max(err, abs(_cv_iou(pg[a], pb[b], 1e-12) - float(iou[a, b])))
        pos += 1
    assert pos >= 100 and err < 1e-3, (pos, err)
