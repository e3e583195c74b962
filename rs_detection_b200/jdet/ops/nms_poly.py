"""jdet.ops.nms_poly -- python/jdet/ops/nms_poly.py:187-252."""
import numpy as np
import torch

from ... import core
from ..._lib import NMS_MERGE, NMS_POLY
from ._io import back, dev


def poly_nms(boxes, nms_overlap_thresh):
    """boxes (n,9) [x1..y4,score] -> kept indices in DESCENDING-SCORE order (`order_t[keep]`, :232)."""
    b, fl = dev(boxes)
    assert b.dim() == 2 and b.shape[1] == 9
    res = core.nms(NMS_POLY, b[:, :8], b[:, 8], float(nms_overlap_thresh), want_mask=False, want_sorted=False,
                   want_score=True)
    return back(res.score_idx, fl)


def multiclass_poly_nms(bboxes, scores, labels, thresh):
    """nms_poly.py:234-245: classes are separated by translating every coordinate by label*(range+1)."""
    b, fl = dev(bboxes)
    s, _ = dev(scores)
    l, _ = dev(labels, torch.int64)
    max_coordinate = b.max() - b.min()
    offsets = l.to(b.dtype) * (max_coordinate + 1)
    bboxes_for_nms = b + offsets[:, None]
    keep = poly_nms(torch.cat([bboxes_for_nms, s[:, None]], dim=1), thresh)
    dets = torch.cat([b[keep], s[keep][:, None]], dim=1)
    return back(dets, fl), back(l[keep], fl)


def iou_poly(poly1, poly2):
    """nms_poly.py:247-252 (Shapely in the reference): float64 polygon IoU of two quads, evaluated on the
    device by the merge predicate's clipper.  Returns a python float."""
    p = torch.as_tensor(np.asarray(poly1, np.float64).reshape(1, 8), device="cuda")
    q = torch.as_tensor(np.asarray(poly2, np.float64).reshape(1, 8), device="cuda")
    return float(core.iou_poly_pairs(p, q)[0].item())
