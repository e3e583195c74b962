"""jdet.ops.nms_rotated -- python/jdet/ops/nms_rotated.py:495-596.

Same names, argument order and return conventions as the reference:
  nms_rotated / ml_nms_rotated   -> kept indices, ascending ORIGINAL index (`jt.where(keep)[0]`)
  nms_rotated_cpu / _cuda        -> bool keep mask in ORIGINAL index space
  multiclass_nms_rotated         -> ((k,6) [cx,cy,w,h,theta,score] by descending score, (k,) 0-based labels)
Everything runs on the GPU; `nms_rotated_cpu` keeps the reference CPU body's `>=` decision rule,
`nms_rotated_cuda` the CUDA kernel's `>` (nms_rotated.py:444 vs :403).
"""
import torch

from ... import core
from ..._lib import NMS_ROTATED, NMS_ROTATED_GE
from ._io import back, dev


def _keep_from_order(dets, order_t, iou_threshold, box_length, kind):
    """The reference passes an explicit score order; rank positions stand in for the scores."""
    d, fl = dev(dets)
    o, _ = dev(order_t, torch.int64)
    n = d.shape[0]
    assert d.dim() == 2 and d.shape[1] == box_length
    pseudo = torch.empty((n,), dtype=torch.float32, device=d.device)
    pseudo[o] = torch.arange(n, 0, -1, dtype=torch.float32, device=d.device)
    labels = d[:, 5].to(torch.int32) if box_length == 6 else None
    res = core.nms(kind, d[:, :5], pseudo, float(iou_threshold), labels=labels, want_mask=True, want_sorted=False)
    return back(res.keep_mask, fl)


def nms_rotated_cpu(dets, order_t, iou_threshold, box_length=6):
    return _keep_from_order(dets, order_t, iou_threshold, box_length, NMS_ROTATED_GE)


def nms_rotated_cuda(dets, order_t, iou_threshold, box_length=6):
    return _keep_from_order(dets, order_t, iou_threshold, box_length, NMS_ROTATED)


def ml_nms_rotated(dets, scores, labels, iou_threshold):
    d, fl = dev(dets)
    assert d.numel() > 0 and d.dim() == 2
    assert dets.dtype == scores.dtype
    s, _ = dev(scores)
    l, _ = dev(labels, torch.int32)
    res = core.nms(NMS_ROTATED, d, s, float(iou_threshold), labels=l, want_mask=False, want_sorted=True)
    return back(res.sorted_idx, fl)


def nms_rotated(dets, scores, iou_threshold):
    d, fl = dev(dets)
    if d.numel() == 0:
        return back(torch.zeros((0,), dtype=torch.int64, device=d.device), fl)  # `jt.array([])`, :528-529
    assert d.dim() == 2
    assert dets.dtype == scores.dtype
    s, _ = dev(scores)
    res = core.nms(NMS_ROTATED, d, s, float(iou_threshold), want_mask=False, want_sorted=True)
    return back(res.sorted_idx, fl)


def multiclass_nms_rotated(multi_bboxes, multi_scores, score_thr, nms_cfg, max_num=-1, score_factors=None):
    """NMS for multi-class bboxes (nms_rotated.py:540-596); column 0 of multi_scores is background.
    One fused device pipeline; the only host sync is reading the output count."""
    mb, fl = dev(multi_bboxes)
    ms, _ = dev(multi_scores)
    sf = dev(score_factors)[0] if score_factors is not None else None
    nms_cfg_ = dict(nms_cfg)
    nms_cfg_.pop('type', 'nms')
    iou_thr = nms_cfg_.pop('iou_thr', 0.1)
    dets, labels, cnt = core.multiclass_nms_rotated(mb, ms, float(score_thr), float(iou_thr), int(max_num), sf)
    k = int(cnt.item())
    return back(dets[:k], fl), back(labels[:k], fl)
