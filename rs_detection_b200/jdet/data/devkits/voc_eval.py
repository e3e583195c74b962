"""jdet.data.devkits.voc_eval -- python/jdet/data/devkits/voc_eval.py.

`voc_eval_dota(dets, gts, iou_func, ovthresh, use_07_metric)` keeps the reference's signature (:236) and
numpy-in / (rec, prec, ap)-out contract; the double Python loop with one Shapely call per (detection, gt)
pair is replaced by one device pass (`rsdet_voc_match`).  `voc_ap` is the reference's (:50-82), host numpy.
`iou_func` is accepted for signature parity and ignored: the polygon IoU is the engine's float64 `iou_poly`.
"""
import numpy as np
import torch

from .... import core
from ...._lib import require_cuda


def voc_ap(rec, prec, use_07_metric=False):
    """:50-82"""
    if use_07_metric:
        ap = 0.
        for t in np.arange(0., 1.1, 0.1):
            p = 0 if np.sum(rec >= t) == 0 else np.max(prec[rec >= t])
            ap = ap + p / 11.
    else:
        mrec = np.concatenate(([0.], rec, [1.]))
        mpre = np.concatenate(([0.], prec, [0.]))
        for i in range(mpre.size - 1, 0, -1):
            mpre[i - 1] = np.maximum(mpre[i - 1], mpre[i])
        i = np.where(mrec[1:] != mrec[:-1])[0]
        ap = np.sum((mrec[i + 1] - mrec[i]) * mpre[i + 1])
    return ap


def voc_match(dets, gts, ovthresh=0.5):
    """dets (nd, 10) [img_id, 8 coords, confidence]; gts {img_id: {"box": (G,8), "difficult": (G,) bool}} ->
    (tp, fp) float arrays in descending-confidence order, npos."""
    require_cuda()
    dets = np.array(np.asarray(dets).tolist(), dtype=np.float64).reshape(-1, 10)
    ids = sorted(gts.keys())
    remap = {k: i for i, k in enumerate(ids)}
    boxes = [np.asarray(gts[k]["box"], np.float64).reshape(-1, 8) for k in ids]
    diff = [np.asarray(gts[k]["difficult"]).astype(bool).reshape(-1) for k in ids]
    npos = int(sum((~d).sum() for d in diff))
    start = np.zeros(len(ids) + 1, np.int32)
    start[1:] = np.cumsum([b.shape[0] for b in boxes])
    gt_all = np.concatenate(boxes) if boxes else np.zeros((0, 8))
    diff_all = np.concatenate(diff) if diff else np.zeros((0,), bool)
    order = np.argsort(-dets[:, -1])
    d = dets[order]
    img = np.array([remap.get(int(v), -1) for v in d[:, 0]], np.int32)
    tp, fp, _, _ = core.voc_match(torch.from_numpy(np.ascontiguousarray(d[:, 1:9])).cuda(), torch.from_numpy(img).cuda(),
                                  torch.from_numpy(np.ascontiguousarray(gt_all)).cuda(), torch.from_numpy(start).cuda(),
                                  torch.from_numpy(diff_all.astype(np.uint8)).cuda(), ovthresh)
    return tp.cpu().numpy().astype(np.float64), fp.cpu().numpy().astype(np.float64), npos


def voc_eval_dota(dets, gts, iou_func=None, ovthresh=0.5, use_07_metric=False):
    dets = np.array(np.asarray(dets).tolist())
    npos = sum([sum(~np.asarray(gts[k]["difficult"]).astype(bool)) for k in gts])
    if len(dets) == 0 or npos == 0:
        return 0., 0., 0.
    tp, fp, npos = voc_match(dets, gts, ovthresh)
    fp = np.cumsum(fp)
    tp = np.cumsum(tp)
    rec = tp / float(npos)
    prec = tp / np.maximum(tp + fp, np.finfo(np.float64).eps)
    ap = voc_ap(rec, prec, use_07_metric)
    return rec, prec, ap
