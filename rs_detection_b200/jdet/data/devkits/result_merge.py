"""jdet.data.devkits.result_merge (merge NMS part) -- python/jdet/data/devkits/result_merge.py.

`py_cpu_nms_poly_fast(dets, thresh)` keeps its name, numpy-in / list-out contract (:66-127) so that
`mergebypoly`, `tools/merge_results.py:38` and `nmsbynamedict` (:177-193) can call it unchanged; the
greedy loop and the Shapely call per candidate pair are replaced by the float64 merge predicate of the
device NMS engine.  `merge_detections` is the batched form: ALL (scene, class) groups of a result set
in one launch, per-class thresholds included.
"""
import numpy as np
import torch

from .... import core
from ...._lib import NMS_MERGE, require_cuda

# the thresh for nms when merge image (result_merge.py:24-27)
nms_threshold_0 = 0.1
nms_threshold_1 = {'Roundabout': 0.1, 'Tennis_Court': 0.1, 'Football_Field': 0.1, 'Vehicle': 0.15, 'Ship': 0.2,
                   'Airplane': 0.3, 'Intersection': 0.3, 'Bridge': 0.0001, 'Basketball_Court': 0.1, 'Baseball_Field': 0.1}


def py_cpu_nms_poly_fast(dets, thresh):
    """dets (n,9) float64 [x1..y4,score] -> python list of kept indices in descending-score order."""
    require_cuda()
    d = np.ascontiguousarray(dets, dtype=np.float64).reshape(-1, 9)
    if d.shape[0] == 0:
        return []
    t = torch.from_numpy(d).cuda()
    res = core.nms(NMS_MERGE, t[:, :8], t[:, 8], float(thresh), want_mask=False, want_sorted=False, want_score=True,
                   ws_tag="merge")
    return res.score_idx.cpu().tolist()


def poly2origpoly(poly, x, y, rate):
    """:196-203"""
    origpoly = []
    for i in range(int(len(poly) / 2)):
        origpoly.append(float(poly[i * 2] + x) / float(rate))
        origpoly.append(float(poly[i * 2 + 1] + y) / float(rate))
    return origpoly


def nmsbynamedict(nameboxdict, nms, thresh):
    """:177-193 (kept for API parity; one engine call per scene)."""
    nameboxnmsdict = {x: [] for x in nameboxdict}
    for imgname in nameboxdict:
        keep = nms(np.array(nameboxdict[imgname]), thresh)
        nameboxnmsdict[imgname] = [nameboxdict[imgname][index] for index in keep]
    return nameboxnmsdict


def merge_detections(polys, scores, group_ids, thresh=nms_threshold_0, group_thresh=None):
    """Batched merge NMS: polys (n,8) float64 scene coordinates, scores (n,), group_ids (n,) int -- one id
    per (scene, class) pair; `group_thresh`: optional per-group-id thresholds (len > max id).
    Returns kept indices in descending-score order (device int64 tensor when inputs are CUDA tensors,
    numpy otherwise)."""
    require_cuda()
    host = not (isinstance(polys, torch.Tensor) and polys.is_cuda)
    p = torch.as_tensor(polys, dtype=torch.float64).cuda() if host else polys.to(torch.float64)
    s = torch.as_tensor(scores, dtype=torch.float64).cuda() if host else scores.to(torch.float64)
    g = torch.as_tensor(group_ids).to(torch.int32).cuda() if host else group_ids.to(torch.int32)
    tpl = None
    if group_thresh is not None:
        tpl = torch.as_tensor(group_thresh, dtype=torch.float64).cuda()
    res = core.nms(NMS_MERGE, p, s, float(thresh), labels=g, thr_per_label=tpl, want_mask=False, want_sorted=False,
                   want_score=True, ws_tag="merge")
    keep = res.score_idx
    return keep.cpu().numpy() if host else keep
