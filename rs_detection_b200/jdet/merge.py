"""Top-level `merge.py` of the reference (two-model CSV ensemble), the NMS part: merge.py:14-27."""
import numpy as np
import torch

from .. import core
from .._lib import NMS_HBB, require_cuda


def nms(boxes, thresh):
    """boxes (n,5) float64 [x1,y1,x2,y2,score]; keeps `iou < thresh`; returns np.array of kept indices in
    descending-score order (merge.py:14-27)."""
    require_cuda()
    b = np.ascontiguousarray(boxes, dtype=np.float64).reshape(-1, 5)
    if b.shape[0] == 0:
        return np.array([], dtype=np.int64)
    t = torch.from_numpy(b).cuda()
    res = core.nms(NMS_HBB, t[:, :4], t[:, 4], float(thresh), want_mask=False, want_sorted=False, want_score=True,
                   ws_tag="merge")
    return res.score_idx.cpu().numpy()
