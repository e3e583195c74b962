"""jdet.ops.box_iou_rotated -- python/jdet/ops/box_iou_rotated.py:502-509."""
from ... import core
from ._io import back, dev


def box_iou_rotated(boxes1, boxes2):
    """IoU of rotated boxes [cx,cy,w,h,theta] (counter-clockwise-positive convention of the S2ANet /
    RoI-Transformer family).  boxes1 (N,5), boxes2 (M,5) -> (N,M) float32."""
    assert boxes1.dtype == boxes2.dtype
    b1, fl = dev(boxes1)
    b2, _ = dev(boxes2)
    return back(core.box_iou_rotated(b1.reshape(-1, 5), b2.reshape(-1, 5), version=0), fl)
