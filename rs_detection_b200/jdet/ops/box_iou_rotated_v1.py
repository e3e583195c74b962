"""jdet.ops.box_iou_rotated_v1 -- python/jdet/ops/box_iou_rotated_v1.py:507-524."""
from ... import core
from ._io import back, dev


def box_iou_rotated_v1(boxes1, boxes2):
    """Clockwise-positive angle convention (Oriented R-CNN).  Boxes with a side < 1e-3 get IoU 0 with
    everything (the reference's "bbox size too small" guard, :515-522, applied per box in-kernel)."""
    assert boxes1.dtype == boxes2.dtype
    b1, fl = dev(boxes1)
    b2, _ = dev(boxes2)
    return back(core.box_iou_rotated(b1.reshape(-1, 5), b2.reshape(-1, 5), version=1, zero_tiny=True), fl)
