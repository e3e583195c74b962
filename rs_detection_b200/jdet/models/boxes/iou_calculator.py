"""jdet.models.boxes.iou_calculator (rotated part) -- python/jdet/models/boxes/iou_calculator.py:81-162."""
from ...ops.box_iou_rotated import box_iou_rotated
from ...ops.box_iou_rotated_v1 import box_iou_rotated_v1


def bbox_overlaps_rotated(rboxes1, rboxes2, version=0):
    if version == 0:
        ious = box_iou_rotated(rboxes1.float(), rboxes2.float())
    else:
        ious = box_iou_rotated_v1(rboxes1.float(), rboxes2.float())
    return ious


class _Rotated2D:
    version = 0

    def __call__(self, bboxes1, bboxes2, mode='iou', is_aligned=False):
        assert bboxes1.size(-1) in [0, 5, 6]
        assert bboxes2.size(-1) in [0, 5, 6]
        if bboxes2.size(-1) == 6:
            bboxes2 = bboxes2[..., :5]
        if bboxes1.size(-1) == 6:
            bboxes1 = bboxes1[..., :5]
        assert mode == "iou" and is_aligned == False  # noqa: E712 (as in the reference)
        return bbox_overlaps_rotated(bboxes1, bboxes2, version=self.version)

    def __repr__(self):
        return self.__class__.__name__ + '()'


class BboxOverlaps2D_rotated(_Rotated2D):
    """2D Overlaps Calculator, counter-clockwise convention (iou_calculator.py:81-115)."""
    version = 0


class BboxOverlaps2D_rotated_v1(_Rotated2D):
    """2D Overlaps Calculator, Oriented R-CNN convention (iou_calculator.py:117-155)."""
    version = 1
