# python/jdet/ops/__init__.py:1-2 exports exactly these two; everything else is imported as a submodule
from .box_iou_rotated import box_iou_rotated
from .box_iou_rotated_v1 import box_iou_rotated_v1
