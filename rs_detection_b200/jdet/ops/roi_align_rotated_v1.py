"""jdet.ops.roi_align_rotated_v1 -- python/jdet/ops/roi_align_rotated_v1.py:300-373.

`ROIAlignRotated_v1(output_size, spatial_scale, sampling_ratio=0)` is looked up BY NAME from the
configs (`getattr(roi_align_rotated_v1, layer_type)`, oriented_single_level.py:43-51), so class name,
constructor arguments, the `output_size` tuple attribute and `__repr__` are kept verbatim.
"""
import torch
from torch import nn

from ... import core

__all__ = ["ROIAlign"]  # sic, roi_align_rotated_v1.py:5


def _pair(x):
    return tuple(x) if isinstance(x, (tuple, list)) else (x, x)


class _RotatedROIAlignFn(torch.autograd.Function):
    """Forward + backward of one feature map (the reference's jt.Function, :300-351); rois get no grad."""

    @staticmethod
    def forward(ctx, input, rois, output_size, spatial_scale, sampling_ratio, version):
        assert rois.shape[1] == 6
        cfg = core.make_roi_cfg([tuple(input.shape)], [spatial_scale], output_size, int(sampling_ratio), version)
        ctx.cfg, ctx.shape = cfg, tuple(input.shape)
        ctx.save_for_backward(rois)
        return core.roi_align_rotated_forward(cfg, [input], rois)

    @staticmethod
    def backward(ctx, grad_output):
        (rois,) = ctx.saved_tensors
        grads = core.roi_align_rotated_backward(ctx.cfg, grad_output.contiguous(), rois, [ctx.shape])
        return grads[0], None, None, None, None, None


def roi_align(input, rois, output_size, spatial_scale, sampling_ratio):
    return _RotatedROIAlignFn.apply(input, rois, _pair(output_size), spatial_scale, sampling_ratio, 1)


class ROIAlignRotated_v1(nn.Module):
    def __init__(self, output_size, spatial_scale, sampling_ratio=0):
        super(ROIAlignRotated_v1, self).__init__()
        self.output_size = _pair(output_size)
        self.spatial_scale = spatial_scale
        self.sampling_ratio = sampling_ratio

    def forward(self, input, rois):
        return roi_align(input, rois, self.output_size, self.spatial_scale, self.sampling_ratio)

    execute = forward  # Jittor spelling

    def __repr__(self):
        tmpstr = self.__class__.__name__ + "("
        tmpstr += "output_size=" + str(self.output_size)
        tmpstr += ", spatial_scale=" + str(self.spatial_scale)
        tmpstr += ", sampling_ratio=" + str(self.sampling_ratio)
        tmpstr += ")"
        return tmpstr
