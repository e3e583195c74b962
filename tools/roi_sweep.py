#!/usr/bin/env python
"""Time the RoIAlignRotated forward kernel alone (channels-last pyramid resident) on bench tiles."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench as B, workloads as W
from rs_detection_b200 import core
dev = torch.device("cuda:0")
shapes = W.fpn_shapes()
cfg = core.make_roi_cfg(shapes, [1.0 / s for s in W.STRIDES], 7, 2, 1, B.EXTEND, 56.0, channels_last=True)
tiles = []
for t in range(8):
    fs, r, b, s = B.tile_inputs(t)
    tiles.append(([core.nchw_to_nhwc(torch.from_numpy(f).to(dev)) for f in fs], torch.from_numpy(r).to(dev)))
out = torch.empty((B.K_ROIS, W.CHANNELS, 7, 7), device=dev)
def run():
    for f, r in tiles:
        core.roi_align_rotated_forward(cfg, f, r, out=out)
for _ in range(3): run()
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(10): run()
b.record(); torch.cuda.synchronize()
print(f"carveout={os.environ.get('RSDET_ROI_CARVEOUT','default')}: {a.elapsed_time(b)/80*1000:.1f} us per tile (order kernel + fwd kernel)")
