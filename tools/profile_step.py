#!/usr/bin/env python
"""One tile of the bench workload, repeated: the command ncu wraps (profiles/README.md has the exact lines).
    python tools/profile_step.py [--tiles 2] [--reps 3] [--what all|roi|nms|train]"""
import argparse
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench as B  # noqa: E402
import workloads as W  # noqa: E402
from rs_detection_b200 import core  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--tiles", type=int, default=2)
ap.add_argument("--reps", type=int, default=3)
ap.add_argument("--what", default="all")
ap.add_argument("--roi-batch", type=int, default=1, help="tiles per extractor call (bench.py's timed step uses 4)")
a = ap.parse_args()
dev = torch.device("cuda:0")
shapes = W.fpn_shapes()
cfg = core.make_roi_cfg(shapes, [1.0 / s for s in W.STRIDES], 7, 2, 1, B.EXTEND, 56.0)
tiles = []
for t in range(a.tiles):
    fs, r, b, s = B.tile_inputs(t)
    tiles.append(([torch.from_numpy(f).to(dev) for f in fs], torch.from_numpy(r).to(dev), torch.from_numpy(b).to(dev),
                  torch.from_numpy(s).to(dev)))
out = torch.empty((B.K_ROIS * a.roi_batch, W.CHANNELS, 7, 7), device=dev)
if a.roi_batch > 1:   # the timed step's extractor call: a batch of images, RoIs with their batch index
    assert a.tiles % a.roi_batch == 0
    cfgb = core.make_roi_cfg([(a.roi_batch,) + tuple(sh[1:]) for sh in shapes], [1.0 / s for s in W.STRIDES], 7, 2, 1, B.EXTEND, 56.0)
    groups = []
    for g in range(a.tiles // a.roi_batch):
        ts = tiles[g * a.roi_batch:(g + 1) * a.roi_batch]
        fb = [torch.cat([t[0][l] for t in ts], 0).contiguous() for l in range(len(shapes))]
        rb = torch.cat([torch.cat([torch.full((t[1].shape[0], 1), float(j), device=dev), t[1][:, 1:]], 1) for j, t in enumerate(ts)], 0).contiguous()
        groups.append((fb, rb))
gout = torch.randn((512, W.CHANNELS, 7, 7), device=dev)
for rep in range(a.reps):
    if a.roi_batch > 1 and a.what in ("all", "roi"):
        for fb, rb in groups:
            core.roi_align_rotated_forward(cfgb, fb, rb, out=out)
    for feats, rois, boxes, scores in tiles:
        if a.roi_batch == 1 and a.what in ("all", "roi"):
            core.roi_align_rotated_forward(cfg, feats, rois, out=out)
        if a.what in ("all",):
            core.obb2poly(boxes)
        if a.what in ("all", "nms"):
            core.multiclass_nms_rotated(boxes, scores, B.SCORE_THR, B.IOU_THR, B.MAX_NUM)
        if a.what in ("train",):
            core.roi_align_rotated_forward(cfg, feats, rois[:512].contiguous())
            core.roi_align_rotated_backward(cfg, gout, rois[:512].contiguous(), shapes)
    torch.cuda.synchronize()
print("done")
