#!/usr/bin/env python
"""One full-size RPN proposal-stage call (5 levels of a 1024^2 tile, 261 888 anchors) for ncu launch lists:
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_rpn.csv python tools/profile_rpn.py"""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import workloads as W
from rs_detection_b200 import core
from rs_detection_b200.jdet.models.boxes.anchor_generator import AnchorGenerator
shapes = [(256, 256), (128, 128), (64, 64), (32, 32), (16, 16)]
cls, reg = W.rpn_outputs(shapes, 3, 5)
cls = [torch.from_numpy(x).cuda() for x in cls]
reg = [torch.from_numpy(x).cuda() for x in reg]
anc = AnchorGenerator(strides=[4, 8, 16, 32, 64], ratios=[0.5, 1.0, 2.0], scales=[8]).grid_anchors(shapes)
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 2
for _ in range(reps):
    dets, cnt = core.rpn_proposals(cls, reg, anc, 3, True, 4000, 4000, 0.8, 0)
torch.cuda.synchronize()
print("proposals:", int(cnt.item()))
