#!/usr/bin/env python
"""single-class nms_rotated on n boxes (ncu target): python tools/profile_nms.py N THR [CANVAS]"""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import workloads as W
from rs_detection_b200 import core
from rs_detection_b200._lib import NMS_ROTATED
n, thr = int(sys.argv[1]), float(sys.argv[2])
canvas = int(sys.argv[3]) if len(sys.argv) > 3 else 1024
d = torch.from_numpy(W.rotated_boxes(n, n, canvas=canvas, smin=8, smax=128)).cuda()
s = torch.from_numpy(W.distinct_scores(n, n)).cuda()
for _ in range(2):
    print(core.nms(NMS_ROTATED, d, s, thr).count)
