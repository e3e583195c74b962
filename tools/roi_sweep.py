#!/usr/bin/env python
"""Time the RoIAlignRotated forward paths alone (channels-last pyramid resident, 8 distinct bench tiles back to
back = the measurement behind bench.py's `roofline`).  With a library built with RSDET_TUNING=1 the paths are
switched through RSDET_ROI_PATH (0 pixel-major, 1 bin-major registers, 2 TMA gather4); a shipped build always
runs its default and the sweep degenerates to one line.

    python tools/roi_sweep.py [--paths 0,1,2] [--check]
"""
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
ap.add_argument("--paths", default="0,1")
ap.add_argument("--check", action="store_true", help="compare every path's output with path 1 (bin-major)")
ap.add_argument("--reps", type=int, default=10)
a = ap.parse_args()

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


ref = None
for spec in a.paths.split(","):      # "<path>" or "<path>:<pixels per batch>" (row-window kernel: 4 or 8)
    path, _, pxb = spec.partition(":")
    path = int(path)
    os.environ["RSDET_ROI_PATH"] = str(path)
    if pxb:
        os.environ["RSDET_ROI_PXB"] = pxb
    for _ in range(3):
        run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.reps):
        run()
    e1.record()
    torch.cuda.synchronize()
    msg = f"path {spec}: {e0.elapsed_time(e1) / (8 * a.reps) * 1000:.1f} us per 4000-RoI tile (geometry + order + gather kernels)"
    if a.check:
        o = core.roi_align_rotated_forward(cfg, tiles[0][0], tiles[0][1]).clone()
        if ref is None:
            os.environ["RSDET_ROI_PATH"] = "1"
            ref = core.roi_align_rotated_forward(cfg, tiles[0][0], tiles[0][1]).clone()
        msg += f"; max |diff| vs bin-major {float((o - ref).abs().max()):.3g} (scale {float(ref.abs().max()):.3g})"
    print(msg, flush=True)
