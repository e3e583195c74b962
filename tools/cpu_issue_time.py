#!/usr/bin/env python
"""How long does the host need to ISSUE one 8-tile step (launch-bound or not)?"""
import os, sys, time, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench as B, workloads as W
from rs_detection_b200 import core
dev = torch.device("cuda:0")
shapes = W.fpn_shapes()
cfg = core.make_roi_cfg(shapes, [1.0 / s for s in W.STRIDES], 7, 2, 1, B.EXTEND, 56.0)
tiles = []
for t in range(8):
    fs, r, b, s = B.tile_inputs(t)
    tiles.append(([torch.from_numpy(f).to(dev) for f in fs], torch.from_numpy(r).to(dev), torch.from_numpy(b).to(dev), torch.from_numpy(s).to(dev)))
side = [torch.cuda.Stream() for _ in range(4)]
outs = [torch.empty((B.K_ROIS, 256, 7, 7), device=dev) for _ in range(4)]
def step():
    main = torch.cuda.current_stream()
    for st in side: st.wait_stream(main)
    for i, (f, r, b, s) in enumerate(tiles):
        with torch.cuda.stream(side[i % 4]):
            core.roi_align_rotated_forward(cfg, f, r, out=outs[i % 4])
            core.obb2poly(b)
            core.multiclass_nms_rotated(b, s, B.SCORE_THR, B.IOU_THR, B.MAX_NUM)
    for st in side: main.wait_stream(st)
for _ in range(3): step()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(10): step()
t1 = time.perf_counter()
torch.cuda.synchronize()
t2 = time.perf_counter()
print(f"host issue {1e3*(t1-t0)/10:.3f} ms/step, total {1e3*(t2-t0)/10:.3f} ms/step")

# ---- the same step captured once into a CUDA graph and replayed
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    step()
torch.cuda.synchronize()
for _ in range(3): g.replay()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0 = time.perf_counter()
e0.record()
for _ in range(10): g.replay()
e1.record()
t1 = time.perf_counter()
torch.cuda.synchronize()
print(f"graph replay: host issue {1e3*(t1-t0)/10:.3f} ms/step, device {e0.elapsed_time(e1)/10:.3f} ms/step")
