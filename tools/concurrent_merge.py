#!/usr/bin/env python
"""Several large merge-NMS calls (sparse path: persistent kernel with a grid barrier) and large single-class rotated
NMS calls (cooperative scan) in flight on different streams at once; checks they all finish and agree with the
single-stream result.  Run under a short `timeout`."""
import os, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import workloads as W
from rs_detection_b200 import core
from rs_detection_b200._lib import NMS_MERGE, NMS_ROTATED
sc = W.merge_scene(num_objects=6000, scene=6000, seed=1)
p, s, l = [torch.from_numpy(np.ascontiguousarray(x)).cuda() for x in (sc["polys"], sc["scores"], sc["labels"].astype(np.int32))]
b = torch.from_numpy(W.rotated_boxes(12000, 2, canvas=2048, smin=8, smax=96)).cuda()
bs = torch.from_numpy(W.distinct_scores(12000, 2)).cuda()
ref_m = core.nms(NMS_MERGE, p, s, 0.1, labels=l, want_score=True, ws_tag="m0").score_idx.clone()
ref_r = core.nms(NMS_ROTATED, b, bs, 0.3, want_score=True, ws_tag="r0").score_idx.clone()
torch.cuda.synchronize()
streams = [torch.cuda.Stream() for _ in range(6)]
outs = []
for rep in range(3):
    for i, st in enumerate(streams):
        with torch.cuda.stream(st):
            if i % 2 == 0:
                outs.append(("m", core.nms(NMS_MERGE, p, s, 0.1, labels=l, want_score=True, ws_tag=f"m{i}")))
            else:
                outs.append(("r", core.nms(NMS_ROTATED, b, bs, 0.3, want_score=True, ws_tag=f"r{i}")))
torch.cuda.synchronize()
ok = all(torch.equal(r.score_idx, ref_m if k == "m" else ref_r) for k, r in outs)
print("concurrent calls:", len(outs), "all equal to the single-stream result:", ok)
sys.exit(0 if ok else 1)
