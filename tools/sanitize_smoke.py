#!/usr/bin/env python
"""Small invocations of every kernel family, meant to run under compute-sanitizer (memcheck / racecheck /
initcheck): python tools/sanitize_smoke.py"""
import os, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import workloads as W
from rs_detection_b200 import core
from rs_detection_b200._lib import NMS_HBB, NMS_MERGE, NMS_POLY, NMS_ROTATED, NMS_ROTATED_GE
t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
if "--roi-persistent" in sys.argv:
    # the persistent 7x7 forward kernel with more items than CTAs (every CTA loops: next-record prefetch, in-place list
    # rebuild under the previous bulk store, counter hand-out), single chunk and two chunks, channels-last and NCHW
    for C, K in ((256, 2500), (512, 900)):
        shapes = W.fpn_shapes(1, tile=256, channels=C)
        feats = [torch.randn(sh, device="cuda") for sh in shapes]
        rois = t(W.proposals(K, 4, batch=1, canvas=256))
        for cl in (False, True):
            cfg = core.make_roi_cfg(shapes, [1 / s_ for s_ in W.STRIDES], 7, 2, 1, (1.4, 1.2), 56.0, channels_last=cl)
            f = [core.nchw_to_nhwc(x) for x in feats] if cl else feats
            core.roi_align_rotated_forward(cfg, f, rois)
    torch.cuda.synchronize()
    sys.exit(0)
n = 700
b = W.rotated_boxes(n, 1, canvas=300, smin=8, smax=96)
s = W.distinct_scores(n, 1)
lab = np.random.default_rng(0).integers(0, 5, n).astype(np.int32)
core.box_iou_rotated(t(b[:100]), t(b), 1, True)
core.assign_wrt_overlaps(core.box_iou_rotated(t(b[:37]), t(b), 0), 0.5, (0.1, 0.4), 0.3, True, True, t(lab[:37]))
for kind in (NMS_ROTATED, NMS_ROTATED_GE):
    core.nms(kind, t(b), t(s), 0.3, labels=t(lab), want_score=True).count
core.nms(NMS_ROTATED, t(W.rotated_boxes(9000, 2, canvas=2048, smin=8, smax=96)), t(W.distinct_scores(9000, 2)), 0.3).count  # coop scan
p = core.obb2poly(t(b)); core.obb2hbb(t(b)); core.poly2hbb(p)
core.nms(NMS_POLY, p[:300], t(s[:300]), 0.2, want_score=True).count
sc = W.merge_scene(num_objects=150, scene=1500, seed=2)
core.nms(NMS_MERGE, t(sc["polys"]), t(sc["scores"]), 0.1, labels=t(sc["labels"].astype(np.int32)), thr_per_label=t(np.full(10, 0.2)), want_score=True).count
core.iou_poly_pairs(t(sc["polys"][:50]), t(sc["polys"][1:51]))
hb = np.concatenate([np.random.default_rng(1).uniform(0, 100, (200, 2))] * 2, 1); hb[:, 2:] += 20
core.nms(NMS_HBB, t(hb), t(W.distinct_scores(200, 3).astype(np.float64)), 0.5, want_score=True).count
core.multiclass_nms_rotated(t(b), t(W.class_scores(n, 6, 1)), 0.05, 0.1, 100)[2].item()
core.multiclass_nms_rotated(t(np.concatenate([b] * 7, 1)), t(W.class_scores(n, 6, 1)), 0.05, 0.1, -1)[2].item()
for C in (256, 24, 6):
    shapes = W.fpn_shapes(2, tile=256, channels=C)
    feats = [torch.randn(sh, device="cuda") for sh in shapes]
    rois = t(W.proposals(300, 4, batch=2, canvas=256))
    cfg = core.make_roi_cfg(shapes, [1 / s_ for s_ in W.STRIDES], 7, 2, 1, (1.4, 1.2))
    out = core.roi_align_rotated_forward(cfg, feats, rois)
    core.roi_align_rotated_backward(cfg, torch.randn_like(out), rois, shapes)
cfg = core.make_roi_cfg([(1, 8, 32, 32)], [0.25], (3, 5), 0, 0)
out = core.roi_align_rotated_forward(cfg, [torch.randn((1, 8, 32, 32), device="cuda")], t(W.proposals(40, 5, canvas=128)))
core.roi_align_rotated_backward(cfg, torch.randn_like(out), t(W.proposals(40, 5, canvas=128)), [(1, 8, 32, 32)])
core.oriented_head_results(t(b), torch.randn((n, 11), device="cuda"), torch.randn((n, 5), device="cuda") * 0.1, 10, True, [0.] * 5, [0.1, 0.1, 0.2, 0.2, 0.1], 0.05, 1.5)[2].item()
# staged per-class scan with two 64-word chunks per row (n > 4096 boxes), and the L2-resident fallback (n > 8192)
for nb in (5000, 9000):
    bb = W.rotated_boxes(nb, 7, canvas=2048, smin=8, smax=96)
    core.multiclass_nms_rotated(t(bb), t(W.class_scores(nb, 3, 2)), 0.05, 0.1, 500)[2].item()
# RPN proposal stage (score, sort, decode, offsets, HBB_P1 NMS, output) and VOC matching
from rs_detection_b200.jdet.models.boxes.anchor_generator import AnchorGenerator
shp = ((32, 32), (16, 16), (8, 8))
cls_, reg_ = W.rpn_outputs(shp, 3, 1)
anc = AnchorGenerator(strides=[4, 8, 16], ratios=[0.5, 1.0, 2.0], scales=[8]).grid_anchors(shp)
core.rpn_proposals([t(x) for x in cls_], [t(x) for x in reg_], anc, 3, True, 500, 300, 0.8, 0)[1].item()
cls2, reg2 = W.rpn_outputs(shp, 3, 2, 2)
core.rpn_proposals([t(x) for x in cls2], [t(x) for x in reg2], anc, 3, False, 100, 50, 0.7, 4.0, want_candidates=True)[1].item()
# sparse merge path (n >= 8192: sweep, CSR, persistent fixed-point resolve) and the pair-queue overflow fallback of the
# decision matrix (dense cluster, 2 classes)
big = W.merge_scene(num_objects=2600, scene=4000, seed=3)
core.nms(NMS_MERGE, t(big["polys"]), t(big["scores"]), 0.1, labels=t(big["labels"].astype(np.int32)), want_score=True).count
rng = np.random.default_rng(5)
dc = np.empty((1500, 5), np.float32)
dc[:, :2] = 200 + rng.normal(0, 6.0, (1500, 2)); dc[:, 2] = rng.uniform(60, 90, 1500); dc[:, 3] = rng.uniform(25, 40, 1500)
dc[:, 4] = rng.uniform(-0.4, 0.4, 1500)
core.multiclass_nms_rotated(t(dc), t(W.class_scores(1500, 2, 3, 1.0)), 0.01, 0.6, 500)[2].item()
from rs_detection_b200._lib import NMS_HBB_P1, NMS_HBB_P1_F64
core.nms(NMS_HBB_P1, t(hb.astype(np.float32)), t(W.distinct_scores(200, 3)), 0.5, want_score=True).count
core.nms(NMS_HBB_P1_F64, t(hb), t(W.distinct_scores(200, 3).astype(np.float64)), 0.5, want_score=True).count
torch.cuda.synchronize()
print("sanitize smoke done")
# round 2, second pass: fast multiclass path with the two-chunk scan (n > 4096) and a sort that is not a power of two;
# specialised 7x7 RoI kernel without a locality order (K < 256) and channels-last
bb5 = W.rotated_boxes(5000, 8, canvas=2048, smin=8, smax=96)
core.multiclass_nms_rotated(t(bb5), t(W.class_scores(5000, 3, 4)), 0.05, 0.1, 2000, t(np.random.default_rng(2).uniform(0.5, 1, 5000).astype(np.float32)))[2].item()
shapes = W.fpn_shapes(1, tile=256, channels=256)
cfg = core.make_roi_cfg(shapes, [1 / s_ for s_ in W.STRIDES], 7, 2, 1, (1.4, 1.2), channels_last=True)
feats = [torch.randn((sh[0], sh[2], sh[3], sh[1]), device="cuda") for sh in shapes]
core.roi_align_rotated_forward(cfg, feats, t(W.proposals(100, 6, canvas=256)))
core.roi_align_rotated_forward(cfg, feats, t(W.proposals(400, 7, canvas=256)))
torch.cuda.synchronize()
print("sanitize_smoke: second-pass paths done")
