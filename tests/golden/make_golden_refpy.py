#!/usr/bin/env python
"""Generate tests/golden/refpy_golden.npz by RUNNING the reference's own Python -- imported by path from
/root/reference/python/jdet, nothing copied -- on top of tests/jittor_shim (a torch-backed stand-in for the slice of the
Jittor API these files use; Jittor itself is not installable here):

    ops/bbox_transforms.py        obb2poly, obb2hbb, poly2hbb, regular_theta, regular_obb, rectpoly2obb   (a10, f1, f2)
    models/boxes/coder.py         MidpointOffsetCoder.decode, OrientedDeltaXYWHTCoder.decode               (f1, f2)
    models/boxes/assigner.py      MaxIoUAssigner.assign_wrt_overlaps                                       (a6)
    models/roi_heads/oriented_head.py   OrientedHead.get_bboxes / get_results                              (f1)

What this pins: the reference's Python logic.  What it cannot pin: Jittor's own kernels (float `%`, softmax, exp) --
the shim evaluates them with torch's float32 CPU kernels and states the conventions it assumes in its docstring.

    python tests/golden/make_golden_refpy.py
"""
import importlib.util
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "jittor_shim"))
import jittor as jt  # noqa: E402  (the shim)
import workloads as W  # noqa: E402

REF = os.path.join(os.environ.get("RSDET_REFERENCE", "/root/reference"), "python", "jdet")


def pkg(name):
    m = types.ModuleType(name)
    m.__path__ = []
    sys.modules[name] = m
    return m


def load(name, rel):
    spec = importlib.util.spec_from_file_location(name, os.path.join(REF, rel))
    m = importlib.util.module_from_spec(spec)
    sys.modules[name] = m
    spec.loader.exec_module(m)
    parent, _, leaf = name.rpartition(".")
    setattr(sys.modules[parent], leaf, m)
    return m


for p in ("jdet", "jdet.ops", "jdet.utils", "jdet.models", "jdet.models.boxes", "jdet.models.utils", "jdet.models.roi_heads",
          "jdet.data", "jdet.data.devkits"):
    pkg(p)
REG = load("jdet.utils.registry", "utils/registry.py")
load("jdet.utils.general", "utils/general.py")
BT = load("jdet.ops.bbox_transforms", "ops/bbox_transforms.py")
load("jdet.models.boxes.box_ops", "models/boxes/box_ops.py")
CODER = load("jdet.models.boxes.coder", "models/boxes/coder.py")
# stubs for imports that are not on the fixture path (Shapely-backed merge NMS, conv blocks, IoU calculators)
sm = types.ModuleType("jdet.data.devkits.result_merge"); sm.py_cpu_nms_poly_fast = None
sys.modules[sm.__name__] = sm
mm = types.ModuleType("jdet.models.utils.modules"); mm.ConvModule = object
sys.modules[mm.__name__] = mm


@REG.BOXES.register_module()
class BboxOverlaps2D:       # the assigner's default iou_calculator; assign_wrt_overlaps never calls it
    pass


ASSIGN = load("jdet.models.boxes.assigner", "models/boxes/assigner.py")
HEAD = load("jdet.models.roi_heads.oriented_head", "models/roi_heads/oriented_head.py")

g = {}
rng = np.random.default_rng(20261018)
f32 = lambda a: np.ascontiguousarray(a, np.float32)
out = lambda v: np.ascontiguousarray(v.numpy())

# ---- a10: bbox_transforms.py:602-632
obb = W.rotated_boxes(300, 41, canvas=1024, smin=4, smax=400)
obb[:4, 4] = [0.0, np.pi / 2, -np.pi / 2, np.pi / 4]
g["tf_obb"] = obb
g["tf_obb2poly"] = out(BT.obb2poly(jt.array(obb)))
g["tf_obb2hbb"] = out(BT.obb2hbb(jt.array(obb)))
g["tf_poly2hbb"] = out(BT.poly2hbb(jt.array(g["tf_obb2poly"])))
# ---- :501-519 regular_theta / regular_obb (float `%`)
th = f32(rng.uniform(-7, 7, 400))
th[:6] = [-np.pi / 2, np.pi / 2, 0, np.pi, -np.pi, 3 * np.pi / 2]
g["rt_theta"] = th
g["rt_180"] = out(BT.regular_theta(jt.array(th)))
g["rt_360"] = out(BT.regular_theta(jt.array(th), mode='360', start=-np.pi))
ro = obb.copy(); ro[:, 4] = f32(rng.uniform(-4, 4, 300)); ro[::7, 2:4] = ro[::7, 3:1:-1]
g["ro_in"], g["ro_out"] = ro, out(BT.regular_obb(jt.array(ro)))
# ---- :577-599 rectpoly2obb
rect = g["tf_obb2poly"] + f32(rng.normal(0, 0.01, g["tf_obb2poly"].shape))
g["rp_in"], g["rp_out"] = rect, out(BT.rectpoly2obb(jt.array(rect)))

# ---- f2: coder.py:373-433 MidpointOffsetCoder.decode (oriented RPN)
n = 300
c = rng.uniform(0, 1024, (n, 2)); wh = 2.0 ** rng.uniform(3, 8, (n, 2))
anchors = f32(np.concatenate([c - wh / 2, c + wh / 2], 1))
pred6 = f32(rng.normal(0, 0.5, (n, 6))); pred6[:5, 2:4] = 30.0; pred6[5:10, 4:6] = [[2, -2]] * 5
mo = CODER.MidpointOffsetCoder(target_means=(0., 0., 0., 0., 0., 0.), target_stds=(1., 1., 1., 1., 0.5, 0.5))
g["mo_anchors"], g["mo_pred"] = anchors, pred6
g["mo_decode"] = out(mo.decode(jt.array(anchors), jt.array(pred6)))

# ---- f1: coder.py:477-514 OrientedDeltaXYWHTCoder.decode (class-agnostic and per-class)
rois5 = W.rotated_boxes(n, 42, canvas=1024, smin=8, smax=300)
for tag, k in (("agn", 1), ("cls", 4)):
    d = f32(rng.normal(0, 0.6, (n, 5 * k))); d[:6, 2::5] = 40.0
    od = CODER.OrientedDeltaXYWHTCoder(target_means=(0., 0., 0., 0., 0.), target_stds=(0.1, 0.1, 0.2, 0.2, 0.1))
    g[f"od_{tag}_pred"] = d
    g[f"od_{tag}_decode"] = out(od.decode(jt.array(rois5), jt.array(d)))
g["od_rois"] = rois5

# ---- a6: assigner.py:111-170 on real IoU matrices (rows = GT, columns = proposals) incl. ties and an all-zero column
from oracle import oracle as O  # noqa: E402  (only to produce a realistic overlaps matrix; the assigner under test is the reference's)
props = W.rotated_boxes(700, 43, canvas=512, smin=8, smax=128)
gts = W.jittered_copies(props, 24, 44)
ov = O.box_iou_rotated_v1(gts, props, 1).astype(np.float32)
ov[:, 5] = 0.0
ov[3, 9] = ov[7, 9] = 0.8125            # argmax tie over GTs
ov[2, 11] = ov[2, 12] = ov[2].max()     # gt_argmax tie over proposals
labels = rng.integers(1, 16, 24).astype(np.int32)
g["as_overlaps"], g["as_gt_labels"] = ov, labels
for tag, kw in (("rcnn", dict(pos_iou_thr=0.5, neg_iou_thr=0.5, min_pos_iou=0.5, match_low_quality=False, ignore_iof_thr=-1)),
                ("rpn", dict(pos_iou_thr=0.7, neg_iou_thr=0.3, min_pos_iou=0.3, match_low_quality=True, ignore_iof_thr=-1)),
                ("rpn_one", dict(pos_iou_thr=0.7, neg_iou_thr=0.3, min_pos_iou=0.3, match_low_quality=True, ignore_iof_thr=-1,
                                 gt_max_assign_all=False))):
    a = ASSIGN.MaxIoUAssigner(**kw)
    r = a.assign_wrt_overlaps(jt.array(ov), jt.array(labels))
    g[f"as_{tag}_gt_inds"] = out(r.gt_inds).astype(np.int32)
    g[f"as_{tag}_max_overlaps"] = out(r.max_overlaps)
    g[f"as_{tag}_labels"] = out(r.labels).astype(np.int32)

# ---- f1: oriented_head.py:498-536 get_bboxes -> :279-305 get_results (10 classes, background = LAST softmax column)
K, C = 400, 10
head = object.__new__(HEAD.OrientedHead)
head.loss_cls = object()      # custom_cls_channels is a property reading loss_cls: plain softmax branch
head.bbox_coder = CODER.OrientedDeltaXYWHTCoder(target_means=(0., 0., 0., 0., 0.), target_stds=(0.1, 0.1, 0.2, 0.2, 0.1))
head.start_bbox_type, head.end_bbox_type = 'obb', 'obb'
rois6 = W.proposals(K, 45)
cls_score = f32(rng.standard_normal((K, C + 1)) * 2.0)
for tag, k, thr, scale in (("agn", 1, 0.05, 1.5), ("cls", C, 0.001, [0.5, 0.5, 0.5, 0.5]), ("raw", 1, 0.05, None)):
    head.score_thresh = thr
    bp = f32(rng.standard_normal((K, 5 * k)) * 0.8); bp[:7, 2::5] = 40.0
    dets, labs = head.get_bboxes(jt.array(rois6), jt.array(cls_score), jt.array(bp), (1024, 1024),
                                 scale if scale is not None else 1.0, rescale=scale is not None)
    g[f"hd_{tag}_pred"], g[f"hd_{tag}_dets"], g[f"hd_{tag}_labels"] = bp, out(dets), out(labs).astype(np.int64)
    g[f"hd_{tag}_thr"] = np.float64(thr)
    g[f"hd_{tag}_scale"] = np.array([0.0] if scale is None else np.atleast_1d(scale), np.float64)
g["hd_rois"], g["hd_cls"] = rois6, cls_score

path = os.path.join(HERE, "refpy_golden.npz")
np.savez_compressed(path, **g)
print("wrote", path, os.path.getsize(path), "bytes;", {k: v.shape for k, v in g.items()})
