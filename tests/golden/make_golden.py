#!/usr/bin/env python
"""Generate tests/golden/rotated_path_golden.npz from oracle/_ref -- the reference's OWN C++/CUDA source
strings (python/jdet/ops/{box_iou_rotated,box_iou_rotated_v1,nms_rotated,roi_align_rotated,
roi_align_rotated_v1,nms_poly}.py) compiled for the host by oracle/build_ref.py.  Needs /root/reference
(build container only); the .npz it writes is committed and is what travels.

    python tests/golden/make_golden.py

The `merge_*` entries are the exception: the reference delegates polygon intersection to Shapely/GEOS
(not available), so they are produced by the oracle restatement itself and only guard against drift
(parity UNPINNED, see oracle/rsdet_oracle.c).
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import workloads as W  # noqa: E402
from oracle import oracle as O  # noqa: E402
from oracle import ref as R  # noqa: E402

assert R.ensure_built(), "oracle/_ref could not be built (is /root/reference mounted?)"
g = {}
# --- rotated IoU (both conventions), 48 GT x 160 proposals, plus the reference's self-test boxes
P = W.rotated_boxes(160, 0)
G = W.jittered_copies(P, 48, 1)
g["iou_boxes1"], g["iou_boxes2"] = G, P
g["iou_v0"] = R.box_iou(G, P, 0)
g["iou_v1"] = R.box_iou(G, P, 1)
kat = np.array([[0, 0, 1, 1, 0], [0.5, 0.5, 1, 2, 0]], np.float32)  # box_iou_rotated.py:513-514
g["iou_kat_boxes"], g["iou_kat"] = kat, R.box_iou(kat, kat, 0)
# --- nms_rotated / ml_nms_rotated keep masks (CUDA rule `>` and CPU rule `>=`)
n = 500
d = W.rotated_boxes(n, 3, canvas=400, smin=16, smax=128)
s = W.distinct_scores(n, 3)
lab = np.random.default_rng(0).integers(0, 6, n).astype(np.int32)
order = np.argsort(-s.astype(np.float64), kind="stable").astype(np.int32)
g["nms_dets"], g["nms_scores"], g["nms_labels"] = d, s, lab
for thr in (0.1, 0.5):
    g[f"nms_keep_gt_{thr}"] = R.nms_keep(d, order, thr, 5, ge=False)
    g[f"nms_keep_ge_{thr}"] = R.nms_keep(d, order, thr, 5, ge=True)
d6 = np.concatenate([d, lab[:, None].astype(np.float32)], 1)
g["mlnms_keep_0.1"] = R.nms_keep(d6, order, 0.1, 6, ge=False)
# --- RoIAlignRotated v0 / v1 forward + backward
rng = np.random.default_rng(0)
feat = rng.standard_normal((2, 8, 24, 24)).astype(np.float32)
rois = W.proposals(20, 0, batch=2, canvas=384)
rois[:3, 3:5] = [[0.5, 0.5], [2000, 20], [1, 700]]
grad = rng.standard_normal((20, 8, 7, 7)).astype(np.float32)
g["roi_feat"], g["roi_rois"], g["roi_grad"] = feat, rois, grad
for v in (0, 1):
    g[f"roi_fwd_v{v}"] = R.roi_fwd(feat, rois, (7, 7), 1 / 16, 2, v)
    g[f"roi_bwd_v{v}"] = R.roi_bwd(grad, rois, feat.shape, 1 / 16, 2, v)
g["roi_fwd_v1_adaptive"] = R.roi_fwd(feat, rois, (7, 7), 1 / 16, 0, 1)
# --- poly IoU / poly NMS
pp = O.obb2poly(W.rotated_boxes(120, 5, canvas=300, smin=16, smax=128))
ps = W.distinct_scores(120, 5)
g["poly_polys"], g["poly_scores"] = pp, ps
g["poly_iou"] = R.poly_iou_matrix(pp[:40], pp[40:80])
po = np.argsort(-ps.astype(np.float64), kind="stable")
g["poly_nms_keep_0.1"] = po[R.poly_nms_sorted_keep(np.concatenate([pp, ps[:, None]], 1)[po], 0.1)]
# --- merge stage (oracle restatement; UNPINNED)
sc = W.merge_scene(num_objects=60, scene=1500, seed=9)
dets = np.concatenate([sc["polys"], sc["scores"][:, None]], 1)
g["merge_dets"], g["merge_labels"] = dets, sc["labels"].astype(np.int32)
for thr in (0.1, 0.3):
    g[f"merge_keep_{thr}"] = np.array(O.py_cpu_nms_poly_fast(dets, thr), np.int64)
out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "rotated_path_golden.npz")
np.savez_compressed(out, **g)
print("wrote", out, os.path.getsize(out), "bytes;", len(g), "arrays")
