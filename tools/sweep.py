#!/usr/bin/env python
"""BASELINE.json configs 3-5 on one GPU: nms_rotated / ml_nms_rotated / box_iou_rotated sweeps (config 4),
training-step ops (config 3), full-scene merge NMS (config 5).  Prints one JSON object; device-timed with CUDA
events (min of `reps` after warm-up); CPU figures come from oracle/_ref (the reference's own source, 1 core)
on bounded sizes.
    python tools/sweep.py [--quick] [--cpu]"""
import argparse, json, os, sys, time
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import workloads as W
from rs_detection_b200 import core
from rs_detection_b200._lib import NMS_MERGE, NMS_ROTATED

ap = argparse.ArgumentParser()
ap.add_argument("--quick", action="store_true")
ap.add_argument("--cpu", action="store_true")
a = ap.parse_args()
dev = torch.device("cuda:0")
t = lambda x: torch.from_numpy(np.ascontiguousarray(x)).to(dev)


def timed(fn, reps=5):
    fn(); torch.cuda.synchronize()
    best = 1e30
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best


out = {"nms_rotated": [], "ml_nms_rotated": [], "box_iou_rotated": [], "merge": [], "train": {}}
sizes = [1000, 2000, 5000, 10000, 20000] + ([] if a.quick else [50000, 100000])
for n in sizes:
    for canvas in (1024, 4096):
        d = W.rotated_boxes(n, n, canvas=canvas, smin=8, smax=128)
        s = W.distinct_scores(n, n)
        dt, st = t(d), t(s)
        for thr in (0.1, 0.5):
            res = {}
            ms = timed(lambda: res.update(r=core.nms(NMS_ROTATED, dt, st, thr, want_mask=False, want_sorted=True)))
            kept = res["r"].count
            row = {"n": n, "canvas": canvas, "thr": thr, "ms": ms, "boxes_per_s": n / ms * 1e3, "kept": kept,
                   "pairs_per_s": n * (n - 1) / 2 / ms * 1e3}
            if a.cpu and n <= 10000:
                from oracle import ref as R
                order = np.argsort(-s.astype(np.float64), kind="stable").astype(np.int32)
                t0 = time.perf_counter(); k = R.nms_keep(d, order, thr, 5, ge=False); row["cpu_ref_ms_1core"] = (time.perf_counter() - t0) * 1e3
                row["cpu_kept"] = int(k.sum())
            out["nms_rotated"].append(row)
            print(row, file=sys.stderr)
    lab = t(np.random.default_rng(n).integers(0, 15, n).astype(np.int32))
    d = W.rotated_boxes(n, n, canvas=1024, smin=8, smax=128); s = W.distinct_scores(n, n); dt, st = t(d), t(s)
    ms = timed(lambda: core.nms(NMS_ROTATED, dt, st, 0.1, labels=lab, want_mask=False, want_sorted=True))
    out["ml_nms_rotated"].append({"n": n, "classes": 15, "thr": 0.1, "ms": ms, "boxes_per_s": n / ms * 1e3})
for n in [1000, 2000, 5000, 10000] + ([] if a.quick else [20000]):
    b = t(W.rotated_boxes(n, n + 1, canvas=1024, smin=8, smax=128))
    ms = timed(lambda: core.box_iou_rotated(b, b, 0))
    out["box_iou_rotated"].append({"n": n, "ms": ms, "pairs_per_s": n * n / ms * 1e3})
    print(out["box_iou_rotated"][-1], file=sys.stderr)
# config 3
shapes = W.fpn_shapes()
cfg = core.make_roi_cfg(shapes, [1.0 / s for s in W.STRIDES], 7, 2, 1, (1.4, 1.2), 56.0)
feats = [torch.randn(s, device=dev) for s in shapes]
rois = t(W.proposals(512, 3))
g = torch.randn((512, 256, 7, 7), device=dev)
out["train"]["roi_fwd_512_ms"] = timed(lambda: core.roi_align_rotated_forward(cfg, feats, rois))
out["train"]["roi_bwd_512_ms"] = timed(lambda: core.roi_align_rotated_backward(cfg, g, rois, shapes))
P = W.rotated_boxes(2000, 5)
for G in (8, 64, 512):
    gt = t(W.jittered_copies(P, G, 6)); pp = t(P)
    out["train"][f"iou_assign_G{G}_ms"] = timed(lambda: core.assign_wrt_overlaps(core.box_iou_rotated(gt, pp, 1, True), 0.5, 0.5, 0.5, False))
# config 5
for nobj in ([2000, 20000] if not a.quick else [2000]):
    sc = W.merge_scene(num_objects=nobj, scene=10000, seed=1)
    from rs_detection_b200.jdet.data.devkits.result_merge import nms_threshold_1
    thr = t(np.array([nms_threshold_1[c] for c in W.FAIR1M_CLASSES]))
    p, s, l = t(sc["polys"]), t(sc["scores"]), t(sc["labels"].astype(np.int32))
    res = {}
    ms = timed(lambda: res.update(r=core.nms(NMS_MERGE, p, s, 0.1, labels=l, thr_per_label=thr, want_mask=False, want_sorted=False, want_score=True, ws_tag="merge")), reps=3)
    row = {"objects": nobj, "detections": int(p.shape[0]), "tiles": sc["tiles"], "ms": ms, "boxes_per_s": p.shape[0] / ms * 1e3,
           "scenes_per_s": 1e3 / ms, "kept": res["r"].count, "class_counts": np.bincount(sc["labels"], minlength=10).tolist()}
    if a.cpu and nobj <= 2000:
        from oracle import oracle as O
        t0 = time.perf_counter()
        kk = 0
        for c in range(10):
            idx = np.nonzero(sc["labels"] == c)[0]
            kk += len(O.py_cpu_nms_poly_fast(np.concatenate([sc["polys"][idx], sc["scores"][idx, None]], 1), float(thr[c])))
        row["cpu_oracle_ms_1core"] = (time.perf_counter() - t0) * 1e3
        row["cpu_kept"] = kk
    out["merge"].append(row)
    print(row, file=sys.stderr)
# voc_eval matching, 200 images x 100 gts, 60k detections of one class
rng = np.random.default_rng(0)
nimg, ngt = 200, 100
g = W.obb_to_poly64(W.rotated_boxes(nimg * ngt, 9, canvas=1024, smin=12, smax=90, dtype=np.float64))
gt_start = torch.arange(0, nimg * ngt + 1, ngt, dtype=torch.int32, device=dev)
nd = 60000
src = rng.integers(0, nimg * ngt, nd)
dp = g[src] + rng.normal(0, 2.0, (nd, 8))
dimg = torch.from_numpy((src // ngt).astype(np.int32)).to(dev)
diff = torch.zeros((nimg * ngt,), dtype=torch.uint8, device=dev)
dpt, gtt = t(dp), t(g)
ms = timed(lambda: core.voc_match(dpt, dimg, gtt, gt_start, diff, 0.5))
out["voc_match"] = {"detections": nd, "images": nimg, "gts_per_image": ngt, "ms": ms, "detections_per_s": nd / ms * 1e3}
print(json.dumps(out))
