#!/usr/bin/env python
"""Generate tests/golden/devkit_golden.npz by RUNNING the reference's own merge / evaluation Python -- imported by path
from /root/reference, nothing copied:

    python/jdet/data/devkits/result_merge.py   py_cpu_nms_poly_fast, py_cpu_nms_poly, py_cpu_nms, nmsbynamedict,
                                               poly2origpoly, mergesingle (threshold types 0 and 1), mergebase,
                                               mergebyrec                                            (rows a11, a12)
    tools/merge_results.py                     merge_file, merge_files                               (row a13)
    python/jdet/data/devkits/voc_eval.py       voc_eval_dota, voc_ap                                 (row f3)
    python/jdet/data/devkits/data_merge.py     flip_box, prepare_data (the before_nms writer), data_merge
                                               (= prepare_data + mergebypoly through its 16-process pool)   (row f4)

Their third-party imports are played by shims: `jittor` by tests/jittor_shim, `shapely.geometry.Polygon` by
tests/shapely_shim (exact rational quadrilateral clipping rounded once to float64 -- GEOS itself is not installable;
every decision of the fixtures is checked to lie further than 1e-6 from its threshold, so any float64 polygon library
decides the same).  What this pins: the reference's control flow, parsing, thresholds tables, tile -> scene mapping,
tie-free ordering and text formatting, byte for byte.

    python tests/golden/make_golden_devkit.py
"""
import importlib.util
import json
import os
import sys
import tempfile
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "tests", "jittor_shim"))
sys.path.insert(0, os.path.join(ROOT, "tests", "shapely_shim"))
import jittor as jt  # noqa: E402,F401  (the shim)
import workloads as W  # noqa: E402
from oracle import formats as F  # noqa: E402  (only its before_nms writer, to produce INPUT files)

REFROOT = os.environ.get("RSDET_REFERENCE", "/root/reference")
REF = os.path.join(REFROOT, "python", "jdet")


def pkg(name):
    m = types.ModuleType(name)
    m.__path__ = []
    sys.modules[name] = m
    return m


def load(name, path):
    spec = importlib.util.spec_from_file_location(name, path)
    m = importlib.util.module_from_spec(spec)
    sys.modules[name] = m
    spec.loader.exec_module(m)
    parent, _, leaf = name.rpartition(".")
    if parent:
        setattr(sys.modules[parent], leaf, m)
    return m


for p in ("jdet", "jdet.ops", "jdet.utils", "jdet.config", "jdet.data", "jdet.data.devkits", "jdet.models", "jdet.models.boxes"):
    pkg(p)
CFG = types.SimpleNamespace(merge_nms_threshold_type=0)
sys.modules["jdet.config"].get_cfg = lambda: CFG
load("jdet.utils.registry", os.path.join(REF, "utils/registry.py"))
load("jdet.utils.general", os.path.join(REF, "utils/general.py"))
load("jdet.ops.bbox_transforms", os.path.join(REF, "ops/bbox_transforms.py"))
NP = load("jdet.ops.nms_poly", os.path.join(REF, "ops/nms_poly.py"))          # iou_poly on the Shapely shim
nr = types.ModuleType("jdet.ops.nms_rotated")                                  # only py_cpu_nms_obb needs it (not on this fixture's path)
nr.nms_rotated_cpu = nr.nms_rotated_cuda = None
sys.modules[nr.__name__] = nr
load("jdet.data.devkits.dota_utils", os.path.join(REF, "data/devkits/dota_utils.py"))
RM = load("jdet.data.devkits.result_merge", os.path.join(REF, "data/devkits/result_merge.py"))
load("jdet.data.devkits.dota_to_fair", os.path.join(REF, "data/devkits/dota_to_fair.py"))
VE = load("jdet.data.devkits.voc_eval", os.path.join(REF, "data/devkits/voc_eval.py"))
MR = load("ref_tools_merge_results", os.path.join(REFROOT, "tools/merge_results.py"))
load("jdet.config.constant", os.path.join(REF, "config/constant.py"))
load("jdet.models.boxes.box_ops", os.path.join(REF, "models/boxes/box_ops.py"))
if not hasattr(jt, "load"):
    import pickle
    jt.load = lambda path: pickle.load(open(path, "rb"))     # jt.load of a results pickle
DM = load("jdet.data.devkits.data_merge", os.path.join(REF, "data/devkits/data_merge.py"))

g = {}


def read_dir(d):
    return {f: open(os.path.join(d, f)).read() for f in sorted(os.listdir(d))}


def margin_of(dets, thr):
    """smallest |iou - thr| over the pairs the greedy loop decides (exact arithmetic)"""
    import exact_geometry as X
    d = np.asarray(dets, np.float64)
    _, m = X.greedy_merge_nms_exact(d, thr, np.argsort(d[:, 8], kind="stable")[::-1])
    return m if m is not None else 1.0


with tempfile.TemporaryDirectory() as tmp:
    # ---- inputs: before_nms/<Class>.txt written from synthetic tile results (distinct four-decimal scores: the
    #      reference orders with numpy's unstable argsort, ties are outside what a fixture can pin)
    res = W.tile_results(260, 10, 2600, 2, seed=5)
    src = os.path.join(tmp, "before_nms")
    F.write_before_nms(res, src, W.FAIR1M_CLASSES)
    inputs = read_dir(src)
    g["before_nms"] = json.dumps(inputs)
    # ---- mergesingle / mergebase, threshold types 0 and 1 (result_merge.py:177-243)
    for ttype in (0, 1):
        CFG.merge_nms_threshold_type = ttype
        dst = os.path.join(tmp, "after_nms_%d" % ttype)
        os.makedirs(dst)
        RM.mergebase(src, dst, RM.py_cpu_nms_poly_fast)
        g["after_nms_type%d" % ttype] = json.dumps(read_dir(dst))
    CFG.merge_nms_threshold_type = 0
    # ---- the slow all-pairs variant must agree with the hbb-prefiltered one on a file (result_merge.py:30-64)
    dst = os.path.join(tmp, "after_slow")
    os.makedirs(dst)
    RM.mergesingle(dst, RM.py_cpu_nms_poly, os.path.join(src, "Ship.txt"))
    g["after_nms_slow_ship"] = open(os.path.join(dst, "Ship.txt")).read()
    # ---- horizontal task: four-coordinate rows through py_cpu_nms (mergebyrec, result_merge.py:246-257)
    hsrc = os.path.join(tmp, "before_hbb")
    os.makedirs(hsrc)
    for name in ("Ship.txt", "Bridge.txt"):
        rows = []
        for ln in inputs[name].strip().split("\n"):
            sp = ln.split(" ")
            xy = np.array(list(map(float, sp[2:]))).reshape(4, 2)
            rows.append("%s %s %.4f %.4f %.4f %.4f" % (sp[0], sp[1], xy[:, 0].min(), xy[:, 1].min(), xy[:, 0].max(), xy[:, 1].max()))
        open(os.path.join(hsrc, name), "w").write("\n".join(rows) + "\n")
    g["before_hbb"] = json.dumps(read_dir(hsrc))
    dst = os.path.join(tmp, "after_hbb")
    os.makedirs(dst)
    RM.mergebyrec(hsrc, dst)
    g["after_hbb"] = json.dumps(read_dir(dst))
    # ---- tools/merge_results.py: merge_file on scene-level rows (the after_nms format), one process
    dst = os.path.join(tmp, "tool_out")
    MR.merge_files(os.path.join(tmp, "after_nms_0"), dst, nms_thr=0.05, process_num=1)
    g["tool_merge_files_thr005"] = json.dumps(read_dir(dst))

# ---- data_merge.py: the runner's result list -> before_nms text -> merged files (prepare_data :29-48, data_merge :50-54)
with tempfile.TemporaryDirectory() as tmp:
    import pickle
    res2 = W.tile_results(150, 10, 2400, 2, seed=9)
    # flipped tiles: `w - box[i]` (data_merge.py:14-27) is float64 arithmetic under the reference's NumPy 1.x (python int
    # with a float32 scalar) and float32 arithmetic under NumPy >= 2 (NEP 50) -- this container has NumPy 2.  The
    # flipped tiles therefore carry coordinates on a 1/16 grid, where both are exact; product and oracle follow the
    # NumPy 1.x rule (float64) for anything else.
    for t, mode in ((3, "H"), (5, "HV")):
        (polys, sc, lab), target = res2[t]
        res2[t] = ((np.round(polys * 16) / 16).astype(np.float32), sc, lab), dict(target, flip_mode=mode, ori_img_size=(1024, 1024))
    pickle.dump(res2, open(os.path.join(tmp, "res.pkl"), "wb"))
    g["results_pkl"] = np.frombuffer(pickle.dumps(res2), np.uint8)
    CFG.merge_nms_threshold_type = 0
    DM.data_merge(os.path.join(tmp, "res.pkl"), os.path.join(tmp, "before"), os.path.join(tmp, "after"), "FAIR1M_1_5")
    g["dm_before_nms"] = json.dumps(read_dir(os.path.join(tmp, "before")))
    g["dm_after_nms"] = json.dumps(read_dir(os.path.join(tmp, "after")))
    g["dm_classes"] = json.dumps(list(sys.modules["jdet.config.constant"].get_classes_by_name("FAIR1M_1_5")))

# ---- array level: keep lists + decision margins
d = []
for ln in inputs["Vehicle.txt"].strip().split("\n")[:400]:
    sp = ln.split(" ")
    name, x, y, rate = sp[0].split("__")[0], *[0, 0, 0]
    d.append(list(map(float, sp[2:])) + [float(sp[1])])
dets = np.array(d, np.float64)
g["nms_dets"] = dets
for thr in (0.1, 0.3):
    g["nms_keep_fast_%g" % thr] = np.array(RM.py_cpu_nms_poly_fast(dets, thr), np.int64)
    g["nms_keep_slow_%g" % thr] = np.array(RM.py_cpu_nms_poly(dets, thr), np.int64)
    g["nms_margin_%g" % thr] = np.float64(margin_of(dets, thr))
hb = np.stack([dets[:, 0:8:2].min(1), dets[:, 1:8:2].min(1), dets[:, 0:8:2].max(1), dets[:, 1:8:2].max(1), dets[:, 8]], 1)
g["hbb_dets"] = hb
g["hbb_keep_0.3"] = np.array(RM.py_cpu_nms(hb, 0.3), np.int64)
g["poly2origpoly"] = np.array(RM.poly2origpoly([10.0, 20.5, 30.25, 40.0, 50.0, 60.0, 70.0, 80.0], 824, 1648, "0.5"), np.float64)
byname = {"a": dets[:50].tolist(), "b": dets[50:120].tolist()}
kept = RM.nmsbynamedict(byname, RM.py_cpu_nms_poly_fast, 0.1)
g["nmsbynamedict_counts"] = np.array([len(kept["a"]), len(kept["b"])], np.int64)

# ---- voc_eval_dota (voc_eval.py:236-318) with the reference's own iou_poly on the Shapely shim
rng = np.random.default_rng(31)
gts, dl = {}, []
for im in range(6):
    gb = W.rotated_boxes(40, 900 + im, canvas=700, smin=12, smax=90, dtype=np.float64)
    gts[im] = {"box": W.obb_to_poly64(gb), "difficult": rng.uniform(size=40) < 0.15}
    for k in range(2):
        j = gb.copy()
        j[:, :2] += rng.normal(0, 0.12, (40, 2)) * np.sqrt(j[:, 2:3] * j[:, 3:4])
        j[:, 2:4] *= np.exp(rng.normal(0, 0.15, (40, 2)))
        j[:, 4] += rng.normal(0, 0.1, 40)
        dl.append(np.concatenate([np.full((40, 1), im), W.obb_to_poly64(j), rng.uniform(0.01, 1, (40, 1))], 1))
    fa = W.obb_to_poly64(W.rotated_boxes(25, 950 + im, canvas=700, smin=12, smax=90, dtype=np.float64))
    dl.append(np.concatenate([np.full((25, 1), im), fa, rng.uniform(0.01, 0.6, (25, 1))], 1))
gts[6] = {"box": np.zeros((0, 8)), "difficult": np.zeros((0,), bool)}
dl.append(np.concatenate([np.full((4, 1), 6), fa[:4], rng.uniform(0.5, 1, (4, 1))], 1))
vd = np.concatenate(dl)
vd[:, -1] += np.arange(len(vd)) * 1e-9                       # distinct confidences (np.argsort(-confidence) is unstable)
g["voc_dets"] = vd
g["voc_gt_boxes"] = np.concatenate([gts[k]["box"] for k in sorted(gts)])
g["voc_gt_img"] = np.concatenate([np.full(len(gts[k]["box"]), k) for k in sorted(gts)]).astype(np.int64)
g["voc_gt_difficult"] = np.concatenate([gts[k]["difficult"] for k in sorted(gts)])
for thr in (0.5, 0.3):
    for m07 in (False, True):
        gg = {k: {"box": v["box"].copy(), "difficult": v["difficult"].copy(), "det": [False] * len(v["box"])} for k, v in gts.items()}
        rec, prec, ap = VE.voc_eval_dota(vd.copy(), gg, NP.iou_poly, thr, m07)
        tag = "voc_%g_%s" % (thr, "07" if m07 else "area")
        g[tag + "_rec"], g[tag + "_prec"], g[tag + "_ap"] = np.asarray(rec), np.asarray(prec), np.float64(ap)

np.savez_compressed(os.path.join(HERE, "devkit_golden.npz"), **g)
print("wrote devkit_golden.npz:", {k: (v.shape if hasattr(v, "shape") and getattr(v, "shape", ()) else "scalar/str") for k, v in g.items()})
