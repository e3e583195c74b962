#!/usr/bin/env python
"""Generate tests/golden/merge_golden.npz by RUNNING the reference's own Python for the ensemble / format part of
the merge stage: `/root/reference/merge.py` (`nms`, `poly2obb`, `obb2hbb`,
`read_csv_to_numpy`, `merge_csv_with_class`, `merge_csv_without_class`, `save_to_csv`) and
`/root/reference/python/jdet/data/devkits/dota_to_fair.py` (`pick_res`, `dota_to_fair1m_1_5`).  Both files
import with what this image has (numpy, cv2, pandas, tqdm); they are loaded BY PATH, nothing is copied.  Needs
/root/reference (build container only); the .npz is committed and is what travels to the GPU box.

    python tests/golden/make_golden_merge.py
"""
import contextlib
import importlib.util
import io
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import workloads as W  # noqa: E402

REF = os.environ.get("RSDET_REFERENCE", "/root/reference")


def load(path, name):
    spec = importlib.util.spec_from_file_location(name, path)
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


M = load(os.path.join(REF, "merge.py"), "ref_merge")
D = load(os.path.join(REF, "python/jdet/data/devkits/dota_to_fair.py"), "ref_dota_to_fair")

g = {}
rng = np.random.default_rng(20261017)

# ---- merge.py:14-27 `nms` on horizontal boxes: distinct scores, 4-decimal (tied) scores, touching boxes
def hbbs(n, canvas, tied):
    c = rng.uniform(0, canvas, (n, 2))
    wh = rng.uniform(4, 60, (n, 2))
    s = rng.uniform(0.05, 1.0, n)
    if tied:
        s = np.round(s, 2)            # many exact ties: the reference order is `argsort()[::-1]`
    return np.concatenate([c - wh / 2, c + wh / 2, s[:, None]], 1)


for tag, n, canvas, tied in (("a", 2000, 800, False), ("b", 1000, 320, True), ("c", 64, 60, True)):
    b = hbbs(n, canvas, tied)
    if tag == "c":                      # edge cases: exact duplicates, boxes that only touch, zero-area boxes
        b[1] = b[0]
        b[3, :4] = [b[2, 2], b[2, 1], b[2, 2] + 10, b[2, 3]]
        b[5, 2:4] = b[5, :2]
        b[7, 2:4] = b[7, :2]               # two zero-area boxes: iou = 0/0 = NaN -> `NaN < thresh` is False
        b[9, 2] = b[9, 0]                   # zero width only
    g[f"nms_{tag}_boxes"] = b
    for thr in (0.625, 0.3):
        g[f"nms_{tag}_keep_{thr}"] = M.nms(b.copy(), thr).astype(np.int64)

# ---- merge.py:73-111 poly2obb (cv2.minAreaRect) / obb2hbb
obb = W.rotated_boxes(400, 5, canvas=1000, smin=6, smax=200, rmax=6, dtype=np.float64)
polys = np.round(W.obb_to_poly64(obb), 4)
polys[:5] = np.round(polys[:5])        # axis-aligned-ish / integer corner cases for minAreaRect
g["p2o_polys"] = polys
g["p2o_obb"] = M.poly2obb(polys.copy())
g["p2o_hbb"] = M.obb2hbb(g["p2o_obb"])

# ---- two synthetic submissions -> CSV text -> read back -> ensemble
def submission(seed, n_img=4, per_img=150):
    r = np.random.default_rng(seed)
    rows = []
    base = W.rotated_boxes(per_img, 99, canvas=1000, smin=8, smax=120, rmax=5, dtype=np.float64)   # shared objects
    cls = np.random.default_rng(7).integers(1, 11, per_img)
    for img in range(n_img):
        o = base.copy()
        o[:, :2] += r.normal(0, 1.5, (per_img, 2))
        o[:, 2:4] *= np.exp(r.normal(0, 0.05, (per_img, 2)))
        o[:, 4] += r.normal(0, 0.03, per_img)
        keep = r.uniform(size=per_img) < 0.8
        sc = np.round(r.uniform(0.05, 1.0, per_img), 4)
        p = W.obb_to_poly64(o)
        for k in np.nonzero(keep)[0]:
            rows.append([10 + 3 * img, *p[k], sc[k], cls[k]])
    return np.array(rows)


with tempfile.TemporaryDirectory() as td:
    subs = []
    for i, seed in enumerate((1, 2)):
        path = os.path.join(td, f"s{i}.csv")
        M.save_to_csv(submission(seed), path)
        g[f"csv_text_{i}"] = np.frombuffer(open(path, "rb").read(), np.uint8)
        subs.append(M.read_csv_to_numpy(path))
        g[f"csv_rows_{i}"] = subs[-1]
    with contextlib.redirect_stdout(io.StringIO()), contextlib.redirect_stderr(io.StringIO()):
        g["ens_with_class_0.625"] = M.merge_csv_with_class([s.copy() for s in subs], 0.625)
        thr_d = {c: 0.3 + 0.05 * i for i, c in enumerate(M.FAIR1M_1_5_CLASSES)}
        g["ens_with_class_dict"] = M.merge_csv_with_class([s.copy() for s in subs], thr_d)
        g["ens_without_class_0.9"] = M.merge_csv_without_class([s.copy() for s in subs], 0.9)
    out = os.path.join(td, "merged.csv")
    M.save_to_csv(g["ens_with_class_0.625"], out)
    g["ens_csv_text"] = np.frombuffer(open(out, "rb").read(), np.uint8)

    # ---- dota_to_fair.py:6-33,102-116: after_nms/<Class>.txt -> FAIR1M-1.5 CSV
    img_dir, src_dir, dst_dir = (os.path.join(td, d) for d in ("images", "after_nms", "csv"))
    os.makedirs(img_dir)
    os.makedirs(src_dir)
    scenes = ["P0012", "P0007", "P0130"]
    for s in scenes:
        for x in (0, 824):
            open(os.path.join(img_dir, f"{s}__1.0__{x}___0.png"), "w").close()
    files = {}
    for ci, cname in enumerate(("Vehicle", "Tennis_Court", "Ship")):
        lines = []
        for k in range(40):
            p = np.round(W.obb_to_poly64(W.rotated_boxes(1, 1000 + 50 * ci + k, canvas=3000, smin=8, smax=90, dtype=np.float64))[0], 1)
            sc = round(float(rng.uniform(0.05, 1)), 4)
            lines.append(f"{scenes[k % 3]} {sc} " + " ".join(str(float(v)) for v in p))
        files[cname] = "\n".join(lines) + "\n"
        open(os.path.join(src_dir, cname + ".txt"), "w").write(files[cname])
    D.dota_to_fair1m_1_5(src_dir, dst_dir, img_dir, "sub")
    text = open(os.path.join(dst_dir, "sub.csv")).read()
    # os.walk order is filesystem dependent: the fixture stores the multiset of lines (sorted) and the inputs
    g["d2f_csv_sorted"] = np.frombuffer("".join(sorted(text.splitlines(True))).encode(), np.uint8)
    for cname, t in files.items():
        g[f"d2f_src_{cname}"] = np.frombuffer(t.encode(), np.uint8)
    g["d2f_scenes"] = np.frombuffer(",".join(scenes).encode(), np.uint8)

path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "merge_golden.npz")
np.savez_compressed(path, **g)
print("wrote", path, {k: getattr(v, "shape", None) for k, v in g.items()})
