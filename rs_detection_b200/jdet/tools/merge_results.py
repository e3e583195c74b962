"""tools/merge_results.py:11-59 of the reference: multi-run ensembling of already-merged result files.

`merge_src_files` concatenates the per-class files of several runs; `merge_files` then applies the merge
NMS per scene.  Here every file and scene goes through ONE launch of the device engine.
"""
import glob
import os

import numpy as np
import torch

from ... import core
from ..._lib import NMS_MERGE, require_cuda
from ..data.devkits.result_merge import py_cpu_nms_poly_fast


def _read(src_file):
    names, order, dets = [], [], []
    with open(src_file, "r") as f:
        for line in f:
            sp = line.strip().split(' ')
            if len(sp) < 10:
                continue
            v = [float(x) for x in sp[1:]]
            if sp[0] not in order:
                order.append(sp[0])
            names.append(sp[0])
            dets.append(v[1:] + v[:1])  # [8 coords, score]  (:20-21)
    return names, order, np.asarray(dets, np.float64).reshape(-1, 9)


def merge_files(src_path, dst_path, nms_op=py_cpu_nms_poly_fast, nms_thr=0.1, process_num=37):
    """:38-45.  `process_num` is accepted and ignored (one launch replaces the pool)."""
    _merge(glob.glob(os.path.join(src_path, "*.txt")), dst_path, nms_op, nms_thr)


def merge_file(src_file, dst_path, nms_op, nms_thr=0.1):
    """:11-36"""
    _merge([src_file], dst_path, nms_op, nms_thr)


def _merge(files, dst_path, nms_op, nms_thr):
    if nms_op is not py_cpu_nms_poly_fast:
        raise ValueError("merge_results: only py_cpu_nms_poly_fast is implemented on the device")
    require_cuda()
    os.makedirs(dst_path, exist_ok=True)
    recs = [_read(f) for f in files]
    gid_of, gids = {}, []
    for fi, (names, order, _) in enumerate(recs):
        for sc in order:
            gid_of[(fi, sc)] = len(gid_of)
        gids.append(np.asarray([gid_of[(fi, sc)] for sc in names], np.int32))
    dets = np.concatenate([r[2] for r in recs]) if recs else np.zeros((0, 9))
    per_gid = [[] for _ in gid_of]
    if dets.shape[0]:
        g = np.concatenate(gids)
        t = torch.from_numpy(dets).cuda()
        for r in core.nms_grouped(NMS_MERGE, t[:, :8].contiguous(), t[:, 8].contiguous(), g, float(nms_thr)).tolist():
            per_gid[g[r]].append(r)
    for fi, f in enumerate(files):
        with open(os.path.join(dst_path, os.path.split(f)[-1]), "w") as fo:
            for sc in recs[fi][1]:
                for r in per_gid[gid_of[(fi, sc)]]:
                    d = dets[r].tolist()
                    fo.write(sc + ' ' + str(d[-1]) + ' ' + ' '.join(map(str, d[:-1])) + "\n")


def merge_src_files(src_paths, dst_path):
    """:47-59 -- append the per-class files of several runs (prefix `Task1_` dropped)."""
    os.makedirs(dst_path, exist_ok=True)
    for ipath in src_paths:
        for ff in glob.glob(ipath + "/*.txt"):
            filename = ff.split("/")[-1].replace("Task1_", "")
            with open(os.path.join(dst_path, filename), "a") as wf, open(ff) as f:
                for line in f.readlines():
                    wf.write(line)
