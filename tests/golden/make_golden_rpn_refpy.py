#!/usr/bin/env python
"""Generate tests/golden/rpn_refpy_golden.npz by RUNNING the reference's own oriented RPN proposal stage --
`OrientedRPNHead._get_bboxes_single` (models/roi_heads/oriented_rpn_head.py:132-216), `MidpointOffsetCoder.decode`
(models/boxes/coder.py) and `AnchorGenerator.grid_anchors` (models/boxes/anchor_generator.py), imported by path from
/root/reference, nothing copied -- on tests/jittor_shim.

What is and is not pinned: the head's Python -- level loop, sigmoid / softmax score choice, top
`nms_pre` per level, concatenation order, decode, size filter, `obb2hbb`, the level-offset trick, truncation to
`nms_post` -- is the reference's own code.  Two Jittor BUILTINS it calls are third-party code that is absent here:
`Var.argsort(descending=True)` (played by torch's stable sort) and `jt.nms` (played by `oracle.jt_nms`, a restatement
of Jittor 1.3.4.7's published `misc.nms`).  The head is instantiated WITHOUT running its constructor (the constructor
builds convolution layers and loss modules that the proposal stage never touches); the attributes `_get_bboxes_single`
reads are set by hand to the Oriented R-CNN config values.

    python tests/golden/make_golden_rpn_refpy.py
"""
import importlib.util
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "jittor_shim"))
import jittor as jt  # noqa: E402  (the shim)
import workloads as W  # noqa: E402
from oracle import oracle as O  # noqa: E402

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


# the two Jittor builtins the stage ends in (third party, absent): see the docstring
def _nms(dets, thresh):
    keep = O.jt_nms(dets.numpy(), float(thresh))
    return jt.array(np.ascontiguousarray(keep).astype(np.int64))


jt.nms = _nms
if not hasattr(jt.misc, "_pair"):
    jt.misc._pair = lambda x: tuple(x) if isinstance(x, (tuple, list)) else (x, x)

for p in ("jdet", "jdet.ops", "jdet.utils", "jdet.models", "jdet.models.boxes", "jdet.models.roi_heads"):
    pkg(p)
load("jdet.utils.registry", "utils/registry.py")
load("jdet.utils.general", "utils/general.py")
load("jdet.ops.bbox_transforms", "ops/bbox_transforms.py")
load("jdet.models.boxes.box_ops", "models/boxes/box_ops.py")
CODER = load("jdet.models.boxes.coder", "models/boxes/coder.py")
AG = load("jdet.models.boxes.anchor_generator", "models/boxes/anchor_generator.py")
at = types.ModuleType("jdet.models.boxes.anchor_target")          # training-side helpers, not on this path
at.images_to_levels = at.anchor_inside_flags = None
sys.modules[at.__name__] = at
RPN = load("jdet.models.roi_heads.oriented_rpn_head", "models/roi_heads/oriented_rpn_head.py")

g = {}
SHAPES = ((48, 48), (24, 24), (12, 12), (6, 6), (3, 3))
STRIDES = [4, 8, 16, 32, 64]
gen = AG.AnchorGenerator(strides=STRIDES, ratios=[0.5, 1.0, 2.0], scales=[8])
anchors = gen.grid_anchors([tuple(s) for s in SHAPES])
for l, a in enumerate(anchors):
    g["anchors_l%d" % l] = np.ascontiguousarray(a.numpy(), dtype=np.float32)
for tag, sigmoid, nms_pre, nms_post, min_size, seed in (("sig", True, 600, 400, 0, 21), ("soft", False, 300, 200, 6.0, 22),
                                                         ("all", True, 100000, 100000, -1, 23)):
    cls, reg = W.rpn_outputs(SHAPES, 3, seed, 1 if sigmoid else 2)
    head = object.__new__(RPN.OrientedRPNHead)                     # no constructor: see the docstring
    head.use_sigmoid_cls, head.reg_dim = sigmoid, 6
    head.nms_pre, head.nms_post, head.nms_thresh, head.min_bbox_size = nms_pre, nms_post, 0.8, min_size
    head.bbox_coder = CODER.MidpointOffsetCoder(target_means=[.0, .0, .0, .0, .0, .0], target_stds=[1.0, 1.0, 1.0, 1.0, 0.5, 0.5])
    dets = RPN.OrientedRPNHead._get_bboxes_single(head, [jt.array(c) for c in cls], [jt.array(r) for r in reg],
                                                  anchors, (192, 192, 3))
    g["dets_" + tag] = np.ascontiguousarray(dets.numpy(), dtype=np.float32)
    g["cfg_" + tag] = np.array([int(sigmoid), nms_pre, nms_post, min_size, seed], np.float64)
np.savez_compressed(os.path.join(HERE, "rpn_refpy_golden.npz"), **g)
print({k: v.shape for k, v in g.items()})
