#!/usr/bin/env python
"""Generate tests/golden/rpn_golden.npz: outputs of the oracle's RPN proposal-stage restatement on seeded
inputs.  DRIFT GUARD ONLY -- the stage ends in jt.nms / jt.argsort (Jittor, not under
/root/reference, not installable), so nothing here is pinned to reference output; the anchors are pinned to
the reference's docstring example (tests/test_rpn_oracle.py).

    python tests/golden/make_golden_rpn.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import workloads as W  # noqa: E402
from oracle import oracle as O  # noqa: E402

SHAPES = ((48, 48), (24, 24), (12, 12), (6, 6), (3, 3))
STRIDES = (4, 8, 16, 32, 64)
g = {}
cls, reg = W.rpn_outputs(SHAPES, 3, 21)
anchors = [O.anchor_grid(s, st) for s, st in zip(SHAPES, STRIDES)]
p, s, ids, rows = O.rpn_candidates(cls, reg, anchors, True, 600, 0)
dets, keep, hb = O.rpn_level_offset_nms(p, s, ids, 0.8, 400)
g["cand_obb"], g["cand_score"], g["cand_level"], g["cand_rows"] = p, s, ids.astype(np.int32), rows.astype(np.int32)
g["keep"], g["dets"] = keep.astype(np.int32), dets
g["anchors_l2"] = anchors[2]
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "rpn_golden.npz"), **g)
print({k: v.shape for k, v in g.items()})
