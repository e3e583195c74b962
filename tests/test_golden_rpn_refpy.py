"""CPU: the oracle's RPN proposal stage against fixtures produced by RUNNING the reference's own
`OrientedRPNHead._get_bboxes_single`, `MidpointOffsetCoder.decode` and `AnchorGenerator.grid_anchors` on the Jittor shim
(tests/golden/rpn_refpy_golden.npz, generator tests/golden/make_golden_rpn_refpy.py).  The Python of the stage is the
reference's; `jt.nms` and `Var.argsort` underneath are Jittor builtins (third party, absent) played by the oracle's
restatement and torch's stable sort -- stated in the generator."""
import os

import numpy as np
import pytest

import workloads as W
from helpers import close_report

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "rpn_refpy_golden.npz")
SHAPES = ((48, 48), (24, 24), (12, 12), (6, 6), (3, 3))
STRIDES = (4, 8, 16, 32, 64)


@pytest.fixture(scope="module")
def g():
    return dict(np.load(GOLD))


def test_anchor_grid_vs_reference(oracle, g):
    for l, (s, st) in enumerate(zip(SHAPES, STRIDES)):
        assert np.array_equal(oracle.anchor_grid(s, st), g["anchors_l%d" % l])


@pytest.mark.parametrize("tag", ["sig", "soft", "all"])
def test_rpn_stage_vs_reference(oracle, g, tag):
    sigmoid, nms_pre, nms_post, min_size, seed = g["cfg_" + tag]
    cls, reg = W.rpn_outputs(SHAPES, 3, int(seed), 1 if sigmoid else 2)
    anchors = [g["anchors_l%d" % l] for l in range(len(SHAPES))]
    got = oracle.rpn_get_bboxes_single(cls, reg, anchors, bool(sigmoid), int(nms_pre), int(nms_post), 0.8, float(min_size))
    want = g["dets_" + tag]
    assert got.shape == want.shape and want.shape[0] > 150
    # torch's float32 exp / atan2 / sin / cos (the shim) against numpy's: an ulp or two, coordinates ~2e2
    bad, err, _ = close_report(got[:, :4], want[:, :4], 1e-5, 1e-3)
    assert bad == 0, err
    dth = np.abs(got[:, 4] - want[:, 4])
    assert np.minimum(dth, np.pi - dth).max() < 1e-4
    bad, err, _ = close_report(got[:, 5], want[:, 5], 1e-6, 1e-7)
    assert bad == 0, err
