"""GPU: rsdet_rpn_proposals against the fixtures produced by the reference's OWN oriented RPN proposal stage
(tests/golden/rpn_refpy_golden.npz, generator tests/golden/make_golden_rpn_refpy.py)."""
import os

import numpy as np
import pytest
import torch

import workloads as W
from helpers import close_report

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "rpn_refpy_golden.npz")
SHAPES = ((48, 48), (24, 24), (12, 12), (6, 6), (3, 3))


@pytest.mark.parametrize("tag", ["sig", "soft", "all"])
def test_rpn_proposals_vs_reference_python(cuda, tag):
    from rs_detection_b200 import core
    g = dict(np.load(GOLD))
    sigmoid, nms_pre, nms_post, min_size, seed = g["cfg_" + tag]
    cls, reg = W.rpn_outputs(SHAPES, 3, int(seed), 1 if sigmoid else 2)
    anchors = [g["anchors_l%d" % l] for l in range(len(SHAPES))]
    cu = lambda xs: [torch.from_numpy(np.ascontiguousarray(x)).cuda() for x in xs]
    dets, cnt = core.rpn_proposals(cu(cls), cu(reg), cu(anchors), 3, bool(sigmoid), int(nms_pre), int(nms_post), 0.8, float(min_size))[:2]
    got = dets[: int(cnt.item())].cpu().numpy()
    want = g["dets_" + tag]
    assert got.shape == want.shape
    # CUDA's expf / atan2f / sinf / cosf against torch's CPU kernels: an ulp or two on coordinates ~2e2
    bad, err, _ = close_report(got[:, :4], want[:, :4], 1e-5, 2e-3)
    assert bad == 0, err
    dth = np.abs(got[:, 4] - want[:, 4])
    assert np.minimum(dth, np.pi - dth).max() < 1e-4
    bad, err, _ = close_report(got[:, 5], want[:, 5], 1e-6, 1e-7)
    assert bad == 0, err
