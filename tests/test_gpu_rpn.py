"""GPU parity: oriented RPN proposal stage through rsdet_rpn_proposals."""
import numpy as np
import pytest
import torch

import workloads as W
from helpers import close_report

pytestmark = pytest.mark.gpu

STRIDES = (4, 8, 16, 32, 64)


def _case(oracle, shapes, seed, cpa=1):
    cls, reg = W.rpn_outputs(shapes, 3, seed, cpa)
    anchors = [oracle.anchor_grid(s, st) for s, st in zip(shapes, STRIDES)]
    return cls, reg, anchors


def _cuda(xs):
    return [torch.from_numpy(x).cuda() for x in xs]


@pytest.mark.parametrize("shapes,nms_pre,nms_post,sigmoid,min_size", [
    (((64, 64), (32, 32), (16, 16), (8, 8), (4, 4)), 1000, 1000, True, 0),
    (((64, 64), (32, 32), (16, 16), (8, 8), (4, 4)), 500, 300, False, 6.0),
    (((40, 56), (20, 28), (10, 14)), 2000, 2000, True, -1),
    (((256, 256), (128, 128), (64, 64), (32, 32), (16, 16)), 2000, 2000, True, 0),
])
def test_rpn_proposals_vs_oracle(cuda, oracle, shapes, nms_pre, nms_post, sigmoid, min_size):
    from rs_detection_b200 import core
    cls, reg, anchors = _case(oracle, shapes, 11, 1 if sigmoid else 2)
    dets, cnt, c_obb, c_hbb, c_score, c_level = core.rpn_proposals(
        _cuda(cls), _cuda(reg), _cuda(anchors), 3, sigmoid, nms_pre, nms_post, 0.8, min_size, want_candidates=True)
    k = int(cnt.item())
    dets, c_obb, c_hbb, c_score, c_level = [t.cpu().numpy() for t in (dets, c_obb, c_hbb, c_score, c_level)]
    # stage 1: candidates (scores, top-k choice, decode, size filter) against the float32 numpy restatement.
    # expf / atan2f / sinf / cosf differ from glibc by an ulp or two, hence a tolerance (coordinates ~1e3).
    p, s, ids, rows = oracle.rpn_candidates(cls, reg, anchors, sigmoid, nms_pre, min_size)
    live = np.isfinite(c_score)
    near = 0
    if min_size >= 0:  # rows within 1e-3 of the size threshold may flip under the tolerance above
        near = int((np.abs(c_obb[:, 2:4] - min_size) < 1e-3).any(1).sum())
    if near == 0:
        assert np.array_equal(np.nonzero(live)[0], rows)
        assert np.array_equal(c_level[live], ids)
        bad, err, _ = close_report(c_score[live], s, 1e-6, 1e-7)
        assert bad == 0, err
        bad, err, _ = close_report(c_obb[live][:, :4], p[:, :4], 1e-5, 2e-3)
        assert bad == 0, err
        dth = np.abs(c_obb[live][:, 4] - p[:, 4])
        dth = np.minimum(dth, np.pi - dth)  # regular_theta wraps at +-pi/2
        assert dth.max() < 1e-4
    # stage 2a: obb2hbb + level offsets (sinf / cosf again: tolerance; offsets reach ~1e4 where an ulp is 1e-3)
    _, _, hb = oracle.rpn_level_offset_nms(c_obb[live], c_score[live], c_level[live].astype(np.int64), 0.8, nms_post)
    bad, err, _ = close_report(c_hbb[live], hb, 1e-6, 4e-3)
    assert bad == 0, err
    # stage 2b: jt.nms + truncation on the DEVICE's own offset boxes must match the oracle bit for bit
    keep = oracle.jt_nms(np.concatenate([c_hbb[live], c_score[live][:, None]], 1), 0.8)[:nms_post]
    want = np.concatenate([c_obb[live], c_score[live][:, None]], 1)[keep]
    assert k == want.shape[0] and 0 < k <= nms_post
    assert np.array_equal(dets[:k], want)
    assert not dets[k:].any()
    assert np.all(np.diff(dets[:k, 5]) <= 0)


def test_rpn_mirror_class_and_batch(cuda, oracle):
    from rs_detection_b200.jdet.models.roi_heads.oriented_rpn_head import OrientedRPNProposals
    shapes = ((32, 32), (16, 16), (8, 8), (4, 4), (2, 2))
    head = OrientedRPNProposals(nms_pre=300, nms_post=200)
    per_img = [W.rpn_outputs(shapes, 3, seed) for seed in (1, 2)]
    cls = [np.stack([per_img[b][0][l] for b in range(2)]) for l in range(5)]
    reg = [np.stack([per_img[b][1][l] for b in range(2)]) for l in range(5)]
    out = head.get_bboxes(cls, reg, None)
    assert len(out) == 2 and isinstance(out[0], np.ndarray) and out[0].shape[1] == 6
    anchors = [oracle.anchor_grid(s, st) for s, st in zip(shapes, STRIDES)]
    for b in range(2):
        want = oracle.rpn_get_bboxes_single(per_img[b][0], per_img[b][1], anchors, True, 300, 200, 0.8, 0)
        assert abs(out[b].shape[0] - want.shape[0]) <= 2  # a near-threshold pair may flip under the decode tolerance
        m = min(out[b].shape[0], want.shape[0], 20)
        assert np.allclose(out[b][:m, 5], want[:m, 5], atol=1e-6)


def test_rpn_bad_config(cuda):
    from rs_detection_b200 import core
    x = torch.zeros((3, 4, 4), device="cuda")
    with pytest.raises(AssertionError):
        core.rpn_proposals([x], [torch.zeros((17, 4, 4), device="cuda")], [torch.zeros((48, 4), device="cuda")], 3)


def test_rpn_vs_golden(cuda):
    """Committed fixture (tests/golden/make_golden_rpn.py): scores / levels / top-k rows exact, boxes within
    the decode tolerance, and at least 99 % of the fixture's kept rows kept (near-threshold pairs may flip)."""
    import os
    from rs_detection_b200 import core
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "rpn_golden.npz"))
    shapes = ((48, 48), (24, 24), (12, 12), (6, 6), (3, 3))
    cls, reg = W.rpn_outputs(shapes, 3, 21)
    from rs_detection_b200.jdet.models.boxes.anchor_generator import AnchorGenerator
    anchors = AnchorGenerator(strides=list(STRIDES), ratios=[0.5, 1.0, 2.0], scales=[8]).grid_anchors(shapes)
    assert np.array_equal(anchors[2].cpu().numpy(), g["anchors_l2"])
    dets, cnt, c_obb, _, c_score, c_level = core.rpn_proposals(_cuda(cls), _cuda(reg), anchors, 3, True, 600, 400, 0.8, 0,
                                                                want_candidates=True)
    c_obb, c_score, c_level = c_obb.cpu().numpy(), c_score.cpu().numpy(), c_level.cpu().numpy()
    live_rows = np.nonzero(np.isfinite(c_score))[0]
    # rows whose width / height sits within 1e-3 of the size threshold may fall on either side (sinf / cosf ulps)
    flipped = np.setxor1d(live_rows, g["cand_rows"])
    assert flipped.size <= 4 and np.all(c_obb[flipped][:, 2:4].min(1) < 1e-3)
    common, ia, ib = np.intersect1d(live_rows, g["cand_rows"], return_indices=True)
    assert np.array_equal(c_level[common], g["cand_level"][ib])
    assert np.allclose(c_score[common], g["cand_score"][ib], atol=1e-6)
    # candidates with scores one ulp apart may swap places (expf vs glibc exp): match boxes per level as sets
    lv = g["cand_level"][ib]
    for l in np.unique(lv):
        a, b = c_obb[common][lv == l][:, :4], g["cand_obb"][ib][lv == l][:, :4]
        d = np.abs(a[:, None, :] - b[None, :, :]).max(-1).min(1)
        assert d.max() < 5e-3, (l, d.max())
    k = int(cnt.item())
    got = dets[:k].cpu().numpy()
    assert abs(k - g["dets"].shape[0]) <= 4
    common = np.intersect1d(np.round(got[:, 5], 6), np.round(g["dets"][:, 5], 6)).size
    assert common >= 0.99 * g["dets"].shape[0]


def test_rpn_edge_cases(cuda, oracle):
    """nms_pre <= 0 keeps every anchor in its original order (oriented_rpn_head.py:183); a size filter that drops
    everything yields zero proposals and an all-zero output block; a single level needs no offsets."""
    from rs_detection_b200 import core
    shapes = ((12, 10), (6, 5))
    cls, reg = W.rpn_outputs(shapes, 3, 31)
    anchors = [oracle.anchor_grid(s, st) for s, st in zip(shapes, STRIDES)]
    dets, cnt, c_obb, c_hbb, c_score, c_level = core.rpn_proposals(_cuda(cls), _cuda(reg), _cuda(anchors), 3, True, 0, 50, 0.7, -1,
                                                                    want_candidates=True)
    n_all = sum(h * w * 3 for h, w in shapes)
    assert c_score.shape[0] == n_all
    _, s, ids, _ = oracle.rpn_candidates(cls, reg, anchors, True, 0, -1)
    assert np.allclose(c_score.cpu().numpy(), s, atol=1e-6) and np.array_equal(c_level.cpu().numpy(), ids)
    keep = oracle.jt_nms(np.concatenate([c_hbb.cpu().numpy(), c_score.cpu().numpy()[:, None]], 1), 0.7)[:50]
    k = int(cnt.item())
    assert k == len(keep)
    assert np.array_equal(dets[:k].cpu().numpy(), np.concatenate([c_obb.cpu().numpy(), c_score.cpu().numpy()[:, None]], 1)[keep])
    # everything filtered
    dets, cnt = core.rpn_proposals(_cuda(cls), _cuda(reg), _cuda(anchors), 3, True, 100, 50, 0.7, 1e6)
    assert int(cnt.item()) == 0 and not dets.cpu().numpy().any()
    # single level
    dets, cnt, c_obb, c_hbb, c_score, _ = core.rpn_proposals(_cuda(cls[:1]), _cuda(reg[:1]), _cuda(anchors[:1]), 3, True, 64, 64, 0.5, 0,
                                                              want_candidates=True)
    live = np.isfinite(c_score.cpu().numpy())
    keep = oracle.jt_nms(np.concatenate([c_hbb.cpu().numpy()[live], c_score.cpu().numpy()[live][:, None]], 1), 0.5)[:64]
    assert int(cnt.item()) == len(keep)
    assert np.array_equal(dets[:len(keep)].cpu().numpy(),
                          np.concatenate([c_obb.cpu().numpy()[live], c_score.cpu().numpy()[live][:, None]], 1)[keep])


@pytest.mark.parametrize("thr", [0.3, 0.7, 0.8])
def test_jt_nms_kind_direct(cuda, oracle, thr):
    """rsdet_nms(kind = RSDET_NMS_HBB_P1) is jt.nms: fp32 boxes, '+1' widths, IoU > thr, result in score order."""
    from rs_detection_b200 import core
    from rs_detection_b200._lib import NMS_HBB_P1
    rng = np.random.default_rng(int(thr * 100))
    n = 3000
    c = rng.uniform(0, 600, (n, 2)).astype(np.float32)
    wh = rng.uniform(4, 120, (n, 2)).astype(np.float32)
    hb = np.concatenate([c - wh / 2, c + wh / 2], 1).astype(np.float32)
    hb[:50] = hb[50:100] + rng.normal(0, 1.0, (50, 4)).astype(np.float32)  # near duplicates
    sc = W.distinct_scores(n, 17)
    res = core.nms(NMS_HBB_P1, torch.from_numpy(hb).cuda(), torch.from_numpy(sc).cuda(), thr, want_score=True)
    got = res.score_idx.cpu().numpy()
    want = oracle.jt_nms(np.concatenate([hb, sc[:, None]], 1), thr)
    assert np.array_equal(got, want) and 0 < len(want) < n
