"""CPU: the RPN proposal-stage oracle against the reference's own example and
hand-derivable cases."""
import numpy as np

from oracle import oracle as O


def test_anchor_grid_matches_reference_docstring():
    # python/jdet/models/boxes/anchor_generator.py:122-136 (the class docstring's expected output)
    a = O.anchor_grid((2, 2), 16, base_size=9, scales=(1.,), ratios=(1.,))
    assert np.array_equal(a, np.array([[-4.5, -4.5, 4.5, 4.5], [11.5, -4.5, 20.5, 4.5], [-4.5, 11.5, 4.5, 20.5],
                                       [11.5, 11.5, 20.5, 20.5]], np.float32))
    b = O.anchor_grid((1, 1), 32, base_size=18, scales=(1.,), ratios=(1.,))
    assert np.array_equal(b, np.array([[-9., -9., 9., 9.]], np.float32))
    from rs_detection_b200.jdet.models.boxes.anchor_generator import AnchorGenerator
    g = AnchorGenerator([16, 32], [1.], [1.], [9, 18])
    got = g.grid_anchors([(2, 2), (1, 1)], device="cpu")
    assert np.array_equal(got[0].numpy(), a) and np.array_equal(got[1].numpy(), b)
    g3 = AnchorGenerator(strides=[4, 8], ratios=[0.5, 1.0, 2.0], scales=[8])
    assert g3.num_base_anchors == [3, 3]
    assert np.array_equal(g3.grid_anchors([(5, 7), (3, 2)], device="cpu")[0].numpy(), O.anchor_grid((5, 7), 4))


def test_midpoint_decode_zero_deltas_is_the_anchor():
    anchors = np.array([[10, 20, 50, 40], [0, 0, 16, 64]], np.float32)
    obb = O.midpoint_offset_decode(anchors, np.zeros((2, 6), np.float32))
    # zero deltas: the quadrilateral's vertices are the edge midpoints -> a diamond with equal diagonals
    # after the stretch; rectpoly2obb measures it along its first edge
    assert np.allclose(obb[:, :2], [[30, 30], [8, 32]])
    assert np.all(obb[:, 2] >= obb[:, 3]) and np.all((obb[:, 4] >= -np.pi / 2) & (obb[:, 4] < np.pi / 2))
    # da = +-0.5 (clamped) moves the top midpoint to the corner: the box becomes the anchor itself
    obb = O.midpoint_offset_decode(anchors[:1], np.array([[0, 0, 0, 0, 9.0, 9.0]], np.float32))
    assert np.allclose(obb[0, :4], [30, 30, 40, 20], atol=1e-4) and abs(obb[0, 4]) < 1e-6


def test_jt_nms_by_hand():
    d = np.array([[0, 0, 9, 9, 0.9], [1, 1, 10, 10, 0.8], [100, 100, 109, 109, 0.7], [0, 0, 9, 9, 0.95]], np.float32)
    # boxes 0 and 3 coincide (IoU 1); box 1 overlaps them with IoU 81/119 = 0.68
    assert O.jt_nms(d, 0.7).tolist() == [3, 1, 2]
    assert O.jt_nms(d, 0.5).tolist() == [3, 2]
    assert O.jt_nms(np.zeros((0, 5), np.float32), 0.5).tolist() == []


def test_level_offsets_separate_levels():
    p = np.array([[50, 50, 20, 10, 0.0], [50, 50, 20, 10, 0.0], [50, 50, 20, 10, 0.0]], np.float32)
    s = np.array([0.9, 0.8, 0.7], np.float32)
    dets, keep, _ = O.rpn_level_offset_nms(p, s, np.array([0, 0, 1]), 0.8, 10)
    assert keep.tolist() == [0, 2] and dets.shape == (2, 6)
    assert O.rpn_level_offset_nms(p, s, np.array([0, 0, 0]), 0.8, 10)[1].tolist() == [0]
    assert O.rpn_level_offset_nms(p, s, np.array([0, 1, 2]), 0.8, 2)[1].tolist() == [0, 1]


def test_rpn_golden_drift_guard():
    import os
    import workloads as W
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "rpn_golden.npz"))
    shapes, strides = ((48, 48), (24, 24), (12, 12), (6, 6), (3, 3)), (4, 8, 16, 32, 64)
    cls, reg = W.rpn_outputs(shapes, 3, 21)
    anchors = [O.anchor_grid(s, st) for s, st in zip(shapes, strides)]
    assert np.array_equal(anchors[2], g["anchors_l2"])
    p, s, ids, rows = O.rpn_candidates(cls, reg, anchors, True, 600, 0)
    assert np.array_equal(rows, g["cand_rows"]) and np.array_equal(ids, g["cand_level"])
    assert np.allclose(p, g["cand_obb"], rtol=1e-5, atol=1e-3) and np.allclose(s, g["cand_score"], atol=1e-6)
    dets, keep, _ = O.rpn_level_offset_nms(g["cand_obb"], g["cand_score"], g["cand_level"].astype(np.int64), 0.8, 400)
    assert np.array_equal(keep, g["keep"]) and np.array_equal(dets, g["dets"])
