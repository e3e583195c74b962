"""Host-side logic that needs no GPU: class-shard planning, oracle glue semantics, workloads."""
import numpy as np

import workloads as W


def test_plan_class_shards_balances_dominant_classes():
    from rs_detection_b200.merge import plan_class_shards
    counts = [9000, 25000, 52000, 3000, 5000, 2000, 3000, 7000, 3000, 3000]  # FAIR1M-like
    for world in (1, 2, 4, 8):
        owner = plan_class_shards(counts, world)
        assert len(owner) == 10 and set(owner) <= set(range(world))
        assert owner == plan_class_shards(counts, world)  # deterministic
        if world >= 2:
            # the dominant class owns a rank by itself
            big = owner[2]
            assert [k for k, o in enumerate(owner) if o == big] == [2]
    assert plan_class_shards([0, 0, 0], 2) in ([0, 1, 0], [0, 0, 0], [0, 1, 1], [0, 1, 0])


def test_multiclass_glue_semantics(oracle):
    n, C = 300, 5
    bb = W.rotated_boxes(n, 1, canvas=300, smin=16, smax=100)
    sc = W.class_scores(n, C, 1)
    d_all, l_all = oracle.multiclass_nms_rotated(bb, sc, 0.05, dict(iou_thr=0.1), max_num=10 ** 6)
    # per class, the fused call equals independent nms_rotated (label gate == per-class NMS)
    for c in range(C):
        m = sc[:, c + 1] > 0.05
        keep = oracle.nms_rotated(bb[m], sc[m, c + 1], 0.1)
        got = d_all[l_all == c]
        want = np.concatenate([bb[m][keep], sc[m, c + 1][keep][:, None]], 1)
        assert np.array_equal(got[np.argsort(-got[:, 5], kind="stable")], want[np.argsort(-want[:, 5], kind="stable")])
    # reference quirk (nms_rotated.py:590-591): max_num=-1 drops the last detection
    d_q, _ = oracle.multiclass_nms_rotated(bb, sc, 0.05, dict(iou_thr=0.1), max_num=-1)
    assert d_q.shape[0] == d_all.shape[0] - 1
    assert np.all(np.diff(d_all[:, 5]) <= 0)


def test_assign_semantics(oracle):
    ov = np.array([[0.6, 0.2, 0.0, 0.5], [0.7, 0.4, 0.0, 0.5]], np.float32)
    inds, mx, lab = oracle.max_iou_assign(ov, 0.5, 0.5, 0.5, False, gt_labels=np.array([3, 9]))
    assert inds.tolist() == [2, 0, 0, 1] and lab.tolist() == [9, 0, 0, 3]  # argmax ties -> first gt; default fill 0 (assigner.py:52)
    np.testing.assert_allclose(mx, [0.7, 0.4, 0.0, 0.5])
    _, _, lab = oracle.max_iou_assign(ov, 0.5, 0.5, 0.5, False, gt_labels=np.array([3, 9]), assigned_labels_filled=-1)
    assert lab.tolist() == [9, -1, -1, 3]  # the Oriented R-CNN configs pass -1 (oriented_head.py:70)
    inds, _, _ = oracle.max_iou_assign(ov, 0.9, 0.3, 0.3, True)
    assert inds.tolist() == [2, -1, 0, -1]  # low-quality matching: each gt claims its best column, later gts win


def test_workloads_are_seeded_and_shaped():
    a, b = W.proposals(4000, 5), W.proposals(4000, 5)
    assert np.array_equal(a, b) and a.shape == (4000, 6) and a.dtype == np.float32
    assert [s[2] for s in W.fpn_shapes()] == [256, 128, 64, 32]
    s = W.distinct_scores(1000, 0)
    assert len(np.unique(s)) == 1000
    sc = W.merge_scene(num_objects=50, scene=1500, seed=1)
    assert sc["polys"].shape[1] == 8 and sc["polys"].dtype == np.float64 and len(np.unique(sc["scores"])) == sc["scores"].size
    from oracle import oracle as O
    lv = O.map_roi_levels(O.roi_rescale(a, (1.4, 1.2)), 4)
    assert set(lv.tolist()) == {0, 1, 2, 3}
