"""GPU parity: merge-stage NMS (py_cpu_nms_poly_fast), merge.py hbb nms, box transforms."""
import numpy as np
import pytest
import torch

import workloads as W

pytestmark = pytest.mark.gpu


def _t(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def test_transforms_vs_oracle(cuda, oracle):
    from rs_detection_b200.jdet.ops.bbox_transforms import obb2hbb, obb2poly, poly2hbb
    o = W.rotated_boxes(5000, 1)
    p = obb2poly(_t(o)).cpu().numpy()
    np.testing.assert_allclose(p, oracle.obb2poly(o), rtol=1e-6, atol=1e-4)
    np.testing.assert_allclose(obb2hbb(_t(o)).cpu().numpy(), oracle.obb2hbb(o), rtol=1e-6, atol=1e-4)
    assert np.array_equal(poly2hbb(_t(p)).cpu().numpy(), oracle.poly2hbb(p))
    assert tuple(obb2poly(_t(o.reshape(50, 100, 5))).shape) == (50, 100, 8)
    assert isinstance(obb2poly(o), np.ndarray)


@pytest.mark.parametrize("thr", [0.1, 0.3, 0.0001])
def test_py_cpu_nms_poly_fast_vs_oracle(cuda, oracle, thr):
    from rs_detection_b200.jdet.data.devkits.result_merge import py_cpu_nms_poly_fast
    sc = W.merge_scene(num_objects=400, scene=3000, seed=2)
    dets = np.concatenate([sc["polys"], sc["scores"][:, None]], 1)
    got = py_cpu_nms_poly_fast(dets, thr)
    want = oracle.py_cpu_nms_poly_fast(dets, thr)
    assert isinstance(got, list) and got == want
    assert py_cpu_nms_poly_fast(np.zeros((0, 9)), thr) == []


def test_iou_poly_pairs_bit_exact(cuda, oracle):
    from rs_detection_b200 import core
    from rs_detection_b200.jdet.ops.nms_poly import iou_poly
    sc = W.merge_scene(num_objects=300, scene=2000, seed=3)
    p = sc["polys"]
    q = np.roll(p, 1, axis=0) + np.random.default_rng(0).normal(0, 2.0, p.shape)
    q[::3] = p[::3] + 0.5
    got = core.iou_poly_pairs(_t(p), _t(q)).cpu().numpy()
    want = np.array([oracle.iou_poly(a, b) for a, b in zip(p, q)])
    assert (want > 0.3).sum() > 50
    assert np.array_equal(got, want)
    assert iou_poly(p[0], q[0]) == want[0]


def test_batched_merge_per_class_thresholds(cuda, oracle):
    """All (scene, class) groups in ONE launch with the per-class thresholds of result_merge.py:26-27 ==
    one reference call per class file."""
    from rs_detection_b200.jdet.data.devkits.result_merge import merge_detections, nms_threshold_1
    sc = W.merge_scene(num_objects=1500, scene=4000, seed=4)
    thr = np.array([nms_threshold_1[c] for c in W.FAIR1M_CLASSES])
    got = merge_detections(sc["polys"], sc["scores"], sc["labels"], group_thresh=thr)
    want = []
    for c in range(10):
        idx = np.nonzero(sc["labels"] == c)[0]
        dets = np.concatenate([sc["polys"][idx], sc["scores"][idx, None]], 1)
        want += [idx[k] for k in oracle.py_cpu_nms_poly_fast(dets, thr[c])]
    assert sorted(got.tolist()) == sorted(want)
    s = sc["scores"][got]
    assert np.all(np.diff(s) < 0)  # descending score order
    assert 0 < len(want) < sc["scores"].size


@pytest.mark.parametrize("objects,scene,jitter", [(4000, 6000, 1.0), (2500, 1500, 3.0)])
def test_sparse_merge_path_vs_oracle(cuda, oracle, objects, scene, jitter):
    """>= 8192 detections take the sparse path (sweep-and-prune candidates -> CSR of suppressing pairs -> fixed-point
    greedy); the second case is a crowded scene (long kept/dead dependency chains, many rounds)."""
    from rs_detection_b200.jdet.data.devkits.result_merge import merge_detections, nms_threshold_1
    sc = W.merge_scene(num_objects=objects, scene=scene, seed=11, jitter_px=jitter)
    assert sc["scores"].size >= 8192
    thr = np.array([nms_threshold_1[c] for c in W.FAIR1M_CLASSES])
    got = merge_detections(sc["polys"], sc["scores"], sc["labels"], group_thresh=thr)
    want = []
    for c in range(10):
        idx = np.nonzero(sc["labels"] == c)[0]
        dets = np.concatenate([sc["polys"][idx], sc["scores"][idx, None]], 1)
        want += [idx[k] for k in oracle.py_cpu_nms_poly_fast(dets, thr[c])]
    assert sorted(got.tolist()) == sorted(want)
    assert np.all(np.diff(sc["scores"][got]) < 0) and 0 < len(want) < sc["scores"].size


def test_sparse_merge_overflow_falls_back_to_dense(cuda, oracle):
    """85 clusters of 100 near-identical quads: ~100 candidate pairs per detection, more than the sparse path's
    32 per detection, so its gate opens and the dense kernels produce the result."""
    from rs_detection_b200.jdet.data.devkits.result_merge import py_cpu_nms_poly_fast
    rng = np.random.default_rng(21)
    centres = rng.uniform(200, 5000, (85, 2))
    obb = np.zeros((8500, 5))
    obb[:, :2] = np.repeat(centres, 100, 0) + rng.normal(0, 1.5, (8500, 2))
    obb[:, 2] = rng.uniform(40, 60, 8500)
    obb[:, 3] = rng.uniform(15, 25, 8500)
    obb[:, 4] = np.repeat(rng.uniform(-1.5, 1.5, 85), 100) + rng.normal(0, 0.05, 8500)
    polys = np.round(W.obb_to_poly64(obb), 4)
    scores = rng.permutation(8500) / 8500.0 + 1e-4
    dets = np.concatenate([polys, scores[:, None]], 1)
    got = py_cpu_nms_poly_fast(dets, 0.3)
    assert got == oracle.py_cpu_nms_poly_fast(dets, 0.3) and 85 <= len(got) < 2000


def test_merge_py_hbb_nms(cuda, oracle):
    from rs_detection_b200.jdet.merge import nms
    rng = np.random.default_rng(8)
    n = 3000
    xy = rng.uniform(0, 1000, (n, 2))
    wh = rng.uniform(10, 80, (n, 2))
    b = np.concatenate([xy, xy + wh, W.distinct_scores(n, 8).astype(np.float64)[:, None]], 1)
    for thr in (0.625, 0.3):
        assert np.array_equal(nms(b, thr), oracle.hbb_nms(b, thr))


def test_merge_sharded_single_rank_matches_unsharded(cuda, oracle):
    """rs_detection_b200.merge.merge_sharded on one rank (no process group): same survivors as the single-launch
    engine call, in the canonical (class, scene, score) order; no host synchronisation before `.indices()`."""
    from rs_detection_b200.jdet.data.devkits.result_merge import merge_detections, nms_threshold_1
    from rs_detection_b200.merge import merge_sharded
    sc = W.merge_scene(num_objects=900, scene=4000, seed=12)
    thr = [nms_threshold_1[c] for c in W.FAIR1M_CLASSES]
    p, s, l = [torch.from_numpy(np.ascontiguousarray(x)).cuda() for x in (sc["polys"], sc["scores"], sc["labels"])]
    counts = np.bincount(sc["labels"], minlength=10).tolist()
    res = merge_sharded(p, s, l, class_thr=thr, class_counts=counts)
    kept = res.indices().cpu().numpy()
    want = merge_detections(p, s, l, group_thresh=thr).cpu().numpy()
    assert sorted(kept.tolist()) == sorted(want.tolist()) and 0 < kept.size < s.numel()
    lab, scr = sc["labels"][kept], sc["scores"][kept]
    assert np.all(np.diff(lab) >= 0)
    assert all(np.all(np.diff(scr[lab == c]) <= 0) for c in range(10))
    # against the oracle, class by class
    for c in (1, 2):
        idx = np.nonzero(sc["labels"] == c)[0]
        dets = np.concatenate([sc["polys"][idx], sc["scores"][idx, None]], 1)
        assert [int(idx[k]) for k in oracle.py_cpu_nms_poly_fast(dets, thr[c])] == kept[lab == c].tolist()


def test_merge_more_rows_than_one_engine_call(cuda, oracle):
    """a `before_nms` dump with more than 2^18 rows.  core.nms_grouped splits the (file, scene) groups over
    several engine calls; every group must come out exactly as one reference-style call per group gives it."""
    from rs_detection_b200 import core
    from rs_detection_b200._lib import NMS_MERGE
    rng = np.random.default_rng(5)
    G, per = 2750, 100                                    # 275 000 rows > 2^18
    base = W.obb_to_poly64(W.rotated_boxes(per, 3, canvas=300, smin=8, smax=40, dtype=np.float64))
    polys = np.round(np.concatenate([base + rng.normal(0, 0.6, base.shape) for _ in range(G)]), 4)
    gids = np.repeat(np.arange(G), per)
    perm = rng.permutation(G * per)                        # groups interleaved like tiles in a file
    polys, gids = polys[perm], gids[perm]
    scores = rng.permutation(G * per).astype(np.float64) / (G * per)
    assert polys.shape[0] > core.MAX_NMS_ROWS
    kept = core.nms_grouped(NMS_MERGE, torch.from_numpy(polys).cuda(), torch.from_numpy(scores).cuda(), gids, 0.3)
    assert len(set(kept.tolist())) == kept.size
    for gsel in rng.choice(G, 40, replace=False):
        idx = np.nonzero(gids == gsel)[0]
        want = [int(idx[k]) for k in oracle.py_cpu_nms_poly_fast(np.concatenate([polys[idx], scores[idx, None]], 1), 0.3)]
        got = kept[gids[kept] == gsel].tolist()
        assert got == want
    with pytest.raises(RuntimeError, match="at most"):
        core.nms_grouped(NMS_MERGE, torch.from_numpy(polys).cuda(), torch.from_numpy(scores).cuda(), np.zeros(G * per, np.int64), 0.3)
