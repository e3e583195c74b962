"""Merge-stage polygon IoU: the oracle's float64 clipper against EXACT rational
arithmetic (tests/exact_geometry.py).  Shapely/GEOS -- what the reference calls -- is not available; agreement with the
exact value to 1e-10 (measured: 9.4e-12 at scene coordinates of 2.6e3, i.e. a handful of float64 ulps of the
un-shifted shoelace products, the regime every float64 clipper incl. GEOS works in) bounds the distance to ANY
float64 implementation, and the greedy keep lists are shown to be decided by margins >= 1e-6, four orders of magnitude
above that."""
from fractions import Fraction

import numpy as np
import pytest

import workloads as W
from exact_geometry import areas_exact, greedy_merge_nms_exact, iou_poly_exact, is_convex


def _scene(seed, n_obj):
    sc = W.merge_scene(num_objects=n_obj, num_classes=3, scene=2600, seed=seed)
    return sc["polys"], sc["scores"], sc["labels"]


def test_iou_poly_matches_exact_rational(oracle):
    polys, _, _ = _scene(3, 260)
    x1, y1 = polys[:, 0::2].min(1), polys[:, 1::2].min(1)
    x2, y2 = polys[:, 0::2].max(1), polys[:, 1::2].max(1)
    rng = np.random.default_rng(0)
    n = polys.shape[0]
    worst, checked, overlapping = 0.0, 0, 0
    cand = np.argwhere((np.minimum(x2[:, None], x2[None]) > np.maximum(x1[:, None], x1[None])) &
                       (np.minimum(y2[:, None], y2[None]) > np.maximum(y1[:, None], y1[None])) &
                       (np.arange(n)[:, None] < np.arange(n)[None]))
    assert len(cand) > 500
    for a, b in cand[rng.permutation(len(cand))[:1500]]:
        if not (is_convex(polys[a]) and is_convex(polys[b])):
            continue
        ex = iou_poly_exact(polys[a], polys[b])
        got = oracle.iou_poly(polys[a], polys[b])
        worst = max(worst, abs(got - float(ex)))
        checked += 1
        overlapping += ex > 0
    print(f"{checked} hbb-overlapping pairs ({overlapping} with a real intersection): max |float64 - exact| = {worst:.3g}")
    assert checked > 400 and overlapping > 200 and worst <= 1e-10


def test_areas_exact_sanity():
    sq = [0, 0, 2, 0, 2, 2, 0, 2]
    a, b, i = areas_exact(sq, [1, 1, 3, 1, 3, 3, 1, 3])
    assert (a, b, i) == (4, 4, 1)
    assert iou_poly_exact(sq, [2, 0, 4, 0, 4, 2, 2, 2]) == 0           # shared edge
    assert iou_poly_exact(sq, sq[::-1][1::2] + sq[::-1][0::2]) >= 0     # orientation does not matter
    a, b, i = areas_exact(sq, [1, -0.5, 2.5, 1, 1, 2.5, -0.5, 1])       # diamond |x-1|+|y-1| <= 1.5 over the square
    assert (a, b, i) == (4, Fraction(9, 2), Fraction(7, 2))             # four corner triangles of area 1/8 are cut off


@pytest.mark.parametrize("seed,thr", [(5, 0.1), (6, 0.3), (7, 0.5)])
def test_greedy_keep_lists_match_exact(oracle, seed, thr):
    """py_cpu_nms_poly_fast with the float64 clipper == the same greedy loop with exact IoUs; the decision closest
    to the threshold is reported (the contract excludes pairs within 1e-6 of it; none comes near)."""
    polys, scores, labels = _scene(seed, 160)
    total = 0
    for c in np.unique(labels):
        m = labels == c
        dets = np.concatenate([polys[m], scores[m][:, None]], 1)
        if not all(is_convex(p) for p in dets[:, :8]):
            continue
        order = oracle.score_order(dets[:, 8])
        want, margin = greedy_merge_nms_exact(dets, thr, order)
        got = oracle.py_cpu_nms_poly_fast(dets, thr)
        print(f"class {c}: {len(dets)} detections -> {len(want)} kept, closest decision margin {margin}")
        assert got == want
        assert margin is None or margin > 1e-6
        total += len(dets)
    assert total > 300
