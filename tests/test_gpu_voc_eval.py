"""voc_eval_dota matching on the device vs the python restatement of the reference loop."""
import numpy as np
import pytest

import workloads as W

pytestmark = pytest.mark.gpu


def _scene(seed, nimg=12, ngt=60, ndet_per_gt=3):
    rng = np.random.default_rng(seed)
    gts, dets = {}, []
    for im in range(nimg):
        g = W.rotated_boxes(ngt, seed * 100 + im, canvas=800, smin=12, smax=90, dtype=np.float64)
        gts[im] = {"box": W.obb_to_poly64(g), "difficult": rng.uniform(size=ngt) < 0.15}
        for k in range(ndet_per_gt):
            j = g.copy()
            j[:, :2] += rng.normal(0, 0.12, (ngt, 2)) * np.sqrt(j[:, 2:3] * j[:, 3:4])
            j[:, 2:4] *= np.exp(rng.normal(0, 0.15, (ngt, 2)))
            j[:, 4] += rng.normal(0, 0.1, ngt)
            p = W.obb_to_poly64(j)
            dets.append(np.concatenate([np.full((ngt, 1), im), p, rng.uniform(0.01, 1, (ngt, 1))], 1))
        fa = W.obb_to_poly64(W.rotated_boxes(40, seed * 100 + im + 50, canvas=800, smin=12, smax=90, dtype=np.float64))
        dets.append(np.concatenate([np.full((40, 1), im), fa, rng.uniform(0.01, 0.6, (40, 1))], 1))
    dets.append(np.concatenate([np.full((5, 1), 999), fa[:5], rng.uniform(0.5, 1, (5, 1))], 1))  # image without gts
    gts[7] = {"box": np.zeros((0, 8)), "difficult": np.zeros((0,), bool)}                          # image with no gt of this class
    d = np.concatenate(dets)
    d[:, -1] += np.arange(len(d)) * 1e-9  # distinct confidences
    return d, gts


@pytest.mark.parametrize("thr", [0.5, 0.3])
def test_voc_eval_matches_reference_loop(cuda, oracle, thr):
    from rs_detection_b200.jdet.data.devkits.voc_eval import voc_eval_dota, voc_match
    dets, gts = _scene(3)
    tp, fp, npos = voc_match(dets, gts, thr)
    gts_o = {k: v for k, v in gts.items()}
    gts_o.setdefault(999, {"box": np.zeros((0, 8)), "difficult": np.zeros((0,), bool)})
    wtp, wfp = oracle.voc_match(dets, gts_o, thr)
    assert np.array_equal(tp, wtp) and np.array_equal(fp, wfp)
    assert 100 < tp.sum() < len(dets) and fp.sum() > 100
    rec, prec, ap = voc_eval_dota(dets, gts, None, thr)
    assert 0.05 < ap < 1.0 and rec.shape == (len(dets),) and np.all(np.diff(rec) >= 0)
    ap07 = voc_eval_dota(dets, gts, None, thr, use_07_metric=True)[2]
    assert abs(ap07 - ap) < 0.1
    assert voc_eval_dota(np.zeros((0, 10)), gts, None) == (0., 0., 0.)
