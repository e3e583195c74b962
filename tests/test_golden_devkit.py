"""CPU: the oracle's merge / evaluation restatements against fixtures produced by RUNNING the reference's own
result_merge.py, tools/merge_results.py and voc_eval.py (tests/golden/devkit_golden.npz; generator
tests/golden/make_golden_devkit.py: jittor and shapely are played by shims, GEOS by exact rational arithmetic).
Rows a11-a13 and f3 of SURVEY section 8: byte-identical result files, identical keep lists, identical rec / prec / ap."""
import json
import os

import numpy as np
import pytest

from oracle import formats as F

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "devkit_golden.npz")


@pytest.fixture(scope="module")
def g():
    return dict(np.load(GOLD))


def _write(tmp, files):
    os.makedirs(tmp, exist_ok=True)
    for name, txt in files.items():
        with open(os.path.join(tmp, name), "w") as f:
            f.write(txt)


def _read(d):
    return {f: open(os.path.join(d, f)).read() for f in sorted(os.listdir(d))}


def test_fixture_is_decisive(g):
    # every polygon decision of the array-level fixtures lies far outside the 1e-6 band around its threshold
    assert float(g["nms_margin_0.1"]) > 1e-4 and float(g["nms_margin_0.3"]) > 1e-4
    assert np.array_equal(g["nms_keep_fast_0.1"], g["nms_keep_slow_0.1"])      # the reference's two variants agree
    before, after = json.loads(str(g["before_nms"])), json.loads(str(g["after_nms_type0"]))
    assert sum(len(v.strip().split("\n")) for v in after.values()) < 0.4 * sum(len(v.strip().split("\n")) for v in before.values())


def test_oracle_nms_vs_reference(oracle, g):
    d = g["nms_dets"]
    for thr in (0.1, 0.3):
        assert np.array_equal(np.asarray(oracle.py_cpu_nms_poly_fast(d, thr)), g["nms_keep_fast_%g" % thr])
    assert np.array_equal(np.asarray(oracle.py_cpu_nms(g["hbb_dets"], 0.3)), g["hbb_keep_0.3"])
    assert np.array_equal(np.asarray(oracle.poly2origpoly([10.0, 20.5, 30.25, 40.0, 50.0, 60.0, 70.0, 80.0], 824, 1648, "0.5")),
                          g["poly2origpoly"])


@pytest.mark.parametrize("ttype", [0, 1])
def test_oracle_mergesingle_vs_reference(tmp_path, g, ttype):
    _write(tmp_path / "src", json.loads(str(g["before_nms"])))
    for f in os.listdir(tmp_path / "src"):
        F.mergesingle(tmp_path / "dst", str(tmp_path / "src" / f), nms_threshold_type=ttype)
    assert _read(tmp_path / "dst") == json.loads(str(g["after_nms_type%d" % ttype]))


def test_oracle_mergebyrec_and_tool_vs_reference(tmp_path, g):
    _write(tmp_path / "hsrc", json.loads(str(g["before_hbb"])))
    for f in os.listdir(tmp_path / "hsrc"):
        F.mergesingle(tmp_path / "hdst", str(tmp_path / "hsrc" / f), nms="rec")
    assert _read(tmp_path / "hdst") == json.loads(str(g["after_hbb"]))
    _write(tmp_path / "after", json.loads(str(g["after_nms_type0"])))
    for f in os.listdir(tmp_path / "after"):
        F.merge_file(str(tmp_path / "after" / f), str(tmp_path / "tool"), nms_thr=0.05)
    assert _read(tmp_path / "tool") == json.loads(str(g["tool_merge_files_thr005"]))


def _gts(g):
    out = {}
    for k in np.unique(g["voc_gt_img"]).tolist() + [6]:
        m = g["voc_gt_img"] == k
        out[int(k)] = {"box": g["voc_gt_boxes"][m], "difficult": g["voc_gt_difficult"][m]}
    return out


@pytest.mark.parametrize("thr", [0.5, 0.3])
def test_oracle_voc_eval_vs_reference(oracle, g, thr):
    """voc_eval_dota (voc_eval.py:236-318): TP / FP vectors of the restated loop reproduce the reference's recall and
    precision curves exactly, and voc_ap both metrics."""
    from rs_detection_b200.jdet.data.devkits.voc_eval import voc_ap     # pure numpy, no device work
    dets, gts = g["voc_dets"], _gts(g)
    tp, fp = oracle.voc_match(dets, gts, thr)
    npos = int(sum((~v["difficult"]).sum() for v in gts.values()))
    ctp, cfp = np.cumsum(tp), np.cumsum(fp)
    rec, prec = ctp / float(npos), ctp / np.maximum(ctp + cfp, np.finfo(np.float64).eps)
    assert np.array_equal(rec, g["voc_%g_area_rec" % thr]) and np.array_equal(prec, g["voc_%g_area_prec" % thr])
    assert voc_ap(rec, prec, False) == float(g["voc_%g_area_ap" % thr])
    assert voc_ap(rec, prec, True) == float(g["voc_%g_07_ap" % thr])


def test_before_nms_writer_vs_reference(tmp_path, g):
    """data_merge.prepare_data (data_merge.py:29-48) was RUN on a pickled result list (incl. H and HV flipped tiles):
    the oracle's writer and the product's host-side writer reproduce its files byte for byte."""
    import pickle
    import workloads as W
    from rs_detection_b200.jdet.data.devkits.data_merge import get_classes_by_name, prepare_data
    res = pickle.loads(g["results_pkl"].tobytes())
    classes = json.loads(str(g["dm_classes"]))
    assert classes == list(W.FAIR1M_CLASSES) == list(get_classes_by_name("FAIR1M_1_5"))
    F.write_before_nms(res, tmp_path / "o", classes)
    assert _read(tmp_path / "o") == json.loads(str(g["dm_before_nms"]))
    prepare_data(res, tmp_path / "p", classes)
    assert _read(tmp_path / "p") == json.loads(str(g["dm_before_nms"]))
    # and the merge of those files (the reference ran mergebypoly through its 16-process pool)
    for f in os.listdir(tmp_path / "o"):
        F.mergesingle(tmp_path / "m", str(tmp_path / "o" / f))
    assert _read(tmp_path / "m") == json.loads(str(g["dm_after_nms"]))
