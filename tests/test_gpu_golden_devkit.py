"""GPU: the product's merge / evaluation mirror (device engine kinds MERGE and HBB_P1_F64, `rsdet_voc_match`) against
the fixtures produced by the reference's OWN result_merge.py, tools/merge_results.py and voc_eval.py
(tests/golden/devkit_golden.npz, generator tests/golden/make_golden_devkit.py).  Result files byte for byte."""
import json
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
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


def test_device_nms_vs_reference(cuda, g):
    from rs_detection_b200.jdet.data.devkits import result_merge as RM
    d = g["nms_dets"]
    for thr in (0.1, 0.3):
        assert np.array_equal(np.asarray(RM.py_cpu_nms_poly_fast(d, thr)), g["nms_keep_fast_%g" % thr])
        assert np.array_equal(np.asarray(RM.py_cpu_nms_poly(d, thr)), g["nms_keep_slow_%g" % thr])
    assert np.array_equal(np.asarray(RM.py_cpu_nms(g["hbb_dets"], 0.3)), g["hbb_keep_0.3"])
    byname = {"a": d[:50].tolist(), "b": d[50:120].tolist()}
    kept = RM.nmsbynamedict(byname, RM.py_cpu_nms_poly_fast, 0.1)
    assert [len(kept["a"]), len(kept["b"])] == g["nmsbynamedict_counts"].tolist()


@pytest.mark.parametrize("ttype", [0, 1])
def test_mergebypoly_files_vs_reference(cuda, tmp_path, g, ttype):
    from rs_detection_b200.jdet.data.devkits import result_merge as RM
    _write(tmp_path / "src", json.loads(str(g["before_nms"])))
    RM.mergebypoly(str(tmp_path / "src"), str(tmp_path / "dst"), nms_threshold_type=ttype)      # all files, one launch
    assert _read(tmp_path / "dst") == json.loads(str(g["after_nms_type%d" % ttype]))
    RM.mergesingle(str(tmp_path / "one"), RM.py_cpu_nms_poly_fast, str(tmp_path / "src" / "Ship.txt"), nms_threshold_type=ttype)
    assert open(tmp_path / "one" / "Ship.txt").read() == json.loads(str(g["after_nms_type%d" % ttype]))["Ship.txt"]


def test_mergebyrec_and_tool_vs_reference(cuda, tmp_path, g):
    from rs_detection_b200.jdet.data.devkits import result_merge as RM
    from rs_detection_b200.jdet.tools import merge_results as MR
    _write(tmp_path / "hsrc", json.loads(str(g["before_hbb"])))
    RM.mergebyrec(str(tmp_path / "hsrc"), str(tmp_path / "hdst"))
    assert _read(tmp_path / "hdst") == json.loads(str(g["after_hbb"]))
    _write(tmp_path / "after", json.loads(str(g["after_nms_type0"])))
    MR.merge_files(str(tmp_path / "after"), str(tmp_path / "tool"), nms_thr=0.05, process_num=1)
    assert _read(tmp_path / "tool") == json.loads(str(g["tool_merge_files_thr005"]))


@pytest.mark.parametrize("thr", [0.5, 0.3])
def test_voc_eval_vs_reference(cuda, g, thr):
    from rs_detection_b200.jdet.data.devkits.voc_eval import voc_eval_dota
    gts = {}
    for k in np.unique(g["voc_gt_img"]).tolist() + [6]:
        m = g["voc_gt_img"] == k
        gts[int(k)] = {"box": g["voc_gt_boxes"][m], "difficult": g["voc_gt_difficult"][m]}
    rec, prec, ap = voc_eval_dota(g["voc_dets"], gts, None, thr, False)
    assert np.array_equal(rec, g["voc_%g_area_rec" % thr]) and np.array_equal(prec, g["voc_%g_area_prec" % thr])
    assert ap == float(g["voc_%g_area_ap" % thr])
    assert voc_eval_dota(g["voc_dets"], gts, None, thr, True)[2] == float(g["voc_%g_07_ap" % thr])


def test_data_merge_vs_reference(cuda, tmp_path, g):
    """data_merge.data_merge (data_merge.py:50-54) end to end: result list -> before_nms -> after_nms, against the files
    the reference's own data_merge wrote."""
    import pickle
    from rs_detection_b200.jdet.data.devkits.data_merge import data_merge
    res = pickle.loads(g["results_pkl"].tobytes())
    data_merge(res, str(tmp_path / "before"), str(tmp_path / "after"), "FAIR1M_1_5")
    assert _read(tmp_path / "before") == json.loads(str(g["dm_before_nms"]))
    assert _read(tmp_path / "after") == json.loads(str(g["dm_after_nms"]))
