"""GPU parity, file level: before_nms text -> device merge -> after_nms text, byte for byte against the
oracle's restatement of mergesingle / merge_file / the merge.py ensemble."""
import os

import numpy as np
import pytest

import workloads as W
from oracle import formats as F

pytestmark = pytest.mark.gpu


def _same_dir(a, b):
    assert sorted(os.listdir(a)) == sorted(os.listdir(b)) and len(os.listdir(a)) > 0
    for f in os.listdir(a):
        assert open(os.path.join(a, f)).read() == open(os.path.join(b, f)).read(), f


@pytest.mark.parametrize("threshold_type", [0, 1])
def test_mergebypoly_files_byte_exact(cuda, tmp_path, threshold_type):
    from rs_detection_b200.jdet.data.devkits.data_merge import data_merge
    res = W.tile_results(300, 10, 3000, 2, seed=3)
    data_merge(res, str(tmp_path / "before"), str(tmp_path / "after"), "FAIR1M_1_5", nms_threshold_type=threshold_type)
    for f in os.listdir(tmp_path / "before"):
        F.mergesingle(str(tmp_path / "want"), str(tmp_path / "before" / f), threshold_type)
    _same_dir(tmp_path / "want", tmp_path / "after")
    n_in = sum(len(open(tmp_path / "before" / f).readlines()) for f in os.listdir(tmp_path / "before"))
    n_out = sum(len(open(tmp_path / "after" / f).readlines()) for f in os.listdir(tmp_path / "after"))
    assert 0 < n_out < n_in


def test_mergesingle_and_empty_file(cuda, tmp_path):
    from rs_detection_b200.jdet.data.devkits.result_merge import mergesingle, py_cpu_nms_poly_fast
    res = W.tile_results(100, 2, 2200, 1, seed=4)
    F.write_before_nms(res, tmp_path / "before", W.FAIR1M_CLASSES)
    src = str(tmp_path / "before" / "Ship.txt")
    mergesingle(str(tmp_path / "got"), py_cpu_nms_poly_fast, src)
    F.mergesingle(str(tmp_path / "want"), src)
    _same_dir(tmp_path / "want", tmp_path / "got")
    (tmp_path / "before" / "Bridge.txt").write_text("")
    mergesingle(str(tmp_path / "got"), py_cpu_nms_poly_fast, str(tmp_path / "before" / "Bridge.txt"))
    assert open(tmp_path / "got" / "Bridge.txt").read() == ""
    with pytest.raises(ValueError):
        mergesingle(str(tmp_path / "got"), max, src)


def test_merge_results_tool_byte_exact(cuda, tmp_path):
    from rs_detection_b200.jdet.data.devkits.data_merge import data_merge
    from rs_detection_b200.jdet.tools.merge_results import merge_files, merge_src_files
    for run, seed in (("a", 5), ("b", 6)):
        data_merge(W.tile_results(150, 4, 2200, 2, seed=seed), str(tmp_path / run / "before"), str(tmp_path / run / "Task1"),
                   "FAIR1M_1_5")
        for f in os.listdir(tmp_path / run / "Task1"):
            os.rename(tmp_path / run / "Task1" / f, tmp_path / run / "Task1" / ("Task1_" + f))
    merge_src_files([str(tmp_path / "a" / "Task1"), str(tmp_path / "b" / "Task1")], str(tmp_path / "cat"))
    merge_files(str(tmp_path / "cat"), str(tmp_path / "got"), nms_thr=0.1)
    for f in os.listdir(tmp_path / "cat"):
        assert not f.startswith("Task1_")
        # the two runs share four-decimal score values: ties -> the well-defined rule argsort(kind='stable')[::-1]
        # (numpy's default argsort order of equal scores is an artefact of the numpy build, oracle.score_order)
        F.merge_file(str(tmp_path / "cat" / f), str(tmp_path / "want"), 0.1, stable_ties=True)
    _same_dir(tmp_path / "want", tmp_path / "got")


@pytest.mark.parametrize("thresh", [0.625, {c: 0.3 + 0.05 * i for i, c in enumerate(W.FAIR1M_CLASSES)}])
def test_csv_ensemble_vs_oracle(cuda, tmp_path, thresh):
    from rs_detection_b200.jdet import merge as M
    subs = []
    for seed in (7, 8):
        rng = np.random.default_rng(seed)
        rows = []
        for img in (3, 11, 12):
            o = W.rotated_boxes(120, 40 + img, canvas=1000, smin=16.0, smax=128.0, dtype=np.float64)  # same objects in both runs
            o[:, :2] += rng.normal(0, 2.0, (120, 2))
            p = np.round(W.obb_to_poly64(o), 4)
            sc = (rng.permutation(9000)[:120] + 1 + seed * 0.5) / 10000.0
            cls = np.random.default_rng(img).integers(1, 11, 120)
            rows.append(np.concatenate([np.full((120, 1), img), p, sc[:, None], cls[:, None]], 1))
        subs.append(np.concatenate(rows))
    subs[1] = np.concatenate([subs[1], subs[1][:5] * [99, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1]])  # image absent from run 0: ignored
    got = M.merge_csv_with_class(subs, thresh)
    want = F.ensemble_with_class(subs, thresh)
    assert got.shape == want.shape and np.array_equal(got, want)
    assert 0 < got.shape[0] < subs[0].shape[0] + subs[1].shape[0]
    if not isinstance(thresh, dict):
        got2, want2 = M.merge_csv_without_class(subs, 0.9), F.ensemble_without_class(subs, 0.9)
        assert np.array_equal(got2, want2)
        M.save_to_csv(got, tmp_path / "m.csv")
        back = M.read_csv_to_numpy(tmp_path / "m.csv")
        assert back.shape == got.shape and np.allclose(back, got, atol=5e-5)


def test_py_cpu_nms_and_mergebyrec(cuda, oracle, tmp_path):
    """Horizontal task: py_cpu_nms (float64, '+1' areas, keeps ovr <= thresh) and the file-level mergebyrec."""
    from rs_detection_b200.jdet.data.devkits import result_merge as RM
    rng = np.random.default_rng(3)
    n = 1500
    c = rng.uniform(0, 800, (n, 2))
    wh = rng.uniform(10, 150, (n, 2))
    d = np.concatenate([c - wh / 2, c + wh / 2, W.distinct_scores(n, 5).astype(np.float64)[:, None]], 1)
    d[:40, :4] = d[40:80, :4] + rng.normal(0, 1.0, (40, 4))
    for thr in (0.1, 0.5):
        assert RM.py_cpu_nms(d, thr) == oracle.py_cpu_nms(d, thr)
    assert RM.py_cpu_nms(np.zeros((0, 5)), 0.3) == []
    sc = W.merge_scene(num_objects=200, scene=1500, seed=4)
    dets = np.concatenate([sc["polys"], sc["scores"][:, None]], 1)
    assert RM.py_cpu_nms_poly(dets, 0.1) == oracle.py_cpu_nms_poly_fast(dets, 0.1)
    # file level: `tile score x1 y1 x2 y2`
    lines = []
    for k in range(600):
        x, y = rng.integers(0, 3) * 824, rng.integers(0, 3) * 824
        x1, y1 = rng.uniform(0, 900, 2)
        w, h = rng.uniform(10, 120, 2)
        lines.append("P0003__1.0__%d___%d %.4f %.4f %.4f %.4f %.4f\n" % (x, y, (k + 1) / 1000.0, x1, y1, x1 + w, y1 + h))
    (tmp_path / "src").mkdir()
    (tmp_path / "src" / "Vehicle.txt").write_text("".join(lines))
    RM.mergebyrec(str(tmp_path / "src"), str(tmp_path / "got"), nms_threshold_type=1)
    F.mergesingle(str(tmp_path / "want"), str(tmp_path / "src" / "Vehicle.txt"), 1, nms="rec")
    _same_dir(tmp_path / "want", tmp_path / "got")
    with pytest.raises(ValueError):
        RM.mergebase(str(tmp_path / "src"), str(tmp_path / "x"), max)


def test_py_cpu_nms_obb(cuda, oracle, tmp_path):
    """py_cpu_nms_obb: cv2.minAreaRect boxes -> nms_rotated_cpu (`>=`), ascending kept indices; mergebyobb files."""
    from rs_detection_b200.jdet.data.devkits import result_merge as RM
    from rs_detection_b200.jdet.ops.bbox_transforms import poly2obb
    sc = W.merge_scene(num_objects=250, scene=1800, seed=6)
    dets = np.concatenate([sc["polys"], sc["scores"][:, None]], 1)
    got = RM.py_cpu_nms_obb(dets, 0.3)
    obb = poly2obb(dets[:, :8]).astype(np.float32)
    s32 = dets[:, 8].astype(np.float32)
    order = np.argsort(-s32.astype(np.float64), kind="stable").astype(np.int32)
    want = np.nonzero(oracle.nms_rotated_keep(obb, order, 0.3, 5, ge=True))[0]
    assert np.array_equal(got, want) and 0 < len(want) < len(dets)
    res = W.tile_results(120, 3, 2200, 1, seed=8)
    F.write_before_nms(res, tmp_path / "before", W.FAIR1M_CLASSES)
    RM.mergebyobb(str(tmp_path / "before"), str(tmp_path / "after"))
    n_in = sum(len(open(tmp_path / "before" / f).readlines()) for f in os.listdir(tmp_path / "before"))
    n_out = sum(len(open(tmp_path / "after" / f).readlines()) for f in os.listdir(tmp_path / "after"))
    assert sorted(os.listdir(tmp_path / "before")) == sorted(os.listdir(tmp_path / "after")) and 0 < n_out < n_in
