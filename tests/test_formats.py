"""CPU: on-disk formats around the merge stage -- the oracle restatement against
hand-written expectations, and the product's text layer (no device work) against the oracle."""
import os

import numpy as np

import workloads as W
from oracle import formats as F


def test_before_nms_writer_by_hand(tmp_path):
    res = [((np.array([[1, 2, 3, 4, 5, 6, 7, 8.12345]], np.float32), np.array([0.98765], np.float32), np.array([2])),
            {"img_file": "/x/P0007__1.0__824___0.png"}),
           ((np.array([[10, 20, 30, 40, 50, 60, 70, 80]], np.float32), np.array([0.5], np.float32), np.array([2])),
            {"img_file": "/x/P0007__1.0__0___0.png", "flip_mode": "HV", "ori_img_size": (1024, 1024)})]
    F.write_before_nms(res, tmp_path, W.FAIR1M_CLASSES)
    assert os.listdir(tmp_path) == ["Vehicle.txt"]
    lines = open(tmp_path / "Vehicle.txt").read().split("\n")
    assert lines[0] == "P0007__1.0__824___0 0.9876 1.0000 2.0000 3.0000 4.0000 5.0000 6.0000 7.0000 8.1235"
    assert lines[1] == "P0007__1.0__0___0 0.5000 1014.0000 1004.0000 994.0000 984.0000 974.0000 964.0000 954.0000 944.0000"


def test_product_writer_matches_oracle(tmp_path):
    from rs_detection_b200.jdet.data.devkits.data_merge import prepare_data
    res = W.tile_results(120, 10, 2200, 2, seed=1)
    res[3][1].update(flip_mode="H", ori_img_size=(1024, 1024))
    F.write_before_nms(res, tmp_path / "o", W.FAIR1M_CLASSES)
    prepare_data(res, tmp_path / "p", W.FAIR1M_CLASSES)
    assert sorted(os.listdir(tmp_path / "o")) == sorted(os.listdir(tmp_path / "p"))
    for f in os.listdir(tmp_path / "o"):
        assert open(tmp_path / "o" / f).read() == open(tmp_path / "p" / f).read()


def test_tile_name_parser():
    from rs_detection_b200.jdet.data.devkits.result_merge import parse_tile_name, read_tile_detections
    assert parse_tile_name("P0706__1.5__4120___824") == ("P0706", 4120, 824, 1.5)
    assert parse_tile_name("12__0.5__0___0") == ("12", 0, 0, 0.5)
    assert parse_tile_name("P1__1__0___13") == ("P1", 0, 13, 1.0)


def test_oracle_mergesingle_by_hand(tmp_path):
    # two tiles at rate 1.0 see the same 20x10 box (offset by the tile origin), a third box is far away
    src = tmp_path / "Ship.txt"
    src.write_text("P1__1.0__0___0 0.9000 100.0 100.0 120.0 100.0 120.0 110.0 100.0 110.0\n"
                   "P1__1.0__50___0 0.8000 50.0 100.0 70.0 100.0 70.0 110.0 50.0 110.0\n"
                   "P1__0.5__0___0 0.7000 300.0 300.0 310.0 300.0 310.0 305.0 300.0 305.0\n"
                   "P2__1.0__0___0 0.6000 100.0 100.0 120.0 100.0 120.0 110.0 100.0 110.0\n")
    F.mergesingle(tmp_path / "out", str(src))
    got = open(tmp_path / "out" / "Ship.txt").read().split("\n")
    assert got == ["P1 0.9 100.0 100.0 120.0 100.0 120.0 110.0 100.0 110.0",
                   "P1 0.7 600.0 600.0 620.0 600.0 620.0 610.0 600.0 610.0",
                   "P2 0.6 100.0 100.0 120.0 100.0 120.0 110.0 100.0 110.0", ""]
    # per-class thresholds: Ship uses 0.2 (result_merge.py:26-27); same outcome here, different code path
    F.mergesingle(tmp_path / "out1", str(src), nms_threshold_type=1)
    assert open(tmp_path / "out1" / "Ship.txt").read().split("\n") == got


def test_product_reader_matches_regex_restatement(tmp_path):
    from rs_detection_b200.jdet.data.devkits.result_merge import read_tile_detections
    res = W.tile_results(80, 3, 2200, 2, seed=2)
    F.write_before_nms(res, tmp_path, W.FAIR1M_CLASSES)
    scenes, order, polys, offs, scores = read_tile_detections(str(tmp_path / "Ship.txt"))
    lines = open(tmp_path / "Ship.txt").read().strip().split("\n")
    assert len(scenes) == len(lines) == polys.shape[0] == offs.shape[0] == scores.shape[0]
    assert order == sorted(set(scenes), key=scenes.index)
    for k in (0, len(lines) // 2, len(lines) - 1):
        sp = lines[k].split(" ")
        name, rate, rest = sp[0].split("__", 2)[0], float(sp[0].split("__", 2)[1]), sp[0].split("__", 2)[2]
        assert scenes[k] == name and offs[k].tolist() == [float(rest.split("___")[0]), float(rest.split("___")[1]), rate]
        assert scores[k] == float(sp[1]) and polys[k].tolist() == [float(v) for v in sp[2:]]


def test_fair1m_csv_oracle_by_hand(tmp_path):
    (tmp_path / "img").mkdir()
    (tmp_path / "img" / "P0012__1.0__0___0.png").write_text("")
    (tmp_path / "after").mkdir()
    (tmp_path / "after" / "Tennis_Court.txt").write_text("P0012 0.75 1.0 2.0 3.0 4.0 5.0 6.0 7.0 8.25\n")
    F.fair1m_csv(tmp_path / "after", tmp_path / "csv", tmp_path / "img", "sub")
    # the reference drops the last character of every line (dota_to_fair.py:25): 8.25 -> 8.2
    assert open(tmp_path / "csv" / "sub.csv").read() == "12.tif,Tennis_Court,1.0000,2.0000,3.0000,4.0000,5.0000,6.0000,7.0000,8.2000,0.7500\n"
    from rs_detection_b200.jdet.data.devkits.dota_to_fair import dota_to_fair1m_1_5
    dota_to_fair1m_1_5(tmp_path / "after", tmp_path / "csv2", tmp_path / "img", "sub")
    assert open(tmp_path / "csv2" / "sub.csv").read() == open(tmp_path / "csv" / "sub.csv").read()


def test_csv_roundtrip(tmp_path):
    from rs_detection_b200.jdet.merge import read_csv_to_numpy, save_to_csv
    rows = np.array([[12, 1, 2, 3, 4, 5, 6, 7, 8, 0.5, 3], [7, 10.5, 20.25, 30, 40, 50, 60, 70, 80, 0.1234, 10]], np.float64)
    save_to_csv(rows, tmp_path / "a.csv")
    assert open(tmp_path / "a.csv").read().split("\n")[0] == "12.tif,Vehicle,1.0000,2.0000,3.0000,4.0000,5.0000,6.0000,7.0000,8.0000,0.5000"
    assert np.array_equal(read_csv_to_numpy(tmp_path / "a.csv"), rows)
