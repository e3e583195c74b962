"""The C-ABI library loads, exports every symbol include/rsdet.h declares, and the host side fails
loudly (no CPU fallback).  No compute call is made here (no GPU needed)."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    h = open(os.path.join(ROOT, "include", "rsdet.h")).read()
    h = re.sub(r"/\*.*?\*/", "", h, flags=re.S)
    return sorted(set(re.findall(r"\b(rsdet_[a-z0-9_]+)\s*\(", h)))


def test_header_symbols_exported_and_bound():
    from rs_detection_b200 import _lib
    names = _declared()
    assert len(names) >= 20
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/rsdet.h but not exported by librsdet.so"
    L = _lib.load()
    assert sorted(_lib.SIGNATURES) == names, "python binding table and header disagree"
    assert L.rsdet_version() >= 100
    assert L.rsdet_error_string(0) == b"ok" and b"workspace" in L.rsdet_error_string(-2)


def test_workspace_queries_are_host_only():
    from rs_detection_b200 import _lib, core
    L = _lib.load()
    for kind in range(5):
        small, big = L.rsdet_nms_workspace_bytes(kind, 1000), L.rsdet_nms_workspace_bytes(kind, 100000)
        assert 0 < small < big
    # 100k boxes in one segment: n * ceil(n/64) 64-bit words dominate
    assert L.rsdet_nms_workspace_bytes(0, 100000) > 100000 * 1563 * 8
    assert L.rsdet_box_iou_rotated_workspace_bytes(512, 2000) >= (512 + 2000) * 48
    cfg = core.make_roi_cfg([(1, 256, 256, 256), (1, 256, 128, 128), (1, 256, 64, 64), (1, 256, 32, 32)],
                            [1 / 4, 1 / 8, 1 / 16, 1 / 32], 7, 2, 1, (1.4, 1.2))
    need = L.rsdet_roi_align_rotated_workspace_bytes(ctypes.byref(cfg), 4000, 0)
    assert need >= 87040 * 256 * 4  # a channels-last copy of the pyramid
    cfg.channels_last = 1
    assert L.rsdet_roi_align_rotated_workspace_bytes(ctypes.byref(cfg), 4000, 0) < 1 << 20


def test_argument_errors_without_touching_the_gpu():
    from rs_detection_b200 import _lib
    L = _lib.load()
    assert L.rsdet_obb2poly(None, -1, None, None) == -1
    assert L.rsdet_obb2poly(None, 0, None, None) == 0          # empty input is a no-op
    assert L.rsdet_box_iou_rotated(None, 0, None, 5, 0, 0, None, None, 0, None) == 0
    assert L.rsdet_box_iou_rotated(None, 3, None, 5, 7, 0, None, None, 0, None) == -1   # bad version
    assert L.rsdet_assign_wrt_overlaps(None, 0, 10, 0.5, 0, 0.5, 0.5, 0, 1, None, -1, None, None, None, None, 0, None) == -1
    assert L.rsdet_nms(9, None, None, None, 10, 0.1, None, 0, None, None, None, None, None, 0, None) == -1
    with pytest.raises(ValueError):
        _lib.check(-1, "x")
    with pytest.raises(RuntimeError):
        _lib.check(-2, "x")


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU behaviour")
def test_no_cpu_fallback():
    from rs_detection_b200.jdet.data.devkits.result_merge import py_cpu_nms_poly_fast
    from rs_detection_b200.jdet.ops import box_iou_rotated
    from rs_detection_b200.jdet.ops.nms_rotated import nms_rotated
    b = np.array([[0, 0, 1, 1, 0]], np.float32)
    for call in (lambda: box_iou_rotated(b, b), lambda: nms_rotated(torch.zeros(3, 5), torch.zeros(3), 0.1),
                 lambda: py_cpu_nms_poly_fast(np.zeros((2, 9)), 0.1)):
        with pytest.raises(RuntimeError, match="CUDA"):
            call()


def test_product_does_not_import_the_oracle():
    pkg = os.path.join(ROOT, "rs_detection_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dp, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f"{f} imports the oracle"
                assert "oracle." not in src and "oracle/" not in src, f"{f} mentions the oracle package"
