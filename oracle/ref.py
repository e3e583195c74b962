"""ctypes front-end of `oracle/_ref/` -- the reference's own C++/CUDA source strings compiled
for the host by `oracle/build_ref.py` (TEST INFRASTRUCTURE ONLY).

`/root/reference` is needed only to BUILD these libraries (in the build container); the
built `.so` files travel to the GPU box.  `available()` tells callers whether they exist.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref")

_pf = C.POINTER(C.c_float)
_pi = C.POINTER(C.c_int)
_pu8 = C.POINTER(C.c_uint8)
_cache: dict = {}


def _path(name: str) -> str:
    return os.path.join(REF_DIR, f"lib{name}.so")


def available(name: str = "ref_iou_v0") -> bool:
    return os.path.exists(_path(name))


def ensure_built() -> bool:
    """Build oracle/_ref if the reference tree is present and the libraries are missing."""
    if available("ref_poly") and available("ref_roi_v1") and available("ref_nms6"):
        return True
    from . import build_ref
    if not os.path.isdir(os.path.join(build_ref.DEFAULT_REF, build_ref.OPS)):
        return False
    build_ref.build(verbose=False)
    return True


def _lib(name: str):
    if name not in _cache:
        _cache[name] = C.CDLL(_path(name))
    return _cache[name]


def _c32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _fp(a):
    return a.ctypes.data_as(_pf)


def box_iou(boxes1, boxes2, version=0, cudasort=False, fma=False):
    name = f"ref_iou_v{version}" + ("_cudasort" if cudasort else "") + ("_fma" if fma else "")
    L = _lib(name)
    L.ref_box_iou.argtypes = [_pf, C.c_int, _pf, C.c_int, _pf]
    b1, b2 = _c32(boxes1).reshape(-1, 5), _c32(boxes2).reshape(-1, 5)
    out = np.zeros((b1.shape[0], b2.shape[0]), np.float32)
    if out.size:
        L.ref_box_iou(_fp(b1), b1.shape[0], _fp(b2), b2.shape[0], _fp(out))
    return out


def _flavour(fma, fast):
    """'' = parity build (-O2, no contraction); '_fma' sizes the 1e-6 band; '_fast' (-O3 -march=x86-64-v3) is the
    timed CPU arm of bench.py and is never used for parity.  Falls back to the parity build if absent."""
    return "_fast" if fast else ("_fma" if fma else "")


def nms_keep(dets, order, thr, box_len=5, ge=True, cudasort=False, fma=False, fast=False):
    name = f"ref_nms{box_len}" + ("_cudasort" if cudasort else "") + _flavour(fma, fast)
    if fast and not available(name):
        name = name[:-5]
    L = _lib(name)
    L.ref_nms_greedy.argtypes = [_pf, _pi, C.c_int, C.c_float, C.c_int, _pu8]
    d = _c32(dets).reshape(-1, box_len)
    o = np.ascontiguousarray(order, np.int32)
    keep = np.zeros((d.shape[0],), np.uint8)
    if d.shape[0]:
        L.ref_nms_greedy(_fp(d), o.ctypes.data_as(_pi), d.shape[0], float(thr), int(ge), keep.ctypes.data_as(_pu8))
    return keep.astype(bool)


def roi_fwd(feat, rois, output_size, spatial_scale, sampling_ratio, version=1, fma=False, fast=False):
    name = f"ref_roi_v{version}" + _flavour(fma, fast)
    if fast and not available(name):
        name = name[:-5]
    L = _lib(name)
    L.ref_roi_fwd.argtypes = [_pf, _pf] + [C.c_int] * 4 + [C.c_float] + [C.c_int] * 3 + [_pf]
    feat, rois = _c32(feat), _c32(rois).reshape(-1, 6)
    N, Cc, H, W = feat.shape
    ph, pw = output_size
    out = np.zeros((rois.shape[0], Cc, ph, pw), np.float32)
    if out.size:
        L.ref_roi_fwd(_fp(feat), _fp(rois), rois.shape[0], Cc, H, W, np.float32(spatial_scale), int(sampling_ratio), ph, pw,
                      _fp(out))
    return out


def roi_bwd(grad, rois, feat_shape, spatial_scale, sampling_ratio, version=1, fma=False):
    L = _lib(f"ref_roi_v{version}" + ("_fma" if fma else ""))
    L.ref_roi_bwd.argtypes = [_pf, _pf] + [C.c_int] * 5 + [C.c_float] + [C.c_int] * 3 + [_pf]
    grad, rois = _c32(grad), _c32(rois).reshape(-1, 6)
    N, Cc, H, W = feat_shape
    gin = np.zeros(feat_shape, np.float32)
    L.ref_roi_bwd(_fp(grad), _fp(rois), rois.shape[0], N, Cc, H, W, np.float32(spatial_scale), int(sampling_ratio),
                  grad.shape[2], grad.shape[3], _fp(gin))
    return gin


def poly_iou_matrix(p, q, fma=False):
    L = _lib("ref_poly" + ("_fma" if fma else ""))
    L.ref_poly_iou_matrix.argtypes = [_pf, C.c_int, _pf, C.c_int, C.c_int, _pf]
    p, q = _c32(p), _c32(q)
    out = np.zeros((p.shape[0], q.shape[0]), np.float32)
    if out.size:
        L.ref_poly_iou_matrix(_fp(p), p.shape[0], _fp(q), q.shape[0], p.shape[1], _fp(out))
    return out


def poly_nms_sorted_keep(polys9_sorted, thr, fma=False):
    L = _lib("ref_poly" + ("_fma" if fma else ""))
    L.ref_poly_nms_sorted.argtypes = [_pf, C.c_int, C.c_float, _pu8]
    p = _c32(polys9_sorted).reshape(-1, 9)
    keep = np.zeros((p.shape[0],), np.uint8)
    if p.shape[0]:
        L.ref_poly_nms_sorted(_fp(p), p.shape[0], float(thr), keep.ctypes.data_as(_pu8))
    return keep.astype(bool)
