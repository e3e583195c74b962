#!/usr/bin/env python
"""Build `oracle/_ref/` -- the UNMODIFIED reference arithmetic compiled for the host.

TEST INFRASTRUCTURE ONLY.  Nothing under `rs_detection_b200/` may import this.

The reference (zcablii/RS_detection, a JDet fork) keeps every native routine as a
C++/CUDA *string literal* inside a `.py` file that Jittor JIT-compiles
(`python/jdet/ops/box_iou_rotated.py:3-352`, `box_iou_rotated_v1.py:3-357`,
`nms_rotated.py:2-352`, `roi_align_rotated.py:7-255`, `roi_align_rotated_v1.py:7-299`,
`nms_poly.py:4-185`).  Jittor is not installable here, so this recipe

  1. reads those literals out of the reference files with `ast` (nothing is imported,
     nothing is executed),
  2. writes them -- verbatim, minus `#include <executor.h>` -- into scratch translation
     units under `oracle/_ref/src/` together with a ~20 line `extern "C"` driver, and
  3. compiles each into `oracle/_ref/lib<name>.so` with g++.

`oracle/_ref/` is git-ignored: reference source never enters the repository history.
The CUDA-only routines (RoIAlignRotated, poly IoU) are compiled *for the host* by
defining `__device__/__global__` away and running the reference's grid-stride loop with a
1x1 launch geometry, so that the reference arithmetic itself (not a restatement) pins the
oracle in this GPU-less container.

Usage:  python oracle/build_ref.py [--ref /root/reference] [--force]
"""
from __future__ import annotations

import argparse
import ast
import hashlib
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref")
SRC = os.path.join(OUT, "src")
DEFAULT_REF = os.environ.get("RSDET_REF", "/root/reference")
OPS = "python/jdet/ops"


# --------------------------------------------------------------------------- extraction
def _module_strings(path: str) -> dict:
    """Evaluate module-level `NAME = <str> + NAME + <str> ...` assignments of a file."""
    with open(path, "r", encoding="utf-8") as f:
        tree = ast.parse(f.read(), path)
    env: dict = {}

    def ev(node):
        if isinstance(node, ast.Constant) and isinstance(node.value, str):
            return node.value
        if isinstance(node, ast.Name) and node.id in env:
            return env[node.id]
        if isinstance(node, ast.BinOp) and isinstance(node.op, ast.Add):
            return ev(node.left) + ev(node.right)
        raise ValueError("not a string expression")

    for st in tree.body:
        if isinstance(st, ast.Assign) and len(st.targets) == 1 and isinstance(st.targets[0], ast.Name):
            try:
                env[st.targets[0].id] = ev(st.value)
            except ValueError:
                pass
    return env


def _strip_jittor(s: str) -> str:
    return s.replace("#include <executor.h>", "")


def _cut_kernels(s: str) -> str:
    """Drop everything from the first __global__ kernel on (host builds of CUDA headers)."""
    for marker in ("template <typename T>\n__global__", "__global__ void"):
        i = s.find(marker)
        if i >= 0:
            return s[:i]
    return s


# --------------------------------------------------------------------------- drivers
IOU_DRIVER = r'''
extern "C" void ref_box_iou(const float* b1, int n, const float* b2, int m, float* out) {
  for (int i = 0; i < n; i++)
    for (int j = 0; j < m; j++)
      out[(long)i * m + j] = single_box_iou_rotated<float>(b1 + i * 5, b2 + j * 5);
}
'''

# greedy loop = the body of ML_NMS_ROTATED_CPU_SRC (nms_rotated.py:414-449) with the
# Jittor @alias glue replaced by plain pointers; `ge` selects the CPU (>=) or CUDA (>) test.
NMS_DRIVER = r'''
extern "C" float ref_single_iou(const float* a, const float* b) {
  return single_box_iou_rotated<float>(a, b);
}
extern "C" void ref_nms_greedy(const float* dets, const int* order, int n, float thr, int ge,
                               unsigned char* keep) {
  unsigned char* sup = new unsigned char[n > 0 ? n : 1]();
  for (int i = 0; i < n; i++) keep[i] = 0;
  for (int _i = 0; _i < n; _i++) {
    int i = order[_i];
    if (sup[i] == 1) continue;
    keep[i] = 1;
    for (int _j = _i + 1; _j < n; _j++) {
      int j = order[_j];
      if (sup[j] == 1) continue;
      float ovr = single_box_iou_rotated<float>(dets + (long)i * BOX_LENGTH, dets + (long)j * BOX_LENGTH);
      if (ge ? (ovr >= thr) : (ovr > thr)) sup[j] = 1;
    }
  }
  delete[] sup;
}
'''

CUDA_HOST_SHIM = r'''
#include <cmath>
#include <cstring>
#include <algorithm>
using namespace std;
#define __device__
#define __global__
#define __host__
#define __forceinline__ inline
struct _rs_dim3 { int x, y, z; };
static _rs_dim3 blockIdx = {0, 0, 0}, threadIdx = {0, 0, 0}, blockDim = {1, 1, 1}, gridDim = {1, 1, 1};
static inline float atomicAdd(float* p, float v) { float o = *p; *p = o + v; return o; }
struct float2 { float x, y; };
static inline float2 make_float2(float x, float y) { float2 r; r.x = x; r.y = y; return r; }
'''

ROI_DRIVER = r'''
extern "C" void ref_roi_fwd(const float* feat, const float* rois, int K, int C, int H, int W,
                            float scale, int sample_num, int ph, int pw, float* out) {
  int total = K * C * ph * pw;
  ROIAlignRotatedForward<float>(total, feat, rois, scale, sample_num, C, H, W, ph, pw, out);
}
extern "C" void ref_roi_bwd(const float* grad, const float* rois, int K, int N, int C, int H, int W,
                            float scale, int sample_num, int ph, int pw, float* gin) {
  memset(gin, 0, sizeof(float) * (size_t)N * C * H * W);
  int total = K * C * ph * pw;
  ROIAlignBackward<float>(total, grad, rois, scale, sample_num, C, H, W, ph, pw, gin);
}
'''

POLY_DRIVER = r'''
extern "C" float ref_poly_iou(const float* p, const float* q) { return devPolyIoU(p, q); }
extern "C" void ref_poly_iou_matrix(const float* p, int n, const float* q, int m, int stride, float* out) {
  for (int i = 0; i < n; i++)
    for (int j = 0; j < m; j++) out[(long)i * m + j] = devPolyIoU(p + i * stride, q + j * stride);
}
// mask + greedy scan of nms_poly.py:135-229 run serially: rows are score-sorted (n,9).
extern "C" void ref_poly_nms_sorted(const float* polys, int n, float thr, unsigned char* keep) {
  unsigned char* rem = new unsigned char[n > 0 ? n : 1]();
  for (int i = 0; i < n; i++) {
    keep[i] = 0;
    if (rem[i]) continue;
    keep[i] = 1;
    for (int j = i + 1; j < n; j++)
      if (devPolyIoU(polys + (long)i * 9, polys + (long)j * 9) > thr) rem[j] = 1;
  }
  delete[] rem;
}
'''


def _targets(ref: str) -> dict:
    ops = os.path.join(ref, OPS)
    iou0 = _module_strings(os.path.join(ops, "box_iou_rotated.py"))
    iou1 = _module_strings(os.path.join(ops, "box_iou_rotated_v1.py"))
    nms = _module_strings(os.path.join(ops, "nms_rotated.py"))
    roi0 = _module_strings(os.path.join(ops, "roi_align_rotated.py"))
    roi1 = _module_strings(os.path.join(ops, "roi_align_rotated_v1.py"))
    poly = _module_strings(os.path.join(ops, "nms_poly.py"))

    host_macros = "#define __host__\n#define __device__\n#define __forceinline__ inline\n"
    t = {}
    t["ref_iou_v0"] = "#include <cstring>\n" + _strip_jittor(iou0["IOU_ROTATED_CPU_HEADER"]) + IOU_DRIVER
    t["ref_iou_v1"] = "#include <cstring>\n" + _strip_jittor(iou1["IOU_ROTATED_CPU_HEADER"]) + IOU_DRIVER
    # the CUDA flavour (bubble sort with the 1e-6 tie rule) compiled for the host
    t["ref_iou_v0_cudasort"] = host_macros + _cut_kernels(_strip_jittor(iou0["IOU_ROTATED_CUDA_HEADER"])) + IOU_DRIVER
    t["ref_iou_v1_cudasort"] = host_macros + _cut_kernels(_strip_jittor(iou1["IOU_ROTATED_CUDA_HEADER"])) + IOU_DRIVER
    for bl in (5, 6):
        t[f"ref_nms{bl}"] = (f"#define BOX_LENGTH {bl}\n" + _strip_jittor(nms["ML_NMS_ROTATED_CPU_HEADER"]) + NMS_DRIVER)
        t[f"ref_nms{bl}_cudasort"] = (f"#define BOX_LENGTH {bl}\n" + host_macros
                                      + _cut_kernels(_strip_jittor(nms["ML_NMS_ROTATED_CUDA_HEADER"])) + NMS_DRIVER)
    t["ref_roi_v0"] = CUDA_HOST_SHIM + roi0["CUDA_HEADER"] + ROI_DRIVER
    t["ref_roi_v1"] = CUDA_HOST_SHIM + roi1["CUDA_HEADER"] + ROI_DRIVER
    t["ref_poly"] = CUDA_HOST_SHIM + _cut_kernels(_strip_jittor(poly["HEADER"])) + POLY_DRIVER
    return t


FLAVOURS = {
    # what g++ -O2 does on x86-64 by default: no contraction -> IEEE single ops as written
    "": ["-O2", "-ffp-contract=off"],
    # what nvcc does by default on the device (FMA contraction); used to size the 1e-6 band
    "_fma": ["-O2", "-mfma", "-ffp-contract=fast"],
    # the TIMED CPU arm of bench.py (never used for parity): what Jittor's JIT would build, `-O3 -march=native`
    # (SURVEY 8d).  The libraries are built in this container and run on the GPU box's host, so `native` is
    # replaced by its portable equivalent x86-64-v3 (AVX2 + FMA + BMI2: every CPU that hosts a B200).
    "_fast": ["-O3", "-march=x86-64-v3", "-ffp-contract=fast"],
}


def build(ref: str = DEFAULT_REF, force: bool = False, verbose: bool = True) -> list:
    if not os.path.isdir(os.path.join(ref, OPS)):
        raise FileNotFoundError(f"reference tree not found at {ref}")
    os.makedirs(SRC, exist_ok=True)
    built = []
    for name, code in _targets(ref).items():
        cpp = os.path.join(SRC, name + ".cpp")
        digest = hashlib.sha1(code.encode()).hexdigest()
        stamp = os.path.join(SRC, name + ".sha1")
        fresh = os.path.exists(stamp) and open(stamp).read() == digest
        with open(cpp, "w") as f:
            f.write(code)
        for suffix, flags in FLAVOURS.items():
            so = os.path.join(OUT, f"lib{name}{suffix}.so")
            if fresh and os.path.exists(so) and not force:
                built.append(so)
                continue
            cmd = ["g++", "-std=c++14", "-shared", "-fPIC", "-w", *flags, cpp, "-o", so]
            if verbose:
                print(" ".join(cmd))
            subprocess.check_call(cmd)
            built.append(so)
        with open(stamp, "w") as f:
            f.write(digest)
    return built


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--ref", default=DEFAULT_REF)
    ap.add_argument("--force", action="store_true")
    a = ap.parse_args()
    libs = build(a.ref, a.force)
    print(f"built {len(libs)} reference libraries under {OUT}")
    sys.exit(0)
