"""CPU: rs_detection_b200/jittor_adapter.py without Jittor (the mirror image of oracle/build_ref.py).

A recording stand-in for the `jittor` module captures every `jt.code(shape, dtype, inputs, cuda_header=, cuda_src=)`
the adapter issues; each captured CUDA body is wrapped into a function with the glue variables Jittor's JIT would
declare (`in<i>_p`, `in<i>_shape<j>`, `out<j>_p`, `out<j>`) and compiled with nvcc for sm_100a against
tests/jittor_stub/executor.h and the REAL include/rsdet.h.  This pins the C side of every op to the C ABI: a changed
prototype or struct field in rsdet.h breaks this test instead of a Jittor box at run time."""
import os
import shutil
import subprocess
import sys
import types

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
CTYPE = {"float32": "float", "float": "float", "int32": "int", "uint8": "unsigned char", "float64": "double"}


class V:
    """shape/dtype carrier playing jt.Var"""
    def __init__(self, shape, dtype="float32"):
        self.shape, self.dtype = tuple(shape), dtype
        self.ndim = len(self.shape)

    def numel(self):
        n = 1
        for d in self.shape:
            n *= d
        return n

    def reshape(self, *s):
        return V(s[0] if len(s) == 1 and isinstance(s[0], (tuple, list)) else s, self.dtype)

    def int32(self):
        return V(self.shape, "int32")

    def float32(self):
        return V(self.shape, "float32")

    def bool(self):
        return V(self.shape, "bool")

    def argsort(self, dim=0, descending=False):
        return V(self.shape, "int32"), V(self.shape, self.dtype)

    def item(self):
        return 3

    def __getitem__(self, k):
        if isinstance(k, tuple) and len(k) == 2 and isinstance(k[1], int):
            return V(self.shape[:1], self.dtype)
        if isinstance(k, tuple) and len(k) == 2 and k[1] is None:
            return V(self.shape + (1,), self.dtype)
        if isinstance(k, tuple) and len(k) == 2 and isinstance(k[1], slice):
            n = len(range(*k[1].indices(self.shape[1])))
            return V((self.shape[0], n), self.dtype)
        return V(self.shape, self.dtype)

    def __rsub__(self, o):
        return self

    def __sub__(self, o):
        return self


@pytest.fixture()
def fake_jt(monkeypatch):
    rec = []
    jt = types.ModuleType("jittor")

    def code(shapes, dtypes, inputs, cuda_header="", cuda_src="", **kw):
        multi = isinstance(shapes, list)
        shp = shapes if multi else [shapes]
        dts = dtypes if isinstance(dtypes, list) else [dtypes] * len(shp)
        rec.append(dict(outs=list(zip(shp, dts)), ins=list(inputs), header=cuda_header, src=cuda_src))
        outs = [V(s_, d) for s_, d in zip(shp, dts)]
        return outs if multi else outs[0]

    class Function:
        @classmethod
        def apply(cls, *a):
            f = cls()
            out = f.execute(*a)
            f.grad(V(out.shape, out.dtype))
            return out

    jt.code, jt.Function = code, Function
    jt.zeros = lambda shape, dtype="float32": V(shape, dtype)
    jt.array = lambda a: V((0,))
    jt.where = lambda m: (V(m.shape, "int32"),)
    jt.arange = lambda n: V((n,), "int32")
    jt.concat = lambda xs, dim=0: V((xs[0].shape[0], sum(x.shape[1] for x in xs)), xs[0].dtype)
    monkeypatch.setitem(sys.modules, "jittor", jt)
    return rec


def _wrap(i, r):
    decl = []
    for k, v in enumerate(r["ins"]):
        ct = CTYPE.get(v.dtype, "float")
        decl.append(f"  {ct}* in{k}_p = nullptr; jittor::Var* in{k} = nullptr;")
        decl += [f"  int in{k}_shape{d} = 1;" for d in range(max(len(v.shape), 4))]
    for k, (shape, dt) in enumerate(r["outs"]):
        decl.append(f"  {CTYPE.get(dt, 'float')}* out{k}_p = nullptr; jittor::Var* out{k} = nullptr;")
        decl += [f"  int out{k}_shape{d} = 1;" for d in range(max(len(shape), 4))]
    return r["header"] + f"\nvoid jit_op_{i}() {{\n" + "\n".join(decl) + "\n" + r["src"] + "\n}\n"


@pytest.mark.skipif(not os.path.exists(NVCC), reason="nvcc not available")
def test_every_jt_code_compiles_against_rsdet_h(fake_jt, tmp_path):
    from rs_detection_b200 import jittor_adapter as A
    f4 = lambda *s: V(s)
    A.obb2poly(f4(10, 5)); A.obb2hbb(f4(4, 3, 5)); A.poly2hbb(f4(10, 8))
    A.box_iou_rotated(f4(7, 5), f4(9, 5)); A.box_iou_rotated_v1(f4(7, 5), f4(9, 5))
    A.assign_wrt_overlaps(f4(6, 100), 0.5, 0.5, 0.5, False, True, V((6,), "int32"), -1)
    A.assign_wrt_overlaps(f4(6, 100), 0.7, (0.1, 0.3), 0.3, True)
    A.nms_rotated(f4(50, 5), f4(50), 0.1); A.ml_nms_rotated(f4(50, 5), f4(50), V((50,), "int32"), 0.1)
    A.nms_rotated_cuda(f4(50, 5), V((50,), "int32"), 0.1); A.nms_rotated_cpu(f4(50, 6), V((50,), "int32"), 0.3, 6)
    A.poly_nms(f4(40, 9), 0.1); A.multiclass_poly_nms(f4(40, 8), f4(40), V((40,), "int32"), 0.1)
    A.multiclass_nms_rotated(f4(100, 5), f4(100, 11), 0.05, dict(iou_thr=0.1), 2000)
    A.multiclass_nms_rotated(f4(100, 55), f4(100, 11), 0.05, dict(iou_thr=0.1), -1, f4(100))
    for version in (0, 1):
        A.make_roi_align(version).apply(f4(2, 256, 64, 64), f4(30, 6), (7, 7), 1 / 16., 2)
    fused = A.make_fused_extractor(1)
    fused.apply(f4(1, 256, 256, 256), f4(1, 256, 128, 128), f4(1, 256, 64, 64), f4(1, 256, 32, 32), f4(4000, 6),
                [4, 8, 16, 32], 7, 2, (1.4, 1.2), 56)
    fused.apply(f4(1, 16, 32, 32), f4(30, 6), [4], (7, 7), 2, (1.4, 1.2), 56)
    A.rpn_proposals([f4(3, 64, 64), f4(3, 32, 32)], [f4(18, 64, 64), f4(18, 32, 32)], [f4(12288, 4), f4(3072, 4)], 3)
    assert len(fake_jt) >= 22
    src = tmp_path / "adapter_ops.cu"
    # one translation unit: the (identical) header once, then every op body as its own function
    body = fake_jt[0]["header"] + "\nnamespace jittor { Executor exe; }\n"
    for i, r in enumerate(fake_jt):
        assert r["header"] == fake_jt[0]["header"]
        body += _wrap(i, dict(r, header=""))
    src.write_text(body)
    cmd = [NVCC, "-gencode", "arch=compute_100a,code=sm_100a", "-std=c++17", "-c", str(src), "-o", str(tmp_path / "a.o"),
           "-I", os.path.join(ROOT, "tests", "jittor_stub"), "-Wno-deprecated-gpu-targets"]
    p = subprocess.run(cmd, capture_output=True, text=True)
    assert p.returncode == 0, p.stderr[-4000:]
    # the adapter includes the header instead of re-declaring structs
    text = open(os.path.join(ROOT, "rs_detection_b200", "jittor_adapter.py")).read()
    assert '#include "%s"' in text and "struct Cfg {" not in text and "abort();" not in text


def test_python_side_guards(fake_jt):
    from rs_detection_b200 import jittor_adapter as A
    assert A.box_iou_rotated(V((0, 5)), V((9, 5))).shape == (0, 9) and not fake_jt      # empty: no launch
    d, l = A.multiclass_nms_rotated(V((0, 5)), V((0, 11)), 0.05, {}, 100)
    assert d.shape == (0, 6) and not fake_jt
    with pytest.raises(ValueError):
        A.nms_rotated(V(((1 << 18) + 1, 5)), V(((1 << 18) + 1,)), 0.1)
    with pytest.raises(ValueError):
        A.assign_wrt_overlaps(V((0, 10)), 0.5, 0.5)
