"""A torch-backed stand-in for the slice of the Jittor API that the reference's box/head/assigner Python uses.

TEST INFRASTRUCTURE ONLY (tests/golden/make_golden_refpy.py): Jittor is not installable in this image, so this shim
lets the reference's OWN Python files (`jdet/ops/bbox_transforms.py`, `jdet/models/boxes/{coder,assigner}.py`,
`jdet/models/roi_heads/oriented_head.py`) be imported BY PATH and executed on CPU to produce golden vectors.  What it
pins is the reference's Python logic (operation order, masks, reshapes, clamps); the primitives underneath are
torch's float32 CPU kernels instead of Jittor's, with Jittor's calling conventions restated here:
  * `Var.argmax(dim)` / `jt.argmax` return (index, value); `Var.max(dim)` / `jt.max(x, dim)` return values only;
  * `clamp(min_v=, max_v=)`; `jt.argsort` returns (index, sorted values); comparison results take part in
    arithmetic as 0/1 (`1 - (w > h)`); float `%` is the floored modulo (Jittor: a - b * floor(a / b));
  * assignments through an index cast the value to the destination dtype.
Nothing in the product imports this package.
"""
import numpy as np
import torch

from . import nn, misc, contrib, init  # noqa: F401

_DT = {"float32": torch.float32, "float": torch.float32, "float64": torch.float64, "int32": torch.int32, "int": torch.int32,
       "int64": torch.int64, "bool": torch.bool, "uint8": torch.uint8}


def _dt(d):
    if d is None or isinstance(d, torch.dtype):
        return d
    return _DT[str(d)]


def _num(x):
    """comparison results behave as 0/1 integers in arithmetic"""
    return x.to(torch.int32) if isinstance(x, torch.Tensor) and x.dtype == torch.bool else x


class Var(torch.Tensor):
    @staticmethod
    def __new__(cls, data):
        return torch.Tensor._make_subclass(cls, data.detach() if isinstance(data, torch.Tensor) else torch.as_tensor(data))

    # ---- jittor conventions that differ from torch
    def argmax(self, dim=None, keepdims=False):
        v, i = torch.Tensor.max(self, dim=dim, keepdim=keepdims)
        return i, v

    def argmin(self, dim=None, keepdims=False):
        v, i = torch.Tensor.min(self, dim=dim, keepdim=keepdims)
        return i, v

    def max(self, dim=None, keepdims=False, **kw):
        dim = kw.get("dims", dim)
        return torch.Tensor.max(self) if dim is None else torch.Tensor.max(self, dim=dim, keepdim=keepdims)[0]

    def min(self, dim=None, keepdims=False, **kw):
        dim = kw.get("dims", dim)
        return torch.Tensor.min(self) if dim is None else torch.Tensor.min(self, dim=dim, keepdim=keepdims)[0]

    def argsort(self, dim=-1, descending=False):
        v, i = torch.sort(self, dim=dim, descending=descending, stable=True)   # Jittor: (index, sorted values); ties: lower index first
        return _v(i), _v(v)

    def astype(self, dtype):
        return _v(self.to(_dt(dtype)))

    def clamp(self, min_v=None, max_v=None):
        return torch.Tensor.clamp(self, min=min_v, max=max_v)

    def float(self):
        return self.to(torch.float32)

    def float32(self):
        return self.to(torch.float32)

    def int(self):
        return self.to(torch.int32)

    def int32(self):
        return self.to(torch.int32)

    def bool(self):
        return self.to(torch.bool)

    def any_(self):
        return bool(torch.Tensor.any(self))

    def numpy(self):
        return torch.Tensor.numpy(self.as_subclass(torch.Tensor).detach())

    def sync(self):
        return self

    def stop_grad(self):
        return self

    def unbind(self, dim=0):
        return torch.Tensor.unbind(self, dim)

    def __setitem__(self, key, value):
        if isinstance(value, torch.Tensor) and value.dtype != self.dtype:
            value = value.to(self.dtype)
        return torch.Tensor.__setitem__(self, key, value)

    # comparison results in arithmetic
    def __sub__(self, o):
        return torch.Tensor.__sub__(_num(self), _num(o))

    def __rsub__(self, o):
        return torch.Tensor.__rsub__(_num(self), _num(o))

    def __add__(self, o):
        return torch.Tensor.__add__(_num(self), _num(o))

    __radd__ = __add__

    def __mul__(self, o):
        return torch.Tensor.__mul__(_num(self), _num(o))

    __rmul__ = __mul__

    def __neg__(self):
        return torch.Tensor.__neg__(_num(self))


def _v(x):
    return x.as_subclass(Var) if isinstance(x, torch.Tensor) else x


def array(data, dtype=None):
    if isinstance(data, torch.Tensor):
        t = data.clone()
    else:
        a = np.asarray(data)
        if a.dtype == np.float64 and dtype is None:
            a = a.astype(np.float32)          # jt.array of python floats is float32
        t = torch.from_numpy(np.ascontiguousarray(a))
    if dtype is not None:
        t = t.to(_dt(dtype))
    return Var(t)


def zeros(*shape, dtype="float32"):
    if len(shape) == 1 and isinstance(shape[0], (tuple, list)):
        shape = tuple(shape[0])
    return Var(torch.zeros(shape, dtype=_dt(dtype)))


def ones(*shape, dtype="float32"):
    if len(shape) == 1 and isinstance(shape[0], (tuple, list)):
        shape = tuple(shape[0])
    return Var(torch.ones(shape, dtype=_dt(dtype)))


def full(shape, val, dtype=None):
    if dtype is None:
        dtype = torch.int32 if isinstance(val, int) else torch.float32
    return Var(torch.full(tuple(shape), val, dtype=_dt(dtype)))


def arange(start=0, end=None, step=1, dtype=None):
    if end is None:
        start, end = 0, start
    return _v(torch.arange(start, end, step, dtype=_dt(dtype)))


def zeros_like(x):
    return _v(torch.zeros_like(x))


def ones_like(x):
    return _v(torch.ones_like(x))


def concat(xs, dim=0):
    return _v(torch.cat([_num(x) if False else x for x in xs], dim=dim))


def stack(xs, dim=0):
    return _v(torch.stack(list(xs), dim=dim))


def split(x, sizes, dim=0):
    return tuple(_v(t) for t in torch.split(x, sizes, dim=dim))


def max(x, dim=None, keepdims=False):  # noqa: A001
    return x.max(dim, keepdims)


def min(x, dim=None, keepdims=False):  # noqa: A001
    return x.min(dim, keepdims)


def argmax(x, dim, keepdims=False):
    return x.argmax(dim, keepdims)


def argsort(x, dim=-1, descending=False):
    v, i = torch.sort(x, dim=dim, descending=descending, stable=True)
    return _v(i), _v(v)


def nonzero(x):
    return _v(torch.nonzero(x))


def where(c, a=None, b=None):
    if a is None:
        return tuple(_v(t) for t in torch.where(c))
    return _v(torch.where(c, a, b))


def sync_all(*a, **k):
    return None


def _wrap(fn):
    return lambda *a, **k: _v(fn(*a, **k))


cos, sin, log, exp, sqrt, abs, matmul, arctan2, maximum, minimum, floor = (  # noqa: A001
    _wrap(f) for f in (torch.cos, torch.sin, torch.log, torch.exp, torch.sqrt, torch.abs, torch.matmul, torch.atan2,
                       torch.maximum, torch.minimum, torch.floor))


class _Flags:
    use_cuda = 0


flags = _Flags()


def no_grad(*a, **k):
    import contextlib
    return contextlib.nullcontext()
