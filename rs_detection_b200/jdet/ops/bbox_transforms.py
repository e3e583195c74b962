"""jdet.ops.bbox_transforms -- only the conversions on the hot path:
obb2poly (python/jdet/ops/bbox_transforms.py:612-623), obb2hbb (:626-632), poly2hbb (:602-609)."""
from ... import core
from ._io import back, dev


def obb2poly(obboxes):
    o, fl = dev(obboxes)
    return back(core.obb2poly(o), fl)


def obb2hbb(obboxes):
    o, fl = dev(obboxes)
    return back(core.obb2hbb(o), fl)


def poly2hbb(polys):
    p, fl = dev(polys)
    return back(core.poly2hbb(p), fl)
