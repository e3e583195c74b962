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


def poly2obb(polys):
    """:549-575 -- the reference's own host loop over `cv2.minAreaRect` (OpenCV, float32 points), le90-style
    normalisation.  Third-party arithmetic on the host in the reference too; returns a numpy (…, 5) array."""
    import cv2
    import numpy as np
    p = polys.detach().cpu().numpy() if hasattr(polys, "detach") else np.asarray(polys)
    order = p.shape[:-1]
    pts = p.reshape(-1, p.shape[-1] // 2, 2).astype(np.float32)
    out = []
    for poly in pts:
        (x, y), (w, h), angle = cv2.minAreaRect(poly)
        if w >= h:
            angle = -angle
        else:
            w, h = h, w
            angle = -90 - angle
        out.append([x, y, w, h, angle / 180 * np.pi])
    out = np.array(out) if out else np.zeros((0, 5))
    return out.reshape(*order, 5)
