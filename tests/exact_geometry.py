"""Exact-rational convex-quad geometry (fractions.Fraction) -- the arbiter for the merge stage's polygon IoU.

The reference delegates `iou_poly` (python/jdet/ops/nms_poly.py:247-252) to Shapely 1.8.2 / GEOS, which is neither
under /root/reference nor installable here, so the float64 clipper of the oracle (and of the device engine) cannot
be compared with GEOS itself.  This module computes the SAME mathematical quantity without rounding at all: float64
inputs are converted to Fractions exactly, the intersection of two convex quadrilaterals is clipped exactly
(Sutherland-Hodgman), areas are exact shoelace sums.  Any correctly rounded float64 implementation -- GEOS included
-- lies within a few ulps of these values, so agreement with them to 1e-12 is a stronger statement than agreement
with one particular float64 library.  Test infrastructure only.
"""
from fractions import Fraction


def _pts(poly8):
    return [(Fraction(float(poly8[2 * i])), Fraction(float(poly8[2 * i + 1]))) for i in range(4)]


def _area2(p):
    """twice the signed area"""
    s = Fraction(0)
    for i in range(len(p)):
        x1, y1 = p[i]
        x2, y2 = p[(i + 1) % len(p)]
        s += x1 * y2 - x2 * y1
    return s


def _ccw(p):
    return p if _area2(p) >= 0 else p[::-1]


def _clip(subject, a, b):
    """keep the part of `subject` on the left of (or on) the directed line a->b"""
    def side(pt):
        return (b[0] - a[0]) * (pt[1] - a[1]) - (b[1] - a[1]) * (pt[0] - a[0])
    out = []
    n = len(subject)
    for i in range(n):
        cur, nxt = subject[i], subject[(i + 1) % n]
        sc, sn = side(cur), side(nxt)
        if sc >= 0:
            out.append(cur)
        if (sc > 0 and sn < 0) or (sc < 0 and sn > 0):
            t = sc / (sc - sn)
            out.append((cur[0] + t * (nxt[0] - cur[0]), cur[1] + t * (nxt[1] - cur[1])))
    return out


def is_convex(poly8):
    p = _ccw(_pts(poly8))
    for i in range(4):
        a, b, c = p[i], p[(i + 1) % 4], p[(i + 2) % 4]
        if (b[0] - a[0]) * (c[1] - b[1]) - (b[1] - a[1]) * (c[0] - b[0]) < 0:
            return False
    return True


def areas_exact(poly_a, poly_b):
    """(area A, area B, area of the intersection) as Fractions; both quads must be convex"""
    pa, pb = _ccw(_pts(poly_a)), _ccw(_pts(poly_b))
    inter = pa
    for i in range(4):
        if not inter:
            break
        inter = _clip(inter, pb[i], pb[(i + 1) % 4])
    ia = abs(_area2(inter)) / 2 if len(inter) >= 3 else Fraction(0)
    return abs(_area2(pa)) / 2, abs(_area2(pb)) / 2, ia


def iou_poly_exact(poly_a, poly_b):
    """nms_poly.py:247-252 with exact arithmetic: inter / max(A + B - inter, 0.01)"""
    a, b, i = areas_exact(poly_a, poly_b)
    return i / max(a + b - i, Fraction(1, 100))


def greedy_merge_nms_exact(dets, thresh, order):
    """py_cpu_nms_poly_fast (result_merge.py:66-127) with exact polygon IoUs.  Returns (keep list, smallest
    |iou - thresh| met in a decision).  The hbb prefilter (`hbb_ovr > 0`) is evaluated in float64 exactly as the
    reference does (it only decides WHICH pairs reach the polygon test; pairs it drops have iou 0 <= thresh)."""
    import numpy as np
    d = np.asarray(dets, np.float64)
    x1, y1 = d[:, 0:8:2].min(1), d[:, 1:8:2].min(1)
    x2, y2 = d[:, 0:8:2].max(1), d[:, 1:8:2].max(1)
    areas = (x2 - x1 + 1) * (y2 - y1 + 1)
    thr = Fraction(float(thresh))
    order = list(order)
    keep, margin = [], None
    while order:
        i = order[0]
        keep.append(int(i))
        rest = []
        for j in order[1:]:
            w = max(0.0, min(x2[i], x2[j]) - max(x1[i], x1[j]))
            h = max(0.0, min(y2[i], y2[j]) - max(y1[i], y1[j]))
            hb = w * h
            if hb / (areas[i] + areas[j] - hb) > 0:
                iou = iou_poly_exact(d[i, :8], d[j, :8])
                m = abs(iou - thr)
                margin = m if margin is None or m < margin else margin
                if iou <= thr:
                    rest.append(j)
            else:
                rest.append(j)
        order = rest
    return keep, (float(margin) if margin is not None else None)
