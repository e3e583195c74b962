"""shapely.geometry.Polygon for convex quadrilaterals, exact arithmetic (see the package docstring)."""
import os
import sys
from fractions import Fraction

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import exact_geometry as X  # noqa: E402


class _Region:
    def __init__(self, area):
        self._area = area

    @property
    def area(self):
        return float(self._area)


class Polygon:
    def __init__(self, shell):
        pts = [(float(p[0]), float(p[1])) for p in shell]
        assert len(pts) == 4, "the shim handles the quadrilaterals of the rotated-box path only"
        self._poly8 = [c for p in pts for c in p]
        self._pts = X._ccw(X._pts(self._poly8))

    @property
    def area(self):
        return float(abs(X._area2(self._pts)) / 2)

    def intersection(self, other):
        inter = self._pts
        for i in range(4):
            if not inter:
                break
            inter = X._clip(inter, other._pts[i], other._pts[(i + 1) % 4])
        return _Region(abs(X._area2(inter)) / 2 if len(inter) >= 3 else Fraction(0))
