"""Stand-in for the slice of Shapely 1.8 the reference's merge / evaluation stage uses (`shapely.geometry.Polygon`:
`.area`, `.intersection(other).area` of convex quadrilaterals).  Shapely / GEOS is neither under /root/reference nor
installable here; this package plays its part with EXACT rational arithmetic (tests/exact_geometry.py) rounded once to
float64, i.e. the value any correctly rounded float64 implementation approximates.  Test infrastructure only: it
exists so that the reference's OWN Python (result_merge.py, voc_eval.py, tools/merge_results.py) can be executed to
generate fixtures (tests/golden/make_golden_devkit.py)."""
