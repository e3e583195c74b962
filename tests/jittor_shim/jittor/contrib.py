def concat(xs, dim=0):
    from . import concat as _c
    return _c(xs, dim)
