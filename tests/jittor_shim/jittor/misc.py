def _pair(x):
    return tuple(x) if isinstance(x, (tuple, list)) else (x, x)
