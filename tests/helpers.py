import numpy as np


def band_pairs(iou, thr, band=1e-6):
    """number of pairs whose IoU lies inside the parity contract's exclusion band around thr"""
    return int((np.abs(iou.astype(np.float64) - thr) <= band).sum())


def close_report(got, want, rtol, atol):
    """(#violations of |got-want| <= atol + rtol*|want|, max abs err, max |want|)"""
    got = np.asarray(got, np.float64)
    want = np.asarray(want, np.float64)
    err = np.abs(got - want)
    bad = err > (atol + rtol * np.abs(want))
    return int(bad.sum()), float(err.max() if err.size else 0.0), float(np.abs(want).max() if want.size else 0.0)
