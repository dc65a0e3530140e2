import os, sys
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def rel_err(a, b):
    """max |a-b| / max |b|  (the 'relative' of the north_star tolerances)."""
    import numpy as np
    a = np.asarray(a, float); b = np.asarray(b, float)
    den = np.abs(b).max()
    if den == 0.0:
        return float(np.abs(a).max())
    return float(np.abs(a - b).max() / den)


def rel_err_rows(a, b, floor=1e-6):
    """Row-wise relative error: max over rows of max|a_row - b_row| / max(max|b_row|, floor * max|b|).  Unlike rel_err (one
    global norm) a small-magnitude row -- the moments of a nearly flat plate, the forces of a quiet corner -- is held to its
    own size; `floor` only keeps rows that are pure rounding noise (below 1e-6 of the largest entry) from dividing by ~0."""
    import numpy as np
    a = np.asarray(a, float).reshape(len(a), -1); b = np.asarray(b, float).reshape(len(b), -1)
    top = np.abs(b).max()
    if top == 0.0:
        return float(np.abs(a).max())
    den = np.maximum(np.abs(b).max(axis=1), floor * top)
    return float((np.abs(a - b).max(axis=1) / den).max())
