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
