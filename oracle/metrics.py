"""Oracle: prediction scores.  TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Restates `rfest/metrics.py:8-20`.  PINNED against the live reference module.
"""

import numpy as np


def r2(y, y_pred):
    """1 - SS_res / SS_tot (metrics.py:8-12)."""
    centre = y.mean()
    total = ((y - centre) ** 2).sum()
    resid = ((y - y_pred) ** 2).sum()
    return np.asarray(1 - (resid / total))


def mse(y, y_hat):
    """Mean squared error (metrics.py:15-16)."""
    return np.asarray(np.mean((y - y_hat) ** 2))


def corrcoef(y, y_pred):
    """Pearson r through `np.corrcoef` (metrics.py:19-20; fp64 inside NumPy)."""
    return np.corrcoef(y, y_pred)[0, 1]


def compute_score(y, y_pred, metric):
    """Dispatcher of `GLM.compute_score` (rfest/GLM/_glm.py:997-1013)."""
    if metric == "r2":
        return r2(y, y_pred)
    if metric == "mse":
        return mse(y, y_pred)
    if metric == "corrcoef":
        return corrcoef(y, y_pred)
    print(f"Metric `{metric}` is not supported.")
    return None
