"""Scores on host arrays (the reference's `rfest/metrics.py:8-20` is host NumPy as well).

During `GLM.fit` the engine does not use these: the per-iteration scores come from sums
accumulated inside the streaming sweep (see csrc/update.cuh).  They serve the static
`GLM.compute_score(y, y_pred, metric)` and `GLM.score(...)`, whose inputs are host arrays.
"""

import numpy as np


def r2(y, y_pred):
    y = np.asarray(y)
    y_pred = np.asarray(y_pred)
    ss_res = np.sum((y - y_pred) ** 2)
    ss_tot = np.sum((y - y.mean()) ** 2)
    return np.asarray(1 - ss_res / ss_tot)


def mse(y, y_hat):
    return np.asarray(np.mean((np.asarray(y) - np.asarray(y_hat)) ** 2))


def corrcoef(y, y_pred):
    return np.corrcoef(np.asarray(y), np.asarray(y_pred))[0, 1]


def score_from_sums(metric, n, sy, syy, sr, srr, syr, sse):
    """The same scores from the sums the sweep accumulates (what update.cuh computes)."""
    if metric == "mse":
        return sse / n
    if metric == "r2":
        return 1.0 - sse / (syy - sy * sy / n)
    if metric == "corrcoef":
        return (syr - sy * sr / n) / np.sqrt((syy - sy * sy / n) * (srr - sr * sr / n))
    return None
