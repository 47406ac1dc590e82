"""Model checks of the reference's `rfest/check.py` that are numerics rather than plots:
held-out scores, the stimulus permutation test, the Wald significance test and the residuals.

Same names and arguments as the reference.  The permutation test is the one that is worth running
on the device: every permutation is a full projection (K1) + forward pass (K4) of the test set, so
the frames are uploaded once and permuted in HBM instead of being re-copied 100 times.
"""

import numpy as np
import scipy.stats

__all__ = ["eval_model_score", "compute_permutation_test", "significance", "residuals_pearson",
           "residuals_deviance"]


def eval_model_score(model, X, y, stim=True, history=True, metric="corrcoef", w_type="opt"):
    """check.py:9-21: score of the fitted model on (X, y); `stim` / `history` choose which of the two
    filters take part (the history filter is driven by the response itself)."""
    if not (stim or history):
        raise ValueError()
    inputs = {}
    if stim:
        inputs["stimulus"] = X
    if history:
        inputs["history"] = y
    return model.score(X_test=inputs, y_test=y, metric=metric, w_type=w_type)


def compute_permutation_test(model, X_test, y_test, n_perm=100, history=True, metric="corrcoef",
                             w_type="opt"):
    """check.py:24-53: score on the true test stimulus vs. scores with the frames shuffled in time.

    The permutations are drawn from `np.random.permutation` exactly as in the reference (seed the
    global NumPy generator to reproduce a run); the frames are gathered on the device."""
    import torch

    score_trueX = eval_model_score(model=model, X=X_test, y=y_test, stim=True, history=history,
                                   metric=metric, w_type=w_type)
    dev = torch.device("cuda", model.engine.device)
    if isinstance(X_test, torch.Tensor):
        frames = X_test.to(dev, dtype=torch.float32)
    else:
        frames = torch.from_numpy(np.ascontiguousarray(X_test, dtype=np.float32)).to(dev)
    ci, model.compute_ci = model.compute_ci, False      # the bands of 100 shuffled predictions are never read
    try:
        score_permX = np.full(n_perm, np.nan)
        for i in range(n_perm):
            perm = torch.from_numpy(np.random.permutation(np.arange(frames.shape[0]))).to(dev)
            score_permX[i] = eval_model_score(model=model, X=frames[perm], y=y_test, stim=True,
                                              history=history, metric=metric, w_type=w_type)
    finally:
        model.compute_ci = ci
    return score_trueX, score_permX


def significance(model, w_type="opt", show_results=False):
    """check.py:103-126: Wald test of every filter's coefficients against zero,
    W = p^T V^-1 p ~ chi2(sum(df))."""
    W_values, p_values = {}, {}
    for name in model.filter_names:
        coef = np.asarray(model.p[w_type][name], dtype=np.float64)
        cov = np.asarray(model.V[w_type][name], dtype=np.float64)
        W_values[name] = np.squeeze(coef.T @ np.linalg.solve(cov, coef))
        p_values[name] = scipy.stats.chi2.sf(W_values[name], df=sum(model.df[name]))
        if show_results:
            verdict = "significant" if p_values[name] < 0.05 else "not significant"
            print(f"{name}: \n\t{verdict} \n\tW={W_values[name]:.3f}, p_value={p_values[name]:.3f}")
    return W_values, p_values


def residuals_pearson(y, y_pred):
    """check.py:129-132."""
    y, y_pred = np.asarray(y), np.asarray(y_pred)
    return (y - y_pred) / np.sqrt(y_pred)


def residuals_deviance(y, y_pred):
    """check.py:135-142."""
    y, y_pred = np.asarray(y, dtype=np.float64), np.asarray(y_pred, dtype=np.float64)
    quo = y / y_pred
    rsd = y - y_pred
    return np.sign(rsd) * np.sqrt(2 * (y * np.log(quo, out=np.zeros_like(quo), where=(quo != 0)) - rsd))
