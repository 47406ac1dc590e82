"""Host-side logic of `rfest_b200.GLM.optimize` (the early-stopping scan, batching, model
selection) driven through a mock device model -- no GPU, no CUDA library.  The mock plays back
given traces the way the device writes them: after `executed` iterations rows `< executed`
of cost_train and rows `< executed - 1` of the post-update traces exist; `fit_finish(i)`
completes row i.  Reference rule: rfest/GLM/_glm.py:748-775 (only for i > min_iters)."""

import numpy as np
import pytest

from rfest_b200.glm import GLM


class MockDevice:
    def __init__(self, cost_train, cost_dev=None):
        self.ct = np.asarray(cost_train, float)
        self.cd = None if cost_dev is None else np.asarray(cost_dev, float)
        self.executed = 0
        self.finished = None
        self.runs = []

    def set_params(self, b, ic):
        pass

    def fit_begin(self, steps, alpha, beta, metric):
        self.n_max = len(steps)

    def fit_run(self, n):
        self.executed += n
        self.runs.append(n)

    def fit_traces(self, k):
        done = max(self.executed - 1, 0) if self.finished is None else self.finished + 1
        post = np.full(k, np.nan)
        post[:min(done, k)] = 0.5
        cd = np.full(k, np.nan)
        if self.cd is not None:
            cd[:min(done, k)] = self.cd[:min(done, k)]
        return self.ct[:k].copy(), cd, post.copy(), (post.copy() if self.cd is not None else np.full(k, np.nan))

    def fit_finish(self, i):
        self.finished = i

    def fit_params_all(self, k):
        return np.zeros((k, 3), np.float32), np.zeros((k, 2))

    def fit_predictions_device(self, kind):
        return np.zeros(4, np.float32)

    def fit_end(self):
        pass


def _glm(dev, has_dev):
    m = GLM.__new__(GLM)
    m._dm = dev
    m._sync = lambda: None
    m._flatten = lambda p: (np.zeros(3, np.float32), np.zeros(2))
    m._unflatten = lambda b, ic: {"stimulus": b.reshape(3, 1), "intercept": {"stimulus": ic[0], "global": ic[1]}}
    m.y = {"train": np.zeros(4)}
    if has_dev:
        m.y["dev"] = np.zeros(4)
    m.y_pred = {"opt": {}}
    m.alpha, m.beta = 1, 0.01
    return m


def _reference_stop(ct, cd, num_iters, tolerance, atol, min_iters):
    """The loop of _glm.py:720-782 on given traces -> (last i, stop)."""
    for i in range(num_iters):
        if tolerance and i > min_iters:
            if cd is not None and np.all(np.diff(cd[i - tolerance:i]) > 0):
                return i, "dev_stop"
            if np.all(-np.diff(ct[i - tolerance:i]) < atol):
                return i, "train_stop"
    return num_iters - 1, "maxiter_stop"


@pytest.mark.parametrize("verbose", [0, 1, True, 7])
def test_strictly_decreasing_cost_never_stops_early(verbose, capsys):
    """with verbose > 0 the scan used to restart at i = 1 and stop after 2 iterations."""
    n = 500
    ct = 100.0 * 0.99 ** np.arange(n)
    for has_dev in (False, True):
        dev = MockDevice(ct, 50.0 * 0.99 ** np.arange(n) if has_dev else None)
        m = _glm(dev, has_dev)
        m.optimize({}, n, "corrcoef", 0.1, 10, verbose, None)
        assert m.train_stop == "maxiter_stop" and len(m.cost_train) == n
    capsys.readouterr()


@pytest.mark.parametrize("verbose", [0, 1, 5])
@pytest.mark.parametrize("min_iters", [0, 20, 63, 64, 100, 300])
def test_flat_cost_stops_where_the_reference_does(verbose, min_iters, capsys):
    n = 700
    ct = np.full(n, 3.0)
    dev = MockDevice(ct)
    m = _glm(dev, False)
    m.optimize({}, n, "mse", 0.1, 10, verbose, None, atol=1e-5, min_iters=min_iters)
    i_ref, stop_ref = _reference_stop(ct, None, n, 10, 1e-5, min_iters)
    assert (len(m.cost_train) - 1, m.train_stop) == (i_ref, stop_ref) == (min_iters + 1, "train_stop")
    capsys.readouterr()


@pytest.mark.parametrize("verbose", [0, 3])
def test_rising_dev_cost_stops_at_the_reference_iteration(verbose, capsys):
    n, turn = 900, 417
    rng = np.random.default_rng(0)
    ct = 10.0 - 1e-3 * np.arange(n)
    cd = np.where(np.arange(n) < turn, 5.0 - 1e-3 * np.arange(n) + 1e-4 * rng.standard_normal(n),
                  5.0 + 1e-3 * (np.arange(n) - turn))
    dev = MockDevice(ct, cd)
    m = _glm(dev, True)
    m.optimize({}, n, "corrcoef", 0.1, 10, verbose, None, min_iters=300)
    i_ref, stop_ref = _reference_stop(ct, cd, n, 10, 1e-5, 300)
    assert stop_ref == "dev_stop"
    assert (len(m.cost_train) - 1, m.train_stop) == (i_ref, stop_ref)
    assert m.best_iteration == int(np.argmin(cd[:i_ref + 1]))
    m2 = _glm(MockDevice(ct, cd), True)
    m2.optimize({}, n, "corrcoef", 0.1, 10, verbose, "last", min_iters=300)
    assert m2.best_iteration == i_ref - 10                                   # _glm.py:805-806
    capsys.readouterr()
