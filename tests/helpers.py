"""Synthetic data for the tests, in the style of the reference's fixtures.

`complex_small_rf` / `data_3d_stim` follow `rfest/simulate.py:99-124` (V1complex_2d)
and `rfest/generate_test_data.py:38-69` (generate_data_3d_stim with
noise='white', rf_kind='complex_small', y_distr='none'), except that the white
noise is seeded (the reference leaves it unseeded).
"""

import numpy as np
import scipy.stats

from oracle import design


def complex_small_rf():
    nt, nx = 30, 40
    dt = 1 / 60
    tt = np.arange(-nt * dt, 0, dt)
    kt1 = scipy.stats.gamma.pdf(-tt, nt / 7.5, scale=0.025)
    kt2 = scipy.stats.gamma.pdf(-tt, nx / 6, scale=0.03)
    kt1 /= np.linalg.norm(kt1)
    kt2 /= -np.linalg.norm(kt2)
    xx = np.linspace(-2, 2, nx)
    env = np.exp(-1 / (2 * 0.35 ** 2) * xx ** 2)
    kx1 = np.cos(2 * np.pi * xx / 2 + np.pi / 5) * env
    kx2 = np.sin(2 * np.pi * xx / 2 + np.pi / 5) * env
    kx1 /= np.linalg.norm(kx1)
    kx2 /= np.linalg.norm(kx2)
    k = np.vstack([kt1, kt2]).T @ np.vstack([kx1, kx2])
    k = k / np.linalg.norm(k)
    return k[23:30, 14:22]                                   # (7, 8)


def data_3d_stim(n_samples=1000, seed=2046):
    """-> w_true (6,7,8), stim (n,7,8), y (n,), dims."""
    frame = complex_small_rf()
    temporal = np.array([-1, -0.5, 0.1, 0.5, 1, 0.1])
    w_true = np.outer(temporal, frame).reshape((temporal.size,) + frame.shape)
    dims = w_true.shape
    rng = np.random.default_rng(seed)
    stim = rng.standard_normal((n_samples,) + dims[1:])
    X = design.lagged_design(stim, dims[0], shift=0)
    y = X @ w_true.flatten()
    return w_true, stim, y, dims


def split(X, y, frac_train, frac_dev):
    """Contiguous train/dev/test split (what `split_data`, rfest/utils.py:139-201, does)."""
    n = len(y)
    a = int(round(n * frac_train))
    b = a + int(round(n * frac_dev))
    return (X[:a], y[:a]), (X[a:b], y[a:b]), (X[b:], y[b:])


def df_rule(dims):
    """tests/test_glm.py:9-11 of the reference."""
    return [int(np.maximum(np.ceil(d / 2), 3)) for d in dims]


def uvec(x):
    return x / np.linalg.norm(x)


def lnp_data(n_raw, dims, seed=2046, rate=0.3, history=True):
    """Poisson spikes from a separable filter + optional negative spike-history
    kernel (small sizes).  -> stim (n,[nx,ny]), y (n,) float32."""
    rng = np.random.default_rng(seed)
    nlag = dims[0]
    spatial = dims[1:]
    stim = rng.standard_normal((n_raw,) + tuple(spatial)).astype(np.float32)
    t = np.arange(nlag)
    kt = np.sin(2 * np.pi * t / nlag) * np.exp(-(nlag - 1 - t) / (nlag / 3))
    w = kt
    for d in spatial:
        g = np.exp(-0.5 * ((np.arange(d) - (d - 1) / 2) / (d / 6)) ** 2)
        w = np.multiply.outer(w, g)
    w = w / np.linalg.norm(w)
    drive = design.lagged_design(stim.astype(np.float64), nlag) @ w.flatten()
    drive = 1.5 * drive / drive.std() + np.log(np.expm1(rate))
    y = np.zeros(n_raw)
    hk = -1.5 * np.exp(-np.arange(1, 21) / 4.0) if history else np.zeros(20)
    for i in range(n_raw):
        h = 0.0
        for k in range(1, min(i, 20) + 1):
            h += hk[k - 1] * y[i - k]
        lam = np.log1p(np.exp(drive[i] + h))
        y[i] = rng.poisson(lam)
    return stim, y.astype(np.float32), w
