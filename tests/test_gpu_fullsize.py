"""BASELINE.json's full sizes (config 3: 2^25 x (288 + 8); config 5: 2^24 x 1200), where no CPU
oracle can follow: size-independent properties of the sweep, each anchored on something computed
another way -- closed forms at zero parameters, the independent fp64 Gram kernel, finite
differences, bit-reproducibility."""

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _free_gb():
    import torch

    return torch.cuda.mem_get_info()[0] / 1e9


def _build(log2n, dims, df, distr, out_nl, history):
    from rfest_b200 import GLM
    from rfest_b200.engine import Engine
    from rfest_b200.synth import GaussianStimulus, TensorRows, gaussian_response, poisson_response

    n_raw = 1 << log2n
    eng = Engine.get()
    stim = GaussianStimulus(n_raw, dims[1:], seed=2046, device=eng.device)
    y = (poisson_response if distr == "poisson" else gaussian_response)(eng, stim, dims, 0, n_raw)
    m = GLM(distr=distr, output_nonlinearity=out_nl)
    m.add_design_matrix(stim, dims=dims, df=df, smooth="cr", name="stimulus")
    if history:
        m.add_design_matrix(TensorRows(y, 0, n_raw), dims=[20], df=[8], smooth="cr", shift=1, name="history")
    m.initialize(method="random", random_seed=2046, compute_ci=False)
    m.alpha, m.beta = 1.0, 0.0
    m.y["train"] = y[dims[0] - 1:].cpu().numpy()
    m._dirty = True
    m._sync()
    return m


def _zero_params(m):
    p = {name: np.zeros_like(m.p0[name]) for name in m.filter_names}
    p["intercept"] = {k: 0.0 for k in m.p0["intercept"]}
    return p


def _columns(m):
    """Design columns (in the Gram kernel's concatenated, padded layout) of every coefficient."""
    cols = []
    for name in m.filter_names:
        c0 = m._slot_col0[m._slot_of[name]]
        cols += list(range(c0, c0 + m.n_features[name]))
    return np.asarray(cols)


def test_config3_full_size_properties():
    if _free_gb() < 60:
        pytest.skip("needs ~45 GB of free HBM")
    m = _build(25, [25, 20, 15], [8, 6, 6], "poisson", "softplus", history=True)
    n = m._n_trim["train"]
    assert n == (1 << 25) - 24 and m.n_features == {"stimulus": 288, "history": 8}
    y = m.y["train"].astype(np.float64)

    # (1) zero parameters: r = softplus(0) + 1e-7 for every sample -> closed-form loss
    p0 = _zero_params(m)
    loss, g = m.value_and_grad(p0, "train", penalize=False)
    r32 = np.float32(np.log(np.float32(1) + np.exp(np.float32(0)))) + np.float32(1e-7)
    want = -float(np.log(r32)) * y.sum() + n * float(r32)
    assert abs(loss - want) < 1e-6 * abs(want)
    # (2) and the gradient is 0.5 * (XS^T 1 - XS^T y / r): compare with the fp64 Gram kernel
    _, colsum, xty = m._dm.gram("train")
    cols = _columns(m)
    gb = np.concatenate([g[name].ravel() for name in m.filter_names])
    want_g = 0.5 * (colsum[cols] - xty[cols] / float(r32))
    assert np.abs(gb - want_g).max() < 1e-5 * np.abs(want_g).max()
    assert abs(g["intercept"]["global"] - 0.5 * (n - y.sum() / float(r32))) < 1e-5 * n

    # (3) bit-reproducible (no atomics anywhere)
    loss2, g2 = m.value_and_grad(p0, "train", penalize=False)
    assert loss2 == loss and all(np.array_equal(g[k], g2[k]) for k in m.filter_names)

    # (4) directional finite difference at the benchmark's initial parameters
    p = {name: (0.1 * np.random.default_rng(2046 + i).standard_normal(m.p0[name].shape)).astype(np.float32)
         for i, name in enumerate(m.filter_names)}
    p["intercept"] = {k: 0.0 for k in m.p0["intercept"]}
    _, g = m.value_and_grad(p, "train", penalize=False)
    rng = np.random.default_rng(1)
    d = {name: rng.standard_normal(p[name].shape).astype(np.float32) for name in m.filter_names}
    eps = 1e-2
    lp = m.cost({**{k: p[k] + eps * d[k] for k in d}, "intercept": p["intercept"]}, "train", penalize=False)
    lm = m.cost({**{k: p[k] - eps * d[k] for k in d}, "intercept": p["intercept"]}, "train", penalize=False)
    fd = (lp - lm) / (2 * eps)
    an = sum(float((g[k] * d[k]).sum()) for k in d)
    assert abs(fd - an) < 2e-3 * abs(an)

    # (5) a few fit iterations: finite, decreasing, reproducible traces
    m.p0 = p
    m.fit(num_iters=5, step_size=0.01, beta=0.01, verbose=0, tolerance=0)
    first = m.cost_train.copy()
    assert np.all(np.isfinite(first)) and np.all(np.diff(first) < 0)
    m.fit(num_iters=5, step_size=0.01, beta=0.01, verbose=0, tolerance=0)
    assert np.array_equal(first, m.cost_train)


def test_config5_full_size_properties():
    if _free_gb() < 110:
        pytest.skip("needs ~90 GB of free HBM")
    m = _build(24, [40, 32, 32], [12, 10, 10], "gaussian", "none", history=False)
    n = m._n_trim["train"]
    assert n == (1 << 24) - 39 and m.n_features == {"stimulus": 1200}
    y = m.y["train"].astype(np.float64)
    p0 = _zero_params(m)
    loss, g = m.value_and_grad(p0, "train", penalize=False)
    assert abs(loss - np.mean(y * y)) < 1e-6 * np.mean(y * y)             # r = 0: MSE = mean(y^2)
    _, colsum, xty = m._dm.gram("train")
    want_g = -2.0 / n * xty[_columns(m)]                                  # d/db mean((y - XS b)^2) at b = 0
    assert np.abs(g["stimulus"].ravel() - want_g).max() < 1e-5 * np.abs(want_g).max()
    assert abs(g["intercept"]["global"] + 2.0 * y.mean()) < 1e-5 * max(1.0, abs(2.0 * y.mean()))
    loss2, g2 = m.value_and_grad(p0, "train", penalize=False)
    assert loss2 == loss and np.array_equal(g["stimulus"], g2["stimulus"])


def _build_lnln(log2n, subunits):
    """BASELINE config 4's train design: 4 softplus subunits on the [25,20,15]/[8,6,6] block + history."""
    from rfest_b200 import GLM
    from rfest_b200.engine import Engine
    from rfest_b200.synth import GaussianStimulus, TensorRows, poisson_response

    dims, df = [25, 20, 15], [8, 6, 6]
    n_raw = 1 << log2n
    eng = Engine.get()
    stim = GaussianStimulus(n_raw, dims[1:], seed=2046, device=eng.device)
    y = poisson_response(eng, stim, dims, 0, n_raw)
    m = GLM(distr="poisson", output_nonlinearity="softplus")
    m.add_design_matrix(stim, dims=dims, df=df, smooth="cr", filter_nonlinearity="softplus", name="stimulus")
    m.add_design_matrix(TensorRows(y, 0, n_raw), dims=[20], df=[8], smooth="cr", shift=1, name="history")
    m.initialize(num_subunits=subunits, method="random", random_seed=2046, compute_ci=False)
    m.alpha, m.beta = 0.5, 0.0
    m.y["train"] = y[dims[0] - 1:].cpu().numpy()
    m._dirty = True
    m._sync()
    return m


def test_config4_full_size_properties():
    """The subunit sweep at BASELINE config 4's size (2^24 samples, 4 subunits + history)."""
    if _free_gb() < 40:
        pytest.skip("needs ~25 GB of free HBM")
    from rfest_b200.engine import Engine

    K = 4
    m = _build_lnln(24, K)
    n = m._n_trim["train"]
    assert n == (1 << 24) - 24 and len(m.filter_names) == K + 1
    y = m.y["train"].astype(np.float64)

    # (1) zero parameters: every subunit gives softplus(0) + 1e-7, the history filter 0 -> closed forms
    p0 = _zero_params(m)
    loss, g = m.value_and_grad(p0, "train", penalize=False)
    f32 = np.float32
    sp0 = f32(np.log(f32(1) + np.exp(f32(0)))) + f32(1e-7)
    u = f32(0)
    for _ in range(K):
        u = f32(u + sp0)
    e = np.exp(u)
    r32 = f32(np.log(f32(1) + e)) + f32(1e-7)
    want = -float(np.log(r32)) * y.sum() + n * float(r32)
    assert abs(loss - want) < 1e-6 * abs(want)
    # dL/du_t = (1 - y_t / r) sigmoid(u); subunits see it through phi'(0) = 1/2, the history filter directly
    _, colsum, xty = m._dm.gram("train")
    sig = float(e / (f32(1) + e))
    base = sig * (colsum - xty / float(r32))
    c_stim = m._slot_col0[m._slot_of["stimulus_s0"]]
    c_hist = m._slot_col0[m._slot_of["history"]]
    for k in range(K):
        want_g = 0.5 * base[c_stim:c_stim + 288]
        assert np.abs(g[f"stimulus_s{k}"].ravel() - want_g).max() < 1e-5 * np.abs(want_g).max()
    want_h = base[c_hist:c_hist + 8]
    assert np.abs(g["history"].ravel() - want_h).max() < 1e-5 * np.abs(want_h).max()
    assert abs(g["intercept"]["global"] - sig * (n - y.sum() / float(r32))) < 1e-5 * n

    # (2) bit-reproducible
    loss2, g2 = m.value_and_grad(p0, "train", penalize=False)
    assert loss2 == loss and all(np.array_equal(g[k], g2[k]) for k in m.filter_names)

    # (3) directional finite difference at distinct random subunits
    p = {name: (0.1 * np.random.default_rng(2046 + i).standard_normal(m.p0[name].shape)).astype(np.float32)
         for i, name in enumerate(m.filter_names)}
    p["intercept"] = {k: 0.0 for k in m.p0["intercept"]}
    lossp, g = m.value_and_grad(p, "train", penalize=False)
    rng = np.random.default_rng(1)
    d = {name: rng.standard_normal(p[name].shape).astype(np.float32) for name in m.filter_names}
    eps = 1e-2
    lp = m.cost({**{k: p[k] + eps * d[k] for k in d}, "intercept": p["intercept"]}, "train", penalize=False)
    lm = m.cost({**{k: p[k] - eps * d[k] for k in d}, "intercept": p["intercept"]}, "train", penalize=False)
    fd = (lp - lm) / (2 * eps)
    an = sum(float((g[k] * d[k]).sum()) for k in d)
    assert abs(fd - an) < 2e-3 * abs(an)

    # (4) the row-per-warp kernels (the mapping it replaced) agree at the same parameters
    del m
    Engine.get().set_option("sweep_sub", 0)
    try:
        m0 = _build_lnln(24, K)
        loss0, g0 = m0.value_and_grad(p, "train", penalize=False)
    finally:
        Engine.get().set_option("sweep_sub", 3)
    assert abs(loss0 - lossp) < 1e-7 * abs(lossp)
    for k in g0:
        if k != "intercept":
            assert np.abs(g0[k] - g[k]).max() < 1e-5 * np.abs(g0[k]).max()
