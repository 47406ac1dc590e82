"""The JAX part of the path is 'parity unpinned' (no JAX in this image); these
tests are what pins the oracle's restatement of it instead:

* analytic gradient vs central finite differences and vs torch.autograd on the
  cost of rfest/GLM/_glm.py:539-614 + rfest/loss.py:4-23 written with torch ops;
* Adam recurrence vs torch.optim.Adam (algebraically the same update);
* the recovery thresholds of the reference's tests/test_glm.py.
"""

import numpy as np
import pytest
import torch

from oracle.glm import OracleGLM, adam_step, fnl, fnl_vjp

import helpers


def _model(distr, out_nl, filt_nl="none", subunits=1, alpha=0.5, beta=0.01, dtype=np.float64,
           n=300, seed=3, dims=(6, 4, 3), df=(3, 3, 3)):
    rng = np.random.default_rng(seed)
    stim = rng.standard_normal((n,) + tuple(dims[1:]))
    y = rng.poisson(0.6, n).astype(np.float64) if distr == "poisson" else rng.standard_normal(n)
    m = OracleGLM(distr=distr, output_nonlinearity=out_nl, dtype=dtype)
    m.add_design_matrix(stim, dims=list(dims), df=list(df), smooth="cr",
                        filter_nonlinearity=filt_nl, name="stimulus")
    m.add_design_matrix(y, dims=[5], df=[3], smooth="cr", shift=1, name="history")
    m.initialize(num_subunits=subunits, method="random", random_seed=11)
    # small, distinct parameters and non-zero intercepts
    k = 0
    for name in m.filter_names:
        k += 1
        m.p0[name] = (0.05 * k * np.random.default_rng(100 + k).standard_normal(m.p0[name].shape)
                      ).astype(dtype)
        m.p0["intercept"][name] = 0.01 * k
    m.p0["intercept"]["global"] = -0.2
    m.alpha, m.beta = alpha, beta
    m.y["train"] = y[m.burn_in:].astype(dtype)
    return m


def _flatten(p, names):
    vec = [np.asarray(p[n], dtype=np.float64).ravel() for n in names]
    vec += [np.atleast_1d(np.float64(p["intercept"][k])) for k in sorted(p["intercept"])]
    return np.concatenate(vec)


def _unflatten(vec, like, names):
    p, pos = {}, 0
    for n in names:
        k = np.asarray(like[n]).size
        p[n] = vec[pos:pos + k].reshape(np.asarray(like[n]).shape)
        pos += k
    p["intercept"] = {}
    for k in sorted(like["intercept"]):
        p["intercept"][k] = vec[pos]
        pos += 1
    return p


CASES = [("poisson", "softplus", "none", 1), ("poisson", "exponential", "none", 1),
         ("gaussian", "none", "none", 1), ("gaussian", "exponential", "none", 1),
         ("poisson", "softplus", "softplus", 3), ("gaussian", "none", "tanh", 2),
         ("poisson", "softplus", "sigmoid", 1), ("gaussian", "relu", "leaky_relu", 1)]


@pytest.mark.parametrize("distr,out_nl,filt_nl,subunits", CASES)
def test_gradient_finite_differences(distr, out_nl, filt_nl, subunits):
    m = _model(distr, out_nl, filt_nl, subunits)
    names = m.filter_names
    loss, g = m.value_and_grad(m.p0, "train")
    assert np.isclose(loss, m.cost(m.p0, "train"), rtol=1e-14)
    gvec = _flatten(g, names)
    x0 = _flatten(m.p0, names)
    rng = np.random.default_rng(0)
    for _ in range(6):
        d = rng.standard_normal(x0.size)
        d /= np.linalg.norm(d)
        h = 1e-6
        fp = m.cost(_unflatten(x0 + h * d, m.p0, names), "train")
        fm = m.cost(_unflatten(x0 - h * d, m.p0, names), "train")
        fd = (fp - fm) / (2 * h)
        assert abs(fd - gvec @ d) <= 2e-6 * max(1.0, abs(fd))


def _torch_fnl(x, kind):
    if kind == "softplus":
        return torch.log(1 + torch.exp(x)) + 1e-7
    if kind == "exponential":
        return torch.exp(x)
    if kind == "sigmoid":
        return 1 / (1 + torch.exp(-x))
    if kind == "tanh":
        return torch.tanh(x)
    if kind == "relu":
        return torch.where(x > 0.0, x, torch.full_like(x, 1e-7))
    if kind == "leaky_relu":
        return torch.where(x > 0.0, x, x * 0.01)
    return x


@pytest.mark.parametrize("distr,out_nl,filt_nl,subunits", CASES)
def test_gradient_torch_autograd(distr, out_nl, filt_nl, subunits):
    m = _model(distr, out_nl, filt_nl, subunits)
    names = m.filter_names
    tp = {n: torch.tensor(np.asarray(m.p0[n]), dtype=torch.float64, requires_grad=True) for n in names}
    ti = {k: torch.tensor(float(v), dtype=torch.float64, requires_grad=True)
          for k, v in m.p0["intercept"].items()}
    outs = []
    for name in m.X["train"]:
        XS = torch.tensor(m.XS["train"][name])
        outs.append(_torch_fnl((XS @ tp[name]).squeeze(-1) + ti[name], m.filter_nonlinearity[name]))
    r = _torch_fnl(torch.stack(outs).sum(0) + ti["global"], out_nl)
    y = torch.tensor(m.y["train"])
    if distr == "gaussian":
        loss = torch.mean((y - r) ** 2)
    else:
        rr = torch.where(r != float("inf"), r, torch.zeros_like(r))
        rr = torch.maximum(rr, torch.full_like(rr, 1e-20))
        loss = -torch.log(rr / 1.0) @ y + rr.sum()
    w = torch.hstack([tp[n].flatten() for n in names])
    loss = loss + m.beta * ((1 - m.alpha) * torch.linalg.norm(w, 2) + m.alpha * torch.linalg.norm(w, 1))
    loss.backward()

    o_loss, g = m.value_and_grad(m.p0, "train")
    assert np.isclose(loss.item(), float(o_loss), rtol=1e-13)
    for n in names:
        np.testing.assert_allclose(g[n], tp[n].grad.numpy(), rtol=1e-10, atol=1e-12)
    for k in ti:
        np.testing.assert_allclose(g["intercept"][k], ti[k].grad.numpy(), rtol=1e-10, atol=1e-12)


def test_fnl_vjp_matches_numeric():
    x = np.linspace(-3, 3, 40)                     # no sample on the kinks at 0
    for kind in ["softplus", "exponential", "sigmoid", "tanh", "leaky_relu", "none", "softmax"]:
        ct = np.cos(x)
        h = 1e-6
        num = np.array([(ct @ (fnl(x + h * e, kind) - fnl(x - h * e, kind))) / (2 * h)
                        for e in np.eye(x.size)])
        np.testing.assert_allclose(fnl_vjp(x, ct, kind), num, rtol=1e-6, atol=1e-8)


def test_softplus_is_the_naive_form():
    x = np.array([-30.0, -17.0, 0.0, 5.0], dtype=np.float32)
    out = fnl(x, "softplus")
    assert out.dtype == np.float32
    assert out[0] == np.float32(1e-7) and out[1] == np.float32(1e-7)     # log(1 + tiny) == 0 in fp32
    assert np.isclose(out[2], np.log(2.0) + 1e-7, rtol=1e-6)


def test_adam_matches_torch():
    rng = np.random.default_rng(5)
    x = rng.standard_normal(17)
    tx = torch.tensor(x.copy(), requires_grad=True)
    opt = torch.optim.Adam([tx], lr=0.03, betas=(0.9, 0.999), eps=1e-8)
    m = np.zeros_like(x)
    v = np.zeros_like(x)
    for i in range(50):
        g = np.sin(x * (i + 1)) + 0.1 * x
        tx.grad = torch.tensor(np.sin(tx.detach().numpy() * (i + 1)) + 0.1 * tx.detach().numpy())
        opt.step()
        x, m, v = adam_step(i, x, g, m, v, 0.03, np.float64)
        np.testing.assert_allclose(x, tx.detach().numpy(), rtol=1e-12, atol=1e-14)


def test_adam_first_step_known_answer():
    # one step from zero state: x - lr * g / (|g| + eps)
    x, m, v = adam_step(0, np.array([1.0, -2.0]), np.array([0.5, -4.0]), np.zeros(2), np.zeros(2),
                        0.1, np.float64)
    np.testing.assert_allclose(x, [1.0 - 0.1, -2.0 + 0.1], rtol=1e-7)
    np.testing.assert_allclose(m, [0.05, -0.4], rtol=1e-12)
    np.testing.assert_allclose(v, [0.00025, 0.016], rtol=1e-12)


def _fit_like_reference(Xy_train, dims, Xy_dev=None, history=False, out_nl="none", subunits=1,
                        cls=OracleGLM, **kw):
    """`_fit_glm` of the reference's tests/test_glm.py:14-73."""
    df = helpers.df_rule(dims)
    model = cls(distr="gaussian", output_nonlinearity=out_nl, **kw)
    model.add_design_matrix(Xy_train[0], dims=list(dims), df=df, smooth="cr", kind="train",
                            filter_nonlinearity="none", name="stimulus")
    if Xy_dev is not None:
        model.add_design_matrix(Xy_dev[0], dims=list(dims), df=df, name="stimulus", kind="dev")
    if history:
        model.add_design_matrix(Xy_train[1], dims=[5], df=[3], smooth="cr", kind="train",
                                filter_nonlinearity="none", name="history")
        if Xy_dev is not None:
            model.add_design_matrix(Xy_dev[1], dims=[5], df=[3], kind="dev", name="history")
    fit_y = {"train": Xy_train[1]}
    if Xy_dev is not None:
        fit_y["dev"] = Xy_dev[1]
    model.initialize(num_subunits=subunits, dt=1.0, method="mle", random_seed=42,
                     compute_ci=False, y=Xy_train[1])
    model.fit(y=fit_y, num_iters=300, verbose=0, step_size=0.03, beta=0.001, metric="mse")
    return model


def test_recovery_3d_stim():
    w_true, X, y, dims = helpers.data_3d_stim()
    model = _fit_like_reference((X, y), dims)
    assert model.score(X, y, metric="corrcoef") > 0.4
    assert np.mean((helpers.uvec(model.w["opt"]["stimulus"]).ravel()
                    - helpers.uvec(w_true.flatten())) ** 2) < 0.01
    assert len(model.cost_train) <= 300 and model.train_stop in ("maxiter_stop", "train_stop")


def test_recovery_train_dev_test_history():
    w_true, X, y, dims = helpers.data_3d_stim()
    tr, dv, te = helpers.split(X, y, 0.6, 0.2)
    model = _fit_like_reference(tr, dims, Xy_dev=dv, history=True)
    assert model.return_model == "best_dev_cost"
    assert model.score({"stimulus": tr[0], "history": tr[1]}, tr[1]) > 0.4
    assert model.score({"stimulus": dv[0], "history": dv[1]}, dv[1]) > 0.3
    assert model.score({"stimulus": te[0], "history": te[1]}, te[1]) > 0.2
    assert np.mean((helpers.uvec(model.w["opt"]["stimulus"]).ravel()
                    - helpers.uvec(w_true.flatten())) ** 2) < 0.01
    assert model.best_iteration == int(np.argmin(model.cost_dev))


def test_recovery_two_subunits_and_exponential():
    w_true, X, y, dims = helpers.data_3d_stim()
    model = _fit_like_reference((X, y), dims, subunits=2)
    w = model.w["opt"]["stimulus_s0"] + model.w["opt"]["stimulus_s1"]
    assert np.mean((helpers.uvec(w).ravel() - helpers.uvec(w_true.flatten())) ** 2) < 0.01
    model = _fit_like_reference((X, y), dims, out_nl="exponential")
    assert model.score(X, y, metric="corrcoef") > 0.4


def test_fp32_mode_tracks_fp64():
    stim, y, _ = helpers.lnp_data(3000, (8, 5, 4), seed=9)
    traces = {}
    for dt in (np.float64, np.float32):
        m = OracleGLM(distr="poisson", output_nonlinearity="softplus", dtype=dt)
        m.add_design_matrix(stim, dims=[8, 5, 4], df=[4, 3, 3], smooth="cr", name="stimulus")
        m.add_design_matrix(y, dims=[6], df=[4], smooth="cr", shift=1, name="history")
        m.initialize(method="random", random_seed=1)
        for n in m.filter_names:
            m.p0[n] = (0.1 * m.p0[n]).astype(dt)
        m.fit(y=y, num_iters=40, step_size=0.05, beta=0.01, alpha=1, metric="corrcoef",
              verbose=0, tolerance=0)
        traces[dt] = (m.cost_train.copy(), np.asarray(m.p["opt"]["stimulus"], np.float64))
    c64, b64 = traces[np.float64]
    c32, b32 = traces[np.float32]
    assert np.max(np.abs(c32 - c64) / np.abs(c64)) < 1e-5
    assert np.linalg.norm(b32 - b64) / np.linalg.norm(b64) < 1e-3
    assert np.all(np.diff(c64[:10]) < 0)


def test_trace_semantics():
    """cost_train[i] is evaluated BEFORE update i; metric/cost_dev AFTER."""
    stim, y, _ = helpers.lnp_data(1500, (6, 3, 3), seed=4)
    m = OracleGLM(distr="poisson", output_nonlinearity="softplus")
    m.add_design_matrix(stim[:1000], dims=[6, 3, 3], df=[3, 3, 3], smooth="cr", name="stimulus")
    m.add_design_matrix(stim[1000:], name="stimulus", kind="dev")
    m.initialize(method="random", random_seed=3)
    m.p0["stimulus"] = 0.05 * m.p0["stimulus"]
    m.fit(y={"train": y[:1000], "dev": y[1000:]}, num_iters=5, step_size=0.01, beta=0.01,
          verbose=0, tolerance=0, metric="r2")
    assert np.isclose(m.cost_train[0], m.cost(m.p0, "train"))
    for i in range(4):
        assert np.isclose(m.cost_train[i + 1], m.cost(m.all_params[i], "train"))
    for i in range(5):
        assert np.isclose(m.cost_dev[i], m.cost(m.all_params[i], "dev", penalize=False))
        rp = m.forwardpass(m.all_params[i], "train")
        assert np.isclose(m.metric_train[i], m.compute_score(m.y["train"], rp, "r2"))
    assert np.array_equal(m.y_pred["opt"]["train"], m.forwardpass(m.all_params[-1], "train"))


def test_response_band_restatement():
    """_get_response_variance (_glm.py:1166-1232): the reference's band lines use the raw lagged design
    X @ w, the centre line XS @ b.  w = S b makes them the same number -- the identity the engine's
    lag-free band kernel rests on -- and a monotone nonlinearity orders lower <= centre <= upper."""
    rng = np.random.default_rng(5)
    n = 400
    stim = rng.standard_normal((n, 4, 3))
    y = rng.poisson(0.5, n).astype(np.float64)
    m = OracleGLM(distr="poisson", output_nonlinearity="softplus", dtype=np.float64)
    m.add_design_matrix(stim, dims=[6, 4, 3], df=[3, 3, 3], smooth="cr", name="stimulus")
    m.add_design_matrix(y, dims=[5], df=[3], smooth="cr", shift=1, name="history")
    m.initialize(method="random", random_seed=4, compute_ci=True)
    m.fit(y=y, num_iters=10, step_size=0.05, beta=0.01, verbose=0)
    for name in m.filter_names:
        lin_raw = m.X["train"][name] @ m.w["opt"][name]
        lin_xs = m.XS["train"][name] @ m.b["opt"][name]
        assert np.abs(lin_raw - lin_xs).max() < 1e-12 * max(1.0, np.abs(lin_xs).max())
        # the quadratic form behind y_se, written per row
        V = m.V["opt"][name]
        M = m.XS["train"][name]
        q = np.einsum("ti,ij,tj->t", M, V, M)
        assert np.allclose(np.sqrt(q), np.sqrt(np.sum(M @ V * M, 1)), rtol=1e-12)
    mid, up, lo = (np.asarray(d["opt"]["train"]) for d in (m.y_pred, m.y_pred_upper, m.y_pred_lower))
    assert mid.shape == up.shape == lo.shape == (n - 5,)
    assert np.all(lo <= mid + 1e-12) and np.all(mid <= up + 1e-12) and np.any(up - lo > 1e-6)
    assert np.allclose(mid, m.forwardpass(m.p["opt"], "train"))
