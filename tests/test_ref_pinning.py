"""Pinning the CPU oracle's GLM half (`oracle/glm.py`) to the REFERENCE'S OWN CODE.

`rfest/GLM/_glm.py` and `rfest/loss.py` are JAX code and JAX cannot be installed here, so
they are executed unmodified through a test-only `jax` package (`oracle/_jaxshim/`: NumPy, a
reverse-mode tape, the published Adam) by `oracle/ref_glm.py`.  Three layers:

1. the shim itself: its autodiff against `torch.autograd` and finite differences, its Adam
   against `torch.optim.Adam`;
2. live (build container only): on every case of `tests/glm_cases.py` the oracle equals the
   shim-run reference to <= 1e-10 relative in float64 (measured <= 3e-12, mostly 1e-15) and
   to float32 rounding in float32; the committed fixtures are what the live reference
   produces; the reference's own `tests/test_glm.py` passes on the shim-run reference;
3. everywhere (GPU box included): the oracle reproduces the committed fixtures
   `tests/golden/ref_glm_golden.npz` (written by `oracle/make_golden_glm.py`).
"""

import importlib.util
import os
import sys

import numpy as np
import pytest
import torch

from oracle import ref_glm
from oracle.glm import OracleGLM

import glm_cases as gc

GOLDEN_GLM = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_glm_golden.npz")
F64_RTOL = 1e-10
F32_RTOL = 2e-6

needs_reference = pytest.mark.skipif(not ref_glm.available(), reason="reference tree not present (GPU box)")


@pytest.fixture(scope="module")
def ref():
    GLM, loss, backend = ref_glm.load()
    return GLM, loss, backend


@pytest.fixture(scope="module")
def golden_glm():
    return np.load(GOLDEN_GLM)


def _rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    assert a.shape == b.shape, (a.shape, b.shape)
    live = np.isfinite(a) | np.isfinite(b)
    assert np.array_equal(np.isnan(a), np.isnan(b))
    if not live.any():
        return 0.0
    return float(np.max(np.abs(a[live] - b[live])) / max(np.max(np.abs(b[live])), 1e-300))


def _compare(got, want, rtol, skip=()):
    for key, w in want.items():
        assert key in got, key
        g = got[key]
        if w.dtype.kind in "US" or w.dtype.kind == "i":
            assert np.array_equal(g, w), (key, g, w)
            continue
        if any(key.startswith(s) for s in skip):
            continue
        assert _rel(g, w) <= rtol, (key, _rel(g, w))


def _softmax_skip(name):
    """A softmax does not depend on a shift of its argument: the gradient of the intercepts below it is rounding
    noise (1e-17 of the cost in float64), which Adam's normalisation turns into ~1e-9 of drift.  Those intercepts
    have no effect on any prediction and are not compared; everything they could influence is."""
    c = gc.CASES[name]
    return ("last.intercept", "opt.intercept") if "softmax" in (c["out_nl"], c["filt_nl"]) else ()


def _golden_case(z, name, tag):
    prefix = f"{name}/{tag}/"
    return {k[len(prefix):]: z[k] for k in z.files if k.startswith(prefix)}


# ---------------------------------------------------------------- 1. the shim itself
def _shim():
    ref_glm._ensure_jax()
    import jax
    import jax.numpy as jnp
    from jax.example_libraries import optimizers

    return jax, jnp, optimizers


@pytest.mark.parametrize("distr,out_nl,filt_nl", [
    ("poisson", "softplus", "none"), ("poisson", "exponential", "tanh"), ("gaussian", "none", "softplus"),
    ("gaussian", "relu", "leaky_relu"), ("poisson", "softplus", "sigmoid"), ("gaussian", "softmax", "none"),
    ("poisson", "softplus", "softmax")])
def test_shim_autodiff_matches_torch_autograd(distr, out_nl, filt_nl):
    """The cost of _glm.py:539-614 + loss.py:4-23 written with jnp ops (differentiated by the
    shim's tape) and with torch ops (differentiated by torch.autograd)."""
    jax, jnp, _ = _shim()
    if not getattr(jax, "IS_RFEST_B200_TEST_SHIM", False):
        pytest.skip("real jax present")
    rng = np.random.default_rng(4)
    n, pa, pb = 257, 7, 5
    A, B = rng.standard_normal((n, pa)), rng.standard_normal((n, pb))
    y = rng.poisson(0.7, n).astype(np.float64) if distr == "poisson" else rng.standard_normal(n)
    p = {"a": 0.2 * rng.standard_normal((pa, 1)), "b": 0.2 * rng.standard_normal((pb, 1)),
         "intercept": {"a": 0.03, "b": -0.02, "global": 0.1}}
    alpha, beta = 0.4, 0.05

    def nl(x, kind, lib):
        if kind == "softplus":
            return lib.log(1 + lib.exp(x)) + 1e-7
        if kind == "exponential":
            return lib.exp(x)
        if kind == "softmax":
            z = lib.exp(x)
            return z / z.sum()
        if kind == "sigmoid":
            return 1 / (1 + lib.exp(-x))
        if kind == "tanh":
            return lib.tanh(x)
        if kind == "relu":
            return lib.where(x > 0.0, x, 1e-7 + 0 * x)
        if kind == "leaky_relu":
            return lib.where(x > 0.0, x, x * 0.01)
        return x

    def cost_jnp(p):
        fa = nl(jnp.squeeze(A @ p["a"]) + p["intercept"]["a"], filt_nl, jnp)
        fb = nl(jnp.squeeze(B @ p["b"]) + p["intercept"]["b"], "none", jnp)
        r = nl(jnp.array([fa, fb]).sum(0) + p["intercept"]["global"], out_nl, jnp)
        if distr == "poisson":
            r = jnp.maximum(jnp.where(r != jnp.inf, r, 0.0), 1e-20)
            loss = -jnp.log(r / 1.0) @ y + jnp.sum(r)
        else:
            loss = jnp.mean((y - r) ** 2)
        w = jnp.hstack([p["a"].flatten(), p["b"].flatten()])
        loss += beta * ((1 - alpha) * jnp.linalg.norm(w, 2) + alpha * jnp.linalg.norm(w, 1))
        return jnp.squeeze(loss)

    val, g = jax.value_and_grad(cost_jnp)(p)

    tp = {"a": torch.tensor(p["a"], requires_grad=True), "b": torch.tensor(p["b"], requires_grad=True)}
    ti = {k: torch.tensor(v, dtype=torch.float64, requires_grad=True) for k, v in p["intercept"].items()}
    tA, tB, ty = torch.tensor(A), torch.tensor(B), torch.tensor(y)
    fa = nl((tA @ tp["a"]).squeeze() + ti["a"], filt_nl, torch)
    fb = (tB @ tp["b"]).squeeze() + ti["b"]
    r = nl(fa + fb + ti["global"], out_nl, torch)
    if distr == "poisson":
        r = torch.clamp(r, min=1e-20)
        tl = -(torch.log(r) @ ty) + r.sum()
    else:
        tl = ((ty - r) ** 2).mean()
    w = torch.cat([tp["a"].flatten(), tp["b"].flatten()])
    tl = tl + beta * ((1 - alpha) * torch.linalg.norm(w, 2) + alpha * torch.linalg.norm(w, 1))
    tl.backward()
    assert abs(float(val) - float(tl.detach())) <= 1e-12 * abs(float(tl.detach()))
    for k in ("a", "b"):
        assert _rel(g[k], tp[k].grad.numpy()) < 1e-11
    for k in ti:
        assert abs(float(g["intercept"][k]) - float(ti[k].grad)) <= 1e-11 * max(1.0, abs(float(ti[k].grad)))


def test_shim_promotion_and_cotangent_dtypes():
    """float32 parameters with a float64 intercept leaf: the forward promotes to float64 (JAX
    under x64 = NumPy NEP 50), gradients come back in each leaf's own dtype; Python-float leaves
    are weak (float32 arithmetic, float64 gradient)."""
    jax, jnp, _ = _shim()
    if not getattr(jax, "IS_RFEST_B200_TEST_SHIM", False):
        pytest.skip("real jax present")
    rng = np.random.default_rng(0)
    A = rng.standard_normal((50, 4)).astype(np.float32)
    seen = {}

    def f(p):
        u = jnp.squeeze(A @ p["b"]) + p["c"]
        seen["dtype"] = u.dtype
        return jnp.sum(jnp.exp(u).astype(np.float32))

    for c, want in ((0.25, np.float32), (np.float64(0.25), np.float64)):
        val, g = jax.value_and_grad(f)({"b": rng.standard_normal((4, 1)).astype(np.float32), "c": c})
        assert seen["dtype"] == want
        assert g["b"].dtype == np.float32 and g["b"].shape == (4, 1)
        assert np.asarray(g["c"]).dtype == np.float64 and np.ndim(g["c"]) == 0
        assert np.asarray(val).dtype == np.float32


def test_shim_adam_matches_torch_optim():
    _, _, optimizers = _shim()
    rng = np.random.default_rng(1)
    x0 = {"w": rng.standard_normal((6, 1)), "intercept": {"c": 0.0}}
    grads = [{"w": rng.standard_normal((6, 1)), "intercept": {"c": float(rng.standard_normal())}}
             for _ in range(25)]
    init, update, get = optimizers.adam(step_size=0.05)
    st = init(x0)
    tw = torch.tensor(x0["w"].copy(), requires_grad=True)
    tc = torch.zeros((), dtype=torch.float64, requires_grad=True)
    opt = torch.optim.Adam([tw, tc], lr=0.05, betas=(0.9, 0.999), eps=1e-8)
    for i, g in enumerate(grads):
        st = update(i, g, st)
        tw.grad, tc.grad = torch.tensor(g["w"]), torch.tensor(g["intercept"]["c"], dtype=torch.float64)
        opt.step()
    p = get(st)
    # torch divides sqrt(v) by sqrt(bias2) and adds eps afterwards: same update up to eps placement
    assert _rel(p["w"], tw.detach().numpy()) < 1e-6
    assert abs(p["intercept"]["c"] - float(tc.detach())) < 1e-6


def test_debug_nans_raises_like_jax(ref):
    GLM, _, _ = ref
    rng = np.random.default_rng(0)
    m = GLM(distr="poisson", output_nonlinearity="exponential")
    m.add_design_matrix(rng.standard_normal(300), dims=[6], df=[3], smooth="cr", name="stimulus")
    m.initialize(method="random", random_seed=0, compute_ci=False)
    m.p0["stimulus"] = np.full_like(m.p0["stimulus"], np.nan)
    with pytest.raises(FloatingPointError):
        m.fit(y=rng.poisson(0.3, 300).astype(float), num_iters=2, verbose=0, tolerance=0)


# ---------------------------------------------------------------- 2. live reference
@needs_reference
@pytest.mark.parametrize("name", sorted(gc.CASES))
def test_oracle_equals_live_reference(ref, golden_glm, name):
    GLM, _, backend = ref
    data = gc.case_data(gc.CASES[name])
    assert gc.data_digest(data) == str(golden_glm[f"{name}/digest"])
    for tag, dtype, rtol in (("f64", np.float64, F64_RTOL), ("f32", np.float32, F32_RTOL)):
        r = gc.summarize(*gc.run_case(lambda d, o: GLM(distr=d, output_nonlinearity=o, dtype=dtype),
                                      name, dtype, reference_quirks=True, data=data))
        o = gc.summarize(*gc.run_case(lambda d, o_: OracleGLM(distr=d, output_nonlinearity=o_, dtype=dtype),
                                      name, dtype, data=data))
        # float32: intercepts of 'none' filters are interchangeable directions -- Adam normalises
        # rounding noise there; they are checked through the predictions instead
        skip = () if tag == "f64" else ("last.intercept", "opt.intercept", "V.", "b_se.", "w_se.")
        skip = skip + _softmax_skip(name)
        _compare(o, r, rtol, skip=skip)
        if backend == "shim":          # the committed fixtures are what the live reference gives
            want = _golden_case(golden_glm, name, tag)
            _compare({k: v.astype(want[k].dtype) if v.dtype.kind == "f" else v for k, v in r.items()},
                     want, 1e-13 if tag == "f64" else 1e-6)


@needs_reference
def test_reference_own_glm_tests_pass_on_the_shim(ref, capsys):
    """/root/reference/tests/test_glm.py, unmodified, against the shim-run reference."""
    spec = importlib.util.spec_from_file_location("reference_test_glm",
                                                  os.path.join(ref_glm.ref_loader.REFERENCE_ROOT, "tests", "test_glm.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    ran = 0
    for k in sorted(dir(mod)):
        if k.startswith("test_"):
            getattr(mod, k)()
            ran += 1
    assert ran == 5


# ---------------------------------------------------------------- 3. committed fixtures
@pytest.mark.parametrize("name", sorted(gc.CASES))
def test_oracle_reproduces_reference_fixtures(golden_glm, name):
    """Runs everywhere: OracleGLM (float64) against what the reference produced."""
    data = gc.case_data(gc.CASES[name])
    if gc.data_digest(data) != str(golden_glm[f"{name}/digest"]):
        pytest.fail("this platform regenerates different inputs than the fixtures were made from")
    o = gc.summarize(*gc.run_case(lambda d, o_: OracleGLM(distr=d, output_nonlinearity=o_, dtype=np.float64),
                                  name, np.float64, data=data))
    _compare(o, _golden_case(golden_glm, name, "f64"), F64_RTOL, skip=_softmax_skip(name))


def test_fixture_inventory(golden_glm):
    assert sorted(golden_glm["cases"].tolist()) == sorted(gc.CASES)
    assert "backend=" in str(golden_glm["env"])
    for name in gc.CASES:
        for tag in ("f64", "f32"):
            c = _golden_case(golden_glm, name, tag)
            assert len(c["cost_train"]) >= 1 and "last.intercept" in c
