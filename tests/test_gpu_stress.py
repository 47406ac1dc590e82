"""Randomised launch geometry as a stand-in for `compute-sanitizer --tool racecheck` (closed on this pool).

Every hand-rolled pipeline of the engine (per-warp bulk-copy rings with mbarriers, the pooled stages with a
producer warp, the tensor-memory accumulators, the host->device ring with events, the persistent kernel's
grid barrier, the fused projection's Z ring) is exercised under randomly drawn warps / CTAs / stage depths / slab sizes / thread counts.
A synchronisation bug shows up as a result that depends on the geometry; here every geometry must
reproduce the default geometry's loss, gradient, predictions and short fit to fp32 summation noise, and
itself bit for bit when repeated.
"""

import os
import zlib

import numpy as np
import pytest

import helpers

pytestmark = pytest.mark.gpu

DEFAULTS = dict(sweep_warps=4, sweep_ctas=-1, sweep_stages=0, sweep_staged=1, sweep_sub=3, sweep_pool_ctas=0,
                project_slab_rows=0, h2d_threads=8, h2d_overlap=1, fit_persistent=1, fit_small_ctas=0,
                project_fused=0)


def _draw(rng):
    return dict(sweep_warps=int(rng.integers(1, 5)), sweep_ctas=int(rng.choice([-1, 0, 1, 2, 3])),
                sweep_stages=int(rng.choice([0, 2, 3, 4, 5])), sweep_staged=int(rng.choice([0, 1, 1, 1])),
                sweep_sub=int(rng.choice([0, 3, 4, 5, 6, 7])), sweep_pool_ctas=int(rng.choice([0, 1, 2, 5])),
                project_slab_rows=int(rng.choice([0, 4096, 4100, 9973])), h2d_threads=int(rng.integers(1, 9)),
                h2d_overlap=int(rng.choice([0, 1, 1])), fit_persistent=int(rng.choice([0, 1, 2, 3])),
                fit_small_ctas=int(rng.choice([0, 1, 3, 17, 148])), project_fused=int(rng.choice([0, 1])))


def _apply(opts):
    from rfest_b200.engine import Engine

    for k, v in opts.items():
        Engine.get().set_option(k, v)


MODELS = {
    # name: (dims, df, subunits, filt_nl, out_nl, distr, n)
    "lnp": ((8, 5, 4), (4, 3, 3), 1, "none", "softplus", "poisson", 20011),
    "lnln4": ((10, 8, 8), (8, 6, 6), 4, "softplus", "softplus", "poisson", 9973),
    "lg2": ((6, 3, 3), (3, 3, 3), 2, "tanh", "none", "gaussian", 12007),
}


def _run(model, opts):
    from rfest_b200 import GLM

    dims, df, subunits, filt_nl, out_nl, distr, n = MODELS[model]
    stim, y, _ = helpers.lnp_data(n, dims, seed=11)
    if distr == "gaussian":
        y = (y - y.mean()).astype(np.float32)
    n_tr = n - 3001
    _apply(opts)
    m = GLM(distr=distr, output_nonlinearity=out_nl)
    m.add_design_matrix(stim[:n_tr], dims=list(dims), df=list(df), smooth="cr", filter_nonlinearity=filt_nl,
                        name="stimulus")
    m.add_design_matrix(stim[n_tr:], name="stimulus", kind="dev")
    m.add_design_matrix(y[:n_tr], dims=[6], df=[4], smooth="cr", shift=1, name="history")
    m.add_design_matrix(y[n_tr:], name="history", kind="dev")
    m.initialize(num_subunits=subunits, method="random", random_seed=3)
    for k, name in enumerate(m.filter_names):
        m.p0[name] = (0.1 * np.random.default_rng(50 + k).standard_normal(m.p0[name].shape)).astype(np.float32)
    xs = np.asarray(m._blocks["train"][m.filter_names[0]])
    m.alpha, m.beta = 0.5, 0.01
    m.y["train"], m.y["dev"] = y[:n_tr][m.burn_in:], y[n_tr:][m.burn_in:]
    m._dirty = True
    loss, grad = m.value_and_grad(m.p0, "train")
    pred = np.asarray(m.forwardpass(m.p0, "dev"))
    m.fit(y={"train": y[:n_tr], "dev": y[n_tr:]}, num_iters=70, alpha=0.5, beta=0.01, step_size=0.02,
          tolerance=0, verbose=0)
    return dict(xs=xs, loss=loss, grad=grad, pred=pred, cost_train=m.cost_train.copy(), cost_dev=m.cost_dev.copy(),
                last={k: np.asarray(m.all_params[-1][k]).copy() for k in m.filter_names})


@pytest.mark.parametrize("model", sorted(MODELS))
def test_results_do_not_depend_on_launch_geometry(model):
    rng = np.random.default_rng(zlib.crc32(model.encode()))
    try:
        base = _run(model, DEFAULTS)
        for trial in range(int(os.environ.get("RFB_STRESS_TRIALS", "6"))):      # (soak runs: more draws)
            opts = _draw(rng)
            got = _run(model, opts)
            again = _run(model, opts) if trial < 2 else None
            tag = (model, opts)
            assert np.array_equal(got["xs"], base["xs"]), tag               # projection: bit-exact whatever the slabs
            assert abs(got["loss"] - base["loss"]) < 1e-6 * abs(base["loss"]), tag
            for name, g in base["grad"].items():
                if name != "intercept":
                    assert np.abs(got["grad"][name] - g).max() < 3e-5 * np.abs(g).max(), tag
            assert np.abs(got["pred"] - base["pred"]).max() <= 1e-5 * np.abs(base["pred"]).max(), tag
            assert np.max(np.abs(got["cost_train"] - base["cost_train"]) / np.abs(base["cost_train"])) < 2e-6, tag
            assert np.max(np.abs(got["cost_dev"] - base["cost_dev"]) / np.abs(base["cost_dev"])) < 2e-6, tag
            for name, w in base["last"].items():
                assert np.linalg.norm(got["last"][name] - w) < 2e-5 * np.linalg.norm(w), tag
            if again is not None:                                           # same geometry twice: bit for bit
                assert again["loss"] == got["loss"] and np.array_equal(again["cost_train"], got["cost_train"]), tag
                assert all(np.array_equal(again["grad"][k], got["grad"][k]) for k in got["grad"] if k != "intercept")
    finally:
        _apply(DEFAULTS)
