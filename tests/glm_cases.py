"""Named GLM fit cases, written once and run on three implementations of `rfest.GLM`:

* the reference's own class (`/root/reference/rfest/GLM/_glm.py`, unmodified, executed
  through `oracle/ref_glm.py`) -- in the build container, by `oracle/make_golden_glm.py`
  (which writes `tests/golden/ref_glm_golden.npz`) and by `tests/test_ref_pinning.py`;
* the CPU oracle (`oracle.glm.OracleGLM`) -- pinned to the two above;
* the CUDA engine (`rfest_b200.GLM`) -- `tests/test_gpu_ref_golden.py`, against the
  committed outputs of the reference.

Every case only uses the public `rfest.GLM` API (`add_design_matrix`, `initialize`, `fit`,
`predict`/`score`, the `p0` attribute), so the same function drives all three.
"""

import hashlib

import numpy as np

import helpers

# name -> spec.  Data: "lnp" = helpers.lnp_data(n, dims, seed) (Poisson spikes, float32 stimulus),
# "g3d" = helpers.data_3d_stim (the reference's tests/test_glm.py fixture, seeded).
_BASE = dict(data="lnp", n=2603, dims=(8, 5, 4), df=(4, 3, 3), seed=31, n_train=2000,
             hist=(6, 4), init="random", p0_scale=0.1, icpt=0.02, iters=60, tolerance=0,
             min_iters=300, return_model=None, compute_ci=False, test=False, verbose=0, smooth="cr",
             hist_smooth="cr")


def _case(**kw):
    c = dict(_BASE)
    c.update(kw)
    return c


CASES = {
    # the eight model configurations of tests/test_gpu_fit.py::CONFIGS
    "lnp_softplus_l1_dev": _case(distr="poisson", out_nl="softplus", filt_nl="none", subunits=1,
                                 alpha=1.0, beta=0.01, step=0.05, metric="corrcoef", history=True, dev=True),
    "lnp_exp_enet": _case(distr="poisson", out_nl="exponential", filt_nl="none", subunits=1,
                          alpha=0.5, beta=0.01, step=0.01, metric="r2", history=True, dev=False),
    "lg_l1_dev": _case(distr="gaussian", out_nl="none", filt_nl="none", subunits=1,
                       alpha=1.0, beta=0.001, step=0.03, metric="mse", history=True, dev=True),
    "lg_2sub": _case(distr="gaussian", out_nl="none", filt_nl="none", subunits=2,
                     alpha=1.0, beta=0.001, step=0.03, metric="mse", history=False, dev=False),
    "lnln_4sub_softplus_dev": _case(distr="poisson", out_nl="softplus", filt_nl="softplus", subunits=4,
                                    alpha=0.5, beta=0.01, step=0.02, metric="corrcoef", history=True, dev=True),
    "g_softplus_tanh_2sub_l2_dev": _case(distr="gaussian", out_nl="softplus", filt_nl="tanh", subunits=2,
                                         alpha=0.0, beta=0.01, step=0.02, metric="r2", history=False, dev=True),
    "lnp_sigmoid_nopen": _case(distr="poisson", out_nl="softplus", filt_nl="sigmoid", subunits=1,
                               alpha=1.0, beta=0.0, step=0.02, metric="corrcoef", history=True, dev=False),
    "g_relu_leaky_dev": _case(distr="gaussian", out_nl="relu", filt_nl="leaky_relu", subunits=1,
                              alpha=0.3, beta=0.02, step=0.01, metric="mse", history=True, dev=True),
    # BASELINE config 1 / README recipe shape (1-D, dims [25] df [8] + history [20]/[8] shift 1), shortened
    "config1_readme": _case(distr="poisson", out_nl="softplus", filt_nl="none", subunits=1, n=6000,
                            dims=(25,), df=(8,), seed=2046, n_train=4800, hist=(20, 8), alpha=1, beta=0.01,
                            step=0.1, metric="corrcoef", history=True, dev=True, iters=120, icpt=None),
    # early stopping (dev cost rises) and the 'last' rule after a dev stop
    "early_stop_dev": _case(distr="poisson", out_nl="softplus", filt_nl="none", subunits=1, n=1400,
                            dims=(6, 3, 3), df=(3, 3, 3), seed=12, n_train=1000, alpha=1.0, beta=0.0,
                            step=0.05, metric="corrcoef", history=False, dev=True, iters=400, tolerance=5,
                            min_iters=20, p0_scale=0.3, icpt=None),
    "early_stop_last": _case(distr="poisson", out_nl="softplus", filt_nl="none", subunits=1, n=1400,
                             dims=(6, 3, 3), df=(3, 3, 3), seed=12, n_train=1000, alpha=1.0, beta=0.0,
                             step=0.05, metric="corrcoef", history=False, dev=True, iters=400, tolerance=5,
                             min_iters=20, p0_scale=0.3, icpt=None, return_model="last"),
    "train_stop_flat": _case(distr="gaussian", out_nl="none", filt_nl="none", subunits=1, n=900,
                             dims=(6, 3, 3), df=(3, 3, 3), seed=5, n_train=900, alpha=1.0, beta=0.0,
                             step=0.05, metric="mse", history=False, dev=False, iters=2000, tolerance=10,
                             min_iters=30, p0_scale=0.01, icpt=None, atol=1e-3),
    "best_train_metric": _case(distr="poisson", out_nl="softplus", filt_nl="none", subunits=1, n=1400,
                               dims=(6, 3, 3), df=(3, 3, 3), seed=12, n_train=1000, alpha=1.0, beta=0.0,
                               step=0.05, metric="corrcoef", history=False, dev=True, iters=40,
                               p0_scale=0.3, icpt=None, return_model="best_train_metric"),
    # a step-size schedule (jax schedules are callables i -> step)
    "schedule": _case(distr="poisson", out_nl="softplus", filt_nl="none", subunits=1, n=1200,
                      dims=(6, 3, 3), df=(3, 3, 3), seed=2, n_train=1200, alpha=1.0, beta=0.01,
                      step="sched", metric="corrcoef", history=False, dev=False, iters=30, icpt=None),
    # least-squares initialisation (compute_mle) = the reference's tests/test_glm.py recipes
    "mle_3d": _case(data="g3d", distr="gaussian", out_nl="none", filt_nl="none", subunits=1, split=(1.0, 0.0),
                    hist=(5, 3), init="mle", alpha=1, beta=0.001, step=0.03, metric="mse", history=False,
                    dev=False, iters=300, tolerance=10, test=True),
    "mle_3d_dev_history": _case(data="g3d", distr="gaussian", out_nl="none", filt_nl="none", subunits=1,
                                split=(0.6, 0.2), hist=(5, 3), init="mle", alpha=1, beta=0.001, step=0.03,
                                metric="mse", history=True, dev=True, iters=300, tolerance=10, test=True),
    "mle_3d_exp": _case(data="g3d", distr="gaussian", out_nl="exponential", filt_nl="none", subunits=1,
                        split=(1.0, 0.0), hist=(5, 3), init="mle", alpha=1, beta=0.001, step=0.03,
                        metric="mse", history=False, dev=False, iters=300, tolerance=10, test=True),
    "mle_3d_2sub": _case(data="g3d", distr="gaussian", out_nl="none", filt_nl="none", subunits=2,
                         split=(1.0, 0.0), hist=(5, 3), init="mle", alpha=1, beta=0.001, step=0.03,
                         metric="mse", history=False, dev=False, iters=300, tolerance=10, test=True),
    # confidence intervals (compute_ci=True: V, b_se, w_se, response bands) + test-set prediction
    "ci_poisson": _case(distr="poisson", out_nl="softplus", filt_nl="none", subunits=1, n=1800,
                        dims=(6, 3, 3), df=(3, 3, 3), seed=8, n_train=1300, hist=(5, 3), alpha=1.0, beta=0.01,
                        step=0.03, metric="corrcoef", history=True, dev=True, iters=40, compute_ci=True,
                        test=True, n_test=200),
    "ci_gaussian_mle": _case(data="g3d", distr="gaussian", out_nl="none", filt_nl="none", subunits=1,
                             split=(0.6, 0.2), hist=(5, 3), init="mle", alpha=1, beta=0.001, step=0.03,
                             metric="mse", history=True, dev=True, iters=50, compute_ci=True, test=True),
    # 'softmax' (GLM.fnl, _glm.py:134-140: a normalisation over all samples) as filter and as output nonlinearity
    "softmax_filter_2sub_dev": _case(distr="poisson", out_nl="softplus", filt_nl="softmax", subunits=2,
                                     alpha=1.0, beta=0.01, step=0.02, metric="corrcoef", history=True, dev=True,
                                     iters=40),
    "softmax_output": _case(distr="poisson", out_nl="softmax", filt_nl="none", subunits=1,
                            alpha=1.0, beta=0.01, step=0.02, metric="corrcoef", history=True, dev=True, iters=40),
    # the other spline bases (rfest/splines.py:6-36,97-189): B-splines, cyclic cubic and thin-plate regression splines
    "lnp_bspline_dev": _case(distr="poisson", out_nl="softplus", filt_nl="none", subunits=1, smooth="bs",
                             hist_smooth="bs", alpha=1.0, beta=0.01, step=0.05, metric="corrcoef", history=True,
                             dev=True, iters=40),
    "lg_thin_plate_2sub": _case(distr="gaussian", out_nl="none", filt_nl="tanh", subunits=2, smooth="tp",
                                alpha=0.5, beta=0.001, step=0.03, metric="mse", history=True, dev=True, iters=40),
    "lnp_cyclic": _case(distr="poisson", out_nl="exponential", filt_nl="none", subunits=1, smooth="cc",
                        alpha=1.0, beta=0.01, step=0.01, metric="r2", history=False, dev=False, iters=40),
    "softmax_both_gaussian": _case(distr="gaussian", out_nl="softmax", filt_nl="softmax", subunits=1,
                                   alpha=1.0, beta=0.0, step=0.005, metric="mse", history=True, dev=False,
                                   iters=40),
}


def _sched(i):
    return 0.05 * 0.97 ** i


def case_data(c):
    """-> dict(train=(stim, y), dev=(stim, y) | None, test=(stim, y) | None, w_true)."""
    if c["data"] == "lnp":
        stim, y, w = helpers.lnp_data(c["n"], c["dims"], seed=c["seed"])
        if c["distr"] == "gaussian":
            y = (y - y.mean()).astype(np.float32)
        n_tr = c["n_train"] if (c["dev"] or c["test"]) else c["n"]
        n_te = c.get("n_test", 0) if c["test"] else 0
        n_dv_end = c["n"] - n_te
        out = {"train": (stim[:n_tr], y[:n_tr]), "dev": None, "test": None, "w_true": w}
        if c["dev"]:
            out["dev"] = (stim[n_tr:n_dv_end], y[n_tr:n_dv_end])
        if n_te:
            out["test"] = (stim[n_dv_end:], y[n_dv_end:])
        return out
    w_true, X, y, dims = helpers.data_3d_stim()
    tr, dv, te = helpers.split(X, y, *c["split"])
    return {"train": tr, "dev": dv if c["dev"] else None,
            "test": te if len(te[1]) else tr, "w_true": w_true, "dims": dims}


def data_digest(data):
    h = hashlib.sha256()
    for k in ("train", "dev", "test"):
        if data[k] is not None:
            for a in data[k]:
                a = np.ascontiguousarray(a)
                h.update(str(a.dtype).encode() + str(a.shape).encode() + a.tobytes())
    return h.hexdigest()[:16]


def run_case(make, name, dtype, reference_quirks=False, data=None):
    """Build, initialise and fit case `name` with the GLM class factory `make(distr, out_nl)`.
    `reference_quirks`: work around the reference's latent crash for subunit models with a dev
    set under random initialisation (`_initialize_subunits` keys its dev clean-up on `self.y`,
    which is still empty, _glm.py:359-364) by announcing the dev responses first -- the public
    `y` dictionary, nothing in the reference is modified."""
    c = CASES[name]
    data = data or case_data(c)
    dims = list(data.get("dims", c["dims"]))
    df = helpers.df_rule(dims) if c["data"] == "g3d" else list(c["df"])
    hl, hdf = c["hist"]
    hshift = 1 if c["data"] == "lnp" else 0
    m = make(c["distr"], c["out_nl"])
    tr, dv = data["train"], data["dev"]
    m.add_design_matrix(tr[0], dims=dims, df=df, smooth=c["smooth"], filter_nonlinearity=c["filt_nl"],
                        kind="train", name="stimulus")
    if dv is not None:
        m.add_design_matrix(dv[0], dims=dims, df=df, name="stimulus", kind="dev")
    if c["history"]:
        m.add_design_matrix(tr[1], dims=[hl], df=[hdf], smooth=c["hist_smooth"], shift=hshift, kind="train",
                            filter_nonlinearity="none", name="history")
        if dv is not None:
            m.add_design_matrix(dv[1], dims=[hl], df=[hdf], kind="dev", name="history")
    fit_y = {"train": tr[1]}
    if dv is not None:
        fit_y["dev"] = dv[1]
    if c["init"] == "mle":
        m.initialize(num_subunits=c["subunits"], dt=1.0, method="mle", random_seed=42,
                     compute_ci=c["compute_ci"], y=fit_y if c["compute_ci"] else tr[1])
    else:
        if reference_quirks and c["subunits"] != 1 and dv is not None:
            m.y["dev"] = np.asarray(dv[1])[m.burn_in:].astype(dtype)
        m.initialize(num_subunits=c["subunits"], method="random", random_seed=3,
                     compute_ci=c["compute_ci"])
        for k, fname in enumerate(m.filter_names):
            shape = np.asarray(m.p0[fname]).shape
            b = (c["p0_scale"] * np.random.default_rng(77 + k).standard_normal(shape)).astype(np.float32)
            m.p0[fname] = b.astype(dtype)
            if c["icpt"] is not None:
                m.p0["intercept"][fname] = c["icpt"] * (k + 1)
        if c["icpt"] is not None:
            m.p0["intercept"]["global"] = -2 * c["icpt"]
    kw = {}
    if "atol" in c:
        kw["atol"] = c["atol"]
    m.fit(y=fit_y, num_iters=c["iters"], alpha=c["alpha"], beta=c["beta"], metric=c["metric"],
          step_size=_sched if c["step"] == "sched" else c["step"], tolerance=c["tolerance"],
          verbose=c["verbose"], return_model=c["return_model"], min_iters=c["min_iters"], **kw)
    extra = {}
    if c["test"]:
        te = data["test"]
        Xt = {"stimulus": te[0]}
        if c["history"]:
            Xt["history"] = te[1]
        s, yp = m.score(Xt, te[1], metric=c["metric"], return_prediction=True)
        extra["test_score"] = float(s)
        extra["test_pred"] = np.asarray(yp, dtype=np.float64)
        if c["compute_ci"]:
            extra["test_upper"] = np.asarray(m.y_pred_upper["opt"]["test"], dtype=np.float64)
            extra["test_lower"] = np.asarray(m.y_pred_lower["opt"]["test"], dtype=np.float64)
    return m, extra


def _scalar(x):
    return float(np.asarray(x, dtype=np.float64).reshape(-1)[0])


def summarize(m, extra):
    """Everything a fit leaves behind that parity is judged on, as plain float64 arrays."""
    out = {
        "filter_names": np.array(m.filter_names),
        "cost_train": np.asarray(m.cost_train, np.float64),
        "cost_dev": np.asarray(m.cost_dev, np.float64),
        "metric_train": np.asarray(m.metric_train, np.float64),
        "metric_dev": np.asarray(m.metric_dev, np.float64),
        "best_iteration": np.int64(m.best_iteration),
        "train_stop": np.array(m.train_stop),
        "return_model": np.array(m.return_model),
    }
    for tag, p in (("last", m.all_params[-1]), ("opt", m.p["opt"]), ("p0", m.p0)):
        for fname in m.filter_names:
            out[f"{tag}.{fname}"] = np.asarray(p[fname], np.float64).ravel()
        out[f"{tag}.intercept"] = np.array([_scalar(p["intercept"][k]) for k in m.filter_names + ["global"]])
    for fname in m.filter_names:
        out[f"w.{fname}"] = np.asarray(m.w["opt"][fname], np.float64).ravel()
    for kind, r in m.y_pred["opt"].items():
        if kind in ("train", "dev"):
            out[f"y_pred.{kind}"] = np.asarray(r, np.float64).ravel()
    if getattr(m, "compute_ci", False):
        for fname in m.filter_names:
            out[f"V.{fname}"] = np.asarray(m.V["opt"][fname], np.float64)
            out[f"b_se.{fname}"] = np.asarray(m.b_se["opt"][fname], np.float64).ravel()
            out[f"w_se.{fname}"] = np.asarray(m.w_se["opt"][fname], np.float64).ravel()
        for kind in ("train", "dev"):
            if kind in m.y_pred_upper.get("opt", {}):
                out[f"upper.{kind}"] = np.asarray(m.y_pred_upper["opt"][kind], np.float64).ravel()
                out[f"lower.{kind}"] = np.asarray(m.y_pred_lower["opt"][kind], np.float64).ravel()
    for k, v in extra.items():
        out[k] = np.asarray(v)
    return out
