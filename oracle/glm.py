"""Oracle: the `rfest.GLM` model object and its fit loop, on the CPU.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  Pinned to the reference's own code:
`oracle/ref_glm.py` runs `rfest/GLM/_glm.py` + `rfest/loss.py` unmodified (through the
test-only `jax` package of `oracle/_jaxshim/`, JAX itself cannot be installed here) and
`tests/test_ref_pinning.py` holds this file to that run (<= 1e-10 relative in float64,
float32 rounding in float32) and to the committed outputs `tests/golden/ref_glm_golden.npz`.
Not pinnable without real JAX: `jax.random` draws and XLA's summation order.

Restates, with NumPy only:
  rfest/GLM/_glm.py:103-170   nonlinearities (naive softplus + 1e-7, ...)
  rfest/GLM/_glm.py:172-301   add_design_matrix (burn-in, dense projection)
  rfest/GLM/_glm.py:303-436   initialize / subunits / p0
  rfest/GLM/_glm.py:438-537   compute_mle
  rfest/GLM/_glm.py:539-614   forwardpass / cost
  rfest/GLM/_glm.py:644-836   optimize (traces, early stop, model selection)
  rfest/GLM/_glm.py:838-937   fit / _extract_opt_params
  rfest/GLM/_glm.py:939-1056  predict / score
  rfest/loss.py:4-23          Poisson NLL, MSE, elastic-net penalty
  jax.example_libraries.optimizers.adam (not in /root/reference; published
  recurrence, called at _glm.py:698-699,694-696).

Two arithmetic modes:
  dtype=np.float64  everything in double -- the reference's default dtype and
                    the "truth" the fp32 paths are measured against.
  dtype=np.float32  the reference's fp32 path with `jax_enable_x64` on (set by
                    `import rfest`): arrays are float32; Python-float intercept leaves
                    are weak-typed, so iteration 0 runs entirely in float32, but their
                    Adam state is float64 and from the first update on they are float64
                    arrays -- `XS @ b + intercept` and the nonlinearities then promote
                    to float64 until `forwardpass` casts the prediction back to float32
                    (see `_scalar`).

The autodiff gradient is restated in the order JAX's reverse mode evaluates
it (cotangent / s * e for the softplus, etc.), which only matters at the ulp
level.
"""

import time

import numpy as np

from . import design as _design
from . import metrics as _metrics
from . import splines as _splines

__all__ = ["OracleGLM", "adam_step"]

ADAM_B1 = 0.9
ADAM_B2 = 0.999
ADAM_EPS = 1e-8


# ----------------------------------------------------------------------------
# nonlinearities: value and derivative (rfest/GLM/_glm.py:103-170)
# ----------------------------------------------------------------------------
def fnl(x, kind):
    dt = x.dtype.type
    if kind == "softplus":
        return np.log(dt(1) + np.exp(x)) + dt(1e-7)               # _glm.py:124-129
    if kind == "exponential":
        return np.exp(x)                                          # _glm.py:131-132
    if kind == "softmax":
        z = np.exp(x)                                             # _glm.py:134-140
        return z / z.sum()
    if kind == "sigmoid":
        return dt(1) / (dt(1) + np.exp(-x))                       # _glm.py:142-147
    if kind == "tanh":
        return np.tanh(x)                                         # _glm.py:149-150
    if kind == "relu":
        return np.where(x > 0.0, x, dt(1e-7))                     # _glm.py:152-157
    if kind == "leaky_relu":
        return np.where(x > 0.0, x, x * dt(0.01))                 # _glm.py:159-164
    if kind == "none":
        return x                                                  # _glm.py:165-166
    raise NotImplementedError(f"Input filter nonlinearity `{kind}` is not supported.")


def fnl_vjp(x, ct, kind):
    """Cotangent of `fnl(x, kind)` pulled back to x (what JAX's autodiff emits)."""
    dt = x.dtype.type
    if kind == "softplus":
        e = np.exp(x)
        s = dt(1) + e
        return (ct / s) * e
    if kind == "exponential":
        return ct * np.exp(x)
    if kind == "softmax":
        z = np.exp(x)
        Z = z.sum()
        return (ct / Z - (ct * z).sum() / (Z * Z)) * z
    if kind == "sigmoid":
        e = np.exp(-x)
        s = dt(1) + e
        return (ct / (s * s)) * e
    if kind == "tanh":
        t = np.tanh(x)
        return ct * (dt(1) - t * t)
    if kind == "relu":
        return np.where(x > 0.0, ct, dt(0))
    if kind == "leaky_relu":
        return np.where(x > 0.0, ct, ct * dt(0.01))
    if kind == "none":
        return ct
    raise NotImplementedError(f"Input filter nonlinearity `{kind}` is not supported.")


# ----------------------------------------------------------------------------
# losses (rfest/loss.py:4-23)
# ----------------------------------------------------------------------------
def loss_neglogli(y, r):
    """Poisson negative log-likelihood with dt = 1 (loss.py:4-11)."""
    dt = r.dtype.type
    r = np.where(r != np.inf, r, dt(0.0))
    r = np.maximum(r, dt(1e-20))
    with np.errstate(divide="ignore", invalid="ignore"):
        return -np.log(r / dt(1.0)) @ y + np.sum(r)


def loss_neglogli_vjp(y, r):
    """d loss / d r (zero where the `where` / `maximum` of loss.py:5-6 cut r off)."""
    dt = r.dtype.type
    finite = r != np.inf
    r1 = np.where(finite, r, dt(0.0))
    live = finite & (r1 > dt(1e-20))
    r2 = np.maximum(r1, dt(1e-20))
    with np.errstate(divide="ignore", invalid="ignore"):
        d = -(y / r2) + dt(1)
    return np.where(live, d, dt(0))


def loss_mse(y, r):
    return np.mean((y - r) ** 2)                                  # loss.py:14-16


def loss_mse_vjp(y, r):
    dt = r.dtype.type
    return (dt(2) * (r - y)) / dt(r.shape[0])


def loss_penalty(w, alpha, beta):
    """beta * ((1 - alpha) * ||w||_2 + alpha * ||w||_1)  (loss.py:19-23).
    NB the L2 term is the plain norm, not its square."""
    dt = w.dtype.type
    l1 = np.sum(np.abs(w)) if alpha != 0.0 else dt(0.0)
    l2 = np.sqrt(np.sum(w * w)) if alpha != 1.0 else dt(0.0)
    return dt(beta) * (dt(1.0 - alpha) * l2 + dt(alpha) * l1)


def loss_penalty_vjp(w, alpha, beta):
    dt = w.dtype.type
    g = np.zeros_like(w)
    if alpha != 0.0:
        g = g + dt(beta) * dt(alpha) * np.sign(w)
    if alpha != 1.0:
        with np.errstate(divide="ignore", invalid="ignore"):
            g = g + dt(beta) * dt(1.0 - alpha) * (w / np.sqrt(np.sum(w * w)))
    return g


# ----------------------------------------------------------------------------
# Adam (jax.example_libraries.optimizers.adam; b1=0.9, b2=0.999, eps=1e-8)
# ----------------------------------------------------------------------------
def adam_step(i, x, g, m, v, step_size, dtype):
    """One Adam update of a single leaf, in `dtype` arithmetic.

        m    <- (1 - b1) g + b1 m
        v    <- (1 - b2) g^2 + b2 v
        mhat  = m / (1 - b1^(i+1)),  vhat = v / (1 - b2^(i+1))
        x    <- x - step_size * mhat / (sqrt(vhat) + eps)

    `i` is the 0-based iteration index passed to `step` (_glm.py:721).  The
    constants are Python floats in JAX, i.e. they take the leaf's dtype.
    """
    t = np.dtype(dtype).type
    m = t(1 - ADAM_B1) * g + t(ADAM_B1) * m
    v = t(1 - ADAM_B2) * (g * g) + t(ADAM_B2) * v
    # jnp.asarray(b1, m.dtype) ** (i + 1): power of the dtype-rounded constant.
    # Evaluated in double and rounded = the correctly rounded power.
    bc1 = t(1) - t(float(t(ADAM_B1)) ** (i + 1))
    bc2 = t(1) - t(float(t(ADAM_B2)) ** (i + 1))
    mhat = m / bc1
    vhat = v / bc2
    x = x - t(step_size) * mhat / (np.sqrt(vhat) + t(ADAM_EPS))
    return x, m, v


def _step_size_at(step_size, i):
    return float(step_size(i)) if callable(step_size) else float(step_size)


class OracleGLM:
    """CPU restatement of `rfest.GLM` (rfest/GLM/_glm.py:24)."""

    def __init__(self, distr="poisson", output_nonlinearity="none", dtype=np.float64):
        self.dtype = np.dtype(dtype).type
        self.distr = distr
        self.output_nonlinearity = output_nonlinearity

        self.X, self.S, self.XS, self.y, self.y_pred = {}, {}, {}, {}, {}
        self.p, self.b, self.w, self.intercept = {}, {}, {}, {}
        self.df, self.dims, self.n_features, self.shift = {}, {}, {}, {}
        self.filter_nonlinearity = {}
        self.filter_names = []
        self.burn_in = None
        self.num_subunits = None
        self.p0 = None
        self.mle_computed = False
        self.alpha = self.beta = self.metric = None

    # ---- design --------------------------------------------------------
    def add_design_matrix(self, X, dims=None, df=None, smooth=None, lag=True,
                          filter_nonlinearity="none", kind="train", name="stimulus",
                          shift=0, burn_in=None):
        """_glm.py:172-301.  Dense, like the reference: lagged matrix then `@ S`."""
        X = np.asarray(X)
        X = (X[:, np.newaxis] if X.ndim == 1 else X).astype(self.dtype)      # :231-234

        self.X.setdefault(kind, {})
        if kind == "train":                                                 # :239-245
            self.filter_nonlinearity[name] = filter_nonlinearity
            self.filter_names.append(name)
            dims = dims if type(dims) is not int else [dims]
            self.dims[name] = dims
            self.shift[name] = shift
        else:                                                               # :246-248
            dims = self.dims[name]
            shift = self.shift[name]

        if self.burn_in is None:                                            # :250-253
            self.burn_in = dims[0] - 1 if burn_in is None else burn_in

        if lag:                                                             # :255-258
            self.X[kind][name] = _design.trimmed_design(
                X, dims[0], shift, self.burn_in, dtype=self.dtype)
        else:                                                               # :259-261
            self.burn_in = 0
            self.X[kind][name] = X

        if smooth is None:
            if name in self.S:                                              # :268-272
                self.XS.setdefault(kind, {})
                self.XS[kind][name] = self.X[kind][name] @ self.S[name]
            elif kind == "test":                                            # :274-281
                self.XS.setdefault(kind, {})
                if self.num_subunits is not None and self.num_subunits > 1:
                    S = self.S["stimulus_s0"]
                else:
                    S = self.S[name]
                self.XS[kind][name] = self.X[kind][name] @ S
            elif kind == "train":                                           # :284-285
                self.n_features[name] = self.X["train"][name].shape[1]
        else:                                                               # :287-301
            self.XS.setdefault(kind, {})
            self.df[name] = df if type(df) is not int else [df]
            S = _splines.build_spline_matrix(dims, self.df[name], smooth, dtype=self.dtype)
            self.S[name] = S
            self.XS[kind][name] = self.X[kind][name] @ S
            if kind == "train":
                self.n_features[name] = self.XS["train"][name].shape[1]

    # ---- initialisation --------------------------------------------------
    def _initialize_random(self, random_seed=2046):
        """_glm.py:303-322 with NumPy's generator in place of `jax.random`
        (threefry draws cannot be reproduced without JAX): filter i draws
        `default_rng(random_seed + i).standard_normal`."""
        self.b["random"], self.w["random"], self.intercept["random"] = {}, {}, {}
        for i, name in enumerate(self.filter_names):
            self.intercept["random"][name] = 0.0
            rng = np.random.default_rng(random_seed + i)
            if name in self.S:
                n_b = self.XS["train"][name].shape[1]
                self.b["random"][name] = rng.standard_normal((n_b, 1)).astype(self.dtype)
                self.w["random"][name] = self.S[name] @ self.b["random"][name]
            else:
                n_w = self.X["train"][name].shape[1]
                self.w["random"][name] = rng.standard_normal((n_w, 1)).astype(self.dtype)
        self.intercept["random"]["global"] = 0.0

    def _initialize_subunits(self, num_subunits, method, filter_names):
        """_glm.py:324-367.  Every subunit aliases the same X / XS / S and starts
        from the same b.  Dev entries are cleaned up whenever dev DESIGNS exist
        (the reference keys this on `self.y`, which is empty before `fit`; see
        SURVEY.md section 7 'latent reference crash')."""
        filter_names.remove("stimulus")
        filter_names = [f"stimulus_s{i}" for i in range(num_subunits)] + filter_names
        for name in filter_names:
            if "stimulus" not in name:
                continue
            self.dims[name] = self.dims["stimulus"]
            self.df[name] = self.dims["stimulus"]              # sic (_glm.py:331)
            self.shift[name] = self.shift["stimulus"]
            self.filter_nonlinearity[name] = self.filter_nonlinearity["stimulus"]
            self.intercept[method][name] = self.intercept[method]["stimulus"]
            self.w[method][name] = self.w[method]["stimulus"]
            self.X["train"][name] = self.X["train"]["stimulus"]
            if "dev" in self.X:
                self.X["dev"][name] = self.X["dev"]["stimulus"]
            if "stimulus" in self.S:
                self.b[method][name] = self.b[method]["stimulus"]
                self.XS["train"][name] = self.XS["train"]["stimulus"]
                if "dev" in self.XS:
                    self.XS["dev"][name] = self.XS["dev"]["stimulus"]
                self.S[name] = self.S["stimulus"]
        self.b[method].pop("stimulus", None)
        self.w[method].pop("stimulus")
        self.intercept[method].pop("stimulus")
        self.X["train"].pop("stimulus")
        if "dev" in self.X:
            self.X["dev"].pop("stimulus", None)
        if self.XS != {}:
            self.XS["train"].pop("stimulus")
            if "dev" in self.XS:
                self.XS["dev"].pop("stimulus", None)
            self.S.pop("stimulus")
        self.filter_names = filter_names

    def _initialize_p0(self, method, add_noise_to_mle=False, random_seed=2046):
        """_glm.py:369-393."""
        p0 = {}
        for i, name in enumerate(self.filter_names):
            leaf = self.b[method][name] if name in self.S else self.w[method][name]
            self.p[method][name] = leaf
            rng = np.random.default_rng(random_seed + i)
            noise = add_noise_to_mle * rng.standard_normal(leaf.shape).astype(self.dtype)
            p0[name] = (leaf + noise).astype(self.dtype)
        self.p[method]["intercept"] = self.intercept[method]
        p0["intercept"] = dict(self.intercept[method])
        self.p0 = p0

    def initialize(self, y=None, num_subunits=1, dt=0.033, method="random", compute_ci=True,
                   random_seed=2046, verbose=0, add_noise_to_mle=False):
        """_glm.py:395-436."""
        self.dt = dt
        self.init_method = method
        self.num_subunits = num_subunits
        self.compute_ci = compute_ci
        if method == "random":
            self._initialize_random(random_seed=random_seed)
        elif method == "mle":
            if not self.mle_computed:
                self.compute_mle(y)
        else:
            raise ValueError(f"`{method}` is not supported.")
        filter_names = self.filter_names.copy()
        if num_subunits != 1:
            self._initialize_subunits(num_subunits, method, filter_names)
        self.p[method] = {}
        self._initialize_p0(method, add_noise_to_mle=add_noise_to_mle, random_seed=random_seed)

    def compute_mle(self, y):
        """_glm.py:438-537: least squares on the ones-augmented column stack
        [1 | 1 XS_f1 | 1 XS_f2 ...] through the normal equations + lstsq."""
        if type(y) is dict:
            y_train = np.asarray(y["train"])
            if len(y_train) == 0:
                raise ValueError("Training set is empty after burned in.")
            y_dev = y.get("dev", None)
        else:
            y_train = np.asarray(y)
            y = {"train": y_train}
            y_dev = None
        n = len(y_train) - self.burn_in
        ones = np.ones((n, 1))
        cols = []
        for name in self.filter_names:
            M = self.XS["train"][name] if name in self.S else self.X["train"][name]
            cols.append(np.hstack([ones, M]))
        A = np.hstack([ones] + cols)                                        # :461-475
        XtX = A.T @ A
        Xty = A.T @ y_train[self.burn_in:]
        mle = np.linalg.lstsq(XtX, Xty, rcond=None)[0]                      # :480

        self.b["mle"], self.w["mle"], self.intercept["mle"] = {}, {}, {}
        edges = np.cumsum([0] + [self.n_features[name] + 1 for name in self.filter_names])
        for i, name in enumerate(self.filter_names):                        # :495-502
            seg = mle[edges[i]:edges[i + 1]][:, np.newaxis].astype(self.dtype)
            self.intercept["mle"][name] = seg[0]
            if name in self.S:
                self.b["mle"][name] = seg[1:]
                self.w["mle"][name] = self.S[name] @ self.b["mle"][name]
            else:
                self.w["mle"][name] = seg[1:]
        self.intercept["mle"]["global"] = mle[0]                            # :504
        self.p["mle"] = {name: (self.b["mle"][name] if name in self.S else self.w["mle"][name])
                         for name in self.filter_names}
        self.p["mle"]["intercept"] = self.intercept["mle"]
        self.y["train"] = y_train[self.burn_in:]
        if len(self.y["train"]) == 0:
            raise ValueError("Training set is empty after burned in.")
        self.y_pred["mle"] = {"train": self.forwardpass(self.p["mle"], "train")}
        if getattr(self, "compute_ci", None):                                # :521-524
            self._get_filter_variance("mle")
            self._get_response_variance("mle", "train")
        if y_dev is not None:
            self.y["dev"] = np.asarray(y_dev)[self.burn_in:]
            if len(self.y["dev"]) == 0:
                raise ValueError("Dev set is empty after burned in.")
            self.y_pred["mle"]["dev"] = self.forwardpass(self.p["mle"], "dev")
            if getattr(self, "compute_ci", None):                            # :532-533
                self._get_response_variance("mle", "dev")
        self.mle_computed = True

    # ---- forward / cost / gradient ---------------------------------------
    def _scalar(self, c):
        """An intercept leaf entering array arithmetic, with JAX's promotion under
        `jax_enable_x64` (= NumPy's NEP 50): a Python float (the initial 0.0 of
        _glm.py:311,322) is weak and takes the array dtype; an array leaf keeps its own
        dtype -- after the first Adam update the intercepts are float64 arrays, so in
        float32 mode `XS @ b + intercept` and every nonlinearity after it run in float64
        until `forwardpass` casts the prediction back (`.astype(self.dtype)`, :568-571).
        Confirmed against the reference run through oracle/_jaxshim."""
        if type(c) is float or type(c) is int:
            return self.dtype(c)
        return np.asarray(c).reshape(())[()]

    def _forward_parts(self, p, kind):
        icpt = p["intercept"]
        acts, outs = {}, []
        for name in self.X[kind]:                                           # _glm.py:555
            M = self.XS[kind][name] if name in self.S else self.X[kind][name]
            a = np.squeeze(M @ np.asarray(p[name], dtype=self.dtype), axis=-1) \
                + self._scalar(icpt[name])
            acts[name] = a
            outs.append(fnl(a, self.filter_nonlinearity[name]))            # :561-565
        u = np.array(outs).sum(0) + self._scalar(icpt["global"])            # :567-569
        r = fnl(u, self.output_nonlinearity).astype(self.dtype)             # :568-571
        return acts, u, r

    def forwardpass(self, p, kind):
        """_glm.py:539-571."""
        with np.errstate(over="ignore"):
            return self._forward_parts(p, kind)[2]

    def _data_loss(self, y, r):
        if self.distr == "gaussian":
            return loss_mse(y, r)
        if self.distr == "poisson":
            return loss_neglogli(y, r)
        raise NotImplementedError(self.distr)

    def _penalised(self, kind, penalize):
        return penalize and (self.beta is not None and self.beta != 0) and kind == "train"

    def cost(self, p, kind="train", precomputed=None, penalize=True):
        """_glm.py:573-614."""
        y = self.y[kind]
        r = self.forwardpass(p, kind) if precomputed is None else precomputed
        loss = self._data_loss(y, r)
        if self._penalised(kind, penalize):                                 # :606-612
            w = np.hstack([np.asarray(p[name]).flatten() for name in self.filter_names])
            loss = loss + loss_penalty(w.astype(self.dtype), self.alpha, self.beta)
        return np.squeeze(loss)

    def value_and_grad(self, p, kind="train"):
        """`value_and_grad(self.cost)(p)` of _glm.py:692-696, by hand."""
        y = self.y[kind]
        with np.errstate(over="ignore", invalid="ignore", divide="ignore"):
            acts, u, r = self._forward_parts(p, kind)
            loss = self._data_loss(y, r)
            if self.distr == "gaussian":
                d_r = loss_mse_vjp(y, r)
            else:
                d_r = loss_neglogli_vjp(y, r)
            # cotangents take the dtype of the value they belong to (the transpose of a
            # promotion is a cast back): u and the activations may be float64 in float32 mode
            d_u = fnl_vjp(u, d_r.astype(u.dtype), self.output_nonlinearity)
            d_u = np.where(np.isnan(d_u) & (d_r == 0), u.dtype.type(0), d_u)

            grads = {"intercept": {"global": np.sum(d_u)}}
            for name in self.X[kind]:
                M = self.XS[kind][name] if name in self.S else self.X[kind][name]
                a = acts[name]
                d_a = fnl_vjp(a, d_u.astype(a.dtype), self.filter_nonlinearity[name])
                grads[name] = M.T @ d_a.astype(M.dtype)[:, np.newaxis]   # (p, n) @ (n, 1), as the transpose of XS @ b
                grads["intercept"][name] = np.sum(d_a)

            if self._penalised(kind, True):
                names = self.filter_names
                w = np.hstack([np.asarray(p[n]).flatten() for n in names]).astype(self.dtype)
                loss = loss + loss_penalty(w, self.alpha, self.beta)
                gw = loss_penalty_vjp(w, self.alpha, self.beta)
                pos = 0
                for n in names:
                    k = np.asarray(p[n]).size
                    grads[n] = grads[n] + gw[pos:pos + k][:, np.newaxis]
                    pos += k
        return np.squeeze(loss), grads

    # ---- optimise ---------------------------------------------------------
    def compute_score(self, y, y_pred, metric):
        return _metrics.compute_score(y, y_pred, metric)

    def optimize(self, p0, num_iters, metric, step_size, tolerance, verbose, return_model,
                 atol=1e-5, min_iters=300):
        """_glm.py:644-836."""
        if return_model is None:                                            # :682-686
            return_model = "best_dev_cost" if "dev" in self.y else "best_train_cost"
        assert ("dev" in self.y) or ("dev" not in return_model), \
            "Cannot use dev set if dev set is not given."

        # optimiser state: filter leaves in self.dtype.  An intercept leaf keeps the type it
        # arrives with: a Python float (random init, _glm.py:311,322) is a weak float64
        # scalar -- its Adam state is float64 and after the first update it is a float64
        # array; an array leaf (the (1,)-shaped intercepts of the mle init, :497) stays in
        # its own dtype and shape.
        names = list(self.filter_names)
        x = {n: np.asarray(p0[n], dtype=self.dtype).copy() for n in names}
        xi = {k: (float(v) if type(v) in (float, int) else np.asarray(v).copy())
              for k, v in p0["intercept"].items()}
        ti = {k: (np.float64 if type(c) is float else c.dtype.type) for k, c in xi.items()}
        m = {n: np.zeros_like(x[n]) for n in names}
        v = {n: np.zeros_like(x[n]) for n in names}
        mi = {k: np.zeros_like(c, dtype=ti[k]) for k, c in xi.items()}
        vi = {k: np.zeros_like(c, dtype=ti[k]) for k, c in xi.items()}

        def pack():
            p = {n: x[n].copy() for n in names}
            p["intercept"] = dict(xi)
            return p

        cost_train = np.full(num_iters, np.nan)
        cost_dev = np.full(num_iters, np.nan)
        metric_train = np.full(num_iters, np.nan)
        metric_dev = np.full(num_iters, np.nan)
        params_list = []
        has_dev = "dev" in self.y

        t0 = time.time()
        i = 0
        stop = "maxiter_stop"
        y_pred_train = y_pred_dev = None
        total_time_elapsed = 0.0
        for i in range(num_iters):                                          # :720
            loss, g = self.value_and_grad(pack(), "train")
            cost_train[i] = loss                                            # pre-update
            if not np.isfinite(loss) or any(not np.all(np.isfinite(g[n])) for n in names):
                raise FloatingPointError("invalid value (nan/inf) encountered in cost/grad")
            lr = _step_size_at(step_size, i)
            for n in names:
                x[n], m[n], v[n] = adam_step(i, x[n], g[n].astype(self.dtype), m[n], v[n],
                                             lr, self.dtype)
            for k in xi:
                gk = np.asarray(g["intercept"][k], dtype=ti[k]).reshape(np.shape(xi[k]))
                xi[k], mi[k], vi[k] = adam_step(i, xi[k], gk, mi[k], vi[k], lr, ti[k])
            params = pack()                                                 # post-update
            params_list.append(params)

            y_pred_train = self.forwardpass(params, "train")                # :726-727
            metric_train[i] = self.compute_score(self.y["train"], y_pred_train, metric)
            if has_dev:                                                     # :729-734
                y_pred_dev = self.forwardpass(params, "dev")
                cost_dev[i] = self.cost(params, "dev", precomputed=y_pred_dev, penalize=False)
                metric_dev[i] = self.compute_score(self.y["dev"], y_pred_dev, metric)

            if tolerance and i > min_iters:                                 # :748-775
                total_time_elapsed = time.time() - t0
                if has_dev and np.all(np.diff(cost_dev[i - tolerance:i]) > 0):
                    stop = "dev_stop"
                    break
                if np.all(-np.diff(cost_train[i - tolerance:i]) < atol):
                    stop = "train_stop"
                    break
        else:
            total_time_elapsed = time.time() - t0
            stop = "maxiter_stop"

        best = select_best(return_model, metric, stop, i, tolerance,
                           cost_train, cost_dev, metric_train, metric_dev)

        self.best_iteration = best
        self.return_model = return_model
        self.train_stop = stop
        self.cost_train = cost_train[: i + 1]
        self.cost_dev = cost_dev[: i + 1]
        self.metric_train = metric_train[: i + 1]
        self.metric_dev = metric_dev[: i + 1]
        self.metric_dev_opt = metric_dev[best]
        self.total_time_elapsed = total_time_elapsed
        self.all_params = params_list[: i + 1]
        self.y_pred["opt"]["train"] = y_pred_train                          # :832-834
        if has_dev:
            self.y_pred["opt"]["dev"] = y_pred_dev
        return params_list[best]

    def fit(self, y=None, num_iters=3, alpha=1, beta=0.01, metric="corrcoef", step_size=1e-3,
            tolerance=10, verbose=True, return_model=None, atol=1e-5, min_iters=300):
        """_glm.py:838-918."""
        self.alpha, self.beta, self.metric = alpha, beta, metric
        if y is None:
            if "train" not in self.y:
                raise ValueError("No `y` is provided.")
        elif type(y) is dict:
            self.y["train"] = np.asarray(y["train"])[self.burn_in:].astype(self.dtype)
            if "dev" in y:
                self.y["dev"] = np.asarray(y["dev"])[self.burn_in:].astype(self.dtype)
        else:
            self.y["train"] = np.asarray(y)[self.burn_in:].astype(self.dtype)
        self.y_pred["opt"] = {}
        self.p["opt"] = self.optimize(self.p0, num_iters, metric, step_size, tolerance, verbose,
                                      return_model, atol, min_iters)
        self._extract_opt_params()

    def _get_filter_variance(self, w_type="opt"):
        """_glm.py:1058-1164, spline and raw filters alike."""
        self.V = getattr(self, "V", {})
        self.b_se = getattr(self, "b_se", {})
        self.w_se = getattr(self, "w_se", {})
        V, b_se, w_se = {}, {}, {}
        p = self.p[w_type]
        if self.distr == "gaussian":
            y = self.y["train"]
            rss = np.sum((y - self.forwardpass(p, "train")) ** 2)                # :1068-1071
        for name in self.filter_names:
            M = (self.XS["train"][name] if name in self.S else self.X["train"][name]).astype(np.float64)
            if self.distr == "gaussian":
                G = M.T @ M
                try:
                    Vn = np.linalg.inv(G)
                except np.linalg.LinAlgError:
                    Vn = np.linalg.pinv(G)
                Vn = Vn * (rss / (len(y) - sum(self.df[name])))                     # :1073,1088
            else:
                b = np.asarray(p[name], dtype=np.float64)
                u = fnl((M @ b).ravel(), self.filter_nonlinearity[name])          # :1126-1128
                U = 1 / fnl(u, self.output_nonlinearity) ** 2                      # :1129-1131
                G = (M.T * U) @ M
                try:
                    Vn = np.linalg.inv(G)                                           # :1134
                except np.linalg.LinAlgError:
                    Vn = np.linalg.pinv(G)
            V[name] = np.abs(Vn)
            se = np.sqrt(np.diag(V[name]))
            if name in self.S:
                b_se[name] = se
                w_se[name] = self.S[name] @ se
            else:
                w_se[name] = se
        self.V[w_type], self.b_se[w_type], self.w_se[w_type] = V, b_se, w_se

    def _get_response_variance(self, w_type="opt", kind="train"):
        """_glm.py:1166-1232.  Followed literally: the band lines use the raw lagged design
        X @ w (:1205-1212) while the centre line uses XS @ b (:1187-1190)."""
        self.y_pred_upper = getattr(self, "y_pred_upper", {})
        self.y_pred_lower = getattr(self, "y_pred_lower", {})
        X, XS = self.X[kind], self.XS.get(kind, None)
        b, w, V = self.b[w_type], self.w[w_type], self.V[w_type]
        icpt = self.intercept[w_type]
        mid, up, lo = [], [], []
        with np.errstate(over="ignore", invalid="ignore"):
            for name in X:
                if name in self.S:
                    M = XS[name]
                    y_se = np.sqrt(np.sum(M @ V[name].astype(M.dtype) * M, 1))          # :1186
                    c = (M @ np.asarray(b[name], dtype=M.dtype)).reshape(-1) + self._scalar(icpt[name])
                else:
                    M = X[name]
                    y_se = np.sqrt(np.sum(M @ V[name].astype(M.dtype) * M, 1))          # :1194-1196
                    c = (M @ np.asarray(w[name], dtype=M.dtype)).reshape(-1) + self._scalar(icpt[name])
                raw = (X[name] @ np.asarray(w[name], dtype=X[name].dtype)).reshape(-1) \
                    + self._scalar(icpt[name])
                nl = self.filter_nonlinearity[name]
                mid.append(fnl(c, nl))
                up.append(fnl(raw + 2 * y_se, nl))                                      # :1205-1208
                lo.append(fnl(raw - 2 * y_se, nl))                                      # :1209-1212
            g = self._scalar(icpt["global"])
            out = self.output_nonlinearity
            self.y_pred.setdefault(w_type, {})[kind] = fnl(np.array(mid).sum(0) + g, out)      # :1214-1218
            self.y_pred_upper.setdefault(w_type, {})[kind] = fnl(np.array(up).sum(0) + g, out)
            self.y_pred_lower.setdefault(w_type, {})[kind] = fnl(np.array(lo).sum(0) + g, out)

    def _extract_opt_params(self):
        """_glm.py:920-933."""
        self.b["opt"], self.w["opt"] = {}, {}
        for name in self.filter_names:
            if name in self.S:
                self.b["opt"][name] = self.p["opt"][name]
                self.w["opt"][name] = self.S[name] @ self.b["opt"][name]
            else:
                self.w["opt"][name] = self.p["opt"][name]
        self.intercept["opt"] = self.p["opt"]["intercept"]
        if self.compute_ci:
            self._get_filter_variance("opt")
            self._get_response_variance("opt", "train")                     # :934-937
            if "dev" in self.y:
                self._get_response_variance("opt", "dev")

    # ---- predict / score ---------------------------------------------------
    def predict(self, X, w_type="opt"):
        """_glm.py:939-995."""
        p = self.p[w_type]
        self.X["test"], self.XS["test"] = {}, {}
        if type(X) is dict:
            for name in X:
                self.add_design_matrix(X[name], dims=self.dims[name], shift=self.shift[name],
                                       name=name, kind="test")
        else:
            self.add_design_matrix(X, dims=self.dims["stimulus"], shift=self.shift["stimulus"],
                                   name="stimulus", kind="test")
        if self.num_subunits != 1:                                          # :983-989
            for name in self.filter_names:
                if "stimulus" in name:
                    self.X["test"][name] = self.X["test"]["stimulus"]
                    self.XS["test"][name] = self.XS["test"]["stimulus"]
            self.X["test"].pop("stimulus")
            self.XS["test"].pop("stimulus")
        y_pred = self.forwardpass(p, "test")
        if self.compute_ci:                                                 # :991-993
            self._get_response_variance(w_type, "test")
        return y_pred

    def score(self, X_test, y_test, metric="corrcoef", w_type="opt", return_prediction=False):
        """_glm.py:1015-1056."""
        y_test = np.asarray(y_test)[self.burn_in:].flatten()
        y_pred = self.predict(X_test if type(X_test) is dict else {"stimulus": X_test}, w_type)
        s = self.compute_score(y_test, y_pred, metric)
        return (s, y_pred) if return_prediction else s


def select_best(return_model, metric, stop, i, tolerance, cost_train, cost_dev,
                metric_train, metric_dev):
    """Model-selection rule of _glm.py:784-808."""
    lower_is_better = metric in ["mse", "gcv"]
    if return_model == "best_dev_cost":
        return int(np.argmin(cost_dev[: i + 1]))
    if return_model == "best_train_cost":
        return int(np.argmin(cost_train[: i + 1]))
    if return_model == "best_dev_metric":
        pick = np.argmin if lower_is_better else np.argmax
        return int(pick(metric_dev[: i + 1]))
    if return_model == "best_train_metric":
        pick = np.argmin if lower_is_better else np.argmax
        return int(pick(metric_train[: i + 1]))
    if return_model != "last":
        print("Provided `return_model` is not supported. Fallback to `last`")
    return i - tolerance if stop == "dev_stop" else i
