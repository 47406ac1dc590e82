"""`rfest_b200.GLM`: the reference's `rfest.GLM` surface on the B200 engine.

Same constructor, methods, keyword arguments, attribute dictionaries and trace semantics
as `rfest/GLM/_glm.py:24` (cited per method), so it drops in for the JAX path.  Underneath
nothing is shared with it:

* `add_design_matrix` never builds the lagged (Hankel) matrix; a fused CUDA kernel projects
  the raw stimulus onto the separable spline basis straight into HBM (csrc/project.cuh);
* `fit` replays one CUDA graph per iteration: a single streaming sweep over XS that yields
  loss, gradient and all monitoring sums, an all-reduce when the samples are sharded over
  several GPUs, and a small penalty + Adam kernel (csrc/sweep.cuh, csrc/update.cuh).

Arithmetic is float32 with float64 accumulation (the reference's fp32 JAX path); there is
no CPU fallback -- constructing a GLM without the CUDA library / a B200 raises.

Differences from the reference that are part of the design (documented in DESIGN.md):
`self.X[kind][name]` holds a lazy descriptor instead of the n x prod(dims) matrix;
`self.XS[kind][name]` is a device block (np.asarray() copies it back); `self.S[name]` is a
Kronecker-factored basis (np.asarray() gives the dense matrix); `method='random'` draws from
NumPy's generator because `jax.random` cannot be reproduced without JAX.
"""

import time
import warnings

import numpy as np

from . import _lib, metrics as _metrics
from .engine import Block, DeviceModel, Engine, _is_cuda_tensor
from .shard import front_padding, shard_rows, source_rows
from .splines import KronBasis

__all__ = ["GLM"]


def _is_row_source(x):
    """A row source streams a recording that is too large (or not yet materialised) to hand
    over as one array: `shape` = (n_raw, *frame) and `rows(lo, hi)` -> float32 array or CUDA
    tensor with raw rows [lo, hi).  np.memmap-backed files, generators and per-rank shards
    all fit this; the engine only ever asks for the rows it projects next."""
    return hasattr(x, "rows") and callable(x.rows) and hasattr(x, "shape")


class LaggedDesign:
    """Stand-in for the reference's `self.X[kind][name]` (the n x nlag*n_pix matrix that the
    engine never materialises).  Knows its shape; `materialize()` builds it on the host for
    small problems / debugging, following rfest/utils.py:50-67 + the burn-in trim."""

    def __init__(self, raw, nlag, shift, burn_in, lag):
        self._raw = raw
        self.nlag, self.shift, self.burn_in, self.lag = nlag, shift, burn_in, lag
        n_raw = raw.shape[0]
        n_pix = int(np.prod(raw.shape[1:])) if len(raw.shape) > 1 else 1
        self.shape = (n_raw - burn_in, nlag * n_pix) if lag else (n_raw, n_pix)

    def __len__(self):
        return self.shape[0]

    def materialize(self, dtype=np.float32):
        if _is_row_source(self._raw):
            self._raw = self._raw.rows(0, self._raw.shape[0])
        raw = self._raw.cpu().numpy() if _is_cuda_tensor(self._raw) else np.asarray(self._raw)
        flat = raw.reshape(raw.shape[0], -1)
        if not self.lag:
            return flat.astype(dtype)
        n, n_pix = flat.shape
        front = front_padding(self.nlag, self.shift)
        back = -self.shift if self.shift < 0 else 0
        padded = np.zeros((front + n + back, n_pix), dtype=np.float64)
        padded[front:front + n] = flat
        out = np.hstack([padded[i:i + n] for i in range(self.nlag)])
        return out[self.burn_in:].astype(dtype)

    def __array__(self, dtype=None, copy=None):
        return self.materialize(dtype or np.float32)


class ParamsList:
    """`all_params` (rfest/GLM/_glm.py:724,828-830) as a lazy sequence of parameter dicts."""

    def __init__(self, b_all, ic_all, unflatten):
        self._b, self._ic, self._unflatten = b_all, ic_all, unflatten

    def __len__(self):
        return self._b.shape[0]

    def __getitem__(self, i):
        if isinstance(i, slice):
            return [self[k] for k in range(*i.indices(len(self)))]
        if i < 0:
            i += len(self)
        return self._unflatten(self._b[i], self._ic[i])

    def __iter__(self):
        return (self[i] for i in range(len(self)))


class GLM:
    def __init__(self, distr="poisson", output_nonlinearity="none", dtype=np.float32, device=None):
        """rfest/GLM/_glm.py:25-101.  `dtype` is accepted for API compatibility; the engine
        always computes in float32 (the reference's fp32 JAX path)."""
        if distr not in _lib.DISTR:
            raise NotImplementedError(distr)                                   # _glm.py:603
        if np.dtype(dtype) != np.float32:
            warnings.warn(f"rfest_b200.GLM computes in float32 (the reference's fp32 path); dtype="
                          f"{np.dtype(dtype).name} is accepted for API compatibility only", stacklevel=2)
        if output_nonlinearity not in _lib.NL:
            raise NotImplementedError(
                f"Input filter nonlinearity `{output_nonlinearity}` is not supported.")
        self.engine = Engine.get(device).init_distributed()
        self._dm = DeviceModel(self.engine, distr, output_nonlinearity)
        self._dirty = True
        self._y_changed = set()  # kinds whose response alone changed since the last _sync
        self._synced_y = set()   # kinds whose response the device model holds
        self._blocks = {}        # kind -> name -> Block
        self._n_rows = {}        # kind -> rows owned by this rank
        self._row_range = {}     # kind -> (lo, hi) trimmed rows owned by this rank
        self._n_trim = {}        # kind -> trimmed rows over all ranks
        self._order = []         # filter order of the device model

        # Optimization
        self.init_method = None
        self.p0 = None
        self.metric = None
        self.dt = None
        self.compute_ci = None
        self.num_subunits = None
        self.burn_in = None
        self.beta = None
        self.alpha = None
        self.idx = None
        self.cost_dev = None
        self.cost_train = None
        self.metric_dev = None
        self.metric_train = None
        self.metric_dev_opt = None
        self.all_params = None
        self.return_model = None
        self.best_iteration = None
        self.train_stop = None
        self.total_time_elapsed = None

        # Data
        self.X, self.S, self.XS = {}, {}, {}
        self.XtX, self.Xty = {}, {}
        self.y, self.y_pred = {}, {}
        self.y_pred_upper, self.y_pred_lower = {}, {}

        # Model parameters
        self.p, self.b, self.b_se, self.w, self.w_se, self.V, self.intercept = {}, {}, {}, {}, {}, {}, {}

        # Model hyper-parameters
        self.df, self.dims, self.n_features, self.shift = {}, {}, {}, {}
        self.filter_names = []
        self.filter_nonlinearity = {}
        self.output_nonlinearity = output_nonlinearity
        self.distr = distr

        self.mle_computed = False
        self.scores, self.r2pseudo, self.corrcoef = {}, {}, {}
        self.dtype = np.float32

    # ------------------------------------------------------------------------------------
    # design
    # ------------------------------------------------------------------------------------
    def add_design_matrix(self, X, dims=None, df=None, smooth=None, lag=True,
                          filter_nonlinearity="none", kind="train", name="stimulus", shift=0,
                          burn_in=None, global_rows=None):
        """rfest/GLM/_glm.py:172-301.

        `global_rows=(src0, n_raw_total)` (extension, multi-GPU): `X` holds only raw rows
        [src0, src0 + len(X)) of a recording of n_raw_total rows -- this rank's samples plus
        the halo.  Without it `X` is the whole recording and the engine picks its shard.
        """
        if filter_nonlinearity not in _lib.NL:
            raise NotImplementedError(
                f"Input filter nonlinearity `{filter_nonlinearity}` is not supported.")
        streamed = _is_row_source(X)
        if streamed and global_rows is not None:
            raise ValueError("a row source already addresses the whole recording; drop `global_rows`")
        if not streamed and not _is_cuda_tensor(X):
            X = np.asarray(X)
            if X.dtype != np.float32:
                X = X.astype(np.float32)                                       # _glm.py:231-234
        n_given = int(X.shape[0])
        n_pix = int(np.prod(X.shape[1:])) if len(X.shape) > 1 else 1
        src0, n_raw = (0, n_given) if global_rows is None else (int(global_rows[0]), int(global_rows[1]))

        self.X.setdefault(kind, {})
        if kind == "train":                                                    # _glm.py:239-245
            self.filter_nonlinearity[name] = filter_nonlinearity
            self.filter_names.append(name)
            dims = dims if type(dims) is not int else [dims]
            self.dims[name] = dims
            self.shift[name] = shift
        else:                                                                  # _glm.py:246-248
            dims = self.dims[name]
            shift = self.shift[name]
        dims = [int(d) for d in dims]

        if self.burn_in is None:                                               # _glm.py:250-253
            self.burn_in = dims[0] - 1 if burn_in is None else burn_in
        if not lag:                                                            # _glm.py:259-261
            self.burn_in = 0
        burn = self.burn_in

        # which basis?  (_glm.py:265-301)
        basis = None
        register_features = False
        if smooth is None:
            if name in self.S:
                basis = self.S[name]
            elif kind == "test":
                if self.num_subunits is not None and self.num_subunits > 1:
                    basis = self.S["stimulus_s0"]
                else:
                    basis = self.S.get(name)
            register_features = basis is None and kind == "train"
        else:
            self.df[name] = df if type(df) is not int else [df]
            basis = KronBasis(dims, self.df[name], smooth)
            self.S[name] = basis
            register_features = kind == "train"

        nlag = dims[0] if lag else 1
        eff_shift = shift if lag else 0
        if basis is not None:
            if lag:
                St, Sxy = basis.temporal_f32(), basis.spatial_f32()
            else:   # instantaneous filter: the whole basis acts on the frame
                St, Sxy = np.ones((1, 1), np.float32), basis.dense(np.float32)
                if n_pix == 1 and Sxy.shape == (1, 1):
                    St, Sxy = Sxy, None
            n_cols = basis.shape[1]
            expect_pix = basis.shape[0] // (dims[0] if lag else 1)
            if expect_pix != n_pix:
                raise ValueError(f"X has {n_pix} values per frame but dims={dims} needs {expect_pix}")
        else:       # no basis: the lagged stimulus itself is the design
            St = np.eye(nlag, dtype=np.float32)
            Sxy = None if n_pix == 1 else np.eye(n_pix, dtype=np.float32)
            n_cols = nlag * n_pix

        # rows of the (burn-in trimmed) design owned by this rank
        n_trim = n_raw - burn
        if n_trim <= 0:
            raise ValueError("Training set is empty after burned in." if kind == "train"
                             else "Dev set is empty after burned in.")             # _glm.py:454,531
        lo, hi = shard_rows(n_trim, self.engine.rank, self.engine.world)
        if kind in self._row_range and self._row_range[kind] != (lo, hi):
            raise ValueError(f"all `{kind}` inputs must have the same number of samples")
        self._row_range[kind] = (lo, hi)
        self._n_trim[kind] = n_trim
        n_out = hi - lo
        first_t = burn + lo
        need_lo, need_hi = source_rows(first_t, n_out, nlag, eff_shift, n_raw)

        block = Block(self.engine, n_out, n_cols)
        common = dict(col0=0, n_raw_total=n_raw, n_pix=n_pix, nlag=nlag, shift=eff_shift, St=St, Sxy=Sxy)
        if streamed:
            # pull only the raw rows each slab of design rows reads (this rank's rows + halo)
            step = int(getattr(X, "chunk_rows", 1 << 18))
            for k0 in range(0, n_out, step):
                kc = min(step, n_out - k0)
                s_lo, s_hi = source_rows(first_t + k0, kc, nlag, eff_shift, n_raw)
                rows = X.rows(s_lo, s_hi)
                rows = rows.reshape(rows.shape[0], -1) if len(rows.shape) != 2 else rows
                block.project(rows, dst_row0=k0, n_out=kc, src0=s_lo, first_t=first_t + k0, **common)
                del rows
        else:
            if global_rows is None and self.engine.world > 1 and need_hi > need_lo:
                X_part, part0 = X[need_lo:need_hi], need_lo      # ship only what this rank reads
            else:
                X_part, part0 = X, src0
            X_part = X_part.reshape(X_part.shape[0], -1) if len(X_part.shape) != 2 else X_part
            block.project(X_part, dst_row0=0, n_out=n_out, src0=part0, first_t=first_t, **common)

        self.X[kind][name] = LaggedDesign(X, nlag, eff_shift, burn, lag)
        self._blocks.setdefault(kind, {})[name] = block
        self._n_rows[kind] = n_out
        if basis is not None:
            self.XS.setdefault(kind, {})[name] = block
        if register_features:
            self.n_features[name] = n_cols
        self._dirty = True

    # ------------------------------------------------------------------------------------
    # initialisation
    # ------------------------------------------------------------------------------------
    def _initialize_random(self, verbose=0, random_seed=2046):
        """_glm.py:303-322 with NumPy's generator instead of `jax.random` (not reproducible
        without JAX): filter i draws default_rng(random_seed + i).standard_normal."""
        self.b["random"], self.w["random"], self.intercept["random"] = {}, {}, {}
        if verbose:
            print("Initializing model parameters randomly...")
        for i, name in enumerate(self.filter_names):
            self.intercept["random"][name] = 0.0
            draw = np.random.default_rng(random_seed + i).standard_normal(
                (self.n_features[name], 1)).astype(np.float32)
            if name in self.S:
                self.b["random"][name] = draw
                self.w["random"][name] = (self.S[name] @ draw).astype(np.float32)
            else:
                self.w["random"][name] = draw
        self.intercept["random"]["global"] = 0.0

    def _initialize_subunits(self, num_subunits, method, filter_names):
        """_glm.py:324-367.  Subunits alias the same device block.  The stale 'stimulus'
        entry of the dev set is removed whenever dev DESIGNS exist (the reference keys this
        on `self.y`, which is still empty here, and then crashes in forwardpass)."""
        filter_names.remove("stimulus")
        filter_names = [f"stimulus_s{i}" for i in range(num_subunits)] + filter_names
        for name in filter_names:
            if "stimulus" not in name:
                continue
            self.dims[name] = self.dims["stimulus"]
            self.df[name] = self.dims["stimulus"]                      # sic, _glm.py:331
            self.shift[name] = self.shift["stimulus"]
            self.filter_nonlinearity[name] = self.filter_nonlinearity["stimulus"]
            self.n_features[name] = self.n_features["stimulus"]
            self.intercept[method][name] = self.intercept[method]["stimulus"]
            self.w[method][name] = self.w[method]["stimulus"]
            for kind in ("train", "dev"):
                if kind in self.X and "stimulus" in self.X[kind]:
                    self.X[kind][name] = self.X[kind]["stimulus"]
                    self._blocks[kind][name] = self._blocks[kind]["stimulus"]
            if "stimulus" in self.S:
                self.b[method][name] = self.b[method]["stimulus"]
                for kind in ("train", "dev"):
                    if kind in self.XS and "stimulus" in self.XS[kind]:
                        self.XS[kind][name] = self.XS[kind]["stimulus"]
                self.S[name] = self.S["stimulus"]
        self.b[method].pop("stimulus", None)
        self.w[method].pop("stimulus")
        self.intercept[method].pop("stimulus")
        for kind in ("train", "dev"):
            for store in (self.X, self.XS, self._blocks):
                if kind in store:
                    store[kind].pop("stimulus", None)
        self.S.pop("stimulus", None)
        self.filter_names = filter_names
        self._dirty = True

    def _initialize_p0(self, method, add_noise_to_mle=False, random_seed=2046):
        """_glm.py:369-393."""
        p0 = {}
        for i, name in enumerate(self.filter_names):
            leaf = self.b[method][name] if name in self.S else self.w[method][name]
            leaf = np.asarray(leaf, dtype=np.float32)
            self.p[method][name] = leaf
            noise = add_noise_to_mle * np.random.default_rng(random_seed + i).standard_normal(
                leaf.shape).astype(np.float32)
            p0[name] = (leaf + noise).astype(np.float32)
        self.p[method]["intercept"] = self.intercept[method]
        p0["intercept"] = dict(self.intercept[method])
        self.p0 = p0

    def initialize(self, y=None, num_subunits=1, dt=0.033, method="random", compute_ci=True,
                   random_seed=2046, verbose=0, add_noise_to_mle=False):
        """_glm.py:395-436.  `compute_ci` gives the coefficient covariances `V`, `b_se`, `w_se`
        (`_get_filter_variance`) and the response confidence bands `y_pred_upper / y_pred_lower`
        (`_get_response_variance`)."""
        self.dt = dt
        self.init_method = method
        self.num_subunits = num_subunits
        self.compute_ci = compute_ci
        if method == "random":
            self._initialize_random(verbose=verbose, random_seed=random_seed)
        elif method == "mle":
            if verbose:
                print("Initializing model parameters with maximum likelihood...")
            if not self.mle_computed:
                self.compute_mle(y)
        else:
            raise ValueError(f"`{method}` is not supported.")                   # _glm.py:420
        if verbose:
            print("Finished.")
        filter_names = self.filter_names.copy()
        if num_subunits != 1:
            self._initialize_subunits(num_subunits, method, filter_names)
        self.p[method] = {}
        self._initialize_p0(method, add_noise_to_mle=add_noise_to_mle, random_seed=random_seed)

    def compute_mle(self, y, compute_ci=True):
        """_glm.py:438-537: least squares on the ones-augmented column stack
        A = [1 | 1 XS_f1 | 1 XS_f2 ...] through the normal equations and `lstsq`.

        The device computes the float64 Gram pieces XS^T XS, XS^T 1, XS^T y in one pass over the
        designs (csrc/gram.cuh; all-reduced over ranks); the (P+F+1)^2 system is assembled and
        solved on the host in float64, as the reference does (its ones are float64)."""
        if type(y) is dict:
            y_train = y["train"]
            if len(y_train) == 0:
                raise ValueError("Training set is empty after burned in.")
            y_dev = y.get("dev", None)
        else:
            y_train, y_dev = y, None
        if y_train is None:
            raise ValueError("`y` is needed for the maximum-likelihood initialisation.")
        self.y["train"] = self._trim_y(y_train, "train")
        if len(self.y["train"]) == 0:
            raise ValueError("Training set is empty after burned in.")
        if y_dev is not None:
            self.y["dev"] = self._trim_y(y_dev, "dev")
            if len(self.y["dev"]) == 0:
                raise ValueError("Dev set is empty after burned in.")
        self._dirty = True
        self._sync()
        G, colsum, xty = self._dm.gram("train")
        n = float(self._n_trim["train"])
        sum_y = float(np.sum(self.y["train"], dtype=np.float64))
        if self.engine.world > 1:
            sum_y = float(self.engine.allreduce_f64([sum_y])[0])

        # column k of A is either a ones column (-1) or design column c of the concatenated blocks
        cols = [-1]
        for name in self.filter_names:
            c0 = self._slot_col0[self._slot_of[name]]
            cols += [-1] + list(range(c0, c0 + self.n_features[name]))
        cols = np.asarray(cols)
        ones = cols < 0
        dc = np.where(ones, 0, cols)
        XtX = G[np.ix_(dc, dc)]
        XtX[ones, :] = np.where(ones, n, colsum[dc])[None, :]
        XtX[:, ones] = np.where(ones, n, colsum[dc])[:, None]
        Xty = np.where(ones, sum_y, xty[dc])
        mle = np.linalg.lstsq(XtX, Xty, rcond=None)[0]                            # _glm.py:480

        self.b["mle"], self.w["mle"], self.intercept["mle"] = {}, {}, {}
        edges = np.cumsum([0] + [self.n_features[name] + 1 for name in self.filter_names])
        self.idx = [np.array((edges[i], edges[i + 1])) for i in range(len(edges) - 1)]
        for i, name in enumerate(self.filter_names):                             # _glm.py:495-502
            seg = mle[edges[i]:edges[i + 1]][:, np.newaxis].astype(np.float32)
            self.intercept["mle"][name] = seg[0]
            if name in self.S:
                self.b["mle"][name] = seg[1:]
                self.w["mle"][name] = (self.S[name] @ seg[1:]).astype(np.float32)
            else:
                self.w["mle"][name] = seg[1:]
        self.intercept["mle"]["global"] = mle[0]
        self.p["mle"] = {name: (self.b["mle"][name] if name in self.S else self.w["mle"][name])
                         for name in self.filter_names}
        self.p["mle"]["intercept"] = self.intercept["mle"]
        self.y_pred["mle"] = {"train": self.forwardpass(self.p["mle"], kind="train")}
        if self.compute_ci is None:
            self.compute_ci = compute_ci
        if self.compute_ci:                                                       # _glm.py:521-524
            self._get_filter_variance(w_type="mle")
            self._get_response_variance(w_type="mle", kind="train")
        if y_dev is not None:
            self.y_pred["mle"]["dev"] = self.forwardpass(self.p["mle"], kind="dev")
            if self.compute_ci:                                                   # _glm.py:532-533
                self._get_response_variance(w_type="mle", kind="dev")
        self.mle_computed = True

    def _get_filter_variance(self, w_type="opt"):
        """_glm.py:1058-1164: covariance and standard errors of every filter's coefficients.

        Gaussian: V = inv(XS^T XS) * rss / (n - sum(df)); otherwise V = inv(XS^T diag(U) XS) with
        U = 1 / fnl(fnl(XS b, filter_nl), output_nl)^2.  The (weighted) Gram matrices are computed on
        the device in float64 (csrc/gram.cuh), the small inverses on the host."""
        self._sync()
        p = self.p[w_type]
        b_flat, ic = self._flatten(p)
        n = float(self._n_trim["train"])
        gaussian = self.distr == "gaussian"
        if gaussian:
            # _glm.py:1068-1071: the residuals of `y_pred[w_type]['train']` -- after a fit that is the
            # LAST iteration's prediction (:832), not the selected model's; reproduced as is
            held = self.y_pred.get(w_type, {}).get("train")
            if held is not None and len(held) == self._n_rows["train"]:
                rss, _ = self._dm.loss_of_predictions("train", held)
            else:
                _, sums = self._dm.eval("train", b_flat, ic)
                rss = sums[_lib.SUM_SQERR]
            # (both entry points return sums over all ranks)
        V, b_se, w_se = {}, {}, {}
        for i, name in enumerate(self.filter_names):
            k = self.n_features[name]
            if n < (sum(self.df[name]) if name in self.S else k):
                print("Sample size is too small for getting reasonable confidence interval.")
            G = self._dm.filter_gram("train", i, k, b=b_flat, weighted=not gaussian)
            try:
                Vn = np.linalg.inv(G)
            except np.linalg.LinAlgError:
                Vn = np.linalg.pinv(G)
            if gaussian:
                Vn = Vn * (rss / (n - sum(self.df[name] if name in self.df else self.dims[name])))
            V[name] = np.abs(Vn)                                                 # _glm.py:1095,1138
            se = np.sqrt(np.diag(V[name]))
            if name in self.S:
                b_se[name] = se
                w_se[name] = self.S[name] @ se
            else:
                w_se[name] = se
        self.V[w_type], self.b_se[w_type], self.w_se[w_type] = V, b_se, w_se

    def _get_response_variance(self, w_type="opt", kind="train"):
        """_glm.py:1166-1232: prediction with its +/- 2 standard-error band.

        y_se_f = sqrt(diag(XS_f V_f XS_f^T)) per filter; the band pushes XS_f b_f + c_f +/- 2 y_se_f
        through the filter and output nonlinearities.  The reference writes the band's linear part
        as X_f w_f with the n x prod(dims) lagged matrix; w_f = S_f b_f makes that XS_f b_f, which is
        what the device kernel (csrc/band.cuh) uses.  Results stay in HBM (`DeviceVector`)."""
        self._sync()
        b_flat, ic = self._flatten(self.p[w_type])
        V = self.V[w_type]
        mats = [V.get(name) for name in self.filter_names]
        pred, upper, lower = self._dm.response_band(kind, b_flat, ic, mats, self._n_local(kind))
        self.y_pred.setdefault(w_type, {})[kind] = pred
        self.y_pred_upper.setdefault(w_type, {})[kind] = upper
        self.y_pred_lower.setdefault(w_type, {})[kind] = lower

    def _n_local(self, kind):
        """This rank's number of samples of `kind`."""
        if kind in self._row_range:
            lo, hi = self._row_range[kind]
            return hi - lo
        return self._n_trim[kind]

    # ------------------------------------------------------------------------------------
    # device model plumbing
    # ------------------------------------------------------------------------------------
    def _trim_y(self, y, kind):
        """y[burn_in:] as float32 (fit, _glm.py:899-903), then this rank's rows.  With several
        ranks `y` may be the whole recording or already this rank's trimmed rows."""
        if _is_cuda_tensor(y):
            y = y.detach().float().cpu().numpy()
        y = np.asarray(y)
        if self.engine.world > 1 and kind in self._row_range:
            lo, hi = self._row_range[kind]
            if len(y) - self.burn_in == self._n_trim[kind]:
                return y[self.burn_in:][lo:hi].astype(np.float32, copy=False)
            if len(y) == hi - lo:
                return y.astype(np.float32, copy=False)
            raise ValueError(f"`y` for `{kind}` has {len(y)} samples; expected the whole recording "
                             f"({self._n_trim[kind] + self.burn_in}) or this rank's rows ({hi - lo})")
        return y[self.burn_in:].astype(np.float32, copy=False)    # a view when y is float32 already

    def _sync(self):
        """(Re)describe filters and data to the device model when something changed."""
        if not self._dirty:
            # only responses changed (a second fit(y=...) on the same design): keep plan and scratch
            for kind in sorted(self._y_changed):
                y = self.y.get(kind)
                n = self._n_rows.get(kind)
                if y is None or n is None:
                    continue
                if len(y) != n:
                    raise ValueError(f"`y` for `{kind}` has {len(y)} samples after burn-in, the design has {n}")
                self._dm.set_y(kind, y, n)
            self._y_changed.clear()
            return
        self._y_changed.clear()
        self._synced_y = set()
        dm = self._dm
        dm.clear_filters()
        train = self._blocks.get("train", {})
        slots, slot_of = [], {}
        for name in self.filter_names:
            blk = train[name]
            for s, other in enumerate(slots):
                if other is blk:
                    slot_of[name] = s
                    break
            else:
                slots.append(blk)
                slot_of[name] = len(slots) - 1
        for name in self.filter_names:
            dm.add_filter(slot_of[name], self.filter_nonlinearity[name], self.n_features[name])
        self._order = list(self.filter_names)
        self._slot_of = slot_of
        self._slot_col0 = [0]
        for blk in slots:                       # padded column offset of every slot in the concatenated design
            self._slot_col0.append(self._slot_col0[-1] + blk.ld)
        for kind, blocks in self._blocks.items():
            per_slot = [None] * len(slots)
            for name, blk in blocks.items():
                if name in slot_of:
                    per_slot[slot_of[name]] = blk
            if all(b is None for b in per_slot):
                continue
            y = self.y.get(kind)
            n = self._n_rows[kind]
            if y is not None and len(y) != n:
                raise ValueError(f"`y` for `{kind}` has {len(y)} samples after burn-in, the design has {n}")
            dm.set_data(kind, per_slot, y, n)
            if y is not None:
                self._synced_y.add(kind)
        self._dirty = False

    def _flatten(self, p):
        names = self.filter_names
        b = np.concatenate([np.asarray(p[n], dtype=np.float32).ravel() for n in names]) \
            if names else np.zeros(0, np.float32)
        icpt = p["intercept"]
        ic = np.array([float(np.asarray(icpt[n]).reshape(-1)[0]) for n in names]
                      + [float(np.asarray(icpt["global"]).reshape(-1)[0])], dtype=np.float64)
        return b, ic

    def _unflatten(self, b, ic):
        p, pos = {}, 0
        for name in self.filter_names:
            k = self.n_features[name]
            p[name] = np.array(b[pos:pos + k], dtype=np.float32).reshape(k, 1)
            pos += k
        p["intercept"] = {name: float(ic[i]) for i, name in enumerate(self.filter_names)}
        p["intercept"]["global"] = float(ic[len(self.filter_names)])
        return p

    # ------------------------------------------------------------------------------------
    # forward / cost
    # ------------------------------------------------------------------------------------
    def fnl(self, x, kind, params=None):
        """GLM.fnl on host arrays (rfest/GLM/_glm.py:103-170); the engine evaluates the same
        functions inside the sweep kernel."""
        x = np.asarray(x)
        one = x.dtype.type(1) if x.dtype.kind == "f" else 1.0
        if kind == "softplus":
            return np.log(one + np.exp(x)) + 1e-7
        if kind == "exponential":
            return np.exp(x)
        if kind == "softmax":
            z = np.exp(x)
            return z / z.sum()
        if kind == "sigmoid":
            return one / (one + np.exp(-x))
        if kind == "tanh":
            return np.tanh(x)
        if kind == "relu":
            return np.where(x > 0.0, x, 1e-7)
        if kind == "leaky_relu":
            return np.where(x > 0.0, x, x * 0.01)
        if kind == "none":
            return x
        raise NotImplementedError(f"Input filter nonlinearity `{kind}` is not supported.")

    def forwardpass(self, p, kind):
        """_glm.py:539-571: prediction for data set `kind` under parameters `p`
        (this rank's samples when the data are sharded)."""
        self._sync()
        b, ic = self._flatten(p)
        r, _ = self._dm.eval(kind, b, ic, want_r=True, n=self._n_rows[kind])
        return r

    def _loss_from_sums(self, sums, kind):
        if self.distr == "poisson":
            return sums[_lib.SUM_LOSS_A] + sums[_lib.SUM_LOSS_B]                  # loss.py:8-10
        return sums[_lib.SUM_LOSS_A] / self._n_trim[kind]                                   # loss.py:15

    def cost(self, p, kind="train", precomputed=None, penalize=True):
        """_glm.py:573-614."""
        if kind not in self.y:
            raise ValueError(f"No `y` is provided for `{kind}`.")
        self._sync()
        b, ic = self._flatten(p)
        if precomputed is not None:                                               # _glm.py:596-599
            if len(precomputed) != self._n_rows[kind]:
                raise ValueError(f"`precomputed` has {len(precomputed)} samples, `{kind}` has {self._n_rows[kind]}")
            la, lb = self._dm.loss_of_predictions(kind, precomputed)
            sums = {_lib.SUM_LOSS_A: la, _lib.SUM_LOSS_B: lb}
        else:
            _, sums = self._dm.eval(kind, b, ic)
        loss = self._loss_from_sums(sums, kind)
        if penalize and (self.beta is not None and self.beta != 0) and kind == "train":
            w = b.astype(np.float32)                                              # loss.py:19-23
            l1 = np.sum(np.abs(w)) if self.alpha != 0.0 else np.float32(0)
            l2 = np.sqrt(np.sum(w * w)) if self.alpha != 1.0 else np.float32(0)
            loss = loss + float(np.float32(self.beta) * (np.float32(1.0 - self.alpha) * l2
                                                         + np.float32(self.alpha) * l1))
        return np.float64(loss)

    def value_and_grad(self, p, kind="train", penalize=True):
        """`value_and_grad(self.cost)(p)` (_glm.py:692-696) in one sweep: -> (cost, gradient dict with
        the layout of `p`).  Extension of the reference surface (there it is a local of `optimize`)."""
        self._sync()
        b, ic = self._flatten(p)
        sums, gb, gi = self._dm.value_and_grad(kind, b, ic)
        loss = self._loss_from_sums(sums, kind)
        gb = gb.astype(np.float64)
        if penalize and (self.beta is not None and self.beta != 0) and kind == "train":
            w = b.astype(np.float64)
            if self.alpha != 0.0:
                loss += self.beta * self.alpha * np.sum(np.abs(w))
                gb = gb + self.beta * self.alpha * np.sign(w)
            if self.alpha != 1.0:
                nrm = np.sqrt(np.sum(w * w))
                loss += self.beta * (1.0 - self.alpha) * nrm
                gb = gb + self.beta * (1.0 - self.alpha) * w / nrm
        g, pos = {}, 0
        for name in self.filter_names:
            k = self.n_features[name]
            g[name] = gb[pos:pos + k].reshape(k, 1)
            pos += k
        g["intercept"] = {name: float(gi[i]) for i, name in enumerate(self.filter_names)}
        g["intercept"]["global"] = float(gi[len(self.filter_names)])
        return np.float64(loss), g

    # ------------------------------------------------------------------------------------
    # optimisation
    # ------------------------------------------------------------------------------------
    @staticmethod
    def print_progress(i, time_elapsed, c_train=None, c_dev=None, m_train=None, m_dev=None):
        """_glm.py:616-629."""
        opt_info = f"{i}".ljust(13) + f"{time_elapsed:>.3f}".ljust(13)
        for v in (c_train, c_dev, m_train, m_dev):
            if v is not None and np.isfinite(v):
                opt_info += f"{v:.3g}".ljust(16)
        print(opt_info)

    @staticmethod
    def print_progress_header(c_train=False, c_dev=False, m_train=False, m_dev=False):
        """_glm.py:631-642."""
        opt_title = "Iters".ljust(13) + "Time (s)".ljust(13)
        for flag, title in ((c_train, "Cost (train)"), (c_dev, "Cost (dev)"),
                            (m_train, "Metric (train)"), (m_dev, "Metric (dev)")):
            if flag:
                opt_title += title.ljust(16)
        print(opt_title)

    def optimize(self, p0, num_iters, metric, step_size, tolerance, verbose, return_model,
                 atol=1e-5, min_iters=300):
        """_glm.py:644-836 -- same traces, early stopping and model selection; the loop body
        is a CUDA graph replayed on the device, checked from the host every `check` iterations."""
        has_dev = "dev" in self.y
        if return_model is None:                                                  # _glm.py:682-686
            return_model = "best_dev_cost" if has_dev else "best_train_cost"
        assert has_dev or ("dev" not in return_model), "Cannot use dev set if dev set is not given."
        if metric not in _lib.METRIC:
            print(f"Metric `{metric}` is not supported.")                         # _glm.py:1012-1013
            metric = None

        self._sync()
        dm = self._dm
        b0, ic0 = self._flatten(p0)
        dm.set_params(b0, ic0)
        steps = np.array([float(step_size(i)) if callable(step_size) else float(step_size)
                          for i in range(num_iters)], dtype=np.float64)
        dm.fit_begin(steps, self.alpha if self.alpha is not None else 1.0,
                     self.beta if self.beta is not None else 0.0, metric)

        if verbose:
            self.print_progress_header(c_train=True, c_dev=has_dev, m_train=metric is not None,
                                       m_dev=metric is not None and has_dev)
        time_start = time.time()
        check = num_iters
        if tolerance and num_iters > min_iters + 1:
            check = 64
        if verbose:
            check = min(check, max(int(verbose), 1))
        executed, i_stop, stop = 0, None, "maxiter_stop"
        next_print, next_scan = 0, min_iters + 1
        stamps = np.zeros(num_iters)
        while executed < num_iters and i_stop is None:
            n = min(check, num_iters - executed)
            dm.fit_run(n)
            executed += n
            if not (tolerance or verbose):
                continue
            c_tr, c_dv, m_tr, m_dv = dm.fit_traces(executed)       # synchronises
            stamps[executed - n:executed] = time.time() - time_start
            if verbose:
                while next_print <= executed - 2:                    # row i is complete after i+1
                    self.print_progress(next_print, stamps[next_print], c_tr[next_print],
                                        c_dv[next_print], m_tr[next_print], m_dv[next_print])
                    next_print += int(verbose)
            if tolerance:
                for i in range(next_scan, executed):                 # _glm.py:748-775
                    if has_dev and np.all(np.diff(c_dv[i - tolerance:i]) > 0):
                        i_stop, stop = i, "dev_stop"
                        break
                    if np.all(-np.diff(c_tr[i - tolerance:i]) < atol):
                        i_stop, stop = i, "train_stop"
                        break
                next_scan = max(next_scan, executed)       # never before min_iters + 1 (_glm.py:748)
        i = executed - 1 if i_stop is None else i_stop
        dm.fit_finish(i)                     # completes trace row i, keeps its predictions
        cost_train, cost_dev, metric_train, metric_dev = dm.fit_traces(i + 1)
        total_time_elapsed = time.time() - time_start
        if verbose:
            while next_print <= i:
                self.print_progress(next_print, total_time_elapsed, cost_train[next_print],
                                    cost_dev[next_print], metric_train[next_print], metric_dev[next_print])
                next_print += int(verbose)
            if stop == "dev_stop":
                print(f"Stop at {i} steps: cost (dev) has been monotonically increasing for {tolerance} steps.")
            elif stop == "train_stop":
                print(f"Stop at {i} steps: cost (train) has been changing less than {atol} for {tolerance} steps.")
            else:
                print("Stop: reached {0} steps.".format(num_iters))
            print("Total time elapsed: {0:.3f}s.".format(total_time_elapsed))

        if not np.all(np.isfinite(cost_train)):
            dm.fit_end()
            # the reference runs with jax_debug_nans=True (_glm.py:18)
            raise FloatingPointError("invalid value (nan/inf) encountered in the cost")

        lower = metric in ["mse", "gcv"]                                          # _glm.py:784-808
        if return_model == "best_dev_cost":
            best = int(np.argmin(cost_dev))
        elif return_model == "best_train_cost":
            best = int(np.argmin(cost_train))
        elif return_model == "best_dev_metric":
            best = int(np.argmin(metric_dev) if lower else np.argmax(metric_dev))
        elif return_model == "best_train_metric":
            best = int(np.argmin(metric_train) if lower else np.argmax(metric_train))
        else:
            if return_model != "last":
                print("Provided `return_model` is not supported. Fallback to `last`")
            best = i - tolerance if stop == "dev_stop" else i
        if verbose:
            print(f"Returning model: {return_model} at iteration {best} of {i} (Max: {num_iters}).\n")

        b_all, ic_all = dm.fit_params_all(i + 1)
        params_list = ParamsList(b_all, ic_all, self._unflatten)
        self.best_iteration = best                                               # _glm.py:815-834
        self.return_model = return_model
        self.train_stop = stop
        self.cost_train, self.cost_dev = cost_train, cost_dev
        self.metric_train, self.metric_dev = metric_train, metric_dev
        self.metric_dev_opt = metric_dev[best]
        self.total_time_elapsed = total_time_elapsed
        self.all_params = params_list
        # like the reference's JAX arrays the predictions stay on the device until they are read
        self.y_pred["opt"]["train"] = dm.fit_predictions_device("train")
        if has_dev:
            self.y_pred["opt"]["dev"] = dm.fit_predictions_device("dev")
        dm.fit_end()
        params = params_list[best]
        dm.set_params(*self._flatten(params))
        return params

    def fit(self, y=None, num_iters=3, alpha=1, beta=0.01, metric="corrcoef", step_size=1e-3,
            tolerance=10, verbose=True, return_model=None, atol=1e-5, min_iters=300):
        """_glm.py:838-918."""
        self.alpha, self.beta, self.metric = alpha, beta, metric
        if y is None:
            if "train" not in self.y:
                raise ValueError("No `y` is provided.")                          # _glm.py:896
        else:
            changed = ["train"]
            if type(y) is dict:
                self.y["train"] = self._trim_y(y["train"], "train")
                if "dev" in y:
                    self.y["dev"] = self._trim_y(y["dev"], "dev")
                    changed.append("dev")
            else:
                self.y["train"] = self._trim_y(y, "train")
            if any(k not in self._synced_y for k in changed):
                self._dirty = True          # first response of a data set: full (re)description
            else:
                self._y_changed.update(changed)
        if self.p0 is None:
            raise ValueError("call initialize() (or set `p0`) before fit()")
        self.y_pred["opt"] = {}
        self.p["opt"] = self.optimize(self.p0, num_iters, metric, step_size, tolerance, verbose,
                                      return_model, atol, min_iters)
        self._extract_opt_params()

    def _extract_opt_params(self):
        """_glm.py:920-930."""
        self.b["opt"], self.w["opt"] = {}, {}
        for name in self.filter_names:
            if name in self.S:
                self.b["opt"][name] = self.p["opt"][name]
                self.w["opt"][name] = (self.S[name] @ self.b["opt"][name]).astype(np.float32)
            else:
                self.w["opt"][name] = self.p["opt"][name]
        self.intercept["opt"] = self.p["opt"]["intercept"]
        if self.compute_ci:                                                       # _glm.py:932-937
            self._get_filter_variance(w_type="opt")
            self._get_response_variance(w_type="opt", kind="train")
            if "dev" in self.y:
                self._get_response_variance(w_type="opt", kind="dev")

    # ------------------------------------------------------------------------------------
    # prediction / scoring
    # ------------------------------------------------------------------------------------
    def predict(self, X, w_type="opt"):
        """_glm.py:939-995."""
        p = self.p[w_type]
        for store in (self.X, self.XS, self._blocks):
            store["test"] = {}
        self._row_range.pop("test", None)
        self._n_trim.pop("test", None)
        if type(X) is dict:
            for name in X:
                self.add_design_matrix(X[name], dims=self.dims[name], shift=self.shift[name],
                                       name=name, kind="test")
        else:
            self.add_design_matrix(X, dims=self.dims["stimulus"], shift=self.shift["stimulus"],
                                   name="stimulus", kind="test")
        if self.num_subunits != 1:                                               # _glm.py:983-989
            for name in self.filter_names:
                if "stimulus" in name:
                    for store in (self.X, self.XS, self._blocks):
                        store["test"][name] = store["test"]["stimulus"]
            for store in (self.X, self.XS, self._blocks):
                store["test"].pop("stimulus")
        self.y.pop("test", None)
        self._dirty = True
        y_pred = self.forwardpass(p, kind="test")
        if self.compute_ci and w_type in self.V:                                 # _glm.py:991-993
            self._get_response_variance(w_type=w_type, kind="test")
        return y_pred

    @staticmethod
    def compute_score(y, y_pred, metric):
        """_glm.py:997-1013."""
        if metric == "r2":
            return _metrics.r2(y, y_pred)
        elif metric == "mse":
            return _metrics.mse(y, y_pred)
        elif metric == "corrcoef":
            return _metrics.corrcoef(y, y_pred)
        else:
            print(f"Metric `{metric}` is not supported.")

    def score(self, X_test, y_test, metric="corrcoef", w_type="opt", return_prediction=False):
        """_glm.py:1015-1056."""
        y_test = np.asarray(y_test)[self.burn_in:].flatten()
        if type(X_test) is dict:
            y_pred = self.predict(X_test, w_type)
        else:
            y_pred = self.predict({"stimulus": X_test}, w_type)
        if self.engine.world > 1:
            # every rank predicts its shard of the test set; the score is the global one, from
            # all-reduced sums (the same closed forms the device uses for the fit's metric traces)
            lo, hi = self._row_range["test"]
            y_loc = np.asarray(y_test[lo:hi], dtype=np.float64)
            r = np.asarray(y_pred, dtype=np.float64)
            sums = self.engine.allreduce_f64([len(y_loc), y_loc.sum(), (y_loc * y_loc).sum(), r.sum(),
                                              (r * r).sum(), (y_loc * r).sum(), ((y_loc - r) ** 2).sum()])
            s = _metrics.score_from_sums(metric, *sums)
            if s is None:
                print(f"Metric `{metric}` is not supported.")
        else:
            s = self.compute_score(y_test, y_pred, metric)
        return (s, y_pred) if return_prediction else s
