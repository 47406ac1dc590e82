"""Spline bases for the B200 engine (host side, float64, tiny).

Produces the same bases as the reference (`rfest/splines.py`: cr :39-94, bs :6-36,
cc :97-171, tp :174-189, build_spline_matrix :213-305) -- bit-identical, because the
floating-point operations are evaluated in the same order -- but keeps the
N-dimensional basis in its separable (Kronecker) form: the engine never needs the
dense (prod(dims) x prod(df)) matrix, only the per-axis factors.
"""

import numpy as np

__all__ = ["cr", "bs", "cc", "tp", "spline_factors", "build_spline_matrix", "KronBasis"]


def _unit(a):
    """Scale to unit Frobenius norm (`uvec`, rfest/utils.py:119-121)."""
    return a / np.linalg.norm(a)


def _knots_by_percentile(x, df):
    qs = np.linspace(0, 1, df)[1:-1] * 100            # df - 2 inner knots (splines.py:50-52)
    inner = np.percentile(np.unique(x), qs)
    return np.unique(np.concatenate(([np.min(x), np.max(x)], inner)))


def _hermite_weights(x, knots):
    """Per-sample interval and cubic weights; the spacing is the mean one (splines.py:60-66)."""
    j = np.maximum(np.searchsorted(knots, x) - 1, 0)
    h = np.mean(knots[1:] - knots[:-1])
    right = knots[j + 1] - x
    left = x - knots[j]
    return (j, h, right / h, left / h,
            (right ** 3 / h - h * right) / 6, (left ** 3 / h - h * left) / 6)


def _second_derivative_map(df, h):
    """F = [0; B^-1 D; 0] mapping knot values to knot second derivatives (splines.py:68-87)."""
    m = df - 2
    B = np.diag(2 * h * np.ones(m) / 3)
    if m > 1:
        B += np.diag(1 * h * np.ones(m - 1) / 6, 1) + np.diag(1 * h * np.ones(m - 1) / 6, -1)
    D = np.zeros((m, df))
    invh = np.eye(m) / h
    D[:, 0:m] += 1 * invh
    D[:, 1:m + 1] += -2 * invh
    D[:, 2:m + 2] += 1 * invh
    return np.vstack([np.zeros(df), np.linalg.solve(B, D), np.zeros(df)])


def cr(x, df):
    """Natural cubic regression spline basis on sample positions `x` -> (len(x), df)."""
    x = np.asarray(x)
    knots = _knots_by_percentile(x, df)
    j, h, ajm, ajp, cjm, cjp = _hermite_weights(x, knots)
    F = _second_derivative_map(df, h)
    I = np.eye(df)
    cols = ajm * I[j, :].T + ajp * I[j + 1, :].T + cjm * F[j, :].T + cjp * F[j + 1, :].T
    return _unit(cols.T)


def bs(x, df, degree=3):
    """B-spline basis, clamped, inner knots by percentile -> (len(x), df)."""
    from scipy.interpolate import BSpline

    x = np.asarray(x)
    order = degree + 1
    n_inner = int(df) - order
    qs = np.linspace(0, 1, n_inner + 2)[1:-1] * 100
    knots = np.sort(np.hstack(([np.min(x), np.max(x)] * order, np.percentile(x, qs))))
    nb = len(knots) - order
    eye = np.eye(nb)
    return _unit(np.vstack([BSpline(knots, eye[k], degree)(x) for k in range(nb)]).T)


def cc(x, df):
    """Cyclic cubic regression spline basis -> (len(x), df - 1)."""
    x = np.array(x, copy=True)
    knots = _knots_by_percentile(x, df)
    lo, hi = min(knots), max(knots)
    above, below = x > hi, x < lo
    x[above] = lo + (x[above] - hi) % (hi - lo)
    x[below] = hi - (lo - x[below]) % (hi - lo)
    n = df - 1

    j, h, ajm, ajp, cjm, cjp = _hermite_weights(x, knots)
    kd = knots[1:] - knots[:-1]
    Bm = np.zeros((n, n))
    Dm = np.zeros((n, n))
    Bm[0, 0] = (kd[n - 1] + kd[0]) / 3.0
    Bm[0, n - 1] = Bm[n - 1, 0] = kd[n - 1] / 6.0
    Dm[0, 0] = -1.0 / kd[0] - 1.0 / kd[n - 1]
    Dm[0, n - 1] = Dm[n - 1, 0] = 1.0 / kd[n - 1]
    for k in range(1, n):
        Bm[k, k] = (kd[k - 1] + kd[k]) / 3.0
        Bm[k, k - 1] = Bm[k - 1, k] = kd[k - 1] / 6.0
        Dm[k, k] = -1.0 / kd[k - 1] - 1.0 / kd[k]
        Dm[k, k - 1] = Dm[k - 1, k] = 1.0 / kd[k - 1]
    F = np.linalg.solve(Bm, Dm)
    I = np.eye(n)
    jn = j + 1
    jn[jn == n] = 0
    cols = ajm * I[j, :].T + ajp * I[jn, :].T + cjm * F[j, :].T + cjp * F[jn, :].T
    return _unit(cols.T)


def tp(x, df):
    """Truncated thin-plate regression spline basis -> (len(x), df)."""
    x = np.asarray(x).ravel()
    r = np.abs(x - x.reshape(-1, 1))
    U = np.linalg.svd(r ** 2 * np.log(r + 1e-10))[0]
    return _unit(U[:, : int(df)])


_KINDS = {"cr": cr, "cc": cc, "bs": bs, "tp": tp}


def spline_factors(dims, df, smooth):
    """One (d, df_d) basis per axis, on the integer grid arange(d)."""
    dims, df = list(dims), list(df)
    if len(df) != len(dims):
        raise ValueError("`df` must have the same length as `dims`")
    if smooth not in _KINDS:
        raise ValueError("Input method `{}` is not supported.".format(smooth))
    if not 1 <= len(dims) <= 3:
        raise NotImplementedError(len(dims))
    return [_KINDS[smooth](np.arange(d).ravel(), f) for d, f in zip(dims, df)]


def _kron_all(factors):
    out = factors[-1]
    for f in reversed(factors[:-1]):
        out = np.kron(f, out)
    return out


def build_spline_matrix(dims, df, smooth, dtype=np.float64):
    """Dense basis matrix, (prod(dims), prod(df)) -- API parity with the reference.
    The engine itself uses `KronBasis` and never builds this."""
    return _unit(_kron_all(spline_factors(dims, df, smooth))).astype(dtype)


class KronBasis:
    """S = kron(St, kron(Sx, Sy)) / ||.||_F kept as factors.

    Behaves like the dense matrix where the GLM API needs it (`shape`, `S @ b`,
    `np.asarray(S)`), and hands the engine what the projection kernel wants:
    the temporal factor (with the global normalisation folded in) and the spatial
    Kronecker factor, both float32.
    """

    def __init__(self, dims, df, smooth):
        self.dims = [int(d) for d in dims]
        self.df = [int(f) for f in df]
        self.smooth = smooth
        self.factors = spline_factors(self.dims, self.df, smooth)
        # cc drops one column per axis: take the sizes from the factors themselves
        self.n_coef_axes = [f.shape[1] for f in self.factors]
        self.shape = (int(np.prod(self.dims)), int(np.prod(self.n_coef_axes)))
        # the reference re-normalises the Kronecker product (splines.py:305); the norm of a
        # Kronecker product is the product of the factor norms
        self.norm = float(np.prod([np.linalg.norm(f) for f in self.factors]))
        self._dense = None

    @property
    def n_coef(self):
        return self.shape[1]

    def temporal_f32(self):
        """(dims[0], df_t) float32, carrying the 1/||S|| factor."""
        return np.ascontiguousarray((self.factors[0] / self.norm).astype(np.float32))

    def spatial_f32(self):
        """(prod(dims[1:]), prod(df[1:])) float32, or None for a purely temporal filter."""
        if len(self.factors) == 1:
            return None
        return np.ascontiguousarray(_kron_all(self.factors[1:]).astype(np.float32))

    def dense(self, dtype=np.float64):
        if self._dense is None:
            self._dense = _unit(_kron_all(self.factors))
        return self._dense.astype(dtype, copy=False)

    def __array__(self, dtype=None, copy=None):
        return self.dense(dtype or np.float64)

    def __matmul__(self, b):
        """w = S @ b through the factors (never forms S)."""
        b = np.asarray(b, dtype=np.float64)
        tail = b.shape[1:]
        t = b.reshape(self.n_coef_axes + [-1])
        for axis, f in enumerate(self.factors):
            t = np.moveaxis(np.tensordot(f, t, axes=([1], [axis])), 0, axis)
        return (t / self.norm).reshape((self.shape[0],) + tail)
