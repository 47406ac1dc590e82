"""Oracle: spline bases.  TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Restates `rfest/splines.py` (cr :39-94, bs :6-36, cc :97-171, tp :174-189,
build_spline_matrix :213-305) and `uvec` (`rfest/utils.py:119-121`).  PINNED
against the live reference module and `tests/golden/`.

The floating-point operation order of the reference is kept on purpose (it
decides the last bit), only the way the small matrices are assembled differs.
"""

import numpy as np


def unit_frobenius(a):
    """`uvec`: divide by the 2-norm of the flattened array (utils.py:119-121)."""
    return a / np.linalg.norm(a)


def _percentile_knots(x, df):
    """Boundary knots + `df - 2` inner knots at equally spaced percentiles of the
    unique sample positions (splines.py:49-56 for cr, :113-122 for cc)."""
    q = np.linspace(0, 1, (df - 2) + 2)[1:-1] * 100
    inner = np.percentile(np.unique(x), q)
    return np.unique(np.concatenate(([np.min(x), np.max(x)], inner)))


def _cubic_pieces(x, knots):
    """Interval index and the four cubic-Hermite-like weights (splines.py:60-66,
    :151-157).  `h` is the MEAN knot spacing, used for every interval."""
    j = np.maximum(np.searchsorted(knots, x) - 1, 0)
    h = np.mean(knots[1:] - knots[:-1])
    lo = knots[j]
    hi = knots[j + 1]
    ajm = (hi - x) / h
    ajp = (x - lo) / h
    cjm = ((hi - x) ** 3 / h - h * (hi - x)) / 6
    cjp = ((x - lo) ** 3 / h - h * (x - lo)) / 6
    return j, h, ajm, ajp, cjm, cjp


def cr(x, df):
    """Natural cubic regression spline basis, (len(x), df) (splines.py:39-94)."""
    x = np.asarray(x)
    knots = _percentile_knots(x, df)
    j, h, ajm, ajp, cjm, cjp = _cubic_pieces(x, knots)

    m = df - 2
    # B: (m, m) tridiagonal, h/6 off the diagonal and 2h/3 on it (splines.py:68-75).
    # The entries are formed exactly as the reference does: c*h*1.0 / d.
    B = np.zeros((m, m))
    off = 1 * h * np.ones(m - 1) / 6
    B[np.arange(1, m), np.arange(m - 1)] = off
    B[np.arange(m), np.arange(m)] = 2 * h * np.ones(m) / 3
    B[np.arange(m - 1), np.arange(1, m)] = off
    # D: (m, df) second-difference operator scaled by 1/h (splines.py:76-85).
    D = np.zeros((m, df))
    inv_h = np.eye(m) / h
    rows = np.arange(m)
    D[rows, rows] += (1 * inv_h)[rows, rows]
    D[rows, rows + 1] += (-2 * inv_h)[rows, rows]
    D[rows, rows + 2] += (1 * inv_h)[rows, rows]

    F = np.vstack([np.zeros(df), np.linalg.solve(B, D), np.zeros(df)])  # splines.py:87
    eye = np.eye(df)
    # splines.py:90-92 -- left-to-right sum of the four terms, shape (df, len(x)).
    basis = ajm * eye[j, :].T + ajp * eye[j + 1, :].T + cjm * F[j, :].T + cjp * F[j + 1, :].T
    return unit_frobenius(basis.T)


def bs(x, df, degree=3):
    """B-spline basis with percentile knots (splines.py:6-36)."""
    from scipy.interpolate import BSpline

    x = np.asarray(x)
    df = np.asarray(df)
    order = degree + 1
    q = np.linspace(0, 1, (df - order) + 2)[1:-1] * 100           # splines.py:19-21
    inner = np.percentile(x, q)
    knots = np.sort(np.hstack(([np.min(x), np.max(x)] * order, inner)))
    n_bases = len(knots) - (degree + 1)
    unit = np.eye(n_bases)
    cols = [BSpline(knots, unit[i], degree)(x) for i in range(n_bases)]  # splines.py:34
    return unit_frobenius(np.vstack(cols).T)


def cc(x, df):
    """Cyclic cubic regression spline basis (splines.py:97-171)."""
    x = np.asarray(x)
    knots = _percentile_knots(x, df)
    lb, ub = min(knots), max(knots)
    # wrap into [lb, ub] (splines.py:106-111)
    x = np.copy(x)
    over = x > ub
    x[over] = lb + (x[over] - ub) % (ub - lb)
    under = x < lb
    x[under] = ub - (lb - x[under]) % (ub - lb)
    df = df - 1

    j, h, ajm, ajp, cjm, cjp = _cubic_pieces(x, knots)

    # cyclic penalty pair (splines.py:124-146)
    kd = knots[1:] - knots[:-1]
    n = knots.size - 1
    b = np.zeros((n, n))
    d = np.zeros((n, n))
    b[0, 0] = (kd[n - 1] + kd[0]) / 3.0
    b[0, n - 1] = kd[n - 1] / 6.0
    b[n - 1, 0] = kd[n - 1] / 6.0
    d[0, 0] = -1.0 / kd[0] - 1.0 / kd[n - 1]
    d[0, n - 1] = 1.0 / kd[n - 1]
    d[n - 1, 0] = 1.0 / kd[n - 1]
    for i in range(1, n):
        b[i, i] = (kd[i - 1] + kd[i]) / 3.0
        b[i, i - 1] = kd[i - 1] / 6.0
        b[i - 1, i] = kd[i - 1] / 6.0
        d[i, i] = -1.0 / kd[i - 1] - 1.0 / kd[i]
        d[i, i - 1] = 1.0 / kd[i - 1]
        d[i - 1, i] = 1.0 / kd[i - 1]
    F = np.linalg.solve(b, d)

    eye = np.eye(df)
    j1 = j + 1
    j1[j1 == df] = 0                                              # splines.py:165-166
    basis = ajm * eye[j, :].T + ajp * eye[j1, :].T + cjm * F[j, :].T + cjp * F[j1, :].T
    return unit_frobenius(basis.T)


def tp(x, df):
    """Truncated thin-plate regression spline basis (splines.py:174-189)."""
    x = np.asarray(x)
    r = np.abs(x.ravel() - x.ravel().reshape(-1, 1))
    E = r ** 2 * np.log(r + 1e-10)
    U, _, _ = np.linalg.svd(E)
    return unit_frobenius(U[:, : int(df)])


_BASES = {"cr": cr, "cc": cc, "bs": bs, "tp": tp}


def spline_factors(dims, df, smooth):
    """The per-dimension 1-D bases on the grids `arange(d)` (splines.py:266-298)."""
    if len(df) != len(dims):
        raise ValueError("`df` must have the same length as `dims`")
    if smooth not in _BASES:
        raise ValueError("Input method `{}` is not supported.".format(smooth))
    if len(dims) not in (1, 2, 3):
        raise NotImplementedError(len(dims))
    return [_BASES[smooth](np.arange(d).ravel(), f) for d, f in zip(dims, df)]


def build_spline_matrix(dims, df, smooth, dtype=np.float64):
    """kron(St, kron(Sx, Sy)) re-normalised and cast (splines.py:278-305)."""
    factors = spline_factors(dims, df, smooth)
    S = factors[0]
    if len(factors) == 2:
        S = np.kron(factors[0], factors[1])
    elif len(factors) == 3:
        S = np.kron(factors[0], np.kron(factors[1], factors[2]))
    return unit_frobenius(S).astype(dtype)
