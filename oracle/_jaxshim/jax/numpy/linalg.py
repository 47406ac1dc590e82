"""`jax.numpy.linalg` stand-in (norm is differentiable; the rest is NumPy's).
TEST INFRASTRUCTURE ONLY."""

import numpy as _np
from numpy.linalg import inv, pinv, lstsq, solve, svd, eigh, det, slogdet, cholesky  # noqa: F401

from .. import _autodiff as _ad


def norm(x, ord=None):  # noqa: A002
    """Vector norms as `jnp.linalg.norm` builds them: ord=1 -> sum(abs(x)),
    ord=2/None -> sqrt(sum(x*x)) (gradient x/||x||, NaN at 0 like JAX's)."""
    if not _ad.is_tracer(x):
        with _np.errstate(all="ignore"):
            return _np.linalg.norm(x, ord)
    if x.ndim != 1:
        raise NotImplementedError("shim norm: vectors only")
    if ord == 1:
        return _ad.sum_(_ad.abs_(x))
    if ord in (None, 2):
        return _ad.sqrt(_ad.sum_(_ad.mul(x, x)))
    raise NotImplementedError(f"shim norm: ord={ord}")
