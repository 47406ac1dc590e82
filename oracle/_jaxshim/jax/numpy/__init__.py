"""`jax.numpy` stand-in: NumPy, with the functions the reference applies to parameters
inside `value_and_grad` dispatching to the tape when an operand is traced.
TEST INFRASTRUCTURE ONLY (see ../__init__.py)."""

import builtins as _bi

import numpy as _np
from numpy import *  # noqa: F401,F403
from numpy import float32, float64, int32, int64, newaxis, inf, nan, pi  # noqa: F401

from .. import _autodiff as _ad
import importlib as _il

linalg = _il.import_module(__name__ + ".linalg")   # not `from . import`: numpy's own `linalg` came in with *

ndarray = _np.ndarray


def _any_tracer(*xs):
    return _bi.any(_ad.is_tracer(x) for x in xs)


def _quiet(fn):
    def f(*a, **k):
        with _np.errstate(all="ignore"):
            return fn(*a, **k)
    f.__name__ = fn.__name__
    return f


def _unary(np_fn, ad_fn):
    q = _quiet(np_fn)

    def f(x):
        return ad_fn(x) if _ad.is_tracer(x) else q(x)
    f.__name__ = np_fn.__name__
    return f


exp = _unary(_np.exp, _ad.exp)
log = _unary(_np.log, _ad.log)
tanh = _unary(_np.tanh, _ad.tanh)
sqrt = _unary(_np.sqrt, _ad.sqrt)
abs = absolute = _unary(_np.abs, _ad.abs_)
square = _unary(_np.square, _ad.square)


def where(c, x, y):
    if _any_tracer(c, x, y):
        return _ad.where(c, x, y)
    return _np.where(c, x, y)


def maximum(x, y):
    if _any_tracer(x, y):
        return _ad.maximum(x, y)
    return _quiet(_np.maximum)(x, y)


def squeeze(x, axis=None):
    return _ad.squeeze(x, axis) if _ad.is_tracer(x) else _np.squeeze(x, axis=axis)


def sum(x, axis=None):  # noqa: A001
    return _ad.sum_(x, axis) if _ad.is_tracer(x) else _quiet(_np.sum)(x, axis=axis)


def mean(x, axis=None):
    return _ad.mean(x, axis) if _ad.is_tracer(x) else _quiet(_np.mean)(x, axis=axis)


def array(x, dtype=None):
    if isinstance(x, (list, tuple)) and _any_tracer(*x):
        out = _ad.stack(list(x))
        return out if dtype is None else out.astype(dtype)
    if _ad.is_tracer(x):
        return x if dtype is None else x.astype(dtype)
    return _np.array(x, dtype=dtype)


def asarray(x, dtype=None):
    if _ad.is_tracer(x):
        return x if dtype is None else x.astype(dtype)
    return _np.asarray(x, dtype=dtype)


def hstack(xs):
    xs = list(xs)
    if _any_tracer(*xs):
        nd = _bi.max(_np.ndim(_ad.val(t)) for t in xs)
        return _ad.concatenate(xs, axis=0 if nd <= 1 else 1)
    return _np.hstack(xs)


def matmul(a, b):
    return _ad.matmul(a, b) if _any_tracer(a, b) else _np.matmul(a, b)


def dot(a, b):
    return _ad.matmul(a, b) if _any_tracer(a, b) else _np.dot(a, b)
