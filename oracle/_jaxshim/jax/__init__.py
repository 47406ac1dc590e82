"""A test-only stand-in for the `jax` package, just wide enough to run the reference's
`rfest/GLM/_glm.py` and `rfest/loss.py` UNMODIFIED on NumPy.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  Real JAX is neither installed nor
installable in this image (no network, no wheel), so the reference's GLM cannot execute
here as shipped.  This package provides the ~20 names it imports
(`rfest/GLM/_glm.py:3-11`, `rfest/loss.py:1`):

  jax.numpy            NumPy, with the functions `GLM.cost` applies to parameters lifted
                       to a reverse-mode tape (`_autodiff.py`)
  jax.value_and_grad   value and gradient over a pytree of parameters (the tape)
  jax.jit              identity (plus the `jax_debug_nans` output check)
  jax.config           records flags; `jax_debug_nans` makes NaN outputs raise
                       FloatingPointError as JAX does
  jax.random           PRNGKey / normal on NumPy's generator -- NOT threefry: random
                       draws differ from JAX's, so parity runs share an explicit p0
  jax.example_libraries.optimizers.adam
                       the published recurrence (see that module)

It models `jax_enable_x64 = True`, which `import rfest` switches on
(`rfest/GLM/_base.py:23` via `rfest/__init__.py:4`): float64 is a real dtype and Python
floats are weak-typed.  What it cannot reproduce is XLA's instruction selection /
summation order (ulp level) and threefry.

It is put on `sys.path` only by `oracle/ref_glm.py`, only when no real `jax` is
importable, and never by anything under `rfest_b200/`.
"""

import numpy as _np

from . import _autodiff as _ad
from . import tree_util  # noqa: F401
from .tree_util import tree_flatten, tree_unflatten, tree_map  # noqa: F401

__version__ = "0.0-shim"
IS_RFEST_B200_TEST_SHIM = True


class Array:
    """Nothing is ever an instance: SciPy / scikit-learn probe `jax.Array` and `jax.float0`
    once a module called `jax` is loaded (array-API compatibility layer)."""


float0 = _np.dtype([("float0", _np.void, 0)])


class _Config:
    def __init__(self):
        self.values = {"jax_enable_x64": True, "jax_debug_nans": False}

    def update(self, name, value):
        if name == "jax_enable_x64" and not value:
            raise NotImplementedError("the shim only models jax_enable_x64=True")
        self.values[name] = value

    def __getattr__(self, name):
        try:
            return self.__dict__["values"][name]
        except KeyError:
            raise AttributeError(name)


config = _Config()


def _check_nans(tree, where):
    if not config.values.get("jax_debug_nans"):
        return
    for leaf in tree_flatten(tree)[0]:
        a = _np.asarray(_ad.val(leaf))
        if a.dtype.kind == "f" and _np.isnan(a).any():
            raise FloatingPointError(f"invalid value (nan) encountered in {where}")


def jit(fun=None, **_kw):
    """Identity.  Like JAX under `jax_debug_nans`, the outputs of the jitted function
    are checked for NaN."""
    if fun is None:
        return lambda f: jit(f)

    def wrapped(*args, **kwargs):
        out = fun(*args, **kwargs)
        _check_nans(out, getattr(fun, "__name__", "jit"))
        return out

    wrapped.__wrapped__ = fun
    return wrapped


def value_and_grad(fun, argnums=0):
    """value_and_grad over the pytree in positional argument `argnums` (an int)."""

    def wrapped(*args, **kwargs):
        leaves, treedef = tree_flatten(args[argnums])
        tracers = [_ad.Tracer(v if isinstance(v, (float, int)) and not isinstance(v, _np.generic)
                              else _np.asarray(v)) for v in leaves]
        call = list(args)
        call[argnums] = tree_unflatten(treedef, tracers)
        out = fun(*call, **kwargs)
        if not _ad.is_tracer(out):
            raise TypeError("shim value_and_grad: the output does not depend on the parameters")
        if _np.shape(out.value) != ():
            raise TypeError("value_and_grad needs a scalar output")
        grads = _ad.backward(out, tracers)
        grads = [g if _np.ndim(g) else _np.asarray(g)[()] for g in grads]
        return out.value, tree_unflatten(treedef, grads)

    return wrapped


def grad(fun, argnums=0):
    vg = value_and_grad(fun, argnums)
    return lambda *a, **k: vg(*a, **k)[1]


from . import numpy  # noqa: E402,F401
from . import random  # noqa: E402,F401
