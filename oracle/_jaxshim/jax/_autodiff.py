"""Reverse-mode autodiff over NumPy values -- the engine behind the test-only `jax` shim.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py and oracle/_jaxshim/jax/__init__.py).

A `Tracer` wraps a NumPy value (or a Python float: a weak-typed leaf) and records,
for every operation the reference's `GLM.cost` performs on it, the vector-Jacobian
product of that operation.  Type promotion is NumPy's (NEP 50), which for the
operations used here coincides with JAX's under `jax_enable_x64`: Python scalars are
weak, float32 array + float64 array -> float64, and the cotangent of a promoted
operand is cast back to the operand's dtype (the transpose of
`convert_element_type`).  Tie/kink conventions follow JAX: `maximum` splits the
cotangent 0.5/0.5 at equality, `abs'(0) = 0`, `where` routes the cotangent to the
selected branch only.
"""

import itertools

import numpy as np

_counter = itertools.count()


def is_tracer(x):
    return isinstance(x, Tracer)


def val(x):
    return x.value if isinstance(x, Tracer) else x


def _dtype(v):
    if isinstance(v, (float, int)) and not isinstance(v, np.generic):
        return np.dtype(np.float64)          # weak Python scalar: float64 under x64
    return np.asarray(v).dtype


def _shape(v):
    return np.shape(v)


def _fit(ct, like):
    """Cotangent of a (broadcast, promoted) operand: sum the broadcast axes, cast back."""
    shape = _shape(like)
    ct = np.asarray(ct)
    if ct.shape != shape:
        extra = ct.ndim - len(shape)
        if extra > 0:
            ct = ct.sum(axis=tuple(range(extra)))
        axes = tuple(i for i, s in enumerate(shape) if s == 1 and ct.shape[i] != 1)
        if axes:
            ct = ct.sum(axis=axes, keepdims=True)
        ct = ct.reshape(shape)
    dt = _dtype(like)
    if np.issubdtype(dt, np.floating) and ct.dtype != dt:
        ct = ct.astype(dt)
    return ct


class Tracer:
    __array_ufunc__ = None        # ndarray (op) Tracer defers to the reflected method
    __array_priority__ = 1000
    __hash__ = object.__hash__

    def __init__(self, value, parents=()):
        self.value = value
        self.parents = parents
        self.id = next(_counter)

    # ---- array attributes -------------------------------------------------
    @property
    def shape(self):
        return _shape(self.value)

    @property
    def ndim(self):
        return len(self.shape)

    @property
    def size(self):
        return int(np.prod(self.shape, dtype=np.int64))

    @property
    def dtype(self):
        return _dtype(self.value)

    def __len__(self):
        return self.shape[0]

    @property
    def T(self):
        return transpose(self)

    # ---- comparisons: plain boolean arrays, nothing to differentiate ---------
    def __gt__(self, o):
        return np.asarray(self.value) > val(o)

    def __ge__(self, o):
        return np.asarray(self.value) >= val(o)

    def __lt__(self, o):
        return np.asarray(self.value) < val(o)

    def __le__(self, o):
        return np.asarray(self.value) <= val(o)

    def __eq__(self, o):
        return np.asarray(self.value) == val(o)

    def __ne__(self, o):
        return np.asarray(self.value) != val(o)

    # ---- arithmetic ------------------------------------------------------------
    def __add__(self, o):
        return add(self, o)

    def __radd__(self, o):
        return add(o, self)

    def __sub__(self, o):
        return sub(self, o)

    def __rsub__(self, o):
        return sub(o, self)

    def __mul__(self, o):
        return mul(self, o)

    def __rmul__(self, o):
        return mul(o, self)

    def __truediv__(self, o):
        return div(self, o)

    def __rtruediv__(self, o):
        return div(o, self)

    def __neg__(self):
        return neg(self)

    def __pow__(self, k):
        return power(self, k)

    def __matmul__(self, o):
        return matmul(self, o)

    def __rmatmul__(self, o):
        return matmul(o, self)

    def __getitem__(self, idx):
        return getitem(self, idx)

    # ---- methods the reference calls ----------------------------------------------
    def astype(self, dtype):
        return astype(self, dtype)

    def flatten(self):
        return reshape(self, (-1,))

    def ravel(self):
        return reshape(self, (-1,))

    def reshape(self, *shape):
        if len(shape) == 1 and isinstance(shape[0], (tuple, list)):
            shape = tuple(shape[0])
        return reshape(self, shape)

    def sum(self, axis=None):
        return sum_(self, axis)

    def mean(self, axis=None):
        return mean(self, axis)

    def squeeze(self, axis=None):
        return squeeze(self, axis)

    def __repr__(self):
        return f"Tracer({self.value!r})"


def _node(value, *edges):
    """edges: (operand, vjp) pairs; operands that are not tracers are dropped."""
    return Tracer(value, tuple((p, f) for p, f in edges if isinstance(p, Tracer)))


def _np(fn, *args, **kw):
    with np.errstate(all="ignore"):
        return fn(*args, **kw)


# ---- binary ------------------------------------------------------------------------
def add(a, b):
    x, y = val(a), val(b)
    return _node(_np(np.add, x, y), (a, lambda ct: _fit(ct, x)), (b, lambda ct: _fit(ct, y)))


def sub(a, b):
    x, y = val(a), val(b)
    return _node(_np(np.subtract, x, y), (a, lambda ct: _fit(ct, x)),
                 (b, lambda ct: _fit(_np(np.negative, ct), y)))


def mul(a, b):
    x, y = val(a), val(b)
    return _node(_np(np.multiply, x, y), (a, lambda ct: _fit(_np(np.multiply, ct, y), x)),
                 (b, lambda ct: _fit(_np(np.multiply, x, ct), y)))


def div(a, b):
    """x / y.  JAX: d/dx = ct / y ; d/dy = -ct * x / y**2 (mul then div by the square)."""
    x, y = val(a), val(b)
    return _node(_np(np.divide, x, y), (a, lambda ct: _fit(_np(np.divide, ct, y), x)),
                 (b, lambda ct: _fit(_np(np.multiply, _np(np.negative, ct),
                                         _np(np.divide, _np(np.divide, x, y), y)), y)))


def neg(a):
    return _node(_np(np.negative, val(a)), (a, lambda ct: _np(np.negative, ct)))


def power(a, k):
    if is_tracer(k):
        raise NotImplementedError("shim: traced exponents are not needed by the GLM path")
    x = val(a)
    return _node(_np(np.power, x, k),
                 (a, lambda ct: _fit(_np(np.multiply, ct, k * _np(np.power, x, k - 1)), x)))


def matmul(a, b):
    x, y = np.asarray(val(a)), np.asarray(val(b))
    out = _np(np.matmul, x, y)

    def two_d(ct):
        ct = np.asarray(ct)
        x2 = x[np.newaxis, :] if x.ndim == 1 else x
        y2 = y[:, np.newaxis] if y.ndim == 1 else y
        # contiguous: NumPy picks its BLAS / non-BLAS matmul loop (= summation order) by strides
        return np.ascontiguousarray(ct.reshape(x2.shape[0], y2.shape[1])), x2, y2

    def vjp_a(ct):
        ct2, x2, y2 = two_d(ct)
        return _fit(_np(np.matmul, ct2, y2.T).reshape(x.shape), x)

    def vjp_b(ct):
        ct2, x2, y2 = two_d(ct)
        return _fit(_np(np.matmul, x2.T, ct2).reshape(y.shape), y)

    return _node(out, (a, vjp_a), (b, vjp_b))


def maximum(a, b):
    x, y = val(a), val(b)
    out = _np(np.maximum, x, y)
    wa = _np(np.where, np.asarray(x) > y, 1.0, _np(np.where, np.asarray(x) == y, 0.5, 0.0))
    return _node(out, (a, lambda ct: _fit(ct * wa.astype(np.asarray(ct).dtype), x)),
                 (b, lambda ct: _fit(ct * (1.0 - wa).astype(np.asarray(ct).dtype), y)))


def where(c, a, b):
    c = np.asarray(val(c))
    x, y = val(a), val(b)
    out = _np(np.where, c, x, y)
    zero = lambda ct: np.zeros((), dtype=np.asarray(ct).dtype)
    return _node(out, (a, lambda ct: _fit(np.where(c, ct, zero(ct)), x)),
                 (b, lambda ct: _fit(np.where(c, zero(ct), ct), y)))


# ---- unary ---------------------------------------------------------------------------
def exp(a):
    out = _np(np.exp, val(a))
    return _node(out, (a, lambda ct: _np(np.multiply, ct, out)))


def log(a):
    x = val(a)
    return _node(_np(np.log, x), (a, lambda ct: _np(np.divide, ct, x)))


def tanh(a):
    out = _np(np.tanh, val(a))
    return _node(out, (a, lambda ct: _np(np.multiply, ct, 1 - out * out)))


def sqrt(a):
    out = _np(np.sqrt, val(a))
    return _node(out, (a, lambda ct: _np(np.divide, ct, 2 * out)))


def abs_(a):
    x = val(a)
    return _node(_np(np.abs, x), (a, lambda ct: _np(np.multiply, ct, np.sign(x))))


def square(a):
    x = val(a)
    return _node(_np(np.square, x), (a, lambda ct: _np(np.multiply, ct, 2 * x)))


# ---- shape ---------------------------------------------------------------------------
def squeeze(a, axis=None):
    x = np.asarray(val(a))
    return _node(np.squeeze(x, axis=axis), (a, lambda ct: np.asarray(ct).reshape(x.shape)))


def reshape(a, shape):
    x = np.asarray(val(a))
    return _node(x.reshape(shape), (a, lambda ct: np.asarray(ct).reshape(x.shape)))


def transpose(a):
    return _node(np.asarray(val(a)).T, (a, lambda ct: np.asarray(ct).T))


def getitem(a, idx):
    x = np.asarray(val(a))

    def vjp(ct):
        g = np.zeros(x.shape, dtype=np.asarray(ct).dtype)
        np.add.at(g, idx, ct)
        return g

    return _node(x[idx], (a, vjp))


def astype(a, dtype):
    x = np.asarray(val(a))
    return _node(x.astype(dtype), (a, lambda ct: np.asarray(ct).astype(x.dtype)))


def sum_(a, axis=None):
    x = np.asarray(val(a))

    def vjp(ct):
        ct = np.asarray(ct)
        if axis is not None:
            ct = np.expand_dims(ct, axis)
        return np.broadcast_to(ct, x.shape).astype(ct.dtype)

    return _node(_np(np.sum, x, axis=axis), (a, vjp))


def mean(a, axis=None):
    x = np.asarray(val(a))
    n = x.size if axis is None else x.shape[axis]
    return div(sum_(a, axis), n)


def stack(items):
    """`jnp.array([t0, t1, ...])` of equal-shaped operands."""
    vals = [np.asarray(val(t)) for t in items]
    out = np.array(vals)
    edges = [(t, (lambda ct, k=k, v=v: _fit(np.asarray(ct)[k], v)))
             for k, (t, v) in enumerate(zip(items, vals))]
    return _node(out, *edges)


def concatenate(items, axis=0):
    vals = [np.atleast_1d(np.asarray(val(t))) for t in items]
    out = np.concatenate(vals, axis=axis)
    edges, pos = [], 0
    for t, v in zip(items, vals):
        k = v.shape[axis]
        sl = [slice(None)] * out.ndim
        sl[axis] = slice(pos, pos + k)
        edges.append((t, (lambda ct, sl=tuple(sl), v=v: _fit(np.asarray(ct)[sl].reshape(v.shape), v))))
        pos += k
    return _node(out, *edges)


# ---- backward ---------------------------------------------------------------------------
def backward(out, leaves):
    """Cotangents of `leaves` for d out (a scalar tracer), nodes visited newest first."""
    reach, stack_ = {}, [out]
    while stack_:
        n = stack_.pop()
        if n.id in reach:
            continue
        reach[n.id] = n
        stack_.extend(p for p, _ in n.parents)
    cts = {out.id: np.ones(_shape(out.value), dtype=_dtype(out.value))}
    for nid in sorted(reach, reverse=True):
        ct = cts.get(nid)
        if ct is None:
            continue
        for parent, vjp in reach[nid].parents:
            g = vjp(ct)
            cts[parent.id] = g if parent.id not in cts else cts[parent.id] + g
    res = []
    for leaf in leaves:
        g = cts.get(leaf.id)
        if g is None:
            g = np.zeros(_shape(leaf.value), dtype=_dtype(leaf.value))
        res.append(_fit(g, leaf.value))
    return res
