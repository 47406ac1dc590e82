"""`jax.example_libraries.optimizers.adam` for the shim.  TEST INFRASTRUCTURE ONLY.

The real module is not part of /root/reference (JAX is an unpinned dependency,
`pyproject.toml:19`); the reference calls it at `rfest/GLM/_glm.py:698-699,694-696,723`.
Restated from the published recurrence:

    init(x0):          m0 = v0 = zeros_like(x0)
    update(i, g, (x, m, v)):
        m = (1 - b1) * g + b1 * m
        v = (1 - b2) * g**2 + b2 * v
        mhat = m / (1 - asarray(b1, m.dtype) ** (i + 1))
        vhat = v / (1 - asarray(b2, m.dtype) ** (i + 1))
        x = x - step_size(i) * mhat / (sqrt(vhat) + eps)

applied leaf by leaf over the parameter pytree; `step_size` is a float or a schedule
`i -> float`.  Python-float leaves (the GLM's intercepts) keep a float64 state.
"""

import numpy as np

from ..tree_util import tree_flatten, tree_unflatten


def constant(step_size):
    return lambda i: step_size


def make_schedule(scalar_or_schedule):
    if callable(scalar_or_schedule):
        return scalar_or_schedule
    if np.ndim(scalar_or_schedule) == 0:
        return constant(scalar_or_schedule)
    raise TypeError(type(scalar_or_schedule))


class OptimizerState:
    def __init__(self, packed_state, treedef):
        self.packed_state = packed_state
        self.treedef = treedef


def adam(step_size, b1=0.9, b2=0.999, eps=1e-8):
    step_size = make_schedule(step_size)

    def init(x0):
        return x0, np.zeros_like(x0), np.zeros_like(x0)

    def update(i, g, state):
        x, m, v = state
        with np.errstate(all="ignore"):
            m = (1 - b1) * g + b1 * m
            v = (1 - b2) * np.square(g) + b2 * v
            mhat = m / (1 - np.asarray(b1, m.dtype) ** (i + 1))
            vhat = v / (1 - np.asarray(b2, m.dtype) ** (i + 1))
            x = x - step_size(i) * mhat / (np.sqrt(vhat) + eps)
        return x, m, v

    def tree_init(x0_tree):
        leaves, treedef = tree_flatten(x0_tree)
        return OptimizerState([init(x) for x in leaves], treedef)

    def tree_update(i, grad_tree, opt_state):
        grads, treedef = tree_flatten(grad_tree)
        if len(grads) != len(opt_state.packed_state):
            raise TypeError("optimizer update: gradient tree does not match the parameters")
        return OptimizerState([update(i, g, s) for g, s in zip(grads, opt_state.packed_state)],
                              opt_state.treedef)

    def tree_get_params(opt_state):
        return tree_unflatten(opt_state.treedef, [s[0] for s in opt_state.packed_state])

    return tree_init, tree_update, tree_get_params
