"""`jax.random` stand-in.  TEST INFRASTRUCTURE ONLY.

NOT threefry: `normal(PRNGKey(s), shape)` draws `default_rng(s).standard_normal(shape)`
(what oracle/glm.py::_initialize_random does), so randomly initialised parameters differ
from real JAX's.  Parity runs never depend on it: they inject an explicit `p0`."""

import numpy as np


class _Key:
    def __init__(self, seed):
        self.seed = int(seed)


def PRNGKey(seed):
    return _Key(seed)


def normal(key, shape=(), dtype=np.float64):
    return np.random.default_rng(key.seed).standard_normal(shape).astype(dtype)
