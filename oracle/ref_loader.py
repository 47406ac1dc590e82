"""Import the reference's NumPy-only modules without importing JAX.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

`import rfest` fails in this image because `rfest/__init__.py` pulls in JAX.
`rfest.utils`, `rfest.splines` and `rfest.metrics` only need NumPy / SciPy /
scikit-learn, so they import fine when `rfest` is registered as a bare
namespace package first (probe 1).

The reference tree only exists in the build container (`/root/reference`);
on the GPU box `available()` is False and callers must fall back to the
committed golden vectors under `tests/golden/`.
"""

import importlib
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("RFEST_REFERENCE_ROOT", "/root/reference")


def available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "rfest"))


def load():
    """Return (utils, splines, metrics) modules of the reference."""
    if not available():
        raise RuntimeError(f"reference tree not found under {REFERENCE_ROOT}")
    name = "rfest"
    existing = sys.modules.get(name)
    if existing is None or not hasattr(existing, "__path__"):
        pkg = types.ModuleType(name)
        pkg.__path__ = [os.path.join(REFERENCE_ROOT, "rfest")]
        sys.modules[name] = pkg
    utils = importlib.import_module("rfest.utils")
    splines = importlib.import_module("rfest.splines")
    metrics = importlib.import_module("rfest.metrics")
    return utils, splines, metrics
