"""Run the reference's own `rfest/GLM/_glm.py` + `rfest/loss.py`, UNMODIFIED, as the
reference-executed oracle.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

The two files are imported from where they lie under /root/reference; the only thing
supplied from this repository is the test-only `jax` package of `oracle/_jaxshim/`
(NumPy + a reverse-mode tape + the published Adam), which is put on `sys.path` only
when no real `jax` can be imported.  With real JAX present the same loader returns the
reference on real JAX.

`/root/reference` exists only in the build container.  On the GPU box `available()` is
False and the tests use the fixtures this reference produced
(`tests/golden/ref_glm_*.npz`, written by `oracle/make_golden_glm.py`).
"""

import importlib
import importlib.util
import os
import sys
import types

from . import ref_loader

SHIM_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_jaxshim")


def available():
    return ref_loader.available()


def _ensure_jax():
    """Real JAX if it exists, else the shim.  Returns 'jax' or 'shim'."""
    mod = sys.modules.get("jax")
    if mod is None:
        if importlib.util.find_spec("jax") is None:
            sys.path.insert(0, SHIM_DIR)
        mod = importlib.import_module("jax")
    return "shim" if getattr(mod, "IS_RFEST_B200_TEST_SHIM", False) else "jax"


def load():
    """Return (GLM class, loss module, backend) of the reference, backend in {'shim','jax'}."""
    if not available():
        raise RuntimeError(f"reference tree not found under {ref_loader.REFERENCE_ROOT}")
    backend = _ensure_jax()
    utils, splines, metrics = ref_loader.load()          # registers the bare `rfest` package
    pkg = sys.modules["rfest"]
    if "rfest.GLM" not in sys.modules:                     # bare sub-package: skip GLM/__init__.py
        sub = types.ModuleType("rfest.GLM")                # (it pulls the legacy classes and matplotlib)
        sub.__path__ = [os.path.join(ref_loader.REFERENCE_ROOT, "rfest", "GLM")]
        sys.modules["rfest.GLM"] = sub
        pkg.GLM_pkg = sub
    loss = importlib.import_module("rfest.loss")
    glm = importlib.import_module("rfest.GLM._glm")
    # what `from rfest import ...` gives the reference's helpers and tests
    pkg.build_design_matrix = utils.build_design_matrix
    pkg.build_spline_matrix = splines.build_spline_matrix
    pkg.GLM = glm.GLM
    return glm.GLM, loss, backend
