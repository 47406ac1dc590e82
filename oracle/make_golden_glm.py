"""Generate `tests/golden/ref_glm_golden.npz`: outputs of the REFERENCE's own
`rfest/GLM/_glm.py` + `rfest/loss.py`, executed unmodified (`oracle/ref_glm.py`: real JAX
if it exists, else the test-only NumPy `jax` shim of `oracle/_jaxshim/`).

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  Run in the build container, the only
place `/root/reference` exists:

    python -m oracle.make_golden_glm

For every named case of `tests/glm_cases.py` (eight model configurations, the README /
BASELINE config-1 recipe, early stopping and model selection, a step-size schedule, the
least-squares initialisation recipes of the reference's tests/test_glm.py, confidence
intervals, test-set prediction) the reference is fitted in float64 and in float32 and what
it leaves behind is stored: the four traces, stop reason, selected iteration, initial / last /
selected parameters, filters `w`, predictions, and with `compute_ci` the covariances, standard
errors and response bands.  Inputs are regenerated from seeds by the tests; their SHA-256
prefix is stored so a platform that regenerates different inputs fails loudly.

Everything stored here was computed by reference code.  The archive travels to the GPU box
with the repository; the reference does not.
"""

import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "tests", "golden", "ref_glm_golden.npz")


def main():
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import glm_cases as gc
    import scipy

    from . import ref_glm

    GLM, _loss, backend = ref_glm.load()
    store = {"env": np.array(f"backend={backend} numpy={np.__version__} scipy={scipy.__version__}"),
             "cases": np.array(sorted(gc.CASES))}
    for name in sorted(gc.CASES):
        data = gc.case_data(gc.CASES[name])
        store[f"{name}/digest"] = np.array(gc.data_digest(data))
        for tag, dtype in (("f64", np.float64), ("f32", np.float32)):
            def make(distr, out_nl, dtype=dtype):
                return GLM(distr=distr, output_nonlinearity=out_nl, dtype=dtype)

            model, extra = gc.run_case(make, name, dtype, reference_quirks=True, data=data)
            for key, val in gc.summarize(model, extra).items():
                if tag == "f32" and val.dtype == np.float64 and val.size > 64:
                    val = val.astype(np.float32)        # long vectors: the dtype they were computed in
                store[f"{name}/{tag}/{key}"] = val
            print(f"{name:32s} {tag} iters {len(model.cost_train):4d} {model.train_stop:13s} "
                  f"best {model.best_iteration:4d} cost {model.cost_train[0]:.6g} -> {model.cost_train[-1]:.6g}")
    np.savez_compressed(OUT, **store)
    print(f"wrote {OUT}: {len(store)} arrays, {os.path.getsize(OUT) / 1024:.0f} KB ({backend})")


if __name__ == "__main__":
    main()
