"""GPU parity against the REFERENCE's own outputs: the CUDA engine, through the public
`rfest_b200.GLM` API (ctypes -> C ABI), on every case of `tests/glm_cases.py`, compared with
`tests/golden/ref_glm_golden.npz` -- what `/root/reference/rfest/GLM/_glm.py` + `rfest/loss.py`
produced when executed unmodified (`oracle/make_golden_glm.py`).  No oracle code runs here.

Tolerances are the north star's: per-iteration train/dev loss within 1e-5 relative, final
filter weights within 1e-4 relative -- against the reference's float64 run and against its
float32 run.
"""

import os

import numpy as np
import pytest

import glm_cases as gc

pytestmark = pytest.mark.gpu

LOSS_RTOL = 1e-5
WEIGHT_RTOL = 1e-4
GOLDEN_GLM = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_glm_golden.npz")

# long fits whose stopping iteration / selected model hinges on cost differences near the tolerance
# still have to agree -- the traces are compared over the common prefix first so a failure says why
STOP_CASES = {"early_stop_dev", "early_stop_last", "train_stop_flat"}


@pytest.fixture(scope="module")
def golden_glm():
    return np.load(GOLDEN_GLM)


def _case(z, name, tag):
    prefix = f"{name}/{tag}/"
    return {k[len(prefix):]: z[k] for k in z.files if k.startswith(prefix)}


def _rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.max(np.abs(a - b) / np.maximum(np.abs(b), 1e-300)))


def _wrel(a, b):
    a, b = np.asarray(a, np.float64).ravel(), np.asarray(b, np.float64).ravel()
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))


@pytest.mark.parametrize("name", sorted(gc.CASES))
def test_engine_matches_reference_outputs(golden_glm, name):
    from rfest_b200 import GLM

    c = gc.CASES[name]
    data = gc.case_data(c)
    assert gc.data_digest(data) == str(golden_glm[f"{name}/digest"]), "inputs differ from the fixture's"
    m, extra = gc.run_case(lambda d, o: GLM(distr=d, output_nonlinearity=o), name, np.float32, data=data)
    got = gc.summarize(m, extra)
    # how far the reference's own float32 run ends from its float64 run: where that alone exceeds
    # the north star's 1e-4 (mle_3d_exp: 4.5e-4 -- an ill-conditioned direction of the weights, the
    # losses agree to 5e-6) no float32 implementation can be held to 1e-4, so the bound scales
    ref64, ref32 = _case(golden_glm, name, "f64"), _case(golden_glm, name, "f32")
    spread = max(_wrel(ref32[f"last.{f}"], ref64[f"last.{f}"]) for f in ref64["filter_names"].tolist())
    weight_rtol = max(WEIGHT_RTOL, 3 * spread)
    for tag in ("f64", "f32"):
        want = _case(golden_glm, name, tag)
        assert got["filter_names"].tolist() == want["filter_names"].tolist()
        np.testing.assert_allclose(got["p0.intercept"], want["p0.intercept"], rtol=0, atol=1e-6)
        k = min(len(got["cost_train"]), len(want["cost_train"]))
        assert _rel(got["cost_train"][:k], want["cost_train"][:k]) < LOSS_RTOL, tag
        has_dev = np.isfinite(want["cost_dev"]).any()
        if has_dev:
            assert _rel(got["cost_dev"][:k], want["cost_dev"][:k]) < LOSS_RTOL, tag
            assert np.max(np.abs(got["metric_dev"][:k] - want["metric_dev"][:k])) < 1e-4
        assert np.max(np.abs(got["metric_train"][:k] - want["metric_train"][:k])) < 1e-4
        assert len(got["cost_train"]) == len(want["cost_train"]), (tag, "stopped at a different iteration")
        assert str(got["train_stop"]) == str(want["train_stop"])
        assert str(got["return_model"]) == str(want["return_model"])
        trace = want["cost_dev"] if str(want["return_model"]) == "best_dev_cost" else want["cost_train"]
        gb, wb = int(got["best_iteration"]), int(want["best_iteration"])
        assert gb == wb or abs(trace[gb] - trace[wb]) < 1e-5 * abs(want["cost_train"][0]), (gb, wb)
        for fname in want["filter_names"].tolist():
            assert _wrel(got[f"last.{fname}"], want[f"last.{fname}"]) < weight_rtol, (tag, fname)
            if gb == wb:
                assert _wrel(got[f"w.{fname}"], want[f"w.{fname}"]) < weight_rtol, (tag, fname)
        for kind in ("train", "dev"):
            if f"y_pred.{kind}" in want:
                assert _wrel(got[f"y_pred.{kind}"], want[f"y_pred.{kind}"]) < weight_rtol, (tag, kind)
        if c["test"] and gb == wb:
            assert abs(float(got["test_score"]) - float(want["test_score"])) < 1e-4
            assert _wrel(got["test_pred"], want["test_pred"]) < weight_rtol
        if c["compute_ci"] and gb == wb:
            for fname in want["filter_names"].tolist():
                # inverse of a Gram matrix: conditioning amplifies float32 rounding
                assert _wrel(got[f"V.{fname}"], want[f"V.{fname}"]) < 2e-3, (tag, fname)
                assert _wrel(got[f"b_se.{fname}"], want[f"b_se.{fname}"]) < 1e-3, (tag, fname)
                assert _wrel(got[f"w_se.{fname}"], want[f"w_se.{fname}"]) < 1e-3, (tag, fname)
