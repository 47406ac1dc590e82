"""Multi-GPU parity: a fit time-sharded over 2 (and 4) GPUs with the in-graph NCCL all-reduce must
reproduce the single-GPU fit (only the summation order differs) and the oracle."""

import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _n_gpus():
    import torch

    return torch.cuda.device_count()


@pytest.mark.parametrize("world,xchg,out_nl,filt_nl", [(2, 1, "softplus", "none"), (2, 0, "softplus", "none"),
                                                       (2, 1, "softmax", "none"), (2, 1, "softplus", "softmax"),
                                                       (2, 3, "softplus", "none"),
                                                       (4, 1, "softplus", "none"), (8, 1, "softplus", "none")])
def test_sharded_fit_matches_single_gpu(world, xchg, out_nl, filt_nl, tmp_path):
    """xchg=1: gradient exchange by the fused one-shot NVLink kernel (3: its one-CTA variant); xchg=0: NCCL all-reduce.  With a
    softmax output the normaliser and the Jacobian scalar are all-reduced inside the iteration as well; with a
    softmax FILTER the per-filter normalisers and the vectors of the Jacobian correction are (NCCL path)."""
    if _n_gpus() < world:
        pytest.skip(f"needs {world} GPUs")
    out = tmp_path / "dist.npz"
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(29500 + world),
           os.path.join(ROOT, "tests", "dist_fit_worker.py"), str(out)]
    env = dict(os.environ, RFB_XCHG=str(xchg), RFB_TEST_OUT_NL=out_nl, RFB_TEST_FILT_NL=filt_nl)
    proc = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env)
    assert proc.returncode == 0, proc.stdout[-3000:] + proc.stderr[-3000:]
    got = np.load(out)

    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import helpers
    from dist_fit_worker import build_and_fit
    from oracle.glm import OracleGLM
    from rfest_b200 import GLM

    stim, y, _ = helpers.lnp_data(5003, (8, 5, 4), seed=31)
    one = build_and_fit(GLM, stim, y, 4001, 40, device=0, out_nl=out_nl, filt_nl=filt_nl)
    rel = lambda a, b: np.max(np.abs(np.asarray(a) - np.asarray(b)) / np.abs(np.asarray(b)))  # noqa: E731
    assert rel(got["cost_train"], one.cost_train) < 1e-6
    assert rel(got["cost_dev"], one.cost_dev) < 1e-6
    assert np.max(np.abs(got["metric_train"] - one.metric_train)) < 1e-6
    assert np.max(np.abs(got["metric_dev"] - one.metric_dev)) < 1e-6
    last = one.all_params[-1]
    assert np.linalg.norm(got["b_stim"] - last["stimulus"]) / np.linalg.norm(last["stimulus"]) < 1e-5
    assert np.linalg.norm(got["y_pred_train"] - np.asarray(one.y_pred["opt"]["train"])) / np.linalg.norm(
        np.asarray(one.y_pred["opt"]["train"])) < 1e-5
    assert int(got["best"]) == one.best_iteration

    class _O(OracleGLM):
        def __init__(self, distr, output_nonlinearity, device=None):
            super().__init__(distr, output_nonlinearity, dtype=np.float64)

    o = build_and_fit(_O, stim, y, 4001, 40, out_nl=out_nl, filt_nl=filt_nl)
    assert rel(got["cost_train"], o.cost_train) < 1e-5
    assert rel(got["cost_dev"], o.cost_dev) < 1e-5
