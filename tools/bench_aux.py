"""Time the one-shot statistics around the fit at the BASELINE config-3 shape (2^log2n x (288 + 8)):
least-squares initialisation (Gram), coefficient covariances (weighted Gram) and response bands.

    python tools/bench_aux.py [log2n=24]
"""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rfest_b200 import GLM
from rfest_b200.engine import Engine
from rfest_b200.synth import GaussianStimulus, RowWindow, TensorRows, poisson_response

log2n = int(sys.argv[1]) if len(sys.argv) > 1 else 24
n_raw = 1 << log2n
dims, df = [25, 20, 15], [8, 6, 6]
eng = Engine.get(0)
stim = GaussianStimulus(n_raw, dims[1:], seed=2046, device=0)
y = poisson_response(eng, stim, dims, 0, n_raw)
m = GLM(distr="poisson", output_nonlinearity="softplus", device=0)
m.add_design_matrix(RowWindow(stim, 0, n_raw), dims=dims, df=df, smooth="cr", name="stimulus")
m.add_design_matrix(TensorRows(y, 0, n_raw), dims=[20], df=[8], smooth="cr", shift=1, name="history")
y_host = y.cpu().numpy()
eng.synchronize()


def timed(fn):
    eng.synchronize()
    t0 = time.perf_counter()
    r = fn()
    eng.synchronize()
    return r, 1e3 * (time.perf_counter() - t0)


n = n_raw - 24
out = {"shape": f"2^{log2n} x 296", "n": n}
_, out["compute_mle_ms"] = timed(lambda: m.initialize(y=y_host, method="mle", compute_ci=False))
_, out["gram_ms"] = timed(lambda: m._dm.gram("train"))
out["gram_dfma_per_s"] = n * (296 * 297 / 2 + 2 * 296) / (out["gram_ms"] * 1e-3)
m.p["opt"] = m.p["mle"]
_, out["filter_variance_ms"] = timed(lambda: m._get_filter_variance("mle"))
_, out["response_band_ms"] = timed(lambda: m._get_response_variance("mle", "train"))
out["band_quadform_fma_per_s"] = n * (288 * 288 + 8 * 8) / (out["response_band_ms"] * 1e-3)
print(json.dumps(out), flush=True)
