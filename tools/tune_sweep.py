"""Time the train sweep for several launch shapes on one GPU (tuning aid, not a benchmark).

    python tools/tune_sweep.py [log2n]
"""
import itertools
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from rfest_b200 import GLM
from rfest_b200.engine import Engine
from rfest_b200.synth import GaussianStimulus, TensorRows, poisson_response

log2n = int(sys.argv[1]) if len(sys.argv) > 1 else 24
n_raw = 1 << log2n
DIMS, DF = [25, 20, 15], [8, 6, 6]
eng = Engine.get(0)
stim = GaussianStimulus(n_raw, DIMS[1:], device=0)
y = poisson_response(eng, stim, DIMS, 0, n_raw)
m = GLM(distr="poisson", output_nonlinearity="softplus", device=0)
m.add_design_matrix(stim, dims=DIMS, df=DF, smooth="cr", name="stimulus")
m.add_design_matrix(TensorRows(y, 0, n_raw), dims=[20], df=[8], smooth="cr", shift=1, name="history")
m.initialize(method="random", random_seed=2046, compute_ci=False)
for i, name in enumerate(m.filter_names):
    m.p0[name] = (0.1 * np.random.default_rng(2046 + i).standard_normal(m.p0[name].shape)).astype(np.float32)
m.alpha, m.beta = 1.0, 0.01
m.y["train"] = y[24:].cpu().numpy()
n = n_raw - 24
bytes_ = 4.0 * n * 297
configs = [(1, st, c, w) for st, c, w in
           [(2, 1, 4), (3, 1, 4), (4, 1, 4), (2, 1, 3), (3, 1, 3), (4, 1, 3), (2, 1, 2), (4, 1, 2), (6, 1, 2),
            (2, 2, 2), (3, 2, 2), (4, 2, 2), (2, 2, 3), (3, 2, 3), (2, 2, 4), (2, 3, 2), (3, 3, 2)]]
ref = None
for staged, stages, ctas, warps in configs:
    eng.set_option("sweep_staged", staged)
    eng.set_option("sweep_stages", stages)
    eng.set_option("sweep_ctas", ctas)
    eng.set_option("sweep_warps", warps)
    m._dirty = True
    m._sync()
    dm = m._dm
    dm.set_params(*m._flatten(m.p0))
    dm.fit_begin(np.full(16, 0.01), 1.0, 0.01, "corrcoef")
    dm.fit_run(3)
    ms = dm.fit_run_timed(8)
    dm.fit_finish(10)
    tr = dm.fit_traces(11)[0]
    dm.fit_end()
    if ref is None:
        ref = tr
    dev = float(np.max(np.abs(tr - ref) / np.abs(ref)))
    print(f"staged={staged} stages={stages} ctas/SM={ctas or 'max'} warps/CTA={warps}: sweep {ms[0]:.3f} ms  "
          f"{bytes_ / ms[0] / 1e6:.0f} GB/s  ({bytes_ / ms[0] / 1e6 / 6551:.3f} of measured)  "
          f"loss dev vs first {dev:.1e}", flush=True)
