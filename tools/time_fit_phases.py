"""Where the wall clock of GLM.fit(y=<pinned host array>) goes: every device-model call timed with a
synchronize after it (so the phases add up to more than the un-instrumented wall time printed first).
    python tools/time_fit_phases.py [log2n] [K]"""
import collections
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rfest_b200 import GLM
from rfest_b200.engine import Engine
from rfest_b200.synth import GaussianStimulus, TensorRows, poisson_response

log2n = int(sys.argv[1]) if len(sys.argv) > 1 else 25
K = int(sys.argv[2]) if len(sys.argv) > 2 else 20
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
yh = torch.empty(n_raw, dtype=torch.float32).pin_memory()
yh.copy_(y)
yh = yh.numpy()
kw = dict(step_size=0.01, beta=0.01, alpha=1.0, verbose=0, tolerance=0)
m.fit(y=yh, num_iters=3, **kw)
eng.synchronize()
for rep in range(3):
    t0 = time.perf_counter()
    m.fit(y=yh, num_iters=K, **kw)
    eng.synchronize()
    print(f"fit({K}) wall {1e3 * (time.perf_counter() - t0):.1f} ms")

acc = collections.OrderedDict()
dm = m._dm
for name in ["clear_filters", "add_filter", "set_data", "set_y", "set_params", "fit_begin", "fit_run", "fit_finish",
             "fit_traces", "fit_params_all", "fit_predictions_device", "fit_end"]:
    fn = getattr(dm, name)

    def wrap(fn=fn, name=name):
        def f(*a, **k):
            eng.synchronize()
            t = time.perf_counter()
            r = fn(*a, **k)
            eng.synchronize()
            acc[name] = acc.get(name, 0.0) + time.perf_counter() - t
            return r
        return f
    setattr(dm, name, wrap())
t0 = time.perf_counter()
m.fit(y=yh, num_iters=K, **kw)
eng.synchronize()
tot = time.perf_counter() - t0
print(f"instrumented fit({K}) wall {1e3 * tot:.1f} ms")
for k, v in acc.items():
    print(f"  {k:26s} {1e3 * v:8.2f} ms")
print(f"  {'host code outside':26s} {1e3 * (tot - sum(acc.values())):8.2f} ms")
