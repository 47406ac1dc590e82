"""Tensor-core spatial projection (tcgen05, 3xTF32) vs the CUDA-core path and fp64 NumPy."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rfest_b200.engine import Block, Engine

eng = Engine.get(0)
rng = np.random.default_rng(0)
for n, n_pix, n_q in [(300, 64, 16), (1000, 300, 36), (5000, 1024, 100), (70001, 300, 36)]:
    stim = rng.standard_normal((n, n_pix)).astype(np.float32)
    Sxy = (rng.standard_normal((n_pix, n_q)) / np.sqrt(n_pix)).astype(np.float32)
    St = np.ones((1, 1), np.float32)
    out = {}
    for tc in (0, 1):
        eng.set_option("project_tc", tc)
        blk = Block(eng, n, n_q)
        t0 = time.perf_counter()
        blk.project(stim, dst_row0=0, n_out=n, col0=0, src0=0, n_raw_total=n, n_pix=n_pix, first_t=0,
                    nlag=1, shift=0, St=St, Sxy=Sxy)
        dt = time.perf_counter() - t0
        out[tc] = blk.read()
        print(f"n={n} K={n_pix} N={n_q} tc={tc}: {1e3 * dt:.2f} ms", flush=True)
    ref = stim.astype(np.float64) @ Sxy.astype(np.float64)
    scale = np.abs(ref).max()
    e0, e1 = np.abs(out[0] - ref).max() / scale, np.abs(out[1] - ref).max() / scale
    print(f"   max err / scale: cuda-core {e0:.2e}  tensor-core {e1:.2e}", flush=True)
    assert e1 < 2e-5, e1
    rel = np.abs(out[1] - ref) / scale
    print(f"   tc error: mean {rel.mean():.2e}, mean signed (out-ref)*sign(ref) {np.mean((out[1] - ref) * np.sign(ref)) / scale:.2e}")
print("ok")
