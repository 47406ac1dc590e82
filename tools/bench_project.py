"""Time the projection (K1) on device-resident stimulus: tensor-core vs CUDA-core spatial step.

    python tools/bench_project.py [config3|config5] [log2n]
"""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rfest_b200.engine import Block, Engine
from rfest_b200.splines import KronBasis

which = sys.argv[1] if len(sys.argv) > 1 else "config3"
dims, df = ([25, 20, 15], [8, 6, 6]) if which == "config3" else ([40, 32, 32], [12, 10, 10])
log2n = int(sys.argv[2]) if len(sys.argv) > 2 else (22 if which == "config3" else 20)
n = 1 << log2n
n_pix = dims[1] * dims[2]
eng = Engine.get(0)
kb = KronBasis(dims, df, "cr")
St, Sxy = kb.temporal_f32(), kb.spatial_f32()
stim = torch.randn((n, n_pix), device="cuda", dtype=torch.float32)
burn = dims[0] - 1
n_out = n - burn
blk = Block(eng, n_out, kb.n_coef)
flops_sp = 2.0 * n * n_pix * Sxy.shape[1]
for tc in (0, 1, 1):
    eng.set_option("project_tc", tc)
    eng.synchronize()
    t0 = time.perf_counter()
    blk.project(stim, dst_row0=0, n_out=n_out, col0=0, src0=0, n_raw_total=n, n_pix=n_pix, first_t=burn,
                nlag=dims[0], shift=0, St=St, Sxy=Sxy)
    eng.synchronize()
    dt = time.perf_counter() - t0
    print(f"{which} n=2^{log2n} project_tc={tc}: whole projection {1e3 * dt:.2f} ms "
          f"(spatial step = {flops_sp / 1e9:.1f} GFLOP, input {4.0 * n * n_pix / 1e9:.2f} GB)", flush=True)
