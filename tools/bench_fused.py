"""Time K1 on a device-resident slab: fused single kernel vs the two-kernel path (CUDA events around the call).

    python tools/bench_fused.py [log2n=22]
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rfest_b200.engine import Block, Engine
from rfest_b200.splines import KronBasis

dims, df = [25, 20, 15], [8, 6, 6]
log2n = int(sys.argv[1]) if len(sys.argv) > 1 else 22
n = 1 << log2n
n_pix = dims[1] * dims[2]
eng = Engine.get(0)
kb = KronBasis(dims, df, "cr")
St, Sxy = kb.temporal_f32(), kb.spatial_f32()
stim = torch.randn((n, n_pix), device="cuda", dtype=torch.float32)
burn = dims[0] - 1
n_out = n - burn
blk = Block(eng, n_out, kb.n_coef)
eng.set_option("project_slab_rows", n)
for fused in (0, 1, 0, 1, 1):
    eng.set_option("project_fused", fused)
    eng.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    import time
    t0 = time.perf_counter()
    blk.project(stim, dst_row0=0, n_out=n_out, col0=0, src0=0, n_raw_total=n, n_pix=n_pix, first_t=burn,
                nlag=dims[0], shift=0, St=St, Sxy=Sxy)
    eng.synchronize()
    dt = time.perf_counter() - t0
    gb = 4.0 * n * (n_pix + kb.n_coef) / 1e9
    print(f"2^{log2n} frames project_fused={fused}: {1e3 * dt:.3f} ms wall ({gb / dt:.0f} GB/s in + out)", flush=True)
