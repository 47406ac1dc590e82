"""What the 3xTF32 split of the projection GEMM (csrc/project_tc.cuh) costs and buys: time and error of the spatial
contraction with 3 (default), 2 and 1 tcgen05 products per k-step.  Measurement only -- fewer than three products
does not meet the parity gates.      python tools/tc_products.py"""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rfest_b200.engine import Block, Engine

eng = Engine.get(0)
for name, n, n_pix, n_q in [("config 2-4 shape", 1 << 22, 300, 36), ("config 5 shape", 1 << 20, 1024, 100)]:
    g = torch.Generator(device="cuda").manual_seed(1)
    stim = torch.randn((n, n_pix), device="cuda", dtype=torch.float32, generator=g)
    Sxy = (np.random.default_rng(0).standard_normal((n_pix, n_q)) / np.sqrt(n_pix)).astype(np.float32)
    St = np.ones((1, 1), np.float32)
    sub = stim[:4096].cpu().numpy().astype(np.float64)
    ref = sub @ Sxy.astype(np.float64)
    scale = np.abs(ref).max()
    blk = Block(eng, n, n_q)
    for products in (3, 2, 1, 3):
        os.environ["RFB_TC_PRODUCTS"] = str(products)
        best = 1e9
        for rep in range(3):
            eng.synchronize()
            t0 = time.perf_counter()
            blk.project(stim, dst_row0=0, n_out=n, col0=0, src0=0, n_raw_total=n, n_pix=n_pix, first_t=0,
                        nlag=1, shift=0, St=St, Sxy=Sxy)
            eng.synchronize()
            best = min(best, time.perf_counter() - t0)
        out = blk.read()[:4096].astype(np.float64)
        err = np.abs(out - ref).max() / scale
        print(f"{name}: {n} x {n_pix} -> {n_q}, {products} product(s) per k-step: {1e3 * best:.3f} ms "
              f"(spatial + trivial temporal copy), max error / scale {err:.2e}", flush=True)
os.environ.pop("RFB_TC_PRODUCTS", None)
