"""Race hunt for the tcgen05 projection kernel: many launches over shapes with several tiles per CTA,
each checked against float64.

    python tools/stress_tc.py [repeats=20]
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rfest_b200.engine import Block, Engine

eng = Engine.get(0)
eng.set_option("project_tc", 1)
repeats = int(sys.argv[1]) if len(sys.argv) > 1 else 20
worst = 0.0
for n, n_pix, n_q in [(70001, 300, 36), (131072 + 77, 1024, 100), (40000, 64, 8), (300000, 300, 36)]:
    rng = np.random.default_rng(n)
    stim = rng.standard_normal((n, n_pix)).astype(np.float32)
    Sxy = (rng.standard_normal((n_pix, n_q)) / np.sqrt(n_pix)).astype(np.float32)
    ref = stim.astype(np.float64) @ Sxy.astype(np.float64)
    scale = np.abs(ref).max()
    blk = Block(eng, n, n_q)
    for rep in range(repeats):
        blk.project(stim, dst_row0=0, n_out=n, col0=0, src0=0, n_raw_total=n, n_pix=n_pix, first_t=0, nlag=1,
                    shift=0, St=np.ones((1, 1), np.float32), Sxy=Sxy)
        err = np.abs(blk.read() - ref).max() / scale
        worst = max(worst, err)
        if err > 5e-6:
            print(f"FAIL n={n} K={n_pix} N={n_q} rep={rep}: err {err:.3g}", flush=True)
            sys.exit(1)
    print(f"ok n={n} K={n_pix} N={n_q} x{repeats}", flush=True)
print(f"worst error {worst:.3g} of the scale")
