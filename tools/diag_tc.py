import os, sys
import numpy as np
sys.path.insert(0, '/root/repo')
from rfest_b200.engine import Block, Engine
eng = Engine.get(0)
rng = np.random.default_rng(0)
n, n_pix, n_q = 70001, 300, 36
stim = rng.standard_normal((n, n_pix)).astype(np.float32)
Sxy = (rng.standard_normal((n_pix, n_q)) / np.sqrt(n_pix)).astype(np.float32)
eng.set_option("project_tc", 1)
blk = Block(eng, n, n_q)
blk.project(stim, dst_row0=0, n_out=n, col0=0, src0=0, n_raw_total=n, n_pix=n_pix, first_t=0, nlag=1, shift=0, St=np.ones((1,1),np.float32), Sxy=Sxy)
out = blk.read()
ref = stim.astype(np.float64) @ Sxy.astype(np.float64)
err = np.abs(out - ref).max(1) / np.abs(ref).max()
bad = err > 1e-5
print("bad rows", bad.sum(), "of", n)
tiles = np.arange(n) // 128
bt = np.unique(tiles[bad])
print("bad tiles", len(bt), "first", bt[:20], "last", bt[-10:])
print("tile mod 148 of bad tiles:", np.unique(bt % 148)[:30], " round (tile//148):", np.unique(bt // 148))
r = np.where(bad)[0][:5]
for i in r: print(i, i % 128, out[i, :4], ref[i, :4])
# bad columns?
cb = (np.abs(out - ref) / np.abs(ref).max() > 1e-5).sum(0)
print("bad per column", cb)
