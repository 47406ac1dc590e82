"""The other BASELINE.json configurations on one GPU (parity-test shapes; not the bench line).

    python tools/bench_configs.py 1 2 4 5        -> one JSON line per configuration
"""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rfest_b200 import GLM
from rfest_b200.engine import Engine
from rfest_b200.synth import GaussianStimulus, RowWindow, TensorRows, gaussian_response, poisson_response

CONFIGS = {
    1: dict(name="config 1: 1-D LNP dims [25] df [8] + history [20]/[8], 20k samples, 80/20 train/dev",
            dims=[25], df=[8], n_raw=20000, distr="poisson", out="softplus", filt="none", subunits=1,
            history=True, dev=0.2, alpha=1.0, beta=0.01, step=0.1, iters=1000),
    2: dict(name="config 2: 3-D LNP dims [25,20,15] df [8,6,6] + history, 2^22 samples",
            dims=[25, 20, 15], df=[8, 6, 6], n_raw=1 << 22, distr="poisson", out="softplus", filt="none",
            subunits=1, history=True, dev=0.0, alpha=1.0, beta=0.01, step=0.01, iters=60),
    4: dict(name="config 4: LNLN 4 subunits softplus/softplus, elastic net 0.5/0.01, 2^24 samples, 80/20 train/dev",
            dims=[25, 20, 15], df=[8, 6, 6], n_raw=1 << 24, distr="poisson", out="softplus", filt="softplus",
            subunits=4, history=True, dev=0.2, alpha=0.5, beta=0.01, step=0.01, iters=40),
    # not BASELINE configurations: the head counts between LNP (1) and config 4 (5)
    6: dict(name="2 heads: 1 subunit with softplus filter nonlinearity + history, 2^23 samples",
            dims=[25, 20, 15], df=[8, 6, 6], n_raw=1 << 23, distr="poisson", out="softplus", filt="softplus",
            subunits=1, history=True, dev=0.0, alpha=1.0, beta=0.01, step=0.01, iters=40),
    7: dict(name="3 heads: 2 subunits softplus/softplus + history, 2^23 samples",
            dims=[25, 20, 15], df=[8, 6, 6], n_raw=1 << 23, distr="poisson", out="softplus", filt="softplus",
            subunits=2, history=True, dev=0.0, alpha=1.0, beta=0.01, step=0.01, iters=40),
    8: dict(name="4 heads: 3 subunits softplus/softplus + history, 2^23 samples",
            dims=[25, 20, 15], df=[8, 6, 6], n_raw=1 << 23, distr="poisson", out="softplus", filt="softplus",
            subunits=3, history=True, dev=0.0, alpha=1.0, beta=0.01, step=0.01, iters=40),
    5: dict(name="config 5: LG dims [40,32,32] df [12,10,10] (1200 coefficients), 2^24 samples",
            dims=[40, 32, 32], df=[12, 10, 10], n_raw=1 << 24, distr="gaussian", out="none", filt="none",
            subunits=1, history=False, dev=0.0, alpha=1.0, beta=0.001, step=0.01, iters=30),
}


def run(cid, iters=None, quiet=False):
    """Set up configuration `cid` on GPU 0 with device-generated data, time `iters` graph-replayed fit
    iterations with CUDA events and read the sweeps' in-graph %globaltimer durations.  -> dict."""
    c = CONFIGS[cid]
    eng = Engine.get(0)
    dims, n_raw = c["dims"], c["n_raw"]
    burn = dims[0] - 1
    t0 = time.perf_counter()
    stim = GaussianStimulus(n_raw, dims[1:], seed=2046, device=0)
    resp = poisson_response if c["distr"] == "poisson" else gaussian_response
    y = resp(eng, stim, dims, 0, n_raw)
    n_tr = int(round(n_raw * (1 - c["dev"])))
    parts = {"train": (0, n_tr)}
    if c["dev"] > 0:
        parts["dev"] = (n_tr, n_raw)
    m = GLM(distr=c["distr"], output_nonlinearity=c["out"], device=0)
    for kind, (a, b) in parts.items():
        kw = dict(dims=dims, df=c["df"], smooth="cr", filter_nonlinearity=c["filt"]) if kind == "train" else {}
        m.add_design_matrix(RowWindow(stim, a, b - a), name="stimulus", kind=kind, **kw)
        if c["history"]:
            kw = dict(dims=[20], df=[8], smooth="cr", shift=1) if kind == "train" else {}
            m.add_design_matrix(TensorRows(y[a:b], 0, b - a), name="history", kind=kind, **kw)
    m.initialize(num_subunits=c["subunits"], method="random", random_seed=2046, compute_ci=False)
    for i, name in enumerate(m.filter_names):
        m.p0[name] = (0.1 * np.random.default_rng(2046 + i).standard_normal(m.p0[name].shape)).astype(np.float32)
    eng.synchronize()
    setup_s = time.perf_counter() - t0

    m.alpha, m.beta, m.metric = c["alpha"], c["beta"], "corrcoef"
    for kind, (a, b) in parts.items():
        m.y[kind] = y[a + burn:b].cpu().numpy()
    m._dirty = True
    m._sync()
    dm = m._dm
    dm.set_params(*m._flatten(m.p0))
    K, W = iters or c["iters"], 3
    dm.fit_begin(np.full(W + K, c["step"]), c["alpha"], c["beta"], "corrcoef")
    ext = torch.cuda.ExternalStream(eng.stream, device=torch.device("cuda:0"))
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    dm.fit_run(W)
    eng.synchronize()
    ev0.record(ext)
    dm.fit_run(K)
    ev1.record(ext)
    eng.synchronize()
    ms = ev0.elapsed_time(ev1) / K
    persistent = dm.fit_is_persistent
    if persistent:       # one cooperative kernel runs the whole loop: there are no per-iteration sweeps to stamp
        sweep_tr = sweep_dv = None
    else:
        sweep_tr = float(np.mean(dm.fit_sweep_times("train", W, K)))
        sweep_dv = float(np.mean(dm.fit_sweep_times("dev", W, K))) if c["dev"] > 0 else 0.0
    dm.fit_finish(W + K - 1)
    tr = dm.fit_traces(W + K)
    dm.fit_end()
    n_train = n_tr - burn
    n_dev = (n_raw - n_tr - burn) if c["dev"] > 0 else 0
    C = int(np.prod(c["df"])) + (8 if c["history"] else 0)
    b_iter = 4.0 * (n_train + n_dev) * (C + 1)
    out = {"config": c["name"], "iterations_per_s": 1e3 / ms, "ms_per_iter": ms, "iters_timed": K,
           "sweep_ms_in_graph": {"train": sweep_tr, "dev": sweep_dv},
           "algorithmic_GB_per_iter": b_iter / 1e9,
           "achieved_GBps_whole_iteration": b_iter / ms / 1e6,
           "persistent_kernel": persistent,
           "achieved_GBps_train_sweep": None if persistent else 4.0 * n_train * (C + 1) / sweep_tr / 1e6,
           "frac_of_measured_peak": b_iter / ms / 1e6 / PEAK_GBS,
           "train_sweep_frac_of_measured_peak": None if persistent else 4.0 * n_train * (C + 1) / sweep_tr / 1e6 / PEAK_GBS,
           "setup_s": setup_s,
           "loss_first_last": [float(tr[0][0]), float(tr[0][-1])]}
    if not quiet:
        print(json.dumps(out), flush=True)
    del m, dm, stim, y
    import gc

    gc.collect()
    torch.cuda.empty_cache()
    return out


def projection(dims, df, log2n, reps=3):
    """K1 alone: `add_design_matrix` of a DEVICE-resident stimulus of 2^log2n frames (spatial tcgen05
    GEMM + temporal step), CUDA-event timed.  Bytes = frames read + projected rows written."""
    eng = Engine.get(0)
    n = 1 << log2n
    n_pix = int(np.prod(dims[1:]))
    stim = torch.randn((n, n_pix), device="cuda:0", dtype=torch.float32)
    ext = torch.cuda.ExternalStream(eng.stream, device=torch.device("cuda:0"))
    best = None
    for _ in range(reps):
        m = GLM(distr="gaussian", device=0)
        torch.cuda.synchronize()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record(ext)
        m.add_design_matrix(stim, dims=dims, df=df, smooth="cr", name="stimulus")
        ev1.record(ext)
        eng.synchronize()
        t = ev0.elapsed_time(ev1)
        best = t if best is None else min(best, t)
        del m
    # the kernels alone: project into an existing block (no allocation, no basis construction)
    from rfest_b200.engine import Block
    from rfest_b200.splines import KronBasis

    basis = KronBasis(dims, df, "cr")
    n_out = n - dims[0] + 1
    blk = Block(eng, n_out, basis.shape[1])
    kern = None
    for _ in range(reps):
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        eng.synchronize()
        ev0.record(ext)
        blk.project(stim, dst_row0=0, n_out=n_out, col0=0, src0=0, n_raw_total=n, n_pix=n_pix,
                    first_t=dims[0] - 1, nlag=dims[0], shift=0, St=basis.temporal_f32(), Sxy=basis.spatial_f32())
        ev1.record(ext)
        eng.synchronize()
        t = ev0.elapsed_time(ev1)
        kern = t if kern is None else min(kern, t)
    del blk
    cols = int(np.prod(df))
    byts = 4.0 * n * n_pix + 4.0 * (n - dims[0] + 1) * cols
    flops = 2.0 * n * n_pix * int(np.prod(df[1:])) + 2.0 * n * cols * dims[0]
    del stim
    torch.cuda.empty_cache()
    return {"dims": dims, "df": df, "frames": n, "add_design_matrix_ms": best, "kernels_ms": kern,
            "GBps_in_plus_out": byts / kern / 1e6, "frac_of_measured_peak": byts / kern / 1e6 / PEAK_GBS,
            "separable_TFLOPs": flops / kern / 1e9,
            "note": "add_design_matrix_ms includes allocating the block (cudaMalloc) and building the basis; "
                    "kernels_ms is rfb_block_project into an existing block (basis upload, tcgen05 spatial GEMM, "
                    "temporal step), CUDA events"}


def host_projection(dims, df, log2n):
    """`add_design_matrix(<host ndarray>)`: the frames cross PCIe through the engine's pinned 3-slot ring
    (copy stream overlapped with the projection kernels).  Wall clock, pageable source."""
    eng = Engine.get(0)
    n = 1 << log2n
    frame = tuple(dims[1:])
    host = np.empty((n,) + frame, dtype=np.float32)
    step = 1 << 20
    for k0 in range(0, n, step):                                # generated on the device, copied out
        k1 = min(n, k0 + step)
        host[k0:k1] = torch.randn((k1 - k0,) + frame, device="cuda:0").cpu().numpy()
    out = {}
    for label, overlap, threads in (("ring", 1, 4), ("ring_8_threads", 1, 8), ("ring_2_threads", 1, 2),
                                    ("serial_r1", 0, 4)):
        eng.set_option("h2d_overlap", overlap)
        eng.set_option("h2d_threads", threads)
        best = None
        for _ in range(2):
            m = GLM(distr="gaussian", device=0)
            eng.synchronize()
            t0 = time.perf_counter()
            m.add_design_matrix(host, dims=dims, df=df, smooth="cr", name="stimulus")
            eng.synchronize()
            t = time.perf_counter() - t0
            best = t if best is None else min(best, t)
            del m
        out[label] = {"wall_s": best, "h2d_GBps": host.nbytes / best / 1e9}
    eng.set_option("h2d_overlap", 1)
    eng.set_option("h2d_threads", 4)
    out.update({"frames": n, "host_GB": host.nbytes / 1e9, "source": "pageable numpy array"})
    return out


try:
    PEAK_GBS = float(json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                                                 "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    PEAK_GBS = 6551.0


if __name__ == "__main__":
    for cid in [int(a) for a in sys.argv[1:]] or [1, 2, 4, 5]:
        run(cid)
