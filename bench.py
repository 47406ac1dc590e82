#!/usr/bin/env python
"""Benchmark of the spline-GLM fit hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--log2n 25]

metric   spline-GLM fit iterations / s at the 2^25 x 288 design (BASELINE.json config 3:
         3-D spline LNP, dims [25,20,15], df [8,6,6] + history [20]/[8] shift 1,
         Poisson / softplus, beta 0.01, alpha 1), time-sharded over N GPUs (strong scaling).
step     one fit iteration = one streaming sweep over XS (forward, loss, gradient, metric
         sums), fixed-order reduction, all-reduce over ranks, elastic net + Adam.
value    K iterations timed on the device (CUDA events on the engine's stream, barrier +
         synchronize on both sides, max over ranks), XS resident in HBM.
e2e      the same K iterations through the public API `GLM.fit(y=<pinned host array>, ...)`:
         host->device copy of y, graph capture, K iterations, the final forward pass and the
         device->host read of traces, parameters and predictions are all inside the timed
         region (the design XS stays resident, as it does in the reference after
         add_design_matrix).
roofline train-sweep kernel only: algorithmic bytes 4 * n_local * (296 + 1) per launch over the
         kernel's mean duration IN THE TIMED REGION: every CTA of the sweep folds %globaltimer into a
         per-iteration (first entry, last exit) pair inside the replayed graph (`rfb_fit_sweep_times`).
         `kernel_ms_eager` repeats the iterations eagerly with CUDA events around each kernel as a
         cross-check.
extras   (N=1 only, after the headline; `--no-extras` skips them) `other_configs`: BASELINE configs 2, 4, 5
         and 1 timed the same way; `projection`: the one-time K1 (tcgen05 spatial GEMM + temporal step) on
         device-resident frames; `setup_host`: `add_design_matrix(<host ndarray>)` wall clock through the
         pinned ring against the serial round-1 path.
"""

import argparse
import collections
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

DIMS, DF = [25, 20, 15], [8, 6, 6]
H_DIMS, H_DF = [20], [8]
N_COLS = 288 + 8
STEP_SIZE, BETA, ALPHA = 0.01, 0.01, 1.0


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--log2n", type=int, default=25, help="log2 of the number of raw samples")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-log2n", type=int, default=21)
    ap.add_argument("--no-extras", action="store_true", help="headline only (no other configs / projection)")
    ap.add_argument("--host-log2n", type=int, default=22, help="frames of the host-array projection timing")
    return ap.parse_args()


# --------------------------------------------------------------------------------------------
# CPU arm: the oracle's fp32 restatement of the reference loop on the host cores
# --------------------------------------------------------------------------------------------
def cpu_arm(log2n_workload, log2n_sample, steps, warmup):
    """Times the reference's per-iteration work (value_and_grad + Adam, second train forward
    for the metric; rfest/GLM/_glm.py:720-734) with the oracle's NumPy fp32 restatement, XS
    injected (the reference cannot build the 3-D lagged design at this size: n x 7500 floats).
    Returns iterations/s extrapolated linearly to the workload's n."""
    import torch

    from oracle.glm import OracleGLM

    n = 1 << log2n_sample
    rng = np.random.default_rng(2046)
    m = OracleGLM(distr="poisson", output_nonlinearity="softplus", dtype=np.float32)
    m.filter_names = ["stimulus", "history"]
    m.filter_nonlinearity = {"stimulus": "none", "history": "none"}
    m.X["train"] = {"stimulus": None, "history": None}
    # random rows are drawn for at most 2^21 samples and repeated beyond that (the arithmetic and the
    # memory traffic per iteration do not depend on the values; drawing 10^10 normals would take minutes)
    nb = min(n, 1 << 21)
    blk_s = rng.standard_normal((nb, 288), dtype=np.float32) * 0.06
    blk_h = rng.standard_normal((nb, 8), dtype=np.float32) * 0.1
    blk_y = rng.poisson(0.3, nb).astype(np.float32)
    if nb == n:
        xs_s, xs_h, yy = blk_s, blk_h, blk_y
    else:
        xs_s, xs_h, yy = np.empty((n, 288), np.float32), np.empty((n, 8), np.float32), np.empty(n, np.float32)
        for k0 in range(0, n, nb):
            xs_s[k0:k0 + nb], xs_h[k0:k0 + nb], yy[k0:k0 + nb] = blk_s, blk_h, blk_y
    m.XS["train"] = {"stimulus": xs_s, "history": xs_h}
    m.S = {"stimulus": None, "history": None}
    m.y["train"] = yy
    m.alpha, m.beta = ALPHA, BETA
    m.burn_in = 0
    p0 = {"stimulus": (0.1 * np.random.default_rng(2046).standard_normal((288, 1))).astype(np.float32),
          "history": (0.1 * np.random.default_rng(2047).standard_normal((8, 1))).astype(np.float32),
          "intercept": {"stimulus": 0.0, "history": 0.0, "global": 0.0}}
    m.y_pred["opt"] = {}
    # all host cores for the BLAS calls, whatever OMP_NUM_THREADS says (torchrun sets it to 1 per rank)
    cores = os.cpu_count()
    blas_threads = None
    try:
        from threadpoolctl import threadpool_info, threadpool_limits

        limiter = threadpool_limits(limits=cores)
        blas_threads = max([int(i.get("num_threads", 1)) for i in threadpool_info()] or [1])
    except Exception:
        limiter = None
    m.optimize(p0, warmup, "corrcoef", STEP_SIZE, 0, 0, "last")
    t0 = time.perf_counter()
    m.optimize(p0, steps, "corrcoef", STEP_SIZE, 0, 0, "last")
    dt = time.perf_counter() - t0
    if limiter is not None:
        limiter.restore_original_limits()
    its = steps / dt
    scaled = its * n / float(1 << log2n_workload)
    full = log2n_sample >= log2n_workload
    return {"value": scaled, "unit": "iterations/s", "cores": cores, "kind": "port",
            "threads": blas_threads if blas_threads is not None else torch.get_num_threads(),
            "extrapolated": not full, "sample_ms_per_step": 1e3 * dt / steps, "sample_log2n": log2n_sample,
            "sample": f"oracle fp32 NumPy restatement (pinned to the reference's own _glm.py, see oracle/), XS "
                      f"injected, n=2^{log2n_sample} rows x 296 cols, {steps} iterations in {dt:.2f} s = "
                      f"{its:.3f} it/s" + ("; the full workload, nothing extrapolated" if full else
                                           f"; LINEARLY EXTRAPOLATED to n=2^{log2n_workload}")}


def reference_sample_log2n(log2n_workload, requested):
    """Largest sample the reference arm can hold: the whole workload when the host has the memory for XS
    (4 * n * 296 bytes) plus the loop's n-vectors, else the requested bounded sample."""
    try:
        import psutil

        avail = psutil.virtual_memory().available
    except Exception:
        return requested
    l2 = log2n_workload
    while l2 > requested and (1 << l2) * (4.0 * 296 + 8.0 * 40) * 1.25 > avail:
        l2 -= 1
    return max(l2, requested)


# --------------------------------------------------------------------------------------------
# clocks
# --------------------------------------------------------------------------------------------
class ClockSampler:
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.lines, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.FIELDS}",
                 "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None
        return self

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            parts = [p.strip() for p in ln.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0]))
                smax.append(float(parts[1]))
            except ValueError:
                continue
            for name, val in zip(names, parts[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None,
                "sm_max_mhz": float(max(smax)) if smax else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# --------------------------------------------------------------------------------------------
# GPU arm
# --------------------------------------------------------------------------------------------
def gpu_arm(args):
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local_rank}"))

    from rfest_b200 import GLM
    from rfest_b200.engine import Engine
    from rfest_b200.shard import shard_rows, source_rows
    from rfest_b200.synth import GaussianStimulus, TensorRows, poisson_response

    def barrier():
        if world > 1:
            dist.barrier()

    n_raw = 1 << args.log2n
    burn = DIMS[0] - 1
    eng = Engine.get(local_rank).init_distributed()
    t_setup = time.perf_counter()

    # ---- synthetic recording, generated on the device per shard -------------------------------
    stim = GaussianStimulus(n_raw, DIMS[1:], seed=2046, device=local_rank)
    lo, hi = shard_rows(n_raw - burn, rank, world)
    # spikes: this rank's response rows plus what its history design reads
    h_lo, h_hi = source_rows(burn + lo, hi - lo, H_DIMS[0], 1, n_raw)
    y_lo, y_hi = min(h_lo, burn + lo), max(h_hi, burn + hi)
    y_dev = poisson_response(eng, stim, DIMS, y_lo, y_hi)
    y_rows = TensorRows(y_dev, y_lo, n_raw)

    model = GLM(distr="poisson", output_nonlinearity="softplus", device=local_rank)
    model.add_design_matrix(stim, dims=DIMS, df=DF, smooth="cr", name="stimulus")
    model.add_design_matrix(y_rows, dims=H_DIMS, df=H_DF, smooth="cr", shift=1, name="history")
    model.initialize(num_subunits=1, method="random", random_seed=2046, compute_ci=False)
    for i, name in enumerate(model.filter_names):
        model.p0[name] = (0.1 * np.random.default_rng(2046 + i).standard_normal(
            model.p0[name].shape)).astype(np.float32)
    y_local = y_dev[burn + lo - y_lo: burn + hi - y_lo].contiguous()
    eng.synchronize()
    torch.cuda.synchronize()
    setup_s = time.perf_counter() - t_setup
    n_local = hi - lo

    # ---- device-timed fit iterations (value) ------------------------------------------------------
    K, W = args.steps, max(args.warmup, 3)
    model.alpha, model.beta, model.metric = ALPHA, BETA, "corrcoef"
    model.y["train"] = y_local.cpu().numpy()
    model._dirty = True
    model._sync()
    dm = model._dm
    dm.set_params(*model._flatten(model.p0))
    total = W + K + K
    dm.fit_begin(np.full(total, STEP_SIZE), ALPHA, BETA, "corrcoef")
    ext = torch.cuda.ExternalStream(eng.stream, device=torch.device(f"cuda:{local_rank}"))
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sampler = ClockSampler(local_rank).start() if rank == 0 else None   # samples warm-up + timed region
    dm.fit_run(W)
    eng.synchronize()
    launches0 = eng.launch_count
    barrier()
    eng.synchronize()
    ev0.record(ext)
    dm.fit_run(K)
    ev1.record(ext)
    eng.synchronize()
    barrier()
    ms = ev0.elapsed_time(ev1)
    gpu_launches = eng.launch_count - launches0
    sweep_in_graph = dm.fit_sweep_times("train", W, K)       # the timed iterations themselves
    # per-kernel timing of the same iterations launched eagerly with events around each kernel
    ms4 = dm.fit_run_timed(K)
    clocks = sampler.stop() if sampler else None
    dm.fit_finish(W + 2 * K - 1)
    traces = dm.fit_traces(W + 2 * K)
    dm.fit_end()
    sweep_ms = float(np.mean(sweep_in_graph))
    if world > 1:
        t = torch.tensor([ms, sweep_ms], device=f"cuda:{local_rank}", dtype=torch.float64)
        all_sweeps = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(all_sweeps, t)
        rank_sweep_ms = [float(x[1]) for x in all_sweeps]
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, sweep_ms = float(t[0]), float(t[1])
    else:
        rank_sweep_ms = [sweep_ms]

    # ---- end to end through the public API, host buffers ---------------------------------------------
    y_host = torch.empty(n_local + burn, dtype=torch.float32).pin_memory()
    y_host[burn:] = y_local.cpu()
    y_host_np = y_host.numpy()
    if world > 1:
        y_arg = y_host_np[burn:]          # this rank's trimmed rows
    else:
        y_arg = y_host_np
    model.fit(y=y_arg, num_iters=W, step_size=STEP_SIZE, beta=BETA, alpha=ALPHA, verbose=0, tolerance=0)
    eng.synchronize()
    barrier()
    t0 = time.perf_counter()
    model.fit(y=y_arg, num_iters=K, step_size=STEP_SIZE, beta=BETA, alpha=ALPHA, verbose=0, tolerance=0)
    eng.synchronize()
    e2e_s = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([e2e_s], device=f"cuda:{local_rank}", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t[0])
    # variant that also brings the final predictions to the host (pinned buffer)
    pred_host = torch.empty(n_local, dtype=torch.float32).pin_memory()
    barrier()
    t0 = time.perf_counter()
    model.fit(y=y_arg, num_iters=K, step_size=STEP_SIZE, beta=BETA, alpha=ALPHA, verbose=0, tolerance=0)
    pred_host.copy_(model.y_pred["opt"]["train"].as_torch(), non_blocking=False)
    eng.synchronize()
    e2e_pred_s = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([e2e_pred_s], device=f"cuda:{local_rank}", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_pred_s = float(t[0])
    # where that wall clock goes: one more fit with a synchronize after every device-model call (run after the
    # timed ones; the phases therefore add up to a little more than e2e_s)
    phases = collections.OrderedDict()
    dm_e2e = model._dm
    names = ["set_y", "set_params", "fit_begin", "fit_run", "fit_finish", "fit_traces", "fit_params_all",
             "fit_predictions_device", "fit_end"]
    saved = {nm: getattr(dm_e2e, nm) for nm in names if hasattr(dm_e2e, nm)}

    def _wrap(fn, nm):
        def f(*a, **k):
            eng.synchronize()
            t = time.perf_counter()
            r = fn(*a, **k)
            eng.synchronize()
            phases[nm] = phases.get(nm, 0.0) + 1e3 * (time.perf_counter() - t)
            return r
        return f

    for nm, fn in saved.items():
        setattr(dm_e2e, nm, _wrap(fn, nm))
    barrier()
    t0 = time.perf_counter()
    model.fit(y=y_arg, num_iters=K, step_size=STEP_SIZE, beta=BETA, alpha=ALPHA, verbose=0, tolerance=0)
    eng.synchronize()
    phases["total_instrumented"] = 1e3 * (time.perf_counter() - t0)
    phases["host_code_outside"] = phases["total_instrumented"] - sum(v for k, v in phases.items()
                                                                     if k != "total_instrumented")
    for nm in saved:
        delattr(dm_e2e, nm)                 # back to the class's methods
    P = 296
    h2d = (4 * n_local + 4 * P + 8 * 3 + 8 * K) / K
    d2h = (4 * 8 * K + K * (4 * P + 8 * 3)) / K      # traces + every iteration's parameters

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return None

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    alg_bytes = 4.0 * n_local * (N_COLS + 1)
    achieved = alg_bytes / (sweep_ms * 1e-3) / 1e9
    traffic = None
    try:
        traffic = json.load(open(os.path.join(ROOT, "profiles", "sweep_traffic.json"))).get(
            f"log2n_{args.log2n}_gpus_{world}")
    except Exception:
        pass
    out = {
        "metric": "spline-GLM fit iterations/s at the 2^%d x 288 design" % args.log2n,
        "value": K / (ms * 1e-3),
        "unit": "iterations/s",
        "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": ms / K,
        "higher_is_better": True,
        "scaling": "strong",
        "vs_baseline": None,
        "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": "BASELINE config 3: 3-D spline LNP dims [25,20,15] df [8,6,6] + history "
                               "[20]/[8] shift 1, poisson/softplus, n_raw=2^%d, time-sharded over %d GPU(s)"
                               % (args.log2n, world),
                   "n_raw": n_raw, "design_cols": N_COLS, "rows_per_gpu": n_local,
                   "l2": "inputs (%.1f GB per GPU) are larger than L2; no flush needed"
                         % (alg_bytes / 1e9),
                   "setup_s": setup_s},
        "e2e": {"value": K / e2e_s, "unit": "iterations/s", "h2d_bytes_per_step": h2d,
                "d2h_bytes_per_step": d2h,
                "what": "GLM.fit(y=pinned host array, num_iters=K) wall clock: H2D of y, graph capture, K "
                        "iterations, trailing forward pass, D2H of the loss/metric traces and all parameters; "
                        "predictions stay on the device (as the reference's JAX arrays do) until read. "
                        "K iterations cost K + 1 sweeps (the reference's forward pass at the final parameters "
                        "for metric_train[K-1] / y_pred), so e2e <= K / (K + 1) of `value`",
                "phases_ms": {k: round(v, 3) for k, v in phases.items()},
                "with_predictions": {"value": K / e2e_pred_s, "unit": "iterations/s",
                                     "d2h_bytes_per_step": d2h + 4.0 * n_local / K,
                                     "what": "the same call followed by the D2H copy of y_pred['opt']['train'] "
                                             "into pinned host memory"}},
        "gpu_launches": int(gpu_launches),
        "kernel_ms": {"train_sweep": sweep_ms, "rest_of_step": ms / K - sweep_ms,
                      "how": "train_sweep: in-graph %globaltimer stamps of the K timed iterations (max over "
                             "ranks of the per-rank mean); rest_of_step: ms_per_step minus that"},
        "kernel_ms_eager": {"train_sweep": float(ms4[0]), "dev_sweep": float(ms4[1]),
                            "reduce_allreduce": float(ms4[2]), "update": float(ms4[3]),
                            "how": "the same iterations relaunched eagerly with CUDA events around each kernel"},
        "rank_sweep_ms": rank_sweep_ms,
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": achieved / peak, "traffic": traffic,
                     "peak_source": "MEASURED_PEAKS.json hbm_gbs (of measured)" if peaks else "fallback 6650 (of fallback)",
                     "frac_of_nominal_8000": achieved / 8000.0,
                     "kernel": "rfb::sweep_staged_kernel<1,3,8,true,1>", "bytes_per_launch": alg_bytes,
                     "timed": "inside the timed region (in-graph stamps)"},
        "clocks": clocks,
        "loss_first_last": [float(traces[0][0]), float(traces[0][-1])],
    }
    if world == 1 and not args.no_cpu_baseline:
        out["cpu_baseline"] = cpu_arm(args.log2n, args.cpu_log2n, 60, 3)
    if world == 1 and not args.no_extras:
        # free the headline's 40 GB before the other configurations allocate theirs
        del model, dm, stim, y_dev, y_rows, y_local
        import gc

        gc.collect()
        torch.cuda.empty_cache()
        sys.path.insert(0, os.path.join(ROOT, "tools"))
        import bench_configs as bc

        extras = {}
        for cid, iters in ((2, 40), (4, 40), (5, 20), (1, 1000)):
            try:
                extras[f"config{cid}"] = bc.run(cid, iters=iters, quiet=True)
            except Exception as exc:                      # an extra must not cost the headline
                extras[f"config{cid}"] = {"error": repr(exc)}
        out["other_configs"] = extras
        proj = {}
        for label, dims, df, l2n in (("config3_shape", DIMS, DF, 22), ("config5_shape", [40, 32, 32], [12, 10, 10], 20)):
            try:
                proj[label] = bc.projection(dims, df, l2n)
            except Exception as exc:
                proj[label] = {"error": repr(exc)}
        out["projection"] = proj
        try:
            out["setup_host"] = bc.host_projection(DIMS, DF, args.host_log2n)
        except Exception as exc:
            out["setup_host"] = {"error": repr(exc)}
    if world > 1:
        dist.destroy_process_group()
    return out


def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    if args.impl == "reference":
        if rank != 0:
            return
        steps = max(args.steps, 1)
        sample = reference_sample_log2n(args.log2n, args.cpu_log2n) if os.environ.get(
            "RFB_REFERENCE_FULL", "1") != "0" else args.cpu_log2n
        big = sample > 22        # >= 0.3 s per iteration: keep the whole run within a few minutes
        base = cpu_arm(args.log2n, sample, min(steps, 20 if big else 60), min(max(args.warmup, 1), 2 if big else 5))
        out = {
            "impl": "reference",
            "metric": "spline-GLM fit iterations/s at the 2^%d x 288 design%s" % (
                args.log2n, " (CPU port, extrapolated from a 2^%d-row sample)" % sample if base["extrapolated"] else
                " (CPU port of the reference loop, measured at the full size)"),
            "value": base["value"], "unit": "iterations/s",
            "n_gpus": int(os.environ.get("WORLD_SIZE", "1")), "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 / base["value"], "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "BASELINE config 3 (see the GPU arm); CPU arm: JAX is not installable in "
                                   "this image, so this is the oracle's fp32 restatement of the reference loop"},
            "cpu_baseline": base,
            "e2e": {"value": base["value"], "unit": "iterations/s", "h2d_bytes_per_step": 0,
                    "d2h_bytes_per_step": 0},
            "gpu_launches": 0,
        }
        print(json.dumps(out))
        return
    out = gpu_arm(args)
    if out is not None:
        print(json.dumps(out))


if __name__ == "__main__":
    main()
