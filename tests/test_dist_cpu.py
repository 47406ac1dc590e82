"""World-size-2 checks on the CPU (gloo): the host-side sharding logic of the multi-GPU path.

Each rank builds ONLY its own rows of the design from its raw rows + halo
(`rfest_b200.shard`), evaluates the oracle's loss and gradient on them, and the partial sums
are all-reduced -- exactly the exchange step the engine performs with NCCL.  The result must
equal the unsharded oracle.
"""

import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import design, splines
from oracle.glm import OracleGLM
from rfest_b200.shard import shard_rows, source_rows


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _data(n_raw=700):
    rng = np.random.default_rng(12)
    stim = rng.standard_normal((n_raw, 4, 3))
    y = rng.poisson(0.4, n_raw).astype(np.float64)
    return stim, y


DIMS, DF, HD, HF = [7, 4, 3], [4, 3, 3], [5], [3]


def _full_model(stim, y):
    m = OracleGLM(distr="poisson", output_nonlinearity="softplus")
    m.add_design_matrix(stim, dims=DIMS, df=DF, smooth="cr", name="stimulus")
    m.add_design_matrix(y, dims=HD, df=HF, smooth="cr", shift=1, name="history")
    m.initialize(method="random", random_seed=5)
    for k, name in enumerate(m.filter_names):
        m.p0[name] = 0.1 * np.random.default_rng(50 + k).standard_normal(m.p0[name].shape)
    m.p0["intercept"]["global"] = -0.3
    m.alpha, m.beta = 1.0, 0.0
    m.y["train"] = y[m.burn_in:]
    return m


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    stim, y = _data()
    n_raw = len(y)
    burn = DIMS[0] - 1
    lo, hi = shard_rows(n_raw - burn, rank, world)

    def shard_design(raw, nlag, shift, dims, df):
        """This rank's rows of design[burn:] built from its raw rows + halo only."""
        s_lo, s_hi = source_rows(burn + lo, hi - lo, nlag, shift, n_raw)
        local = np.zeros_like(raw.reshape(n_raw, -1))
        local[s_lo:s_hi] = raw.reshape(n_raw, -1)[s_lo:s_hi]       # everything else is NOT available
        X = design.lagged_design(local, nlag, shift)[burn:][lo:hi]
        return X @ splines.build_spline_matrix(dims, df, "cr")

    m = OracleGLM(distr="poisson", output_nonlinearity="softplus")
    m.filter_names = ["stimulus", "history"]
    m.filter_nonlinearity = {"stimulus": "none", "history": "none"}
    m.XS["train"] = {"stimulus": shard_design(stim, DIMS[0], 0, DIMS, DF),
                     "history": shard_design(y, HD[0], 1, HD, HF)}
    m.X["train"] = {"stimulus": None, "history": None}
    m.S = {"stimulus": None, "history": None}
    m.y["train"] = y[burn:][lo:hi]
    m.alpha, m.beta, m.burn_in = 1.0, 0.0, burn
    full = _full_model(stim, y)
    loss, g = m.value_and_grad(full.p0, "train")
    vec = np.concatenate([[float(loss)], g["stimulus"].ravel(), g["history"].ravel(),
                          [g["intercept"]["global"], g["intercept"]["stimulus"], hi - lo]])
    t = torch.from_numpy(vec.copy())
    dist.all_reduce(t, op=dist.ReduceOp.SUM)                       # the per-iteration exchange step
    if rank == 0:
        out.put(t.numpy())
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_loss_and_gradient_allreduce_to_the_unsharded_ones(world):
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, out)) for r in range(world)]
    for p in procs:
        p.start()
    got = out.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    stim, y = _data()
    full = _full_model(stim, y)
    loss, g = full.value_and_grad(full.p0, "train")
    want = np.concatenate([[float(loss)], g["stimulus"].ravel(), g["history"].ravel(),
                           [g["intercept"]["global"], g["intercept"]["stimulus"], len(y) - full.burn_in]])
    np.testing.assert_allclose(got, want, rtol=1e-11, atol=1e-11)
