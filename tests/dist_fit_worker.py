"""Worker of tests/test_gpu_multi.py: one rank of a time-sharded fit (launched with torchrun)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def build_and_fit(GLM, stim, y, n_tr, iters, device=None, out_nl=None, filt_nl=None):
    out_nl = out_nl or os.environ.get("RFB_TEST_OUT_NL", "softplus")
    filt_nl = filt_nl or os.environ.get("RFB_TEST_FILT_NL", "none")
    m = GLM(distr="poisson", output_nonlinearity=out_nl, device=device)
    m.add_design_matrix(stim[:n_tr], dims=[8, 5, 4], df=[4, 3, 3], smooth="cr", filter_nonlinearity=filt_nl,
                        name="stimulus")
    m.add_design_matrix(y[:n_tr], dims=[6], df=[4], smooth="cr", shift=1, name="history")
    m.add_design_matrix(stim[n_tr:], name="stimulus", kind="dev")
    m.add_design_matrix(y[n_tr:], name="history", kind="dev")
    m.initialize(method="random", random_seed=3, compute_ci=False)
    for k, name in enumerate(m.filter_names):
        m.p0[name] = (0.1 * np.random.default_rng(70 + k).standard_normal(m.p0[name].shape)).astype(np.float32)
    m.fit(y={"train": y[:n_tr], "dev": y[n_tr:]}, num_iters=iters, step_size=0.05, beta=0.01,
          verbose=0, tolerance=0, metric="corrcoef")
    return m


def main():
    out_path = sys.argv[1]
    local_rank = int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local_rank)
    dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local_rank}"))
    import helpers
    from rfest_b200 import GLM

    stim, y, _ = helpers.lnp_data(5003, (8, 5, 4), seed=31)
    m = build_and_fit(GLM, stim, y, 4001, 40, device=local_rank)
    assert m.engine.world == dist.get_world_size()
    last = m.all_params[-1]
    # every rank must hold identical (replicated) parameters and traces
    vec = torch.tensor(np.concatenate([m.cost_train, m.cost_dev, last["stimulus"].ravel()]),
                       device=f"cuda:{local_rank}", dtype=torch.float64)
    ref = vec.clone()
    dist.broadcast(ref, src=0)
    assert torch.equal(vec, ref), "ranks diverged"
    pred = np.asarray(m.y_pred["opt"]["train"])
    gathered = [None] * dist.get_world_size()
    dist.all_gather_object(gathered, pred)
    if dist.get_rank() == 0:
        np.savez(out_path, cost_train=m.cost_train, cost_dev=m.cost_dev, metric_train=m.metric_train,
                 metric_dev=m.metric_dev, b_stim=last["stimulus"], b_hist=last["history"],
                 icpt=np.array([last["intercept"][k] for k in ("stimulus", "history", "global")]),
                 y_pred_train=np.concatenate(gathered), best=m.best_iteration)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
