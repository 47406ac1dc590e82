"""Generate `tests/golden/rfest_numpy_golden.npz` from the REFERENCE's own code.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  Run in the build container
(the only place `/root/reference` exists):

    python -m oracle.make_golden

It imports `rfest.utils`, `rfest.splines`, `rfest.metrics` of the reference
through `oracle/ref_loader.py` (JAX-free modules) and stores their outputs on
small seeded inputs.  The vectors travel to the GPU box with the repository;
the reference does not.  Everything stored here was computed by reference code,
nothing by the oracle or the product.

Recorded with: see the `env` entry of the archive (numpy / scipy versions; the
spline bases go through LAPACK `gesv`, so their last bit depends on the build).
"""

import hashlib
import os

import numpy as np

from . import ref_loader

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                   "tests", "golden", "rfest_numpy_golden.npz")

# (dims, df) pairs used by BASELINE.json configs, the reference README and tests/test_glm.py
SPLINE_CASES_SMALL = {
    "c1_stim": ([25], [8]),
    "c1_hist": ([20], [8]),
    "test_hist": ([5], [3]),
    "test_3d": ([6, 7, 8], [3, 4, 4]),
    "two_d": ([9, 11], [4, 5]),
}
SPLINE_CASES_BIG = {
    "c2_3d": ([25, 20, 15], [8, 6, 6]),
    "c5_3d": ([40, 32, 32], [12, 10, 10]),
}
CR_1D = [(5, 3), (25, 8), (20, 8), (20, 6), (15, 6), (40, 12), (32, 10), (6, 3), (7, 4), (8, 4)]


def sha16(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()[:16]


def main():
    utils, splines, metrics = ref_loader.load()
    import scipy

    g = {}
    g["env"] = np.array(f"numpy {np.__version__} scipy {scipy.__version__}")

    # ---- design indexing (rfest/utils.py:5-67) --------------------------------
    g["design_doctest"] = utils.build_design_matrix(np.array([0, 1, 2])[:, None], 2)
    ramp = np.arange(1, 7, dtype=np.float64)[:, None]
    for sh in (0, 1, -1, 2, -2):
        g[f"design_ramp_nlag3_shift{sh}"] = utils.build_design_matrix(ramp, 3, shift=sh)
    g["design_ramp_nlag2_shift-3"] = utils.build_design_matrix(ramp, 2, shift=-3)
    rng = np.random.default_rng(20461)
    stim2 = rng.standard_normal((40, 3, 4)).astype(np.float32)
    g["design_in_stim2"] = stim2
    g["design_stim2_nlag5_shift0_f32"] = utils.build_design_matrix(stim2, 5, shift=0, dtype=np.float32)
    g["design_stim2_nlag5_shift1_f32"] = utils.build_design_matrix(stim2, 5, shift=1, dtype=np.float32)
    g["design_stim2_nlag4_shift-2_f64"] = utils.build_design_matrix(stim2, 4, shift=-2)

    # ---- spline bases (rfest/splines.py) ----------------------------------------
    for d, f in CR_1D:
        g[f"cr_{d}_{f}"] = splines.cr(np.arange(d), f)
    g["bs_20_7"] = splines.bs(np.arange(20), 7)
    g["cc_20_7"] = splines.cc(np.arange(20), 7)
    g["tp_15_5"] = splines.tp(np.arange(15), 5)
    for key, (dims, df) in SPLINE_CASES_SMALL.items():
        g[f"S_{key}_f64"] = splines.build_spline_matrix(dims, df, "cr")
        g[f"S_{key}_f32"] = splines.build_spline_matrix(dims, df, "cr", dtype=np.float32)
    for key, (dims, df) in SPLINE_CASES_BIG.items():
        S64 = splines.build_spline_matrix(dims, df, "cr")
        S32 = splines.build_spline_matrix(dims, df, "cr", dtype=np.float32)
        # too large to store: keep hashes, sums and a strided sample
        g[f"S_{key}_shape"] = np.array(S64.shape)
        g[f"S_{key}_sha_f64"] = np.array(sha16(S64))
        g[f"S_{key}_sha_f32"] = np.array(sha16(S32))
        g[f"S_{key}_sum"] = np.array(S64.sum())
        g[f"S_{key}_colsum"] = S64.sum(0)
        g[f"S_{key}_sample_f64"] = S64[::211, ::11].copy()
        g[f"S_{key}_sample_f32"] = S32[::211, ::11].copy()
    g["S_bs_2d_f64"] = splines.build_spline_matrix([12, 10], [6, 5], "bs")

    # ---- lag + burn-in + projection, exactly as add_design_matrix does ---------
    # (rfest/GLM/_glm.py:250-258,297-298), with the reference's own functions
    def project(stim, dims, df, shift, burn_in, dtype):
        X = utils.build_design_matrix(stim.astype(dtype), dims[0], shift=shift, dtype=dtype)[burn_in:]
        S = splines.build_spline_matrix(dims, df, "cr", dtype=dtype)
        return X @ S

    stim3 = rng.standard_normal((160, 7, 8)).astype(np.float32)
    g["proj_in_stim3"] = stim3
    g["proj_stim3_f64"] = project(stim3.reshape(160, -1), [6, 7, 8], [3, 4, 4], 0, 5, np.float64)
    g["proj_stim3_f32"] = project(stim3.reshape(160, -1), [6, 7, 8], [3, 4, 4], 0, 5, np.float32)
    stim1 = rng.standard_normal(400).astype(np.float32)
    spikes = rng.poisson(0.3, 400).astype(np.float32)
    g["proj_in_stim1"] = stim1
    g["proj_in_spikes"] = spikes
    g["proj_stim1_f64"] = project(stim1[:, None], [25], [8], 0, 24, np.float64)
    g["proj_stim1_f32"] = project(stim1[:, None], [25], [8], 0, 24, np.float32)
    # history filter: nlag 20, shift 1, burn-in 24 inherited from the stimulus filter
    g["proj_hist_f64"] = project(spikes[:, None], [20], [8], 1, 24, np.float64)
    g["proj_hist_f32"] = project(spikes[:, None], [20], [8], 1, 24, np.float32)
    # user-supplied burn_in = 0 keeps the zero-padded rows
    g["proj_stim1_burn0_f64"] = project(stim1[:, None], [25], [8], 0, 0, np.float64)

    # ---- metrics (rfest/metrics.py:8-20) ----------------------------------------
    ya = rng.standard_normal(500)
    yb = 0.7 * ya + 0.3 * rng.standard_normal(500)
    g["metric_in_y"] = ya
    g["metric_in_yhat"] = yb
    g["metric_r2"] = np.asarray(metrics.r2(ya, yb))
    g["metric_mse"] = np.asarray(metrics.mse(ya, yb))
    g["metric_corrcoef"] = np.asarray(metrics.corrcoef(ya, yb))
    ya32, yb32 = ya.astype(np.float32), yb.astype(np.float32)
    g["metric_r2_f32"] = np.asarray(metrics.r2(ya32, yb32))
    g["metric_mse_f32"] = np.asarray(metrics.mse(ya32, yb32))
    g["metric_corrcoef_f32"] = np.asarray(metrics.corrcoef(ya32, yb32))

    # ---- uvec (rfest/utils.py:119-121) ------------------------------------------
    g["uvec_in"] = rng.standard_normal((6, 4))
    g["uvec_out"] = utils.uvec(g["uvec_in"])

    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    np.savez_compressed(OUT, **g)
    print(f"wrote {OUT}: {len(g)} arrays, {os.path.getsize(OUT) / 1024:.1f} KiB")


if __name__ == "__main__":
    main()
