"""CPU-only checks: the C-ABI library builds/loads and exports every symbol the header
declares; the product's host-side logic (spline bases, sharding, metrics from sums)."""

import ctypes
import os
import re

import numpy as np
import pytest

from rfest_b200 import _lib, shard, splines
from rfest_b200 import metrics as pmetrics

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_functions():
    text = open(os.path.join(ROOT, "include", "rfest_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(rfb_[a-z0-9_]+)\s*\(", text)))


def test_library_loads_and_exports_every_declared_symbol():
    from rfest_b200 import build

    build.build()
    lib = _lib.load()
    names = header_functions()
    assert len(names) >= 30
    for name in names:
        assert hasattr(lib, name), f"{name} declared in include/rfest_b200.h but not exported"
    assert sorted(_lib.SIGNATURES) == names, "ctypes table and header disagree"
    assert lib.rfb_abi_version() == 1


def test_header_constants_match_ctypes_table():
    text = open(os.path.join(ROOT, "include", "rfest_b200.h")).read()
    defs = dict(re.findall(r"#define\s+(RFB_[A-Z0-9_]+)\s+\(?(-?\d+)\)?", text))
    for name, code in _lib.NL.items():
        key = "RFB_NL_" + name.upper()
        assert int(defs[key]) == code
    assert int(defs["RFB_DISTR_GAUSSIAN"]) == _lib.DISTR["gaussian"]
    assert int(defs["RFB_DISTR_POISSON"]) == _lib.DISTR["poisson"]
    for k, v in _lib.KIND.items():
        assert int(defs["RFB_KIND_" + k.upper()]) == v
    assert int(defs["RFB_N_SUMS"]) == _lib.N_SUMS
    assert int(defs["RFB_ERR_INVALID"]) == _lib.ERR_INVALID
    assert int(defs["RFB_ERR_UNSUPPORTED"]) == _lib.ERR_UNSUPPORTED


def test_no_gpu_means_loud_failure():
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from rfest_b200 import GLM

    with pytest.raises(Exception) as ei:
        GLM(distr="poisson", output_nonlinearity="softplus")
    assert "CUDA" in str(ei.value) or "cuda" in str(ei.value)


def test_product_does_not_import_the_oracle():
    pkg = os.path.join(ROOT, "rfest_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f
                assert "/root/reference" not in src, f


# ---- product spline code: identical bases ------------------------------------------------
CR_1D = [(5, 3), (25, 8), (20, 8), (20, 6), (15, 6), (40, 12), (32, 10), (6, 3), (7, 4), (8, 4)]


@pytest.mark.parametrize("d,f", CR_1D)
def test_product_cr_matches_reference_vectors(golden, d, f):
    np.testing.assert_allclose(splines.cr(np.arange(d), f), golden[f"cr_{d}_{f}"], rtol=0, atol=1e-15)


def test_product_bases_bit_identical_to_live_reference(reference_modules):
    _, ref, _ = reference_modules
    for d, f in CR_1D + [(30, 10), (11, 3)]:
        assert np.array_equal(splines.cr(np.arange(d), f), ref.cr(np.arange(d), f))
    assert np.array_equal(splines.bs(np.arange(20), 7), ref.bs(np.arange(20), 7))
    assert np.array_equal(splines.cc(np.arange(20), 7), ref.cc(np.arange(20), 7))
    assert np.array_equal(splines.tp(np.arange(15), 5), ref.tp(np.arange(15), 5))
    for dims, df, kind in [([25], [8], "cr"), ([6, 7, 8], [3, 4, 4], "cr"), ([9, 11], [4, 5], "cr"),
                           ([12, 10], [6, 5], "bs"), ([10, 8], [5, 4], "cc")]:
        for dt in (np.float64, np.float32):
            assert np.array_equal(splines.build_spline_matrix(dims, df, kind, dtype=dt),
                                  ref.build_spline_matrix(dims, df, kind, dtype=dt))


def test_product_spline_matrix_golden(golden):
    for key, dims, df in [("c1_stim", [25], [8]), ("c1_hist", [20], [8]),
                          ("test_3d", [6, 7, 8], [3, 4, 4]), ("two_d", [9, 11], [4, 5])]:
        np.testing.assert_allclose(splines.build_spline_matrix(dims, df, "cr"),
                                   golden[f"S_{key}_f64"], rtol=0, atol=1e-16)
    other = golden["bs_20_7"], golden["cc_20_7"]
    np.testing.assert_allclose(splines.bs(np.arange(20), 7), other[0], rtol=0, atol=1e-15)
    np.testing.assert_allclose(splines.cc(np.arange(20), 7), other[1], rtol=0, atol=1e-15)


def test_kron_basis_is_the_dense_matrix(golden):
    kb = splines.KronBasis([6, 7, 8], [3, 4, 4], "cr")
    assert kb.shape == (336, 48)
    np.testing.assert_allclose(np.asarray(kb), golden["S_test_3d_f64"], rtol=0, atol=1e-16)
    # the separable factors multiply back to the dense basis
    St, Sxy = kb.temporal_f32().astype(np.float64), kb.spatial_f32().astype(np.float64)
    np.testing.assert_allclose(np.kron(St, Sxy), golden["S_test_3d_f64"], rtol=0, atol=2e-8)
    b = np.random.default_rng(0).standard_normal((48, 1))
    np.testing.assert_allclose(kb @ b, golden["S_test_3d_f64"] @ b, rtol=0, atol=1e-14)
    kb1 = splines.KronBasis([25], [8], "cr")
    assert kb1.spatial_f32() is None
    np.testing.assert_allclose(kb1 @ b[:8], golden["S_c1_stim_f64"] @ b[:8], rtol=0, atol=1e-14)
    big = splines.KronBasis([25, 20, 15], [8, 6, 6], "cr")
    assert big.shape == tuple(golden["S_c2_3d_shape"])
    np.testing.assert_allclose((big @ np.ones((288, 1))).sum(), golden["S_c2_3d_sum"], rtol=1e-12)


# ---- sharding ---------------------------------------------------------------------------------
def test_shard_rows_partition():
    for n in (1, 7, 64, 1000, 33554408):
        for world in (1, 2, 3, 4, 8):
            edges = [shard.shard_rows(n, r, world) for r in range(world)]
            assert edges[0][0] == 0 and edges[-1][1] == n
            for a, b in zip(edges, edges[1:]):
                assert a[1] == b[0]
            sizes = [hi - lo for lo, hi in edges]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard.shard_rows(10, 3, 2)


def test_source_rows_cover_exactly_what_the_design_reads():
    from oracle import design

    rng = np.random.default_rng(3)
    n_raw = 60
    x = rng.standard_normal((n_raw, 2))
    for nlag, shift, burn in [(5, 0, 4), (7, 1, 4), (4, -2, 0), (3, 2, 6), (2, -3, 0)]:
        full = design.lagged_design(x, nlag, shift)[burn:]
        n = n_raw - burn
        for world in (1, 2, 3):
            for rank in range(world):
                lo, hi = shard.shard_rows(n, rank, world)
                s_lo, s_hi = shard.source_rows(burn + lo, hi - lo, nlag, shift, n_raw)
                # rebuild the shard from only those raw rows (zero elsewhere)
                masked = np.zeros_like(x)
                masked[s_lo:s_hi] = x[s_lo:s_hi]
                part = design.lagged_design(masked, nlag, shift)[burn:][lo:hi]
                assert np.array_equal(part, full[lo:hi])
                assert shard.front_padding(nlag, shift) == design.front_padding(nlag, shift)


# ---- metrics from sums ---------------------------------------------------------------------------
def test_metrics_from_sums_match_reference_metrics(golden):
    y, r = golden["metric_in_y"], golden["metric_in_yhat"]
    args = dict(n=len(y), sy=y.sum(), syy=(y * y).sum(), sr=r.sum(), srr=(r * r).sum(),
                syr=(y * r).sum(), sse=((y - r) ** 2).sum())
    assert np.isclose(pmetrics.score_from_sums("mse", **args), golden["metric_mse"], rtol=1e-13)
    assert np.isclose(pmetrics.score_from_sums("r2", **args), golden["metric_r2"], rtol=1e-12)
    assert np.isclose(pmetrics.score_from_sums("corrcoef", **args), golden["metric_corrcoef"], rtol=1e-12)
    assert pmetrics.r2(y, r) == golden["metric_r2"]
    assert pmetrics.mse(y, r) == golden["metric_mse"]
    assert pmetrics.corrcoef(y, r) == golden["metric_corrcoef"]
