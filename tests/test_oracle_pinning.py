"""Pin the oracle (design indexing, spline bases, metrics) to the reference.

Two anchors: (1) golden vectors computed by the reference's own NumPy modules
(tests/golden/, oracle/make_golden.py) -- available everywhere; (2) the live
reference modules -- only in the build container (skipped on the GPU box).
"""

import hashlib

import numpy as np
import pytest

from oracle import design, metrics, splines


def sha16(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()[:16]


# ---- design indexing: bit-exact -------------------------------------------------
def test_design_doctest(golden):
    out = design.lagged_design(np.array([0, 1, 2])[:, None], 2)
    assert np.array_equal(out, np.array([[0, 0], [0, 1], [1, 2]]))     # rfest/utils.py:44-46
    assert np.array_equal(out, golden["design_doctest"])


@pytest.mark.parametrize("shift", [0, 1, -1, 2, -2])
def test_design_ramp_shifts(golden, shift):
    ramp = np.arange(1, 7, dtype=np.float64)[:, None]
    out = design.lagged_design(ramp, 3, shift=shift)
    assert np.array_equal(out, golden[f"design_ramp_nlag3_shift{shift}"])


def test_design_known_answers():
    ramp = np.arange(1, 7, dtype=np.float64)[:, None]
    # SURVEY.md section 8a, probed from the reference
    assert design.lagged_design(ramp, 3, shift=1).tolist() == [
        [0, 0, 0], [0, 0, 1], [0, 1, 2], [1, 2, 3], [2, 3, 4], [3, 4, 5]]
    assert design.lagged_design(ramp, 3, shift=-1).tolist() == [
        [0, 1, 2], [1, 2, 3], [2, 3, 4], [3, 4, 5], [4, 5, 6], [5, 6, 0]]


def test_design_saturating_negative_shift(golden):
    ramp = np.arange(1, 7, dtype=np.float64)[:, None]
    assert np.array_equal(design.lagged_design(ramp, 2, shift=-3), golden["design_ramp_nlag2_shift-3"])


def test_design_multipixel(golden):
    stim = golden["design_in_stim2"]
    for nlag, shift, dt, key in [(5, 0, np.float32, "design_stim2_nlag5_shift0_f32"),
                                 (5, 1, np.float32, "design_stim2_nlag5_shift1_f32"),
                                 (4, -2, np.float64, "design_stim2_nlag4_shift-2_f64")]:
        out = design.lagged_design(stim, nlag, shift=shift, dtype=dt)
        assert out.dtype == golden[key].dtype
        assert np.array_equal(out, golden[key])


def test_design_source_row_map(golden):
    stim = golden["design_in_stim2"].reshape(40, -1)
    n, n_pix = stim.shape
    for nlag, shift in [(5, 0), (5, 1), (4, -2), (2, -3)]:
        out = design.lagged_design(stim, nlag, shift=shift, dtype=np.float32)
        for t in (0, 1, 7, n - 2, n - 1):
            for i in range(nlag):
                src = design.lag_source_row(t, i, nlag, shift)
                want = stim[src] if 0 <= src < n else np.zeros(n_pix, np.float32)
                assert np.array_equal(out[t, i * n_pix:(i + 1) * n_pix], want)


def test_design_live_reference(reference_modules):
    utils, _, _ = reference_modules
    rng = np.random.default_rng(7)
    for n, shape, nlag, shift in [(30, (1,), 25, 0), (64, (5,), 7, 1), (50, (3, 2), 4, -1),
                                  (20, (2,), 3, -5), (12, (1,), 1, 0), (9, (4,), 12, 2)]:
        X = rng.standard_normal((n,) + shape).astype(np.float32)
        for dt in (np.float32, np.float64):
            assert np.array_equal(design.lagged_design(X, nlag, shift, dtype=dt),
                                  utils.build_design_matrix(X, nlag, shift=shift, dtype=dt))


# ---- spline bases ----------------------------------------------------------------
CR_1D = [(5, 3), (25, 8), (20, 8), (20, 6), (15, 6), (40, 12), (32, 10), (6, 3), (7, 4), (8, 4)]


@pytest.mark.parametrize("d,f", CR_1D)
def test_cr_golden(golden, d, f):
    out = splines.cr(np.arange(d), f)
    assert out.shape == (d, f)
    np.testing.assert_allclose(out, golden[f"cr_{d}_{f}"], rtol=0, atol=1e-15)
    assert abs(np.linalg.norm(out) - 1) < 1e-14


def test_cr_5_3_known_answer():
    # SURVEY.md section 8c
    want = np.array([[.4826375819, 0, 0], [.1960715177, .3318133376, -.0452472733],
                     [0, .4826375819, 0], [-.0452472733, .3318133376, .1960715177],
                     [0, 0, .4826375819]])
    np.testing.assert_allclose(splines.cr(np.arange(5), 3), want, atol=5e-11)


def test_other_bases_golden(golden):
    np.testing.assert_allclose(splines.bs(np.arange(20), 7), golden["bs_20_7"], rtol=0, atol=1e-15)
    np.testing.assert_allclose(splines.cc(np.arange(20), 7), golden["cc_20_7"], rtol=0, atol=1e-15)
    # tp goes through an SVD: compare up to column signs
    out, ref = splines.tp(np.arange(15), 5), golden["tp_15_5"]
    sign = np.sign((out * ref).sum(0))
    np.testing.assert_allclose(out * sign, ref, atol=1e-12)


@pytest.mark.parametrize("key,dims,df", [("c1_stim", [25], [8]), ("c1_hist", [20], [8]),
                                         ("test_hist", [5], [3]), ("test_3d", [6, 7, 8], [3, 4, 4]),
                                         ("two_d", [9, 11], [4, 5])])
def test_spline_matrix_small_golden(golden, key, dims, df):
    S64 = splines.build_spline_matrix(dims, df, "cr")
    S32 = splines.build_spline_matrix(dims, df, "cr", dtype=np.float32)
    np.testing.assert_allclose(S64, golden[f"S_{key}_f64"], rtol=0, atol=1e-16)
    assert S32.dtype == np.float32
    np.testing.assert_allclose(S32, golden[f"S_{key}_f32"], rtol=0, atol=1e-9)


@pytest.mark.parametrize("key,dims,df", [("c2_3d", [25, 20, 15], [8, 6, 6]),
                                         ("c5_3d", [40, 32, 32], [12, 10, 10])])
def test_spline_matrix_big_golden(golden, key, dims, df):
    S64 = splines.build_spline_matrix(dims, df, "cr")
    assert tuple(golden[f"S_{key}_shape"]) == S64.shape
    np.testing.assert_allclose(S64.sum(), golden[f"S_{key}_sum"], rtol=1e-13)
    np.testing.assert_allclose(S64.sum(0), golden[f"S_{key}_colsum"], rtol=0, atol=1e-13)
    np.testing.assert_allclose(S64[::211, ::11], golden[f"S_{key}_sample_f64"], rtol=0, atol=1e-16)
    S32 = S64.astype(np.float32)
    np.testing.assert_allclose(S32[::211, ::11], golden[f"S_{key}_sample_f32"], rtol=0, atol=1e-9)


def test_spline_matrix_survey_values():
    # sums and one entry recorded from the reference
    S = splines.build_spline_matrix([25], [8], "cr")
    assert abs(S.sum() - 5.394848981839) < 1e-11
    S3 = splines.build_spline_matrix([25, 20, 15], [8, 6, 6], "cr")
    assert S3.shape == (7500, 288)
    assert abs(S3.sum() - 109.335462988895) < 1e-9
    assert abs(S3[1, 1] - 7.966642203383e-03) < 1e-14


def test_spline_hash_tripwire(golden):
    """Bit-for-bit equality with the reference (same LAPACK build as the recording)."""
    for key, dims, df in [("c2_3d", [25, 20, 15], [8, 6, 6])]:
        S64 = splines.build_spline_matrix(dims, df, "cr")
        assert sha16(S64) == str(golden[f"S_{key}_sha_f64"])
        assert sha16(S64.astype(np.float32)) == str(golden[f"S_{key}_sha_f32"])


def test_splines_live_reference(reference_modules):
    _, ref, _ = reference_modules
    for d, f in CR_1D + [(30, 10), (11, 3)]:
        assert np.array_equal(splines.cr(np.arange(d), f), ref.cr(np.arange(d), f))
    assert np.array_equal(splines.bs(np.arange(20), 7), ref.bs(np.arange(20), 7))
    assert np.array_equal(splines.cc(np.arange(20), 7), ref.cc(np.arange(20), 7))
    assert np.array_equal(splines.tp(np.arange(15), 5), ref.tp(np.arange(15), 5))
    for dims, df, kind in [([25], [8], "cr"), ([6, 7, 8], [3, 4, 4], "cr"), ([9, 11], [4, 5], "cr"),
                           ([12, 10], [6, 5], "bs"), ([10, 8], [5, 4], "cc"), ([8, 6], [3, 3], "tp")]:
        for dt in (np.float64, np.float32):
            assert np.array_equal(splines.build_spline_matrix(dims, df, kind, dtype=dt),
                                  ref.build_spline_matrix(dims, df, kind, dtype=dt))


def test_spline_errors():
    with pytest.raises(ValueError):
        splines.build_spline_matrix([5, 5], [3], "cr")
    with pytest.raises(ValueError):
        splines.build_spline_matrix([5], [3], "nope")
    with pytest.raises(NotImplementedError):
        splines.build_spline_matrix([5, 5, 5, 5], [3, 3, 3, 3], "cr")


# ---- projection as add_design_matrix does it ---------------------------------------
def test_projection_golden(golden):
    cases = [("proj_in_stim3", (160, -1), [6, 7, 8], [3, 4, 4], 0, 5, "proj_stim3"),
             ("proj_in_stim1", (400, 1), [25], [8], 0, 24, "proj_stim1"),
             ("proj_in_spikes", (400, 1), [20], [8], 1, 24, "proj_hist")]
    for src, shape, dims, df, shift, burn, key in cases:
        stim = golden[src].reshape(shape)
        for dt, tag, tol in [(np.float64, "f64", 1e-13), (np.float32, "f32", 2e-5)]:
            X = design.trimmed_design(stim.astype(dt), dims[0], shift, burn, dtype=dt)
            S = splines.build_spline_matrix(dims, df, "cr", dtype=dt)
            XS = design.project_dense(X, S)
            assert XS.dtype == dt
            np.testing.assert_allclose(XS, golden[f"{key}_{tag}"], rtol=0, atol=tol)
    X0 = design.trimmed_design(golden["proj_in_stim1"].reshape(400, 1).astype(np.float64), 25, 0, 0)
    np.testing.assert_allclose(X0 @ splines.build_spline_matrix([25], [8], "cr"),
                               golden["proj_stim1_burn0_f64"], rtol=0, atol=1e-13)


def test_history_rows_after_burn_in(golden):
    """nlag 20, shift 1, burn-in 24: trimmed row t' sees y[t'+4 .. t'+23]."""
    y = golden["proj_in_spikes"].astype(np.float64)
    X = design.trimmed_design(y[:, None], 20, 1, 24)
    for tp in (0, 5, 100, len(y) - 25):
        assert np.array_equal(X[tp], y[tp + 4: tp + 24])


# ---- metrics ---------------------------------------------------------------------------
def test_metrics_golden(golden):
    y, yh = golden["metric_in_y"], golden["metric_in_yhat"]
    assert metrics.r2(y, yh) == golden["metric_r2"]
    assert metrics.mse(y, yh) == golden["metric_mse"]
    assert metrics.corrcoef(y, yh) == golden["metric_corrcoef"]
    y32, yh32 = y.astype(np.float32), yh.astype(np.float32)
    assert metrics.r2(y32, yh32) == golden["metric_r2_f32"]
    assert metrics.mse(y32, yh32) == golden["metric_mse_f32"]
    assert metrics.corrcoef(y32, yh32) == golden["metric_corrcoef_f32"]


def test_uvec_golden(golden):
    assert np.array_equal(splines.unit_frobenius(golden["uvec_in"]), golden["uvec_out"])
