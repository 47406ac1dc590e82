"""GPU parity: fused lag + spline projection (K1) against the reference's golden vectors
and the oracle.  Runs on the B200 box (`-m gpu`); everything goes through the C ABI."""

import numpy as np
import pytest

from oracle import design, splines as osplines

pytestmark = pytest.mark.gpu


def _glm(**kw):
    from rfest_b200 import GLM

    return GLM(**kw)


def _oracle_xs(stim, dims, df, shift, burn_in, smooth="cr", dtype=np.float64):
    X = design.trimmed_design(np.asarray(stim, dtype).reshape(len(stim), -1), dims[0], shift, burn_in, dtype=dtype)
    S = osplines.build_spline_matrix(dims, df, smooth, dtype=dtype)
    return X @ S


def test_projection_against_reference_golden(golden):
    """Outputs of the reference's own build_design_matrix(...)[burn:] @ build_spline_matrix(...)."""
    m = _glm(distr="gaussian")
    m.add_design_matrix(golden["proj_in_stim3"], dims=[6, 7, 8], df=[3, 4, 4], smooth="cr", name="stimulus")
    XS = np.asarray(m.XS["train"]["stimulus"])
    assert XS.shape == golden["proj_stim3_f64"].shape and XS.dtype == np.float32
    scale = np.abs(golden["proj_stim3_f64"]).max()
    assert np.abs(XS - golden["proj_stim3_f64"]).max() <= 1e-6 * scale     # vs fp64 truth
    assert np.abs(XS - golden["proj_stim3_f32"]).max() <= 2e-6 * scale     # vs the reference's fp32

    m = _glm(distr="poisson", output_nonlinearity="softplus")
    m.add_design_matrix(golden["proj_in_stim1"], dims=[25], df=[8], smooth="cr", name="stimulus")
    m.add_design_matrix(golden["proj_in_spikes"], dims=[20], df=[8], smooth="cr", shift=1, name="history")
    assert m.burn_in == 24
    for name, key in (("stimulus", "proj_stim1"), ("history", "proj_hist")):
        XS = np.asarray(m.XS["train"][name])
        ref64 = golden[key + "_f64"]
        assert XS.shape == ref64.shape
        scale = np.abs(ref64).max()
        assert np.abs(XS - ref64).max() <= 1e-6 * scale
        assert np.abs(XS - golden[key + "_f32"]).max() <= 2e-6 * scale

    m = _glm(distr="gaussian")
    m.add_design_matrix(golden["proj_in_stim1"], dims=[25], df=[8], smooth="cr", burn_in=0)
    ref = golden["proj_stim1_burn0_f64"]
    assert np.abs(np.asarray(m.XS["train"]["stimulus"]) - ref).max() <= 1e-6 * np.abs(ref).max()


@pytest.mark.parametrize("nlag,shift,burn", [(5, 0, None), (5, 1, None), (4, -1, None), (3, 2, 0),
                                             (2, -3, 0), (7, 0, 2), (1, 0, None)])
def test_lag_shift_indexing_is_bit_exact(nlag, shift, burn):
    """With no basis the design IS the lagged stimulus: integer-valued input must come back
    exactly (the lag/shift/zero-padding map of rfest/utils.py:50-67)."""
    rng = np.random.default_rng(nlag * 10 + shift)
    stim = rng.integers(-9, 10, size=(57, 3)).astype(np.float32)
    m = _glm(distr="gaussian")
    m.add_design_matrix(stim, dims=[nlag, 3], smooth=None, shift=shift, burn_in=burn, name="stimulus")
    got = m._blocks["train"]["stimulus"].read()
    want = design.trimmed_design(stim, nlag, shift, m.burn_in, dtype=np.float32)
    assert got.shape == want.shape
    assert np.array_equal(got, want)
    assert m.n_features["stimulus"] == nlag * 3
    assert np.array_equal(np.asarray(m.X["train"]["stimulus"]), want)


@pytest.mark.parametrize("dims,df,shape", [([8], [4], (300,)), ([6, 5], [3, 3], (211, 5)),
                                           ([7, 4, 6], [4, 3, 3], (150, 4, 6)),
                                           ([25, 20, 15], [8, 6, 6], (100, 20, 15)),
                                           ([12, 9, 13], [5, 4, 6], (130, 9, 13))])
def test_projection_vs_oracle(dims, df, shape):
    rng = np.random.default_rng(sum(dims))
    stim = rng.standard_normal(shape).astype(np.float32)
    for shift in (0, 1):
        m = _glm(distr="gaussian")
        m.add_design_matrix(stim, dims=dims, df=df, smooth="cr", shift=shift, name="stimulus")
        XS = np.asarray(m.XS["train"]["stimulus"])
        ref = _oracle_xs(stim, dims, df, shift, dims[0] - 1)
        assert XS.shape == ref.shape
        assert np.abs(XS - ref).max() <= 2e-6 * np.abs(ref).max()
        assert m.n_features["stimulus"] == int(np.prod(df))


def test_other_bases_and_instantaneous_filter():
    rng = np.random.default_rng(5)
    stim = rng.standard_normal((120, 6)).astype(np.float32)
    for smooth in ("bs", "cc", "tp"):
        m = _glm(distr="gaussian")
        m.add_design_matrix(stim, dims=[10, 6], df=[6, 4], smooth=smooth, name="stimulus")
        ref = _oracle_xs(stim, [10, 6], [6, 4], 0, 9, smooth=smooth)
        XS = np.asarray(m.XS["train"]["stimulus"])
        assert XS.shape == ref.shape
        assert np.abs(XS - ref).max() <= 3e-6 * np.abs(ref).max()
    # lag=False: instantaneous receptive field, no burn-in (_glm.py:259-261)
    m = _glm(distr="gaussian")
    m.add_design_matrix(stim, dims=[6], df=[4], smooth="cr", lag=False, name="stimulus")
    assert m.burn_in == 0
    S = osplines.build_spline_matrix([6], [4], "cr")
    ref = stim.astype(np.float64) @ S
    assert np.abs(np.asarray(m.XS["train"]["stimulus"]) - ref).max() <= 2e-6 * np.abs(ref).max()


def test_dev_and_test_reuse_the_training_basis():
    rng = np.random.default_rng(8)
    tr, dv = rng.standard_normal((90, 4, 3)).astype(np.float32), rng.standard_normal((50, 4, 3)).astype(np.float32)
    m = _glm(distr="gaussian")
    m.add_design_matrix(tr, dims=[5, 4, 3], df=[3, 3, 3], smooth="cr", name="stimulus")
    m.add_design_matrix(dv, name="stimulus", kind="dev")
    ref = _oracle_xs(dv, [5, 4, 3], [3, 3, 3], 0, 4)
    got = np.asarray(m.XS["dev"]["stimulus"])
    assert got.shape == ref.shape and np.abs(got - ref).max() <= 2e-6 * np.abs(ref).max()
    assert m.filter_names == ["stimulus"]


def test_long_recording_crosses_chunk_boundaries():
    """> 2^20 rows: the projection is chunked internally; rows next to the seams must agree."""
    rng = np.random.default_rng(11)
    n = (1 << 20) + 70001
    stim = rng.standard_normal(n).astype(np.float32)
    m = _glm(distr="gaussian")
    m.add_design_matrix(stim, dims=[25], df=[8], smooth="cr", name="stimulus")
    blk = m._blocks["train"]["stimulus"]
    S = osplines.build_spline_matrix([25], [8], "cr")
    for row0 in (0, (1 << 20) - 40, (1 << 20) - 24 - 3, n - 24 - 30):
        got = blk.read(row0, 30)
        want = np.stack([stim[r:r + 25].astype(np.float64) @ S for r in range(row0, row0 + 30)])
        assert np.abs(got - want).max() <= 2e-6 * np.abs(want).max()


def test_sharded_projection_with_halo_equals_whole():
    """What each rank of a multi-GPU run does: project only its rows from its raw rows + halo."""
    from rfest_b200.engine import Block, Engine
    from rfest_b200.shard import shard_rows, source_rows
    from rfest_b200.splines import KronBasis

    eng = Engine.get()
    rng = np.random.default_rng(2)
    n_raw, dims, df = 500, [9, 5, 4], [4, 3, 3]
    stim = rng.standard_normal((n_raw, 20)).astype(np.float32)
    kb = KronBasis(dims, df, "cr")
    for shift, burn in ((0, 8), (1, 8), (-2, 3)):
        n = n_raw - burn
        whole = Block(eng, n, kb.n_coef)
        whole.project(stim, dst_row0=0, n_out=n, col0=0, src0=0, n_raw_total=n_raw, n_pix=20,
                      first_t=burn, nlag=dims[0], shift=shift, St=kb.temporal_f32(), Sxy=kb.spatial_f32())
        ref = whole.read()
        oracle = _oracle_xs(stim, dims, df, shift, burn)
        assert np.abs(ref - oracle).max() <= 2e-6 * np.abs(oracle).max()
        for world in (2, 3):
            parts = []
            for rank in range(world):
                lo, hi = shard_rows(n, rank, world)
                s_lo, s_hi = source_rows(burn + lo, hi - lo, dims[0], shift, n_raw)
                blk = Block(eng, hi - lo, kb.n_coef)
                blk.project(stim[s_lo:s_hi], dst_row0=0, n_out=hi - lo, col0=0, src0=s_lo,
                            n_raw_total=n_raw, n_pix=20, first_t=burn + lo, nlag=dims[0], shift=shift,
                            St=kb.temporal_f32(), Sxy=kb.spatial_f32())
                parts.append(blk.read())
            assert np.array_equal(np.vstack(parts), ref)      # bit-identical to the unsharded run
        # a shard that does not bring its halo is refused
        lo, hi = shard_rows(n, 1, 2)
        blk = Block(eng, hi - lo, kb.n_coef)
        with pytest.raises(ValueError, match="halo"):
            blk.project(stim[burn + lo:], dst_row0=0, n_out=hi - lo, col0=0, src0=burn + lo,
                        n_raw_total=n_raw, n_pix=20, first_t=burn + lo, nlag=dims[0], shift=shift,
                        St=kb.temporal_f32(), Sxy=kb.spatial_f32())


def test_device_resident_input():
    import torch

    rng = np.random.default_rng(4)
    stim = rng.standard_normal((400, 6, 5)).astype(np.float32)
    a = _glm(distr="gaussian")
    a.add_design_matrix(stim, dims=[6, 6, 5], df=[3, 3, 3], smooth="cr")
    b = _glm(distr="gaussian")
    b.add_design_matrix(torch.from_numpy(stim).cuda(), dims=[6, 6, 5], df=[3, 3, 3], smooth="cr")
    assert np.array_equal(np.asarray(a.XS["train"]["stimulus"]), np.asarray(b.XS["train"]["stimulus"]))


def test_design_errors():
    m = _glm(distr="gaussian")
    with pytest.raises(ValueError):
        m.add_design_matrix(np.zeros((10, 3), np.float32), dims=[25, 3], df=[8, 2], smooth="cr")  # empty after burn-in
    with pytest.raises(ValueError):
        _glm().add_design_matrix(np.zeros((50, 4), np.float32), dims=[5, 3], df=[3, 3], smooth="cr")
    with pytest.raises(NotImplementedError):
        _glm().add_design_matrix(np.zeros(50, np.float32), dims=[5], df=[3], smooth="cr",
                                 filter_nonlinearity="swish")
    with pytest.raises(NotImplementedError):
        _glm(distr="binomial")
    with pytest.raises(ValueError):
        _glm().add_design_matrix(np.zeros(50, np.float32), dims=[5], df=[3], smooth="zz")


@pytest.mark.parametrize("n,n_pix,n_q", [(300, 64, 16), (1000, 300, 36), (4097, 1024, 100), (70001, 300, 36),
                                         (513, 128, 128), (260, 96, 8)])
def test_tensor_core_spatial_projection(n, n_pix, n_q):
    """K1a on tcgen05 (TMA-fed 3xTF32, TMEM accumulators) against fp64 and the CUDA-core kernel."""
    from rfest_b200.engine import Block, Engine

    eng = Engine.get()
    rng = np.random.default_rng(n_pix + n_q)
    stim = rng.standard_normal((n, n_pix)).astype(np.float32)
    Sxy = (rng.standard_normal((n_pix, n_q)) / np.sqrt(n_pix)).astype(np.float32)
    out = {}
    try:
        for tc in (0, 1):
            eng.set_option("project_tc", tc)
            blk = Block(eng, n, n_q)
            blk.project(stim, dst_row0=0, n_out=n, col0=0, src0=0, n_raw_total=n, n_pix=n_pix, first_t=0,
                        nlag=1, shift=0, St=np.ones((1, 1), np.float32), Sxy=Sxy)
            out[tc] = blk.read()
    finally:
        eng.set_option("project_tc", 1)
    ref = stim.astype(np.float64) @ Sxy.astype(np.float64)
    scale = np.abs(ref).max()
    assert np.abs(out[0] - ref).max() <= 2e-6 * scale
    assert np.abs(out[1] - ref).max() <= 2e-6 * scale          # fp32-level accuracy from tf32 tensor cores
    assert not np.array_equal(out[0], out[1])                   # i.e. the tensor-core path really ran


# ---- round 2: wide time bases, the host -> device staging ring, block @ matrix ----------------------------
@pytest.mark.parametrize("nlag,n_pix", [(20, 1), (37, 1), (24, 3), (70, 2)])
def test_filters_without_basis_wider_than_16_lags(nlag, n_pix):
    """a filter without a spline basis hands the kernel St = I[nlag]; more than 16 temporal
    columns used to be refused.  Integer input must still come back exactly (rfest/utils.py:50-67)."""
    rng = np.random.default_rng(nlag)
    shape = (333,) if n_pix == 1 else (333, n_pix)
    stim = rng.integers(-9, 10, size=shape).astype(np.float32)
    m = _glm(distr="gaussian")
    m.add_design_matrix(stim, dims=[nlag] if n_pix == 1 else [nlag, n_pix], smooth=None, shift=1, name="history")
    got = m._blocks["train"]["history"].read()
    want = design.trimmed_design(stim.reshape(333, -1), nlag, 1, m.burn_in, dtype=np.float32)
    assert got.shape == want.shape and np.array_equal(got, want)
    assert m.n_features["history"] == nlag * n_pix


def test_spline_basis_with_more_than_16_temporal_functions():
    rng = np.random.default_rng(3)
    stim = rng.standard_normal((400, 5)).astype(np.float32)
    m = _glm(distr="gaussian")
    m.add_design_matrix(stim, dims=[40, 5], df=[21, 3], smooth="cr", name="stimulus")
    ref = _oracle_xs(stim, [40, 5], [21, 3], 0, 39)
    got = np.asarray(m.XS["train"]["stimulus"])
    assert got.shape == ref.shape == (361, 63)
    assert np.abs(got - ref).max() <= 3e-6 * np.abs(ref).max()


@pytest.mark.parametrize("slab_rows", [4096, 5000, 1 << 15])
def test_host_ring_equals_device_input_bit_for_bit(slab_rows):
    """Host arrays travel through the 3-slot pinned ring (copy stream + events) in slabs; the result
    must not depend on the slab size, on pageable vs pinned source memory, or on host vs device input."""
    import torch

    from rfest_b200.engine import Engine, pinned_empty

    rng = np.random.default_rng(21)
    n = 23000
    stim = rng.standard_normal((n, 20, 15)).astype(np.float32)
    eng = Engine.get()
    outs = []
    try:
        for src in ("device", "pageable", "pinned", "serial"):
            eng.set_option("project_slab_rows", slab_rows)
            eng.set_option("h2d_overlap", 0 if src == "serial" else 1)
            if src == "device":
                x = torch.from_numpy(stim).cuda()
            elif src == "pinned":
                x = pinned_empty(stim.size).reshape(stim.shape)
                x[...] = stim
            else:
                x = stim
            m = _glm(distr="gaussian")
            m.add_design_matrix(x, dims=[25, 20, 15], df=[8, 6, 6], smooth="cr", name="stimulus")
            outs.append(np.asarray(m.XS["train"]["stimulus"]))
            if src != "device":
                assert eng.last_project_h2d_bytes >= stim.nbytes
    finally:
        eng.set_option("project_slab_rows", 0)
        eng.set_option("h2d_overlap", 1)
    for o in outs[1:]:
        assert np.array_equal(outs[0], o)
    ref = _oracle_xs(stim[:300], [25, 20, 15], [8, 6, 6], 0, 24)
    assert np.abs(outs[0][:276] - ref).max() <= 2e-6 * np.abs(ref).max()


@pytest.mark.parametrize("n,dims,df,shift,slab_rows", [
    (23000, [25, 20, 15], [8, 6, 6], 0, 0),          # BASELINE config 2-4 shape, one slab
    (23000, [25, 20, 15], [8, 6, 6], 0, 4100),       # several slabs: ranges, warm-up tiles and slab seams
    (9000, [25, 20, 15], [8, 6, 6], -3, 2500),       # negative shift: the window reaches into the future
    (9000, [25, 20, 15], [8, 6, 6], 2, 3000),        # positive shift: the first slab falls back, the rest fuse
    (700, [12, 8, 8], [4, 4, 4], 0, 0),              # 4 time-basis columns, 16 spatial ones, few tiles
    (40000, [64, 16, 16], [8, 8, 8], 0, 0),          # the longest lag window, 64 spatial columns
    (130, [25, 20, 15], [8, 6, 6], 0, 0),            # one and a bit tiles
])
def test_fused_projection_equals_the_two_kernel_path_bit_for_bit(n, dims, df, shift, slab_rows):
    """K1 fused (csrc/project_tc.cuh, FUSED: the temporal Hankel contraction runs inside the tensor-core kernel on
    a shared-memory ring of Z tiles) must reproduce the two-kernel path exactly -- same Z, same order of the lag
    sums -- whatever the slab size, tile ranges per CTA or shift; and both follow the oracle."""
    import torch

    from rfest_b200.engine import Engine

    rng = np.random.default_rng(n + shift)
    stim = rng.standard_normal((n, dims[1], dims[2])).astype(np.float32)
    x = torch.from_numpy(stim).cuda()
    eng = Engine.get()
    outs = {}
    try:
        eng.set_option("project_slab_rows", slab_rows)
        for fused in (1, 0):
            eng.set_option("project_fused", fused)
            m = _glm(distr="gaussian")
            launches = eng.launch_count
            m.add_design_matrix(x, dims=dims, df=df, smooth="cr", shift=shift, name="stimulus")
            outs[fused] = (np.asarray(m.XS["train"]["stimulus"]), eng.launch_count - launches)
    finally:
        eng.set_option("project_slab_rows", 0)
        eng.set_option("project_fused", 0)
    assert outs[1][1] < outs[0][1]                               # fewer kernels: the fused path really ran
    assert np.array_equal(outs[1][0], outs[0][0])
    k = min(n, 400)
    burn = dims[0] - 1                                           # the default burn-in (_glm.py:250-253)
    ref = _oracle_xs(stim[:k], dims, df, shift, burn)
    ok = len(ref) - max(0, -shift)                  # (rows whose window would reach past the truncated copy)
    assert ok > 0 and np.abs(outs[1][0][:ok] - ref[:ok]).max() <= 2e-6 * np.abs(ref).max()


def test_block_matmul_runs_on_the_device():
    rng = np.random.default_rng(2)
    stim = rng.standard_normal((5000, 6, 5)).astype(np.float32)
    m = _glm(distr="gaussian")
    m.add_design_matrix(stim, dims=[7, 6, 5], df=[4, 3, 3], smooth="cr", name="stimulus")
    blk = m.XS["train"]["stimulus"]
    XS = np.asarray(blk).astype(np.float64)
    launches = m.engine.launch_count
    for k in (1, 3, 11):
        B = rng.standard_normal((36, k)).astype(np.float32)
        got = blk @ B
        assert got.shape == (XS.shape[0], k) and got.dtype == np.float32
        ref = XS @ B.astype(np.float64)
        assert np.abs(got - ref).max() <= 2e-6 * np.abs(ref).max()
    v = blk @ B[:, 0]
    assert v.shape == (XS.shape[0],)
    assert m.engine.launch_count > launches                 # a kernel ran; nothing was multiplied on the host
    with pytest.raises(ValueError):
        blk @ np.zeros((35, 1), np.float32)


def test_refusals_are_pinned():
    """Unknown nonlinearities raise as the reference does (_glm.py:167-170); float64 is accepted with a warning."""
    stim = np.zeros((50, 3), np.float32)
    m = _glm(distr="poisson", output_nonlinearity="softmax")
    with pytest.raises(NotImplementedError):
        m.add_design_matrix(stim, dims=[4, 3], df=[3, 3], smooth="cr", filter_nonlinearity="swish")
    with pytest.warns(UserWarning, match="float32"):
        _glm(distr="gaussian", dtype=np.float64)
