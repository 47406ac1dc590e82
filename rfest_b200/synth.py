"""Synthetic stimulus / response generators for benchmarks.

In the style of the reference's `rfest/simulate.py` (`noise2d` :201-244 -- i.i.d. N(0,1)
frames; `get_response` :291-323 -- Poisson spikes through a softplus), but generated ON THE
DEVICE in fixed blocks addressed by (seed, block index), so that any rank can produce
exactly its own rows (plus halo) and the data do not depend on the number of GPUs.
"""

import numpy as np

BLOCK = 1 << 16   # rows per generation block


class GaussianStimulus:
    """Row source of i.i.d. N(0,1) float32 frames living on a CUDA device."""

    def __init__(self, n_raw, frame_shape, seed=2046, device=0, chunk_rows=1 << 18):
        self.shape = (int(n_raw),) + tuple(int(d) for d in frame_shape)
        self.n_pix = int(np.prod(frame_shape)) if len(frame_shape) else 1
        self.seed, self.device, self.chunk_rows = int(seed), device, int(chunk_rows)

    def _block(self, j):
        import torch

        gen = torch.Generator(device=f"cuda:{self.device}")
        gen.manual_seed(self.seed * 1000003 + j)
        return torch.randn((BLOCK, self.n_pix), generator=gen, device=f"cuda:{self.device}",
                           dtype=torch.float32)

    def rows(self, lo, hi):
        import torch

        lo, hi = int(lo), int(hi)
        if hi <= lo:
            return torch.zeros((0, self.n_pix), device=f"cuda:{self.device}")
        parts = []
        for j in range(lo // BLOCK, (hi - 1) // BLOCK + 1):
            blk = self._block(j)
            a, b = max(lo - j * BLOCK, 0), min(hi - j * BLOCK, BLOCK)
            parts.append(blk[a:b])
        return parts[0].contiguous() if len(parts) == 1 else torch.cat(parts, 0)


class TensorRows:
    """Row source over a tensor/array that holds raw rows [row0, row0 + len) of a longer recording."""

    def __init__(self, data, row0, n_raw_total, chunk_rows=1 << 20):
        self.data, self.row0 = data, int(row0)
        self.shape = (int(n_raw_total),) + tuple(data.shape[1:])
        self.chunk_rows = chunk_rows

    def rows(self, lo, hi):
        a, b = lo - self.row0, hi - self.row0
        if a < 0 or b > self.data.shape[0]:
            raise ValueError(f"rows [{lo}, {hi}) requested, [{self.row0}, {self.row0 + self.data.shape[0]}) held")
        return self.data[a:b]


class RowWindow:
    """Rows [row0, row0 + n) of another row source as a recording of its own (train/dev splits)."""

    def __init__(self, source, row0, n):
        self.source, self.row0 = source, int(row0)
        self.shape = (int(n),) + tuple(source.shape[1:])
        self.chunk_rows = getattr(source, "chunk_rows", 1 << 18)
        self.n_pix = getattr(source, "n_pix", None)

    def rows(self, lo, hi):
        return self.source.rows(self.row0 + lo, self.row0 + hi)


def separable_filter(dims):
    """Ground-truth receptive field: biphasic temporal profile x Gaussian bump(s), unit norm
    (in the spirit of `gaussian3d`, rfest/simulate.py:19-22).  -> (kt [nlag], ks [n_pix] or None)."""
    nlag = dims[0]
    t = np.arange(nlag)
    kt = np.sin(2 * np.pi * t / nlag) * np.exp(-(nlag - 1 - t) / (nlag / 3))
    kt /= np.linalg.norm(kt)
    if len(dims) == 1:
        return kt.astype(np.float32), None
    ks = np.ones(1)
    for d in dims[1:]:
        g = np.exp(-0.5 * ((np.arange(d) - (d - 1) / 2) / (d / 6)) ** 2)
        ks = np.multiply.outer(ks, g).reshape(-1)
    ks /= np.linalg.norm(ks)
    return kt.astype(np.float32), ks.astype(np.float32)


def gaussian_response(engine, stim, dims, row_lo, row_hi, noise=0.1, seed=2046):
    """y[t] = x_t . w + N(0, (noise * std)^2)  (`get_response(distr='gaussian')`, rfest/simulate.py:315-316)."""
    return _response(engine, stim, dims, row_lo, row_hi, "gaussian", 1.0, 0.0, noise, seed)


def poisson_response(engine, stim, dims, row_lo, row_hi, gain=1.5, rate=0.3, seed=2046):
    """y[t] ~ Poisson(softplus(gain * (x_t . w) + c)) for raw rows [row_lo, row_hi), computed on
    the device with the engine's own projection kernel (w = kt x ks is rank one, so the drive
    is a one-column 'basis').  Whole generation blocks are drawn so that the spikes of a row
    do not depend on which rank asks for it.  -> CUDA float32 tensor of length row_hi - row_lo."""
    return _response(engine, stim, dims, row_lo, row_hi, "poisson", gain, rate, 0.0, seed)


def _response(engine, stim, dims, row_lo, row_hi, kind, gain, rate, noise, seed):
    import torch

    from .engine import Block
    from .shard import source_rows

    n_raw = stim.shape[0]
    kt, ks = separable_filter(dims)
    b_lo, b_hi = (row_lo // BLOCK) * BLOCK, min(((row_hi - 1) // BLOCK + 1) * BLOCK, n_raw)
    n = b_hi - b_lo
    drive = Block(engine, n, 1)
    step = stim.chunk_rows
    for k0 in range(0, n, step):
        kc = min(step, n - k0)
        s_lo, s_hi = source_rows(b_lo + k0, kc, dims[0], 0, n_raw)
        rows = stim.rows(s_lo, s_hi)
        drive.project(rows, dst_row0=k0, n_out=kc, col0=0, src0=s_lo, n_raw_total=n_raw,
                      n_pix=stim.n_pix, first_t=b_lo + k0, nlag=dims[0], shift=0,
                      St=kt[:, None], Sxy=None if ks is None else ks[:, None])
        del rows
    dev = f"cuda:{engine.device}"
    out = torch.empty(n, device=dev, dtype=torch.float32)
    engine.synchronize()
    flat = drive.as_torch()               # zero-copy view [n x 1] of the block (row stride 4)
    if kind == "poisson":
        u = gain * flat[:, 0] + float(np.log(np.expm1(rate)))
        lam = torch.nn.functional.softplus(u)
    else:
        lam = flat[:, 0].clone()        # unit-norm filter on N(0,1) frames: std of the drive is 1
    for j in range(b_lo // BLOCK, (b_hi - 1) // BLOCK + 1):
        gen = torch.Generator(device=dev)
        gen.manual_seed(seed * 7919 + 13 + j)
        a, b = j * BLOCK - b_lo, min((j + 1) * BLOCK, b_hi) - b_lo
        if kind == "poisson":
            out[a:b] = torch.poisson(lam[a:b], generator=gen)
        else:
            out[a:b] = lam[a:b] + noise * torch.randn(b - a, generator=gen, device=dev)
    return out[row_lo - b_lo: row_hi - b_lo].contiguous()
