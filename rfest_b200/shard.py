"""Time-sharding of a recording across GPU ranks (host-side integer logic).

Samples are independent given the lagged window, so rank g of G owns a contiguous block
of trimmed design rows; to build them it needs its raw rows plus a left halo of
`nlag + shift - 1` (or, for negative shifts, a right one) raw rows -- at build time only,
the fit iterations exchange nothing but the reduced gradient.
"""


def front_padding(nlag, shift):
    """Zero rows the reference puts in front of the input (rfest/utils.py:54-57)."""
    return nlag + shift - 1 if nlag + shift > 0 else 0


def shard_rows(n_rows, rank, world):
    """Contiguous, balanced split of `n_rows`: -> [lo, hi) owned by `rank`."""
    if world < 1 or not 0 <= rank < world:
        raise ValueError("bad rank/world")
    base, extra = divmod(int(n_rows), world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def source_rows(first_t, n_out, nlag, shift, n_raw_total):
    """Raw rows [lo, hi) that design rows [first_t, first_t + n_out) read (clipped)."""
    front = front_padding(nlag, shift)
    lo = max(first_t - front, 0)
    hi = min(first_t + n_out - 1 + nlag - 1 - front, n_raw_total - 1) + 1
    return (lo, hi) if hi > lo else (0, 0)
