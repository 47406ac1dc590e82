"""Oracle: time-lagged design matrix and dense spline projection.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  PINNED against the live
reference module `rfest/utils.py` and the golden vectors in `tests/golden/`.
"""

import numpy as np


def lagged_design(X, nlag, shift=0, dtype=np.float64):
    """Restates `build_design_matrix` (rfest/utils.py:5-67) for n_c == 1.

    Row t, lag block i (i = 0 .. nlag-1), pixel q:
        out[t, i * n_pix + q] = X[t + i - (nlag + shift - 1), q]   (0 outside).
    Zero rows are put in front when nlag + shift > 0 (utils.py:54-55) and
    `-shift` zero rows behind when shift < 0 (utils.py:59-60); the lag blocks
    are stacked side by side, oldest first (utils.py:62).
    """
    X = np.asarray(X)
    n = X.shape[0]
    flat = X.reshape(n, int(np.prod(X.shape[1:], dtype=np.int64)))
    n_pix = flat.shape[1]

    front = nlag + shift - 1 if nlag + shift > 0 else 0          # utils.py:54-57
    back = -shift if shift < 0 else 0                            # utils.py:59-60
    # NB: np.vstack with a float64 zero block promotes to float64 first; the
    # cast to `dtype` happens last (utils.py:67), so do the same here.
    padded = np.zeros((front + n + back, n_pix), dtype=np.result_type(flat.dtype, np.float64))
    padded[front:front + n] = flat

    out = np.empty((n, nlag * n_pix), dtype=padded.dtype)
    for i in range(nlag):                                        # utils.py:62
        out[:, i * n_pix:(i + 1) * n_pix] = padded[i:i + n]
    return out.astype(dtype)


def front_padding(nlag, shift):
    """Zero rows put in front of the input (rfest/utils.py:54-57).

    `nlag + shift - 1` when `nlag + shift > 0`; the reference pads nothing when
    `nlag + shift <= 0` (so very negative shifts saturate -- a quirk that is
    part of the contract).
    """
    return nlag + shift - 1 if nlag + shift > 0 else 0


def lag_source_row(t, i, nlag, shift):
    """Raw-input row feeding design row `t`, lag block `i` (may be out of range,
    which means zero).  The integer map the CUDA projection kernel must
    reproduce bit-exactly."""
    return t + i - front_padding(nlag, shift)


def trimmed_design(X, nlag, shift, burn_in, dtype=np.float64):
    """`build_design_matrix(...)[burn_in:]` as in `add_design_matrix`
    (rfest/GLM/_glm.py:255-258)."""
    return lagged_design(X, nlag, shift=shift, dtype=dtype)[burn_in:]


def project_dense(X_lag, S):
    """`XS = X_lag @ S` in the arrays' own dtype (rfest/GLM/_glm.py:297-298)."""
    return X_lag @ S
