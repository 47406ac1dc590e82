"""rfest_b200 -- B200-native engine for RFEst's spline-GLM fitting hot path.

    from rfest_b200 import GLM          # same surface as rfest.GLM (rfest/GLM/_glm.py:24)

Python host code over hand-written sm_100a CUDA kernels through a ctypes C ABI
(include/rfest_b200.h).  No JAX, no Triton, no CPU fallback: using the engine without
`librfest_b200.so` (build: `python -m rfest_b200.build`) or without a B200 raises.
"""

from .splines import build_spline_matrix, KronBasis  # noqa: F401
from .glm import GLM  # noqa: F401

__version__ = "0.1.0"
__all__ = ["GLM", "build_spline_matrix", "KronBasis"]
