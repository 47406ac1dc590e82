"""ctypes binding of `librfest_b200.so` (declarations mirror include/rfest_b200.h).

There is no CPU fallback: if the library is missing, `load()` raises.
"""

import ctypes as C
import os

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(PKG_DIR, "librfest_b200.so")

ABI_VERSION = 1

# constants of include/rfest_b200.h
OK = 0
ERR_INVALID, ERR_UNSUPPORTED, ERR_CUDA, ERR_NCCL, ERR_STATE = -1, -2, -3, -4, -5
NL = {"none": 0, "softplus": 1, "exponential": 2, "sigmoid": 3, "tanh": 4, "relu": 5,
      "leaky_relu": 6, "softmax": 7}
DISTR = {"gaussian": 0, "poisson": 1}
KIND = {"train": 0, "dev": 1, "test": 2}
METRIC = {None: 0, "mse": 1, "r2": 2, "corrcoef": 3}
SUM_LOSS_A, SUM_LOSS_B, SUM_R, SUM_RR, SUM_YR, SUM_SQERR, SUM_Y, SUM_YY = range(8)
N_SUMS = 8

_p = C.c_void_p
_i = C.c_int
_i64 = C.c_int64
_u64 = C.c_uint64
_d = C.c_double
_fp = C.POINTER(C.c_float)
_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)

# name -> (restype, argtypes); every function include/rfest_b200.h declares
SIGNATURES = {
    "rfb_last_error": (C.c_char_p, []),
    "rfb_abi_version": (_i, []),
    "rfb_engine_create": (_i, [_i, C.POINTER(_p)]),
    "rfb_engine_destroy": (_i, [_p]),
    "rfb_engine_synchronize": (_i, [_p]),
    "rfb_engine_launch_count": (_i, [_p, C.POINTER(_i64)]),
    "rfb_engine_set_option": (_i, [_p, C.c_char_p, _i]),
    "rfb_engine_stream": (_i, [_p, C.POINTER(_u64)]),
    "rfb_host_alloc": (_i, [C.c_size_t, C.POINTER(_p)]),
    "rfb_host_free": (_i, [_p]),
    "rfb_comm_unique_id": (_i, [_p]),
    "rfb_engine_comm_init": (_i, [_p, _p, _i, _i]),
    "rfb_engine_comm_info": (_i, [_p, _ip, _ip]),
    "rfb_engine_xchg_create": (_i, [_p, _p]),
    "rfb_engine_xchg_connect": (_i, [_p, _p]),
    "rfb_engine_allreduce_f64": (_i, [_p, _dp, _i]),
    "rfb_block_create": (_i, [_p, _i64, _i, C.POINTER(_p)]),
    "rfb_block_destroy": (_i, [_p]),
    "rfb_block_shape": (_i, [_p, C.POINTER(_i64), _ip, C.POINTER(_i64)]),
    "rfb_block_device_ptr": (_i, [_p, C.POINTER(_u64)]),
    "rfb_block_read": (_i, [_p, _i64, _i64, _p]),
    "rfb_block_write": (_i, [_p, _i64, _i64, _p, _i]),
    "rfb_block_project": (_i, [_p, _i64, _i64, _i, _p, _i, _i64, _i64, _i64, _i, _i64, _i, _i,
                                _p, _i, _p, _i]),
    "rfb_engine_last_project_h2d_bytes": (_i, [_p, _dp]),
    "rfb_block_matmul": (_i, [_p, _p, _i, _p]),
    "rfb_model_create": (_i, [_p, _i, _i, C.POINTER(_p)]),
    "rfb_model_destroy": (_i, [_p]),
    "rfb_model_clear_filters": (_i, [_p]),
    "rfb_model_add_filter": (_i, [_p, _i, _i, _i]),
    "rfb_model_set_data": (_i, [_p, _i, _i, C.POINTER(_p), _p, _i, _i64]),
    "rfb_model_set_params": (_i, [_p, _p, _p]),
    "rfb_model_get_params": (_i, [_p, _p, _p]),
    "rfb_model_n_params": (_i, [_p, _ip, _ip]),
    "rfb_model_eval": (_i, [_p, _i, _p, _p, _p, _p]),
    "rfb_model_value_and_grad": (_i, [_p, _i, _p, _p, _p, _p, _p]),
    "rfb_model_gram": (_i, [_p, _i, _p, _p, _p, _ip]),
    "rfb_model_filter_gram": (_i, [_p, _i, _i, _p, _i, _p]),
    "rfb_model_response_band": (_i, [_p, _i, _p, _p, _p, _p, _p, _p, _i]),
    "rfb_fit_begin": (_i, [_p, _i, _p, _d, _d, _i]),
    "rfb_fit_run": (_i, [_p, _i]),
    "rfb_fit_run_timed": (_i, [_p, _i, _dp]),
    "rfb_fit_is_persistent": (_i, [_p, _ip]),
    "rfb_fit_sweep_times": (_i, [_p, _i, _i, _i, _dp]),
    "rfb_fit_iterations": (_i, [_p, _ip]),
    "rfb_fit_finish": (_i, [_p, _i]),
    "rfb_fit_traces": (_i, [_p, _i, _p, _p, _p, _p]),
    "rfb_fit_params_at": (_i, [_p, _i, _p, _p]),
    "rfb_fit_params_all": (_i, [_p, _i, _p, _p]),
    "rfb_fit_predictions": (_i, [_p, _i, _p]),
    "rfb_fit_predictions_ptr": (_i, [_p, _i, C.POINTER(_u64), C.POINTER(_i64)]),
    "rfb_fit_predictions_swap": (_i, [_p, _i, _p]),
    "rfb_model_set_y": (_i, [_p, _i, _p, _i, _i64]),
    "rfb_model_loss_of_predictions": (_i, [_p, _i, _p, _i, C.POINTER(C.c_double)]),
    "rfb_fit_end": (_i, [_p]),
}

_lib = None


class RfbError(RuntimeError):
    def __init__(self, code, message):
        super().__init__(message)
        self.code = code


def load():
    """Load the shared library (once).  Raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -m rfest_b200.build` "
            "(rfest_b200 has no CPU fallback)")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    if lib.rfb_abi_version() != ABI_VERSION:
        raise ImportError(f"librfest_b200.so has ABI {lib.rfb_abi_version()}, expected {ABI_VERSION}")
    _lib = lib
    return lib


_EXC = {ERR_INVALID: ValueError, ERR_UNSUPPORTED: NotImplementedError}


def check(rc):
    """Map a status code to the exception type the reference raises for that situation."""
    if rc >= 0:
        return rc
    msg = load().rfb_last_error().decode("utf-8", "replace")
    exc = _EXC.get(rc)
    if exc is not None:
        raise exc(msg)
    raise RfbError(rc, msg)
