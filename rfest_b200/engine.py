"""Python handles over the C ABI: engine (one per GPU), design blocks, device model."""

import ctypes as C
import os
import weakref

import numpy as np

from . import _lib
from ._lib import check


def _is_cuda_tensor(x):
    return hasattr(x, "data_ptr") and getattr(x, "is_cuda", False)


def as_f32(x):
    """-> (pointer, on_device, keepalive, shape) for a float32 C-contiguous array / CUDA tensor."""
    if _is_cuda_tensor(x):
        import torch

        t = x.detach()
        if t.dtype != torch.float32 or not t.is_contiguous():
            t = t.to(torch.float32).contiguous()
        # the engine runs on its own stream: whatever torch still has in flight for this
        # tensor must have landed before the pointer is handed over
        torch.cuda.current_stream(t.device).synchronize()
        return C.c_void_p(t.data_ptr()), 1, t, tuple(t.shape)
    a = np.ascontiguousarray(np.asarray(x), dtype=np.float32)
    return a.ctypes.data_as(C.c_void_p), 0, a, a.shape


class PinnedArray(np.ndarray):
    """ndarray over page-locked host memory owned by the engine library (freed with the array)."""

    _holder = None


def pinned_empty(n, dtype=np.float32):
    lib = _lib.load()
    dtype = np.dtype(dtype)
    ptr = C.c_void_p()
    check(lib.rfb_host_alloc(int(n) * dtype.itemsize, C.byref(ptr)))
    buf = (C.c_char * (int(n) * dtype.itemsize)).from_address(ptr.value)
    arr = np.frombuffer(buf, dtype=dtype, count=int(n)).view(PinnedArray)

    class _Holder:
        pass

    holder = _Holder()
    weakref.finalize(holder, lib.rfb_host_free, ptr)
    arr._holder = (holder, buf)
    return arr


class Engine:
    """One CUDA device + stream (+ NCCL communicator).  Use `Engine.get()`."""

    _instances = {}

    def __init__(self, device):
        self.lib = _lib.load()
        h = C.c_void_p()
        check(self.lib.rfb_engine_create(int(device), C.byref(h)))
        self.handle = h
        self.device = int(device)
        self.rank, self.world = 0, 1
        self._finalizer = weakref.finalize(self, self.lib.rfb_engine_destroy, h)

    @classmethod
    def get(cls, device=None):
        if device is None:
            device = int(os.environ.get("RFB_DEVICE", os.environ.get("LOCAL_RANK", "0")))
        if device not in cls._instances:
            cls._instances[device] = cls(device)
        return cls._instances[device]

    def synchronize(self):
        check(self.lib.rfb_engine_synchronize(self.handle))

    def set_option(self, key, value):
        """Sweep launch-shape knobs (see rfb_engine_set_option); affects models planned afterwards."""
        check(self.lib.rfb_engine_set_option(self.handle, key.encode(), int(value)))

    @property
    def launch_count(self):
        n = C.c_int64()
        check(self.lib.rfb_engine_launch_count(self.handle, C.byref(n)))
        return n.value

    @property
    def last_project_h2d_bytes(self):
        """Bytes the last `Block.project` call on this engine copied host -> device."""
        v = C.c_double()
        check(self.lib.rfb_engine_last_project_h2d_bytes(self.handle, C.byref(v)))
        return v.value

    @property
    def stream(self):
        s = C.c_uint64()
        check(self.lib.rfb_engine_stream(self.handle, C.byref(s)))
        return s.value

    def init_distributed(self):
        """Join the ranks of an initialised `torch.distributed` process group with NCCL.

        torch.distributed is only the bootstrap (it broadcasts the 128-byte NCCL id); the
        per-iteration all-reduce is issued by the engine itself inside its CUDA graph."""
        try:
            import torch.distributed as dist
        except ImportError:          # torch is only the multi-GPU bootstrap; single-GPU use needs NumPy alone
            return self

        if not dist.is_initialized() or dist.get_world_size() == 1 or self.world > 1:
            return self
        rank, world = dist.get_rank(), dist.get_world_size()
        uid = (C.c_char * 128)()
        if rank == 0:
            check(self.lib.rfb_comm_unique_id(uid))
        box = [bytes(uid)]
        dist.broadcast_object_list(box, src=0)
        buf = (C.c_char * 128).from_buffer_copy(box[0])
        check(self.lib.rfb_engine_comm_init(self.handle, buf, rank, world))
        self.rank, self.world = rank, world
        # one-shot NVLink exchange buffers (CUDA IPC); if the mapping fails the engine keeps NCCL
        if world <= 8 and os.environ.get("RFB_XCHG", "1") != "0":
            h = (C.c_char * 64)()
            check(self.lib.rfb_engine_xchg_create(self.handle, h))
            handles = [None] * world
            dist.all_gather_object(handles, bytes(h))
            blob = (C.c_char * (64 * world)).from_buffer_copy(b"".join(handles))
            rc = self.lib.rfb_engine_xchg_connect(self.handle, blob)
            ok = [None] * world
            dist.all_gather_object(ok, int(rc))
            if any(r != 0 for r in ok):          # all ranks must agree on the path they take
                self.set_option("xchg", 0)
        return self

    def allreduce_f64(self, values):
        a = np.ascontiguousarray(values, dtype=np.float64)
        check(self.lib.rfb_engine_allreduce_f64(self.handle, a.ctypes.data_as(_lib._dp), a.size))
        return a


class Block:
    """A projected design block XS_f resident in HBM: float32 (n_rows, n_cols)."""

    def __init__(self, engine, n_rows, n_cols):
        self.engine = engine
        self.lib = engine.lib
        h = C.c_void_p()
        check(self.lib.rfb_block_create(engine.handle, int(n_rows), int(n_cols), C.byref(h)))
        self.handle = h
        self.shape = (int(n_rows), int(n_cols))
        self.dtype = np.dtype(np.float32)
        self._finalizer = weakref.finalize(self, self.lib.rfb_block_destroy, h)

    @property
    def ld(self):
        ld = C.c_int64()
        check(self.lib.rfb_block_shape(self.handle, None, None, C.byref(ld)))
        return ld.value

    @property
    def device_ptr(self):
        p = C.c_uint64()
        check(self.lib.rfb_block_device_ptr(self.handle, C.byref(p)))
        return p.value

    def project(self, stim, *, dst_row0, n_out, col0, src0, n_raw_total, n_pix, first_t, nlag,
                shift, St, Sxy):
        """Fused lag + spline projection of raw rows into block rows (K1)."""
        ptr, on_dev, keep, shape = as_f32(stim)
        n_src = int(shape[0]) if len(shape) else 0
        St = np.ascontiguousarray(St, dtype=np.float32)
        df_t = St.shape[1]
        if Sxy is not None:
            Sxy = np.ascontiguousarray(Sxy, dtype=np.float32)
            n_q = Sxy.shape[1]
            sxy_ptr = Sxy.ctypes.data_as(C.c_void_p)
        else:
            n_q, sxy_ptr = 1, None
        check(self.lib.rfb_block_project(
            self.handle, int(dst_row0), int(n_out), int(col0), ptr, on_dev, int(src0), n_src,
            int(n_raw_total), int(n_pix), int(first_t), int(nlag), int(shift),
            St.ctypes.data_as(C.c_void_p), int(df_t), sxy_ptr, int(n_q)))
        del keep

    def write(self, row0, data):
        ptr, on_dev, keep, shape = as_f32(data)
        check(self.lib.rfb_block_write(self.handle, int(row0), int(shape[0]), ptr, on_dev))
        del keep

    def read(self, row0=0, n=None):
        n = self.shape[0] - row0 if n is None else n
        out = np.empty((n, self.shape[1]), dtype=np.float32)
        check(self.lib.rfb_block_read(self.handle, int(row0), int(n), out.ctypes.data_as(C.c_void_p)))
        return out

    def as_torch(self):
        """Zero-copy torch view (n_rows, n_cols) of the block's HBM storage (strided by ld)."""
        import torch

        class _View:
            pass

        v = _View()
        v.__cuda_array_interface__ = {
            "shape": self.shape, "typestr": "<f4", "data": (self.device_ptr, False), "version": 2,
            "strides": (self.ld * 4, 4)}
        t = torch.as_tensor(v, device=f"cuda:{self.engine.device}")
        t._rfb_keepalive = self
        return t

    # NumPy interoperability: np.asarray(model.XS['train']['stimulus']) copies the block back
    def __array__(self, dtype=None, copy=None):
        a = self.read()
        return a if dtype is None else a.astype(dtype)

    def __len__(self):
        return self.shape[0]

    def __matmul__(self, other):
        """`XS @ b` (rfest/GLM/_glm.py:562): computed on the device (rfb_block_matmul); only the
        (n_rows, k) product comes back, the block never leaves HBM."""
        B = np.asarray(other, dtype=np.float32)
        vec = B.ndim == 1
        B2 = np.ascontiguousarray(B.reshape(B.shape[0], -1))
        if B2.shape[0] != self.shape[1]:
            raise ValueError(f"matmul: block has {self.shape[1]} columns, operand {B2.shape[0]} rows")
        out = np.empty((self.shape[0], B2.shape[1]), dtype=np.float32)
        check(self.lib.rfb_block_matmul(self.handle, B2.ctypes.data_as(C.c_void_p), int(B2.shape[1]),
                                        out.ctypes.data_as(C.c_void_p)))
        return out[:, 0] if vec else out


class DeviceVector:
    """A float32 vector resident in HBM -- what `y_pred['opt'][kind]` holds after a fit (the
    reference keeps JAX device arrays there).  Behaves like an array where the GLM API needs it:
    `shape`, `len()`, indexing and `np.asarray()` copy it back (once), `as_torch()` is zero-copy."""

    def __init__(self, engine, device_ptr, n, padded_source=False):
        self.n = int(n)
        self.shape = (self.n,)
        self.dtype = np.dtype(np.float32)
        self._block = Block(engine, max(1, (self.n + 3) // 4), 4)   # contiguous n floats (+ padding)
        if device_ptr is not None and self.n > 0:
            # rfb_block_write moves whole 4-float rows: the source must hold the padded length
            if self.n % 4 and not padded_source:
                raise ValueError("DeviceVector: n is not a multiple of 4 and the source is not padded; "
                                 "pass padded_source=True if it holds (n + 3) // 4 * 4 floats")
            check(engine.lib.rfb_block_write(self._block.handle, 0, (self.n + 3) // 4,
                                             C.c_void_p(int(device_ptr)), 1))
        self._host = None

    @classmethod
    def empty(cls, engine, n):
        """Uninitialised vector for a kernel to fill through `device_ptr`."""
        return cls(engine, None, n)

    @classmethod
    def adopt(cls, block, n, recycle=None):
        """Vector over an existing [rows x 4] block (no allocation, no copy).  `recycle(block)` is called
        when the vector is garbage-collected, so the owner can reuse the storage."""
        v = cls.__new__(cls)
        v.n = int(n)
        v.shape = (v.n,)
        v.dtype = np.dtype(np.float32)
        v._block = block
        v._host = None
        if recycle is not None:
            weakref.finalize(v, recycle, block)
        return v

    @property
    def device_ptr(self):
        return self._block.device_ptr

    def __len__(self):
        return self.n

    def __array__(self, dtype=None, copy=None):
        if self._host is None:
            self._host = self._block.read().reshape(-1)[: self.n]
        return self._host if dtype is None else self._host.astype(dtype)

    def __getitem__(self, idx):
        return np.asarray(self)[idx]

    def as_torch(self):
        return self._block.as_torch().reshape(-1)[: self.n]

    # arithmetic falls back to the host copy
    def __sub__(self, other):
        return np.asarray(self) - np.asarray(other)

    def __rsub__(self, other):
        return np.asarray(other) - np.asarray(self)


class DeviceModel:
    """rfb_model handle: filters, bound data, parameters, the fit loop."""

    def __init__(self, engine, distr, output_nl):
        self.engine = engine
        self.lib = engine.lib
        h = C.c_void_p()
        check(self.lib.rfb_model_create(engine.handle, _lib.DISTR[distr], _lib.NL[output_nl],
                                        C.byref(h)))
        self.handle = h
        self._keep = {}
        self._spare = {}         # kind -> spare [rows x 4] blocks for the prediction hand-over
        self._finalizer = weakref.finalize(self, self.lib.rfb_model_destroy, h)

    def clear_filters(self):
        check(self.lib.rfb_model_clear_filters(self.handle))

    def add_filter(self, slot, nl, n_coef):
        return check(self.lib.rfb_model_add_filter(self.handle, int(slot), _lib.NL[nl], int(n_coef)))

    def set_data(self, kind, blocks, y, n):
        arr = (C.c_void_p * len(blocks))(*[b.handle.value if b is not None else None for b in blocks])
        if y is None:
            ptr, on_dev, keep = None, 0, None
        else:
            ptr, on_dev, keep, _ = as_f32(y)
        check(self.lib.rfb_model_set_data(self.handle, _lib.KIND[kind], len(blocks), arr, ptr,
                                          on_dev, int(n)))
        self._keep[kind] = list(blocks)      # blocks must outlive the binding
        del keep

    def set_y(self, kind, y, n):
        """Replace the response of an already bound data set (rfb_model_set_y)."""
        ptr, on_dev, keep, _ = as_f32(y)
        check(self.lib.rfb_model_set_y(self.handle, _lib.KIND[kind], ptr, on_dev, int(n)))
        del keep

    def loss_of_predictions(self, kind, r):
        """-> (loss_a, loss_b) of given predictions against the bound response (rfb_model_loss_of_predictions)."""
        if isinstance(r, DeviceVector):
            ptr, on_dev, keep = C.c_void_p(r.device_ptr), 1, r
        else:
            ptr, on_dev, keep, _ = as_f32(r)
        out = (C.c_double * 2)()
        check(self.lib.rfb_model_loss_of_predictions(self.handle, _lib.KIND[kind], ptr, on_dev, out))
        del keep
        return float(out[0]), float(out[1])

    def n_params(self):
        a, b = C.c_int(), C.c_int()
        check(self.lib.rfb_model_n_params(self.handle, C.byref(a), C.byref(b)))
        return a.value, b.value

    def set_params(self, b, icpt):
        b = np.ascontiguousarray(b, dtype=np.float32)
        icpt = np.ascontiguousarray(icpt, dtype=np.float64)
        check(self.lib.rfb_model_set_params(self.handle, b.ctypes.data_as(C.c_void_p),
                                            icpt.ctypes.data_as(C.c_void_p)))

    def eval(self, kind, b=None, icpt=None, want_r=False, n=None):
        sums = np.zeros(_lib.N_SUMS, dtype=np.float64)
        r = np.empty(n, dtype=np.float32) if want_r else None
        if b is not None:
            b = np.ascontiguousarray(b, dtype=np.float32)
            icpt = np.ascontiguousarray(icpt, dtype=np.float64)
        check(self.lib.rfb_model_eval(
            self.handle, _lib.KIND[kind],
            b.ctypes.data_as(C.c_void_p) if b is not None else None,
            icpt.ctypes.data_as(C.c_void_p) if b is not None else None,
            r.ctypes.data_as(C.c_void_p) if want_r else None,
            sums.ctypes.data_as(C.c_void_p)))
        return r, sums

    def value_and_grad(self, kind, b, icpt):
        """-> (sums, d loss/d b, d loss/d intercepts) of the data term at (b, icpt)."""
        P, K = self.n_params()
        b = np.ascontiguousarray(b, dtype=np.float32)
        icpt = np.ascontiguousarray(icpt, dtype=np.float64)
        sums = np.zeros(_lib.N_SUMS, dtype=np.float64)
        gb, gi = np.empty(P, dtype=np.float64), np.empty(K, dtype=np.float64)
        check(self.lib.rfb_model_value_and_grad(
            self.handle, _lib.KIND[kind], b.ctypes.data_as(C.c_void_p), icpt.ctypes.data_as(C.c_void_p),
            sums.ctypes.data_as(C.c_void_p), gb.ctypes.data_as(C.c_void_p), gi.ctypes.data_as(C.c_void_p)))
        return sums, gb, gi

    def gram(self, kind):
        """-> (G [ctot x ctot], colsum [ctot], xty [ctot]) in float64 over the padded, concatenated
        design columns (slot after slot); see rfb_model_gram."""
        c = C.c_int()
        check(self.lib.rfb_model_gram(self.handle, _lib.KIND[kind], None, None, None, C.byref(c)))
        G = np.empty((c.value, c.value), dtype=np.float64)
        s = np.empty(c.value, dtype=np.float64)
        t = np.empty(c.value, dtype=np.float64)
        check(self.lib.rfb_model_gram(self.handle, _lib.KIND[kind], G.ctypes.data_as(C.c_void_p),
                                      s.ctypes.data_as(C.c_void_p), t.ctypes.data_as(C.c_void_p), C.byref(c)))
        return G, s, t

    def filter_gram(self, kind, filter_index, n_coef, b=None, weighted=False):
        """XS_f^T diag(U) XS_f of one filter in float64 (see rfb_model_filter_gram)."""
        G = np.empty((n_coef, n_coef), dtype=np.float64)
        bb = np.ascontiguousarray(b, dtype=np.float32) if b is not None else None
        check(self.lib.rfb_model_filter_gram(
            self.handle, _lib.KIND[kind], int(filter_index),
            bb.ctypes.data_as(C.c_void_p) if bb is not None else None, int(bool(weighted)),
            G.ctypes.data_as(C.c_void_p)))
        return G

    def response_band(self, kind, b, intercepts, V_list, n):
        """-> (pred, upper, lower) DeviceVectors of this rank's n samples (rfb_model_response_band).
        V_list holds one float32 [n_coef x n_coef] covariance (or None) per filter."""
        bb = np.ascontiguousarray(b, dtype=np.float32)
        ic = np.ascontiguousarray(intercepts, dtype=np.float64)
        mats = [None if v is None else np.ascontiguousarray(v, dtype=np.float32) for v in V_list]
        ptrs = (C.c_void_p * len(mats))(*[None if v is None else v.ctypes.data for v in mats])
        out = [DeviceVector.empty(self.engine, n) for _ in range(3)]
        check(self.lib.rfb_model_response_band(
            self.handle, _lib.KIND[kind], bb.ctypes.data_as(C.c_void_p), ic.ctypes.data_as(C.c_void_p),
            ptrs, C.c_void_p(out[0].device_ptr), C.c_void_p(out[1].device_ptr),
            C.c_void_p(out[2].device_ptr), 1))
        return tuple(out)

    # ---- fit loop ------------------------------------------------------------------------
    def fit_begin(self, step_sizes, alpha, beta, metric):
        steps = np.ascontiguousarray(step_sizes, dtype=np.float64)
        check(self.lib.rfb_fit_begin(self.handle, steps.size, steps.ctypes.data_as(C.c_void_p),
                                     float(alpha), float(beta), _lib.METRIC[metric]))

    def fit_run(self, n):
        check(self.lib.rfb_fit_run(self.handle, int(n)))

    def fit_run_timed(self, n):
        """-> mean ms of (train sweep, dev sweep, reduction + all-reduce, update) over n iterations."""
        ms = np.zeros(4, dtype=np.float64)
        check(self.lib.rfb_fit_run_timed(self.handle, int(n), ms.ctypes.data_as(_lib._dp)))
        return ms

    @property
    def fit_is_persistent(self):
        """True when the fit in progress runs as the persistent small-model kernel (csrc/fit_small.cuh)."""
        v = C.c_int()
        check(self.lib.rfb_fit_is_persistent(self.handle, C.byref(v)))
        return bool(v.value)

    def fit_sweep_times(self, kind, i0, n):
        """-> ms of the `kind` sweep kernel in iterations [i0, i0 + n), stamped inside the replayed graph."""
        ms = np.zeros(n, dtype=np.float64)
        check(self.lib.rfb_fit_sweep_times(self.handle, _lib.KIND[kind], int(i0), int(n), ms.ctypes.data_as(_lib._dp)))
        return ms

    def fit_finish(self, i_last):
        check(self.lib.rfb_fit_finish(self.handle, int(i_last)))

    def fit_traces(self, n):
        out = [np.empty(n, dtype=np.float64) for _ in range(4)]
        check(self.lib.rfb_fit_traces(self.handle, int(n), *[o.ctypes.data_as(C.c_void_p) for o in out]))
        return out

    def fit_params_all(self, n):
        P, K = self.n_params()
        b = np.empty((n, P), dtype=np.float32)
        ic = np.empty((n, K), dtype=np.float64)
        check(self.lib.rfb_fit_params_all(self.handle, int(n), b.ctypes.data_as(C.c_void_p),
                                          ic.ctypes.data_as(C.c_void_p)))
        return b, ic

    def fit_predictions(self, kind, n):
        r = np.empty(n, dtype=np.float32)
        check(self.lib.rfb_fit_predictions(self.handle, _lib.KIND[kind], r.ctypes.data_as(C.c_void_p)))
        return r

    def fit_predictions_device(self, kind):
        """Predictions of the final forward pass as a DeviceVector (no host copy until asked).  The vector
        takes over the model's prediction buffer and the model gets a spare one in exchange
        (rfb_fit_predictions_swap); vectors that are garbage-collected return their storage to the
        spares, so repeated fits neither allocate nor copy."""
        ptr, n = C.c_uint64(), C.c_int64()
        check(self.lib.rfb_fit_predictions_ptr(self.handle, _lib.KIND[kind], C.byref(ptr), C.byref(n)))
        rows = (n.value + 4 + 3) // 4
        spares = self._spare.setdefault(kind, [])
        blk = None
        while spares and blk is None:
            cand = spares.pop()
            if cand.shape == (rows, 4):
                blk = cand
        if blk is None:
            blk = Block(self.engine, rows, 4)
        check(self.lib.rfb_fit_predictions_swap(self.handle, _lib.KIND[kind], blk.handle))
        me = weakref.ref(self)

        def recycle(block, me=me, kind=kind):
            dm = me()
            if dm is not None and len(dm._spare.setdefault(kind, [])) < 2:
                dm._spare[kind].append(block)

        return DeviceVector.adopt(blk, n.value, recycle)

    def fit_end(self):
        check(self.lib.rfb_fit_end(self.handle))
