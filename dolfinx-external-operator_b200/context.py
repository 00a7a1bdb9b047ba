"""Device context: one `Context` = one B200 = one compute stream (include/eo_b200.h).

Mirrors how the reference is deployed - one MPI rank per process, each owning its
cell partition (external_operator.py:368-370) - with one rank per GPU.
"""

from __future__ import annotations

import ctypes as C
import os
import weakref

import numpy as np

from . import _lib
from ._lib import EOError, Stats


def _ptr(a):
    """void* of a numpy array / DeviceArray / None."""
    if a is None:
        return None
    if isinstance(a, int):
        return a
    if isinstance(a, DeviceArray):
        return a.ptr
    if isinstance(a, np.ndarray):
        return a.ctypes.data
    if hasattr(a, "__cuda_array_interface__"):
        return a.__cuda_array_interface__["data"][0]
    raise TypeError(f"cannot take a pointer of {type(a)}")


class DeviceArray:
    """A typed view of device memory allocated through a Context.  An owning array returns its memory to the device
    when it is garbage collected (or earlier, by .free()); views made by .reshape() keep their owner alive.  Memory
    that outlives an explicitly closed context is left to the process."""

    def __init__(self, ctx: "Context", ptr: int, shape, dtype, owner: bool = True, base: "DeviceArray | None" = None):
        self.ctx = ctx
        self.ptr = ptr
        self.shape = tuple(int(s) for s in shape)
        self.dtype = np.dtype(dtype)
        self._owner = owner
        self._base = base  # keeps the owning array (and with it the allocation) alive
        self._release = None
        if owner and ptr:
            self._release = weakref.finalize(self, Context._dev_release, ctx._lib, ctx._finalizer, ctx._h, ptr)
            self._release.atexit = False

    @property
    def size(self) -> int:
        return int(np.prod(self.shape, dtype=np.int64)) if self.shape else 1

    @property
    def nbytes(self) -> int:
        return self.size * self.dtype.itemsize

    @property
    def __cuda_array_interface__(self):
        return {
            "shape": self.shape,
            "typestr": self.dtype.str,
            "data": (self.ptr or 0, False),
            "version": 3,
            "strides": None,
            "stream": int(self.ctx.stream or 0) or None,
        }

    def reshape(self, *shape):
        if len(shape) == 1 and isinstance(shape[0], (tuple, list)):
            shape = tuple(shape[0])
        shape = tuple(shape)
        if -1 in shape:
            known = int(np.prod([s for s in shape if s != -1], dtype=np.int64))
            shape = tuple(self.size // max(known, 1) if s == -1 else s for s in shape)
        if int(np.prod(shape, dtype=np.int64)) != self.size:
            raise ValueError(f"cannot reshape device array of size {self.size} into {shape}")
        return DeviceArray(self.ctx, self.ptr, shape, self.dtype, owner=False, base=self)

    def copy_from(self, src) -> "DeviceArray":
        if isinstance(src, np.ndarray):
            src = np.ascontiguousarray(src, dtype=self.dtype)
            if src.size != self.size:
                raise ValueError(f"size mismatch: device {self.size} vs host {src.size}")
        self.ctx.copy(self, src, self.nbytes)
        return self

    def to_host(self, out: np.ndarray | None = None) -> np.ndarray:
        if out is None:
            out = np.empty(self.shape, dtype=self.dtype)
        if out.nbytes != self.nbytes or not out.flags.c_contiguous:
            raise ValueError("to_host: `out` must be C-contiguous with matching size")
        self.ctx.copy(out, self, self.nbytes)
        return out

    def free(self):
        if self._owner and self.ptr:
            if self._release is not None:
                self._release.detach()
            self.ctx._dev_free(self.ptr)
            self.ptr = None


class Context:
    def __init__(self, device: int | None = None):
        self._lib = _lib.load()
        if device is None:
            device = int(os.environ.get("LOCAL_RANK", "0"))
        h = C.c_void_p()
        rc = self._lib.eo_create(int(device), C.byref(h))
        if rc != 0:
            raise EOError(rc, (self._lib.eo_last_error(None) or b"").decode())
        self._h = h
        self.device = int(device)
        self._finalizer = weakref.finalize(self, Context._destroy, self._lib, h)
        # not at interpreter exit: models / tabulators that still hold this context are torn down in arbitrary order
        # then, and their destroy calls must not find a freed context (process teardown releases the GPU anyway)
        self._finalizer.atexit = False

    @property
    def alive(self) -> bool:
        """False once `close()` has destroyed the native context (handles created on it must not be used any more)."""
        return self._finalizer.alive

    @staticmethod
    def _destroy(lib, h):
        lib.eo_destroy(h)

    @staticmethod
    def _dev_release(lib, ctx_finalizer, h, ptr):
        if ctx_finalizer.alive:
            lib.eo_dev_free(h, ptr)

    @staticmethod
    def _host_free(lib, ctx_finalizer, h, ptr):
        if ctx_finalizer.alive:  # a pinned array that outlives an explicitly closed context is left to the process
            lib.eo_host_free(h, ptr)

    def close(self):
        self._finalizer()

    # -------------------------------------------------------------- plumbing
    def check(self, rc: int):
        if rc != 0:
            raise EOError(rc, (self._lib.eo_last_error(self._h) or b"").decode())

    @property
    def handle(self):
        return self._h

    @property
    def lib(self):
        return self._lib

    @property
    def stream(self):
        return self._lib.eo_stream(self._h)

    def sync(self):
        self.check(self._lib.eo_sync(self._h))

    def set_chunk(self, n_qp: int):
        self.check(self._lib.eo_set_chunk(self._h, int(n_qp)))

    @property
    def launch_count(self) -> int:
        return int(self._lib.eo_launch_count(self._h))

    # -------------------------------------------------------------- memory
    def empty(self, shape, dtype=np.float64) -> DeviceArray:
        shape = (shape,) if np.isscalar(shape) else tuple(shape)
        nbytes = int(np.prod(shape, dtype=np.int64)) * np.dtype(dtype).itemsize
        p = C.c_void_p()
        self.check(self._lib.eo_dev_alloc(self._h, nbytes, C.byref(p)))
        return DeviceArray(self, p.value, shape, dtype)

    def zeros(self, shape, dtype=np.float64) -> DeviceArray:
        a = self.empty(shape, dtype)
        self.check(self._lib.eo_dev_memset(self._h, a.ptr, 0, a.nbytes))
        return a

    def to_device(self, host: np.ndarray, dtype=None) -> DeviceArray:
        host = np.ascontiguousarray(host, dtype=dtype or host.dtype)
        a = self.empty(host.shape, host.dtype)
        a.copy_from(host)
        return a

    def _dev_free(self, ptr: int):
        if self.alive:  # after close() the native context (and its allocations' stream) is gone
            self.check(self._lib.eo_dev_free(self._h, ptr))

    def copy(self, dst, src, nbytes: int):
        self.check(self._lib.eo_copy(self._h, _ptr(dst), _ptr(src), int(nbytes)))

    def pinned_empty(self, shape, dtype=np.float64) -> np.ndarray:
        """NumPy array in page-locked host memory (freed when the array dies)."""
        shape = (shape,) if np.isscalar(shape) else tuple(shape)
        dtype = np.dtype(dtype)
        nbytes = int(np.prod(shape, dtype=np.int64)) * dtype.itemsize
        if nbytes == 0:
            return np.empty(shape, dtype=dtype)
        p = C.c_void_p()
        self.check(self._lib.eo_host_alloc(self._h, nbytes, C.byref(p)))
        buf = (C.c_char * nbytes).from_address(p.value)
        arr = np.frombuffer(buf, dtype=dtype).reshape(shape)
        weakref.finalize(buf, Context._host_free, self._lib, self._finalizer, self._h, p.value)
        return arr

    def register(self, host: np.ndarray):
        """Page-lock an existing array in place (e.g. `ref_coefficient.x.array`)."""
        self.check(self._lib.eo_host_register(self._h, host.ctypes.data, host.nbytes))

    def unregister(self, host: np.ndarray):
        self.check(self._lib.eo_host_unregister(self._h, host.ctypes.data))

    # -------------------------------------------------------------- timing
    def event(self) -> int:
        e = C.c_void_p()
        self.check(self._lib.eo_event_create(self._h, C.byref(e)))
        return e.value

    def record(self, ev: int):
        self.check(self._lib.eo_event_record(self._h, ev))

    def elapsed_ms(self, e0: int, e1: int) -> float:
        ms = C.c_float()
        self.check(self._lib.eo_event_elapsed_ms(self._h, e0, e1, C.byref(ms)))
        return float(ms.value)

    def flush_l2(self, nbytes: int = 256 << 20):
        self.check(self._lib.eo_flush_l2(self._h, int(nbytes)))

    def fp64_peak_tflops(self, iters: int = 1 << 16) -> float:
        """Measured FP64 DFMA throughput of this GPU (the roofline denominator of the Newton-bound kernels)."""
        t = C.c_double()
        self.check(self._lib.eo_fp64_peak(self._h, int(iters), C.byref(t)))
        return float(t.value)

    def fp32_peak_tflops(self, iters: int = 1 << 16, variant: int | None = None) -> float:
        """Measured FP32 FMA throughput of this GPU (the roofline denominator of the Isihara network kernel).
        variant None / 0: scalar FFMA with uniform operands; 1: scalar FFMA, three register operands; 2: packed FFMA2."""
        t = C.c_double()
        if variant is None:
            self.check(self._lib.eo_fp32_peak(self._h, int(iters), C.byref(t)))
        else:
            self.check(self._lib.eo_fp32_peak_variant(self._h, int(iters), int(variant), C.byref(t)))
        return float(t.value)

    # -------------------------------------------------------------- statistics
    def stats_reset(self):
        self.check(self._lib.eo_stats_reset(self._h))

    def stats(self) -> dict:
        """The LOCAL record (this GPU's points since the last stats_reset)."""
        s = Stats()
        self.check(self._lib.eo_stats_read(self._h, C.byref(s)))
        return self._stats_dict(s)

    def stats_global(self) -> dict:
        """The GLOBAL record written by the last statistics collective (`parallel.allreduce_stats_device`,
        `eo_allreduce_stats`); waits for that collective."""
        s = Stats()
        self.check(self._lib.eo_stats_read_global(self._h, C.byref(s)))
        return self._stats_dict(s)

    @staticmethod
    def _stats_dict(s) -> dict:
        hist = np.array(s.niter_hist[:], dtype=np.int64)
        return {
            "n_points": int(s.n_points),
            "n_plastic": int(s.n_plastic),
            "n_nonconverged": int(s.n_nonconverged),
            "n_nonfinite": int(s.n_nonfinite),
            "niter_hist": hist,
            "niter_max": float(s.niter_max),
            "f_max": float(s.f_max),
            "res_max": float(s.res_max),
        }

    @property
    def stats_device_ptr(self) -> int:
        return self._lib.eo_stats_device_ptr(self._h)


_default_ctx: Context | None = None

# Counter of evaluate_operands / evaluate_external_operators rounds: callables that cache results across the requests
# of ONE round (HeatFlux: q, dq/dT, dq/dsigma from one launch) key their cache on it, so operand buffers that are
# refilled in place between rounds are never served stale.
_round = 0


def new_evaluation_round() -> int:
    global _round
    _round += 1
    return _round


def evaluation_round() -> int:
    return _round


def default_context() -> Context:
    """Process-wide context on cuda:$LOCAL_RANK (created on first use)."""
    global _default_ctx
    if _default_ctx is None:
        _default_ctx = Context()
    return _default_ctx
