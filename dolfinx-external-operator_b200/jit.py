"""Run-time compiled (NVRTC, sm_100a) external functions with forward-mode derivatives.

replaces: user-written `external_function(derivatives)` callables in general
(external_operator.py:432; README.md:16-25 "any array library ... automatic
differentiation").  Where the reference user writes a NumPy/JAX/PyTorch function
and lets the library differentiate it, the user here writes the per-quadrature-
point arithmetic once as a CUDA C++ function template

    template <class T>
    __device__ void model(const T* x, const double* state, const double* prm, T* y, T* aux);

(`x` = all operands concatenated, `state` = per-point history fields, `y` = the
operator's value, `aux` = extra per-point outputs) and `JitModel` serves every
derivative multi-index of total order <= 2 from it: T = double for the value,
`eo::dual<N>` / `eo::dual<N, eo::dual<M>>` (include/eo_dual.h) for derivatives.
Layout of a derivative: `[point][*value_shape][*operand_a_shape][*operand_b_shape]`,
the reference's `space_shape + operand_shape` convention (external_operator.py:117-121).

The callable protocol is the reference's: `model(derivatives)` returns a callable
taking one array per operand, shape `(n_cells, n_points, *operand_shape)`, and
returning a flat array - or, when `returns=` lists more fields, a tuple whose
element 0 is that array (:435-438), e.g. `("out", "value", "aux0")` to mirror
`return C_tang.reshape(-1), sigma.reshape(-1), dp.reshape(-1)` of demo_vm:352.
There is no CPU fallback: evaluation needs libeo_b200.so and a B200; compilation
alone (`compile_only=True`) works without a GPU.
"""

from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from ._lib import EO_JIT_MAX_ARGS, EO_JIT_MAX_PARAMS, EOError, JitDesc
from .context import Context, DeviceArray, _ptr, default_context


def _sizes(shapes):
    return [int(np.prod(s, dtype=np.int64)) if np.ndim(s) else int(s) for s in shapes]


class JitModel:
    """A per-quadrature-point model compiled on demand for each derivative multi-index.

    operand_shapes : value shape of every operand, e.g. `[(4,)]` or `[(), (2,)]`
    out_shape      : value shape of the operator
    state_shapes   : per-point history fields (resident in HBM: `state[i]` are DeviceArrays,
                     set with `set_state`), read-only inside the model
    aux_shapes     : extra per-point outputs (value part only)
    params         : up to 32 doubles, passed in the kernel's constant bank
    returns        : what the inner callable returns, names out | value | aux<i> | state<i>;
                     default: ("out",) -> a bare flat array
    output         : "host" - results in pinned host arrays owned by the model (the protocol's default);
                     "device" - results stay in HBM as `DeviceArray`s (consumed by
                     `evaluate_external_operators`, which downloads / gathers them into the coefficient)
    supported      : optional set of derivative tuples to accept; others raise NotImplementedError
                     like the reference demos do (demo_vm:364-368)
    """

    def __init__(self, source: str, entry: str, operand_shapes, out_shape=(), *, state_shapes=(), aux_shapes=(),
                 params=(), fmad: bool = True, returns=("out",), supported=None, output: str = "host",
                 ctx: Context | None = None, compile_only: bool = False):
        self.lib = _lib.load()
        self.ctx = None if compile_only else (ctx or default_context())
        self.operand_shapes = list(operand_shapes)
        self.operand_sizes = _sizes(operand_shapes)
        self.out_size = _sizes([out_shape])[0]
        self.state_sizes = _sizes(state_shapes)
        self.aux_sizes = _sizes(aux_shapes)
        self.params = np.ascontiguousarray(np.asarray(params, dtype=np.float64).reshape(-1))
        if max(len(self.operand_sizes), len(self.state_sizes), len(self.aux_sizes)) > EO_JIT_MAX_ARGS:
            raise ValueError(f"at most {EO_JIT_MAX_ARGS} operands / state fields / aux outputs")
        if self.params.size > EO_JIT_MAX_PARAMS:
            raise ValueError(f"at most {EO_JIT_MAX_PARAMS} parameters")
        self.returns = tuple(returns)
        if output not in ("host", "device"):
            raise ValueError("output must be 'host' or 'device'")
        self.output = output
        self._dev_out: dict = {}
        self.supported = None if supported is None else {tuple(d) for d in supported}
        self._src = source.encode()
        self._entry = entry.encode()
        d = JitDesc()
        d.source, d.entry = self._src, self._entry
        d.n_operands = len(self.operand_sizes)
        d.n_state = len(self.state_sizes)
        d.n_aux = len(self.aux_sizes)
        for i, s in enumerate(self.operand_sizes):
            d.operand_size[i] = s
        for i, s in enumerate(self.state_sizes):
            d.state_size[i] = s
        for i, s in enumerate(self.aux_sizes):
            d.aux_size[i] = s
        d.out_size = self.out_size
        d.n_params = self.params.size
        d.fmad = 1 if fmad else 0
        h = C.c_void_p()
        rc = self.lib.eo_jit_create(self.ctx.handle if self.ctx else None, C.byref(d), C.byref(h))
        if rc != 0:
            raise EOError(rc, (self.lib.eo_last_error(self.ctx.handle if self.ctx else None) or b"").decode())
        self._h = h
        self.state: list[DeviceArray | None] = [None] * len(self.state_sizes)
        self._host_out: dict = {}
        self.n_qp = None

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h and (self.ctx is None or self.ctx.alive):
            try:
                self.lib.eo_jit_destroy(h)
            except Exception:
                pass

    # ------------------------------------------------------------------ compilation
    def _check(self, rc: int):
        if rc != 0:
            text = (self.lib.eo_jit_last_error(self._h) or b"").decode()
            log = (self.lib.eo_jit_log(self._h) or b"").decode()
            raise EOError(rc, text + ("\n" + log if log and rc == -1 else ""))

    def _deriv(self, derivatives):
        derivatives = tuple(int(x) for x in derivatives)
        if len(derivatives) != len(self.operand_sizes):
            raise ValueError(f"derivative multi-index {derivatives} needs one entry per operand ({len(self.operand_sizes)})")
        return derivatives, (C.c_int * len(derivatives))(*derivatives)

    def compile(self, derivatives) -> int:
        """Compile the kernel for `derivatives` now; returns the CUBIN size in bytes."""
        _, d = self._deriv(derivatives)
        nbytes = C.c_size_t(0)
        self._check(self.lib.eo_jit_compile(self._h, d, C.byref(nbytes)))
        return int(nbytes.value)

    def cubin(self, derivatives) -> bytes:
        """The sm_100a CUBIN of the kernel for `derivatives` (cuobjdump -sass reads it)."""
        n = self.compile(derivatives)
        _, d = self._deriv(derivatives)
        buf = C.create_string_buffer(n)
        self._check(self.lib.eo_jit_cubin(self._h, d, buf, n))
        return buf.raw

    @property
    def log(self) -> str:
        return (self.lib.eo_jit_log(self._h) or b"").decode()

    def out_width(self, derivatives) -> int:
        _, d = self._deriv(derivatives)
        w = self.lib.eo_jit_out_width(self._h, d)
        if w < 0:
            self._check(w)
        return w

    # ------------------------------------------------------------------ state
    def set_state(self, i: int, values) -> None:
        """Upload history field i given in the reference's flat layout ([qp][components])."""
        v = np.ascontiguousarray(values, dtype=np.float64).reshape(-1)
        n = v.size // self.state_sizes[i]
        if self.state[i] is None or self.state[i].size != v.size:
            self.state[i] = self.ctx.empty((v.size,))
        self.state[i].copy_from(v)
        self.n_qp = n

    def get_state(self, i: int) -> np.ndarray:
        return self.state[i].to_host()

    # ------------------------------------------------------------------ callable protocol
    def __call__(self, derivatives):
        derivatives, d = self._deriv(derivatives)
        if self.supported is not None and derivatives not in self.supported:
            raise NotImplementedError(f"No external function is defined for the requested derivative {derivatives}.")
        if sum(derivatives) > 2:
            raise NotImplementedError(f"derivative {derivatives}: total order > 2 is not implemented")
        width = self.out_width(derivatives)

        def impl(*operands):
            return self._evaluate(derivatives, d, width, operands)

        impl.__name__ = f"{self._entry.decode()}_{''.join(map(str, derivatives))}"
        return impl

    def _buf(self, name: str, size: int) -> np.ndarray:
        """Pinned result buffer, one per (field, size): a Newton iteration alternates between derivative orders
        (different `out` widths) and must not re-allocate page-locked memory at every call."""
        a = self._host_out.get((name, size))
        if a is None:
            a = self.ctx.pinned_empty(size)
            self._host_out[(name, size)] = a
        return a

    def _evaluate(self, derivatives, d, width, operands, device_out: dict | None = None):
        if self.ctx is None:
            raise EOError(-5, "this JitModel was created with compile_only=True")
        if len(operands) != len(self.operand_sizes):
            raise TypeError(f"expected {len(self.operand_sizes)} operand arrays, got {len(operands)}")
        from .tabulation import LazyOperand

        lazy = [isinstance(a, LazyOperand) for a in operands]
        if all(lazy) and len({(a.tab.n_cells, a.tab.nq) for a in operands}) == 1:
            return self._evaluate_fused(derivatives, d, width, operands, device_out)
        if any(lazy):  # mixed: tabulate the lazy ones after all
            operands = [a.materialize() if z else a for a, z in zip(operands, lazy)]
        ins, n = [], None
        for a, s in zip(operands, self.operand_sizes):
            if not isinstance(a, DeviceArray):
                a = np.asarray(a)
                if a.dtype != np.float64 or not a.flags.c_contiguous:
                    a = np.ascontiguousarray(a, dtype=np.float64)
            m = a.size // s
            if m * s != a.size or (n is not None and m != n):
                raise ValueError("operand arrays disagree on the number of quadrature points")
            n = m
            ins.append(a)
        for i, st in enumerate(self.state):
            if st is None:
                raise ValueError(f"state field {i} has not been set (set_state)")
            if st.size != n * self.state_sizes[i]:
                raise ValueError(f"state field {i} holds {st.size // self.state_sizes[i]} points, the operands {n}")
        order = sum(derivatives)
        if n == 0:  # empty partition (a rank without cells): nothing to launch
            res = [np.empty(0) for _ in self.returns]
            return res[0] if len(res) == 1 else tuple(res)
        out_ptr, bufs = self._out_binder(device_out)
        p_out, p_val, p_aux = self._out_ptrs(out_ptr, order, width, n)
        ops = (C.c_void_p * len(ins))(*[_ptr(a) for a in ins])
        sts = (C.c_void_p * max(1, len(self.state)))(*[s.ptr for s in self.state])
        aux = (C.c_void_p * max(1, len(p_aux)))(*p_aux)
        prm = self.params.ctypes.data if self.params.size else None
        self._check(self.lib.eo_jit_eval(self._h, d, prm, ops, sts, p_out, p_val, aux, n))
        self.ctx.sync()
        return self._results(order, bufs)

    def _out_binder(self, device_out):
        """(out_ptr(name, size) -> void*, bufs): where each result lands - a caller-given DeviceArray, a
        model-owned DeviceArray (output='device') or a model-owned pinned host array."""
        device_out = device_out or {}
        bufs = {}

        def out_ptr(name, size):
            if name in device_out:
                bufs[name] = device_out[name]
                return device_out[name].ptr
            if self.output == "device" and (name == "out" or name in self.returns):
                a = self._dev_out.get((name, size))
                if a is None:
                    a = self._dev_out[(name, size)] = self.ctx.empty((size,))
                bufs[name] = a
                return a.ptr
            if name == "out" or name in self.returns:
                bufs[name] = self._buf(name, size)
                return _ptr(bufs[name])
            return None

        return out_ptr, bufs

    def _out_ptrs(self, out_ptr, order, width, n):
        p_out = out_ptr("out", width * n)
        p_val = out_ptr("value", self.out_size * n) if order >= 1 else None
        p_aux = [out_ptr(f"aux{i}", s * n) for i, s in enumerate(self.aux_sizes)]
        return p_out, p_val, p_aux

    def _results(self, order, bufs):
        res = []
        for name in self.returns:
            if name == "value" and order == 0:
                name = "out"
            if name.startswith("state"):
                res.append(self.state[int(name[5:])])
            else:
                res.append(bufs[name])
        return res[0] if len(res) == 1 else tuple(res)

    def _evaluate_fused(self, derivatives, d, width, operands, device_out=None):
        """All operands are `LazyOperand`s: tabulate them inside the model's kernel (eo_jit_eval_tabulated)."""
        n = operands[0].tab.n_cells * operands[0].tab.nq
        for a, s in zip(operands, self.operand_sizes):
            if a.tab.ncomp(a.kind_id) != s:
                raise ValueError("operand kind does not have the number of components the model expects")
        for i, st in enumerate(self.state):
            if st is None or st.size != n * self.state_sizes[i]:
                raise ValueError(f"state field {i} is not set or does not hold {n} points")
        order = sum(derivatives)
        if n == 0:
            res = [np.empty(0) for _ in self.returns]
            return res[0] if len(res) == 1 else tuple(res)
        out_ptr, bufs = self._out_binder(device_out)
        p_out, p_val, p_aux = self._out_ptrs(out_ptr, order, width, n)
        coeffs = [a.tab._coeff(a.coefficient) for a in operands]  # keep the arrays alive across the call
        tabs = (C.c_void_p * len(operands))(*[a.tab._h for a in operands])
        kinds = (C.c_int * len(operands))(*[a.kind_id for a in operands])
        us = (C.c_void_p * len(operands))(*[_ptr(u) for u in coeffs])
        sts = (C.c_void_p * max(1, len(self.state)))(*[s.ptr for s in self.state])
        aux = (C.c_void_p * max(1, len(p_aux)))(*p_aux)
        prm = self.params.ctypes.data if self.params.size else None
        self._check(self.lib.eo_jit_eval_tabulated(self._h, d, prm, tabs, kinds, us, sts, p_out, p_val, aux))
        self.ctx.sync()
        return self._results(order, bufs)

    def eval_device(self, derivatives, operands, out: DeviceArray, value: DeviceArray | None = None, aux=()):
        """All-device evaluation, asynchronous on the ctx stream (for device-side consumers and benchmarks)."""
        derivatives, d = self._deriv(derivatives)
        n = operands[0].size // self.operand_sizes[0]
        ops = (C.c_void_p * len(operands))(*[a.ptr for a in operands])
        sts = (C.c_void_p * max(1, len(self.state)))(*[s.ptr for s in self.state])
        pa = [a.ptr if a is not None else None for a in aux] + [None] * (len(self.aux_sizes) - len(aux))
        auxp = (C.c_void_p * max(1, len(pa)))(*pa)
        prm = self.params.ctypes.data if self.params.size else None
        self._check(self.lib.eo_jit_eval(self._h, d, prm, ops, sts, out.ptr, value.ptr if value is not None else None,
                                         auxp, n))
