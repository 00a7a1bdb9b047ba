"""`external_function` factories backed by the sm_100a kernels.

Each class is a drop-in for one of the reference demos' `*_external(derivatives)`
higher-order functions (protocol: external_operator.py:432 - the outer call
selects by derivative multi-index, the inner callable receives one array per
operand, shape `(n_cells, n_points, *operand_shape)`, and returns a flat array or
a tuple whose element 0 is that array, :435-438).

Differences from the reference callables, all behind the same protocol:
  * the arithmetic runs in `libeo_b200.so` (C ABI, include/eo_b200.h); the
    returned arrays live in page-locked host memory owned by the model object
    (the protocol says the callee keeps ownership - the caller copies, :289-290)
    or - `bind_outputs` - ARE the coefficient arrays, so no extra copy is made;
  * operands may be `DeviceArray`s produced by the GPU `evaluate_operands`, in
    which case nothing is uploaded;
  * history (`sigma_n`, `p`) may stay resident in HBM (`history="resident"`,
    committed on device by `commit()` = demo_vm:564-565) instead of being re-read
    from host `fem.Function`s at every call (`history=` two objects with
    `.x.array`, the closure style of demo_vm:347-348).
There is no CPU fallback: without the library / a B200 these raise `EOError`.
"""

from __future__ import annotations

import ctypes as C

import numpy as np

from ._lib import EO_LAYOUT_AOS, EO_LAYOUT_SOA, McParams, VmParams
from .context import Context, DeviceArray, _ptr, default_context, evaluation_round


def _n_points(arr) -> int:
    shape = arr.shape
    return int(np.prod(shape[:-1], dtype=np.int64)) if len(shape) > 1 else int(shape[0])


def _as_input(arr):
    """Borrow an operand without copying when it is already usable."""
    if isinstance(arr, DeviceArray):
        return arr
    if hasattr(arr, "materialize"):  # tabulation.LazyOperand handed to a callable that does not fuse the tabulation
        return arr.materialize()
    a = np.asarray(arr)
    if a.dtype != np.float64 or not a.flags.c_contiguous:
        a = np.ascontiguousarray(a, dtype=np.float64)
    return a


class _ModelBase:
    def __init__(self, ctx: Context | None):
        self.ctx = ctx or default_context()
        self._host_out: dict[str, np.ndarray] = {}
        self._bound: dict[str, np.ndarray] = {}
        self._last_stats = None

    def local_stats(self) -> dict:
        """Statistics record of this rank's last evaluation (points, plastic / non-converged / non-finite counts,
        Newton-iteration histogram, maxima) - what the reference prints per call, demo_mc:584-591."""
        if self._last_stats is None:
            raise RuntimeError("no evaluation yet")
        return self._last_stats

    def global_stats(self, group=None) -> dict:
        """The same record over ALL ranks of the torch.distributed `group` (default: the world group; one rank per
        GPU along the cell partition, external_operator.py:368-370): counts summed, maxima maximised - the one
        collective of the hot path (`parallel.allreduce_stats`: NCCL all-gather of the 1.7 KB record on the device,
        host tensors for gloo).  Collective call: every rank of the group must make it after its evaluation.
        Without an initialised process group it is the local record."""
        import torch.distributed as dist

        st = self.local_stats()
        if not (dist.is_available() and dist.is_initialized()):
            return st
        from .parallel import allreduce_stats

        return allreduce_stats(self.ctx, st, group)

    def _out(self, name: str, size: int) -> np.ndarray:
        """Result buffer `name` of `size` f64: a bound coefficient array, or pinned memory we own."""
        b = self._bound.get(name)
        if b is not None:
            if b.size != size:
                raise ValueError(f"bound output '{name}' has size {b.size}, the evaluation produces {size}")
            return b
        a = self._host_out.get(name)
        if a is None or a.size != size:
            a = self.ctx.pinned_empty(size)
            self._host_out[name] = a
        return a

    def bind_outputs(self, **arrays: np.ndarray):
        """Let results land directly in long-lived host arrays (e.g.
        `J_op.ref_coefficient.x.array`): they are page-locked in place and
        returned by the callable, so `evaluate_external_operators` skips the copy
        of external_operator.py:289-290."""
        for name, arr in arrays.items():
            if not (isinstance(arr, np.ndarray) and arr.dtype == np.float64 and arr.flags.c_contiguous):
                raise TypeError(f"bind_outputs({name}=...): need a C-contiguous float64 ndarray")
            self.ctx.register(arr)
            self._bound[name] = arr


# ---------------------------------------------------------------------------------------
class VonMises(_ModelBase):
    """von Mises radial return (plane strain, Mandel 4-vectors) with linear isotropic hardening.

    replaces: `return_mapping` + `C_tang_impl` + `sigma_external`,
    doc/demo/demo_plasticity_von_mises.py:298-368.  Parameters as :185-191.
    """

    def __init__(
        self,
        E: float = 70e3,
        nu: float = 0.3,
        E_tangent: float | None = None,
        sigma_0: float = 250.0,
        *,
        history="resident",
        n_qp: int | None = None,
        state_layout: str = "aos",
        ctx: Context | None = None,
    ):
        super().__init__(ctx)
        E_tangent = E / 100.0 if E_tangent is None else E_tangent
        self.E, self.nu, self.E_tangent, self.sigma_0 = E, nu, E_tangent, sigma_0
        self.H = E * E_tangent / (E - E_tangent)  # demo_vm:187
        self.lmbda = E * nu / (1.0 + nu) / (1.0 - 2.0 * nu)  # :190
        self.mu = E / 2.0 / (1.0 + nu)  # :191
        self._prm = VmParams(self.lmbda, self.mu, self.H, sigma_0)
        self.state_layout = {"aos": EO_LAYOUT_AOS, "soa": EO_LAYOUT_SOA}[state_layout]
        self._resident = isinstance(history, str) and history == "resident"
        self._history_objs = None if self._resident else tuple(history)  # (sigma_n, p) with .x.array
        self.n_qp = None
        self.sigma_n_dev = self.p_dev = self.sigma_dev = self.dp_dev = None
        if self._resident and n_qp is not None:
            self._alloc_state(int(n_qp))

    # ------------------------------------------------------------ state
    def _alloc_state(self, n: int):
        self.n_qp = n
        self.sigma_n_dev = self.ctx.zeros((n * 4,))
        self.p_dev = self.ctx.zeros((n,))
        self.sigma_dev = self.ctx.zeros((n * 4,))
        self.dp_dev = self.ctx.zeros((n,))

    def _to_state_layout(self, a: np.ndarray, ncomp: int) -> np.ndarray:
        a = np.asarray(a, dtype=np.float64).reshape(-1, ncomp)
        return np.ascontiguousarray(a.T if self.state_layout == EO_LAYOUT_SOA else a)

    def _from_state_layout(self, a: np.ndarray, ncomp: int) -> np.ndarray:
        if self.state_layout == EO_LAYOUT_SOA:
            return np.ascontiguousarray(a.reshape(ncomp, -1).T).reshape(-1)
        return a.reshape(-1)

    def set_history(self, sigma_n: np.ndarray, p: np.ndarray):
        """Upload history given in the reference's flat layout ([qp][4], [qp])."""
        n = int(np.asarray(p).size)
        if self.n_qp != n:
            self._alloc_state(n)
        self.sigma_n_dev.copy_from(self._to_state_layout(sigma_n, 4).reshape(-1))
        self.p_dev.copy_from(np.asarray(p, dtype=np.float64).reshape(-1))

    def get_history(self):
        """(sigma_n, p) in the reference's flat layout."""
        return self._from_state_layout(self.sigma_n_dev.to_host(), 4), self.p_dev.to_host()

    def commit(self):
        """End of a converged load step: p += dp ; sigma_n <- sigma  (demo_vm:564-565), on device."""
        if not self._resident:
            raise RuntimeError("commit() is for history='resident'; host history is updated by the caller")
        c = self.ctx
        c.check(c.lib.eo_commit_history(c.handle, self.sigma_n_dev.ptr, self.sigma_dev.ptr, self.p_dev.ptr,
                                        self.dp_dev.ptr, self.n_qp, 4))

    # ------------------------------------------------------------ callable protocol
    def __call__(self, derivatives):
        if derivatives == (1,):
            return self.C_tang_impl
        raise NotImplementedError(f"No external function is defined for the requested derivative {derivatives}.")

    def C_tang_impl(self, deps):
        """deps (n_cells, n_pts, 4) -> (C_tang.reshape(-1), sigma.reshape(-1), dp.reshape(-1)), demo_vm:343-352."""
        from .tabulation import LazyOperand

        if isinstance(deps, LazyOperand):
            if self._resident and self.state_layout == EO_LAYOUT_AOS and deps.kind_id == 2:
                return self._C_tang_fused(deps)
            deps = deps.materialize()
        deps = _as_input(deps)
        n = _n_points(deps) if len(deps.shape) > 1 else deps.shape[0] // 4
        c = self.ctx
        C_tang = self._out("C_tang", 16 * n)
        sigma = self._out("sigma", 4 * n)
        dp = self._out("dp", n)
        c.stats_reset()
        if self._resident:
            if self.n_qp is None:
                self._alloc_state(n)
            if self.n_qp != n:
                raise ValueError(f"operand has {n} quadrature points, the resident history {self.n_qp}")
            if self.state_layout == EO_LAYOUT_AOS:
                # tangent streams to the host through the chunk pipeline; the small
                # candidates stay in HBM for commit() and are copied out afterwards
                c.check(c.lib.eo_vm_eval(c.handle, C.byref(self._prm), _ptr(deps), self.sigma_n_dev.ptr,
                                         self.p_dev.ptr, _ptr(C_tang), self.sigma_dev.ptr, self.dp_dev.ptr, n))
                self.sigma_dev.to_host(sigma)
            else:
                d_deps = deps if isinstance(deps, DeviceArray) else c.to_device(deps.reshape(-1))
                d_Ct = c.empty((16 * n,))
                c.check(c.lib.eo_vm_eval_resident(c.handle, C.byref(self._prm), d_deps.ptr, self.sigma_n_dev.ptr,
                                                  self.p_dev.ptr, d_Ct.ptr, self.sigma_dev.ptr, self.dp_dev.ptr, n,
                                                  self.state_layout))
                d_Ct.to_host(C_tang)
                sigma[:] = self._from_state_layout(self.sigma_dev.to_host(), 4)
                d_Ct.free()
                if d_deps is not deps:
                    d_deps.free()
            self.dp_dev.to_host(dp)
        else:
            sigma_n_f, p_f = self._history_objs
            sn = _as_input(sigma_n_f.x.array)
            pp = _as_input(p_f.x.array)
            if sn.size != 4 * n or pp.size != n:
                raise ValueError("history arrays do not match the operand's quadrature-point count")
            c.check(c.lib.eo_vm_eval(c.handle, C.byref(self._prm), _ptr(deps), _ptr(sn), _ptr(pp), _ptr(C_tang),
                                     _ptr(sigma), _ptr(dp), n))
        self._last_stats = c.stats()  # synchronises
        return C_tang, sigma, dp

    def _C_tang_fused(self, lazy):
        """The operand is an un-tabulated Mandel strain: tabulation + radial return in one kernel (eo_tab_vm_fused,
        exact arithmetic: bit-identical to the two-step path), the strain never touches HBM."""
        n = lazy.tab.n_cells * lazy.tab.nq
        c = self.ctx
        d_Ct = getattr(self, "_fused_Ct", None)
        if d_Ct is None or d_Ct.size != 16 * n:
            d_Ct = self._fused_Ct = c.empty((16 * n,))
        c.stats_reset()
        lazy.tab.vm_fused(self, lazy.coefficient, C_tang=d_Ct, exact=True)
        C_tang, sigma, dp = self._out("C_tang", 16 * n), self._out("sigma", 4 * n), self._out("dp", n)
        d_Ct.to_host(C_tang)
        self.sigma_dev.to_host(sigma)
        self.dp_dev.to_host(dp)
        self._last_stats = c.stats()  # synchronises
        return C_tang, sigma, dp

    def eval_device(self, deps: DeviceArray, C_tang: DeviceArray):
        """All-device evaluation (asynchronous on the ctx stream): the tangent is written to
        `C_tang`, the candidates to the resident `sigma_dev`/`dp_dev`.  For device-side consumers."""
        n = deps.size // 4
        if self.n_qp is None:
            self._alloc_state(n)
        c = self.ctx
        c.check(c.lib.eo_vm_eval_resident(c.handle, C.byref(self._prm), deps.ptr, self.sigma_n_dev.ptr,
                                          self.p_dev.ptr, C_tang.ptr, self.sigma_dev.ptr, self.dp_dev.ptr, n,
                                          self.state_layout))


# ---------------------------------------------------------------------------------------
class MohrCoulomb(_ModelBase):
    """Non-associative Mohr-Coulomb plasticity with Abbo-Sloan rounding and apex smoothing, per-point
    local Newton solve, tangent = derivative of the Newton iteration (what `jax.jacfwd` through
    `lax.while_loop` returns).

    replaces: `return_mapping`, `dsigma_ddeps`, `dsigma_ddeps_vec`, `C_tang_impl`, `sigma_external`,
    doc/demo/demo_plasticity_mohr_coulomb.py:474-533, :555, :574-620.  Parameters as :110-116, :469.

    The callable returns `(C_tang.reshape(-1), sigma.reshape(-1))` exactly like :593.  The aux state of
    :533 is available afterwards as `.niter`, `.yielding`, `.norm_res`, `.dlambda`, and the per-call
    summary the reference prints (:584-591) as `.summary()` (computed on the device, one small read).
    """

    def __init__(
        self,
        E: float = 6778.0,
        nu: float = 0.25,
        c: float = 3.45,
        phi: float = 30 * np.pi / 180,
        psi: float = 30 * np.pi / 180,
        theta_T: float = 26 * np.pi / 180,
        a: float | None = None,
        tol: float = 1e-8,
        Nitermax: int = 200,
        *,
        history="resident",
        n_qp: int | None = None,
        aux: bool = True,
        verbose: bool = False,
        ctx: Context | None = None,
    ):
        super().__init__(ctx)
        a = 0.26 * c / np.tan(phi) if a is None else a  # demo_mc:116
        self.E, self.nu, self.c, self.phi, self.psi, self.theta_T, self.a = E, nu, c, phi, psi, theta_T, a
        self.tol, self.Nitermax = tol, int(Nitermax)
        self._prm = McParams(E, nu, c, phi, psi, theta_T, a, tol, int(Nitermax))
        self._resident = isinstance(history, str) and history == "resident"
        self._history_objs = None if self._resident else (history,)  # sigma_n with .x.array (demo_mc:579)
        self.want_aux = aux
        self.verbose = verbose
        self.n_qp = None
        self.sigma_n_dev = self.sigma_dev = None
        self.niter = self.yielding = self.norm_res = self.dlambda = None
        if self._resident and n_qp is not None:
            self._alloc_state(int(n_qp))

    def _alloc_state(self, n: int):
        self.n_qp = n
        self.sigma_n_dev = self.ctx.zeros((n * 4,))
        self.sigma_dev = self.ctx.zeros((n * 4,))

    def set_history(self, sigma_n: np.ndarray):
        sigma_n = np.ascontiguousarray(sigma_n, dtype=np.float64).reshape(-1)
        n = sigma_n.size // 4
        if self.n_qp != n:
            self._alloc_state(n)
        self.sigma_n_dev.copy_from(sigma_n)

    def get_history(self) -> np.ndarray:
        return self.sigma_n_dev.to_host()

    def commit(self):
        """End of a converged load step: sigma_n <- sigma (demo_mc:728), on device."""
        if not self._resident:
            raise RuntimeError("commit() is for history='resident'; host history is updated by the caller")
        c = self.ctx
        c.check(c.lib.eo_commit_history(c.handle, self.sigma_n_dev.ptr, self.sigma_dev.ptr, None, None, self.n_qp, 4))

    def __call__(self, derivatives):
        if derivatives == (1,):
            return self.C_tang_impl
        raise NotImplementedError(f"No external function is defined for the requested derivative {derivatives}.")

    def C_tang_impl(self, deps):
        """deps (n_cells, n_pts, 4) -> (C_tang.reshape(-1), sigma.reshape(-1)), demo_mc:577-593."""
        deps = _as_input(deps)
        n = _n_points(deps) if len(deps.shape) > 1 else deps.shape[0] // 4
        c = self.ctx
        C_tang = self._out("C_tang", 16 * n)
        sigma = self._out("sigma", 4 * n)
        if self.want_aux:
            self.niter = self._out_typed("niter", n, np.int32)
            self.yielding, self.norm_res, self.dlambda = (self._out(k, n) for k in ("yielding", "norm_res", "dlambda"))
            aux = [_ptr(self.niter), _ptr(self.yielding), _ptr(self.norm_res), _ptr(self.dlambda)]
        else:
            aux = [None, None, None, None]
        c.stats_reset()
        if self._resident:
            if self.n_qp is None:
                self._alloc_state(n)
            if self.n_qp != n:
                raise ValueError(f"operand has {n} quadrature points, the resident history {self.n_qp}")
            c.check(c.lib.eo_mc_eval(c.handle, C.byref(self._prm), _ptr(deps), self.sigma_n_dev.ptr, _ptr(C_tang),
                                     self.sigma_dev.ptr, *aux, n))
            self.sigma_dev.to_host(sigma)
        else:
            sn = _as_input(self._history_objs[0].x.array)
            if sn.size != 4 * n:
                raise ValueError("history array does not match the operand's quadrature-point count")
            c.check(c.lib.eo_mc_eval(c.handle, C.byref(self._prm), _ptr(deps), _ptr(sn), _ptr(C_tang), _ptr(sigma),
                                     *aux, n))
        c.sync()
        self._last_stats = c.stats()
        if self.verbose:
            print(self.summary_text())
        return C_tang, sigma

    def _out_typed(self, name: str, size: int, dtype) -> np.ndarray:
        a = self._host_out.get(name)
        if a is None or a.size != size or a.dtype != np.dtype(dtype):
            a = self.ctx.pinned_empty(size, dtype)
            self._host_out[name] = a
        return a

    def summary(self, group=None, *, reduce: bool | None = None) -> dict:
        """The inner-Newton summary of demo_mc:584-591 for the last call (from the device statistics
        record): unique iteration counts, their counts, max f(trial), max ||res||.  Under torch.distributed (one rank
        per GPU) pass `group` or `reduce=True` for the figures of ALL ranks - the scalar all-reduce that turns the
        reference's rank-local prints into global ones (collective call, see `global_stats`)."""
        if reduce is None:
            reduce = group is not None
        st = self.global_stats(group) if reduce else self.local_stats()
        it = np.nonzero(st["niter_hist"])[0]
        return {"unique_iters": it.astype(np.int32), "counts": st["niter_hist"][it], "max_f": st["f_max"],
                "max_residual": st["res_max"], "n_points": st["n_points"], "n_plastic": st["n_plastic"],
                "n_nonconverged": st["n_nonconverged"], "n_nonfinite": st["n_nonfinite"]}

    def summary_text(self) -> str:
        s = self.summary()
        return ("\tInner Newton summary:\n"
                f"\t\tUnique number of iterations: {s['unique_iters']}\n"
                f"\t\tCounts of unique number of iterations: {s['counts']}\n"
                f"\t\tMaximum f: {s['max_f']}\n"
                f"\t\tMaximum residual: {s['max_residual']}")

    def stress_update(self, deps: np.ndarray, sigma_n: np.ndarray) -> np.ndarray:
        """One stress update with explicit history (host arrays); used to walk stress paths."""
        deps = np.ascontiguousarray(deps, dtype=np.float64).reshape(-1, 4)
        sigma_n = np.ascontiguousarray(sigma_n, dtype=np.float64).reshape(-1, 4)
        n = deps.shape[0]
        Ct, sig = np.empty(16 * n), np.empty((n, 4))
        c = self.ctx
        c.check(c.lib.eo_mc_eval(c.handle, C.byref(self._prm), _ptr(deps), _ptr(sigma_n), _ptr(Ct), _ptr(sig),
                                 None, None, None, None, n))
        c.sync()
        return sig


# ---------------------------------------------------------------------------------------
class HeatConductivity(_ModelBase):
    """k(T) = 1/(A + B T) and its derivative.

    replaces: `k_impl`/`dkdT_impl`/`k_external`, demo_nonlinear_heat_equation_part1.py:247-303."""

    def __init__(self, A: float = 1.0, B: float = 1.0, *, ctx: Context | None = None):
        super().__init__(ctx)
        self.A, self.B = float(A), float(B)

    def __call__(self, derivatives):
        if derivatives == (0,):
            return self.k_impl
        elif derivatives == (1,):
            return self.dkdT_impl
        raise NotImplementedError(f"No external function is defined for the requested derivative {derivatives}.")

    def _run(self, T, which):
        T = _as_input(T)
        n = T.size
        out = self._out(which, n)
        c = self.ctx
        args = {"k": None, "dk": None}
        args[which] = _ptr(out)
        c.check(c.lib.eo_heat_eval(c.handle, self.A, self.B, _ptr(T), None, args["k"], args["dk"], None, None, None, n))
        c.sync()
        return out

    def k_impl(self, T):
        return self._run(T, "k")

    def dkdT_impl(self, T):
        return self._run(T, "dk")


class HeatFlux(_ModelBase):
    """q(T, sigma) = -k(T) sigma and its two derivatives (gdim = 2).

    replaces: `q_impl`/`dqdT_impl`/`dqdsigma_impl`/`q_external`, part2.py:209-284.
    `fused=True` evaluates all three in ONE kernel launch on the first request of a
    Newton iteration and serves the other two from the cached results as long as the
    operand arrays are the same objects (the reference evaluates k(T) three times)."""

    def __init__(self, A: float = 1.0, B: float = 1.0, *, fused: bool = True, ctx: Context | None = None):
        super().__init__(ctx)
        self.A, self.B = float(A), float(B)
        self.fused = fused
        self._cache_key = None

    def __call__(self, derivatives):
        if derivatives == (0, 0):
            return self.q_impl
        elif derivatives == (1, 0):
            return self.dqdT_impl
        elif derivatives == (0, 1):
            return self.dqdsigma_impl
        raise NotImplementedError(f"No external function is defined for the requested derivative {derivatives}.")

    def invalidate(self):
        self._cache_key = None

    def _run(self, T, sigma, want: str):
        # the cache holds strong references to the operand OBJECTS it was computed from, so a
        # later array cannot be confused with them by address reuse; in-place mutation of an
        # operand array between requests of ONE evaluate_external_operators round needs an explicit invalidate()
        # (a new round - context.new_evaluation_round - always recomputes).  The returned arrays are model-owned
        # buffers that the next evaluation overwrites; the reference copies them into the coefficient (:289-290).
        key = (T, sigma, evaluation_round())
        T = _as_input(T)
        sigma = _as_input(sigma)
        n = T.size
        if sigma.size != 2 * n:
            raise ValueError("sigma must hold 2 components per quadrature point")
        c = self.ctx
        if self.fused:
            if not (self._cache_key is not None and self._cache_key[0] is key[0] and self._cache_key[1] is key[1]
                    and self._cache_key[2] == key[2]):
                q, dT, ds = self._out("q", 2 * n), self._out("dqdT", 2 * n), self._out("dqdsigma", 4 * n)
                c.check(c.lib.eo_heat_eval(c.handle, self.A, self.B, _ptr(T), _ptr(sigma), None, None, _ptr(q),
                                           _ptr(dT), _ptr(ds), n))
                c.sync()
                self._cache_key = key
            return self._out(want, {"q": 2, "dqdT": 2, "dqdsigma": 4}[want] * n)
        out = self._out(want, {"q": 2, "dqdT": 2, "dqdsigma": 4}[want] * n)
        ptrs = {"q": None, "dqdT": None, "dqdsigma": None}
        ptrs[want] = _ptr(out)
        c.check(c.lib.eo_heat_eval(c.handle, self.A, self.B, _ptr(T), _ptr(sigma), None, None, ptrs["q"],
                                   ptrs["dqdT"], ptrs["dqdsigma"], n))
        c.sync()
        return out

    def eval_device(self, T: DeviceArray, sigma: DeviceArray, q: DeviceArray | None = None,
                    dqdT: DeviceArray | None = None, dqdsigma: DeviceArray | None = None):
        """All-device evaluation (asynchronous on the ctx stream) of any subset of q (2/pt), dq/dT (2/pt), dq/dsigma
        (4/pt, row-major 2x2) - for device-side consumers (`QuadratureForms.vector / .action / .matrix`)."""
        n = T.size
        if sigma.size != 2 * n:
            raise ValueError("sigma must hold 2 components per quadrature point")
        c = self.ctx
        c.check(c.lib.eo_heat_eval(c.handle, self.A, self.B, T.ptr, sigma.ptr, None, None, _ptr(q), _ptr(dqdT),
                                   _ptr(dqdsigma), n))

    def q_impl(self, T, sigma):
        return self._run(T, sigma, "q")

    def dqdT_impl(self, T, sigma):
        return self._run(T, sigma, "dqdT")

    def dqdsigma_impl(self, T, sigma):
        return self._run(T, sigma, "dqdsigma")
