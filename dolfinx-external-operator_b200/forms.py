"""Device-side consumers of the external-operator values (C ABI `eo_form_*`, csrc/form.cu).

In the reference, `evaluate_external_operators` copies stress and tangent into `ref_coefficient.x.array`
(external_operator.py:289-290) and DOLFINx then integrates them - `assemble_vector(b, F)` / `assemble_matrix(A, J)`
inside the SNES callbacks (petsc/petsc.py:60-64, 86-88; forms at demo_plasticity_von_mises.py:253, 390-398).  At
168 B per point that copy across PCIe is the end-to-end bottleneck once the kernels run at the HBM roofline
(DESIGN.md section 8).  `QuadratureForms` evaluates those two integrals on the device, where the values already
are, so only DOF vectors cross the link:

    forms = QuadratureForms(tab, weights)                    # tab: the Tabulator of the displacement space
    b = forms.vm_residual(vm, Du)                            # constitutive update + int sigma . eps(v) dx, one kernel
    y = forms.action("mandel_strain", "mandel_strain", forms.C_tang, x)     # J @ x for a Krylov method
    A = forms.matrix("mandel_strain", "mandel_strain", forms.C_tang)        # CSR values on the device

Boundary terms (the pressure load of demo_vm:253), Dirichlet lifting and the linear solve stay with the caller.
"""

from __future__ import annotations

import ctypes as C

import numpy as np

from ._lib import VmParams
from .context import DeviceArray, _ptr
from .tabulation import KINDS, Tabulator


class QuadratureForms:
    def __init__(self, tab: Tabulator, weights):
        if type(tab) is not Tabulator:
            raise TypeError("QuadratureForms needs the affine-simplex Tabulator (eo_tab) of the test/trial space")
        self.tab, self.ctx = tab, tab.ctx
        self.weights = np.ascontiguousarray(weights, dtype=np.float64).reshape(-1)
        if self.weights.size != tab.nq:
            raise ValueError(f"{self.weights.size} quadrature weights for {tab.nq} evaluation points per cell")
        h = C.c_void_p()
        self.ctx.check(self.ctx.lib.eo_form_create(tab._h, self.weights.ctypes.data, C.byref(h)))
        self._h = h
        self.n_scalar_dofs = tab.bs * tab.n_dofs
        self.C_tang: DeviceArray | None = None  # tangent of the last vm_residual call, resident
        self.row_ptr = self.col = None

    def close(self):
        if self._h is not None and self.ctx.alive:
            self.ctx.lib.eo_form_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ------------------------------------------------------------------ helpers
    def _kind(self, kind) -> int:
        k = KINDS[kind] if isinstance(kind, str) else int(kind)
        self.tab.ncomp(k)  # raises for kinds that do not fit the element
        return k

    def _vec_in(self, x):
        if isinstance(x, DeviceArray):
            if x.size != self.n_scalar_dofs:
                raise ValueError("vector size does not match the dofmap")
            return x
        a = x.x.array if hasattr(x, "x") else x
        a = np.ascontiguousarray(a, dtype=np.float64).reshape(-1)
        if a.size != self.n_scalar_dofs:
            raise ValueError(f"vector has {a.size} entries, the dofmap addresses {self.n_scalar_dofs}")
        return a

    def _vec_out(self, out, output):
        if out is None:
            return self.ctx.empty((self.n_scalar_dofs,)) if output == "device" else np.empty(self.n_scalar_dofs)
        if isinstance(out, np.ndarray) and (out.dtype != np.float64 or not out.flags.c_contiguous):
            raise ValueError("`out` must be a C-contiguous float64 array")
        if out.size != self.n_scalar_dofs:
            raise ValueError("`out` has the wrong size")
        return out

    def _points(self, a, ncomp, n_cells):
        if not isinstance(a, DeviceArray):
            raise TypeError("point values must be a DeviceArray (they are consumed where the operator kernels wrote them)")
        if a.size != self.tab.n_cells * self.tab.nq * ncomp:
            raise ValueError(f"point-value array has {a.size} entries, expected {self.tab.n_cells * self.tab.nq * ncomp}")
        return a

    def _cells(self, n_cells):
        n = self.tab.n_cells if n_cells is None else int(n_cells)
        if not 0 <= n <= self.tab.n_cells:
            raise ValueError("n_cells out of range")
        return n

    # ------------------------------------------------------------------ the integrals
    def vector(self, kind_test, coef: DeviceArray, out=None, n_cells=None, accumulate=False, output="host"):
        """b = assemble_vector(inner(coef, OP_test(v)) dx) over cells 0..n_cells-1 (default: all)."""
        k = self._kind(kind_test)
        coef = self._points(coef, self.tab.ncomp(k), n_cells)
        out = self._vec_out(out, output)
        c = self.ctx
        c.check(c.lib.eo_form_vector(self._h, k, coef.ptr, self._cells(n_cells), _ptr(out), int(bool(accumulate))))
        return out

    def action(self, kind_test, kind_trial, D: DeviceArray, x, out=None, n_cells=None, accumulate=False, output="host"):
        """y = A x with A = assemble_matrix(inner(D OP_trial(u_hat), OP_test(v)) dx), never formed."""
        kt, ki = self._kind(kind_test), self._kind(kind_trial)
        D = self._points(D, self.tab.ncomp(kt) * self.tab.ncomp(ki), n_cells)
        x = self._vec_in(x)
        out = self._vec_out(out, output)
        c = self.ctx
        c.check(c.lib.eo_form_action(self._h, kt, ki, D.ptr, _ptr(x), self._cells(n_cells), _ptr(out),
                                     int(bool(accumulate))))
        return out

    def vm_residual(self, vm, u=None, out=None, n_cells=None, accumulate=False, exact=False, output="host",
                    tangent="full"):
        """Constitutive update of the von Mises demo + residual in ONE kernel: the Mandel strain of `u` (default: the
        tabulator's coefficient) goes through the radial return (`vm`: a resident-history `VonMises`), tangent /
        stress / dp stay in HBM (`self.C_tang`, `vm.sigma_dev`, `vm.dp_dev`) and b = int sigma . eps(v) dx comes back.
        tangent="factored": the tangent is kept as 6 numbers per point (`self.T6`: v, cn, cd with
        C_t = C_elas - cn v v^T - cd dev) for `vm_action` - 48 instead of 128 bytes per point; `expand_tangent()` gives
        the reference's (n, 4, 4) array when an assembler needs it."""
        if tangent not in ("full", "factored"):
            raise ValueError("tangent must be 'full' or 'factored'")
        t = self.tab
        n = t.n_cells * t.nq
        if vm.n_qp is None:
            vm._alloc_state(n)
        if vm.n_qp != n:
            raise ValueError(f"mesh has {n} quadrature points, the resident history {vm.n_qp}")
        c = self.ctx
        u = t._coeff(u)
        out = self._vec_out(out, output)
        prm = VmParams(vm.lmbda, vm.mu, vm.H, vm.sigma_0)
        if tangent == "factored":
            if getattr(self, "T6", None) is None or self.T6.size != 6 * n:
                self.T6 = c.empty((6 * n,))
            self._T6_prm, self._T6_exact = prm, bool(exact)
            c.check(c.lib.eo_form_vm_step_factored(self._h, C.byref(prm), _ptr(u), vm.sigma_n_dev.ptr, vm.p_dev.ptr,
                                                   self.T6.ptr, vm.sigma_dev.ptr, vm.dp_dev.ptr, self._cells(n_cells),
                                                   _ptr(out), int(bool(accumulate)), int(bool(exact))))
            return out
        if self.C_tang is None or self.C_tang.size != 16 * n:
            self.C_tang = c.empty((16 * n,))
        c.check(c.lib.eo_form_vm_step(self._h, C.byref(prm), _ptr(u), vm.sigma_n_dev.ptr, vm.p_dev.ptr, self.C_tang.ptr,
                                      vm.sigma_dev.ptr, vm.dp_dev.ptr, self._cells(n_cells), _ptr(out),
                                      int(bool(accumulate)), int(bool(exact))))
        return out

    def vm_action(self, x, out=None, n_cells=None, accumulate=False, output="host"):
        """y = J x with the factored tangent of the last `vm_residual(..., tangent="factored")`: the matrix-free form of
        `assemble_matrix(J)` with J = inner(C_t eps(u_hat), eps(v)) dx (demo_plasticity_von_mises.py:390-398)."""
        if getattr(self, "T6", None) is None:
            raise RuntimeError("vm_action needs vm_residual(..., tangent='factored') first")
        x = self._vec_in(x)
        out = self._vec_out(out, output)
        c = self.ctx
        c.check(c.lib.eo_form_action_vm_factored(self._h, C.byref(self._T6_prm), self.T6.ptr, _ptr(x), self._cells(n_cells),
                                                 _ptr(out), int(bool(accumulate))))
        return out

    def expand_tangent(self, out: DeviceArray = None):
        """The factored tangent as the reference's (n, 4, 4) array on the device (bit-identical to what
        `vm_residual(tangent="full")` stores with the same `exact` flag)."""
        if getattr(self, "T6", None) is None:
            raise RuntimeError("expand_tangent needs vm_residual(..., tangent='factored') first")
        c = self.ctx
        n = self.T6.size // 6
        if out is None:
            if self.C_tang is None or self.C_tang.size != 16 * n:
                self.C_tang = c.empty((16 * n,))
            out = self.C_tang
        c.check(c.lib.eo_vm_expand_tangent(c.handle, C.byref(self._T6_prm), self.T6.ptr, out.ptr, n, int(self._T6_exact)))
        return out

    def mc_residual(self, mc, u=None, out=None, n_cells=None, accumulate=False, output="host", fused=False):
        """The Mohr-Coulomb counterpart of `vm_residual` (demo_plasticity_mohr_coulomb.py:679-688 + assemble_vector):
        Mandel strain of `u` -> local Newton return mapping (`mc`: a resident-history `MohrCoulomb`; tangent and stress
        stay in HBM as `self.C_tang`, `mc.sigma_dev`) -> b = int sigma . eps(v) dx.  Default: three launches
        (tabulation, the two-pass Mohr-Coulomb kernels, the stress integral - it needs both passes' stresses).
        fused=True tabulates the strain inside pass 1 (`eo_mc_eval_tabulated`: kept only for the plastic points, the
        32 B/point strain array of the mesh is never allocated); measured 2.85 ms per 2e7 points either way with the
        final kernels (3.30 against 3.07 earlier in round 2), so it is a memory, not a time saving."""
        t = self.tab
        n = t.n_cells * t.nq
        if mc.n_qp is None:
            mc._alloc_state(n)
        if mc.n_qp != n:
            raise ValueError(f"mesh has {n} quadrature points, the resident history {mc.n_qp}")
        c = self.ctx
        if self.C_tang is None or self.C_tang.size != 16 * n:
            self.C_tang = c.empty((16 * n,))
        c.stats_reset()
        if fused and type(t).__name__ == "Tabulator" and t.gdim == 2 and t.bs == 2 and t.nb in (3, 6, 10):
            # strain tabulated inside pass 1 of the Mohr-Coulomb kernels: never stored
            t.mc_fused(mc, u, C_tang=self.C_tang)
        else:
            if getattr(self, "_strain", None) is None or self._strain.size != 4 * n:
                self._strain = c.empty((t.n_cells, t.nq, 4))
            t.evaluate("mandel_strain", u, out=self._strain)
            c.check(c.lib.eo_mc_eval(c.handle, C.byref(mc._prm), self._strain.ptr, mc.sigma_n_dev.ptr, self.C_tang.ptr,
                                     mc.sigma_dev.ptr, None, None, None, None, n))
        return self.vector("mandel_strain", mc.sigma_dev, out=out, n_cells=n_cells, accumulate=accumulate, output=output)

    # ------------------------------------------------------------------ assembled matrix (CSR on the device)
    def set_pattern(self, row_ptr=None, col=None):
        """CSR pattern over scalar dofs; default: every pair of dofs sharing a cell (what DOLFINx's
        `create_sparsity_pattern` gives for a cell integral), built on the host once."""
        if row_ptr is None:
            row_ptr, col = cell_sparsity(self.tab.dofmap, self.tab.bs, self.tab.n_dofs)
        self.row_ptr = np.ascontiguousarray(row_ptr, dtype=np.int32)
        self.col = np.ascontiguousarray(col, dtype=np.int32)
        if self.row_ptr.size != self.n_scalar_dofs + 1:
            raise ValueError("row_ptr must have bs * n_dofs + 1 entries")
        c = self.ctx
        c.check(c.lib.eo_form_set_pattern(self._h, self.row_ptr.ctypes.data, self.col.ctypes.data, int(self.col.size)))
        return self.row_ptr, self.col

    def matrix(self, kind_test, kind_trial, D: DeviceArray, vals: DeviceArray | None = None, n_cells=None,
               accumulate=False):
        """CSR values (DeviceArray, nnz) of assemble_matrix(inner(D OP_trial(u_hat), OP_test(v)) dx)."""
        if self.row_ptr is None:
            self.set_pattern()
        kt, ki = self._kind(kind_test), self._kind(kind_trial)
        D = self._points(D, self.tab.ncomp(kt) * self.tab.ncomp(ki), n_cells)
        c = self.ctx
        if vals is None:
            vals = c.empty((int(self.col.size),))
        elif vals.size != self.col.size:
            raise ValueError("`vals` does not match the pattern")
        c.check(c.lib.eo_form_matrix(self._h, kt, ki, D.ptr, self._cells(n_cells), vals.ptr, int(bool(accumulate))))
        return vals


def cell_sparsity(dofmap: np.ndarray, bs: int, n_dofs: int):
    """(row_ptr, col) int32 of the scalar-dof CSR pattern in which two dofs are coupled iff they share a cell."""
    dofmap = np.asarray(dofmap)
    nc, nb = dofmap.shape
    nd = nb * bs
    sd = (bs * dofmap[:, :, None].astype(np.int64) + np.arange(bs)[None, None, :]).reshape(nc, nd)
    n = bs * n_dofs
    key = np.unique((sd[:, :, None] * n + sd[:, None, :]).reshape(-1))
    if key.size >= 2**31 - 1:
        raise ValueError("pattern too large for int32 CSR")
    rows, col = key // n, key % n
    row_ptr = np.zeros(n + 1, dtype=np.int64)
    np.add.at(row_ptr, rows + 1, 1)
    return np.cumsum(row_ptr).astype(np.int32), col.astype(np.int32)
