"""GPU operand tabulation behind `evaluate_operands` (reference: external_operator.py:338-404).

`Tabulator` owns the device-resident mesh arrays and basis tables of ONE coefficient function on ONE mesh
(what the reference caches as a compiled `fem.Expression`, :386-399) and produces operand arrays of shape
`(n_entities, n_points, *operand_shape)` - as `DeviceArray`s that the GPU callables consume in place, or as
NumPy arrays.  Operands are declared explicitly (`register`): recognising arbitrary UFL is the job of the
reference's own `fem.Expression` path, which `evaluate_operands` still uses for every operand without a plan.

    tab = Tabulator.from_function_space(V, Q, coefficient=Du)       # needs dolfinx + basix
    tab = Tabulator(dofmap=..., x_dofmap=..., x=..., phi=..., dphi=..., bs=2, coefficient=Du)   # plain arrays
    tab.register(epsilon(Du), "mandel_strain")
    evaluated = evaluate_operands(F_external_operators, tabulator=tab)
"""

from __future__ import annotations

import ctypes as C

import numpy as np

from . import elements
from ._lib import (EO_OPERAND_DEF_GRAD, EO_OPERAND_GRAD, EO_OPERAND_MANDEL_STRAIN, EO_OPERAND_VALUE, TabDesc,
                   VmParams)
from .context import Context, DeviceArray, _ptr, default_context

KINDS = {"value": EO_OPERAND_VALUE, "grad": EO_OPERAND_GRAD, "mandel_strain": EO_OPERAND_MANDEL_STRAIN,
         "def_grad": EO_OPERAND_DEF_GRAD}


class Tabulator:
    def __init__(self, *, dofmap, x_dofmap, x, phi, dphi, bs: int = 1, n_dofs: int | None = None, dpsi=None,
                 coefficient=None, ctx: Context | None = None):
        self.ctx = ctx or default_context()
        self.dofmap = np.ascontiguousarray(dofmap, dtype=np.int32)
        self.x_dofmap = np.ascontiguousarray(x_dofmap, dtype=np.int32)
        x = np.asarray(x, dtype=np.float64)
        if x.shape[1] == 2:
            x = np.concatenate([x, np.zeros((x.shape[0], 1))], axis=1)
        self.x = np.ascontiguousarray(x)
        self.phi = np.ascontiguousarray(phi, dtype=np.float64)
        self.dphi = np.ascontiguousarray(dphi, dtype=np.float64)
        self.gdim = int(self.dphi.shape[0])
        self.nq, self.nb = (int(s) for s in self.phi.shape)
        self.bs = int(bs)
        self.n_cells = int(self.dofmap.shape[0])
        self.n_dofs = int(n_dofs) if n_dofs is not None else (int(self.dofmap.max()) + 1 if self.dofmap.size else 0)
        self.dpsi = np.ascontiguousarray(dpsi if dpsi is not None else elements.p1_geometry_derivatives(self.gdim),
                                         dtype=np.float64)
        if self.x_dofmap.shape != (self.n_cells, self.gdim + 1):
            raise ValueError("x_dofmap must be (n_cells, gdim + 1): affine simplex cells only")
        if self.dphi.shape != (self.gdim, self.nq, self.nb) or self.dofmap.shape[1] != self.nb:
            raise ValueError("table / dofmap shapes are inconsistent")
        d = TabDesc(self.gdim, self.bs, self.nb, self.nq, self.n_cells, self.n_dofs, self.x.shape[0],
                    self.dofmap.ctypes.data, self.x_dofmap.ctypes.data, self.x.ctypes.data, self.phi.ctypes.data,
                    self.dphi.ctypes.data, self.dpsi.ctypes.data)
        h = C.c_void_p()
        self.ctx.check(self.ctx.lib.eo_tab_create(self.ctx.handle, C.byref(d), C.byref(h)))
        self._h = h
        self.coefficient = coefficient  # object with .x.array (a fem.Function), a NumPy array or a DeviceArray
        self._plans: dict = {}

    # ------------------------------------------------------------------ construction from DOLFINx objects
    @classmethod
    def from_function_space(cls, V, Q, coefficient=None, ctx: Context | None = None):
        """Build from a DOLFINx function space `V` (the coefficient's) and the quadrature space `Q` of the
        operator (its interpolation points are the evaluation points, external_operator.py:200)."""
        mesh = V.mesh
        points = Q.element.interpolation_points
        tab = V.element.basix_element.tabulate(1, points)  # (1 + tdim, nq, nb, 1)
        tab = np.asarray(tab).reshape(tab.shape[0], tab.shape[1], -1)
        tdim = mesh.topology.dim
        return cls(dofmap=V.dofmap.list, x_dofmap=mesh.geometry.dofmap, x=mesh.geometry.x, phi=tab[0],
                   dphi=tab[1:1 + tdim], bs=V.dofmap.index_map_bs,
                   n_dofs=V.dofmap.index_map.size_local + V.dofmap.index_map.num_ghosts, coefficient=coefficient, ctx=ctx)

    def close(self):
        if self._h is not None and self.ctx.alive:
            self.ctx.lib.eo_tab_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ------------------------------------------------------------------ plans (evaluate_operands protocol)
    def register(self, operand, kind: str, coefficient=None, output: str = "device"):
        """Declare that UFL `operand` is `kind` of `coefficient` (default: the tabulator's).  output: 'device'
        (DeviceArray), 'host' (ndarray) or 'lazy' (a `LazyOperand`: tabulation fused into the consuming kernel)."""
        if kind not in KINDS:
            raise ValueError(f"unknown operand kind {kind!r}; known: {sorted(KINDS)}")
        self._plans[operand] = _Plan(self, KINDS[kind], coefficient, output)
        return self

    def plan_for(self, external_operator, operand):
        return self._plans.get(operand)

    # ------------------------------------------------------------------ evaluation
    def ncomp(self, kind) -> int:
        k = KINDS[kind] if isinstance(kind, str) else int(kind)
        n = self.ctx.lib.eo_tab_ncomp(self._h, k)
        if n < 0:
            raise ValueError(f"operand kind {kind!r} does not fit this element (gdim={self.gdim}, bs={self.bs})")
        return n

    def _coeff(self, coefficient):
        c = coefficient if coefficient is not None else self.coefficient
        if c is None:
            raise ValueError("no coefficient given")
        if isinstance(c, DeviceArray):
            if c.size != self.bs * self.n_dofs:
                raise ValueError("coefficient size does not match the dofmap")
            return c
        a = c.x.array if hasattr(c, "x") else c
        a = np.ascontiguousarray(a, dtype=np.float64).reshape(-1)
        if a.size != self.bs * self.n_dofs:
            raise ValueError(f"coefficient has {a.size} entries, the dofmap addresses {self.bs * self.n_dofs}")
        return a

    def _shape(self, kind_id: int, n_cells: int):
        if kind_id == EO_OPERAND_VALUE:
            tail = () if self.bs == 1 else (self.bs,)
        elif kind_id == EO_OPERAND_MANDEL_STRAIN:
            tail = (4,)
        elif kind_id == EO_OPERAND_GRAD and self.bs == 1:
            tail = (self.gdim,)
        else:
            tail = (self.bs, self.gdim)
        return (n_cells, self.nq) + tail

    def evaluate(self, kind, coefficient=None, entities=None, output: str = "device", out=None):
        """Operand of `kind` on `entities` (None = all local + ghost cells, :365-371; else an int32 array of
        cell indices).  output='device' returns a DeviceArray, 'host' a NumPy array."""
        kind_id = KINDS[kind] if isinstance(kind, str) else int(kind)
        ncomp = self.ncomp(kind_id)
        u = self._coeff(coefficient)
        cells = None
        n = self.n_cells
        if entities is not None:
            ent = np.asarray(entities)
            if ent.ndim != 1:
                raise NotImplementedError("(cell, local_facet) entities are not tabulated on the GPU; use the "
                                          "reference's fem.Expression path for codim-1 entities")
            cells = np.ascontiguousarray(ent, dtype=np.int32)
            n = int(cells.size)
        shape = self._shape(kind_id, n)
        c = self.ctx
        if out is None:
            out = c.empty(shape) if output == "device" else np.empty(shape)
        elif int(np.prod(out.shape)) != n * self.nq * ncomp:
            raise ValueError("`out` has the wrong size")
        c.check(c.lib.eo_tabulate(self._h, kind_id, _ptr(u), _ptr(cells), n, _ptr(out)))
        return out

    def vm_fused(self, vm, coefficient=None, C_tang: DeviceArray | None = None, strain: DeviceArray | None = None,
                 exact: bool = False):
        """Tabulate the Mandel strain and run the von Mises update in ONE kernel (history resident in `vm`).
        Returns the tangent DeviceArray; stress / dp candidates land in vm.sigma_dev / vm.dp_dev.
        exact=True reproduces `vm((1,))(tab.evaluate("mandel_strain"))` bit for bit; the default uses the
        cheaper downstream algebra (same flags, a few ulp in the values)."""
        n = self.n_cells * self.nq
        if vm.n_qp is None:
            vm._alloc_state(n)
        if vm.n_qp != n:
            raise ValueError(f"mesh has {n} quadrature points, the resident history {vm.n_qp}")
        c = self.ctx
        if C_tang is None:
            C_tang = c.empty((16 * n,))
        u = self._coeff(coefficient)
        prm = VmParams(vm.lmbda, vm.mu, vm.H, vm.sigma_0)
        c.check(c.lib.eo_tab_vm_fused(self._h, C.byref(prm), _ptr(u), vm.sigma_n_dev.ptr, vm.p_dev.ptr, C_tang.ptr,
                                      vm.sigma_dev.ptr, vm.dp_dev.ptr, _ptr(strain), int(bool(exact))))
        return C_tang


    def mc_fused(self, mc, coefficient=None, C_tang: DeviceArray | None = None, aux: dict | None = None):
        """Tabulate the Mandel strain and run the Mohr-Coulomb return mapping without ever storing the strain
        (`eo_mc_eval_tabulated`: the strain is evaluated inside pass 1 and kept only for the plastic points).
        History resident in `mc`; returns the tangent DeviceArray, the stress lands in `mc.sigma_dev`.  `aux`: optional
        dict of device arrays niter (int32) / yielding / norm_res / dlambda."""
        n = self.n_cells * self.nq
        if mc.n_qp is None:
            mc._alloc_state(n)
        if mc.n_qp != n:
            raise ValueError(f"mesh has {n} quadrature points, the resident history {mc.n_qp}")
        c = self.ctx
        if C_tang is None:
            C_tang = c.empty((16 * n,))
        u = self._coeff(coefficient)
        a = aux or {}
        c.check(c.lib.eo_mc_eval_tabulated(c.handle, C.byref(mc._prm), self._h, _ptr(u), mc.sigma_n_dev.ptr, C_tang.ptr,
                                           mc.sigma_dev.ptr, _ptr(a.get("niter")), _ptr(a.get("yielding")),
                                           _ptr(a.get("norm_res")), _ptr(a.get("dlambda"))))
        return C_tang


class LazyOperand:
    """An operand that has NOT been tabulated: (tabulator, kind, coefficient) for all cells.  `evaluate_operands`
    returns it for plans registered with output='lazy'; callables that can fuse the tabulation into their own kernel
    (`JitModel`: eo_jit_eval_tabulated; `VonMises`: eo_tab_vm_fused) consume it as is - the operand array is then
    never written to HBM - and every other consumer calls `materialize()`."""

    def __init__(self, tab: "Tabulator", kind_id: int, coefficient):
        self.tab, self.kind_id, self.coefficient = tab, kind_id, coefficient

    @property
    def shape(self):
        return self.tab._shape(self.kind_id, self.tab.n_cells)

    @property
    def size(self) -> int:
        return int(np.prod(self.shape))

    def materialize(self, output: str = "device"):
        return self.tab.evaluate(self.kind_id, self.coefficient, None, output)


class _Plan:
    def __init__(self, tab: Tabulator, kind_id: int, coefficient, output: str):
        self.tab, self.kind_id, self.coefficient, self.output = tab, kind_id, coefficient, output

    def evaluate(self, entities=None):
        if self.output == "lazy":
            if entities is None and type(self.tab) is Tabulator:
                return LazyOperand(self.tab, self.kind_id, self.coefficient)
            return self.tab.evaluate(self.kind_id, self.coefficient, entities, "device")
        return self.tab.evaluate(self.kind_id, self.coefficient, entities, self.output)


class GeneralTabulator(Tabulator):
    """Tabulation for everything beyond the affine-simplex fast path (`eo_gtab_*`, csrc/tabg.cu): any element given
    by its tables (quadrilaterals, hexahedra, higher degrees), non-affine geometry (`dgeo`: derivative tables of the
    geometry element, Jacobian per point) and codimension-1 entities: pass one table set per local facet
    (`phi.shape == (n_facets, nq, nb)`) and `entities` of shape (n, 2) = (cell, local facet), the form the reference
    uses for boundary operators (test_codim_external_operator.py:75-109).

        phi (n_sets, nq, nb) | (nq, nb);  dphi (n_sets, gdim, nq, nb) | (gdim, nq, nb);  dgeo likewise with ng columns
    """

    def __init__(self, *, dofmap, x_dofmap, x, phi, dphi, dgeo, bs: int = 1, n_dofs: int | None = None,
                 coefficient=None, ctx: Context | None = None):
        from ._lib import GTabDesc

        self.ctx = ctx or default_context()
        self.dofmap = np.ascontiguousarray(dofmap, dtype=np.int32)
        self.x_dofmap = np.ascontiguousarray(x_dofmap, dtype=np.int32)
        x = np.asarray(x, dtype=np.float64)
        if x.shape[1] == 2:
            x = np.concatenate([x, np.zeros((x.shape[0], 1))], axis=1)
        self.x = np.ascontiguousarray(x)
        phi, dphi, dgeo = (np.asarray(a, dtype=np.float64) for a in (phi, dphi, dgeo))
        if phi.ndim == 2:
            phi, dphi, dgeo = phi[None], dphi[None], dgeo[None]
        self.phi, self.dphi, self.dgeo = (np.ascontiguousarray(a) for a in (phi, dphi, dgeo))
        self.n_sets, self.nq, self.nb = (int(s) for s in self.phi.shape)
        self.gdim = int(self.dphi.shape[1])
        self.ng = int(self.dgeo.shape[3])
        self.bs = int(bs)
        self.n_cells = int(self.dofmap.shape[0])
        self.n_dofs = int(n_dofs) if n_dofs is not None else (int(self.dofmap.max()) + 1 if self.dofmap.size else 0)
        if self.dphi.shape != (self.n_sets, self.gdim, self.nq, self.nb) or self.dofmap.shape[1] != self.nb:
            raise ValueError("table / dofmap shapes are inconsistent")
        if self.dgeo.shape != (self.n_sets, self.gdim, self.nq, self.ng) or self.x_dofmap.shape != (self.n_cells, self.ng):
            raise ValueError("geometry table / x_dofmap shapes are inconsistent")
        d = GTabDesc(self.gdim, self.bs, self.nb, self.nq, self.ng, self.n_sets, self.n_cells, self.n_dofs, self.x.shape[0],
                     self.dofmap.ctypes.data, self.x_dofmap.ctypes.data, self.x.ctypes.data, self.phi.ctypes.data,
                     self.dphi.ctypes.data, self.dgeo.ctypes.data)
        h = C.c_void_p()
        self.ctx.check(self.ctx.lib.eo_gtab_create(self.ctx.handle, C.byref(d), C.byref(h)))
        self._h = h
        self.coefficient = coefficient
        self._plans = {}

    @classmethod
    def from_function_space(cls, V, Q, coefficient=None, facets: bool = False, ctx: Context | None = None):
        """From DOLFINx objects: `Q`'s interpolation points are the evaluation points - points of the cell, or
        (facets=True, Q on a codim-1 submesh) points of the reference facet, mapped onto every local facet with
        `elements.facet_points` (basix's sub-entity ordering)."""
        mesh = V.mesh
        tdim = mesh.topology.dim
        pts = np.asarray(Q.element.interpolation_points, dtype=np.float64)
        cell = mesh.topology.cell_name()
        Xs = elements.facet_points(cell, pts) if facets else pts[None]
        phi, dphi, dgeo = [], [], []
        for X in Xs:
            t = np.asarray(V.element.basix_element.tabulate(1, X))
            t = t.reshape(t.shape[0], t.shape[1], -1)
            g = np.asarray(mesh.geometry.cmap.tabulate(1, X))
            g = g.reshape(g.shape[0], g.shape[1], -1)
            phi.append(t[0]), dphi.append(t[1:1 + tdim]), dgeo.append(g[1:1 + tdim])
        return cls(dofmap=V.dofmap.list, x_dofmap=mesh.geometry.dofmap, x=mesh.geometry.x, phi=np.stack(phi),
                   dphi=np.stack(dphi), dgeo=np.stack(dgeo), bs=V.dofmap.index_map_bs,
                   n_dofs=V.dofmap.index_map.size_local + V.dofmap.index_map.num_ghosts, coefficient=coefficient, ctx=ctx)

    def close(self):
        if self._h is not None and self.ctx.alive:
            self.ctx.lib.eo_gtab_destroy(self._h)
            self._h = None

    def ncomp(self, kind) -> int:
        k = KINDS[kind] if isinstance(kind, str) else int(kind)
        n = self.ctx.lib.eo_gtab_ncomp(self._h, k)
        if n < 0:
            raise ValueError(f"operand kind {kind!r} does not fit this element (gdim={self.gdim}, bs={self.bs})")
        return n

    def evaluate(self, kind, coefficient=None, entities=None, output: str = "device", out=None):
        """Operand of `kind` on `entities`: None = all cells (:365-371), (n,) int cells, (n, 2) (cell, local facet)."""
        kind_id = KINDS[kind] if isinstance(kind, str) else int(kind)
        ncomp = self.ncomp(kind_id)
        u = self._coeff(coefficient)
        ent, width, n = None, 0, self.n_cells
        if entities is not None:
            ent = np.ascontiguousarray(entities, dtype=np.int32)
            if ent.ndim == 1:
                width = 1
            elif ent.ndim == 2 and ent.shape[1] == 2:
                width = 2
            else:
                raise ValueError("entities must have shape (n,) or (n, 2)")
            n = int(ent.shape[0])
        shape = self._shape(kind_id, n)
        c = self.ctx
        if out is None:
            out = c.empty(shape) if output == "device" else np.empty(shape)
        elif int(np.prod(out.shape)) != n * self.nq * ncomp:
            raise ValueError("`out` has the wrong size")
        c.check(c.lib.eo_gtab_tabulate(self._h, kind_id, _ptr(u), _ptr(ent), width, n, _ptr(out)))
        return out

    def vm_fused(self, *a, **k):
        raise NotImplementedError("the fused tabulate + von Mises kernel is the affine-triangle fast path (Tabulator)")
