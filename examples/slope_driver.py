"""The Mohr-Coulomb demo problem of the reference, driven end to end through the device-side consumers.

replaces (as a self-contained driver, no DOLFINx / PETSc): doc/demo/demo_plasticity_mohr_coulomb.py:110-127 (soil
parameters, L x H = 1.2 x 1.0 rectangle), :129-145 (bottom and right boundary clamped), :617-625 (residual
inner(sigma, eps(v)) dx - inner(q, v) dx with the self-weight q = (0, -gamma)), :679-688 (constitutive update),
:708-731 (load steps in gamma up to the plateau, Du carried from step to step, sigma_n <- sigma).

Known answer (demo_mc:743-770): the slope stability factor l_lim = gamma_lim H / c = 6.69 of limit analysis (Chen) for
phi = 30 degrees - the self-weight at which the load-displacement curve of the crest point (0, H) reaches its plateau.
The demo's own 25 x 25 P2 mesh still converges at gamma = 22.99 (l = 6.66) with the displacement running away.

`backend` as in `thick_walled_cylinder`: residual(Du) -> int sigma(Du) . eps(v) dx (updates the tangent), tangent_csr(),
commit(), plus body_force() -> int (0, -1) . v dx.  `GpuBackend` is the product path (`QuadratureForms.mc_residual`,
`.matrix`, `.vector`); sparse LU (SciPy) and the Dirichlet elimination stay on the host like PETSc in the reference.
"""

from __future__ import annotations

import numpy as np

from dolfinx_external_operator_b200 import elements as el
from dolfinx_external_operator_b200 import synthetic as syn

L, H = 1.2, 1.0  # demo_mc:119
C_COHESION = 3.45  # :112
L_LIM = 6.69  # :764
GAMMA_LIM = L_LIM / H * C_COHESION  # :765


def slope_mesh(nx: int = 25, ny: int = 25):
    """P2 triangulation of the L x H rectangle (:119-127) with the demo's boundary data: `fixed` = all displacement
    dofs on the bottom (y = 0) and right (x = L) boundaries (:129-145), `probe` = u_x at the crest (0, H) (:714)."""
    m = syn.triangle_mesh(nx, ny, 2, lx=L, ly=H)
    xy = m["dof_coords"]
    clamped = np.nonzero(np.isclose(xy[:, 1], 0.0) | np.isclose(xy[:, 0], L))[0]
    m["fixed"] = np.sort(np.concatenate([2 * clamped, 2 * clamped + 1]))
    m["probe"] = 2 * int(np.nonzero(np.isclose(xy[:, 0], 0.0) & np.isclose(xy[:, 1], H))[0][0])
    X = el.triangle_quadrature(2)
    m["phi"], m["dphi"] = el.lagrange_triangle(2, X)
    m["weights"] = el.triangle_quadrature_weights(2)
    return m


class GpuBackend:
    """The product path: Tabulator + MohrCoulomb (history resident) + QuadratureForms on one B200."""

    def __init__(self, mesh, ctx=None):
        from dolfinx_external_operator_b200 import MohrCoulomb, QuadratureForms, Tabulator

        self.tab = Tabulator(dofmap=mesh["dofmap"], x_dofmap=mesh["x_dofmap"], x=mesh["x"], phi=mesh["phi"],
                             dphi=mesh["dphi"], bs=2, n_dofs=mesh["n_dofs"], ctx=ctx)
        self.ctx = self.tab.ctx
        self.forms = QuadratureForms(self.tab, mesh["weights"])
        self.nq = 3 * mesh["dofmap"].shape[0]
        self.mc = MohrCoulomb(n_qp=self.nq, aux=False, ctx=self.ctx)
        self.pattern = self.forms.set_pattern()
        self._vals = None

    def body_force(self):
        g = np.zeros((self.nq, 2))
        g[:, 1] = -1.0
        return self.forms.vector("value", self.ctx.to_device(g))

    def residual(self, Du):
        return self.forms.mc_residual(self.mc, Du)

    def tangent_csr(self):
        self._vals = self.forms.matrix("mandel_strain", "mandel_strain", self.forms.C_tang, vals=self._vals)
        return self._vals.to_host()

    def plastic_fraction(self) -> float:
        st = self.ctx.stats()
        return st["n_plastic"] / max(st["n_points"], 1)

    def commit(self):
        self.mc.commit()


def solve(mesh, backend, load_steps=None, rtol: float = 1e-8, atol: float = 1e-8, max_it: int = 100,
          verbose: bool = False):
    """Load stepping of demo_mc:708-731.  Stops (without raising) at the first load step whose Newton iteration fails:
    that is the collapse.  Returns dict(load, u_probe (= -u_x at the crest, :731), newton_iterations, plastic_fraction,
    n_converged)."""
    import scipy.sparse as sp
    from scipy.sparse.linalg import splu

    if load_steps is None:
        load_steps = np.concatenate([np.linspace(2, 22.9, 50), [22.96, 22.99]])  # :708-710
    load_steps = np.asarray(load_steps, dtype=np.float64)
    n = 2 * mesh["n_dofs"]
    free = np.setdiff1d(np.arange(n), mesh["fixed"])
    row_ptr, col = backend.pattern
    f_unit = backend.body_force()  # int (0, -1) . v dx ; the load of step k is load_steps[k] * gamma * f_unit (:718)
    u = np.zeros(n)
    # the demo starts from Du = 1 (:645), a constant whose strain is rounding residue; a strain-carrying seed instead
    # (J2 = 0 is 0/0 in the Lode angle, :290-292)
    xy = mesh["dof_coords"]
    Du = 1e-9 * np.stack([xy[:, 0] * xy[:, 1], xy[:, 0] + xy[:, 1] ** 2], 1).reshape(-1)
    Du[mesh["fixed"]] = 0.0
    K = len(load_steps)
    out = {"load": load_steps, "u_probe": np.full(K, np.nan), "newton_iterations": np.zeros(K, dtype=int),
           "plastic_fraction": np.full(K, np.nan), "n_converged": 0}
    for k, load in enumerate(load_steps):
        hist, ok = [], False
        Du_start = Du.copy()
        for it in range(max_it + 1):
            r = backend.residual(Du) - load * f_unit
            nrm = float(np.linalg.norm(r[free]))
            hist.append(nrm)
            if not np.isfinite(nrm) or it == max_it or nrm > 1e6 * max(hist[0], 1.0):
                break
            if nrm <= atol or (it > 0 and nrm <= rtol * hist[0]):
                ok = True
                break
            A = sp.csr_matrix((backend.tangent_csr(), col, row_ptr), shape=(n, n))
            Du[free] -= splu(A[free][:, free].tocsc()).solve(r[free])
        if not ok:
            Du = Du_start
            if verbose:
                print(f"step {k:2d} gamma {load:.3f}: Newton failed after {len(hist) - 1} iterations - collapse")
            break
        out["plastic_fraction"][k] = backend.plastic_fraction()
        backend.commit()
        u += Du
        out["u_probe"][k] = -u[mesh["probe"]]
        out["newton_iterations"][k] = len(hist) - 1
        out["n_converged"] = k + 1
        if verbose:
            print(f"step {k:2d} gamma {load:.3f} l = {load * H / C_COHESION:.3f} its {len(hist) - 1} -u_x(0,H) "
                  f"{out['u_probe'][k]:.6e} plastic {out['plastic_fraction'][k]:.3f}")
    out["u"] = u
    return out
