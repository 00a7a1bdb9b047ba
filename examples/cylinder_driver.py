"""The von Mises demo problem of the reference, driven end to end through the device-side consumers.

replaces (as a self-contained driver, no DOLFINx / PETSc): the load-stepping Newton loop of
doc/demo/demo_plasticity_von_mises.py:183-191 (parameters), :205-222 (symmetry conditions), :253 (residual form with
the pressure on the inner boundary), :390-398 (tangent form), :500-513 (SNES residual callback = constitutive update +
assemble_vector), :542-565 (load steps up to 1.1 q_lim, history update) on a quarter of a thick-walled cylinder in
plane strain, P2 displacement, 3-point quadrature.

The pieces on the hot path are injected as a `backend` with three methods, so that the SAME driver runs on the GPU
(`GpuBackend`: `QuadratureForms.vm_residual` + `.matrix`, history resident in HBM - the product path) and, in the
tests, on a NumPy restatement:

    backend.residual(Du) -> b            internal force vector int sigma(Du) . eps(v) dx   (also updates the tangent)
    backend.tangent_csr() -> vals        CSR values of int C_tang eps(u_hat) . eps(v) dx on `backend.pattern`
    backend.commit()                     p += dp ; sigma_n <- sigma                          (demo_vm:564-565)

What stays on the host here - sparse LU (SciPy), the boundary load vector, Dirichlet elimination - is what stays with
DOLFINx/PETSc in the reference.  Known answers this problem offers (tests/test_cylinder_*.py): the Lame solution
while the response is elastic, quadratic Newton convergence with the consistent tangent, and the analytic collapse
load q_lim = 2/sqrt(3) sigma_0 ln(R_e/R_i) (demo_vm:542) at which the plastic zone reaches the outer radius.
"""

from __future__ import annotations

import numpy as np

from dolfinx_external_operator_b200 import elements as el
from dolfinx_external_operator_b200 import synthetic as syn

R_E, R_I = 1.3, 1.0  # demo_vm:183
E, NU, SIGMA_0 = 70e3, 0.3, 250.0  # :185-188
Q_LIM = 2.0 / np.sqrt(3.0) * np.log(R_E / R_I) * SIGMA_0  # :542


def quarter_ring_mesh(n_r: int, n_t: int):
    """Structured P2 triangulation of the quarter ring R_i <= r <= R_e, 0 <= theta <= pi/2 (straight-sided cells).
    Returns the mesh dict of `synthetic.triangle_mesh` plus the boundary data the demo gets from gmsh tags:
    `fixed` (scalar dofs with u_y = 0 on y = 0 and u_x = 0 on x = 0), `load` (scalar-dof vector of the unit pressure on
    the inner boundary, traction = -n, :253) and `probe` (scalar dof of u_x at (R_i, 0), :534)."""
    m = syn.triangle_mesh(n_r, n_t, 2, lx=1.0, ly=1.0)
    par = m["dof_coords"].copy()  # (s, t) in [0,1]^2: r = R_i + s (R_e - R_i), theta = t pi/2
    xv = m["x"]
    r, th = R_I + xv[:, 0] * (R_E - R_I), xv[:, 1] * (np.pi / 2)
    x = np.zeros_like(xv)
    x[:, 0], x[:, 1] = r * np.cos(th), r * np.sin(th)
    x[np.isclose(xv[:, 1], 0.0), 1] = 0.0
    x[np.isclose(xv[:, 1], 1.0), 0] = 0.0
    m["x"] = x
    # P2 nodes of straight-sided cells: vertices, then the midpoints of edge0=(v1,v2), edge1=(v0,v2), edge2=(v0,v1)
    dc = np.zeros((m["n_dofs"], 2))
    xd, dm = m["x_dofmap"], m["dofmap"]
    dc[dm[:, 0:3].reshape(-1)] = x[xd.reshape(-1), :2]
    for loc, (a, b) in zip((3, 4, 5), ((1, 2), (0, 2), (0, 1))):
        dc[dm[:, loc]] = 0.5 * (x[xd[:, a], :2] + x[xd[:, b], :2])
    m["dof_coords"] = dc
    on_bottom, on_left, on_inner = np.isclose(par[:, 1], 0.0), np.isclose(par[:, 1], 1.0), np.isclose(par[:, 0], 0.0)
    m["fixed"] = np.concatenate([2 * np.nonzero(on_bottom)[0] + 1, 2 * np.nonzero(on_left)[0]])
    # consistent nodal forces of a unit pressure on the inner boundary: per straight P2 edge (Simpson weights 1/6, 4/6,
    # 1/6) x length x (-n), n the outward normal of the material (towards the axis)
    load = np.zeros(2 * m["n_dofs"])
    for c in range(dm.shape[0]):
        for loc, (a, b) in zip((3, 4, 5), ((1, 2), (0, 2), (0, 1))):
            na, nb, nm = dm[c, a], dm[c, b], dm[c, loc]
            if on_inner[na] and on_inner[nb] and on_inner[nm]:
                t = dc[nb] - dc[na]
                nrm = np.array([t[1], -t[0]])
                if nrm @ dc[nm] < 0:  # make it point away from the axis (= -n)
                    nrm = -nrm
                for node, wgt in ((na, 1.0 / 6.0), (nb, 1.0 / 6.0), (nm, 4.0 / 6.0)):
                    load[2 * node:2 * node + 2] += wgt * nrm
    m["load"] = load
    m["probe"] = 2 * int(np.nonzero(on_inner & on_bottom)[0][0])
    X = el.triangle_quadrature(2)
    m["phi"], m["dphi"] = el.lagrange_triangle(2, X)
    m["weights"] = el.triangle_quadrature_weights(2)
    return m


def lame_inner_displacement(q: float) -> float:
    """u_r(R_i) of the elastic thick-walled cylinder under internal pressure q, plane strain."""
    mu = E / 2.0 / (1.0 + NU)
    a2, b2 = R_I**2, R_E**2
    A = q * a2 / (b2 - a2)
    return A / (2.0 * mu) * ((1.0 - 2.0 * NU) * R_I + b2 / R_I)


class GpuBackend:
    """The product path: Tabulator + VonMises (history resident) + QuadratureForms on one B200."""

    def __init__(self, mesh, ctx=None, exact: bool = True):
        from dolfinx_external_operator_b200 import QuadratureForms, Tabulator, VonMises

        self.tab = Tabulator(dofmap=mesh["dofmap"], x_dofmap=mesh["x_dofmap"], x=mesh["x"], phi=mesh["phi"],
                             dphi=mesh["dphi"], bs=2, n_dofs=mesh["n_dofs"], ctx=ctx)
        self.forms = QuadratureForms(self.tab, mesh["weights"])
        self.vm = VonMises(E=E, nu=NU, sigma_0=SIGMA_0, n_qp=3 * mesh["dofmap"].shape[0], ctx=self.tab.ctx)
        self.pattern = self.forms.set_pattern()
        self.exact = exact
        self._vals = None

    def residual(self, Du):
        return self.forms.vm_residual(self.vm, Du, exact=self.exact)

    def tangent_csr(self):
        self._vals = self.forms.matrix("mandel_strain", "mandel_strain", self.forms.C_tang, vals=self._vals)
        return self._vals.to_host()

    def plastic_fraction(self) -> float:
        return float((self.vm.dp_dev.to_host() > 0).mean())

    def commit(self):
        self.vm.commit()


def solve(mesh, backend, n_steps: int = 20, max_load: float = 1.1, rtol: float = 1e-8, atol: float = 1e-8,
          max_it: int = 50, verbose: bool = False):
    """Load stepping of demo_vm:539-565.  Returns dict(load (n_steps,), u_probe, newton_iterations, residual_histories,
    plastic_fraction, u (final displacement))."""
    import scipy.sparse as sp
    from scipy.sparse.linalg import splu

    n = 2 * mesh["n_dofs"]
    free = np.setdiff1d(np.arange(n), mesh["fixed"])
    row_ptr, col = backend.pattern
    loads = Q_LIM * np.linspace(0.0, max_load, n_steps) ** 0.5  # :544-545
    u = np.zeros(n)
    out = {"load": loads / Q_LIM, "u_probe": np.zeros(n_steps), "newton_iterations": np.zeros(n_steps, dtype=int),
           "residual_histories": [], "plastic_fraction": np.zeros(n_steps)}
    for k, q in enumerate(loads):
        # :551 sets Du = eps so that sigma_eq != 0 (0/0 in the radial return, :317).  A CONSTANT field has zero strain in
        # exact arithmetic - the demo lives on the rounding residue of sum_a dphi_a - so the seed here is eps * (1 + x):
        # strain eps * I whatever the summation order.
        Du = np.finfo(np.float64).eps * (1.0 + mesh["dof_coords"]).reshape(-1)
        Du[mesh["fixed"]] = 0.0
        hist = []
        for it in range(max_it + 1):
            r = backend.residual(Du) - q * mesh["load"]
            nrm = float(np.linalg.norm(r[free]))
            hist.append(nrm)
            if nrm <= atol or (it > 0 and nrm <= rtol * hist[0]):
                break
            if it == max_it:
                raise RuntimeError(f"Newton did not converge at load step {k}: {hist}")
            if not np.isfinite(nrm):
                raise RuntimeError(f"non-finite residual at load step {k}, iteration {it}")
            A = sp.csr_matrix((backend.tangent_csr(), col, row_ptr), shape=(n, n))
            Du[free] -= splu(A[free][:, free].tocsc()).solve(r[free])
        out["plastic_fraction"][k] = backend.plastic_fraction()
        backend.commit()
        u += Du
        out["u_probe"][k] = u[mesh["probe"]]
        out["newton_iterations"][k] = len(hist) - 1
        out["residual_histories"].append(hist)
        if verbose:
            print(f"step {k:2d} q/q_lim {q / Q_LIM:.3f} its {len(hist) - 1} u_x(R_i,0) {u[mesh['probe']]:.6e} "
                  f"plastic {out['plastic_fraction'][k]:.3f}")
    out["u"] = u
    return out
