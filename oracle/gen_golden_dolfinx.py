#!/usr/bin/env python
"""TEST INFRASTRUCTURE - golden vectors for the operand tabulation and the device-side consumers
(row f1) made by DOLFINx ITSELF.  Runs only where fenics-dolfinx >= 0.10, < 0.11 (+ basix, ufl, mpi4py) is installed -
not in the build container, which is why tests/golden/ has no tab_dolfinx_*.npz yet and DESIGN.md calls those rows
"parity unpinned".  One serial run

    python oracle/gen_golden_dolfinx.py [out_dir = tests/golden]

writes tests/golden/tab_dolfinx_<case>.npz; tests/test_dolfinx_golden_cpu.py (oracle + kernel core on the CPU) and
tests/test_dolfinx_golden_gpu.py (the sm_100a kernels through the C ABI) pick the files up when present and are
skipped otherwise.

What is recorded per case, exactly as the reference obtains it (src/dolfinx_external_operator/external_operator.py:
365-402): the mesh arrays the kernels take (V.dofmap.list, mesh.geometry.dofmap, mesh.geometry.x), the basix tables at
the quadrature element's interpolation points, a coefficient vector, and
    fem.Expression(operand, points, dtype).eval(mesh, cells)
for the four operand kinds; plus, for the vector spaces, the assembled residual  inner(s, OP(v)) dx  and the action of the
assembled Jacobian  inner(D OP(u_hat), OP(v)) dx  on a vector, with s / D given quadrature fields
(reference/test/test_external_operators_evaluation.py:20-45 compares the same kind of assembled objects)."""

import os
import sys

import numpy as np


def cases():
    from dolfinx import mesh as dmesh
    from mpi4py import MPI

    comm = MPI.COMM_SELF
    yield "tri_p1_vec", dmesh.create_unit_square(comm, 7, 5), 1, 2, 2
    yield "tri_p2_vec", dmesh.create_unit_square(comm, 7, 5), 2, 2, 2
    yield "tri_p2_scalar", dmesh.create_unit_square(comm, 6, 6), 2, 1, 2
    yield "tri_p3_vec", dmesh.create_unit_square(comm, 4, 3), 3, 2, 3
    yield "tet_p1_vec", dmesh.create_unit_cube(comm, 3, 2, 2), 1, 3, 2
    yield "tet_p2_vec", dmesh.create_unit_cube(comm, 2, 2, 2), 2, 3, 2
    yield "quad_q2_vec", dmesh.create_unit_square(comm, 5, 4, dmesh.CellType.quadrilateral), 2, 2, 3
    yield "hex_q1_vec", dmesh.create_unit_cube(comm, 2, 2, 2, dmesh.CellType.hexahedron), 1, 3, 2


def run_case(name, domain, degree, bs, qdeg):
    import basix
    import basix.ufl
    import ufl
    from dolfinx import fem

    gdim = domain.geometry.dim
    shape = () if bs == 1 else (bs,)
    V = fem.functionspace(domain, ("P", degree, shape))
    u = fem.Function(V)
    rng = np.random.default_rng(0)
    u.x.array[:] = rng.normal(size=u.x.array.size)
    cell = domain.topology.cell_name()
    Qe = basix.ufl.quadrature_element(cell, degree=qdeg, value_shape=())
    points, weights = basix.make_quadrature(Qe.cell_type, qdeg)
    map_c = domain.topology.index_map(domain.topology.dim)
    cells = np.arange(map_c.size_local + map_c.num_ghosts, dtype=np.int32)

    def evaluate(expr):
        return np.asarray(fem.Expression(expr, points, dtype=np.float64).eval(domain, cells))

    out = {
        "dofmap": np.asarray(V.dofmap.list, dtype=np.int32), "x_dofmap": np.asarray(domain.geometry.dofmap, dtype=np.int32),
        "x": np.asarray(domain.geometry.x, dtype=np.float64), "bs": bs, "degree": degree, "gdim": gdim, "cell": cell,
        "n_dofs": V.dofmap.index_map.size_local + V.dofmap.index_map.num_ghosts, "u": u.x.array.copy(),
        "points": np.asarray(points), "weights": np.asarray(weights),
    }
    # the element tables the kernels are given: basix tabulate(1, points) of the SCALAR element and of the geometry element
    tab = np.asarray(V.element.basix_element.tabulate(1, points))
    out["phi"], out["dphi"] = tab[0].reshape(tab.shape[1], -1), tab[1:1 + gdim].reshape(gdim, tab.shape[1], -1)
    gtab = np.asarray(domain.geometry.cmap.tabulate(1, points))
    out["dgeo"] = gtab[1:1 + gdim].reshape(gdim, gtab.shape[1], -1)
    out["value"] = evaluate(u)
    out["grad"] = evaluate(ufl.grad(u))
    if bs == gdim:
        out["def_grad"] = evaluate(ufl.Identity(gdim) + ufl.grad(u))
    if bs == 2 and gdim == 2:
        g = ufl.grad(u)
        out["mandel_strain"] = evaluate(ufl.as_vector([g[0, 0], g[1, 1], 0.0, np.sqrt(2.0) * 0.5 * (g[0, 1] + g[1, 0])]))
    # assembled residual / Jacobian action with given quadrature fields (gradient operand on both sides)
    if bs > 1:
        Qs = fem.functionspace(domain, basix.ufl.quadrature_element(cell, degree=qdeg, value_shape=(bs, gdim)))
        QD = fem.functionspace(domain, basix.ufl.quadrature_element(cell, degree=qdeg, value_shape=(bs, gdim, bs, gdim)))
        s, D = fem.Function(Qs), fem.Function(QD)
        s.x.array[:] = rng.normal(size=s.x.array.size)
        D.x.array[:] = rng.normal(size=D.x.array.size)
        dx = ufl.Measure("dx", domain=domain, metadata={"quadrature_scheme": "default", "quadrature_degree": qdeg})
        v, w = ufl.TestFunction(V), ufl.TrialFunction(V)
        i, j, k, l = ufl.indices(4)
        b = fem.assemble_vector(fem.form(ufl.inner(s, ufl.grad(v)) * dx))
        A = fem.assemble_matrix(fem.form(D[i, j, k, l] * ufl.grad(w)[k, l] * ufl.grad(v)[i, j] * dx))
        xvec = rng.normal(size=u.x.array.size)
        out.update(s=s.x.array.copy(), D=D.x.array.copy(), b_grad=b.array.copy(), xvec=xvec,
                   y_grad_grad=A.to_scipy() @ xvec if hasattr(A, "to_scipy") else A.to_dense() @ xvec)
    return out


def main():
    out_dir = sys.argv[1] if len(sys.argv) > 1 else os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                                                                "tests", "golden")
    os.makedirs(out_dir, exist_ok=True)
    import dolfinx

    for name, domain, degree, bs, qdeg in cases():
        data = run_case(name, domain, degree, bs, qdeg)
        data["dolfinx_version"] = dolfinx.__version__
        np.savez_compressed(os.path.join(out_dir, f"tab_dolfinx_{name}.npz"), **data)
        print("wrote", name, {k: np.shape(v) for k, v in data.items() if hasattr(v, "shape")})


if __name__ == "__main__":
    main()
