"""Lagrange tables on the reference simplices, NumPy only.

The tabulation kernels take their basis tables as INPUTS: in production they are what basix returns
(`V.element.basix_element.tabulate(1, Q.element.interpolation_points)`, cf. external_operator.py:145,200).
basix is not installable in the build container or on the GPU box, so the synthetic benchmarks and the tests
use these closed-form P1/P2 tables instead.  Node ordering follows the basix convention recalled in SURVEY.md
appendix B (vertices first, then edges, edge e opposite to... edge0=(v1,v2), edge1=(v0,v2), edge2=(v0,v1)); parity
of kernel vs oracle only needs a CONSISTENT (table, dofmap) pair.
"""

from __future__ import annotations

import numpy as np


def triangle_quadrature(degree: int) -> np.ndarray:
    """Reference points of the default Gauss-Jacobi-free simplex rules used by the demos:
    degree <= 1: centroid; degree 2: the 3-point rule (1/6,1/6), (1/6,2/3), (2/3,1/6)."""
    if degree <= 1:
        return np.array([[1.0 / 3.0, 1.0 / 3.0]])
    if degree == 2:
        return np.array([[1.0 / 6.0, 1.0 / 6.0], [1.0 / 6.0, 2.0 / 3.0], [2.0 / 3.0, 1.0 / 6.0]])
    raise NotImplementedError("only the 1- and 3-point triangle rules are tabulated here; pass basix points")


def triangle_quadrature_weights(degree: int) -> np.ndarray:
    """Weights of `triangle_quadrature(degree)` on the reference triangle (they sum to its area, 1/2)."""
    if degree <= 1:
        return np.array([0.5])
    if degree == 2:
        return np.full(3, 1.0 / 6.0)
    raise NotImplementedError("only the 1- and 3-point triangle rules are tabulated here; pass basix weights")


def lagrange_triangle(degree: int, X: np.ndarray):
    """(phi (nq, nb), dphi (2, nq, nb)) of P1 / P2 / P3 on the reference triangle at points X (nq, 2)."""
    x, y = X[:, 0], X[:, 1]
    one, zero = np.ones_like(x), np.zeros_like(x)
    if degree == 1:
        phi = np.stack([1 - x - y, x, y], axis=1)
        dx = np.stack([-one, one, zero], axis=1)
        dy = np.stack([-one, zero, one], axis=1)
    elif degree == 2:
        l0 = 1 - x - y
        phi = np.stack([l0 * (2 * l0 - 1), x * (2 * x - 1), y * (2 * y - 1), 4 * x * y, 4 * y * l0, 4 * x * l0], axis=1)
        dx = np.stack([-(4 * l0 - 1), 4 * x - 1, zero, 4 * y, -4 * y, 4 * (l0 - x)], axis=1)
        dy = np.stack([-(4 * l0 - 1), zero, 4 * y - 1, 4 * x, 4 * (l0 - y), -4 * x], axis=1)
    elif degree == 3:
        # equispaced P3: vertices, two nodes per edge (edge0=(v1,v2), edge1=(v0,v2), edge2=(v0,v1), from the lower to the
        # higher local vertex), interior node.  (DOLFINx's default P3 variant is GLL-warped: production tables come from
        # basix; this closed form only has to be consistent with the dofmaps of the synthetic meshes.)
        l = [1 - x - y, x, y]  # noqa: E741
        dl = [(-one, -one), (one, zero), (zero, one)]
        funcs = []  # (value, d/dx, d/dy) built from barycentric products

        def vert(i):
            li = l[i]
            v = 0.5 * li * (3 * li - 1) * (3 * li - 2)
            dv = 0.5 * ((3 * li - 1) * (3 * li - 2) + 3 * li * (3 * li - 2) + 3 * li * (3 * li - 1))
            return v, dv * dl[i][0], dv * dl[i][1]

        def edge(i, j):  # node at 1/3 from vertex i towards vertex j
            li, lj = l[i], l[j]
            v = 4.5 * li * lj * (3 * li - 1)
            dvi, dvj = 4.5 * lj * (6 * li - 1), 4.5 * li * (3 * li - 1)
            return v, dvi * dl[i][0] + dvj * dl[j][0], dvi * dl[i][1] + dvj * dl[j][1]

        funcs += [vert(0), vert(1), vert(2)]
        funcs += [edge(1, 2), edge(2, 1), edge(0, 2), edge(2, 0), edge(0, 1), edge(1, 0)]
        b = 27.0 * l[0] * l[1] * l[2]
        db = [27.0 * (l[1] * l[2] * dl[0][k] + l[0] * l[2] * dl[1][k] + l[0] * l[1] * dl[2][k]) for k in (0, 1)]
        funcs.append((b, db[0], db[1]))
        phi = np.stack([f[0] for f in funcs], axis=1)
        dx = np.stack([f[1] for f in funcs], axis=1)
        dy = np.stack([f[2] for f in funcs], axis=1)
    else:
        raise NotImplementedError("closed-form tables for P1, P2 and P3; pass basix tables for higher degrees")
    return np.ascontiguousarray(phi), np.ascontiguousarray(np.stack([dx, dy]))


def lagrange_triangle_nodes(degree: int) -> np.ndarray:
    """Reference coordinates of the nodes of `lagrange_triangle(degree, .)`, in its ordering."""
    if degree == 1:
        return np.array([[0.0, 0.0], [1.0, 0.0], [0.0, 1.0]])
    if degree == 2:
        return np.array([[0.0, 0.0], [1.0, 0.0], [0.0, 1.0], [0.5, 0.5], [0.0, 0.5], [0.5, 0.0]])
    if degree == 3:
        t = 1.0 / 3.0
        return np.array([[0.0, 0.0], [1.0, 0.0], [0.0, 1.0], [2 * t, t], [t, 2 * t], [0.0, t], [0.0, 2 * t], [t, 0.0],
                         [2 * t, 0.0], [t, t]])
    raise NotImplementedError


_TET_EDGES = ((2, 3), (1, 3), (1, 2), (0, 3), (0, 2), (0, 1))  # basix sub-entity order of the reference tetrahedron


def lagrange_tetrahedron(degree: int, X: np.ndarray):
    """(phi (nq, nb), dphi (3, nq, nb)) of P1 (nb = 4) / P2 (nb = 10: vertices, then the edge midpoints in the order
    (v2,v3), (v1,v3), (v1,v2), (v0,v3), (v0,v2), (v0,v1)) on the reference tetrahedron."""
    x, y, z = X[:, 0], X[:, 1], X[:, 2]
    one, zero = np.ones_like(x), np.zeros_like(x)
    lam = [1 - x - y - z, x, y, z]
    dlam = [(-one, -one, -one), (one, zero, zero), (zero, one, zero), (zero, zero, one)]
    if degree == 1:
        phi = np.stack(lam, axis=1)
        d = np.stack([np.stack([dlam[a][k] for a in range(4)], 1) for k in range(3)])
    elif degree == 2:
        vals = [lam[i] * (2 * lam[i] - 1) for i in range(4)] + [4 * lam[i] * lam[j] for i, j in _TET_EDGES]
        ders = []
        for k in range(3):
            dk = [(4 * lam[i] - 1) * dlam[i][k] for i in range(4)]
            dk += [4 * (lam[i] * dlam[j][k] + lam[j] * dlam[i][k]) for i, j in _TET_EDGES]
            ders.append(np.stack(dk, 1))
        phi, d = np.stack(vals, axis=1), np.stack(ders)
    else:
        raise NotImplementedError("closed-form tetrahedron tables for P1 and P2; pass basix tables")
    return np.ascontiguousarray(phi), np.ascontiguousarray(d)


def lagrange_tetrahedron_nodes(degree: int) -> np.ndarray:
    """Reference coordinates of the nodes of `lagrange_tetrahedron(degree, .)`, in its ordering."""
    v = np.array([[0.0, 0, 0], [1.0, 0, 0], [0.0, 1, 0], [0.0, 0, 1]])
    if degree == 1:
        return v
    if degree == 2:
        return np.concatenate([v, np.array([0.5 * (v[i] + v[j]) for i, j in _TET_EDGES])])
    raise NotImplementedError


def p1_geometry_derivatives(gdim: int) -> np.ndarray:
    """d psi_v / d X_k of the affine geometry element, shape (gdim, gdim + 1)."""
    d = np.zeros((gdim, gdim + 1))
    d[:, 0] = -1.0
    for k in range(gdim):
        d[k, k + 1] = 1.0
    return d


# ----------------------------------------------------------------------------- tensor-product cells
def lagrange_interval(degree: int, t: np.ndarray):
    """(phi (n, degree+1), dphi (n, degree+1)) of the equispaced Lagrange basis on [0, 1], nodes in increasing order."""
    nodes = np.linspace(0.0, 1.0, degree + 1)
    t = np.asarray(t, dtype=np.float64)
    phi = np.ones((t.size, degree + 1))
    dphi = np.zeros((t.size, degree + 1))
    for a in range(degree + 1):
        for m in range(degree + 1):
            if m != a:
                phi[:, a] *= (t - nodes[m]) / (nodes[a] - nodes[m])
        for m in range(degree + 1):
            if m == a:
                continue
            term = np.full(t.size, 1.0 / (nodes[a] - nodes[m]))
            for l in range(degree + 1):  # noqa: E741
                if l != a and l != m:
                    term *= (t - nodes[l]) / (nodes[a] - nodes[l])
            dphi[:, a] += term
    return phi, dphi


def lagrange_quadrilateral(degree: int, X: np.ndarray):
    """(phi (nq, nb), dphi (2, nq, nb)), nb = (degree+1)^2, TENSOR ordering: basis (i, j) -> i * (degree+1) + j with
    i along x, j along y (NOT basix's ordering: production tables come from basix together with DOLFINx's dofmaps;
    this closed form is paired with `synthetic.quad_mesh`, whose dofmaps use the same tensor ordering)."""
    px, dx = lagrange_interval(degree, X[:, 0])
    py, dy = lagrange_interval(degree, X[:, 1])
    nq = X.shape[0]
    phi = np.einsum("qi,qj->qij", px, py).reshape(nq, -1)
    d0 = np.einsum("qi,qj->qij", dx, py).reshape(nq, -1)
    d1 = np.einsum("qi,qj->qij", px, dy).reshape(nq, -1)
    return np.ascontiguousarray(phi), np.ascontiguousarray(np.stack([d0, d1]))


def lagrange_hexahedron(degree: int, X: np.ndarray):
    """(phi (nq, nb), dphi (3, nq, nb)), nb = (degree+1)^3, tensor ordering (i, j, k) -> (i * n + j) * n + k."""
    px, dx = lagrange_interval(degree, X[:, 0])
    py, dy = lagrange_interval(degree, X[:, 1])
    pz, dz = lagrange_interval(degree, X[:, 2])
    nq = X.shape[0]
    phi = np.einsum("qi,qj,qk->qijk", px, py, pz).reshape(nq, -1)
    d0 = np.einsum("qi,qj,qk->qijk", dx, py, pz).reshape(nq, -1)
    d1 = np.einsum("qi,qj,qk->qijk", px, dy, pz).reshape(nq, -1)
    d2 = np.einsum("qi,qj,qk->qijk", px, py, dz).reshape(nq, -1)
    return np.ascontiguousarray(phi), np.ascontiguousarray(np.stack([d0, d1, d2]))


# reference vertices (DOLFINx / basix ordering) and the vertices of every local facet
REFERENCE_VERTICES = {
    "triangle": np.array([[0.0, 0.0], [1.0, 0.0], [0.0, 1.0]]),
    "tetrahedron": np.array([[0.0, 0, 0], [1.0, 0, 0], [0, 1.0, 0], [0, 0, 1.0]]),
    "quadrilateral": np.array([[0.0, 0.0], [1.0, 0.0], [0.0, 1.0], [1.0, 1.0]]),
    "hexahedron": np.array([[x, y, z] for z in (0.0, 1.0) for y in (0.0, 1.0) for x in (0.0, 1.0)]),
}
FACET_VERTICES = {
    "triangle": [(1, 2), (0, 2), (0, 1)],
    "tetrahedron": [(1, 2, 3), (0, 2, 3), (0, 1, 3), (0, 1, 2)],
    "quadrilateral": [(0, 1), (0, 2), (1, 3), (2, 3)],
    "hexahedron": [(0, 1, 2, 3), (0, 1, 4, 5), (0, 2, 4, 6), (1, 3, 5, 7), (2, 3, 6, 7), (4, 5, 6, 7)],
}


def facet_points(cell: str, Xf: np.ndarray) -> np.ndarray:
    """Points `Xf` (nq, tdim-1) of the reference FACET mapped onto every local facet of the reference cell:
    (n_facets, nq, tdim).  This is where `fem.Expression.eval` evaluates for (cell, local_facet) entities
    (test_codim_external_operator.py:75-109): X = v0 + sum_k s_k (v_{k+1} - v0), with the facet's vertices in the
    order of the table above (for quadrilateral faces of a hexahedron: the first three vertices span the face)."""
    V = REFERENCE_VERTICES[cell]
    Xf = np.asarray(Xf, dtype=np.float64).reshape(len(Xf), -1)
    out = []
    for fv in FACET_VERTICES[cell]:
        v0 = V[fv[0]]
        X = np.tile(v0, (Xf.shape[0], 1))
        for k in range(Xf.shape[1]):
            X = X + Xf[:, k : k + 1] * (V[fv[k + 1]] - v0)
        out.append(X)
    return np.stack(out)


def tabulate_on(cell: str, degree: int, X: np.ndarray):
    """Closed-form (phi, dphi) of the Lagrange element of `degree` on `cell` at X - stands in for
    `basix_element.tabulate(1, X)` in the synthetic tests."""
    if cell == "triangle":
        return lagrange_triangle(degree, X)
    if cell == "tetrahedron":
        return lagrange_tetrahedron(degree, X)
    if cell == "quadrilateral":
        return lagrange_quadrilateral(degree, X)
    if cell == "hexahedron":
        return lagrange_hexahedron(degree, X)
    raise ValueError(cell)


def table_sets(cell: str, degree: int, geometry_degree: int, X: np.ndarray, facets: bool = False):
    """(phi (n_sets, nq, nb), dphi (n_sets, tdim, nq, nb), dgeo (n_sets, tdim, nq, ng)) for `eo_gtab_create`:
    one set at the cell points X, or one per local facet at the facet points X (facets=True)."""
    Xs = facet_points(cell, X) if facets else np.asarray(X, dtype=np.float64)[None]
    phi, dphi, dgeo = [], [], []
    for Xc in Xs:
        p, d = tabulate_on(cell, degree, Xc)
        _, g = tabulate_on(cell, geometry_degree, Xc)
        phi.append(p), dphi.append(d), dgeo.append(g)
    return np.ascontiguousarray(np.stack(phi)), np.ascontiguousarray(np.stack(dphi)), np.ascontiguousarray(np.stack(dgeo))
