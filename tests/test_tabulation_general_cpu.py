"""The general tabulation oracle (oracle/tabulation.py::tabulate_general: table sets, per-point Jacobians, facet
entities) against analytic fields on distorted quadrilateral / hexahedral / simplex meshes, and against the
affine-simplex oracle where both apply.  (The arithmetic of the reference lives in DOLFINx/FFCx/basix: parity
unpinned, see the oracle's header.)"""

import numpy as np
import pytest

from oracle import tabulation as ot
from tab_util import general_case, tri_case


def _linear_field(m, bs, seed=0):
    rng = np.random.default_rng(seed)
    gdim = m["gdim"]
    A, b = rng.normal(size=(bs, gdim)), rng.normal(size=bs)
    u = m["dof_coords"][:, :gdim] @ A.T + b  # (n_dofs, bs)
    return A, b, u.reshape(-1)


@pytest.mark.parametrize("cell,degree,bs", [("triangle", 2, 2), ("quadrilateral", 1, 1), ("quadrilateral", 2, 2),
                                            ("hexahedron", 1, 3), ("hexahedron", 2, 1), ("tetrahedron", 1, 3)])
def test_linear_fields_are_reproduced_on_distorted_cells(cell, degree, bs):
    m = general_case(cell, degree, bs)
    A, b, u = _linear_field(m, bs)
    args = (u, m["dofmap"], bs, m["x"], m["x_dofmap"], m["phi"], m["dphi"], m["dgeo"])
    val = ot.tabulate_general(ot.VALUE, *args)
    np.testing.assert_allclose(val, m["xq"][0] @ A.T + b, rtol=1e-12, atol=1e-12)
    grad = ot.tabulate_general(ot.GRAD, *args)
    np.testing.assert_allclose(grad, np.broadcast_to(A.reshape(-1), grad.shape), rtol=1e-11, atol=1e-11)
    if bs == m["gdim"]:
        F = ot.tabulate_general(ot.DEF_GRAD, *args)
        np.testing.assert_allclose(F, np.broadcast_to((A + np.eye(bs)).reshape(-1), F.shape), rtol=1e-11, atol=1e-11)


@pytest.mark.parametrize("cell,degree", [("triangle", 1), ("triangle", 2), ("quadrilateral", 2), ("hexahedron", 1),
                                         ("tetrahedron", 1)])
def test_facet_entities_evaluate_on_the_facets(cell, degree):
    """(cell, local facet) pairs (test_codim_external_operator.py:75-109): the value of u = x + y (+ z) - the field
    that test interpolates - at the facet points, and the gradient there."""
    m = general_case(cell, degree, 1, facets=True)
    gdim = m["gdim"]
    u = m["dof_coords"][:, :gdim].sum(axis=1)
    n_f = m["phi"].shape[0]
    rng = np.random.default_rng(1)
    ent = np.stack([rng.integers(0, m["dofmap"].shape[0], 40), rng.integers(0, n_f, 40)], axis=1).astype(np.int32)
    args = (u, m["dofmap"], 1, m["x"], m["x_dofmap"], m["phi"], m["dphi"], m["dgeo"])
    val = ot.tabulate_general(ot.VALUE, *args, entities=ent)
    xq = m["xq"][ent[:, 1], ent[:, 0]]  # (n, nq, gdim) physical facet points
    np.testing.assert_allclose(val[..., 0], xq.sum(axis=2), rtol=1e-12, atol=1e-12)
    grad = ot.tabulate_general(ot.GRAD, *args, entities=ent)
    np.testing.assert_allclose(grad, np.ones_like(grad), rtol=1e-11, atol=1e-11)
    # the facet points really lie on the cell boundary: one barycentric / tensor coordinate is 0 or 1
    from dolfinx_external_operator_b200 import elements as el

    Xs = el.facet_points(cell, m["X"])
    for X in Xs:
        lam = np.concatenate([X, 1 - X.sum(1, keepdims=True)], axis=1) if cell in ("triangle", "tetrahedron") else np.concatenate([X, 1 - X], axis=1)
        assert np.all(np.isclose(lam, 0.0).any(axis=1))


def test_general_oracle_equals_affine_oracle_on_triangles():
    m = tri_case()
    g = general_case("triangle", 2, 2)
    rng = np.random.default_rng(5)
    u = rng.normal(size=2 * g["n_dofs"])
    cells = rng.permutation(g["dofmap"].shape[0])[:33].astype(np.int32)
    for kind in (ot.VALUE, ot.GRAD, ot.MANDEL_STRAIN, ot.DEF_GRAD):
        a = ot.tabulate(kind, u, g["dofmap"], 2, g["x"], g["x_dofmap"], g["phi"][0], g["dphi"][0], m["dpsi"], cells=cells)
        b = ot.tabulate_general(kind, u, g["dofmap"], 2, g["x"], g["x_dofmap"], g["phi"], g["dphi"], g["dgeo"], entities=cells)
        np.testing.assert_allclose(b, a, rtol=1e-13, atol=1e-13)
