"""CPU: the tabulation oracle (oracle/tabulation.py - parity UNPINNED, see its header) against analytic
fields that P1/P2 interpolate exactly, including the fields the reference's own tests use
(test_operands_evaluation.py:20, part1.py:187); and the kernel's core header (tab_core.cuh, host build)
against the oracle."""

import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from dolfinx_external_operator_b200 import synthetic as syn
from oracle import tabulation as ot
from tab_util import tet_case, tri_case

HERE = os.path.dirname(os.path.abspath(__file__))


def _tab(m, kind, u, bs, cells=None):
    return ot.tabulate(kind, u, m["dofmap"], bs, m["x"], m["x_dofmap"], m["phi"], m["dphi"], m["dpsi"], cells)


def test_reference_test_field_gradient():
    """u = (0.1 x, 0.3 y) (test_operands_evaluation.py:20): F = I + grad u, tr(F^T F) at every point."""
    m = tri_case(degree=1, qdeg=2)
    xy = m["dof_coords"]
    u = np.stack([0.1 * xy[:, 0], 0.3 * xy[:, 1]], 1).reshape(-1)
    F = _tab(m, ot.DEF_GRAD, u, 2).reshape(-1, 2, 2)
    np.testing.assert_allclose(F, np.broadcast_to(np.array([[1.1, 0.0], [0.0, 1.3]]), F.shape), atol=1e-14)
    np.testing.assert_allclose(np.einsum("nij,nij->n", F, F), 1.1**2 + 1.3**2, rtol=1e-14)  # :32-35


def test_heat_field_value_and_gradient():
    """T = x^2 + y (part1.py:187) is in P2: value and gradient are exact at the quadrature points."""
    m = tri_case(degree=2, qdeg=2)
    xy, xq = m["dof_coords"], m["xq"]
    T = xy[:, 0] ** 2 + xy[:, 1]
    np.testing.assert_allclose(_tab(m, ot.VALUE, T, 1)[..., 0], xq[..., 0] ** 2 + xq[..., 1], atol=1e-14)
    g = _tab(m, ot.GRAD, T, 1)
    np.testing.assert_allclose(g[..., 0], 2 * xq[..., 0], atol=1e-13)
    np.testing.assert_allclose(g[..., 1], 1.0, atol=1e-13)


def test_mandel_strain_of_a_quadratic_displacement():
    m = tri_case(degree=2, qdeg=2)
    xy, xq = m["dof_coords"], m["xq"]
    u = np.stack([0.1 * xy[:, 0] + xy[:, 0] * xy[:, 1], 0.3 * xy[:, 1] + xy[:, 0] ** 2], 1).reshape(-1)
    e = _tab(m, ot.MANDEL_STRAIN, u, 2)
    g00, g01, g10, g11 = 0.1 + xq[..., 1], xq[..., 0], 2 * xq[..., 0], 0.3 + 0 * xq[..., 0]
    np.testing.assert_allclose(e[..., 0], g00, atol=1e-13)
    np.testing.assert_allclose(e[..., 1], g11, atol=1e-13)
    assert np.all(e[..., 2] == 0.0)
    np.testing.assert_allclose(e[..., 3], np.sqrt(2.0) * 0.5 * (g01 + g10), atol=1e-13)  # demo_vm:227


def test_entities_subset_and_tetrahedra():
    m = tri_case()
    u = syn.smooth_displacement(m["dof_coords"], seed=2).reshape(-1)
    full = _tab(m, ot.GRAD, u, 2)
    cells = np.array([5, 0, 17, 5, 3], dtype=np.int32)
    assert np.array_equal(_tab(m, ot.GRAD, u, 2, cells), full[cells])
    t = tet_case()
    xyz = t["dof_coords"]
    A = np.array([[0.1, 0.2, -0.3], [0.0, 0.5, 0.4], [0.7, -0.1, 0.2]])
    ulin = (xyz @ A.T).reshape(-1)  # linear field: P1-exact, grad u = A
    g = ot.tabulate(ot.GRAD, ulin, t["dofmap"], 3, t["x"], t["x_dofmap"], t["phi"], t["dphi"], t["dpsi"])
    np.testing.assert_allclose(g.reshape(-1, 3, 3), np.broadcast_to(A, (g.shape[0] * g.shape[1], 3, 3)), atol=1e-12)


@pytest.fixture(scope="module")
def hc():
    subprocess.check_call(["make", "-C", os.path.join(HERE, "hostcheck")], stdout=subprocess.DEVNULL)
    return C.CDLL(os.path.join(HERE, "hostcheck", "libhostcheck.so"))


def _hc_tab(hc, m, kind, u, bs, gdim, ncomp):
    p = lambda a: a.ctypes.data_as(C.c_void_p)  # noqa: E731
    nq, nb = m["phi"].shape
    out = np.empty((m["dofmap"].shape[0], nq, ncomp))
    u = np.ascontiguousarray(u)
    rc = hc.hostcheck_tab(gdim, bs, nb, nq, kind, p(m["phi"]), p(m["dphi"]), p(m["dpsi"]), p(m["dofmap"]),
                          p(m["x_dofmap"]), p(np.ascontiguousarray(m["x"])), p(u), C.c_int64(m["dofmap"].shape[0]), p(out))
    assert rc == 0
    return out


@pytest.mark.parametrize("kind,ncomp", [(ot.VALUE, 2), (ot.GRAD, 4), (ot.MANDEL_STRAIN, 4), (ot.DEF_GRAD, 4)])
def test_tab_core_against_oracle_p2_vector(hc, kind, ncomp):
    m = tri_case(nx=13, ny=11)
    u = syn.smooth_displacement(m["dof_coords"], seed=4).reshape(-1)
    ref = _tab(m, kind, u, 2)
    np.testing.assert_allclose(_hc_tab(hc, m, kind, u, 2, 2, ncomp), ref, rtol=1e-12, atol=1e-12 * np.abs(ref).max())


def test_tab_core_against_oracle_p1_scalar_and_tets(hc):
    m = tri_case(degree=1)
    T = np.sin(m["dof_coords"][:, 0]) + m["dof_coords"][:, 1] ** 2
    for kind, nc in ((ot.VALUE, 1), (ot.GRAD, 2)):
        ref = _tab(m, kind, T, 1)
        np.testing.assert_allclose(_hc_tab(hc, m, kind, T, 1, 2, nc), ref, rtol=1e-12, atol=1e-13)
    t = tet_case()
    u = np.random.default_rng(0).normal(size=t["n_dofs"] * 3)
    ref = ot.tabulate(ot.GRAD, u, t["dofmap"], 3, t["x"], t["x_dofmap"], t["phi"], t["dphi"], t["dpsi"])
    np.testing.assert_allclose(_hc_tab(hc, t, ot.GRAD, u, 3, 3, 9), ref, rtol=1e-11, atol=1e-11 * np.abs(ref).max())


def test_p3_triangle_reproduces_cubic_fields(hc):
    """P3 (nb = 10, the largest element of the fast path): a cubic field and its gradient are reproduced exactly by the
    oracle and by the kernels' core (tab_core.cuh, instantiation <2, bs, 10>)."""
    from tab_util import tri_case_discontinuous

    m = tri_case_discontinuous(degree=3)
    x, y = m["dof_coords"][:, 0], m["dof_coords"][:, 1]
    xq, yq = m["xq"][..., 0], m["xq"][..., 1]
    f = lambda x, y: x**3 - 2 * x * x * y + 0.5 * y**3 + x * y - 0.3 * y  # noqa: E731
    fx = lambda x, y: 3 * x * x - 4 * x * y + y  # noqa: E731
    fy = lambda x, y: -2 * x * x + 1.5 * y * y + x - 0.3  # noqa: E731
    T = f(x, y)
    val, grad = _tab(m, ot.VALUE, T, 1), _tab(m, ot.GRAD, T, 1)
    np.testing.assert_allclose(val[..., 0], f(xq, yq), atol=1e-13)
    np.testing.assert_allclose(grad, np.stack([fx(xq, yq), fy(xq, yq)], -1), atol=1e-11)
    np.testing.assert_allclose(_hc_tab(hc, m, ot.VALUE, T, 1, 2, 1), val, rtol=0, atol=1e-13)
    np.testing.assert_allclose(_hc_tab(hc, m, ot.GRAD, T, 1, 2, 2), grad, rtol=0, atol=1e-12)
    u = np.stack([f(x, y), fy(x, y)], 1).reshape(-1)  # vector field, bs = 2
    ref = _tab(m, ot.MANDEL_STRAIN, u, 2)
    np.testing.assert_allclose(ref[..., 0], fx(xq, yq), atol=1e-11)
    np.testing.assert_allclose(_hc_tab(hc, m, ot.MANDEL_STRAIN, u, 2, 2, 4), ref, rtol=0, atol=1e-12 * np.abs(ref).max())


def test_p2_tetrahedra_reproduce_quadratic_fields(hc):
    """P2 tetrahedra (nb = 10, gdim = 3): quadratic fields and F = I + grad u reproduced exactly by the oracle and by the
    kernels' core (instantiations <3, 1, 10>, <3, 3, 10>)."""
    from tab_util import tet_case_discontinuous

    m = tet_case_discontinuous()
    X = m["dof_coords"]
    xq = m["xq"]
    f = lambda p: p[..., 0] ** 2 - p[..., 0] * p[..., 2] + 0.5 * p[..., 1] * p[..., 2] + p[..., 1]  # noqa: E731
    gf = lambda p: np.stack([2 * p[..., 0] - p[..., 2], 0.5 * p[..., 2] + 1.0, -p[..., 0] + 0.5 * p[..., 1]], -1)  # noqa: E731
    T = f(X)
    geo = (m["x"], m["x_dofmap"], m["phi"], m["dphi"], m["dpsi"])
    val = ot.tabulate(ot.VALUE, T, m["dofmap"], 1, *geo)
    grad = ot.tabulate(ot.GRAD, T, m["dofmap"], 1, *geo)
    np.testing.assert_allclose(val[..., 0], f(xq), atol=1e-13)
    np.testing.assert_allclose(grad, gf(xq), atol=1e-11)
    np.testing.assert_allclose(_hc_tab(hc, m, ot.VALUE, T, 1, 3, 1), val, rtol=0, atol=1e-13)
    np.testing.assert_allclose(_hc_tab(hc, m, ot.GRAD, T, 1, 3, 3), grad, rtol=0, atol=1e-12)
    u = np.stack([f(X), 0.3 * X[:, 0] * X[:, 1], X[:, 2] ** 2], 1).reshape(-1)
    F = ot.tabulate(ot.DEF_GRAD, u, m["dofmap"], 3, *geo)
    np.testing.assert_allclose(F[..., 0:3], gf(xq) + np.array([1.0, 0, 0]), atol=1e-11)
    np.testing.assert_allclose(_hc_tab(hc, m, ot.DEF_GRAD, u, 3, 3, 9), F, rtol=0, atol=1e-12 * np.abs(F).max())


@pytest.mark.parametrize("order", ["shuffled", "rcm", "morton"])
def test_renumbered_mesh_is_the_same_mesh(order):
    """synthetic.renumber permutes cells / dofs / nodes consistently: the oracle returns the same per-cell values in the
    new cell order, and RCM restores locality (small dof spread per cell) where the shuffle destroys it."""
    from dolfinx_external_operator_b200 import synthetic as syn

    m = tri_case(nx=12, ny=9)
    r = syn.renumber(m, order, seed=2)
    u = syn.smooth_displacement(m["dof_coords"], seed=3)
    ur = np.empty_like(u)
    ur[r["dof_new"]] = u
    a = _tab(m, ot.MANDEL_STRAIN, u.reshape(-1), 2)
    b = _tab(r, ot.MANDEL_STRAIN, ur.reshape(-1), 2)
    assert np.array_equal(b, a[r["cell_old"]])
    assert sorted(r["dof_new"]) == list(range(m["n_dofs"])) and sorted(r["cell_old"]) == list(range(m["dofmap"].shape[0]))
    spread = lambda d: (d.max(axis=1) - d.min(axis=1)).mean()  # noqa: E731
    if order in ("rcm", "morton"):
        assert spread(r["dofmap"]) < 0.5 * spread(syn.renumber(m, "shuffled", seed=2)["dofmap"])
