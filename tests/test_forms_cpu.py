"""Known-answer anchors of the forms oracle (oracle/forms.py - parity unpinned against DOLFINx, see its header):
adjointness with the analytically checked tabulation, exactly integrable fields, rigid-body modes and symmetry of
the elastic stiffness, CSR assembly == action, and the reference's Taylor test of residual vs tangent
(demo_plasticity_mohr_coulomb.py:1203-1235) with the von Mises oracle."""

import numpy as np
import pytest

from dolfinx_external_operator_b200 import elements as el
from dolfinx_external_operator_b200 import forms as pkg_forms
from dolfinx_external_operator_b200 import synthetic as syn
from oracle import constitutive as oc
from oracle import forms as of
from oracle import tabulation as ot
from tab_util import tet_case, tri_case

W3 = el.triangle_quadrature_weights(2)


def _geo(m):
    return (m["x"], m["x_dofmap"], m["phi"], m["dphi"], m["dpsi"])


@pytest.mark.parametrize("kind", [ot.VALUE, ot.GRAD, ot.MANDEL_STRAIN, ot.DEF_GRAD])
def test_vector_is_the_weighted_transpose_of_the_tabulation(kind):
    m = tri_case(nx=7, ny=5)
    rng = np.random.default_rng(0)
    u = rng.normal(size=2 * m["n_dofs"])
    op = ot.tabulate(kind, u, m["dofmap"], 2, *_geo(m))
    if kind == ot.DEF_GRAD:
        op = op - np.eye(2).reshape(-1)[None, None]
    s = rng.normal(size=op.shape)
    b = of.assemble_vector(kind, s, W3, m["dofmap"], 2, m["n_dofs"], *_geo(m))
    _, adet = of.operand_matrix(kind, m["dofmap"], 2, *_geo(m))
    lhs = np.einsum("cqk,cqk,q,c->", op, s, W3, adet)
    assert abs(lhs - u @ b) <= 1e-12 * abs(lhs)


def test_exact_integrals_on_the_unit_square():
    m = tri_case(nx=6, ny=4, degree=2)  # jittered interior, the domain is still [0,1]^2
    xq = m["xq"]
    one = np.ones(xq.shape[:2] + (1,))
    b = of.assemble_vector(ot.VALUE, one, W3, m["dofmap"], 1, m["n_dofs"], *_geo(m))
    assert abs(b.sum() - 1.0) < 1e-14  # sum_i int phi_i = |domain|
    N = (xq[..., 0] * xq[..., 1] + xq[..., 1] ** 2)[..., None]  # degree 2: the 3-point rule is exact for sum_i b_i
    b = of.assemble_vector(ot.VALUE, N, W3, m["dofmap"], 1, m["n_dofs"], *_geo(m))
    assert abs(b.sum() - (0.25 + 1.0 / 3.0)) < 1e-14
    # int grad(v) . c over the domain against v = a P1 field interpolated into P2: = int c . grad(x) dx
    c = np.broadcast_to(np.array([0.3, -1.1]), xq.shape[:2] + (2,))
    b = of.assemble_vector(ot.GRAD, c, W3, m["dofmap"], 1, m["n_dofs"], *_geo(m))
    assert abs(b @ m["dof_coords"][:, 0] - 0.3) < 1e-14 and abs(b @ m["dof_coords"][:, 1] + 1.1) < 1e-14


def test_elastic_stiffness_rigid_modes_symmetry_and_csr():
    m = tri_case(nx=5, ny=4)
    nc = m["dofmap"].shape[0]
    Ce = oc.elastic_stiffness(oc.VonMisesParams().lmbda, oc.VonMisesParams().mu)
    D = np.broadcast_to(Ce.reshape(-1), (nc, 3, 16)).copy()
    args = (W3, m["dofmap"], 2, m["n_dofs"], *_geo(m))
    xy = m["dof_coords"]
    scale = None
    rng = np.random.default_rng(1)
    xr, yr = rng.normal(size=2 * m["n_dofs"]), rng.normal(size=2 * m["n_dofs"])
    Ax, Ay = (of.apply_action(ot.MANDEL_STRAIN, ot.MANDEL_STRAIN, D, v, *args) for v in (xr, yr))
    scale = np.abs(Ax).max()
    for mode in (np.stack([np.ones(len(xy)), np.zeros(len(xy))], 1), np.stack([np.zeros(len(xy)), np.ones(len(xy))], 1),
                 np.stack([-xy[:, 1], xy[:, 0]], 1)):
        y = of.apply_action(ot.MANDEL_STRAIN, ot.MANDEL_STRAIN, D, mode.reshape(-1), *args)
        assert np.abs(y).max() < 1e-12 * scale
    assert abs(yr @ Ax - xr @ Ay) < 1e-12 * abs(yr @ Ax)
    assert xr @ Ax > 0
    # assembled CSR == action; the package's pattern builder == the oracle's
    rp, col = of.sparsity_pattern(m["dofmap"], 2, m["n_dofs"])
    rp2, col2 = pkg_forms.cell_sparsity(m["dofmap"], 2, m["n_dofs"])
    assert np.array_equal(rp, rp2) and np.array_equal(col, col2)
    vals = of.assemble_matrix(ot.MANDEL_STRAIN, ot.MANDEL_STRAIN, D, *args, rp, col)
    rows = np.repeat(np.arange(rp.size - 1), np.diff(rp))
    y = np.zeros(2 * m["n_dofs"])
    np.add.at(y, rows, vals * xr[col])
    np.testing.assert_allclose(y, Ax, rtol=0, atol=1e-12 * scale)


def test_tets_and_cell_subset():
    m = tet_case(2)
    w = np.array([0.1, 1.0 / 6.0 - 0.1])
    rng = np.random.default_rng(2)
    nc = m["dofmap"].shape[0]
    s = rng.normal(size=(nc, 2, 9))
    u = rng.normal(size=3 * m["n_dofs"])
    b = of.assemble_vector(ot.GRAD, s, w, m["dofmap"], 3, m["n_dofs"], *_geo(m))
    op = ot.tabulate(ot.GRAD, u, m["dofmap"], 3, *_geo(m))
    _, adet = of.operand_matrix(ot.GRAD, m["dofmap"], 3, *_geo(m))
    lhs = np.einsum("cqk,cqk,q,c->", op, s, w, adet)
    assert abs(lhs - u @ b) <= 1e-12 * abs(lhs)
    b_half = of.assemble_vector(ot.GRAD, s, w, m["dofmap"], 3, m["n_dofs"], *_geo(m), n_cells=nc // 2)
    b_rest = of.assemble_vector(ot.GRAD, s[nc // 2:], w, m["dofmap"][nc // 2:], 3, m["n_dofs"], m["x"],
                                m["x_dofmap"][nc // 2:], m["phi"], m["dphi"], m["dpsi"])
    np.testing.assert_allclose(b_half + b_rest, b, rtol=0, atol=1e-13 * np.abs(b).max())


def test_taylor_remainder_of_residual_vs_tangent_has_slope_two():
    """demo_mc:1203-1235 for the von Mises demo: |F(u + h du) - F(u) - h J(u) du| = O(h^2)."""
    m = tri_case(nx=8, ny=6)
    prm = oc.VonMisesParams()
    nq = m["dofmap"].shape[0] * 3
    rng = np.random.default_rng(5)
    sigma_n = rng.normal(0.0, 30.0, (nq, 4))
    p = np.abs(rng.normal(0.0, 1e-3, nq))
    u = syn.smooth_displacement(m["dof_coords"], scale=2e-2, seed=3).reshape(-1)  # deep in the plastic range
    du = syn.smooth_displacement(m["dof_coords"], scale=2e-2, seed=4).reshape(-1)
    args = (W3, m["dofmap"], 2, m["n_dofs"], *_geo(m))

    def residual(v):
        eps = ot.tabulate(ot.MANDEL_STRAIN, v, m["dofmap"], 2, *_geo(m)).reshape(-1, 4)
        Ct, sig, dp = oc.vm_return_mapping(eps, sigma_n, p, prm)
        return of.assemble_vector(ot.MANDEL_STRAIN, sig, *args), Ct, dp

    F0, Ct, dp = residual(u)
    assert (dp > 0).mean() > 0.9
    Jdu = of.apply_action(ot.MANDEL_STRAIN, ot.MANDEL_STRAIN, Ct, du, *args)
    hs = np.array([1e-2, 5e-3, 2.5e-3, 1.25e-3])
    r0 = np.array([np.linalg.norm(residual(u + h * du)[0] - F0) for h in hs])
    r1 = np.array([np.linalg.norm(residual(u + h * du)[0] - F0 - h * Jdu) for h in hs])
    slope0 = np.polyfit(np.log(hs), np.log(r0), 1)[0]
    slope1 = np.polyfit(np.log(hs), np.log(r1), 1)[0]
    assert abs(slope0 - 1.0) < 0.05 and slope1 > 1.9


# ---------------------------------------------------------------------------------------------------------------
# the kernels' own per-cell algebra (csrc/form_core.cuh, compiled for the host by tests/hostcheck/) against the oracle
@pytest.fixture(scope="module")
def hc():
    import ctypes as C
    import os
    import subprocess

    here = os.path.dirname(os.path.abspath(__file__))
    subprocess.check_call(["make", "-C", os.path.join(here, "hostcheck")], stdout=subprocess.DEVNULL)
    return C.CDLL(os.path.join(here, "hostcheck", "libhostcheck.so"))


def _host_form(hc, m, bs, weights, kind_test, kind_trial, D, x=None):
    import ctypes as C

    p = lambda a: None if a is None else a.ctypes.data_as(C.c_void_p)  # noqa: E731
    gdim, (nq, nb) = m["dphi"].shape[0], m["phi"].shape
    arrs = [np.ascontiguousarray(a, dtype=np.float64) for a in (m["phi"], m["dphi"], m["dpsi"], weights)]
    dm, xd = np.ascontiguousarray(m["dofmap"], dtype=np.int32), np.ascontiguousarray(m["x_dofmap"], dtype=np.int32)
    xg, D = np.ascontiguousarray(m["x"]), np.ascontiguousarray(D, dtype=np.float64)
    xin = None if x is None else np.ascontiguousarray(x, dtype=np.float64)
    y = np.zeros(bs * m["n_dofs"])
    fk = lambda k: ot.GRAD if k == ot.DEF_GRAD else k  # noqa: E731  (form.cu: form_kind)
    rc = hc.hostcheck_form(gdim, bs, nb, nq, fk(kind_test), fk(kind_trial), *(p(a) for a in arrs), p(dm), p(xd), p(xg), p(D),
                           p(xin), C.c_int64(dm.shape[0]), p(y))
    assert rc == 0
    return y


@pytest.mark.parametrize("kind", [ot.VALUE, ot.GRAD, ot.MANDEL_STRAIN, ot.DEF_GRAD])
def test_form_core_vector_against_oracle(hc, kind):
    m = tri_case(nx=9, ny=7)
    s = np.random.default_rng(0).normal(size=(m["dofmap"].shape[0], 3, of.ncomp(kind, 2, 2)))
    got = _host_form(hc, m, 2, W3, kind, 0, s)
    ref = of.assemble_vector(kind, s, W3, m["dofmap"], 2, m["n_dofs"], *_geo(m))
    np.testing.assert_allclose(got, ref, rtol=0, atol=1e-13 * np.abs(ref).max())


@pytest.mark.parametrize("case", [("tri2", ot.MANDEL_STRAIN, ot.MANDEL_STRAIN), ("tri1s", ot.GRAD, ot.VALUE),
                                  ("tri1s", ot.GRAD, ot.GRAD), ("tet", ot.GRAD, ot.DEF_GRAD)])
def test_form_core_action_against_oracle(hc, case):
    name, kt, ki = case
    if name == "tri2":
        m, bs, w = tri_case(nx=9, ny=7), 2, W3
    elif name == "tri1s":
        m, bs, w = tri_case(nx=9, ny=7, degree=1), 1, W3
    else:
        m, bs, w = tet_case(2), 3, np.array([0.1, 1.0 / 6.0 - 0.1])
    gdim = m["dphi"].shape[0]
    rng = np.random.default_rng(1)
    D = rng.normal(size=(m["dofmap"].shape[0], m["phi"].shape[0], of.ncomp(kt, bs, gdim) * of.ncomp(ki, bs, gdim)))
    x = rng.normal(size=bs * m["n_dofs"])
    got = _host_form(hc, m, bs, w, kt, ki, D, x)
    ref = of.apply_action(kt, ki, D, x, w, m["dofmap"], bs, m["n_dofs"], *_geo(m))
    np.testing.assert_allclose(got, ref, rtol=0, atol=1e-13 * np.abs(ref).max())


def test_three_point_rule_is_exact_for_the_p2_stiffness():
    """Known answer independent of the 3-point rule: on affine cells the P2 strain-displacement products are quadratic, so
    the demos' 3-point rule integrates the constant-tangent stiffness exactly - compare with the 6-point degree-4
    Dunavant rule on the same tables machinery (different points, different weights)."""
    a1, b1, w1 = 0.445948490915965, 0.108103018168070, 0.223381589678011
    a2, b2, w2 = 0.091576213509771, 0.816847572980459, 0.109951743655322
    X6 = np.array([[a1, a1], [a1, b1], [b1, a1], [a2, a2], [a2, b2], [b2, a2]])
    W6 = 0.5 * np.array([w1, w1, w1, w2, w2, w2])
    m = tri_case(nx=6, ny=5)
    phi6, dphi6 = el.lagrange_triangle(2, X6)
    Ce = oc.elastic_stiffness(oc.VonMisesParams().lmbda, oc.VonMisesParams().mu)
    nc = m["dofmap"].shape[0]
    K3 = of.element_matrices(ot.MANDEL_STRAIN, ot.MANDEL_STRAIN, np.broadcast_to(Ce.reshape(-1), (nc, 3, 16)), W3, m["dofmap"], 2,
                             m["x"], m["x_dofmap"], m["phi"], m["dphi"], m["dpsi"])
    K6 = of.element_matrices(ot.MANDEL_STRAIN, ot.MANDEL_STRAIN, np.broadcast_to(Ce.reshape(-1), (nc, 6, 16)), W6, m["dofmap"], 2,
                             m["x"], m["x_dofmap"], phi6, dphi6, m["dpsi"])
    np.testing.assert_allclose(K3, K6, rtol=0, atol=1e-12 * np.abs(K6).max())
    # and the mass-like integral of a product of two P2 functions (degree 4) is NOT exact with 3 points: the check above
    # is a statement about the rule, not a tautology of the machinery
    M3 = of.element_matrices(ot.VALUE, ot.VALUE, np.broadcast_to(np.eye(2).reshape(-1), (nc, 3, 4)), W3, m["dofmap"], 2,
                             m["x"], m["x_dofmap"], m["phi"], m["dphi"], m["dpsi"])
    M6 = of.element_matrices(ot.VALUE, ot.VALUE, np.broadcast_to(np.eye(2).reshape(-1), (nc, 6, 4)), W6, m["dofmap"], 2,
                             m["x"], m["x_dofmap"], phi6, dphi6, m["dpsi"])
    assert np.abs(M3 - M6).max() > 1e-3 * np.abs(M6).max()
    assert abs(M6.sum() - 2 * 1.0) < 1e-12  # sum_ij int phi_i phi_j = |domain| per component


def test_c_cell_loop_against_the_numpy_oracles():
    """oracle/csrc/forms_oracle.c (the OpenMP CPU baseline of the tab / fused / step / action bench legs) reproduces the
    NumPy oracles: strain, radial return (bit-exact flags), residual and tangent action."""
    from oracle import native

    m = tri_case(nx=11, ny=8)
    n = m["dofmap"].shape[0] * 3
    rng = np.random.default_rng(3)
    sigma_n, p = rng.normal(0.0, 100.0, (n, 4)), np.abs(rng.normal(0.0, 1e-3, n))
    u = syn.smooth_displacement(m["dof_coords"], scale=6e-4, seed=3).reshape(-1)
    prm = oc.VonMisesParams()
    eps_ref = ot.tabulate(ot.MANDEL_STRAIN, u, m["dofmap"], 2, *_geo(m))
    eps = native.forms_p2_cells("tab", m, W3, u)
    np.testing.assert_allclose(eps, eps_ref, rtol=0, atol=1e-13 * np.abs(eps_ref).max())
    rCt, rsig, rdp = oc.vm_return_mapping(eps_ref.reshape(-1, 4), sigma_n, p, prm)
    b, Ct, sig, dp = native.forms_p2_cells("step", m, W3, u, prm, sigma_n, p)
    assert np.array_equal(dp > 0, np.asarray(rdp).reshape(-1) > 0) and 0.2 < (dp > 0).mean() < 0.8
    np.testing.assert_allclose(Ct, rCt, rtol=0, atol=1e-11 * np.abs(rCt).max())
    np.testing.assert_allclose(sig, rsig, rtol=0, atol=1e-12 * np.abs(rsig).max())
    b_ref = of.assemble_vector(ot.MANDEL_STRAIN, rsig, W3, m["dofmap"], 2, m["n_dofs"], *_geo(m))
    np.testing.assert_allclose(b, b_ref, rtol=0, atol=1e-12 * np.abs(b_ref).max())
    Ct2, sig2, dp2 = native.forms_p2_cells("fused", m, W3, u, prm, sigma_n, p)
    assert np.array_equal(Ct2, Ct) and np.array_equal(sig2, sig) and np.array_equal(dp2, dp)
    x = rng.normal(size=u.size)
    y = native.forms_p2_cells("action", m, W3, x, C_tang=Ct)
    y_ref = of.apply_action(ot.MANDEL_STRAIN, ot.MANDEL_STRAIN, Ct, x, W3, m["dofmap"], 2, m["n_dofs"], *_geo(m))
    np.testing.assert_allclose(y, y_ref, rtol=0, atol=1e-12 * np.abs(y_ref).max())


def test_form_core_p3_against_oracle(hc):
    """nb = 10 (P3 triangles, cell-wise numbering): the kernels' core vs the oracle for vector and action, and the adjoint
    identity with the P3 tabulation."""
    from tab_util import tri_case_discontinuous

    m = tri_case_discontinuous(degree=3)
    rng = np.random.default_rng(4)
    nc = m["dofmap"].shape[0]
    for bs, kt, ki in ((2, ot.MANDEL_STRAIN, ot.MANDEL_STRAIN), (1, ot.GRAD, ot.VALUE)):
        s = rng.normal(size=(nc, 3, of.ncomp(kt, bs, 2)))
        ref = of.assemble_vector(kt, s, W3, m["dofmap"], bs, m["n_dofs"], *_geo(m))
        np.testing.assert_allclose(_host_form(hc, m, bs, W3, kt, 0, s), ref, rtol=0, atol=1e-13 * np.abs(ref).max())
        u = rng.normal(size=bs * m["n_dofs"])
        op = ot.tabulate(kt, u, m["dofmap"], bs, *_geo(m))
        _, adet = of.operand_matrix(kt, m["dofmap"], bs, *_geo(m))
        lhs = np.einsum("cqk,cqk,q,c->", op, s, W3, adet)
        assert abs(lhs - u @ ref) <= 1e-12 * abs(lhs)
        D = rng.normal(size=(nc, 3, of.ncomp(kt, bs, 2) * of.ncomp(ki, bs, 2)))
        yref = of.apply_action(kt, ki, D, u, W3, m["dofmap"], bs, m["n_dofs"], *_geo(m))
        np.testing.assert_allclose(_host_form(hc, m, bs, W3, kt, ki, D, u), yref, rtol=0, atol=1e-13 * np.abs(yref).max())


def test_form_core_p2_tetrahedra_against_oracle(hc):
    """nb = 10, gdim = 3 (P2 tetrahedra): P : grad(v) and the dP/dF action of a 3-d hyperelastic residual / Jacobian."""
    from tab_util import tet_case_discontinuous

    m = tet_case_discontinuous()
    w = np.array([0.05, 0.04, 0.03, 1.0 / 6.0 - 0.12])
    rng = np.random.default_rng(5)
    nc, nq = m["dofmap"].shape[0], m["phi"].shape[0]
    P = rng.normal(size=(nc, nq, 9))
    ref = of.assemble_vector(ot.GRAD, P, w, m["dofmap"], 3, m["n_dofs"], *_geo(m))
    np.testing.assert_allclose(_host_form(hc, m, 3, w, ot.GRAD, 0, P), ref, rtol=0, atol=1e-13 * np.abs(ref).max())
    D = rng.normal(size=(nc, nq, 81))
    x = rng.normal(size=3 * m["n_dofs"])
    yref = of.apply_action(ot.GRAD, ot.DEF_GRAD, D, x, w, m["dofmap"], 3, m["n_dofs"], *_geo(m))
    np.testing.assert_allclose(_host_form(hc, m, 3, w, ot.GRAD, ot.DEF_GRAD, D, x), yref, rtol=0, atol=1e-13 * np.abs(yref).max())
