"""The reference's own known-answer check of assembled operator forms (demo_nonlinear_heat_equation_part2.py:283-300:
np.allclose of the external-operator residual / Jacobian with their pure-UFL counterparts; same structure as
test_external_operators_evaluation.py:40-45), here for the forms oracle: b = int q . grad(v), A = d b / d T through
dq/dT (test grad x trial value) and dq/dsigma (test grad x trial grad) against the hand-written explicit forms."""

import numpy as np

from dolfinx_external_operator_b200 import elements as el
from heat_util import explicit_heat_forms
from oracle import constitutive as oc
from oracle import forms as of
from oracle import tabulation as ot
from tab_util import tri_case


def heat_case():
    m = tri_case(nx=10, ny=10, degree=1, jitter=0.25, seed=2)  # part2.py:126 uses the 10 x 10 unit square, P1
    xy = m["dof_coords"]
    T = xy[:, 0] ** 2 + xy[:, 1]  # part2.py:146
    return m, T


def oracle_heat_forms(m, T):
    geo = (m["x"], m["x_dofmap"], m["phi"], m["dphi"], m["dpsi"])
    W3 = el.triangle_quadrature_weights(2)
    nc = m["dofmap"].shape[0]
    Tq = ot.tabulate(ot.VALUE, T, m["dofmap"], 1, *geo).reshape(nc, 3)  # (n_cells, n_pts), part2.py:219
    sq = ot.tabulate(ot.GRAD, T, m["dofmap"], 1, *geo).reshape(nc, 6)
    q, dT, ds = oc.heat_q(Tq, sq), oc.heat_dqdT(Tq, sq), oc.heat_dqdsigma(Tq, sq)
    args = (W3, m["dofmap"], 1, m["n_dofs"], *geo)
    b = of.assemble_vector(ot.GRAD, q, *args)
    rp, col = of.sparsity_pattern(m["dofmap"], 1, m["n_dofs"])
    vals = of.assemble_matrix(ot.GRAD, ot.VALUE, dT, *args, rp, col) + of.assemble_matrix(ot.GRAD, ot.GRAD, ds, *args, rp, col)
    return b, vals, rp, col


def dense(vals, rp, col):
    n = rp.size - 1
    A = np.zeros((n, n))
    A[np.repeat(np.arange(n), np.diff(rp)), col] = vals
    return A


def test_heat_residual_and_jacobian_match_the_explicit_forms():
    m, T = heat_case()
    b, vals, rp, col = oracle_heat_forms(m, T)
    b_ex, A_ex = explicit_heat_forms(m, T)
    np.testing.assert_allclose(b, b_ex, rtol=0, atol=1e-13 * np.abs(b_ex).max())
    np.testing.assert_allclose(dense(vals, rp, col), A_ex, rtol=0, atol=1e-13 * np.abs(A_ex).max())
    # the Jacobian is the derivative of the residual (central differences)
    rng = np.random.default_rng(0)
    dT, h = rng.normal(size=T.size), 1e-6
    fd = (explicit_heat_forms(m, T + h * dT)[0] - explicit_heat_forms(m, T - h * dT)[0]) / (2 * h)
    np.testing.assert_allclose(A_ex @ dT, fd, rtol=0, atol=1e-8 * np.abs(fd).max())
