"""The 'pure UFL' side of the reference's own check of the heat demo, restated by hand for P1 triangles
(demo_nonlinear_heat_equation_part2.py:283-300): F_explicit = inner(-k(T) grad T, grad v) dx and
J_manual = inner(B k^2 grad T  T_hat, grad v) dx + inner(-k I grad T_hat, grad v) dx, integrated with the 3-point rule.
Written from the triangle's vertex coordinates (closed-form barycentric gradients), i.e. independently of the operand
matrices of oracle/forms.py and of the kernels' tables."""

import numpy as np


def explicit_heat_forms(m, T, A=1.0, B=1.0):
    """m: P1 triangle mesh dict (x, x_dofmap == dofmap), T nodal values.  Returns (b (n,), A_dense (n, n))."""
    X = np.array([[1 / 6, 1 / 6], [1 / 6, 2 / 3], [2 / 3, 1 / 6]])
    w = np.full(3, 1 / 6)
    n = m["n_dofs"]
    b, Amat = np.zeros(n), np.zeros((n, n))
    for cell in m["dofmap"]:
        (x0, y0), (x1, y1), (x2, y2) = m["x"][cell, :2]
        det = (x1 - x0) * (y2 - y0) - (x2 - x0) * (y1 - y0)
        g = np.array([[y1 - y2, x2 - x1], [y2 - y0, x0 - x2], [y0 - y1, x1 - x0]]) / det  # grad of the barycentric functions
        sigma = T[cell] @ g
        for (xi, eta), wq in zip(X, w):
            phi = np.array([1 - xi - eta, xi, eta])
            k = 1.0 / (A + B * (phi @ T[cell]))
            q = -k * sigma
            dx = wq * abs(det)
            b[cell] += dx * (g @ q)
            Amat[np.ix_(cell, cell)] += dx * (np.outer(g @ (B * k * k * sigma), phi) - k * (g @ g.T))
    return b, Amat
