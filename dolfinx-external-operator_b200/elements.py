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


def lagrange_triangle(degree: int, X: np.ndarray):
    """(phi (nq, nb), dphi (2, nq, nb)) of P1 / P2 on the reference triangle at points X (nq, 2)."""
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
    else:
        raise NotImplementedError("closed-form tables for P1 and P2 only; pass basix tables for higher degrees")
    return np.ascontiguousarray(phi), np.ascontiguousarray(np.stack([dx, dy]))


def lagrange_tetrahedron(degree: int, X: np.ndarray):
    """(phi (nq, 4), dphi (3, nq, 4)) of P1 on the reference tetrahedron."""
    if degree != 1:
        raise NotImplementedError("closed-form tetrahedron tables for P1 only; pass basix tables")
    x, y, z = X[:, 0], X[:, 1], X[:, 2]
    one, zero = np.ones_like(x), np.zeros_like(x)
    phi = np.stack([1 - x - y - z, x, y, z], axis=1)
    d = np.stack([np.stack([-one, one, zero, zero], 1), np.stack([-one, zero, one, zero], 1),
                  np.stack([-one, zero, zero, one], 1)])
    return np.ascontiguousarray(phi), np.ascontiguousarray(d)


def p1_geometry_derivatives(gdim: int) -> np.ndarray:
    """d psi_v / d X_k of the affine geometry element, shape (gdim, gdim + 1)."""
    d = np.zeros((gdim, gdim + 1))
    d[:, 0] = -1.0
    for k in range(gdim):
        d[k, k + 1] = 1.0
    return d
