"""Shared fixtures for the tabulation tests: small synthetic meshes with consistent tables."""

import numpy as np

from dolfinx_external_operator_b200 import elements as el
from dolfinx_external_operator_b200 import synthetic as syn


def tri_case(nx=9, ny=7, degree=2, qdeg=2, jitter=0.3, seed=1):
    m = syn.triangle_mesh(nx, ny, degree, jitter=jitter, seed=seed)
    X = el.triangle_quadrature(qdeg)
    phi, dphi = el.lagrange_triangle(degree, X)
    m.update(phi=phi, dphi=dphi, dpsi=el.p1_geometry_derivatives(2), X=X)
    P1phi, _ = el.lagrange_triangle(1, X)
    m["xq"] = np.einsum("cvi,qv->cqi", m["x"][m["x_dofmap"]][:, :, :2], P1phi)  # physical evaluation points
    return m


def tet_case(n=3, seed=0):
    """A few P1 tetrahedra: the 6-tet split of each cube of an n^3 grid (jittered)."""
    rng = np.random.default_rng(seed)
    g = np.arange(n + 1)
    Z, Y, Xc = np.meshgrid(g, g, g, indexing="ij")
    x = np.stack([Xc.reshape(-1), Y.reshape(-1), Z.reshape(-1)], 1).astype(float) / n
    x += 0.1 / n * rng.uniform(-1, 1, x.shape)
    vid = lambda i, j, k: (k * (n + 1) + j) * (n + 1) + i  # noqa: E731
    cells = []
    for k in range(n):
        for j in range(n):
            for i in range(n):
                v = [vid(i + a, j + b, k + c) for c in (0, 1) for b in (0, 1) for a in (0, 1)]
                for t in ((0, 1, 3, 7), (0, 1, 5, 7), (0, 2, 3, 7), (0, 2, 6, 7), (0, 4, 5, 7), (0, 4, 6, 7)):
                    cells.append([v[t[0]], v[t[1]], v[t[2]], v[t[3]]])
    xd = np.array(cells, dtype=np.int32)
    X = np.array([[0.25, 0.25, 0.25], [0.1, 0.2, 0.3]])
    phi, dphi = el.lagrange_tetrahedron(1, X)
    return {"x": x, "x_dofmap": xd, "dofmap": xd.copy(), "n_dofs": x.shape[0], "phi": phi, "dphi": dphi,
            "dpsi": el.p1_geometry_derivatives(3), "X": X, "dof_coords": x}
