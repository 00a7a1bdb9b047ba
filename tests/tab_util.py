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


def tri_case_discontinuous(degree=3, nx=5, ny=4, jitter=0.3, seed=1, qdeg=2):
    """Jittered triangles with a cell-wise (discontinuous) numbering of a degree-`degree` Lagrange space: every cell owns
    its nb nodes, dofmap = arange.  Enough to exercise gather / contraction / scatter of elements the structured
    generators do not number globally (P3: nb = 10)."""
    g = syn.triangle_mesh(nx, ny, 1, jitter=jitter, seed=seed)
    X = el.triangle_quadrature(qdeg)
    phi, dphi = el.lagrange_triangle(degree, X)
    nodes = el.lagrange_triangle_nodes(degree)
    P1n, _ = el.lagrange_triangle(1, nodes)  # affine map of the reference nodes
    nc, nb = g["x_dofmap"].shape[0], nodes.shape[0]
    xv = g["x"][g["x_dofmap"]][:, :, :2]
    dof_coords = np.einsum("cvi,av->cai", xv, P1n).reshape(nc * nb, 2)
    P1phi, _ = el.lagrange_triangle(1, X)
    return {"x": g["x"], "x_dofmap": g["x_dofmap"], "dofmap": np.arange(nc * nb, dtype=np.int32).reshape(nc, nb),
            "n_dofs": nc * nb, "dof_coords": dof_coords, "phi": phi, "dphi": dphi, "dpsi": el.p1_geometry_derivatives(2),
            "X": X, "xq": np.einsum("cvi,qv->cqi", xv, P1phi)}


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


def tet_case_discontinuous(degree=2, n=2, seed=0):
    """Jittered tetrahedra (the 6-tet cube split of `tet_case`) with a cell-wise numbering of a degree-`degree` Lagrange
    space (P2: nb = 10)."""
    g = tet_case(n, seed)
    X = np.array([[0.25, 0.25, 0.25], [0.1, 0.2, 0.3], [0.5, 0.2, 0.1], [0.15, 0.6, 0.2]])
    phi, dphi = el.lagrange_tetrahedron(degree, X)
    nodes = el.lagrange_tetrahedron_nodes(degree)
    P1n, _ = el.lagrange_tetrahedron(1, nodes)
    nc, nb = g["x_dofmap"].shape[0], nodes.shape[0]
    xv = g["x"][g["x_dofmap"]]
    P1phi, _ = el.lagrange_tetrahedron(1, X)
    return {"x": g["x"], "x_dofmap": g["x_dofmap"], "dofmap": np.arange(nc * nb, dtype=np.int32).reshape(nc, nb),
            "n_dofs": nc * nb, "dof_coords": np.einsum("cvi,av->cai", xv, P1n).reshape(nc * nb, 3), "phi": phi, "dphi": dphi,
            "dpsi": el.p1_geometry_derivatives(3), "X": X, "xq": np.einsum("cvi,qv->cqi", xv, P1phi)}


def general_case(cell: str, degree: int, bs: int, facets: bool = False, seed: int = 0):
    """Synthetic mesh + consistent table sets for the general tabulation path.  Returns the mesh dict extended with
    phi / dphi / dgeo (n_sets, ...), the reference points X, the cell name and physical evaluation points `xq`
    (n_sets, n_cells, nq, gdim)."""
    if cell == "triangle":
        m = syn.triangle_mesh(7, 5, degree, jitter=0.3, seed=seed)
        Xc = el.triangle_quadrature(2)
        Xf = np.array([[0.2113248654051871], [0.7886751345948129]])
    elif cell == "quadrilateral":
        m = syn.quad_mesh(6, 5, degree, jitter=0.3, seed=seed)
        g = np.array([0.2113248654051871, 0.7886751345948129])
        Xc = np.array([[a, b] for a in g for b in g])
        Xf = g[:, None]
    elif cell == "hexahedron":
        m = syn.hex_mesh(3, degree, jitter=0.25, seed=seed)
        g = np.array([0.2113248654051871, 0.7886751345948129])
        Xc = np.array([[a, b, c] for a in g for b in g for c in g])
        Xf = np.array([[a, b] for a in g for b in g])
    elif cell == "tetrahedron":
        m = tet_case(3, seed)
        Xc = np.array([[0.25, 0.25, 0.25], [0.1, 0.2, 0.3], [0.5, 0.2, 0.1]])
        Xf = np.array([[1 / 3, 1 / 3], [0.2, 0.6]])
    else:
        raise ValueError(cell)
    X = Xf if facets else Xc
    phi, dphi, dgeo = el.table_sets(cell, degree, 1, X, facets=facets)
    gphi = np.stack([el.tabulate_on(cell, 1, Xs)[0] for Xs in (el.facet_points(cell, X) if facets else X[None])])
    gdim = dphi.shape[1]
    m = dict(m)
    m.update(phi=phi, dphi=dphi, dgeo=dgeo, X=X, cell=cell, bs=bs, gdim=gdim,
             xq=np.einsum("cvi,sqv->scqi", m["x"][m["x_dofmap"]][:, :, :gdim], gphi))
    return m
