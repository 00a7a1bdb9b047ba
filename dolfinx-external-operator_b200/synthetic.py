"""Seeded synthetic quadrature-point batches.
Shared by the golden generator, the parity tests and bench.py so that every leg sees the
same inputs for a given (kind, n, seed)."""

from __future__ import annotations

import numpy as np


def vm_batch(n: int, seed: int = 0, kind: str = "mixed"):
    """deps (n,4) with eps_zz = 0, sigma_n (n,4), p (n,).  'mixed' gives ~53 % plastic points,
    'elastic' scales strain and stress by 0.01, 'plastic' by 3."""
    rng = np.random.default_rng(seed)
    scale = {"mixed": 1.0, "elastic": 0.01, "plastic": 3.0}[kind]
    deps = rng.normal(0.0, 2e-3, (n, 4)) * scale
    deps[:, 2] = 0.0
    sigma_n = rng.normal(0.0, 100.0, (n, 4)) * scale
    p = np.abs(rng.normal(0.0, 1e-3, n))
    return deps, sigma_n, p


def heat_batch(n: int, seed: int = 0):
    """T ~ U(0,2) (so A + B T > 0), sigma = grad T ~ N(0,1)^2."""
    rng = np.random.default_rng(seed)
    return rng.uniform(0.0, 2.0, n), rng.normal(0.0, 1.0, (n, 2))


def isihara_batch(n: int, seed: int = 0):
    """Deformation gradients F = I + N(0, 0.05) (det F > 0), flat [F11, F12, F21, F22]."""
    rng = np.random.default_rng(seed)
    F = rng.normal(0.0, 0.05, (n, 4))
    F[:, 0] += 1.0
    F[:, 3] += 1.0
    return F


# ----------------------------------------------------------------------------- Mohr-Coulomb
MC_E, MC_NU = 6778.0, 0.25  # demo_plasticity_mohr_coulomb.py:110-111


def mc_compliance(E: float = MC_E, nu: float = MC_NU) -> np.ndarray:
    """S_elas = inv(C_elas) of demo_mc:405-416 in closed form (plane-strain Mandel 4x4)."""
    S = np.zeros((4, 4))
    S[:3, :3] = -nu / E
    S[0, 0] = S[1, 1] = S[2, 2] = 1.0 / E
    S[3, 3] = (1.0 + nu) / E  # 1/(2 mu)
    return S


def mc_rotate(v: np.ndarray, omega: np.ndarray) -> np.ndarray:
    """Rotate plane-strain Mandel vectors [xx, yy, zz, sqrt2 xy] by the in-plane angle omega
    (the model is isotropic, so the return mapping commutes with this rotation; it only serves to
    populate the shear component, which the demo's principal-axes stress paths leave at zero)."""
    c, s = np.cos(omega), np.sin(omega)
    xx, yy, zz, xy = v[:, 0], v[:, 1], v[:, 2], v[:, 3] / np.sqrt(2.0)
    out = np.empty_like(v)
    out[:, 0] = c * c * xx + s * s * yy - 2.0 * s * c * xy
    out[:, 1] = s * s * xx + c * c * yy + 2.0 * s * c * xy
    out[:, 2] = zz
    out[:, 3] = np.sqrt(2.0) * (s * c * (xx - yy) + (c * c - s * s) * xy)
    return out


def mc_path_increments(theta: np.ndarray, R) -> np.ndarray:
    """Stress-path increment in principal Haigh-Westergaard coordinates, demo_mc:868-871."""
    d = np.zeros((theta.size, 4))
    d[:, 0] = (R / np.sqrt(2)) * (np.cos(theta) + np.sin(theta) / np.sqrt(3))
    d[:, 1] = (R / np.sqrt(2)) * (-2 * np.sin(theta) / np.sqrt(3))
    d[:, 2] = (R / np.sqrt(2)) * (np.sin(theta) / np.sqrt(3) - np.cos(theta))
    return d


def mc_batch(n: int, seed: int = 0, stepper=None, max_level: int = 16, rotate: bool = True):
    """The demo's yield-surface tracing family (demo_mc:853-930), randomised:
    Lode angle theta ~ U(-pi/6+1e-5, pi/6-1e-5), path radius R ~ U(0.1, 0.7), hydrostatic offset
    p ~ U(-1, 1), load level k ~ U{0..max_level} (16 -> ~34 % plastic points, 2-5 Newton iterations).  deps = S_elas @ dsigma_path (:903); sigma_n is the
    stress reached after k increments of the path, each followed by the re-projection onto the
    deviatoric plane of :921-923.

    `stepper(deps (m,4), sigma_n (m,4)) -> sigma (m,4)` performs one stress update; the tests pass
    the CPU oracle, bench.py passes the GPU kernel itself.  Returns deps (n,4), sigma_n (n,4)."""
    if stepper is None:
        raise ValueError("mc_batch needs a `stepper` (stress update) to walk the stress paths")
    rng = np.random.default_rng(seed)
    eps = 1e-5
    theta = rng.uniform(-np.pi / 6 + eps, np.pi / 6 - eps, n)
    R = rng.uniform(0.1, 0.7, n)
    p = rng.uniform(-1.0, 1.0, n)
    level = rng.integers(0, max_level + 1, n)
    omega = rng.uniform(0.0, np.pi, n) if rotate else np.zeros(n)
    dsig = mc_path_increments(theta, R)
    deps = dsig @ mc_compliance().T
    sigma_n = np.zeros((n, 4))
    sigma_n[:, :3] = p[:, None]
    tr = np.array([1.0, 1.0, 1.0, 0.0])
    for k in range(max_level):
        active = level > k
        if not active.any():
            break
        sig = np.asarray(stepper(deps[active], sigma_n[active])).reshape(-1, 4)
        dpp = sig @ tr / 3.0 - p[active]
        sigma_n[active] = sig - np.outer(dpp, tr)
    if rotate:
        deps, sigma_n = mc_rotate(deps, omega), mc_rotate(sigma_n, omega)
    return np.ascontiguousarray(deps), np.ascontiguousarray(sigma_n)


def mc_demo_path(n_angles: int = 50, n_loads: int = 9, R: float = 0.7, p: float = 0.1, stepper=None):
    """The demo's tracing driver itself (demo_mc:853-930): returns the (deps, sigma_n) pairs of every
    (load level, angle), shapes (n_loads*n_angles, 4)."""
    eps = 1e-5
    theta = np.linspace(-np.pi / 6 + eps, np.pi / 6 - eps, n_angles)
    dsig = mc_path_increments(theta, R)
    deps = dsig @ mc_compliance().T
    sigma_n = np.zeros((n_angles, 4))
    sigma_n[:, :3] = p
    tr = np.array([1.0, 1.0, 1.0, 0.0])
    D, S = [], []
    for _ in range(n_loads):
        D.append(deps.copy())
        S.append(sigma_n.copy())
        sig = np.asarray(stepper(deps, sigma_n)).reshape(-1, 4)
        dpp = sig @ tr / 3.0 - p
        sigma_n = sig - np.outer(dpp, tr)
    return np.concatenate(D), np.concatenate(S)


# ----------------------------------------------------------------------------- structured meshes
def triangle_mesh(nx: int, ny: int, degree: int = 2, lx: float = 1.0, ly: float = 1.0, jitter: float = 0.0,
                  seed: int = 0):
    """Structured triangulation of [0,lx]x[0,ly] (each grid square split along its (i,j)-(i+1,j+1) diagonal),
    numbered arithmetically so that meshes with 10^7-10^8 cells are generated in seconds.

    Returns dict(x (n_nodes,3), x_dofmap (n_cells,3) int32, dofmap (n_cells,nb) int32, n_dofs, dof_coords
    (n_dofs,2)).  P2 dofs: vertices first, then horizontal, vertical and diagonal edge midpoints; the local
    order is vertices then edges, edge0=(v1,v2), edge1=(v0,v2), edge2=(v0,v1).  `jitter` displaces interior
    vertices by a fraction of the grid spacing so that the cells are not all congruent."""
    nvx, nvy = nx + 1, ny + 1
    i, j = np.meshgrid(np.arange(nx, dtype=np.int64), np.arange(ny, dtype=np.int64), indexing="xy")
    i, j = i.reshape(-1), j.reshape(-1)
    v00, v10, v01, v11 = j * nvx + i, j * nvx + i + 1, (j + 1) * nvx + i, (j + 1) * nvx + i + 1
    X, Y = np.meshgrid(np.linspace(0.0, lx, nvx), np.linspace(0.0, ly, nvy), indexing="xy")
    if jitter > 0.0:
        rng = np.random.default_rng(seed)
        dx, dy = lx / nx, ly / ny
        X[1:-1, 1:-1] += jitter * dx * rng.uniform(-1, 1, (nvy - 2, nvx - 2))
        Y[1:-1, 1:-1] += jitter * dy * rng.uniform(-1, 1, (nvy - 2, nvx - 2))
    x = np.zeros((nvx * nvy, 3))
    x[:, 0], x[:, 1] = X.reshape(-1), Y.reshape(-1)
    n_cells = 2 * nx * ny
    x_dofmap = np.empty((n_cells, 3), dtype=np.int32)
    x_dofmap[0::2] = np.stack([v00, v10, v11], axis=1)
    x_dofmap[1::2] = np.stack([v00, v11, v01], axis=1)
    if degree == 1:
        return {"x": x, "x_dofmap": x_dofmap, "dofmap": x_dofmap.copy(), "n_dofs": nvx * nvy, "dof_coords": x[:, :2].copy()}
    if degree != 2:
        raise NotImplementedError
    nv = nvx * nvy
    nH, nV = nx * nvy, nvx * ny
    H = lambda ii, jj: nv + jj * nx + ii  # noqa: E731  edge (ii,jj)-(ii+1,jj)
    V = lambda ii, jj: nv + nH + jj * nvx + ii  # noqa: E731  edge (ii,jj)-(ii,jj+1)
    D = lambda ii, jj: nv + nH + nV + jj * nx + ii  # noqa: E731  edge (ii,jj)-(ii+1,jj+1)
    dofmap = np.empty((n_cells, 6), dtype=np.int32)
    # T0 = (v00, v10, v11): edge0=(v10,v11)=V(i+1,j), edge1=(v00,v11)=D(i,j), edge2=(v00,v10)=H(i,j)
    dofmap[0::2] = np.stack([v00, v10, v11, V(i + 1, j), D(i, j), H(i, j)], axis=1)
    # T1 = (v00, v11, v01): edge0=(v11,v01)=H(i,j+1), edge1=(v00,v01)=V(i,j), edge2=(v00,v11)=D(i,j)
    dofmap[1::2] = np.stack([v00, v11, v01, H(i, j + 1), V(i, j), D(i, j)], axis=1)
    n_dofs = nv + nH + nV + nx * ny
    dof_coords = np.empty((n_dofs, 2))
    dof_coords[:nv] = x[:, :2]
    xy = x[:, :2]
    dof_coords[H(i, j)] = 0.5 * (xy[v00] + xy[v10])
    dof_coords[H(i, j + 1)] = 0.5 * (xy[v01] + xy[v11])
    dof_coords[V(i, j)] = 0.5 * (xy[v00] + xy[v01])
    dof_coords[V(i + 1, j)] = 0.5 * (xy[v10] + xy[v11])
    dof_coords[D(i, j)] = 0.5 * (xy[v00] + xy[v11])
    return {"x": x, "x_dofmap": x_dofmap, "dofmap": dofmap, "n_dofs": int(n_dofs), "dof_coords": dof_coords}


def renumber(mesh: dict, order: str = "shuffled", seed: int = 0) -> dict:
    """The same mesh under another numbering of cells, dofs and geometry nodes - the structured generators number
    everything row-major, which is the best case for the per-cell gathers of the tabulation / form kernels.

    order = "shuffled": independent random permutations of cells, dofs and nodes (worst case: no two neighbouring cells
            share a cache line);
            "rcm": reverse Cuthill-McKee on the dof connectivity of a SHUFFLED mesh (nothing of the structured order
            survives), cells sorted by their lowest dof - the locality an unstructured (gmsh) mesh has after the graph
            reordering DOLFINx applies to dofs and cells when it builds a mesh / function space.
            "morton": dofs, nodes and cells sorted along a Z-order space-filling curve of their coordinates (needs
            `dof_coords`; 2-d) - the locality of a geometric partitioner / reordering, computed in O(n log n) so that
            10^7-cell meshes can be renumbered for the bench.
    Returns a new dict with permuted x, x_dofmap, dofmap, dof_coords, plus `dof_new` (old dof -> new dof), `node_new`
    and `cell_old` (new cell k is old cell cell_old[k]): per-cell results are the old ones in the order cell_old, bit
    for bit; a dof vector maps as u_new[dof_new] = u_old (rows of bs components)."""
    rng = np.random.default_rng(seed)
    dofmap, x_dofmap, x = np.asarray(mesh["dofmap"]), np.asarray(mesh["x_dofmap"]), np.asarray(mesh["x"])
    nc, n_dofs, n_nodes = dofmap.shape[0], int(mesh["n_dofs"]), x.shape[0]
    dof_new, node_new, cell_old = rng.permutation(n_dofs), rng.permutation(n_nodes), rng.permutation(nc)
    if order == "rcm":
        from scipy.sparse import coo_matrix
        from scipy.sparse.csgraph import reverse_cuthill_mckee

        def rcm(conn, n, pre):
            c = pre[conn].astype(np.int64)  # shuffled labels
            k = c.shape[1]
            rows, cols = np.repeat(c, k, axis=1).reshape(-1), np.tile(c, (1, k)).reshape(-1)
            A = coo_matrix((np.ones(rows.size, dtype=np.int8), (rows, cols)), shape=(n, n)).tocsr()
            perm = reverse_cuthill_mckee(A, symmetric_mode=True)  # position k holds shuffled label perm[k]
            inv = np.empty(n, dtype=np.int64)
            inv[perm] = np.arange(n)
            return inv[pre]  # old label -> new label

        dof_new = rcm(dofmap, n_dofs, dof_new)
        node_new = rcm(x_dofmap, n_nodes, node_new)
        cell_old = np.argsort(dof_new[dofmap].min(axis=1), kind="stable")
    elif order == "morton":
        def zkey(xy):
            lo, hi = xy.min(axis=0), xy.max(axis=0)
            q = np.minimum(((xy - lo) / np.maximum(hi - lo, 1e-300) * 65535.0).astype(np.uint64), 65535)

            def spread(v):  # 16 bits -> every second bit of 32
                v = (v | (v << 8)) & np.uint64(0x00FF00FF)
                v = (v | (v << 4)) & np.uint64(0x0F0F0F0F)
                v = (v | (v << 2)) & np.uint64(0x33333333)
                return (v | (v << 1)) & np.uint64(0x55555555)

            return spread(q[:, 0]) | (spread(q[:, 1]) << np.uint64(1))

        def rank(keys):
            order_ = np.argsort(keys, kind="stable")
            r = np.empty(order_.size, dtype=np.int64)
            r[order_] = np.arange(order_.size)
            return r

        dof_new = rank(zkey(np.asarray(mesh["dof_coords"])[:, :2]))
        node_new = rank(zkey(x[:, :2]))
        cell_old = np.argsort(zkey(x[x_dofmap][:, :, :2].mean(axis=1)), kind="stable")
    elif order != "shuffled":
        raise ValueError(f"unknown numbering {order!r}")
    out = dict(mesh)
    out["dofmap"] = np.ascontiguousarray(dof_new[dofmap][cell_old], dtype=np.int32)
    out["x_dofmap"] = np.ascontiguousarray(node_new[x_dofmap][cell_old], dtype=np.int32)
    xn = np.empty_like(x)
    xn[node_new] = x
    out["x"] = xn
    if "dof_coords" in mesh:
        dc = np.empty_like(mesh["dof_coords"])
        dc[dof_new] = mesh["dof_coords"]
        out["dof_coords"] = dc
    if "xq" in mesh:
        out["xq"] = np.asarray(mesh["xq"])[cell_old]
    out.update(dof_new=dof_new, node_new=node_new, cell_old=cell_old, order=order)
    return out


def smooth_displacement(dof_coords: np.ndarray, scale: float = 1e-3, seed: int = 0) -> np.ndarray:
    """A smooth random vector field sampled at the dof coordinates, blocked layout [node][comp]."""
    rng = np.random.default_rng(seed)
    a = rng.normal(0.0, 1.0, (2, 6))
    x, y = dof_coords[:, 0], dof_coords[:, 1]
    u = np.empty((dof_coords.shape[0], 2))
    for c in range(2):
        u[:, c] = scale * (a[c, 0] * x + a[c, 1] * y + a[c, 2] * np.sin(3 * x + a[c, 3]) * np.cos(2 * y) + a[c, 4] * x * y
                           + a[c, 5] * y * y)
    return u


def quad_mesh(nx: int, ny: int, degree: int = 1, jitter: float = 0.0, seed: int = 0):
    """Structured quadrilateral mesh of the unit square with Q1 (bilinear, non-affine when jittered) geometry and a
    Q`degree` Lagrange space, all in TENSOR ordering (elements.lagrange_quadrilateral): local node (i, j) -> i*(d+1)+j.
    Returns dict(x, x_dofmap (n_cells, 4), dofmap (n_cells, (d+1)^2), n_dofs, dof_coords)."""
    nvx, nvy = nx + 1, ny + 1
    X, Y = np.meshgrid(np.linspace(0.0, 1.0, nvx), np.linspace(0.0, 1.0, nvy), indexing="ij")  # [i][j]
    if jitter > 0.0:
        rng = np.random.default_rng(seed)
        X[1:-1, 1:-1] += jitter / nx * rng.uniform(-1, 1, (nvx - 2, nvy - 2))
        Y[1:-1, 1:-1] += jitter / ny * rng.uniform(-1, 1, (nvx - 2, nvy - 2))
    x = np.zeros((nvx * nvy, 3))
    x[:, 0], x[:, 1] = X.reshape(-1), Y.reshape(-1)
    ci, cj = np.meshgrid(np.arange(nx), np.arange(ny), indexing="ij")
    ci, cj = ci.reshape(-1), cj.reshape(-1)
    vid = lambda i, j: i * nvy + j  # noqa: E731
    x_dofmap = np.stack([vid(ci + a, cj + b) for a in (0, 1) for b in (0, 1)], axis=1).astype(np.int32)
    d = degree
    ndx, ndy = d * nx + 1, d * ny + 1
    did = lambda i, j: i * ndy + j  # noqa: E731
    dofmap = np.stack([did(d * ci + a, d * cj + b) for a in range(d + 1) for b in range(d + 1)], axis=1).astype(np.int32)
    # dof coordinates: the bilinear geometry map of each cell evaluated at the local node positions
    from . import elements as el

    t = np.linspace(0.0, 1.0, d + 1)
    Xn = np.array([[a, b] for a in t for b in t])
    gphi, _ = el.lagrange_quadrilateral(1, Xn)  # (nb, 4)
    dof_coords = np.zeros((ndx * ndy, 2))
    dof_coords[dofmap] = np.einsum("cvi,av->cai", x[x_dofmap][:, :, :2], gphi)
    return {"x": x, "x_dofmap": x_dofmap, "dofmap": dofmap, "n_dofs": ndx * ndy, "dof_coords": dof_coords}


def hex_mesh(n: int, degree: int = 1, jitter: float = 0.0, seed: int = 0):
    """Structured hexahedral mesh of the unit cube, Q1 geometry (trilinear), Q`degree` space, tensor ordering
    (i, j, k) -> (i*(d+1) + j)*(d+1) + k."""
    nv = n + 1
    g = np.linspace(0.0, 1.0, nv)
    X, Y, Z = np.meshgrid(g, g, g, indexing="ij")
    if jitter > 0.0:
        rng = np.random.default_rng(seed)
        for A in (X, Y, Z):
            A[1:-1, 1:-1, 1:-1] += jitter / n * rng.uniform(-1, 1, (nv - 2,) * 3)
    x = np.stack([X.reshape(-1), Y.reshape(-1), Z.reshape(-1)], axis=1)
    ci, cj, ck = (a.reshape(-1) for a in np.meshgrid(np.arange(n), np.arange(n), np.arange(n), indexing="ij"))
    vid = lambda i, j, k: (i * nv + j) * nv + k  # noqa: E731
    x_dofmap = np.stack([vid(ci + a, cj + b, ck + c) for a in (0, 1) for b in (0, 1) for c in (0, 1)], axis=1).astype(np.int32)
    d = degree
    nd = d * n + 1
    did = lambda i, j, k: (i * nd + j) * nd + k  # noqa: E731
    dofmap = np.stack([did(d * ci + a, d * cj + b, d * ck + c) for a in range(d + 1) for b in range(d + 1)
                       for c in range(d + 1)], axis=1).astype(np.int32)
    from . import elements as el

    t = np.linspace(0.0, 1.0, d + 1)
    Xn = np.array([[a, b, c] for a in t for b in t for c in t])
    gphi, _ = el.lagrange_hexahedron(1, Xn)
    dof_coords = np.zeros((nd**3, 3))
    dof_coords[dofmap] = np.einsum("cvi,av->cai", x[x_dofmap], gphi)
    return {"x": x, "x_dofmap": x_dofmap, "dofmap": dofmap, "n_dofs": nd**3, "dof_coords": dof_coords}
