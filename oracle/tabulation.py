"""TEST INFRASTRUCTURE - NumPy restatement (the oracle) of the operand tabulation,
i.e. of what `fem.Expression(operand, eval_points).eval(mesh, entities)` computes inside
`evaluate_operands` (src/dolfinx_external_operator/external_operator.py:386-402) for the operand
expressions of the reference demos.

PARITY UNPINNED: the arithmetic lives in un-vendored third-party code (fenics-dolfinx >=0.10,<0.11 with
FFCx and basix, pyproject.toml:15-16) that is not installable here, and the reference's own tests at this
boundary (test_operands_evaluation.py:65-66) compare DOLFINx with itself.  This is the textbook
affine-simplex algorithm; it is checked against analytic fields - the ones the
reference's tests use (u = (0.1x, 0.3y): test_operands_evaluation.py:20; T = x^2 + y: part1.py:187) - for
which P1/P2 interpolation is exact.
"""

from __future__ import annotations

import numpy as np

VALUE, GRAD, MANDEL_STRAIN, DEF_GRAD = 0, 1, 2, 3


def tabulate(kind, u, dofmap, bs, x, x_dofmap, phi, dphi, dpsi, cells=None):
    """out (n_cells, nq, ncomp).  u: flat blocked coefficient (bs * n_dofs); dofmap (n_cells, nb);
    x (n_nodes, 3); x_dofmap (n_cells, gdim+1); phi (nq, nb); dphi (gdim, nq, nb); dpsi (gdim, gdim+1)."""
    gdim = dphi.shape[0]
    if cells is not None:
        dofmap, x_dofmap = dofmap[cells], x_dofmap[cells]
    w = np.asarray(u, dtype=np.float64).reshape(-1, bs)[dofmap]  # (n_cells, nb, bs)   gather through the dofmap
    xv = x[x_dofmap][:, :, :gdim]  # (n_cells, nv, gdim)
    J = np.einsum("cvi,jv->cij", xv, dpsi)  # J_ij = sum_v x_vi dpsi_v/dX_j
    K = np.linalg.inv(J)
    if kind == VALUE:
        return np.einsum("cab,qa->cqb", w, phi)
    G = np.einsum("cab,kqa->cqbk", w, dphi)  # reference gradient
    g = np.einsum("cqbk,ckj->cqbj", G, K)  # physical gradient (identity pull-back, :362)
    nc, nq = g.shape[:2]
    if kind == GRAD:
        return g.reshape(nc, nq, bs * gdim)
    if kind == MANDEL_STRAIN:  # demo_vm:225-227, demo_mc:148-157
        out = np.zeros((nc, nq, 4))
        out[:, :, 0] = g[:, :, 0, 0]
        out[:, :, 1] = g[:, :, 1, 1]
        out[:, :, 3] = np.sqrt(2.0) * 0.5 * (g[:, :, 0, 1] + g[:, :, 1, 0])
        return out
    if kind == DEF_GRAD:  # demo_hyperelasticity.py:479
        return (g + np.eye(gdim)[None, None]).reshape(nc, nq, bs * gdim)
    raise ValueError(kind)


def tabulate_general(kind, u, dofmap, bs, x, x_dofmap, phi, dphi, dgeo, entities=None):
    """General restatement: any element given by table sets, Jacobian per evaluation point (non-affine cells),
    `entities` None / (n,) cells / (n, 2) (cell, local facet) pairs (external_operator.py:365-371,
    test_codim_external_operator.py:75-109).  phi (n_sets, nq, nb); dphi (n_sets, gdim, nq, nb);
    dgeo (n_sets, gdim, nq, ng) derivatives of the geometry element.  out (n_entities, nq, ncomp).
    PARITY UNPINNED like `tabulate` (same third-party arithmetic); checked against analytic fields."""
    gdim = dphi.shape[1]
    if entities is None:
        cells = np.arange(dofmap.shape[0])
        sets = np.zeros(cells.size, dtype=np.int64)
    else:
        ent = np.asarray(entities)
        cells = ent if ent.ndim == 1 else ent[:, 0]
        sets = np.zeros(cells.size, dtype=np.int64) if ent.ndim == 1 else ent[:, 1]
    w = np.asarray(u, dtype=np.float64).reshape(-1, bs)[dofmap[cells]]  # (ne, nb, bs)
    xv = x[x_dofmap[cells]][:, :, :gdim]  # (ne, ng, gdim)
    J = np.einsum("evi,ejqv->eqij", xv, dgeo[sets])  # J_ij(q) = sum_v x_vi dpsi_v/dX_j (q)
    K = np.linalg.inv(J)
    if kind == VALUE:
        return np.einsum("eab,eqa->eqb", w, phi[sets])
    G = np.einsum("eab,ekqa->eqbk", w, dphi[sets])
    g = np.einsum("eqbk,eqkj->eqbj", G, K)
    ne, nq = g.shape[:2]
    if kind == GRAD:
        return g.reshape(ne, nq, bs * gdim)
    if kind == MANDEL_STRAIN:
        out = np.zeros((ne, nq, 4))
        out[:, :, 0] = g[:, :, 0, 0]
        out[:, :, 1] = g[:, :, 1, 1]
        out[:, :, 3] = np.sqrt(2.0) * 0.5 * (g[:, :, 0, 1] + g[:, :, 1, 0])
        return out
    if kind == DEF_GRAD:
        return (g + np.eye(gdim)[None, None]).reshape(ne, nq, bs * gdim)
    raise ValueError(kind)
