"""TEST INFRASTRUCTURE - NumPy restatement (the oracle) of the device-side consumers (csrc/form.cu): the two
integrals the reference hands to DOLFINx right after `evaluate_external_operators`,

    assemble_vector(b, F),  F = inner(N, OP(v)) dx           petsc/petsc.py:64; demo_plasticity_von_mises.py:253
    assemble_matrix(A, J),  J = inner(dN OP(u_hat), OP(v)) dx  petsc/petsc.py:88; demo_vm:390-398

with N, dN the quadrature-space coefficients written by the external operators.

PARITY UNPINNED: like the tabulation (oracle/tabulation.py) the arithmetic lives in un-vendored third-party code
(DOLFINx 0.10 assemblers + FFCx kernels) that cannot run here.  This is the textbook element loop
    b_e[a, c] = sum_q w_q |det J_cell| sum_k N[cell, q, k] OP(phi_a e_c)[k](x_q),      A_e = B^T W D B
written through the SAME operand matrix B as `oracle.tabulation.tabulate` (so b = B^T W N exactly), and it is
anchored on known answers instead: adjointness <B u, W s> = <u, b(s)> against the analytically checked tabulation,
integrals of polynomials the quadrature rule integrates exactly (sum_i b_i = |domain|, the reference's own
comparison of an assembled operator form with its pure-UFL value, test_external_operators_evaluation.py:40-45),
rigid-body modes in the kernel of the elastic stiffness, symmetry, and the reference's Taylor test (residual vs
tangent action, demo_plasticity_mohr_coulomb.py:1203-1235, remainder slope 2)."""

from __future__ import annotations

import numpy as np

from .tabulation import DEF_GRAD, GRAD, MANDEL_STRAIN, VALUE


def ncomp(kind, bs, gdim):
    return bs if kind == VALUE else (4 if kind == MANDEL_STRAIN else bs * gdim)


def operand_matrix(kind, dofmap, bs, x, x_dofmap, phi, dphi, dpsi):
    """B[cell, q, k, a, c] = component k of OP(phi_a e_c) at point q (the linear part of the operand: DEF_GRAD = GRAD),
    and |det J| per cell."""
    gdim = dphi.shape[0]
    nq, nb = phi.shape
    xv = x[x_dofmap][:, :, :gdim]
    J = np.einsum("cvi,jv->cij", xv, dpsi)
    K = np.linalg.inv(J)
    adet = np.abs(np.linalg.det(J))
    nc = dofmap.shape[0]
    B = np.zeros((nc, nq, ncomp(kind, bs, gdim), nb, bs))
    if kind == VALUE:
        for c in range(bs):
            B[:, :, c, :, c] = phi[None]
        return B, adet
    g = np.einsum("kqa,ckj->cqaj", dphi, K)  # physical gradient of phi_a
    if kind in (GRAD, DEF_GRAD):
        for c in range(bs):
            for j in range(gdim):
                B[:, :, c * gdim + j, :, c] = g[:, :, :, j]
        return B, adet
    if kind == MANDEL_STRAIN:  # demo_vm:225-227
        r = np.sqrt(2.0) * 0.5
        B[:, :, 0, :, 0] = g[:, :, :, 0]
        B[:, :, 1, :, 1] = g[:, :, :, 1]
        B[:, :, 3, :, 0] = r * g[:, :, :, 1]
        B[:, :, 3, :, 1] = r * g[:, :, :, 0]
        return B, adet
    raise ValueError(kind)


def _scatter(be, dofmap, bs, n_dofs):
    b = np.zeros((n_dofs, bs))
    np.add.at(b, dofmap, be)  # be (nc, nb, bs)
    return b.reshape(-1)


def assemble_vector(kind, coef, weights, dofmap, bs, n_dofs, x, x_dofmap, phi, dphi, dpsi, n_cells=None):
    """b (bs * n_dofs).  coef (n_cells, nq, ncomp)."""
    nc = dofmap.shape[0] if n_cells is None else n_cells
    B, adet = operand_matrix(kind, dofmap[:nc], bs, x, x_dofmap[:nc], phi, dphi, dpsi)
    s = np.asarray(coef, dtype=np.float64).reshape(-1, phi.shape[0], B.shape[2])[:nc]
    be = np.einsum("q,c,cqk,cqkab->cab", weights, adet, s, B)
    return _scatter(be, dofmap[:nc], bs, n_dofs)


def element_matrices(kind_test, kind_trial, D, weights, dofmap, bs, x, x_dofmap, phi, dphi, dpsi):
    Bt, adet = operand_matrix(kind_test, dofmap, bs, x, x_dofmap, phi, dphi, dpsi)
    Bi, _ = operand_matrix(kind_trial, dofmap, bs, x, x_dofmap, phi, dphi, dpsi)
    Dm = np.asarray(D, dtype=np.float64).reshape(dofmap.shape[0], phi.shape[0], Bt.shape[2], Bi.shape[2])
    return np.einsum("q,c,cqkab,cqkl,cqlde->cabde", weights, adet, Bt, Dm, Bi)  # (nc, nb, bs, nb, bs)


def apply_action(kind_test, kind_trial, D, xvec, weights, dofmap, bs, n_dofs, x, x_dofmap, phi, dphi, dpsi, n_cells=None):
    """y = A x without forming A."""
    nc = dofmap.shape[0] if n_cells is None else n_cells
    Ae = element_matrices(kind_test, kind_trial, np.asarray(D).reshape(dofmap.shape[0], -1)[:nc], weights, dofmap[:nc], bs,
                          x, x_dofmap[:nc], phi, dphi, dpsi)
    w = np.asarray(xvec, dtype=np.float64).reshape(-1, bs)[dofmap[:nc]]
    ye = np.einsum("cabde,cde->cab", Ae, w)
    return _scatter(ye, dofmap[:nc], bs, n_dofs)


def sparsity_pattern(dofmap, bs, n_dofs):
    """(row_ptr int32, col int32) of the scalar-dof CSR pattern: every pair of dofs sharing a cell."""
    nb = dofmap.shape[1]
    sd = (bs * dofmap[:, :, None] + np.arange(bs)[None, None, :]).reshape(dofmap.shape[0], nb * bs).astype(np.int64)
    rows = np.repeat(sd, nb * bs, axis=1).reshape(-1)
    cols = np.tile(sd, (1, nb * bs)).reshape(-1)
    n = bs * n_dofs
    key = np.unique(rows * n + cols)
    r, c = key // n, key % n
    row_ptr = np.zeros(n + 1, dtype=np.int64)
    np.add.at(row_ptr, r + 1, 1)
    return np.cumsum(row_ptr).astype(np.int32), c.astype(np.int32)


def assemble_matrix(kind_test, kind_trial, D, weights, dofmap, bs, n_dofs, x, x_dofmap, phi, dphi, dpsi, row_ptr, col):
    """CSR values (nnz,) on the given pattern."""
    Ae = element_matrices(kind_test, kind_trial, D, weights, dofmap, bs, x, x_dofmap, phi, dphi, dpsi)
    nc, nb = dofmap.shape
    nd = nb * bs
    sd = (bs * dofmap[:, :, None] + np.arange(bs)[None, None, :]).reshape(nc, nd).astype(np.int64)
    Ae = Ae.reshape(nc, nd, nd)
    vals = np.zeros(col.size)
    n = bs * n_dofs
    rowkey = np.repeat(np.arange(n, dtype=np.int64), np.diff(row_ptr)) * n + col
    pos = np.searchsorted(rowkey, (sd[:, :, None] * n + sd[:, None, :]).reshape(-1))
    np.add.at(vals, pos, Ae.reshape(-1))
    return vals
