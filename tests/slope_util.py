"""NumPy / C++ oracle backend (tabulation + Mohr-Coulomb + forms restatements) for the slope-stability driver."""

import numpy as np

from dolfinx_external_operator_b200 import elements as el
from oracle import constitutive as oc
from oracle import forms as of
from oracle import native
from oracle import tabulation as ot


class OracleBackend:
    def __init__(self, mesh):
        self.m = mesh
        self.geo = (mesh["x"], mesh["x_dofmap"], mesh["phi"], mesh["dphi"], el.p1_geometry_derivatives(2))
        self.nq = 3 * mesh["dofmap"].shape[0]
        self.sigma_n = np.zeros((self.nq, 4))
        self.pattern = of.sparsity_pattern(mesh["dofmap"], 2, mesh["n_dofs"])
        self.prm = oc.MohrCoulombParams()
        self.args = (mesh["weights"], mesh["dofmap"], 2, mesh["n_dofs"], *self.geo)

    def body_force(self):
        g = np.zeros((self.nq, 2))
        g[:, 1] = -1.0
        return of.assemble_vector(ot.VALUE, g, *self.args)

    def residual(self, Du):
        eps = ot.tabulate(ot.MANDEL_STRAIN, Du, self.m["dofmap"], 2, *self.geo).reshape(-1, 4)
        self.out = native.mc_return_mapping(eps, self.sigma_n, self.prm, parallel=True)
        return of.assemble_vector(ot.MANDEL_STRAIN, self.out["sigma"], *self.args)

    def tangent_csr(self):
        return of.assemble_matrix(ot.MANDEL_STRAIN, ot.MANDEL_STRAIN, self.out["C_tang"], *self.args, *self.pattern)

    def plastic_fraction(self):
        return float((np.asarray(self.out["yielding"]) > 0).mean())

    def commit(self):
        self.sigma_n = np.asarray(self.out["sigma"]).reshape(-1, 4).copy()
