"""NumPy backend (oracle: tabulation + von Mises + forms restatements) for the thick-walled cylinder driver."""

import numpy as np

from dolfinx_external_operator_b200 import elements as el
from oracle import constitutive as oc
from oracle import forms as of
from oracle import tabulation as ot


class OracleBackend:
    def __init__(self, mesh):
        self.m = mesh
        self.geo = (mesh["x"], mesh["x_dofmap"], mesh["phi"], mesh["dphi"], el.p1_geometry_derivatives(2))
        self.nq = 3 * mesh["dofmap"].shape[0]
        self.sigma_n, self.p = np.zeros((self.nq, 4)), np.zeros(self.nq)
        self.pattern = of.sparsity_pattern(mesh["dofmap"], 2, mesh["n_dofs"])
        self.prm = oc.VonMisesParams()

    def residual(self, Du):
        m = self.m
        eps = ot.tabulate(ot.MANDEL_STRAIN, Du, m["dofmap"], 2, *self.geo).reshape(-1, 4)
        self.Ct, self.sigma, self.dp = (np.asarray(a) for a in oc.vm_return_mapping(eps, self.sigma_n, self.p, self.prm))
        return of.assemble_vector(ot.MANDEL_STRAIN, self.sigma, m["weights"], m["dofmap"], 2, m["n_dofs"], *self.geo)

    def tangent_csr(self):
        m = self.m
        return of.assemble_matrix(ot.MANDEL_STRAIN, ot.MANDEL_STRAIN, self.Ct, m["weights"], m["dofmap"], 2, m["n_dofs"],
                                  *self.geo, *self.pattern)

    def plastic_fraction(self):
        return float((self.dp.reshape(-1) > 0).mean())

    def commit(self):
        self.p = self.p + self.dp.reshape(-1)
        self.sigma_n = self.sigma.reshape(-1, 4).copy()
