"""TEST INFRASTRUCTURE - NumPy restatement (the oracle) of the constitutive
callables on the reference's hot path.  NOT part of the product: only `tests/`,
`__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / --impl reference legs
may import this module.

Parity status (see DESIGN.md "Oracle"):
  * von Mises, heat, Isihara: PINNED - `tests/golden/*.npz` were produced by
    executing the reference's own source (`oracle/ref_exec.py`,
    `oracle/gen_golden.py`) and this restatement is checked against them.
  * Mohr-Coulomb: pinned against the reference's own source executed over a
    torch.func shim of the JAX API (JAX itself is absent) - see
    `oracle/jax_on_torch.py`; the restatement lives in `oracle/csrc/mc_oracle.cpp`.
  * operand tabulation: parity UNPINNED (arithmetic lives in un-vendored
    fenics-dolfinx 0.10 / FFCx / basix); checked against analytic fields only.

Every function cites the reference lines it follows (relative to /root/reference).
"""

from __future__ import annotations

import dataclasses

import numpy as np


# ------------------------------------------------------------------ von Mises
@dataclasses.dataclass(frozen=True)
class VonMisesParams:
    """doc/demo/demo_plasticity_von_mises.py:185-192."""

    E: float = 70e3
    nu: float = 0.3
    E_tangent: float = 70e3 / 100.0
    sigma_0: float = 250.0

    @property
    def H(self) -> float:  # :187
        return self.E * self.E_tangent / (self.E - self.E_tangent)

    @property
    def lmbda(self) -> float:  # :190
        return self.E * self.nu / (1.0 + self.nu) / (1.0 - 2.0 * self.nu)

    @property
    def mu(self) -> float:  # :191
        return self.E / 2.0 / (1.0 + self.nu)


def elastic_stiffness(lmbda: float, mu: float) -> np.ndarray:
    """4x4 plane-strain Mandel stiffness, demo_vm:193-201 / demo_mc:407-415."""
    C = np.zeros((4, 4))
    C[:3, :3] = lmbda
    C[0, 0] = C[1, 1] = C[2, 2] = lmbda + 2.0 * mu
    C[3, 3] = 2.0 * mu
    return C


def deviatoric_projector() -> np.ndarray:
    """demo_vm:203-204."""
    D = np.eye(4)
    D[:3, :3] -= 1.0 / 3.0
    return D


def vm_return_mapping(deps, sigma_n, p, prm: VonMisesParams = VonMisesParams()):
    """Radial return with linear isotropic hardening, per quadrature point.

    Follows `_kernel` of demo_plasticity_von_mises.py:307-326 line by line,
    vectorised over the leading axis (the reference loops, :328-330).
    deps, sigma_n: (n, 4); p: (n,).  Returns C_tang (n,4,4), sigma (n,4), dp (n,).
    No branches, like the source: the elastic case falls out of f_plus == 0.
    """
    deps = np.asarray(deps, dtype=np.float64).reshape(-1, 4)
    sigma_n = np.asarray(sigma_n, dtype=np.float64).reshape(-1, 4)
    p = np.asarray(p, dtype=np.float64).reshape(-1)
    C = elastic_stiffness(prm.lmbda, prm.mu)
    D = deviatoric_projector()
    mu, H = prm.mu, prm.H

    sigma_el = sigma_n + deps @ C.T  # :308
    s = sigma_el @ D.T  # :309
    sigma_eq = np.sqrt(1.5 * np.einsum("ni,ni->n", s, s))  # :310
    f_el = sigma_eq - prm.sigma_0 - H * p  # :312
    f_plus = (f_el + np.sqrt(f_el**2)) / 2.0  # :313
    dp = f_plus / (3 * mu + H)  # :315
    with np.errstate(invalid="ignore", divide="ignore"):
        n_el = s / sigma_eq[:, None] * f_plus[:, None] / f_el[:, None]  # :317
    beta = 3 * mu * dp / sigma_eq  # :318
    sigma = sigma_el - beta[:, None] * s  # :320
    nn = n_el[:, :, None] * n_el[:, None, :]  # :322
    C_tang = (
        C[None] - (3 * mu * (3 * mu / (3 * mu + H) - beta))[:, None, None] * nn - (2 * mu * beta)[:, None, None] * D[None]
    )  # :323
    return C_tang, sigma, dp


# ------------------------------------------------------------------ heat
def heat_k(T, A=1.0, B=1.0):
    """demo_nonlinear_heat_equation_part1.py:252-256 ; part2.py:215-216."""
    return 1.0 / (A + B * np.asarray(T, dtype=np.float64))


def heat_dkdT(T, A=1.0, B=1.0):
    """part1.py:271-272."""
    return -B * heat_k(T, A, B) ** 2


def heat_q(T, sigma, A=1.0, B=1.0, gdim=2):
    """part2.py:219-230.  T (n_cells, n_pts); sigma (n_cells, n_pts*gdim)."""
    T = np.asarray(T, dtype=np.float64)
    s = np.asarray(sigma, dtype=np.float64).reshape(T.shape[0], -1, gdim)
    return (-heat_k(T, A, B)[:, :, None] * s).reshape(-1)


def heat_dqdT(T, sigma, A=1.0, B=1.0, gdim=2):
    """part2.py:243-247."""
    T = np.asarray(T, dtype=np.float64)
    s = np.asarray(sigma, dtype=np.float64).reshape(T.shape[0], -1, gdim)
    return (B * (heat_k(T, A, B) ** 2)[:, :, None] * s).reshape(-1)


def heat_dqdsigma(T, sigma=None, A=1.0, B=1.0, gdim=2):
    """part2.py:259-261."""
    T = np.asarray(T, dtype=np.float64)
    return (-heat_k(T, A, B)[:, :, None, None] * np.eye(gdim)[None, None]).reshape(-1)


# ------------------------------------------------------------------ Mohr-Coulomb parameters
@dataclasses.dataclass(frozen=True)
class MohrCoulombParams:
    """doc/demo/demo_plasticity_mohr_coulomb.py:110-116, 469."""

    E: float = 6778.0
    nu: float = 0.25
    c: float = 3.45
    phi: float = 30 * np.pi / 180
    psi: float = 30 * np.pi / 180
    theta_T: float = 26 * np.pi / 180
    a: float = 0.26 * 3.45 / np.tan(30 * np.pi / 180)
    tol: float = 1e-8
    Nitermax: int = 200

    @property
    def lmbda(self) -> float:  # :405
        return self.E * self.nu / ((1.0 + self.nu) * (1.0 - 2.0 * self.nu))

    @property
    def mu(self) -> float:  # :406
        return self.E / (2.0 * (1.0 + self.nu))


# ------------------------------------------------------------------ EXTENSION: full 3-D von Mises (6-component Mandel)
def vm3d_return_mapping(deps, sigma_n, p, prm: VonMisesParams = VonMisesParams()):
    """EXTENSION, not in the reference (its demos are plane strain, 4-component Mandel vectors: demo_vm:193-204):
    the SAME statements as `vm_return_mapping` (demo_vm:307-326) written for 6-component Mandel vectors
    [xx, yy, zz, sqrt2 yz, sqrt2 xz, sqrt2 xy] - BASELINE config 5 asks for a "3D" batch; this restatement is its oracle.
    deps, sigma_n: (n, 6); p: (n,).  Returns C_tang (n,6,6), sigma (n,6), dp (n,)."""
    deps = np.asarray(deps, dtype=np.float64).reshape(-1, 6)
    sigma_n = np.asarray(sigma_n, dtype=np.float64).reshape(-1, 6)
    p = np.asarray(p, dtype=np.float64).reshape(-1)
    l, mu, H = prm.lmbda, prm.mu, prm.H
    tr = np.array([1.0, 1.0, 1.0, 0.0, 0.0, 0.0])
    C = l * np.outer(tr, tr) + 2.0 * mu * np.eye(6)
    D = np.eye(6) - np.outer(tr, tr) / 3.0
    sigma_el = sigma_n + deps @ C.T
    s = sigma_el @ D.T
    sigma_eq = np.sqrt(1.5 * np.einsum("ni,ni->n", s, s))
    f_el = sigma_eq - prm.sigma_0 - H * p
    f_plus = (f_el + np.sqrt(f_el**2)) / 2.0
    dp = f_plus / (3 * mu + H)
    with np.errstate(invalid="ignore", divide="ignore"):
        n_el = s / sigma_eq[:, None] * f_plus[:, None] / f_el[:, None]
    beta = 3 * mu * dp / sigma_eq
    sigma = sigma_el - beta[:, None] * s
    nn = n_el[:, :, None] * n_el[:, None, :]
    C_tang = C[None] - (3 * mu * (3 * mu / (3 * mu + H) - beta))[:, None, None] * nn - (2 * mu * beta)[:, None, None] * D[None]
    return C_tang, sigma, dp
