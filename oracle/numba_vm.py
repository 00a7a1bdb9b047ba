"""TEST INFRASTRUCTURE - the von Mises radial return as a serial Numba kernel: the execution model of the reference's
own CPU callable (`@numba.njit` loop over cells and points with small-array NumPy algebra per point and three fresh
result arrays per call, doc/demo/demo_plasticity_von_mises.py:298-332, called from `C_tang_impl` :343-352).
bench.py times it next to the GPU figure as the second CPU baseline ("the reference's CPU callable, serial"); the
OpenMP C port in oracle/csrc is ~30x faster per core and stays the reference arm.  Pinned against
tests/golden/vm_seed0_n1026.npz (made by the reference's kernel) in tests/test_oracle_cpu.py.
Only tests/ and bench.py's CPU-baseline legs may import this; the product never does."""

from __future__ import annotations

import numpy as np

from .constitutive import VonMisesParams, deviatoric_projector, elastic_stiffness


def make_return_mapping(prm: VonMisesParams = VonMisesParams()):
    """Compile (first call: a few seconds) and return `f(deps (n,4), sigma_n (n,4), p (n,)) -> (C_tang, sigma, dp)`."""
    import numba

    C_el = elastic_stiffness(prm.lmbda, prm.mu)
    dev = deviatoric_projector()
    mu3, H, sigma_0, mu = 3.0 * prm.mu, prm.H, prm.sigma_0, prm.mu

    @numba.njit
    def point(de, sn, pn):
        trial = sn + C_el @ de  # :308
        s = dev @ trial  # :309
        seq = np.sqrt(3.0 / 2.0 * np.dot(s, s))  # :310
        f = seq - sigma_0 - H * pn  # :312
        fp = (f + np.sqrt(f**2)) / 2.0  # :313
        dp = fp / (mu3 + H)  # :315
        nrm = s / seq * fp / f  # :317
        beta = mu3 * dp / seq  # :318
        Ct = C_el - mu3 * (mu3 / (mu3 + H) - beta) * np.outer(nrm, nrm) - 2 * mu * beta * dev  # :322-323
        return Ct, trial - beta * s, dp  # :320

    @numba.njit
    def run(deps, sigma_n, p):
        n = deps.shape[0]
        Ct_all = np.empty((n, 4, 4), dtype=np.float64)  # :303-305 fresh arrays every call
        sig_all = np.empty_like(sigma_n)
        dp_all = np.empty_like(p)
        for i in range(n):  # :328-330 serial loop
            Ct_all[i], sig_all[i], dp_all[i] = point(deps[i], sigma_n[i], p[i])
        return Ct_all, sig_all, dp_all

    return run
