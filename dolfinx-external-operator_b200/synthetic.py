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
