"""Shared checker for the Mohr-Coulomb parity tests (CPU host harness and GPU)."""

import numpy as np

from oracle import constitutive as oc

RTOL = 1e-10  # north_star tolerance for local-Newton models


def lode_arg(sig):
    """sin(3 theta) of demo_plasticity_mohr_coulomb.py:290-292 for an array of Mandel stresses."""
    sd = np.asarray(sig).reshape(-1, 4) @ oc.deviatoric_projector().T
    J2 = 0.5 * (sd**2).sum(1)
    J3 = sd[:, 2] * (sd[:, 0] * sd[:, 1] - sd[:, 3] ** 2 / 2)
    with np.errstate(invalid="ignore", divide="ignore"):
        return -(3 * np.sqrt(3) * J3) / (2 * np.sqrt(J2**3))


def corner_points(ref, deps, sigma_n, prm, thresh=1e-6):
    """Plastic points whose Lode angle is within ~0.01 degree of a corner of the Mohr-Coulomb hexagon
    (1 - sin^2(3 theta) < 1e-6 at sigma_n, at the trial stress or at the returned stress).

    There the REFERENCE's own AD evaluates d sin(3 theta)/d arg as cos(3 theta)/sqrt(1 - arg^2) (asin followed
    by sin, demo_mc:292-294, :340-342): a 0/0-type quotient whose relative rounding noise is eps/(1 - arg^2),
    i.e. >= 1e-10 for these points.  Two correct IEEE evaluations of the reference's program (with and without
    FMA contraction, different libm) already differ by that much, so such points are compared with the
    tolerance scaled by that condition number (capped at 1e-6 relative); all other points at 1e-10.
    Returned stresses cluster at the hexagon corners, so ~0.1 % of a batch is affected."""
    w2 = 1.0 - lode_arg(ref["sigma"]) ** 2
    C = oc.elastic_stiffness(prm.lmbda, prm.mu)
    sigma_n = np.asarray(sigma_n).reshape(-1, 4)
    deps = np.asarray(deps).reshape(-1, 4)
    w2 = np.minimum(w2, 1.0 - lode_arg(sigma_n) ** 2)
    w2 = np.minimum(w2, 1.0 - lode_arg(sigma_n + deps @ C.T) ** 2)
    w2 = np.where(np.isfinite(w2), np.maximum(w2, 1e-10), 1.0)
    corner = (w2 < thresh) & (np.asarray(ref["yielding"]) > 0)
    return corner, np.where(corner, thresh / w2, 1.0)


def check_mc(out, ref, deps, sigma_n, prm, rtol=RTOL):
    """Flags and iteration counts bit-exact; tangent / stress / dlambda / yielding within rtol (relative to the
    field's scale, corner points scaled as explained above); ||res|| (a converged, rounding-level quantity)
    within an absolute 1e-10 of the stress scale."""
    assert np.array_equal(out["niter"], ref["niter"])
    assert np.array_equal(np.asarray(out["yielding"]) > 0, np.asarray(ref["yielding"]) > 0)
    corner, scale = corner_points(ref, deps, sigma_n, prm)
    assert corner.mean() < 2e-2 or corner.size < 200, corner.mean()
    for k, w in (("C_tang", 16), ("sigma", 4), ("dlambda", 1), ("yielding", 1)):
        a, b = np.asarray(out[k]).reshape(-1, w), np.asarray(ref[k]).reshape(-1, w)
        tol = rtol * scale[:, None]
        bad = ~(np.abs(a - b) <= tol * (np.abs(b) + np.abs(b[np.isfinite(b)]).max())) & ~(np.isnan(a) & np.isnan(b))
        assert not bad.any(), (k, np.argwhere(bad)[:5], np.nanmax(np.abs(a - b)))
    smax = np.abs(np.asarray(ref["sigma"])[np.isfinite(ref["sigma"])]).max()
    np.testing.assert_allclose(out["norm_res"], ref["norm_res"], rtol=0, atol=1e-10 * smax)
    return int(corner.sum())
