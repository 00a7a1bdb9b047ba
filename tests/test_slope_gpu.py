"""The Mohr-Coulomb demo problem (slope stability) end to end on the device-side consumers - eo_mc_eval + eo_form_vector
+ eo_form_matrix behind `GpuBackend` - against the same driver on the oracle chain and against limit analysis."""

import numpy as np
import pytest

import slope_driver as ss
from slope_util import OracleBackend
from test_slope_cpu import STEPS, check_collapse

pytestmark = pytest.mark.gpu


def test_slope_matches_the_oracle_chain(ctx):
    m = ss.slope_mesh(12, 10)
    ref = ss.solve(m, OracleBackend(m), load_steps=STEPS)
    got = ss.solve(m, ss.GpuBackend(m, ctx=ctx), load_steps=STEPS)
    assert abs(got["n_converged"] - ref["n_converged"]) <= 1  # the last converged step sits on the plateau
    k = int(np.searchsorted(STEPS, 6.3 * ss.C_COHESION / ss.H))  # well before the plateau: a well-conditioned path
    assert np.array_equal(got["newton_iterations"][:k], ref["newton_iterations"][:k])
    np.testing.assert_allclose(got["u_probe"][:k], ref["u_probe"][:k], rtol=1e-7)
    np.testing.assert_allclose(got["plastic_fraction"][:k], ref["plastic_fraction"][:k], rtol=0, atol=2.0 / (3 * m["dofmap"].shape[0]))
    check_collapse(got)


def test_slope_demo_mesh_on_the_gpu(ctx):
    """The demo's own 25 x 25 mesh and load steps (demo_mc:120, 708-710): every step up to gamma = 22.99 converges, as
    in the demo, and the plateau is reached within 4 % of Chen's factor."""
    m = ss.slope_mesh(25, 25)
    got = ss.solve(m, ss.GpuBackend(m, ctx=ctx), load_steps=STEPS)
    assert got["n_converged"] >= 52
    check_collapse(got, lo=0.99, hi=1.04)
