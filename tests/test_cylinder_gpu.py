"""The von Mises demo problem (thick-walled cylinder, demo_plasticity_von_mises.py) end to end on the device-side
consumers - eo_form_vm_step + eo_form_matrix behind `GpuBackend` - against the same driver on the NumPy oracle chain and
against the problem's analytic anchors (Lame, q_lim)."""

import numpy as np
import pytest

from cylinder_util import OracleBackend
import cylinder_driver as twc

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("exact", [True, False])
def test_cylinder_load_stepping_matches_the_oracle_chain(ctx, exact):
    m = twc.quarter_ring_mesh(5, 16)
    ref = twc.solve(m, OracleBackend(m), n_steps=12)
    got = twc.solve(m, twc.GpuBackend(m, ctx=ctx, exact=exact), n_steps=12)
    assert np.array_equal(got["newton_iterations"], ref["newton_iterations"])
    assert np.array_equal(got["plastic_fraction"], ref["plastic_fraction"])  # flags are bit-exact
    np.testing.assert_allclose(got["u_probe"], ref["u_probe"], rtol=1e-9, atol=1e-18)
    np.testing.assert_allclose(got["u"], ref["u"], rtol=0, atol=1e-9 * np.abs(ref["u"]).max())
    # analytic anchors on the GPU path itself
    k = 3
    assert got["plastic_fraction"][k] == 0.0
    assert abs(got["u_probe"][k] - twc.lame_inner_displacement(got["load"][k] * twc.Q_LIM)) < 2e-3 * got["u_probe"][k]
    assert got["plastic_fraction"][got["load"] < 0.9].max() < 1.0 and got["plastic_fraction"][got["load"] > 1.02].min() == 1.0


def test_cylinder_finer_mesh_on_the_gpu(ctx):
    """20 x 64 mesh (2560 cells): Lame to the finer chord error, quadratic convergence in the plastic range."""
    m = twc.quarter_ring_mesh(20, 64)
    got = twc.solve(m, twc.GpuBackend(m, ctx=ctx), n_steps=12)
    k = 3
    assert abs(got["u_probe"][k] - twc.lame_inner_displacement(got["load"][k] * twc.Q_LIM)) < 2e-4 * got["u_probe"][k]
    h = np.array(got["residual_histories"][9])
    assert h[-1] <= 1e-8 * h[0] and len(h) - 1 <= 8
    assert got["plastic_fraction"][got["load"] > 1.02].min() == 1.0
