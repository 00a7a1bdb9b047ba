"""GPU parity of the non-contiguous coefficient assignment: `evaluate_external_operators` with device-resident
operator values (AssignPlan -> eo_assign_gather) against the reference's NumPy statements (oracle/assign.py)."""

import numpy as np
import pytest

import dolfinx_external_operator_b200 as eo
from assign_util import CASES

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("case", sorted(CASES))
def test_device_assignment_matches_reference(ctx, case):
    op = CASES[case]()
    values = np.random.default_rng(9).standard_normal(op.n_values)
    ref = op.ref_coefficient.x.array.copy()
    op._assign_func(values)
    want = op.ref_coefficient.x.array.copy()
    op.ref_coefficient.x.array[:] = ref
    d_values = ctx.to_device(values)
    op.ufl_operands = ["u"]
    op.derivatives = (0,)
    op.external_function = lambda derivatives: (lambda u: d_values)
    l0 = ctx.launch_count
    (out,) = eo.evaluate_external_operators([op], {"u": None})
    assert out is d_values and ctx.launch_count > l0
    assert np.array_equal(op.ref_coefficient.x.array, want)  # bit-identical, duplicates resolved like NumPy
    assert op.ref_coefficient.x.scattered == 1
    # second call reuses the cached plan
    plan = op._b200_assign_plan
    op.ref_coefficient.x.array[:] = ref
    eo.evaluate_external_operators([op], {"u": None})
    assert op._b200_assign_plan is plan and np.array_equal(op.ref_coefficient.x.array, want)


def test_jit_model_on_a_continuous_space(ctx):
    """A JitModel with output='device' feeding a continuous (shared-dof) coefficient: k(T) of part1.py:252-262."""
    from dolfinx_external_operator_b200 import jit_models as jm

    op = CASES["continuous_bs1"]()
    T = np.random.default_rng(2).uniform(0.0, 2.0, op.n_values)
    op.ufl_operands = ["T"]
    op.derivatives = (1,)
    op.external_function = jm.heat_conductivity(ctx=ctx, output="device")
    eo.evaluate_external_operators([op], {"T": T.reshape(-1, 3)})
    want = np.full(op.ref_coefficient.x.array.size, -7.25)
    want[op.unrolled_dofmap] = -1.0 / (1.0 + T) ** 2
    np.testing.assert_allclose(op.ref_coefficient.x.array, want, rtol=1e-14)


def test_device_assignment_size_mismatch(ctx):
    op = CASES["continuous_bs1"]()
    op.ufl_operands, op.derivatives = ["u"], (0,)
    d = ctx.to_device(np.zeros(op.n_values + 2))
    op.external_function = lambda derivatives: (lambda u: d)
    with pytest.raises(ValueError):
        eo.evaluate_external_operators([op], {"u": None})
