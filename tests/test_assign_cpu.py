"""Host logic of the non-contiguous coefficient assignment (external_operator.py:286-335): the index pairs and
the inside-out plan reproduce the reference's NumPy statements (oracle/assign.py) bit for bit, including the
last-write-wins rule for dofs shared between cells."""

import numpy as np
import pytest

from assign_util import CASES
from dolfinx_external_operator_b200.external_operator import assignment_pairs


@pytest.mark.parametrize("case", sorted(CASES))
def test_pairs_and_plan_reproduce_reference_assignment(case):
    op = CASES[case]()
    values = np.random.default_rng(7).standard_normal(op.n_values)
    ref = op.ref_coefficient.x.array.copy()
    op._assign_func(values)  # the reference's statements
    want, op.ref_coefficient.x.array = op.ref_coefficient.x.array, ref
    targets, sources = assignment_pairs(op, values.size)
    seq = ref.copy()
    for t, s in zip(targets, sources):  # the pairs applied one by one, in order
        seq[t] = values[s]
    assert np.array_equal(seq, want)
    src_for_dof = np.full(ref.size, -1, dtype=np.int64)  # what AssignPlan builds
    src_for_dof[targets] = sources
    got = ref.copy()
    touched = src_for_dof >= 0
    got[touched] = values[src_for_dof[touched]]
    assert np.array_equal(got, want)
    if case == "continuous_untouched":
        assert (~touched).sum() >= 11 and np.all(got[~touched] == -7.25)


def test_size_mismatch_raises_value_error():  # :440-444
    op = CASES["continuous_bs1"]()
    with pytest.raises(ValueError):
        assignment_pairs(op, op.n_values + 3)
    op = CASES["mixed_vector_scalar"]()
    with pytest.raises(ValueError):
        assignment_pairs(op, op.n_values + 1)
