"""The reference's numeric-layer API - `evaluate_operands`,
`evaluate_external_operators` - re-implemented so that operands and operator
values may live on the GPU, plus the symbolic layer re-exported UNCHANGED.

Reference: src/dolfinx_external_operator/external_operator.py
  :338-404  evaluate_operands            :407-448  evaluate_external_operators
  :49-335   FEMExternalOperator          :670-684  replace_external_operators

The symbolic layer (`FEMExternalOperator`, `replace_external_operators`) is
set-up-time UFL rewriting; it is not on the hot path and is imported from the
reference package when that is installed (it needs ufl/basix/dolfinx, which are
absent from the build container - the names then raise ImportError on use).

Objects are used through the attributes the reference itself uses
(`ufl_operands`, `derivatives`, `external_function`, `ref_coefficient.x.array`,
`ref_coefficient.x.scatter_forward()`, `_assign_func`, `eval_points`,
`ref_function_space`), so the functions work on real `FEMExternalOperator`s and on
light duck-typed stand-ins alike.
"""

from __future__ import annotations

import numpy as np

from .context import DeviceArray

try:  # pragma: no cover - needs the FEniCS stack
    from dolfinx_external_operator import FEMExternalOperator, replace_external_operators  # noqa: F401

    HAVE_REFERENCE_SYMBOLIC_LAYER = True
except Exception as _exc:  # ImportError or a missing dependency inside it
    HAVE_REFERENCE_SYMBOLIC_LAYER = False
    _why = repr(_exc)

    def _missing(name):
        def fn(*a, **k):
            raise ImportError(
                f"{name} is the reference's symbolic (UFL) layer and is re-exported unchanged from "
                f"`dolfinx_external_operator`, which could not be imported here: {_why}"
            )

        fn.__name__ = name
        return fn

    FEMExternalOperator = _missing("FEMExternalOperator")
    replace_external_operators = _missing("replace_external_operators")


def _is_external_operator(operand) -> bool:
    """`isinstance(operand, ufl.ExternalOperator)` (:383, :427) without importing ufl."""
    return hasattr(operand, "external_function") and hasattr(operand, "ufl_operands")


def evaluate_operands(external_operators, entities=None, *, tabulator=None):
    """Evaluate the operands of external operators (reference :338-404).

    Same signature and return value as the reference (`{operand: array}`, nested
    operators give a sub-dict, empty list gives `{}`).  For every operand a GPU
    tabulation plan is looked up - `tabulator.plan_for(external_operator, operand)`
    with `tabulator` defaulting to `external_operator.b200_tabulator` - and, when
    one exists, the operand is tabulated by the sm_100a kernel (`tabulation.py`)
    and the result is a `DeviceArray` of shape `(n_entities, n_points, *shape)`
    that the GPU callables consume in place.  Operands without a plan are
    evaluated exactly as the reference does, through the operator's cached
    `fem.Expression` (:386-402) - that is the reference's own code path for
    arbitrary UFL, not a re-implementation.
    """
    if len(external_operators) == 0:  # :356-357
        return {}
    evaluated_operands = {}
    for external_operator in external_operators:
        tab = tabulator if tabulator is not None else getattr(external_operator, "b200_tabulator", None)
        for operand in external_operator.ufl_operands:
            if operand in evaluated_operands:  # :380-381 evaluate each unique operand once
                continue
            if _is_external_operator(operand):  # :383-384
                evaluated_operands[operand] = evaluate_operands([operand], entities, tabulator=tabulator)
                continue
            plan = tab.plan_for(external_operator, operand) if tab is not None else None
            if plan is not None:
                evaluated_operands[operand] = plan.evaluate(entities)
            else:
                evaluated_operands[operand] = _reference_expression_eval(external_operator, operand, entities)
    return evaluated_operands


def _reference_expression_eval(external_operator, operand, entities):
    """Reference path :365-402 for operands that have no GPU plan (needs dolfinx)."""
    try:
        import ufl
        from dolfinx import fem
        from dolfinx import mesh as _mesh
    except ImportError as exc:
        raise ImportError(
            "operand has no GPU tabulation plan and dolfinx is not importable, so the reference's "
            "fem.Expression path cannot be used either"
        ) from exc
    ref_function_space = external_operator.ref_function_space
    mesh = ref_function_space.mesh
    assert isinstance(ref_function_space.ufl_element().pullback, ufl.pullback.IdentityPullback)  # :362
    if entities is None:  # :365-371
        entities = getattr(mesh, "_full_cells", None)
        if entities is None:
            map_c = mesh.topology.index_map(mesh.topology.dim)
            entities = np.arange(0, map_c.size_local + map_c.num_ghosts, dtype=np.int32)
            mesh._full_cells = entities
    if not hasattr(external_operator, "_compiled_operands"):
        external_operator._compiled_operands = {}
    cached = external_operator._compiled_operands.get(operand)
    if cached is None:  # :387-399
        operand_domain = ufl.domain.extract_unique_domain(operand)
        if operand_domain == ref_function_space.ufl_domain():
            operand_mesh = mesh
        else:
            operand_mesh = _mesh.Mesh(operand_domain.ufl_cargo(), operand_domain)
        expr = fem.Expression(operand, external_operator.eval_points, dtype=external_operator.ref_coefficient.dtype)
        cached = (expr, operand_mesh)
        external_operator._compiled_operands[operand] = cached
    expr, operand_mesh = cached
    return expr.eval(operand_mesh, entities)


def _assign(external_operator, values) -> None:
    """`external_operator._assign_func(values)` (:440-444) with two additions: values that
    already ARE the coefficient array (a model bound with `bind_outputs`) are not copied again,
    and `DeviceArray` values are downloaded straight into the coefficient array when the
    assignment is the contiguous one (:289-290)."""
    x_array = external_operator.ref_coefficient.x.array
    if isinstance(values, np.ndarray) and values.ctypes.data == x_array.ctypes.data and values.size == x_array.size:
        return
    if isinstance(values, DeviceArray):
        if getattr(external_operator, "unrolled_dofmap", None) is None and not getattr(
            external_operator, "_is_mixed", False
        ):
            if values.size != x_array.size:
                raise ValueError(
                    f"could not broadcast input array from shape ({values.size},) into shape ({x_array.size},)"
                )
            values.to_host(x_array.reshape(values.shape))
            return
        values = values.to_host().reshape(-1)
    external_operator._assign_func(values)


def evaluate_external_operators(external_operators, evaluated_operands):
    """Evaluate external operators and update their coefficients (reference :407-448)."""
    evaluated_operators = []
    for external_operator in external_operators:
        ufl_operands_eval = []
        for operand in external_operator.ufl_operands:
            if _is_external_operator(operand):  # :427-428
                ufl_operands_eval.extend(evaluate_external_operators([operand], evaluated_operands[operand]))
            else:
                ufl_operands_eval.append(evaluated_operands[operand])

        external_operator_eval = external_operator.external_function(external_operator.derivatives)(*ufl_operands_eval)

        if type(external_operator_eval) is tuple:  # :435-438
            values = external_operator_eval[0]
        else:
            values = external_operator_eval

        _assign(external_operator, values)
        external_operator.ref_coefficient.x.scatter_forward()  # :445
        evaluated_operators.append(external_operator_eval)
    return evaluated_operators
