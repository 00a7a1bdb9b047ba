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

from .context import DeviceArray, new_evaluation_round

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


def assignment_pairs(external_operator, n_values: int):
    """The reference's non-contiguous assignments (:286-335) as index pairs: the statement sequence
    `x.array[targets[j]] = values.reshape(-1)[sources[j]]` for j = 0, 1, 2, ... is what `_assign_non_mixed`
    (:286-287), `_assign_mixed_2d` (:292-311) and `_assign_mixed_3d` (:313-335) perform."""
    if not getattr(external_operator, "_is_mixed", False):
        targets = np.asarray(external_operator.unrolled_dofmap, dtype=np.int64).reshape(-1)
        if targets.size != n_values:  # what NumPy raises for x.array[dofmap] = values (:287, re-raised at :440-444)
            raise ValueError(f"shape mismatch: value array of shape ({n_values},) could not be broadcast to indexing "
                             f"result of shape ({targets.size},)")
        return targets, np.arange(n_values, dtype=np.int64)
    npt = int(external_operator._n_points_total)
    comp = int(external_operator._comp_size)
    n_cells = n_values // (npt * comp)
    if n_cells * npt * comp != n_values:
        raise ValueError(f"cannot reshape array of size {n_values} into shape ({n_cells},{npt}" + (f",{comp})" if comp > 1 else ")"))
    cells = np.arange(n_cells, dtype=np.int64)[:, None]
    targets, sources = [], []
    for info in external_operator._mixed_subspace_info:
        offset, n_pts = int(info["offset"]), int(info["n_pts"])
        flat_dofs = np.asarray(info["flat_dofs"], dtype=np.int64).reshape(-1)
        pts = np.arange(n_pts, dtype=np.int64)[None, :]
        if comp == 1:  # :305-311  block = values[:, offset:offset+n_pts]
            src = (cells * npt + offset + pts).reshape(-1)
        else:  # :325-335  block = values[:, offset:offset+n_pts, :val_size].reshape(n_cells, dofs_per_cell)
            vs, dpc = int(info["val_size"]), int(info["dofs_per_cell"])
            if n_pts * vs != dpc:
                raise ValueError(f"cannot reshape array of size {n_cells * n_pts * vs} into shape ({n_cells},{dpc})")
            src = ((cells[:, :, None] * npt + offset + pts[:, :, None]) * comp + np.arange(vs, dtype=np.int64)[None, None, :]).reshape(-1)
        if flat_dofs.size != src.size:
            raise ValueError(f"shape mismatch: value array of shape ({src.size},) could not be broadcast to indexing "
                             f"result of shape ({flat_dofs.size},)")
        targets.append(flat_dofs)
        sources.append(src)
    return np.concatenate(targets), np.concatenate(sources)


class AssignPlan:
    """The scatter of :286-335 turned inside out, once per operator: for every degree of freedom the index of the
    value that NumPy's last-one-wins fancy assignment would leave there.  `apply` is then a race-free device
    gather (`eo_assign_gather`) whose output is the compact coefficient array - the only thing that crosses PCIe."""

    def __init__(self, ctx, external_operator, n_values: int):
        x_array = external_operator.ref_coefficient.x.array
        targets, sources = assignment_pairs(external_operator, n_values)
        if targets.size and (targets.min() < 0 or targets.max() >= x_array.size):
            raise IndexError(f"index {int(targets.max())} is out of bounds for axis 0 with size {x_array.size}")
        src_for_dof = np.full(x_array.size, -1, dtype=np.int64)
        src_for_dof[targets] = sources  # NumPy keeps the last assignment for repeated targets - exactly the rule wanted
        self.ctx, self.n_values, self.n_dofs = ctx, int(n_values), int(x_array.size)
        self.touched = None
        if (src_for_dof < 0).any():  # dofs the operator never writes keep their old values
            self.touched = np.nonzero(src_for_dof >= 0)[0]
            src_for_dof = src_for_dof[self.touched]
            self._tmp = ctx.pinned_empty(self.touched.size)
        self.src = ctx.to_device(np.ascontiguousarray(src_for_dof))

    def apply(self, values: DeviceArray, x_array: np.ndarray) -> None:
        c = self.ctx
        out = x_array if self.touched is None else self._tmp
        c.check(c.lib.eo_assign_gather(c.handle, values.ptr, self.n_values, self.src.ptr, self.src.size, out.ctypes.data))
        c.sync()
        if self.touched is not None:
            x_array[self.touched] = self._tmp


def _assign(external_operator, values) -> None:
    """`external_operator._assign_func(values)` (:440-444) with two additions: values that
    already ARE the coefficient array (a model bound with `bind_outputs`) are not copied again,
    `DeviceArray` values are downloaded straight into the coefficient array when the assignment is the
    contiguous one (:289-290), and go through a cached `AssignPlan` (device gather) when it is not (:286-335)."""
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
        # continuous / mixed coefficient spaces (:286-335): gather on the device, download the compact dof array
        plan = getattr(external_operator, "_b200_assign_plan", None)
        if plan is None or plan.n_values != values.size or plan.n_dofs != x_array.size or plan.ctx is not values.ctx:
            plan = AssignPlan(values.ctx, external_operator, values.size)
            external_operator._b200_assign_plan = plan
        plan.apply(values, x_array)
        return
    external_operator._assign_func(values)


def evaluate_external_operators(external_operators, evaluated_operands):
    """Evaluate external operators and update their coefficients (reference :407-448)."""
    return _evaluate_external_operators(external_operators, evaluated_operands, 0)


def _evaluate_external_operators(external_operators, evaluated_operands, depth):
    if depth == 0:
        new_evaluation_round()  # results cached by a callable for the requests of one round end here
    evaluated_operators = []
    for external_operator in external_operators:
        ufl_operands_eval = []
        for operand in external_operator.ufl_operands:
            if _is_external_operator(operand):  # :427-428
                ufl_operands_eval.extend(_evaluate_external_operators([operand], evaluated_operands[operand], depth + 1))
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
