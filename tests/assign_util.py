"""Duck-typed stand-ins for FEMExternalOperator carrying exactly the attributes the assignment path reads
(external_operator.py:137-205): continuous (non-mixed, unrolled dofmap) and mixed coefficient spaces."""

from types import SimpleNamespace

import numpy as np

from oracle import assign as oa


class _X:
    def __init__(self, n):
        self.array = np.full(n, -7.25)
        self.scattered = 0

    def scatter_forward(self):
        self.scattered += 1


def continuous_operator(n_cells=500, dofs_per_cell=3, bs=2, seed=0, untouched=0):
    """A P1-like continuous space: cells share dofs (many duplicates), block size bs, optionally some dofs that no
    cell touches."""
    rng = np.random.default_rng(seed)
    n_nodes = n_cells // 2 + 5
    cell_nodes = rng.integers(0, n_nodes, (n_cells, dofs_per_cell))
    unrolled = (cell_nodes[:, :, None] * bs + np.arange(bs)[None, None, :]).reshape(-1).astype(np.int32)  # get_unrolled_dofmap
    op = SimpleNamespace(_is_mixed=False, unrolled_dofmap=unrolled, ref_coefficient=SimpleNamespace(x=_X(n_nodes * bs + untouched)))
    op._assign_func = lambda values: oa.assign_non_mixed(op.ref_coefficient.x.array, op.unrolled_dofmap, values)
    op.n_values = unrolled.size
    return op


def mixed_operator(n_cells=300, sub=((3, 1), (6, 1)), seed=1):
    """Mixed space; sub = ((n_pts, val_size), ...).  comp_size = max val_size (2-D values if 1, else 3-D)."""
    rng = np.random.default_rng(seed)
    comp = max(v for _, v in sub)
    infos, offset, n_dofs = [], 0, 0
    for n_pts, vs in sub:
        dpc = n_pts * vs
        n_sub = n_cells * dpc // 3 + 7  # shared dofs -> duplicates
        flat = (n_dofs + rng.integers(0, n_sub, n_cells * dpc)).astype(np.int32)
        infos.append({"n_pts": n_pts, "val_size": vs, "dofs_per_cell": dpc, "flat_dofs": flat, "offset": offset})
        offset += n_pts
        n_dofs += n_sub
    op = SimpleNamespace(_is_mixed=True, _comp_size=comp, _n_points_total=offset, _mixed_subspace_info=infos,
                         ref_coefficient=SimpleNamespace(x=_X(n_dofs)))
    if comp == 1:
        op._assign_func = lambda values: oa.assign_mixed_2d(op.ref_coefficient.x.array, infos, offset, values)
    else:
        op._assign_func = lambda values: oa.assign_mixed_3d(op.ref_coefficient.x.array, infos, offset, comp, values)
    op.n_values = n_cells * offset * comp
    return op


CASES = {
    "continuous_bs1": lambda: continuous_operator(bs=1),
    "continuous_bs2": lambda: continuous_operator(bs=2, seed=3),
    "continuous_untouched": lambda: continuous_operator(bs=1, seed=4, untouched=11),
    "mixed_scalar_scalar": lambda: mixed_operator(sub=((3, 1), (6, 1))),
    "mixed_vector_scalar": lambda: mixed_operator(sub=((3, 2), (3, 1)), seed=5),
    "mixed_tensor_vector": lambda: mixed_operator(sub=((4, 9), (4, 3)), seed=6),
}
