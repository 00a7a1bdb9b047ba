"""CPU: host-side logic of the drop-in API (reference semantics of evaluate_operands /
evaluate_external_operators, external_operator.py:338-448) with duck-typed operators, plus the
partition / statistics reduction helpers (world_size-2 gloo)."""

import os
import subprocess
import sys

import numpy as np
import pytest

from dolfinx_external_operator_b200 import evaluate_external_operators, evaluate_operands
from dolfinx_external_operator_b200 import parallel as par

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class _X:
    def __init__(self, n):
        self.array = np.zeros(n)
        self.scatters = 0

    def scatter_forward(self):
        self.scatters += 1


class _Coeff:
    def __init__(self, n):
        self.x = _X(n)
        self.dtype = np.float64


class FakeOperator:
    """The attributes the reference's numeric layer touches (external_operator.py:375-445)."""

    def __init__(self, operands, n_out, external_function, derivatives):
        self.ufl_operands = tuple(operands)
        self.ref_coefficient = _Coeff(n_out)
        self.external_function = external_function
        self.derivatives = derivatives
        self.unrolled_dofmap = None
        self._is_mixed = False

    def _assign_func(self, values):  # :289-290
        self.ref_coefficient.x.array[:] = values


class FakeTabulator:
    def __init__(self, table):
        self.table = table
        self.calls = 0

    def plan_for(self, op, operand):
        tab = self

        class Plan:
            def evaluate(self, entities):
                tab.calls += 1
                return tab.table[operand]

        return Plan() if operand in self.table else None


def test_empty_lists():  # test_external_operators_construction.py:202-212
    assert evaluate_operands([]) == {}
    assert evaluate_external_operators([], {}) == []


def test_tuple_rule_and_assignment_and_scatter():
    n = 12
    vals = np.arange(n, dtype=float)

    def ext(derivatives):
        assert derivatives == (1,)
        return lambda a: (a.reshape(-1) * 2.0, "aux")

    op = FakeOperator(["u"], n, ext, (1,))
    tab = FakeTabulator({"u": vals.reshape(4, 3)})
    ops = evaluate_operands([op, op], tabulator=tab)
    assert tab.calls == 1  # unique operands evaluated once (:380-381)
    out = evaluate_external_operators([op], ops)
    assert out[0][1] == "aux" and np.array_equal(out[0][0], vals * 2)
    assert np.array_equal(op.ref_coefficient.x.array, vals * 2)
    assert op.ref_coefficient.x.scatters == 1


def test_shape_mismatch_raises_value_error():  # :440-444
    op = FakeOperator(["u"], 5, lambda d: (lambda a: np.zeros(7)), (0,))
    with pytest.raises(ValueError):
        evaluate_external_operators([op], {"u": np.zeros(3)})


def test_not_implemented_derivative_propagates():  # demo_vm:364-368
    def ext(derivatives):
        raise NotImplementedError(f"No external function is defined for the requested derivative {derivatives}.")

    op = FakeOperator(["u"], 3, ext, (2,))
    with pytest.raises(NotImplementedError):
        evaluate_external_operators([op], {"u": np.zeros(3)})


def test_bound_output_is_not_copied_again():
    op = FakeOperator(["u"], 6, None, (0,))
    arr = op.ref_coefficient.x.array

    def ext(derivatives):
        def f(a):
            arr[:] = 7.0
            return arr

        return f

    op.external_function = ext
    op._assign_func = lambda v: (_ for _ in ()).throw(AssertionError("copy must be skipped"))
    evaluate_external_operators([op], {"u": np.zeros(6)})
    assert np.all(arr == 7.0)


def test_nested_operator_recursion():  # :383-384, :427-428
    inner = FakeOperator(["u"], 4, lambda d: (lambda a: a.reshape(-1) + 1.0), (0,))
    outer = FakeOperator([inner], 4, lambda d: (lambda a: a * 10.0), (0,))
    tab = FakeTabulator({"u": np.arange(4.0)})
    ops = evaluate_operands([outer], tabulator=tab)
    assert isinstance(ops[inner], dict) and np.array_equal(ops[inner]["u"], np.arange(4.0))
    out = evaluate_external_operators([outer], ops)
    assert np.array_equal(out[0], (np.arange(4.0) + 1) * 10)
    assert np.array_equal(inner.ref_coefficient.x.array, np.arange(4.0) + 1)


def test_partition_covers_range():
    for n in (0, 1, 7, 100, 10**9 + 3):
        for world in (1, 2, 3, 8):
            spans = [par.partition(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            for a, b in zip(spans, spans[1:]):
                assert a[1] == b[0]
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        par.partition(10, 2, 2)


_GLOO_WORKER = r"""
import os, sys
sys.path.insert(0, {root!r})
import numpy as np, torch.distributed as dist
from dolfinx_external_operator_b200 import parallel as par
dist.init_process_group("gloo")
r, w = dist.get_rank(), dist.get_world_size()
a, b = par.partition(1001, r, w)
hist = np.zeros(par.N_SUM - 4, dtype=np.int64); hist[1 + r] = 10 * (r + 1)
st = dict(n_points=b - a, n_plastic=r + 1, n_nonconverged=r, n_nonfinite=0, niter_hist=hist,
          niter_max=float(3 + r), f_max=0.5 * (r + 1), res_max=1e-9 / (r + 1))
out = par.allreduce_stats_host(st)
assert st["n_points"] == b - a and st["n_plastic"] == r + 1          # the input is not modified
assert par.allreduce_stats_host(st)["n_points"] == out["n_points"]   # ... so repeating the collective changes nothing
assert par.allreduce_stats(None, st)["n_plastic"] == 3               # backend dispatch (gloo -> host tensors)
rec = par.stats_to_record(st); back = par.combine_records(rec[None])
assert back["n_points"] == st["n_points"] and back["f_max"] == st["f_max"] and np.array_equal(back["niter_hist"], hist)
assert out["n_points"] == 1001 and out["n_plastic"] == 3 and out["n_nonconverged"] == 1
assert out["niter_hist"][1] == 10 and out["niter_hist"][2] == 20
assert out["niter_max"] == 4.0 and out["f_max"] == 1.0 and out["res_max"] == 1e-9
dist.destroy_process_group()
print("rank", r, "ok")
"""


def test_stats_allreduce_world2_gloo(tmp_path):
    script = tmp_path / "w.py"
    script.write_text(_GLOO_WORKER.format(root=ROOT))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr",
           "127.0.0.1", "--master-port", str(29400 + os.getpid() % 500), str(script)]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=300)
    assert res.returncode == 0, res.stdout + res.stderr
    assert res.stdout.count("ok") == 2


_GLOO_FORMS_WORKER = r"""
import os, sys
sys.path.insert(0, {root!r}); sys.path.insert(0, os.path.join({root!r}, "tests"))
import numpy as np, torch.distributed as dist
from dolfinx_external_operator_b200 import elements as el, parallel as par
from oracle import forms as of, tabulation as ot
from tab_util import tri_case
dist.init_process_group("gloo")
r, w = dist.get_rank(), dist.get_world_size()
m = tri_case(nx=9, ny=7)
W3 = el.triangle_quadrature_weights(2)
s = np.random.default_rng(0).normal(size=(m["dofmap"].shape[0], 3, 4))      # the same "stress" on every rank
x = np.random.default_rng(1).normal(size=2 * m["n_dofs"])
loc = par.local_submesh(m["dofmap"], m["x_dofmap"], m["x"], r, w)
a, b = loc["cells"]
geo = (loc["x"], loc["x_dofmap"], m["phi"], m["dphi"], m["dpsi"])
b_loc = of.assemble_vector(ot.MANDEL_STRAIN, s[a:b], W3, loc["dofmap"], 2, loc["n_dofs"], *geo)
y_loc = of.apply_action(ot.MANDEL_STRAIN, ot.MANDEL_STRAIN, np.tile(np.eye(4).reshape(-1), (b - a, 3, 1)),
                        x.reshape(-1, 2)[loc["dof_l2g"]].reshape(-1), W3, loc["dofmap"], 2, loc["n_dofs"], *geo)
gb = par.sum_shared(b_loc, loc["dof_l2g"], m["n_dofs"], bs=2)
gy = par.sum_shared(y_loc, loc["dof_l2g"], m["n_dofs"], bs=2)
G = (m["x"], m["x_dofmap"], m["phi"], m["dphi"], m["dpsi"])
ref_b = of.assemble_vector(ot.MANDEL_STRAIN, s, W3, m["dofmap"], 2, m["n_dofs"], *G)
ref_y = of.apply_action(ot.MANDEL_STRAIN, ot.MANDEL_STRAIN, np.tile(np.eye(4).reshape(-1), (m["dofmap"].shape[0], 3, 1)), x,
                        W3, m["dofmap"], 2, m["n_dofs"], *G)
assert np.abs(gb - ref_b).max() <= 1e-13 * np.abs(ref_b).max()
assert np.abs(gy - ref_y).max() <= 1e-13 * np.abs(ref_y).max()
assert loc["n_dofs"] < m["n_dofs"] and (b - a) in (m["dofmap"].shape[0] // 2, m["dofmap"].shape[0] - m["dofmap"].shape[0] // 2)
dist.destroy_process_group()
print("rank", r, "ok")
"""


def test_cell_partitioned_forms_world2_gloo(tmp_path):
    """The device-side consumers shard by cells: each rank integrates its owned block on a locally numbered sub-mesh and
    the interface dofs are summed afterwards (the reference's ghostUpdate(ADD, REVERSE)); the sum equals the global vector."""
    script = tmp_path / "wf.py"
    script.write_text(_GLOO_FORMS_WORKER.format(root=ROOT))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr",
           "127.0.0.1", "--master-port", str(29900 + os.getpid() % 90), str(script)]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=300)
    assert res.returncode == 0, res.stdout + res.stderr
    assert res.stdout.count("ok") == 2
