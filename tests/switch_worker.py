"""Worker of tests/test_switches_gpu.py: evaluates the cell-loop kernels on one fixed problem and stores every result, so
that two processes started with different EO_GEOM_CACHE / EO_QUAD_STORE settings (read once per process by the library)
can be compared.  Usage: python tests/switch_worker.py OUT.npz"""

import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import dolfinx_external_operator_b200 as eo  # noqa: E402
from dolfinx_external_operator_b200 import elements as el  # noqa: E402
from dolfinx_external_operator_b200 import synthetic as syn  # noqa: E402
from tab_util import tri_case  # noqa: E402


def main(out):
    ctx = eo.Context(0)
    # 2 x 37 x 23 triangles x 3 points = 5106 points: not a multiple of 4, 32 or 256 (tail quads, tail warps, tail tiles)
    m = tri_case(nx=37, ny=23)
    tab = eo.Tabulator(dofmap=m["dofmap"], x_dofmap=m["x_dofmap"], x=m["x"], phi=m["phi"], dphi=m["dphi"], bs=2,
                       n_dofs=m["n_dofs"], ctx=ctx)
    forms = eo.QuadratureForms(tab, el.triangle_quadrature_weights(2))
    n = m["dofmap"].shape[0] * 3
    rng = np.random.default_rng(11)
    sigma_n, p = rng.normal(0.0, 100.0, (n, 4)), np.abs(rng.normal(0.0, 1e-3, n))
    u = syn.smooth_displacement(m["dof_coords"], scale=6e-4, seed=5).reshape(-1)
    res = {}
    for kind in ("grad", "mandel_strain", "def_grad"):
        res["tab_" + kind] = tab.evaluate(kind, u, output="host")
    some = np.arange(0, m["dofmap"].shape[0], 3, dtype=np.int32)
    res["tab_subset"] = tab.evaluate("mandel_strain", u, entities=some, output="host")
    for exact in (False, True):
        vm = eo.VonMises(ctx=ctx, n_qp=n)
        vm.set_history(sigma_n, p)
        res[f"fused_Ct_{int(exact)}"] = tab.vm_fused(vm, u, exact=exact).to_host()
        res[f"fused_sig_{int(exact)}"] = vm.sigma_dev.to_host()
        res[f"fused_dp_{int(exact)}"] = vm.dp_dev.to_host()
        vm2 = eo.VonMises(ctx=ctx, n_qp=n)
        vm2.set_history(sigma_n, p)
        res[f"step_b_{int(exact)}"] = forms.vm_residual(vm2, u, exact=exact)
        res[f"step_Ct_{int(exact)}"] = forms.C_tang.to_host()
        res[f"step_sig_{int(exact)}"] = vm2.sigma_dev.to_host()
    x = rng.normal(size=u.size)
    res["action"] = forms.action("mandel_strain", "mandel_strain", forms.C_tang, x)
    res["vector"] = forms.vector("mandel_strain", vm2.sigma_dev)
    np.savez(out, **res)


if __name__ == "__main__":
    main(sys.argv[1])
