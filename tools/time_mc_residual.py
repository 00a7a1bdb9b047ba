"""tools/time_mc_residual.py - device-resident Mohr-Coulomb residual step (QuadratureForms.mc_residual) at ~2e7 points:
fused (strain tabulated inside pass 1) against the three-launch chain tabulate -> eo_mc_eval -> eo_form_vector."""
import ctypes as C
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import dolfinx_external_operator_b200 as eo
from dolfinx_external_operator_b200 import elements as el, synthetic as syn

ctx = eo.Context(0)
n_target = float(sys.argv[1]) if len(sys.argv) > 1 else 2e7
nxy = int(round((n_target / 6.0) ** 0.5))
m = syn.triangle_mesh(nxy, nxy, 2, jitter=0.2, seed=0)
phi, dphi = el.lagrange_triangle(2, el.triangle_quadrature(2))
tab = eo.Tabulator(dofmap=m["dofmap"], x_dofmap=m["x_dofmap"], x=m["x"], phi=phi, dphi=dphi, bs=2, n_dofs=m["n_dofs"], ctx=ctx)
forms = eo.QuadratureForms(tab, el.triangle_quadrature_weights(2))
n = 3 * m["dofmap"].shape[0]
mc = eo.MohrCoulomb(ctx=ctx, n_qp=n)
_, sn = syn.mc_batch(1 << 20, seed=0, stepper=mc.stress_update)
mc.set_history(np.resize(sn, (n, 4)))
d_u = ctx.to_device(syn.smooth_displacement(m["dof_coords"], scale=2e-6, seed=3).reshape(-1))
d_b = ctx.empty((2 * tab.n_dofs,))
strain = ctx.empty((tab.n_cells, 3, 4))


def chain():
    tab.evaluate("mandel_strain", d_u, out=strain)
    ctx.check(ctx.lib.eo_mc_eval(ctx.handle, C.byref(mc._prm), strain.ptr, mc.sigma_n_dev.ptr, forms.C_tang.ptr, mc.sigma_dev.ptr,
                                 None, None, None, None, n))
    forms.vector("mandel_strain", mc.sigma_dev, out=d_b)


for name, fn in (("fused", lambda: forms.mc_residual(mc, d_u, out=d_b)), ("three launches", chain)):
    for _ in range(3):
        fn()
    e0, e1 = ctx.event(), ctx.event()
    ctx.sync()
    ctx.record(e0)
    for _ in range(5):
        fn()
    ctx.record(e1)
    ctx.sync()
    st = ctx.stats()
    print(f"{name}: {ctx.elapsed_ms(e0, e1) / 5:.3f} ms per step, n = {n}, plastic fraction {st['n_plastic'] / max(st['n_points'], 1):.3f}")
