import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import dolfinx_external_operator_b200 as eo
from dolfinx_external_operator_b200 import elements as el, synthetic as syn
ctx = eo.Context(0)
nxy = int(round((1e8 / 6.0) ** 0.5))
m = syn.triangle_mesh(nxy, nxy, 2, jitter=0.2, seed=0)
phi, dphi = el.lagrange_triangle(2, el.triangle_quadrature(2))
tab = eo.Tabulator(dofmap=m["dofmap"], x_dofmap=m["x_dofmap"], x=m["x"], phi=phi, dphi=dphi, bs=2, n_dofs=m["n_dofs"], ctx=ctx)
forms = eo.QuadratureForms(tab, el.triangle_quadrature_weights(2))
n = 3 * tab.n_cells
sig = ctx.zeros((n * 4,)); b = ctx.empty((2 * tab.n_dofs,))
for v in ("1", "0"):
    os.environ["EO_FORM_VECTOR_CELL"] = v
    fn = lambda: forms.vector("mandel_strain", sig, out=b)
    fn(); ctx.sync()
    e0, e1 = ctx.event(), ctx.event()
    ctx.record(e0)
    for _ in range(10): fn()
    ctx.record(e1); ctx.sync()
    t = ctx.elapsed_ms(e0, e1) / 10
    print(f"form_vector per_cell={v} n={n}: {t:.3f} ms  {n/t/1e6:.2f} G QP/s  {(32 + 48/3 + 64/3)*n/t/1e6:.0f} GB/s of 69.3 B/QP")
