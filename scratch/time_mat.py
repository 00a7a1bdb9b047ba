import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import dolfinx_external_operator_b200 as eo
from dolfinx_external_operator_b200 import elements as el, synthetic as syn
ctx = eo.Context(0)
nxy = int(round((2e6 / 6.0) ** 0.5))
m = syn.triangle_mesh(nxy, nxy, 2, jitter=0.2, seed=0)
phi, dphi = el.lagrange_triangle(2, el.triangle_quadrature(2))
tab = eo.Tabulator(dofmap=m["dofmap"], x_dofmap=m["x_dofmap"], x=m["x"], phi=phi, dphi=dphi, bs=2, n_dofs=m["n_dofs"], ctx=ctx)
n = 3 * tab.n_cells
D = ctx.zeros((n * 16,))
for v in ("0", "1"):
    os.environ["EO_FORM_MATRIX_POS"] = v
    forms = eo.QuadratureForms(tab, el.triangle_quadrature_weights(2))
    forms.set_pattern()
    vals = ctx.empty((forms.col.size,))
    fn = lambda: forms.matrix("mandel_strain", "mandel_strain", D, vals=vals)
    fn(); fn(); ctx.sync()
    e0, e1 = ctx.event(), ctx.event()
    ctx.record(e0)
    for _ in range(5): fn()
    ctx.record(e1); ctx.sync()
    t = ctx.elapsed_ms(e0, e1) / 5
    print(f"form_matrix pos_cache={v} n={n} nnz={forms.col.size}: {t:.3f} ms  {n/t/1e6:.3f} G QP/s")
