"""Timings (CUDA events) of the consumer kernels that have no bench leg: eo_form_vector alone, eo_form_matrix, mc_residual."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import dolfinx_external_operator_b200 as eo
from dolfinx_external_operator_b200 import elements as el, synthetic as syn

ctx = eo.Context(0)
def mesh(n):
    nxy = int(round((n / 6.0) ** 0.5))
    m = syn.triangle_mesh(nxy, nxy, 2, jitter=0.2, seed=0)
    phi, dphi = el.lagrange_triangle(2, el.triangle_quadrature(2))
    tab = eo.Tabulator(dofmap=m["dofmap"], x_dofmap=m["x_dofmap"], x=m["x"], phi=phi, dphi=dphi, bs=2, n_dofs=m["n_dofs"], ctx=ctx)
    return m, tab, eo.QuadratureForms(tab, el.triangle_quadrature_weights(2))

def timeit(fn, reps=5):
    fn(); ctx.sync()
    e0, e1 = ctx.event(), ctx.event()
    ctx.record(e0)
    for _ in range(reps): fn()
    ctx.record(e1); ctx.sync()
    return ctx.elapsed_ms(e0, e1) / reps

m, tab, forms = mesh(1e8)
n = 3 * tab.n_cells
sig = ctx.zeros((n * 4,)); b = ctx.empty((2 * tab.n_dofs,))
t = timeit(lambda: forms.vector("mandel_strain", sig, out=b))
print(f"form_vector n={n}: {t:.3f} ms  {n/t/1e6:.2f} G QP/s  {(32 + 48/3 + 64/3)*n/t/1e6:.0f} GB/s of 69.3 B/QP")
del sig, b, forms, tab, m

m, tab, forms = mesh(2e7)
n = 3 * tab.n_cells
D = ctx.zeros((n * 16,))
forms.set_pattern()
vals = ctx.empty((forms.col.size,))
t = timeit(lambda: forms.matrix("mandel_strain", "mandel_strain", D, vals=vals), reps=3)
print(f"form_matrix n={n} nnz={forms.col.size}: {t:.3f} ms  {n/t/1e6:.3f} G QP/s")
from oracle import native, constitutive as oc
mprm = oc.MohrCoulombParams()
mc = eo.MohrCoulomb(ctx=ctx, n_qp=n, aux=False)
_, sn = syn.mc_batch(1 << 20, seed=0, stepper=lambda d, s: mc.stress_update(d, s))
mc.set_history(np.resize(sn, (n, 4)))
u = ctx.to_device(syn.smooth_displacement(m["dof_coords"], scale=2e-6, seed=3).reshape(-1))
b = ctx.empty((2 * tab.n_dofs,))
t = timeit(lambda: forms.mc_residual(mc, u, out=b), reps=3)
st = ctx.stats()
print(f"mc_residual n={n}: {t:.3f} ms  {n/t/1e6:.3f} G QP/s  plastic {st['n_plastic']/max(st['n_points'],1):.3f}")
