"""GPU parity tests of the device-side consumers (csrc/form.cu, C ABI eo_form_*, `QuadratureForms`) against the
NumPy oracle (oracle/forms.py).  Floating point with an order-free atomic scatter: rtol 1e-12 of the vector's
largest entry (north_star: 1e-12 for closed-form paths).  The per-point results of the fused residual step are
bit-identical to the fused tabulate + von Mises kernel."""

import numpy as np
import pytest

import dolfinx_external_operator_b200 as eo
from dolfinx_external_operator_b200 import elements as el
from dolfinx_external_operator_b200 import synthetic as syn
from oracle import constitutive as oc
from oracle import forms as of
from oracle import tabulation as ot
from tab_util import tet_case, tri_case

pytestmark = pytest.mark.gpu
KIND = {"value": ot.VALUE, "grad": ot.GRAD, "mandel_strain": ot.MANDEL_STRAIN, "def_grad": ot.DEF_GRAD}
W3 = el.triangle_quadrature_weights(2)


def _mk(ctx, m, bs, weights=W3):
    tab = eo.Tabulator(dofmap=m["dofmap"], x_dofmap=m["x_dofmap"], x=m["x"], phi=m["phi"], dphi=m["dphi"], bs=bs,
                       n_dofs=m["n_dofs"], ctx=ctx)
    return tab, eo.QuadratureForms(tab, weights)


def _geo(m):
    return (m["x"], m["x_dofmap"], m["phi"], m["dphi"], m["dpsi"])


def _close(a, b, rtol=1e-12):
    a, b = np.asarray(a).reshape(-1), np.asarray(b).reshape(-1)
    np.testing.assert_allclose(a, b, rtol=0, atol=rtol * max(np.abs(b).max(), 1e-300))


@pytest.mark.parametrize("kind", ["value", "grad", "mandel_strain", "def_grad"])
@pytest.mark.parametrize("degree", [1, 2])
def test_vector_p1_p2_vector_triangle(ctx, kind, degree):
    m = tri_case(nx=37, ny=23, degree=degree)
    tab, forms = _mk(ctx, m, 2)
    nc = m["dofmap"].shape[0]
    s = np.random.default_rng(0).normal(size=(nc, 3, tab.ncomp(kind)))
    d_s = ctx.to_device(s)
    ref = of.assemble_vector(KIND[kind], s, W3, m["dofmap"], 2, m["n_dofs"], *_geo(m))
    _close(forms.vector(kind, d_s), ref)
    d_b = forms.vector(kind, d_s, output="device")
    _close(d_b.to_host(), ref)
    # accumulate on top of an existing host vector; owned-cell prefix
    b0 = np.random.default_rng(1).normal(size=ref.size)
    b = b0.copy()
    forms.vector(kind, d_s, out=b, accumulate=True)
    _close(b - b0, ref, rtol=1e-11)
    half = of.assemble_vector(KIND[kind], s, W3, m["dofmap"], 2, m["n_dofs"], *_geo(m), n_cells=nc // 3)
    _close(forms.vector(kind, d_s, n_cells=nc // 3), half)


def test_vector_scalar_space_and_exact_integral(ctx):
    m = tri_case(nx=20, ny=15, degree=2)
    tab, forms = _mk(ctx, m, 1)
    xq = m["xq"]
    N = (xq[..., 0] * xq[..., 1] + xq[..., 1] ** 2)[..., None]
    b = forms.vector("value", ctx.to_device(N))
    assert abs(b.sum() - (0.25 + 1.0 / 3.0)) < 1e-13
    _close(b, of.assemble_vector(ot.VALUE, N, W3, m["dofmap"], 1, m["n_dofs"], *_geo(m)))
    q = np.random.default_rng(3).normal(size=xq.shape[:2] + (2,))  # heat flux against grad(v), part2.py:181
    _close(forms.vector("grad", ctx.to_device(q)), of.assemble_vector(ot.GRAD, q, W3, m["dofmap"], 1, m["n_dofs"], *_geo(m)))


def test_vector_and_action_tetrahedra(ctx):
    m = tet_case(4)
    w = np.array([0.1, 1.0 / 6.0 - 0.1])
    tab, forms = _mk(ctx, m, 3, w)
    nc = m["dofmap"].shape[0]
    rng = np.random.default_rng(2)
    s = rng.normal(size=(nc, 2, 9))
    _close(forms.vector("grad", ctx.to_device(s)), of.assemble_vector(ot.GRAD, s, w, m["dofmap"], 3, m["n_dofs"], *_geo(m)))
    D = rng.normal(size=(nc, 2, 81))
    x = rng.normal(size=3 * m["n_dofs"])
    ref = of.apply_action(ot.GRAD, ot.GRAD, D, x, w, m["dofmap"], 3, m["n_dofs"], *_geo(m))
    _close(forms.action("grad", "def_grad", ctx.to_device(D), x), ref)


@pytest.mark.parametrize("pair", [("mandel_strain", "mandel_strain"), ("grad", "value"), ("value", "grad"), ("grad", "grad"),
                                  ("grad", "def_grad", 2), ("mandel_strain", "grad", 2), ("value", "mandel_strain", 2)])
def test_action_p2_triangle(ctx, pair):
    kt, ki = pair[:2]
    # (grad, value) = dq/dT, (grad, grad) = dq/dsigma of the heat demo; (grad, def_grad) on a vector field = dP/dF of the
    # hyperelasticity demo (demo_hyperelasticity.py:479-500), 4x4 like C_tang
    bs = pair[2] if len(pair) > 2 else (2 if "mandel_strain" in pair else 1)
    m = tri_case(nx=31, ny=17)
    tab, forms = _mk(ctx, m, bs)
    nc = m["dofmap"].shape[0]
    rng = np.random.default_rng(4)
    D = rng.normal(size=(nc, 3, tab.ncomp(kt) * tab.ncomp(ki)))
    x = rng.normal(size=bs * m["n_dofs"])
    ref = of.apply_action(KIND[kt], KIND[ki], D, x, W3, m["dofmap"], bs, m["n_dofs"], *_geo(m))
    d_D = ctx.to_device(D)
    _close(forms.action(kt, ki, d_D, x), ref)
    d_y = forms.action(kt, ki, d_D, ctx.to_device(x), output="device")
    _close(d_y.to_host(), ref)
    part = of.apply_action(KIND[kt], KIND[ki], D, x, W3, m["dofmap"], bs, m["n_dofs"], *_geo(m), n_cells=nc // 2)
    _close(forms.action(kt, ki, d_D, x, n_cells=nc // 2), part)


def test_matrix_csr(ctx):
    m = tri_case(nx=13, ny=9)
    tab, forms = _mk(ctx, m, 2)
    nc = m["dofmap"].shape[0]
    rng = np.random.default_rng(6)
    D = rng.normal(size=(nc, 3, 16))
    rp, col = forms.set_pattern()
    ref = of.assemble_matrix(ot.MANDEL_STRAIN, ot.MANDEL_STRAIN, D, W3, m["dofmap"], 2, m["n_dofs"], *_geo(m), rp, col)
    d_D = ctx.to_device(D)
    vals = forms.matrix("mandel_strain", "mandel_strain", d_D)
    _close(vals.to_host(), ref)
    forms.matrix("mandel_strain", "mandel_strain", d_D, vals=vals, accumulate=True)
    _close(vals.to_host(), 2 * ref)
    # subset of cells on the cached positions; a pattern that misses an element entry is an error
    part = of.assemble_matrix(ot.MANDEL_STRAIN, ot.MANDEL_STRAIN, D[:nc // 2], W3, m["dofmap"][:nc // 2], 2, m["n_dofs"],
                              m["x"], m["x_dofmap"][:nc // 2], m["phi"], m["dphi"], m["dpsi"], rp, col)
    _close(forms.matrix("mandel_strain", "mandel_strain", d_D, n_cells=nc // 2).to_host(), part)
    tab2, forms2 = _mk(ctx, m, 2)
    keep = np.ones(col.size, dtype=bool)
    keep[rp[5] + 1] = False  # drop one off-diagonal entry of row 5
    rp_bad = np.concatenate([[0], np.cumsum(np.add.reduceat(keep.astype(np.int64), rp[:-1]))]).astype(np.int32)
    forms2.set_pattern(rp_bad, col[keep])
    with pytest.raises(eo.EOError):
        forms2.matrix("mandel_strain", "mandel_strain", d_D)
    # the assembled matrix and the matrix-free action are the same operator
    x = rng.normal(size=2 * m["n_dofs"])
    rows = np.repeat(np.arange(rp.size - 1), np.diff(rp))
    y = np.zeros_like(x)
    np.add.at(y, rows, 0.5 * vals.to_host() * x[col])
    _close(forms.action("mandel_strain", "mandel_strain", d_D, x), y, rtol=1e-11)


@pytest.mark.parametrize("exact", [True, False])
def test_vm_residual_step(ctx, exact):
    m = tri_case(nx=41, ny=29)
    tab, forms = _mk(ctx, m, 2)
    n = m["dofmap"].shape[0] * 3
    rng = np.random.default_rng(7)
    sigma_n, p = rng.normal(0.0, 100.0, (n, 4)), np.abs(rng.normal(0.0, 1e-3, n))
    u = syn.smooth_displacement(m["dof_coords"], scale=6e-4, seed=3).reshape(-1)  # ~50 % plastic points
    vm = eo.VonMises(ctx=ctx, n_qp=n)
    vm.set_history(sigma_n, p)
    ctx.stats_reset()
    b = forms.vm_residual(vm, u, exact=exact)
    Ct, sig, dp = forms.C_tang.to_host(), vm.sigma_dev.to_host(), vm.dp_dev.to_host()
    st = ctx.stats()
    # per-point results: identical to the fused tabulate + von Mises kernel
    vm2 = eo.VonMises(ctx=ctx, n_qp=n)
    vm2.set_history(sigma_n, p)
    Ct2 = tab.vm_fused(vm2, u, exact=exact).to_host()
    assert np.array_equal(Ct, Ct2) and np.array_equal(sig, vm2.sigma_dev.to_host()) and np.array_equal(dp, vm2.dp_dev.to_host())
    # ... and within 1e-12 of the oracle, flags bit-exact
    eps = ot.tabulate(ot.MANDEL_STRAIN, u, m["dofmap"], 2, *_geo(m)).reshape(-1, 4)
    rCt, rsig, rdp = oc.vm_return_mapping(eps, sigma_n, p)
    assert np.array_equal(dp > 0, np.asarray(rdp).reshape(-1) > 0) and 0.2 < (dp > 0).mean() < 0.8
    assert st["n_plastic"] == int((dp > 0).sum()) and st["n_points"] == n
    _close(Ct, rCt, 1e-12), _close(sig, rsig, 1e-12)
    _close(b, of.assemble_vector(ot.MANDEL_STRAIN, rsig, W3, m["dofmap"], 2, m["n_dofs"], *_geo(m)))
    # residual step == separate stress integral; tangent action against the oracle with the kernel's own tangent
    _close(forms.vector("mandel_strain", vm.sigma_dev), b)
    x = rng.normal(size=u.size)
    _close(forms.action("mandel_strain", "mandel_strain", forms.C_tang, x),
           of.apply_action(ot.MANDEL_STRAIN, ot.MANDEL_STRAIN, Ct, x, W3, m["dofmap"], 2, m["n_dofs"], *_geo(m)))


@pytest.mark.parametrize("fused", [False, True])
def test_mc_residual_step(ctx, fused):
    """Mohr-Coulomb: tabulate -> local Newton -> stress integral on the device vs the oracle chain (1e-10: Newton model);
    fused=True: the strain is tabulated inside pass 1 of the Mohr-Coulomb kernels and never stored."""
    from oracle import native

    m = tri_case(nx=23, ny=17)
    tab, forms = _mk(ctx, m, 2)
    n = m["dofmap"].shape[0] * 3
    mprm = oc.MohrCoulombParams()
    _, sigma_n = syn.mc_batch(n, seed=5, stepper=lambda d, s: native.mc_stress(d, s, mprm, parallel=True)[0])
    u = syn.smooth_displacement(m["dof_coords"], scale=2e-6, seed=3).reshape(-1)
    mc = eo.MohrCoulomb(ctx=ctx, n_qp=n)
    mc.set_history(sigma_n)
    b = forms.mc_residual(mc, u, fused=fused)
    eps = ot.tabulate(ot.MANDEL_STRAIN, u, m["dofmap"], 2, *_geo(m)).reshape(-1, 4)
    ref = native.mc_return_mapping(eps, sigma_n, mprm, parallel=True)
    assert 0.05 < (np.asarray(ref["yielding"]) > 0).mean() < 0.95 and int(np.max(ref["niter"])) < 20
    _close(mc.sigma_dev.to_host(), ref["sigma"], 1e-10)
    _close(b, of.assemble_vector(ot.MANDEL_STRAIN, ref["sigma"], W3, m["dofmap"], 2, m["n_dofs"], *_geo(m)), 1e-10)
    x = np.random.default_rng(1).normal(size=u.size)
    Ct = forms.C_tang.to_host()
    _close(forms.action("mandel_strain", "mandel_strain", forms.C_tang, x),
           of.apply_action(ot.MANDEL_STRAIN, ot.MANDEL_STRAIN, Ct, x, W3, m["dofmap"], 2, m["n_dofs"], *_geo(m)))
    mc.commit()
    _close(mc.sigma_n_dev.to_host(), ref["sigma"], 1e-10)


def test_heat_demo_forms_match_the_explicit_forms(ctx):
    """demo_nonlinear_heat_equation_part2.py:283-300 on the device: tabulate T and grad T, evaluate q / dq/dT / dq/dsigma
    (eo_heat_eval), integrate b = int q . grad(v) and assemble J (test grad x trial value + test grad x trial grad) into
    CSR; compared like the reference does with the explicit ("pure UFL") forms, and with the oracle chain."""
    from heat_util import explicit_heat_forms
    from test_heat_forms_cpu import dense, heat_case, oracle_heat_forms

    m, T = heat_case()
    tab, forms = _mk(ctx, m, 1)
    nc = m["dofmap"].shape[0]
    d_T = ctx.to_device(T)
    Tq, sq = tab.evaluate("value", d_T), tab.evaluate("grad", d_T)
    q, dT, ds = ctx.empty((nc, 3, 2)), ctx.empty((nc, 3, 2)), ctx.empty((nc, 3, 4))
    eo.HeatFlux(ctx=ctx).eval_device(Tq, sq, q, dT, ds)
    b = forms.vector("grad", q)
    rp, col = forms.set_pattern()
    vals = forms.matrix("grad", "value", dT)
    forms.matrix("grad", "grad", ds, vals=vals, accumulate=True)
    b_ex, A_ex = explicit_heat_forms(m, T)
    _close(b, b_ex), _close(dense(vals.to_host(), rp, col), A_ex)
    b_or, v_or, rp2, col2 = oracle_heat_forms(m, T)
    assert np.array_equal(rp, rp2) and np.array_equal(col, col2)
    _close(b, b_or), _close(vals.to_host(), v_or)
    # the same Jacobian matrix-free: two accumulated actions
    x = np.random.default_rng(2).normal(size=T.size)
    y = forms.action("grad", "value", dT, x)
    forms.action("grad", "grad", ds, x, out=y, accumulate=True)
    _close(y, A_ex @ x, 1e-11)


def test_errors_and_empty(ctx):
    m = tri_case(nx=4, ny=3)
    tab, forms = _mk(ctx, m, 2)
    nc = m["dofmap"].shape[0]
    with pytest.raises(ValueError):
        eo.QuadratureForms(tab, np.ones(2))
    with pytest.raises(TypeError):
        forms.vector("grad", np.zeros((nc, 3, 4)))  # point values must be resident
    with pytest.raises(ValueError):
        forms.vector("grad", ctx.zeros((nc, 3, 3)))
    with pytest.raises(ValueError):
        forms.action("grad", "grad", ctx.zeros((nc, 3, 16)), np.zeros(5))
    x = ctx.zeros((2 * m["n_dofs"],))
    with pytest.raises(eo.EOError):
        forms.action("grad", "grad", ctx.zeros((nc, 3, 16)), x, out=x)  # aliasing
    with pytest.raises(eo.EOError):
        ctx.check(ctx.lib.eo_form_matrix(forms._h, 1, 1, ctx.zeros((nc, 3, 16)).ptr, -1, x.ptr, 0))  # no pattern yet
    b = forms.vector("grad", ctx.zeros((nc, 3, 4)), n_cells=0)
    assert b.shape == (2 * m["n_dofs"],) and not b.any()
    mt, ft = _mk(ctx, tri_case(nx=4, ny=3, degree=1), 1)
    with pytest.raises(ValueError):
        ft.vector("mandel_strain", ctx.zeros((10,)))


def test_full_size_properties(ctx):
    """1e7 points: adjointness with the tabulation kernel, rigid-body modes of the elastic stiffness."""
    m = syn.triangle_mesh(1291, 1291, 2)  # 3 333 362 cells, 1.0e7 points
    X = el.triangle_quadrature(2)
    phi, dphi = el.lagrange_triangle(2, X)
    m.update(phi=phi, dphi=dphi)
    tab, forms = _mk(ctx, m, 2)
    nc = m["dofmap"].shape[0]
    n = nc * 3
    rng = np.random.default_rng(8)
    u = rng.normal(size=2 * m["n_dofs"])
    s = rng.normal(size=(n, 4))
    d_s = ctx.to_device(s)
    eps = tab.evaluate("mandel_strain", u, output="host").reshape(n, 4)
    area = 1.0 / (2 * 1291 * 1291)  # |det J| / 2 ... every cell is congruent: w_q |det J| = area / 3
    lhs = (eps * s).sum() * area / 3.0
    b = forms.vector("mandel_strain", d_s)
    assert abs(lhs - u @ b) < 1e-11 * abs(lhs)
    Ce = oc.elastic_stiffness(oc.VonMisesParams().lmbda, oc.VonMisesParams().mu)
    d_D = ctx.to_device(np.broadcast_to(Ce.reshape(-1), (n, 16)).copy())
    xy = m["dof_coords"]
    scale = np.abs(forms.action("mandel_strain", "mandel_strain", d_D, u)).max()
    rot = np.stack([-xy[:, 1], xy[:, 0]], 1).reshape(-1)
    assert np.abs(forms.action("mandel_strain", "mandel_strain", d_D, rot)).max() < 1e-11 * scale


# ------------------------------------------------------------------ nb = 10 instantiations, renumbered meshes
def test_forms_p3_triangles(ctx):
    """form kernels <2,1,10> / <2,2,10>: vector, action (general + 4x4 TMA path), residual step vs the oracle."""
    from tab_util import tri_case_discontinuous

    m = tri_case_discontinuous(degree=3, nx=11, ny=8)
    nc = m["dofmap"].shape[0]
    rng = np.random.default_rng(4)
    for bs, kt, ki in ((2, "mandel_strain", "mandel_strain"), (1, "grad", "value"), (2, "grad", "def_grad")):
        tab, forms = _mk(ctx, m, bs)
        s = rng.normal(size=(nc, 3, tab.ncomp(kt)))
        _close(forms.vector(kt, ctx.to_device(s)), of.assemble_vector(KIND[kt], s, W3, m["dofmap"], bs, m["n_dofs"], *_geo(m)))
        D = rng.normal(size=(nc, 3, tab.ncomp(kt) * tab.ncomp(ki)))
        x = rng.normal(size=bs * m["n_dofs"])
        ref = of.apply_action(KIND[kt], KIND[ki], D, x, W3, m["dofmap"], bs, m["n_dofs"], *_geo(m))
        _close(forms.action(kt, ki, ctx.to_device(D), x), ref)
    # residual step on P3: per-point results identical to the fused kernel, vector vs oracle
    tab, forms = _mk(ctx, m, 2)
    n = 3 * nc
    u = syn.smooth_displacement(m["dof_coords"], scale=1.5e-3, seed=1).reshape(-1)
    _, sn, p = syn.vm_batch(n, seed=2)
    vm_a, vm_b = eo.VonMises(ctx=ctx), eo.VonMises(ctx=ctx)
    vm_a.set_history(sn, p)
    vm_b.set_history(sn, p)
    b = forms.vm_residual(vm_a, u, exact=True)
    Ct = tab.vm_fused(vm_b, u, exact=True)
    assert np.array_equal(forms.C_tang.to_host(), Ct.to_host()) and np.array_equal(vm_a.sigma_dev.to_host(), vm_b.sigma_dev.to_host())
    ref = of.assemble_vector(ot.MANDEL_STRAIN, vm_b.sigma_dev.to_host().reshape(nc, 3, 4), W3, m["dofmap"], 2, m["n_dofs"], *_geo(m))
    _close(b, ref)


def test_forms_p2_tetrahedra(ctx):
    """form kernels <3,1,10> / <3,3,10>: P : grad(v) and the dP/dF action of a 3-d residual / Jacobian."""
    from tab_util import tet_case_discontinuous

    m = tet_case_discontinuous(n=3)
    w = np.array([0.05, 0.04, 0.03, 1.0 / 6.0 - 0.12])
    rng = np.random.default_rng(5)
    nc, nq = m["dofmap"].shape[0], m["phi"].shape[0]
    tab, forms = _mk(ctx, m, 3, w)
    P = rng.normal(size=(nc, nq, 9))
    _close(forms.vector("grad", ctx.to_device(P)), of.assemble_vector(ot.GRAD, P, w, m["dofmap"], 3, m["n_dofs"], *_geo(m)))
    D = rng.normal(size=(nc, nq, 81))
    x = rng.normal(size=3 * m["n_dofs"])
    _close(forms.action("grad", "def_grad", ctx.to_device(D), x),
           of.apply_action(ot.GRAD, ot.DEF_GRAD, D, x, w, m["dofmap"], 3, m["n_dofs"], *_geo(m)))
    tab1, forms1 = _mk(ctx, m, 1, w)
    q = rng.normal(size=(nc, nq, 3))
    _close(forms1.vector("grad", ctx.to_device(q)), of.assemble_vector(ot.GRAD, q, w, m["dofmap"], 1, m["n_dofs"], *_geo(m)))
    K = rng.normal(size=(nc, nq, 9))
    T = rng.normal(size=m["n_dofs"])
    _close(forms1.action("grad", "grad", ctx.to_device(K), T),
           of.apply_action(ot.GRAD, ot.GRAD, K, T, w, m["dofmap"], 1, m["n_dofs"], *_geo(m)))


@pytest.mark.parametrize("order", ["shuffled", "rcm"])
def test_forms_on_renumbered_mesh(ctx, order):
    """Residual step and tangent action under a random / RCM numbering: vectors equal the structured ones up to the dof
    permutation (atomic scatter: to rounding), per-point results bit for bit."""
    m = tri_case(nx=29, ny=31)
    r = syn.renumber(m, order, seed=5)
    nc = m["dofmap"].shape[0]
    n = 3 * nc
    u = syn.smooth_displacement(m["dof_coords"], scale=1.5e-3, seed=3)
    ur = np.empty_like(u)
    ur[r["dof_new"]] = u
    _, sn, p = syn.vm_batch(n, seed=6)
    co = r["cell_old"]
    sn_r = sn.reshape(nc, 3, 4)[co].reshape(-1, 4)
    p_r = p.reshape(nc, 3)[co].reshape(-1)
    tab_s, forms_s = _mk(ctx, m, 2)
    tab_r, forms_r = _mk(ctx, r, 2)
    vm_s, vm_r = eo.VonMises(ctx=ctx), eo.VonMises(ctx=ctx)
    vm_s.set_history(sn, p)
    vm_r.set_history(sn_r, p_r)
    b_s = forms_s.vm_residual(vm_s, u.reshape(-1), exact=True).copy()
    b_r = forms_r.vm_residual(vm_r, ur.reshape(-1), exact=True).copy()
    assert np.array_equal(forms_r.C_tang.to_host().reshape(nc, -1), forms_s.C_tang.to_host().reshape(nc, -1)[co])
    _close(b_r.reshape(-1, 2), b_s.reshape(-1, 2)[np.argsort(r["dof_new"])], 1e-12)
    x = np.random.default_rng(7).normal(size=u.shape)
    xr = np.empty_like(x)
    xr[r["dof_new"]] = x
    y_s = forms_s.action("mandel_strain", "mandel_strain", forms_s.C_tang, x.reshape(-1)).copy()
    y_r = forms_r.action("mandel_strain", "mandel_strain", forms_r.C_tang, xr.reshape(-1)).copy()
    _close(y_r.reshape(-1, 2), y_s.reshape(-1, 2)[np.argsort(r["dof_new"])], 1e-12)
    ref = of.apply_action(ot.MANDEL_STRAIN, ot.MANDEL_STRAIN, forms_r.C_tang.to_host().reshape(nc, 3, 16), xr.reshape(-1), W3,
                          r["dofmap"], 2, r["n_dofs"], r["x"], r["x_dofmap"], r["phi"], r["dphi"], r["dpsi"])
    _close(y_r, ref)


@pytest.mark.parametrize("order", ["structured", "rcm", "shuffled"])
def test_host_vector_pipeline_equals_device_path(ctx, order):
    """Host vectors in / out go through the dependency-aware chunk pipeline (16 cell chunks x 64 vector pieces, three
    streams) once the mesh has >= 2^17 cells; device vectors take the single-launch path.  Same residual, same tangent
    action (to the rounding of the atomic scatter), same per-point results, whatever the numbering - a shuffled mesh
    makes every chunk depend on every piece (the serial schedule)."""
    m = tri_case(nx=300, ny=230)
    if order != "structured":
        m = syn.renumber(m, order, seed=1)
    nc = m["dofmap"].shape[0]
    assert nc >= 1 << 17
    n = 3 * nc
    u = syn.smooth_displacement(m["dof_coords"], scale=1.5e-3, seed=2).reshape(-1)
    _, sn, p = syn.vm_batch(n, seed=3)
    tab, forms = _mk(ctx, m, 2)
    tab2, forms2 = _mk(ctx, m, 2)
    vm_h, vm_d = eo.VonMises(ctx=ctx), eo.VonMises(ctx=ctx)
    vm_h.set_history(sn, p)
    vm_d.set_history(sn, p)
    u_pin = ctx.pinned_empty(u.shape)
    u_pin[:] = u
    b_h = forms.vm_residual(vm_h, u_pin, exact=True).copy()                      # host in, host out: pipeline
    b_d = forms2.vm_residual(vm_d, ctx.to_device(u), exact=True, output="device").to_host()  # device: one launch
    assert np.array_equal(forms.C_tang.to_host(), forms2.C_tang.to_host())
    assert np.array_equal(vm_h.sigma_dev.to_host(), vm_d.sigma_dev.to_host())
    _close(b_h, b_d)
    b_pageable = forms.vm_residual(vm_h, u.copy(), exact=True).copy()             # pageable host memory: same result
    _close(b_pageable, b_d)
    x = np.random.default_rng(5).normal(size=u.shape)
    y_h = forms.action("mandel_strain", "mandel_strain", forms.C_tang, x).copy()
    y_d = forms2.action("mandel_strain", "mandel_strain", forms2.C_tang, ctx.to_device(x), output="device").to_host()
    _close(y_h, y_d)
    assert ctx.stats()["n_points"] >= n


@pytest.mark.parametrize("exact", [False, True])
@pytest.mark.parametrize("degree", [1, 2])
def test_vm_factored_tangent(ctx, exact, degree):
    """Device-side consumers with the von Mises tangent kept as 6 numbers per point (v, cn, cd): the residual, stress, dp and
    statistics are those of the ordinary step bit for bit, the expanded factors are the ordinary tangent bit for bit, and
    the tangent action rebuilt from the factors agrees with the action on the full tangent and with the oracle."""
    m = tri_case(nx=37, ny=23, degree=degree)
    tab, forms = _mk(ctx, m, 2)
    n = m["dofmap"].shape[0] * 3
    rng = np.random.default_rng(17)
    sigma_n, p = rng.normal(0.0, 100.0, (n, 4)), np.abs(rng.normal(0.0, 1e-3, n))
    u = syn.smooth_displacement(m["dof_coords"], scale=6e-4, seed=4).reshape(-1)
    vm = eo.VonMises(ctx=ctx, n_qp=n)
    vm.set_history(sigma_n, p)
    b_full = forms.vm_residual(vm, u, exact=exact)
    Ct, sig, dp = forms.C_tang.to_host(), vm.sigma_dev.to_host(), vm.dp_dev.to_host()
    assert 0.2 < (dp > 0).mean() < 0.8
    vm2 = eo.VonMises(ctx=ctx, n_qp=n)
    vm2.set_history(sigma_n, p)
    ctx.stats_reset()
    b_fact = forms.vm_residual(vm2, u, exact=exact, tangent="factored")
    st = ctx.stats()
    assert st["n_points"] == n and st["n_plastic"] == int((dp > 0).sum())
    assert np.array_equal(sig, vm2.sigma_dev.to_host()) and np.array_equal(dp, vm2.dp_dev.to_host())
    _close(b_fact, b_full, 1e-13)
    # the factors: elastic points carry v = 0 and cd = 0; expanded, they are the stored tangent
    T6 = forms.T6.to_host().reshape(n, 6)
    assert np.all(T6[dp == 0, :4] == 0.0) and np.all(T6[dp == 0, 5] == 0.0)
    assert np.array_equal(forms.expand_tangent(ctx.empty((16 * n,))).to_host(), Ct)
    # the action from the factors == the action on the full tangent == the oracle
    x = rng.normal(size=u.size)
    y_fact = forms.vm_action(x)
    y_full = forms.action("mandel_strain", "mandel_strain", forms.C_tang, x)
    _close(y_fact, y_full, 1e-12)
    _close(y_fact, of.apply_action(ot.MANDEL_STRAIN, ot.MANDEL_STRAIN, Ct, x, W3, m["dofmap"], 2, m["n_dofs"], *_geo(m)))
    # host vectors through the chunk pipeline and device-resident vectors give the same action
    d_x, d_y = ctx.to_device(x), ctx.empty((u.size,))
    forms.vm_action(d_x, out=d_y)
    _close(d_y.to_host(), y_fact, 1e-13)
    # accumulate: y += J x
    y2 = forms.vm_action(x, out=y_fact.copy(), accumulate=True)
    _close(y2, 2.0 * y_fact, 1e-13)


def test_vm_factored_errors(ctx):
    m = tri_case()
    tab, forms = _mk(ctx, m, 2)
    with pytest.raises(RuntimeError):
        forms.vm_action(np.zeros(2 * m["n_dofs"]))
    with pytest.raises(ValueError):
        forms.vm_residual(eo.VonMises(ctx=ctx, n_qp=m["dofmap"].shape[0] * 3), np.zeros(2 * m["n_dofs"]), tangent="packed")
