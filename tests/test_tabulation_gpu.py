"""GPU parity tests for the operand tabulation kernels (hot path (a)) through the C ABI / Tabulator,
against the NumPy oracle (oracle/tabulation.py; parity unpinned against DOLFINx itself - see DESIGN.md) and
against analytic fields; the fused tabulate + von Mises kernel against the two-step path; and the drop-in
`evaluate_operands` -> `evaluate_external_operators` flow of demo_plasticity_von_mises.py:445-456 with
duck-typed operators."""

import numpy as np
import pytest

import dolfinx_external_operator_b200 as eo
from dolfinx_external_operator_b200 import synthetic as syn
from oracle import constitutive as oc
from oracle import native
from oracle import tabulation as ot
from tab_util import tet_case, tri_case

pytestmark = pytest.mark.gpu
KIND = {"value": ot.VALUE, "grad": ot.GRAD, "mandel_strain": ot.MANDEL_STRAIN, "def_grad": ot.DEF_GRAD}


def _mk(ctx, m, bs, coefficient=None):
    return eo.Tabulator(dofmap=m["dofmap"], x_dofmap=m["x_dofmap"], x=m["x"], phi=m["phi"], dphi=m["dphi"], bs=bs,
                        n_dofs=m["n_dofs"], coefficient=coefficient, ctx=ctx)


def _ref(m, kind, u, bs, cells=None):
    return ot.tabulate(KIND[kind], u, m["dofmap"], bs, m["x"], m["x_dofmap"], m["phi"], m["dphi"], m["dpsi"], cells)


def _close(a, b, rtol=1e-12):
    np.testing.assert_allclose(np.asarray(a).reshape(-1), np.asarray(b).reshape(-1), rtol=rtol,
                               atol=rtol * max(np.abs(b).max(), 1e-300))


@pytest.mark.parametrize("kind", ["value", "grad", "mandel_strain", "def_grad"])
@pytest.mark.parametrize("output", ["host", "device"])
def test_p2_vector_triangle(ctx, kind, output):
    m = tri_case(nx=37, ny=23)
    u = syn.smooth_displacement(m["dof_coords"], seed=3).reshape(-1)
    tab = _mk(ctx, m, 2)
    out = tab.evaluate(kind, u, output=output)
    ref = _ref(m, kind, u, 2)
    assert tuple(out.shape) == {"value": ref.shape, "mandel_strain": ref.shape}.get(kind, ref.shape[:2] + (2, 2))
    got = out.to_host() if output == "device" else out
    _close(got, ref)


def test_scalar_p1_p2_and_reference_fields(ctx):
    for degree in (1, 2):
        m = tri_case(degree=degree)
        xy, xq = m["dof_coords"], m["xq"]
        T = xy[:, 0] ** 2 + xy[:, 1] if degree == 2 else 0.5 * xy[:, 0] - 2.0 * xy[:, 1]  # part1.py:187
        tab = _mk(ctx, m, 1, coefficient=T)
        val = tab.evaluate("value", output="host")
        grad = tab.evaluate("grad", output="host")
        assert val.shape == (m["dofmap"].shape[0], 3) and grad.shape == (m["dofmap"].shape[0], 3, 2)
        _close(val, _ref(m, "value", T, 1))
        _close(grad, _ref(m, "grad", T, 1))
        if degree == 2:
            np.testing.assert_allclose(val, xq[..., 0] ** 2 + xq[..., 1], atol=1e-13)
            np.testing.assert_allclose(grad[..., 0], 2 * xq[..., 0], atol=1e-12)
    m = tri_case(degree=1)
    xy = m["dof_coords"]
    u = np.stack([0.1 * xy[:, 0], 0.3 * xy[:, 1]], 1).reshape(-1)  # test_operands_evaluation.py:20
    F = _mk(ctx, m, 2).evaluate("def_grad", u, output="host")
    np.testing.assert_allclose(F, np.broadcast_to(np.array([[1.1, 0.0], [0.0, 1.3]]), F.shape), atol=1e-14)


def test_entities_empty_and_errors(ctx):
    m = tri_case()
    u = syn.smooth_displacement(m["dof_coords"], seed=5).reshape(-1)
    tab = _mk(ctx, m, 2)
    cells = np.array([5, 0, 17, 5, 3, 125], dtype=np.int32)  # arbitrary order, repeats allowed (the `entities` array)
    _close(tab.evaluate("mandel_strain", u, entities=cells, output="host"), _ref(m, "mandel_strain", u, 2, cells))
    assert tab.evaluate("grad", u, entities=np.zeros(0, dtype=np.int32), output="host").shape == (0, 3, 2, 2)
    with pytest.raises(eo.EOError):
        tab.evaluate("grad", u, entities=np.array([10**6], dtype=np.int32), output="host")
    with pytest.raises(NotImplementedError):
        tab.evaluate("grad", u, entities=np.zeros((2, 2), dtype=np.int32))  # (cell, facet) pairs: reference path
    with pytest.raises(ValueError):
        tab.evaluate("grad", u[:-2], output="host")
    with pytest.raises(ValueError):
        _mk(ctx, tri_case(degree=1), 1).evaluate("mandel_strain", np.zeros(80))
    bad = dict(m)
    bad["dofmap"] = m["dofmap"].copy()
    bad["dofmap"][3, 2] = m["n_dofs"] + 7
    with pytest.raises(eo.EOError):
        _mk(ctx, bad, 2)


def test_tetrahedra(ctx):
    t = tet_case()
    u = np.random.default_rng(0).normal(size=t["n_dofs"] * 3)
    tab = _mk(ctx, t, 3)
    ref = ot.tabulate(ot.GRAD, u, t["dofmap"], 3, t["x"], t["x_dofmap"], t["phi"], t["dphi"], t["dpsi"])
    _close(tab.evaluate("grad", u, output="host"), ref, 1e-11)


def test_fused_tabulate_von_mises_equals_two_steps(ctx):
    """Strain never stored: the fused kernel must give the bits of (tabulate -> eo_vm_eval)."""
    m = tri_case(nx=61, ny=47)
    n = m["dofmap"].shape[0] * 3
    u = syn.smooth_displacement(m["dof_coords"], scale=1.5e-3, seed=7).reshape(-1)
    _, sn, p = syn.vm_batch(n, seed=9)
    tab = _mk(ctx, m, 2)
    vm_a, vm_b = eo.VonMises(ctx=ctx), eo.VonMises(ctx=ctx)
    vm_a.set_history(sn, p)
    vm_b.set_history(sn, p)
    strain = tab.evaluate("mandel_strain", u)  # DeviceArray (n_cells, 3, 4)
    Ct_a, sig_a, dp_a = vm_a((1,))(strain)
    strain_f = ctx.empty((4 * n,))
    Ct_b = tab.vm_fused(vm_b, u, strain=strain_f, exact=True)
    ctx.sync()
    assert np.array_equal(strain_f.to_host(), strain.to_host().reshape(-1))
    assert np.array_equal(Ct_b.to_host(), Ct_a)
    assert np.array_equal(vm_b.sigma_dev.to_host(), sig_a) and np.array_equal(vm_b.dp_dev.to_host(), dp_a)
    assert 0.05 < (dp_a > 0).mean() < 0.95  # both regimes exercised
    # default variant: cheaper downstream algebra, identical flags, values to a few ulp
    vm_c = eo.VonMises(ctx=ctx)
    vm_c.set_history(sn, p)
    Ct_c = tab.vm_fused(vm_c, u).to_host()
    dp_c, sig_c = vm_c.dp_dev.to_host(), vm_c.sigma_dev.to_host()
    assert np.array_equal(dp_c > 0, dp_a > 0)
    _close(Ct_c, Ct_a, 1e-12)
    _close(sig_c, sig_a, 1e-12)
    _close(dp_c, dp_a, 1e-12)
    el = np.repeat(dp_a == 0, 16)
    assert np.array_equal(Ct_c[el], Ct_a[el]) and np.array_equal(sig_c[np.repeat(dp_a == 0, 4)], sig_a[np.repeat(dp_a == 0, 4)])
    # and against the oracle chain
    e_ref = _ref(m, "mandel_strain", u, 2).reshape(-1, 4)
    rC, rs, rdp = native.vm_return_mapping(e_ref, sn, p, oc.VonMisesParams())
    assert np.array_equal(dp_a > 0, rdp > 0)
    _close(Ct_a, rC, 1e-10)  # strain differs by rounding (summation order) -> 1e-12 * cond of the plastic branch
    _close(sig_a, rs, 1e-11)


class _X:
    def __init__(self, n):
        self.array = np.zeros(n)

    def scatter_forward(self):
        pass


class _Coeff:
    def __init__(self, n):
        self.x = _X(n)
        self.dtype = np.float64


class _Op:
    """Attributes of FEMExternalOperator that the numeric layer touches (external_operator.py:375-445)."""

    def __init__(self, operands, n_out, external_function, derivatives):
        self.ufl_operands, self.ref_coefficient = tuple(operands), _Coeff(n_out)
        self.external_function, self.derivatives = external_function, derivatives
        self.unrolled_dofmap, self._is_mixed = None, False

    def _assign_func(self, values):
        self.ref_coefficient.x.array[:] = values


def test_drop_in_constitutive_update_flow(ctx):
    """demo_plasticity_von_mises.py:445-456 with the GPU tabulator and the GPU callable."""
    m = tri_case(nx=21, ny=17)
    n = m["dofmap"].shape[0] * 3
    Du = _Coeff(m["n_dofs"] * 2)
    Du.x.array[:] = syn.smooth_displacement(m["dof_coords"], scale=1.5e-3, seed=8).reshape(-1)
    _, sn, p = syn.vm_batch(n, seed=10)
    vm = eo.VonMises(ctx=ctx)
    vm.set_history(sn, p)
    eps_operand = "epsilon(Du)"  # stands for the UFL expression
    J_op = _Op([eps_operand], 16 * n, vm, (1,))
    tab = _mk(ctx, m, 2, coefficient=Du).register(eps_operand, "mandel_strain")
    evaluated_operands = eo.evaluate_operands([J_op], tabulator=tab)
    assert isinstance(evaluated_operands[eps_operand], eo.DeviceArray)
    ((_, sigma_new, dp_new),) = eo.evaluate_external_operators([J_op], evaluated_operands)
    e_ref = _ref(m, "mandel_strain", Du.x.array, 2).reshape(-1, 4)
    rC, rs, rdp = native.vm_return_mapping(e_ref, sn, p, oc.VonMisesParams())
    _close(J_op.ref_coefficient.x.array, rC, 1e-10)
    _close(sigma_new, rs, 1e-11)
    assert np.array_equal(dp_new > 0, rdp > 0)


def test_lazy_operand_is_materialised_by_callables_that_do_not_fuse(ctx):
    """`register(..., output="lazy")` hands an un-tabulated operand to the callable; a callable without a fused kernel
    (Mohr-Coulomb) tabulates it after all and gives the same result as with an explicit tabulation."""
    from dolfinx_external_operator_b200 import synthetic as syn
    from dolfinx_external_operator_b200.tabulation import LazyOperand
    from tab_util import tri_case

    m = tri_case(nx=9, ny=8)
    tab = eo.Tabulator(dofmap=m["dofmap"], x_dofmap=m["x_dofmap"], x=m["x"], phi=m["phi"], dphi=m["dphi"], bs=2,
                       n_dofs=m["n_dofs"], ctx=ctx)
    u = syn.smooth_displacement(m["dof_coords"], scale=1e-4, seed=3).reshape(-1)
    n = 3 * m["dofmap"].shape[0]
    mc = eo.MohrCoulomb(ctx=ctx)
    mc.set_history(np.tile([-1.0, -1.2, -0.9, 0.1], (n, 1)))
    a = [np.array(x) for x in mc((1,))(tab.evaluate("mandel_strain", u))]
    b = [np.array(x) for x in mc((1,))(LazyOperand(tab, 2, u))]
    for x, y in zip(a, b):
        assert np.array_equal(x, y)


# ------------------------------------------------------------------ nb = 10 instantiations of the fast path
def test_p3_triangles_fast_path(ctx):
    """tab_kernel<2,1,10> / <2,2,10> (P3 triangles, reference/test/test_nested_ex_op.py:93-105 sweeps these element
    families): kernel vs oracle, and a cubic field reproduced exactly."""
    from tab_util import tri_case_discontinuous

    m = tri_case_discontinuous(degree=3, nx=13, ny=9)
    x, y = m["dof_coords"][:, 0], m["dof_coords"][:, 1]
    xq, yq = m["xq"][..., 0], m["xq"][..., 1]
    f = lambda x, y: x**3 - 2 * x * x * y + 0.5 * y**3 + x * y - 0.3 * y  # noqa: E731
    fx = lambda x, y: 3 * x * x - 4 * x * y + y  # noqa: E731
    fy = lambda x, y: -2 * x * x + 1.5 * y * y + x - 0.3  # noqa: E731
    T = f(x, y)
    tab1 = _mk(ctx, m, 1)
    val, grad = tab1.evaluate("value", T, output="host"), tab1.evaluate("grad", T, output="host")
    _close(val, _ref(m, "value", T, 1))
    _close(grad, _ref(m, "grad", T, 1), 1e-11)
    np.testing.assert_allclose(val, f(xq, yq), atol=1e-13)
    np.testing.assert_allclose(grad, np.stack([fx(xq, yq), fy(xq, yq)], -1), atol=1e-11)
    u = np.stack([f(x, y), fy(x, y)], 1).reshape(-1)
    tab2 = _mk(ctx, m, 2)
    for kind in ("value", "grad", "mandel_strain", "def_grad"):
        _close(tab2.evaluate(kind, u, output="host"), _ref(m, kind, u, 2), 1e-11)
    cells = np.array([7, 0, 3, 3, m["dofmap"].shape[0] - 1], dtype=np.int32)
    _close(tab2.evaluate("mandel_strain", u, entities=cells, output="host"), _ref(m, "mandel_strain", u, 2, cells), 1e-11)
    # fused tabulate + von Mises on P3 (tab_vm_kernel<10, 3, *>) == two steps, bit for bit
    n = m["dofmap"].shape[0] * 3
    us = 2e-3 * u
    _, sn, p = syn.vm_batch(n, seed=4)
    vm_a, vm_b = eo.VonMises(ctx=ctx), eo.VonMises(ctx=ctx)
    vm_a.set_history(sn, p)
    vm_b.set_history(sn, p)
    Ct_a, sig_a, dp_a = vm_a((1,))(tab2.evaluate("mandel_strain", us))
    Ct_b = tab2.vm_fused(vm_b, us, exact=True)
    assert np.array_equal(Ct_b.to_host(), Ct_a) and np.array_equal(vm_b.sigma_dev.to_host(), sig_a)
    assert np.array_equal(vm_b.dp_dev.to_host(), dp_a) and 0.02 < (dp_a > 0).mean() < 0.98


def test_p2_tetrahedra_fast_path(ctx):
    """tab_kernel<3,1,10> / <3,3,10> (P2 tetrahedra): kernel vs oracle, quadratic fields reproduced exactly."""
    from tab_util import tet_case_discontinuous

    m = tet_case_discontinuous(n=3)
    X, xq = m["dof_coords"], m["xq"]
    f = lambda p: p[..., 0] ** 2 - p[..., 0] * p[..., 2] + 0.5 * p[..., 1] * p[..., 2] + p[..., 1]  # noqa: E731
    gf = lambda p: np.stack([2 * p[..., 0] - p[..., 2], 0.5 * p[..., 2] + 1.0, -p[..., 0] + 0.5 * p[..., 1]], -1)  # noqa: E731
    T = f(X)
    tab1 = _mk(ctx, m, 1)
    val, grad = tab1.evaluate("value", T, output="host"), tab1.evaluate("grad", T, output="host")
    _close(val, _ref(m, "value", T, 1))
    _close(grad, _ref(m, "grad", T, 1), 1e-11)
    np.testing.assert_allclose(val, f(xq), atol=1e-13)
    np.testing.assert_allclose(grad, gf(xq), atol=1e-11)
    u = np.stack([f(X), 0.3 * X[:, 0] * X[:, 1], X[:, 2] ** 2], 1).reshape(-1)
    tab3 = _mk(ctx, m, 3)
    for kind in ("value", "grad", "def_grad"):
        out = tab3.evaluate(kind, u, output="device")
        _close(out.to_host(), _ref(m, kind, u, 3), 1e-11)
    F = tab3.evaluate("def_grad", u, output="host")
    np.testing.assert_allclose(F[..., 0, :], gf(xq) + np.array([1.0, 0, 0]), atol=1e-11)


# ------------------------------------------------------------------ numbering-independent results
@pytest.mark.parametrize("order", ["shuffled", "rcm"])
def test_renumbered_mesh_gives_the_same_bits(ctx, order):
    """The structured generators number dofs row-major (best case for the gathers).  Under a random / RCM numbering of
    cells, dofs and nodes the kernels must return the SAME per-cell bits (in the new cell order) and agree with the oracle
    evaluated on the renumbered mesh."""
    m = tri_case(nx=33, ny=29)
    r = syn.renumber(m, order, seed=2)
    u = syn.smooth_displacement(m["dof_coords"], scale=1.5e-3, seed=3)
    ur = np.empty_like(u)
    ur[r["dof_new"]] = u
    base = _mk(ctx, m, 2).evaluate("mandel_strain", u.reshape(-1), output="host")
    tab = _mk(ctx, r, 2)
    got = tab.evaluate("mandel_strain", ur.reshape(-1), output="host")
    assert np.array_equal(got, base[r["cell_old"]])
    _close(got, _ref(r, "mandel_strain", ur.reshape(-1), 2))
    n = 3 * m["dofmap"].shape[0]
    _, sn, p = syn.vm_batch(n, seed=9)
    vm_a, vm_b = eo.VonMises(ctx=ctx), eo.VonMises(ctx=ctx)
    vm_a.set_history(sn, p)
    vm_b.set_history(sn, p)
    Ct_a, _, dp_a = vm_a((1,))(got)
    Ct_b = tab.vm_fused(vm_b, ur.reshape(-1), exact=True)
    assert np.array_equal(Ct_b.to_host(), Ct_a) and np.array_equal(vm_b.dp_dev.to_host(), dp_a)
