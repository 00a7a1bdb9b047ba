"""GPU parity tests of the run-time compiled (NVRTC) generic path: the sm_100a kernels generated from a user
model, called through the C ABI / the callable protocol, against (a) the golden vectors made by the
reference's own code, (b) the hard-wired kernels, (c) the host build of the same model text (jit_util)."""

import os

import numpy as np
import pytest

import dolfinx_external_operator_b200 as eo
from dolfinx_external_operator_b200 import jit_models as jm
from dolfinx_external_operator_b200.jit import JitModel
from jit_util import host_eval
from oracle import inputs
from test_jit_cpu import ELEMENTARY

pytestmark = pytest.mark.gpu


def _close(a, b, rtol):
    b = np.asarray(b).reshape(-1)
    np.testing.assert_allclose(np.asarray(a).reshape(-1), b, rtol=rtol, atol=rtol * np.abs(b).max())


@pytest.mark.parametrize("kind", ["mixed", "elastic", "plastic"])
def test_jit_von_mises_against_reference_golden(ctx, golden_dir, kind):
    g = np.load(os.path.join(golden_dir, "vm_seed0_n1026.npz"))
    deps, sn, p = g[f"{kind}_deps"], g[f"{kind}_sigma_n"], g[f"{kind}_p"]
    m = jm.von_mises(ctx=ctx)
    m.set_state(0, sn)
    m.set_state(1, p)
    Ct, sig, dp = m((1,))(deps.reshape(-1, 3, 4))  # the protocol call of demo_vm:343-352
    assert np.array_equal(dp > 0, g[f"{kind}_dp"] > 0)  # plastic flags bit-exact
    _close(Ct, g[f"{kind}_C_tang"], 1e-12)
    _close(sig, g[f"{kind}_sigma"], 1e-12)
    _close(dp, g[f"{kind}_dp"], 1e-12)
    m0 = jm.von_mises(ctx=ctx, returns=("out", "aux0"))
    m0.state = m.state
    sig0, dp0 = m0((0,))(deps)
    assert np.array_equal(sig0, sig) and np.array_equal(dp0, dp)  # same primal in every instantiation


@pytest.mark.parametrize("n", [1, 31, 257, 100_003, 2_500_000])
def test_jit_von_mises_against_hardwired_kernel(ctx, n):
    """Generic path vs vm_kernel on the same seeded batch (pageable host arrays -> chunked pipeline for the
    largest size); the AD tangent and the closed-form tangent agree to 1e-11."""
    deps, sn, p = inputs.vm_batch(n, seed=n)
    vm = eo.VonMises(ctx=ctx)
    vm.set_history(sn, p)
    rC, rs, rdp = (np.array(a) for a in vm((1,))(deps.reshape(-1, 1, 4)))
    m = jm.von_mises(ctx=ctx)
    m.set_state(0, sn)
    m.set_state(1, p)
    Ct, sig, dp = m((1,))(deps.reshape(-1, 1, 4))
    assert np.array_equal(dp > 0, rdp > 0)
    _close(Ct, rC, 1e-11)
    _close(sig, rs, 1e-12)
    _close(dp, rdp, 1e-12)


def test_jit_heat_against_reference_golden(ctx, golden_dir):
    g = np.load(os.path.join(golden_dir, "heat_seed0_n4098.npz"))
    T, s = g["T"], g["sigma"]
    m = jm.heat_flux(ctx=ctx)
    _close(m((0, 0))(T, s), g["q"], 1e-13)
    _close(m((1, 0))(T, s), g["dqdT"], 1e-13)
    _close(m((0, 1))(T, s), g["dqdsigma"], 1e-13)
    k = jm.heat_conductivity(ctx=ctx)
    _close(k((0,))(T), g["k"], 1e-13)
    _close(k((1,))(T), g["dk"], 1e-13)
    _close(k((2,))(T), 2.0 / (1.0 + T) ** 3, 1e-13)
    _close(m((1, 1))(T, s), host_eval(m, (1, 1), [T, s])[0], 1e-13)


def test_jit_elementary_functions_match_host_build(ctx):
    """Every function of eo_dual.h, orders 0-2, NVRTC/sm_100a vs g++ on the same text."""
    rng = np.random.default_rng(5)
    n = 4097
    x = rng.uniform(0.2, 1.2, (n, 2))
    x[:, 1] += 0.05 * np.sign(x[:, 1] - x[:, 0])
    m = JitModel(ELEMENTARY, "elem", [(2,)], (7,), params=[0.7], ctx=ctx)
    for d, rtol in [((0,), 1e-13), ((1,), 1e-12), ((2,), 1e-11)]:
        _close(m(d)(x), host_eval(m, d, [x])[0], rtol)


def test_jit_device_operands_and_async_eval(ctx):
    n = 70_001
    deps, sn, p = inputs.vm_batch(n, seed=11)
    m = jm.von_mises(ctx=ctx)
    m.set_state(0, sn)
    m.set_state(1, p)
    ref = [np.array(a) for a in m((1,))(deps)]
    d_deps = ctx.to_device(deps.reshape(-1))
    got = m((1,))(d_deps)  # DeviceArray operand: nothing uploaded
    for a, b in zip(got, ref):
        assert np.array_equal(a, b)
    d_C, d_s, d_dp = ctx.empty((16 * n,)), ctx.empty((4 * n,)), ctx.empty((n,))
    l0 = ctx.launch_count
    m.eval_device((1,), [d_deps], d_C, d_s, [d_dp])
    ctx.sync()
    assert ctx.launch_count == l0 + 1
    assert np.array_equal(d_C.to_host(), ref[0]) and np.array_equal(d_s.to_host(), ref[1])
    assert np.array_equal(d_dp.to_host(), ref[2])


def test_jit_empty_and_errors(ctx):
    m = jm.heat_conductivity(ctx=ctx)
    assert m((0,))(np.zeros((0, 3))).size == 0
    with pytest.raises(ValueError):
        jm.heat_flux(ctx=ctx)((0, 0))(np.zeros(6), np.zeros(10))  # 6 vs 5 points
    with pytest.raises(ValueError):
        jm.von_mises(ctx=ctx)((1,))(np.zeros(8))  # state not set


# ---------------------------------------------------------------------------------------------------------------
# fused: operands tabulated inside the model's kernel (eo_jit_eval_tabulated), never stored
def _p2_setup(ctx):
    from dolfinx_external_operator_b200 import synthetic as syn
    from tab_util import tri_case

    m = tri_case(nx=23, ny=17)
    tab = eo.Tabulator(dofmap=m["dofmap"], x_dofmap=m["x_dofmap"], x=m["x"], phi=m["phi"], dphi=m["dphi"], bs=2,
                       n_dofs=m["n_dofs"], ctx=ctx)
    u = syn.smooth_displacement(m["dof_coords"], scale=1.5e-3, seed=2).reshape(-1)
    return m, tab, u


def test_fused_tabulation_von_mises_equals_two_step(ctx):
    from types import SimpleNamespace

    m, tab, u = _p2_setup(ctx)
    n = 3 * m["dofmap"].shape[0]
    _, sn, p = inputs.vm_batch(n, seed=4)
    jv = jm.von_mises(ctx=ctx, fmad=False)  # like tab_kernel (-fmad=false): fused and two-step strains are the same bits
    jv.set_state(0, sn)
    jv.set_state(1, p)
    strain = tab.evaluate("mandel_strain", u)  # two-step: tabulate, then evaluate
    ref = [np.array(a) for a in jv((1,))(strain)]
    # the reference's two calls, with the operand registered as lazy -> one fused kernel
    tab.coefficient = u
    tab.register("eps", "mandel_strain", output="lazy")
    op = SimpleNamespace(ufl_operands=["eps"], derivatives=(1,), external_function=jv, b200_tabulator=tab,
                         ref_coefficient=SimpleNamespace(x=SimpleNamespace(array=np.zeros(16 * n), scatter_forward=lambda: None)))
    op._assign_func = lambda values: op.ref_coefficient.x.array.__setitem__(slice(None), values)  # :289-290
    ops = eo.evaluate_operands([op])
    assert type(ops["eps"]).__name__ == "LazyOperand"
    l0 = ctx.launch_count
    ((Ct, sig, dp),) = eo.evaluate_external_operators([op], ops)
    assert ctx.launch_count - l0 == (n + (1 << 20) - 1) // (1 << 20)  # one launch per pipeline chunk, no tabulation launch
    for a, b in zip((Ct, sig, dp), ref):
        assert np.array_equal(a, b)  # same arithmetic, same bits
    assert np.array_equal(op.ref_coefficient.x.array, ref[0])
    # chunked host pipeline: chunks that are not multiples of nq (the kernel maps global point -> (cell, q))
    ctx.set_chunk(1000)
    try:
        ((Ct3, sig3, dp3),) = eo.evaluate_external_operators([op], ops)
        for a, b in zip((Ct3, sig3, dp3), ref):
            assert np.array_equal(a, b)
    finally:
        ctx.set_chunk(1 << 20)
    # the hard-wired von Mises callable fuses the same way (eo_tab_vm_fused, exact variant)
    vm = eo.VonMises(ctx=ctx)
    vm.set_history(sn, p)
    two = [np.array(a) for a in vm((1,))(strain)]
    op.external_function = vm
    ((Ct2, sig2, dp2),) = eo.evaluate_external_operators([op], ops)
    for a, b in zip((Ct2, sig2, dp2), two):
        assert np.array_equal(a, b)


def test_fused_tabulation_two_operands_heat(ctx):
    """q(T, grad T): two operands of different kinds tabulated from the same P2 scalar coefficient."""
    from dolfinx_external_operator_b200 import elements as el
    from tab_util import tri_case

    m = tri_case(nx=11, ny=13)
    tab = eo.Tabulator(dofmap=m["dofmap"], x_dofmap=m["x_dofmap"], x=m["x"], phi=m["phi"], dphi=m["dphi"], bs=1,
                       n_dofs=m["n_dofs"], ctx=ctx)
    x, y = m["dof_coords"][:, 0], m["dof_coords"][:, 1]
    T = x * x + y  # part1.py:187
    h = jm.heat_flux(ctx=ctx, fmad=False)
    from dolfinx_external_operator_b200.tabulation import LazyOperand

    lz = [LazyOperand(tab, 0, T), LazyOperand(tab, 1, T)]
    Tq, gq = tab.evaluate("value", T, output="host"), tab.evaluate("grad", T, output="host")
    for d in [(0, 0), (1, 0), (0, 1), (1, 1)]:
        fused = np.array(h(d)(*lz))
        two = np.array(h(d)(Tq, gq))
        assert np.array_equal(fused, two)
    # analytic: q = -grad T / (1 + T), grad T = (2x, 1) (P2 reproduces x^2 + y exactly)
    xq = m["xq"].reshape(-1, 2)
    k = 1.0 / (1.0 + xq[:, 0] ** 2 + xq[:, 1])
    np.testing.assert_allclose(np.array(h((0, 0))(*lz)).reshape(-1, 2), -k[:, None] * np.stack([2 * xq[:, 0], np.ones(len(xq))], 1),
                               rtol=1e-11, atol=1e-12)


def test_jit_von_mises_3d_extension(ctx):
    """EXTENSION (6-component Mandel): NVRTC kernel vs the NumPy restatement oracle/constitutive.py::vm3d_return_mapping."""
    from oracle import constitutive as oc
    from test_jit_cpu import _vm3d_batch

    n = 100_003
    deps, sn, p = _vm3d_batch(n, 5)
    m = jm.von_mises_3d(ctx=ctx)
    m.set_state(0, sn)
    m.set_state(1, p)
    Ct, sig, dp = m((1,))(deps)
    rC, rs, rdp = oc.vm3d_return_mapping(deps, sn, p)
    assert np.array_equal(dp > 0, rdp > 0)
    _close(Ct, rC, 1e-11)
    _close(sig, rs, 1e-12)
    _close(dp, rdp, 1e-12)


def test_jit_staged_variant_odd_component_counts(ctx):
    """Arrays with 9 and 81 doubles per point go through the staged kernel (TMA bulk copies + mbarrier): full tiles
    staged, the tail direct, host arrays through the chunk pipeline - all against the closed forms."""
    from test_jit_cpu import neo_hookean_batch, neo_hookean_reference

    n = 70_003  # not a multiple of any tile size
    F = neo_hookean_batch(n, 7)
    W, P, A = neo_hookean_reference(F)
    m = jm.neo_hookean_3d(ctx=ctx)
    l0 = ctx.launch_count
    _close(m((1,))(F.reshape(-1, 1, 3, 3)), P, 1e-12)
    assert ctx.launch_count - l0 == 2  # staged tiles + direct tail
    _close(m((2,))(F), A, 1e-11)
    _close(m((0,))(F), W, 1e-13)
    ctx.set_chunk(10_000)  # chunks whose sizes are no multiples of the tile
    try:
        _close(m((1,))(F), P, 1e-12)
    finally:
        ctx.set_chunk(1 << 20)
    d_F = ctx.to_device(F.reshape(-1))
    d_P = ctx.empty((9 * n,))
    m.eval_device((1,), [d_F], d_P)
    ctx.sync()
    _close(d_P.to_host(), P, 1e-12)
    few = m((1,))(F[:17])  # fewer points than a tile: direct kernel only
    _close(few, P[:17], 1e-12)


@pytest.mark.parametrize("n", [1023, 1024, 100_003])
def test_jit_points_per_thread_variant(ctx, n):
    """Scalar-sized models run 2 or 4 points per thread above 1024 points (bulk) + a direct tail; every variant
    must give the same bits as the one-point-per-thread kernel (same model code, same IEEE operations)."""
    import os

    rng = np.random.default_rng(n)
    T, s = rng.uniform(0.0, 2.0, n), rng.normal(size=(n, 2))
    q, k = jm.heat_flux(ctx=ctx), jm.heat_conductivity(ctx=ctx)
    l0 = ctx.launch_count
    got = {d: np.array(q(d)(T, s)) for d in [(0, 0), (1, 0), (0, 1), (1, 1)]}
    gk = {d: np.array(k(d)(T)) for d in [(0,), (1,), (2,)]}
    launches = ctx.launch_count - l0
    assert launches == (7 if n < 1024 else 7 + sum(1 for ppt in (2, 2, 2, 2, 4, 4, 4) if n % ppt))
    os.environ["EO_JIT_PPT"] = "0"
    try:
        q1, k1 = jm.heat_flux(ctx=ctx), jm.heat_conductivity(ctx=ctx)
        for d in got:
            assert np.array_equal(got[d], np.array(q1(d)(T, s))), d
        for d in gk:
            assert np.array_equal(gk[d], np.array(k1(d)(T))), d
    finally:
        del os.environ["EO_JIT_PPT"]
    _close(gk[(1,)], -1.0 / (1.0 + T) ** 2, 1e-13)


def test_fused_tabulation_two_coefficients_on_one_tabulator(ctx):
    """Regression: two lazy operands of ONE tabulator with DIFFERENT host coefficients (u and u_old on one space) each
    get their own staging buffer - q(T_a, grad T_b) must not be tabulated from a single coefficient."""
    from dolfinx_external_operator_b200.tabulation import LazyOperand
    from tab_util import tri_case

    m = tri_case(nx=9, ny=7)
    tab = eo.Tabulator(dofmap=m["dofmap"], x_dofmap=m["x_dofmap"], x=m["x"], phi=m["phi"], dphi=m["dphi"], bs=1,
                       n_dofs=m["n_dofs"], ctx=ctx)
    x, y = m["dof_coords"][:, 0], m["dof_coords"][:, 1]
    Ta, Tb = x * x + y, 0.5 * x * y + 0.25 * y * y
    h = jm.heat_flux(ctx=ctx, fmad=False)
    fused = np.array(h((0, 0))(LazyOperand(tab, 0, Ta), LazyOperand(tab, 1, Tb)))
    two = np.array(h((0, 0))(tab.evaluate("value", Ta, output="host"), tab.evaluate("grad", Tb, output="host")))
    assert np.array_equal(fused, two)
    same = np.array(h((0, 0))(tab.evaluate("value", Tb, output="host"), tab.evaluate("grad", Tb, output="host")))
    assert not np.array_equal(fused, same)
