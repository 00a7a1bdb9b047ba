"""GPU parity tests for the Mohr-Coulomb kernel (run with -m gpu on a B200), through the C ABI:
against the goldens produced by the reference's own source, and against the oracle on seeded batches.
Tolerances (north_star): plastic/elastic flags and iteration counts bit-exact; local-Newton model
stress / tangent / state within rtol 1e-10."""

import ctypes as C
import dataclasses
import os

import numpy as np
import pytest

import dolfinx_external_operator_b200 as eo
from dolfinx_external_operator_b200._lib import McParams
from oracle import constitutive as oc
from oracle import inputs, native
from mc_util import check_mc, check_mc_exact

pytestmark = pytest.mark.gpu
RTOL = 1e-10
PRM = oc.MohrCoulombParams()


def _close(a, b, rtol=RTOL):
    np.testing.assert_allclose(np.asarray(a).reshape(-1), np.asarray(b).reshape(-1), rtol=rtol,
                               atol=rtol * np.abs(b).max())


def _abi(ctx, deps, sn, prm=PRM, scheme=0, aux=True):
    n = deps.shape[0]
    deps, sn = np.ascontiguousarray(deps), np.ascontiguousarray(sn)
    out = {"C_tang": np.full((n, 4, 4), np.nan), "sigma": np.full((n, 4), np.nan),
           "niter": np.full(n, -7, dtype=np.int32), "yielding": np.full(n, np.nan), "norm_res": np.full(n, np.nan),
           "dlambda": np.full(n, np.nan)}
    q = McParams(prm.E, prm.nu, prm.c, prm.phi, prm.psi, prm.theta_T, prm.a, prm.tol, prm.Nitermax)
    v = lambda a: a.ctypes.data  # noqa: E731
    auxp = [v(out[k]) for k in ("niter", "yielding", "norm_res", "dlambda")] if aux else [None] * 4
    ctx.check(ctx.lib.eo_mc_eval_scheme(ctx.handle, C.byref(q), v(deps), v(sn), v(out["C_tang"]), v(out["sigma"]),
                                        *auxp, n, scheme))
    return out


def _check(o, g, deps, sigma_n, prm=PRM):
    return check_mc(o, g, deps, sigma_n, prm)


def _stepper(prm=PRM):
    return lambda d, s: native.mc_stress(d, s, prm, parallel=True)[0]


@pytest.mark.parametrize("scheme", [0, 1, 2, 3, 4])
@pytest.mark.parametrize("name", ["mc_path_10x9.npz", "mc_rand_seed0_n96.npz", "mc_rand_seed1_n2048.npz"])
def test_mc_against_reference_golden(ctx, golden_dir, name, scheme):
    g = np.load(os.path.join(golden_dir, name))
    o = _abi(ctx, g["deps"], g["sigma_n"], scheme=scheme)
    _check(o, g, g["deps"], g["sigma_n"])
    el = g["yielding"] <= 0
    Cel = oc.elastic_stiffness(PRM.lmbda, PRM.mu)
    assert np.array_equal(o["C_tang"][el], np.broadcast_to(Cel, (el.sum(), 4, 4)))  # demo_mc:442-443


@pytest.mark.parametrize("n", [1, 31, 127, 128, 129, 1023, 1024, 1025, 5000, 60_001])
def test_mc_against_oracle_ragged_sizes(ctx, n):
    d, s = inputs.mc_batch(n, seed=n, stepper=_stepper())
    o = _abi(ctx, d, s)
    _check(o, native.mc_return_mapping(d, s, PRM, parallel=True), d, s)


def test_mc_against_the_exact_value_flat_tolerance(ctx):
    """The reference program evaluated in extended precision (oracle_mc_return_mapping_ld) is the exact value: the kernel
    is within the FLAT north_star tolerance 1e-10 of it at every one of 2 x 10^5 points of the demo's stress-path family,
    hexagon corners included (there the reference's own float64 evaluation is off by up to 1.3e-10)."""
    d, s = inputs.mc_batch(200_000, seed=7, stepper=_stepper())
    o = _abi(ctx, d, s)
    worst = check_mc_exact(o, d, s, PRM)
    assert max(worst.values()) < 5e-11, worst
    g = native.mc_return_mapping(d, s, PRM, parallel=True)
    relaxed = _check(o, g, d, s)
    assert relaxed < 0.01 * d.shape[0]


def test_mc_schemes_agree(ctx):
    """The stage-queue kernel and the one-thread-per-point kernel run the same per-point code (mc_core.cuh);
    nvcc may contract FMAs differently in the two contexts, so equality is to rounding, not bitwise."""
    d, s = inputs.mc_batch(20_000, seed=11, stepper=_stepper())
    a, b = _abi(ctx, d, s, scheme=0), _abi(ctx, d, s, scheme=1)
    assert np.array_equal(a["niter"], b["niter"]) and np.array_equal(a["yielding"] > 0, b["yielding"] > 0)
    _check(a, b, d, s)
    # ring scheduler in two passes (4), in one pass without (2) / with (3) sub-partition affinity: the same kernel
    # template, only scheduled differently -> bit-identical stresses, tangents and aux outputs; the lane-class kernel
    # (0, default) inlines the same stage code in another kernel: equal to rounding; the statistics records agree
    st = {}
    ring = _abi(ctx, d, s, scheme=4)
    _check(a, ring, d, s)
    assert np.array_equal(a["niter"], ring["niter"])
    for scheme in (0, 2, 3, 4):
        ctx.stats_reset()
        o = _abi(ctx, d, s, scheme=scheme)
        ctx.sync()
        st[scheme] = ctx.stats()
        ref = a if scheme == 0 else ring
        for key in ("C_tang", "sigma", "niter", "dlambda", "norm_res"):
            assert np.array_equal(o[key], ref[key], equal_nan=True), (scheme, key)
        np.testing.assert_allclose(o["yielding"], a["yielding"], rtol=1e-13, atol=1e-13)
    for scheme in (2, 3, 4):
        for key in ("n_points", "n_plastic", "n_nonconverged", "n_nonfinite", "niter_max", "f_max"):
            assert st[scheme][key] == st[0][key], (scheme, key)
        assert abs(st[scheme]["res_max"] - st[0]["res_max"]) <= 1e-10 * max(st[0]["res_max"], 1e-300)
        assert np.array_equal(st[scheme]["niter_hist"], st[0]["niter_hist"])


def test_mc_demo_tracing_path(ctx):
    """The demo's yield-surface tracing driver (demo_mc:853-930) walked with the GPU kernel as the stress
    update: same iteration histogram as the reference and same stresses as the oracle."""
    mc = eo.MohrCoulomb(ctx=ctx, history=None)
    d, s = inputs.mc_demo_path(50, 9, stepper=mc.stress_update)
    d2, s2 = inputs.mc_demo_path(50, 9, stepper=_stepper())
    _close(s, s2, 1e-10)
    o = _abi(ctx, d, s)
    it, cnt = np.unique(o["niter"], return_counts=True)
    assert dict(zip(it.tolist(), cnt.tolist())) == {1: 276, 2: 99, 3: 57, 4: 15, 5: 3}
    _check(o, native.mc_return_mapping(d, s, PRM, parallel=True), d, s)
    check_mc_exact(o, d, s, PRM)  # the path ends in the hexagon corners: flat 1e-10 against the exact value there too
    # returned plastic stresses lie on the yield surface
    pl = o["yielding"] > 0
    assert np.abs(native.mc_yield(o["sigma"][pl], PRM)).max() < 1e-6


def test_mc_non_associative(ctx):
    prm = dataclasses.replace(PRM, psi=10 * np.pi / 180)
    d, s = inputs.mc_batch(8000, seed=5, stepper=_stepper(prm))
    o = _abi(ctx, d, s, prm)
    _check(o, native.mc_return_mapping(d, s, prm, parallel=True), d, s, prm)


def test_mc_edge_semantics(ctx):
    """IEEE behaviour kept from the reference: zero increment -> ||res0|| = 0 -> zero
    iterations and zero tangent; plastic step from a hydrostatic state -> J2 = 0 -> NaN residual -> zero
    iterations; sign(theta = 0) = +1."""
    d = np.array([[0.0, 0, 0, 0], [0.0, 0, 0, 0], [1e-3, -1e-3, 0, 0]])
    s = np.array([[-1.0, -1.2, -0.9, 0.1], [-1.0, -1, -1, 0], [-1.0, -1, -1, 0]])
    o, r = _abi(ctx, d, s), native.mc_return_mapping(d, s, PRM)
    assert np.array_equal(o["niter"], r["niter"]) and list(o["niter"]) == [0, 0, 0]
    assert np.array_equal(o["sigma"], r["sigma"])
    assert np.array_equal(o["C_tang"][:2], np.zeros((2, 4, 4)))
    assert np.array_equal(np.isnan(o["C_tang"]), np.isnan(r["C_tang"]))
    assert np.array_equal(o["yielding"] > 0, r["yielding"] > 0)


def test_mc_statistics_and_api(ctx):
    """`MohrCoulomb((1,))` is the drop-in for C_tang_impl (demo_mc:577-593): returns (C_tang, sigma) flat,
    history resident, summary = the reference's per-call prints, commit = demo_mc:728."""
    n = 30_000
    d, s = inputs.mc_batch(n, seed=2, stepper=_stepper())
    mc = eo.MohrCoulomb(ctx=ctx)
    mc.set_history(s)
    with pytest.raises(NotImplementedError):
        mc((0,))
    Ct, sig = mc((1,))(d.reshape(-1, 3, 4))
    r = native.mc_return_mapping(d, s, PRM, parallel=True)
    assert Ct.shape == (16 * n,) and sig.shape == (4 * n,)
    _check({"C_tang": Ct, "sigma": sig, "niter": mc.niter, "yielding": mc.yielding, "norm_res": mc.norm_res,
            "dlambda": mc.dlambda}, r, d, s)
    sm = mc.summary()
    it, cnt = np.unique(r["niter"], return_counts=True)
    assert np.array_equal(sm["unique_iters"], it) and np.array_equal(sm["counts"], cnt)
    assert sm["max_f"] == r["yielding"].max()
    np.testing.assert_allclose(sm["max_residual"], r["norm_res"].max(), rtol=0, atol=1e-12)
    assert sm["n_plastic"] == int((r["yielding"] > 0).sum()) and sm["n_nonconverged"] == 0 and sm["n_nonfinite"] == 0
    mc.commit()
    _close(mc.get_history(), r["sigma"], 1e-9)
    # aux pointers are optional
    o = _abi(ctx, d[:1000], s[:1000], aux=False)
    _close(o["sigma"], r["sigma"][:1000], 1e-9)
    assert np.all(o["niter"] == -7)


def test_mc_device_resident_and_bad_arguments(ctx):
    n = 4096
    d, s = inputs.mc_batch(n, seed=4, stepper=_stepper())
    dd, ds = ctx.to_device(d.reshape(-1)), ctx.to_device(s.reshape(-1))
    dC, dsig = ctx.empty((16 * n,)), ctx.empty((4 * n,))
    q = McParams(PRM.E, PRM.nu, PRM.c, PRM.phi, PRM.psi, PRM.theta_T, PRM.a, PRM.tol, PRM.Nitermax)
    ctx.check(ctx.lib.eo_mc_eval(ctx.handle, C.byref(q), dd.ptr, ds.ptr, dC.ptr, dsig.ptr, None, None, None, None, n))
    r = native.mc_return_mapping(d, s, PRM, parallel=True)
    _close(dC.to_host(), r["C_tang"], 1e-9)
    _close(dsig.to_host(), r["sigma"], 1e-9)
    assert ctx.lib.eo_mc_eval(ctx.handle, C.byref(q), dd.ptr, ds.ptr, dC.ptr, dsig.ptr, None, None, None, None, 0) == 0
    assert ctx.lib.eo_mc_eval(ctx.handle, C.byref(q), None, ds.ptr, dC.ptr, dsig.ptr, None, None, None, None, n) == -1
    assert ctx.lib.eo_mc_eval(ctx.handle, C.byref(q), dd.ptr + 8, ds.ptr, dC.ptr, dsig.ptr, None, None, None, None, n - 1) == -1
    bad = McParams(PRM.E, PRM.nu, PRM.c, PRM.phi, PRM.psi, PRM.theta_T, PRM.a, PRM.tol, 500)
    assert ctx.lib.eo_mc_eval(ctx.handle, C.byref(bad), dd.ptr, ds.ptr, dC.ptr, dsig.ptr, None, None, None, None, n) == -1
    for a in (dd, ds, dC, dsig):
        a.free()


def test_mc_scratch_survives_l2_flush(ctx):
    """Regression: eo_flush_l2 growing its buffer must not touch the Mohr-Coulomb scratch (plastic-point list)."""
    d, s = inputs.mc_batch(20_000, seed=5, stepper=_stepper())
    ref = native.mc_return_mapping(d, s, PRM, parallel=True)
    _check(_abi(ctx, d, s), ref, d, s)  # allocates the scratch of the two-pass scheme
    ctx.flush_l2(300 << 20)             # grows the flush buffer
    ctx.flush_l2(400 << 20)
    tmp = [ctx.empty((1 << 20,)) for _ in range(4)]  # would reuse a freed scratch block
    for t in tmp:
        ctx.check(ctx.lib.eo_dev_memset(ctx.handle, t.ptr, 0xFF, t.nbytes))
    _check(_abi(ctx, d[:15_000], s[:15_000]), {k: v[:15_000] for k, v in ref.items()}, d[:15_000], s[:15_000])
    ctx.sync()


@pytest.mark.parametrize("degree", [1, 2])
def test_mc_fused_with_tabulation(ctx, degree):
    """eo_mc_eval_tabulated: the Mandel strain is tabulated inside pass 1 and never stored.  Same iteration counts and
    flags as tabulate -> eo_mc_eval, values to the rounding of the strain (mc.cu contracts FMAs, tab.cu does not), and
    within the tolerance of the oracle chain."""
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from dolfinx_external_operator_b200 import synthetic as syn
    from mc_util import lode_w2
    from oracle import tabulation as ot
    from tab_util import tri_case

    m = tri_case(nx=41, ny=37, degree=degree)
    n = 3 * m["dofmap"].shape[0]
    _, sn = inputs.mc_batch(n, seed=5, stepper=_stepper())
    u = syn.smooth_displacement(m["dof_coords"], scale=2e-6, seed=3).reshape(-1)
    tab = eo.Tabulator(dofmap=m["dofmap"], x_dofmap=m["x_dofmap"], x=m["x"], phi=m["phi"], dphi=m["dphi"], bs=2,
                       n_dofs=m["n_dofs"], ctx=ctx)
    mc_a, mc_b = eo.MohrCoulomb(ctx=ctx, n_qp=n), eo.MohrCoulomb(ctx=ctx, n_qp=n)
    mc_a.set_history(sn)
    mc_b.set_history(sn)
    aux = {"niter": ctx.empty((n,), np.int32), "yielding": ctx.empty((n,)), "norm_res": ctx.empty((n,)), "dlambda": ctx.empty((n,))}
    ctx.stats_reset()
    Ct_f = tab.mc_fused(mc_a, u, aux=aux).to_host()
    st = ctx.stats()
    strain = tab.evaluate("mandel_strain", u, output="host").reshape(-1, 4)
    two = _abi(ctx, strain, sn)
    fused = {"C_tang": Ct_f.reshape(n, 4, 4), "sigma": mc_a.sigma_dev.to_host().reshape(n, 4), "niter": aux["niter"].to_host(),
             "yielding": aux["yielding"].to_host(), "norm_res": aux["norm_res"].to_host(), "dlambda": aux["dlambda"].to_host()}
    assert 0.05 < (two["yielding"] > 0).mean() < 0.95
    _check(fused, two, strain, sn)
    assert st["n_points"] == n and st["n_plastic"] == int((two["yielding"] > 0).sum())
    ref = native.mc_return_mapping(ot.tabulate(ot.MANDEL_STRAIN, u, m["dofmap"], 2, m["x"], m["x_dofmap"], m["phi"], m["dphi"],
                                               m["dpsi"]).reshape(-1, 4), sn, PRM, parallel=True)
    _check(fused, ref, strain, sn)
