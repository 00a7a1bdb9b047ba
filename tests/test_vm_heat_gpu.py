"""GPU parity tests proper (run with -m gpu on a B200): the CUDA path, called through the C ABI,
against the oracle on seeded inputs and against the golden vectors made by the reference's own code.
Tolerances (north_star): plastic/elastic flags bit-exact; closed-form models rtol 1e-12."""

import ctypes as C
import os

import numpy as np
import pytest

import dolfinx_external_operator_b200 as eo
from dolfinx_external_operator_b200._lib import VmParams
from oracle import constitutive as oc
from oracle import inputs, native

pytestmark = pytest.mark.gpu
RTOL = 1e-12
PRM = oc.VonMisesParams()


def _close(a, b, rtol=RTOL):
    np.testing.assert_allclose(np.asarray(a).reshape(-1), np.asarray(b).reshape(-1), rtol=rtol,
                               atol=rtol * np.abs(b).max())


def _vm_abi(ctx, deps, sn, p, n):
    """Straight C-ABI call with pageable host arrays."""
    Ct, sig, dp = np.empty(16 * n), np.empty(4 * n), np.empty(n)
    prm = VmParams(PRM.lmbda, PRM.mu, PRM.H, PRM.sigma_0)
    v = lambda a: a.ctypes.data  # noqa: E731
    ctx.check(ctx.lib.eo_vm_eval(ctx.handle, C.byref(prm), v(deps), v(sn), v(p), v(Ct), v(sig), v(dp), n))
    return Ct, sig, dp


@pytest.mark.parametrize("kind", ["mixed", "elastic", "plastic"])
def test_vm_against_reference_golden(ctx, golden_dir, kind):
    g = np.load(os.path.join(golden_dir, "vm_seed0_n1026.npz"))
    deps, sn, p = (np.ascontiguousarray(g[f"{kind}_{k}"]) for k in ("deps", "sigma_n", "p"))
    Ct, sig, dp = _vm_abi(ctx, deps, sn, p, p.size)
    assert np.array_equal(dp > 0, g[f"{kind}_dp"] > 0)
    _close(Ct, g[f"{kind}_C_tang"])
    _close(sig, g[f"{kind}_sigma"])
    _close(dp, g[f"{kind}_dp"])
    if kind == "elastic":
        assert np.all(dp == 0.0)
        assert np.array_equal(Ct.reshape(-1, 4, 4),
                              np.broadcast_to(oc.elastic_stiffness(PRM.lmbda, PRM.mu), (p.size, 4, 4)))


@pytest.mark.parametrize("n", [1, 2, 31, 32, 33, 255, 257, 1000, 100_003])
def test_vm_against_oracle_ragged_sizes(ctx, n):
    deps, sn, p = inputs.vm_batch(n, seed=n)
    Ct, sig, dp = _vm_abi(ctx, deps, sn, p, n)
    rC, rs, rdp = native.vm_return_mapping(deps, sn, p, PRM)
    assert np.array_equal(dp > 0, rdp > 0)
    # same statement order, no contraction on either side -> identical bits
    assert np.array_equal(Ct, rC.reshape(-1)) and np.array_equal(sig, rs.reshape(-1)) and np.array_equal(dp, rdp)


def test_vm_empty_and_bad_arguments(ctx):
    prm = VmParams(PRM.lmbda, PRM.mu, PRM.H, PRM.sigma_0)
    assert ctx.lib.eo_vm_eval(ctx.handle, C.byref(prm), None, None, None, None, None, None, 0) == 0
    assert ctx.lib.eo_vm_eval(ctx.handle, C.byref(prm), None, None, None, None, None, None, 5) == -1
    assert b"NULL" in ctx.lib.eo_last_error(ctx.handle)
    assert ctx.lib.eo_vm_eval(ctx.handle, None, None, None, None, None, None, None, 5) == -1
    assert ctx.lib.eo_vm_eval(ctx.handle, C.byref(prm), None, None, None, None, None, None, -1) == -1


def test_vm_unaligned_pointers_take_the_scalar_path(ctx):
    n = 1000
    deps, sn, p = inputs.vm_batch(n, seed=5)
    d_deps = ctx.empty((4 * n + 1,))
    d_sn = ctx.empty((4 * n + 1,))
    d_p, d_Ct, d_sig, d_dp = ctx.to_device(p), ctx.empty((16 * n + 1,)), ctx.empty((4 * n + 1,)), ctx.empty((n,))
    ctx.copy(d_deps.ptr + 8, deps, deps.nbytes)
    ctx.copy(d_sn.ptr + 8, sn, sn.nbytes)
    prm = VmParams(PRM.lmbda, PRM.mu, PRM.H, PRM.sigma_0)
    ctx.check(ctx.lib.eo_vm_eval(ctx.handle, C.byref(prm), d_deps.ptr + 8, d_sn.ptr + 8, d_p.ptr, d_Ct.ptr + 8,
                                 d_sig.ptr + 8, d_dp.ptr, n))
    Ct, sig = np.empty(16 * n), np.empty(4 * n)
    ctx.copy(Ct, d_Ct.ptr + 8, Ct.nbytes)
    ctx.copy(sig, d_sig.ptr + 8, sig.nbytes)
    rC, rs, rdp = native.vm_return_mapping(deps, sn, p, PRM)
    assert np.array_equal(Ct, rC.reshape(-1)) and np.array_equal(sig, rs.reshape(-1))
    assert np.array_equal(d_dp.to_host(), rdp)


def test_vm_host_pipeline_many_chunks_and_mixed_sides(ctx):
    n = 200_001
    ctx.set_chunk(4096)  # ~49 chunks through 3 staging slots
    try:
        deps, sn, p = inputs.vm_batch(n, seed=11)
        rC, rs, rdp = native.vm_return_mapping(deps, sn, p, PRM)
        Ct, sig, dp = _vm_abi(ctx, deps, sn, p, n)
        assert np.array_equal(Ct, rC.reshape(-1)) and np.array_equal(sig, rs.reshape(-1)) and np.array_equal(dp, rdp)
        # history on the device, strain and results on the host (pinned)
        d_sn, d_p = ctx.to_device(sn.reshape(-1)), ctx.to_device(p)
        Ct2, sig2, dp2 = ctx.pinned_empty(16 * n), ctx.pinned_empty(4 * n), ctx.pinned_empty(n)
        prm = VmParams(PRM.lmbda, PRM.mu, PRM.H, PRM.sigma_0)
        ctx.check(ctx.lib.eo_vm_eval(ctx.handle, C.byref(prm), deps.ctypes.data, d_sn.ptr, d_p.ptr, Ct2.ctypes.data,
                                     sig2.ctypes.data, dp2.ctypes.data, n))
        assert np.array_equal(Ct2, rC.reshape(-1)) and np.array_equal(sig2, rs.reshape(-1)) and np.array_equal(dp2, rdp)
    finally:
        ctx.set_chunk(1 << 20)


@pytest.mark.parametrize("n", [3 * 4001, 4 * 3001])  # n % 4 == 0: the SoA layout takes the 4-points-per-thread kernel
@pytest.mark.parametrize("layout", ["aos", "soa"])
def test_vm_model_callable_resident_history_and_commit(ctx, layout, n):
    deps, sn, p = inputs.vm_batch(n, seed=2)
    vm = eo.VonMises(ctx=ctx, state_layout=layout)
    vm.set_history(sn, p)
    with pytest.raises(NotImplementedError):
        vm((0,))
    ctx.stats_reset()
    Ct, sig, dp = vm((1,))(deps.reshape(-1, 1, 4))
    rC, rs, rdp = native.vm_return_mapping(deps, sn, p, PRM)
    assert Ct.shape == (16 * n,) and sig.shape == (4 * n,) and dp.shape == (n,)
    assert np.array_equal(Ct, rC.reshape(-1)) and np.array_equal(sig, rs.reshape(-1)) and np.array_equal(dp, rdp)
    st = ctx.stats()
    assert st["n_points"] == n and st["n_plastic"] == int((rdp > 0).sum())
    # load-step commit on device == demo_vm:564-565 on the host
    vm.commit()
    sn2, p2 = vm.get_history()
    assert np.array_equal(sn2, rs.reshape(-1)) and np.array_equal(p2, p + 1.0 * rdp)
    # second increment from the committed state
    Ct_b, sig_b, dp_b = vm((1,))(deps.reshape(-1, 1, 4))
    rC2, rs2, rdp2 = native.vm_return_mapping(deps, rs, p + rdp, PRM)
    assert np.array_equal(sig_b, rs2.reshape(-1)) and np.array_equal(dp_b, rdp2) and np.array_equal(Ct_b, rC2.reshape(-1))


def test_vm_host_history_closure_style_and_bound_outputs(ctx):
    class F:  # a fem.Function stand-in: `.x.array`
        def __init__(self, a):
            self.x = type("X", (), {})()
            self.x.array = a

    n = 999
    deps, sn, p = inputs.vm_batch(n, seed=4)
    sigma_n_f, p_f = F(sn.reshape(-1).copy()), F(p.copy())
    vm = eo.VonMises(ctx=ctx, history=(sigma_n_f, p_f))
    coeff = np.zeros(16 * n)
    vm.bind_outputs(C_tang=coeff)
    Ct, sig, dp = vm((1,))(deps.reshape(-1, 3, 4))
    assert Ct is coeff
    rC, rs, rdp = native.vm_return_mapping(deps, sn, p, PRM)
    assert np.array_equal(coeff, rC.reshape(-1)) and np.array_equal(sig, rs.reshape(-1))
    # the caller updates host history (demo_vm:564-565); the next call must see it
    p_f.x.array[:] += dp
    sigma_n_f.x.array[:] = sig
    _, sig2, _ = vm((1,))(deps.reshape(-1, 3, 4))
    _, rs2, _ = native.vm_return_mapping(deps, rs, p + rdp, PRM)
    assert np.array_equal(sig2, rs2.reshape(-1))
    ctx.unregister(coeff)


def test_vm_full_size_properties(ctx):
    """BASELINE size (1e8 QPs on one GPU, 24 GB): size-independent properties.
    Tile invariance (the batch is a 2^20-point tile repeated; every repetition must reproduce the
    oracle-checked first tile bit for bit), and the closed-form elastic identity."""
    n, tile = 100_000_000, 1 << 20
    deps_t, sn_t, p_t = inputs.vm_batch(tile, seed=9)
    rC, rs, rdp = native.vm_return_mapping(deps_t, sn_t, p_t, PRM, parallel=True)
    vm = eo.VonMises(ctx=ctx, n_qp=n)
    d_deps, d_Ct = ctx.empty((4 * n,)), ctx.empty((16 * n,))
    dt_deps, dt_sn, dt_p = ctx.to_device(deps_t.reshape(-1)), ctx.to_device(sn_t.reshape(-1)), ctx.to_device(p_t)
    for off in range(0, n, tile):
        m = min(tile, n - off)
        ctx.copy(d_deps.ptr + 32 * off, dt_deps, 32 * m)
        ctx.copy(vm.sigma_n_dev.ptr + 32 * off, dt_sn, 32 * m)
        ctx.copy(vm.p_dev.ptr + 8 * off, dt_p, 8 * m)
    ctx.stats_reset()
    vm.eval_device(d_deps, d_Ct)
    ctx.sync()
    st = ctx.stats()
    full, rem = divmod(n, tile)
    assert st["n_points"] == n
    assert st["n_plastic"] == full * int((rdp > 0).sum()) + int((rdp[:rem] > 0).sum())
    buf_C, buf_s, buf_d = np.empty(16 * tile), np.empty(4 * tile), np.empty(tile)
    for off in (0, 17 * tile, (full - 1) * tile, full * tile):
        m = min(tile, n - off)
        if m <= 0:
            continue
        ctx.copy(buf_C, d_Ct.ptr + 128 * off, 128 * m)
        ctx.copy(buf_s, vm.sigma_dev.ptr + 32 * off, 32 * m)
        ctx.copy(buf_d, vm.dp_dev.ptr + 8 * off, 8 * m)
        assert np.array_equal(buf_C[:16 * m], rC.reshape(-1)[:16 * m])
        assert np.array_equal(buf_s[:4 * m], rs.reshape(-1)[:4 * m])
        assert np.array_equal(buf_d[:m], rdp[:m])
    for a in (d_deps, d_Ct, dt_deps, dt_sn, dt_p, vm.sigma_n_dev, vm.p_dev, vm.sigma_dev, vm.dp_dev):
        a.free()


# ------------------------------------------------------------------------------------ heat
@pytest.mark.parametrize("which", ["k", "dk", "q", "dqdT", "dqdsigma"])
def test_heat_against_reference_golden(ctx, golden_dir, which):
    g = np.load(os.path.join(golden_dir, "heat_seed0_n4098.npz"))
    T2, s2 = g["T"].reshape(-1, 3), g["sigma"].reshape(-1, 6)
    if which in ("k", "dk"):
        out = eo.HeatConductivity(ctx=ctx)({"k": (0,), "dk": (1,)}[which])(T2)
    else:
        out = eo.HeatFlux(ctx=ctx, fused=False)({"q": (0, 0), "dqdT": (1, 0), "dqdsigma": (0, 1)}[which])(T2, s2)
    assert np.array_equal(out, g[which])  # bit-exact


@pytest.mark.parametrize("n", [1, 2, 3, 511, 512, 513, 65_537])
def test_heat_fused_ragged(ctx, n):
    T, s = inputs.heat_batch(n, seed=n)
    hf = eo.HeatFlux(ctx=ctx, fused=True)
    launches = ctx.launch_count
    Tn, sn = T.reshape(-1, 1), s.reshape(-1, 2)
    q, dT, ds = hf((0, 0))(Tn, sn), hf((1, 0))(Tn, sn), hf((0, 1))(Tn, sn)
    assert ctx.launch_count - launches == 1  # one fused launch serves the three requests
    assert np.array_equal(q, native.heat("q", T, s))
    assert np.array_equal(dT, native.heat("dqdT", T, s))
    assert np.array_equal(ds, native.heat("dqdsigma", T, s))
    with pytest.raises(NotImplementedError):
        hf((1, 1))


def test_heat_unaligned_device_pointers(ctx):
    n = 1001
    T, s = inputs.heat_batch(n, seed=1)
    d_T, d_s, d_q = ctx.empty((n + 1,)), ctx.empty((2 * n + 1,)), ctx.empty((2 * n + 1,))
    ctx.copy(d_T.ptr + 8, T, T.nbytes)
    ctx.copy(d_s.ptr + 8, s, s.nbytes)
    ctx.check(ctx.lib.eo_heat_eval(ctx.handle, 1.0, 1.0, d_T.ptr + 8, d_s.ptr + 8, None, None, d_q.ptr + 8, None, None, n))
    q = np.empty(2 * n)
    ctx.copy(q, d_q.ptr + 8, q.nbytes)
    assert np.array_equal(q, native.heat("q", T, s))


def test_heat_argument_errors(ctx):
    T = np.ones(4)
    assert ctx.lib.eo_heat_eval(ctx.handle, 1.0, 1.0, T.ctypes.data, None, None, None, None, None, None, 4) == -1
    out = np.empty(8)
    assert ctx.lib.eo_heat_eval(ctx.handle, 1.0, 1.0, T.ctypes.data, None, None, None, out.ctypes.data, None, None, 4) == -1
    assert b"sigma" in ctx.lib.eo_last_error(ctx.handle)


def test_device_arrays_return_their_memory(ctx):
    """Owning DeviceArrays free their allocation when collected (the default drop-in path allocates one operand
    array per evaluate_operands call); views keep the owner alive."""
    import gc

    import torch

    ctx.sync()
    gc.collect()
    free0 = torch.cuda.mem_get_info(ctx.device)[0]
    for _ in range(40):
        a = ctx.empty((1 << 24,))  # 128 MB each: 5 GB in total if leaked
        v = a.reshape(-1, 4)
        del a
        assert v._base is not None and v.ptr
        ctx.check(ctx.lib.eo_dev_memset(ctx.handle, v.ptr, 0, v.nbytes))  # still valid through the view
        del v
    gc.collect()
    free1 = torch.cuda.mem_get_info(ctx.device)[0]
    assert free0 - free1 < (512 << 20)


def test_heat_flux_cache_is_per_evaluation_round(ctx):
    """An operand buffer refilled in place between two evaluate_external_operators rounds is never served stale."""
    from dolfinx_external_operator_b200.context import new_evaluation_round

    hf = eo.HeatFlux(ctx=ctx)
    T, g = inputs.heat_batch(999, seed=3)
    Tb, gb = T.reshape(-1, 3).copy(), g.reshape(-1, 6).copy()
    q1 = np.array(hf((0, 0))(Tb, gb))
    Tb += 0.5
    new_evaluation_round()
    q2 = np.array(hf((0, 0))(Tb, gb))
    np.testing.assert_allclose(q2, native.heat("q", Tb.reshape(-1), gb.reshape(-1, 2)).reshape(-1), rtol=1e-12)
    assert not np.array_equal(q1, q2)
