"""GPU: the kernels at BASELINE.json's full per-GPU size (1e8 quadrature points, EO_FULLSIZE_N to override),
checked through size-independent properties evaluated ON the device (torch is used only to look at the
results): the batch is a seeded tile repeated to full size, so (i) every repetition must reproduce the first
tile bit for bit (no dependence on position, chunking or scheduling), (ii) the first tile must match the CPU
oracle, (iii) the device statistics must be the tile's statistics times the repetition count; for the
tabulation a linear displacement gives a constant, exactly known strain at every point."""

import ctypes as C
import os

import numpy as np
import pytest

import dolfinx_external_operator_b200 as eo
from dolfinx_external_operator_b200 import elements as el
from dolfinx_external_operator_b200 import synthetic as syn
from dolfinx_external_operator_b200._lib import McParams, VmParams
from mc_util import check_mc
from oracle import constitutive as oc
from oracle import native

pytestmark = pytest.mark.gpu
N = int(float(os.environ.get("EO_FULLSIZE_N", "1e8")))
TILE = 1 << 20


def _t(arr, dtype=None):
    import torch

    t = torch.as_tensor(arr, device=f"cuda:{arr.ctx.device}")
    return t


def _fill(ctx, dst, tile, width, n):
    d_tile = ctx.to_device(np.ascontiguousarray(tile).reshape(-1))
    for r in range(0, n, tile.shape[0]):
        m = min(tile.shape[0], n - r)
        ctx.copy(dst.ptr + r * width * 8, d_tile, m * width * 8)
    ctx.sync()
    d_tile.free()


def _periodic(ctx, arr, width, n, tile_n=TILE):
    """All full repetitions of the tile equal the first one, bit for bit (compared as int64 patterns)."""
    import torch

    ctx.sync()
    reps = n // tile_n
    t = _t(arr).view(torch.int64)[: reps * tile_n * width].view(reps, tile_n * width)
    first = t[0]
    for r in range(1, reps):
        if not torch.equal(t[r], first):
            return False
    return True


def test_von_mises_full_size(ctx):
    torch = pytest.importorskip("torch")
    n = N
    deps_t, sn_t, p_t = syn.vm_batch(TILE, seed=3)
    d = {k: ctx.empty((n * w,)) for k, w in (("deps", 4), ("sn", 4), ("p", 1), ("Ct", 16), ("sig", 4), ("dp", 1))}
    _fill(ctx, d["deps"], deps_t, 4, n), _fill(ctx, d["sn"], sn_t, 4, n), _fill(ctx, d["p"], p_t, 1, n)
    prm = oc.VonMisesParams()
    q = VmParams(prm.lmbda, prm.mu, prm.H, prm.sigma_0)
    ctx.stats_reset()
    ctx.check(ctx.lib.eo_vm_eval_resident(ctx.handle, C.byref(q), d["deps"].ptr, d["sn"].ptr, d["p"].ptr, d["Ct"].ptr,
                                          d["sig"].ptr, d["dp"].ptr, n, 0))
    ctx.sync()
    for k, w in (("Ct", 16), ("sig", 4), ("dp", 1)):
        assert _periodic(ctx, d[k], w, n), k
    rC, rs, rdp = native.vm_return_mapping(deps_t, sn_t, p_t, prm, parallel=True)
    assert np.array_equal(_t(d["Ct"])[: 16 * TILE].cpu().numpy(), rC.reshape(-1))  # bit-exact vs the oracle
    assert np.array_equal(_t(d["sig"])[: 4 * TILE].cpu().numpy(), rs.reshape(-1))
    assert np.array_equal(_t(d["dp"])[:TILE].cpu().numpy(), rdp)
    st = ctx.stats()
    reps, rem = divmod(n, TILE)
    assert st["n_points"] == n
    assert st["n_plastic"] == reps * int((rdp > 0).sum()) + int((rdp[:rem] > 0).sum())
    # elastic points carry C_elas exactly, everywhere
    dp_all, Ct_all = _t(d["dp"]), _t(d["Ct"]).view(n, 16)
    Cel = torch.tensor(oc.elastic_stiffness(prm.lmbda, prm.mu).reshape(-1), device=Ct_all.device)
    chunk = 1 << 24
    for a in range(0, n, chunk):
        el_mask = dp_all[a:a + chunk] == 0
        assert bool((Ct_all[a:a + chunk][el_mask] == Cel).all())
    for a in d.values():
        a.free()


def test_mohr_coulomb_full_size(ctx):
    pytest.importorskip("torch")
    n = N
    prm = oc.MohrCoulombParams()
    d_t, s_t = syn.mc_batch(TILE, seed=5, stepper=lambda dd, ss: native.mc_stress(dd, ss, prm, parallel=True)[0])
    d = {k: ctx.empty((n * w,)) for k, w in (("deps", 4), ("sn", 4), ("Ct", 16), ("sig", 4), ("yl", 1), ("nr", 1), ("dl", 1))}
    d_it = ctx.empty((n,), np.int32)
    _fill(ctx, d["deps"], d_t, 4, n), _fill(ctx, d["sn"], s_t, 4, n)
    q = McParams(prm.E, prm.nu, prm.c, prm.phi, prm.psi, prm.theta_T, prm.a, prm.tol, prm.Nitermax)
    ctx.stats_reset()
    ctx.check(ctx.lib.eo_mc_eval(ctx.handle, C.byref(q), d["deps"].ptr, d["sn"].ptr, d["Ct"].ptr, d["sig"].ptr, d_it.ptr,
                                 d["yl"].ptr, d["nr"].ptr, d["dl"].ptr, n))
    ctx.sync()
    # position / scheduling independence: which warp, slot or stage batch a point lands in must not matter
    for k, w in (("Ct", 16), ("sig", 4), ("yl", 1), ("nr", 1), ("dl", 1)):
        assert _periodic(ctx, d[k], w, n), k
    it_all = _t(d_it)
    reps, rem = divmod(n, TILE)
    it0 = it_all[:TILE]
    assert bool((it_all[: reps * TILE].view(reps, TILE) == it0).all())
    # the first 40 000 points against the oracle (the C++ nested-dual restatement costs ~0.2 ms per plastic point)
    m = 40_000
    ref = native.mc_return_mapping(d_t[:m], s_t[:m], prm, parallel=True)
    out = {"C_tang": _t(d["Ct"])[: 16 * m].cpu().numpy().reshape(m, 4, 4), "sigma": _t(d["sig"])[: 4 * m].cpu().numpy().reshape(m, 4),
           "niter": it0[:m].cpu().numpy(), "yielding": _t(d["yl"])[:m].cpu().numpy(), "norm_res": _t(d["nr"])[:m].cpu().numpy(),
           "dlambda": _t(d["dl"])[:m].cpu().numpy()}
    check_mc(out, ref, d_t[:m], s_t[:m], prm)
    # statistics record = tile statistics x repetitions (demo_mc:584-591 at full size, from the device)
    st = ctx.stats()
    it_tile = it0.cpu().numpy()
    hist_tile = np.bincount(it_tile, minlength=208)
    hist_rem = np.bincount(it_tile[:rem], minlength=208)
    assert np.array_equal(st["niter_hist"], reps * hist_tile + hist_rem)
    assert st["n_points"] == n and st["n_nonconverged"] == 0 and st["n_nonfinite"] == 0
    yl_tile = _t(d["yl"])[:TILE].cpu().numpy()
    assert st["f_max"] == yl_tile.max() and st["n_plastic"] == reps * int((yl_tile > 0).sum()) + int((yl_tile[:rem] > 0).sum())
    # returned plastic stresses lie on the yield surface
    sig0 = _t(d["sig"])[: 4 * m].cpu().numpy().reshape(m, 4)
    assert np.abs(native.mc_yield(sig0[out["yielding"] > 0], prm)).max() < 1e-6
    for a in d.values():
        a.free()
    d_it.free()


def test_tabulation_full_size_known_answer(ctx):
    torch = pytest.importorskip("torch")
    nxy = max(2, int(round((N / 6.0) ** 0.5)))
    mesh = syn.triangle_mesh(nxy, nxy, 2, jitter=0.2, seed=1)
    n_cells = mesh["dofmap"].shape[0]
    phi, dphi = el.lagrange_triangle(2, el.triangle_quadrature(2))
    tab = eo.Tabulator(dofmap=mesh["dofmap"], x_dofmap=mesh["x_dofmap"], x=mesh["x"], phi=phi, dphi=dphi, bs=2,
                       n_dofs=mesh["n_dofs"], ctx=ctx)
    xy = mesh["dof_coords"]
    u = np.stack([0.1 * xy[:, 0] + 0.05 * xy[:, 1], 0.3 * xy[:, 1] - 0.02 * xy[:, 0]], 1).reshape(-1)  # linear: P2-exact
    del mesh
    e = tab.evaluate("mandel_strain", u)  # (n_cells, 3, 4) on the device
    ctx.sync()
    t = _t(e).view(-1, 4)
    expect = torch.tensor([0.1, 0.3, 0.0, np.sqrt(2.0) * 0.5 * (0.05 - 0.02)], device=t.device, dtype=torch.float64)
    assert t.shape[0] == 3 * n_cells
    assert float((t - expect).abs().max()) < 1e-10  # gradients of a 1e-1 field on cells of size 1/nxy: eps * nxy
    assert bool((t[:, 2] == 0).all())
    # fused kernel at full size: identical flags and values to a few ulp between its exact and fast variants, and
    # the exact variant bit-equal to the two-step path on the stored strain
    n = 3 * n_cells
    _, sn_t, p_t = syn.vm_batch(TILE, seed=7)
    vm_a, vm_b = eo.VonMises(n_qp=n, ctx=ctx), eo.VonMises(n_qp=n, ctx=ctx)
    for vm in (vm_a, vm_b):
        _fill(ctx, vm.sigma_n_dev, sn_t, 4, n), _fill(ctx, vm.p_dev, p_t, 1, n)
    u2 = syn.smooth_displacement(xy, scale=1.5e-3, seed=1).reshape(-1)
    d_u = ctx.to_device(u2)
    Ct_a = tab.vm_fused(vm_a, d_u, exact=True)
    strain = tab.evaluate("mandel_strain", d_u, out=e)
    Ct_b = ctx.empty((16 * n,))
    vm_b.eval_device(strain.reshape(-1), Ct_b)
    ctx.sync()
    assert torch.equal(_t(Ct_a).view(torch.int64), _t(Ct_b).view(torch.int64))
    assert torch.equal(_t(vm_a.dp_dev).view(torch.int64), _t(vm_b.dp_dev).view(torch.int64))
    Ct_c = tab.vm_fused(vm_b, d_u, C_tang=Ct_b)  # fast variant, overwrites Ct_b
    ctx.sync()
    assert torch.equal(_t(vm_a.dp_dev) > 0, _t(vm_b.dp_dev) > 0)
    scale = float(_t(Ct_a).abs().max())
    assert float((_t(Ct_c) - _t(Ct_a)).abs().max()) <= 1e-12 * scale
    frac = float((_t(vm_a.dp_dev) > 0).double().mean())
    assert 0.2 < frac < 0.8
