"""CPU: the host/device "core" headers of the CUDA kernels (csrc/*_core.cuh), compiled with g++ by
tests/hostcheck/Makefile, against the oracle.  This checks the kernels' ALGEBRA (e.g. the Taylor-jet
restatement of the reference's nested forward-mode AD for Mohr-Coulomb) on a machine without a GPU; the
GPU parity tests proper are tests/test_*_gpu.py.  The harness library is test-only: the product package
never loads it."""

import ctypes as C
import dataclasses
import os
import subprocess

import numpy as np
import pytest

from oracle import constitutive as oc
from oracle import inputs, native
from mc_util import check_mc, check_mc_exact

HERE = os.path.dirname(os.path.abspath(__file__))
MC_PRM = oc.MohrCoulombParams()


@pytest.fixture(scope="module")
def hc():
    subprocess.check_call(["make", "-C", os.path.join(HERE, "hostcheck")], stdout=subprocess.DEVNULL)
    return C.CDLL(os.path.join(HERE, "hostcheck", "libhostcheck.so"))


def _mc(hc, deps, sn, prm):
    n = deps.shape[0]
    deps, sn = np.ascontiguousarray(deps), np.ascontiguousarray(sn)
    q = native._mc_prm(prm)
    out = {"C_tang": np.empty((n, 4, 4)), "sigma": np.empty((n, 4)), "niter": np.empty(n, dtype=np.int32),
           "yielding": np.empty(n), "norm_res": np.empty(n), "dlambda": np.empty(n)}
    p = lambda a: a.ctypes.data_as(C.c_void_p)  # noqa: E731
    hc.hostcheck_mc(C.byref(q), p(deps), p(sn), p(out["C_tang"]), p(out["sigma"]), p(out["niter"]),
                    p(out["yielding"]), p(out["norm_res"]), p(out["dlambda"]), C.c_int64(n))
    return out


@pytest.mark.parametrize("name", ["mc_path_10x9.npz", "mc_rand_seed0_n96.npz", "mc_rand_seed1_n2048.npz"])
def test_mc_core_against_reference_golden(hc, golden_dir, name):
    g = np.load(os.path.join(golden_dir, name))
    o = _mc(hc, g["deps"], g["sigma_n"], MC_PRM)
    check_mc(o, g, g["deps"], g["sigma_n"], MC_PRM)
    check_mc_exact(o, g["deps"], g["sigma_n"], MC_PRM)  # flat 1e-10 at every point, corners included


@pytest.mark.parametrize("psi_deg", [30, 10])
def test_mc_core_against_oracle(hc, psi_deg):
    prm = dataclasses.replace(MC_PRM, psi=psi_deg * np.pi / 180)
    step = lambda d, s: native.mc_stress(d, s, prm, parallel=True)[0]  # noqa: E731
    d, s = inputs.mc_batch(20_000, seed=3, stepper=step)
    o = _mc(hc, d, s, prm)
    check_mc(o, native.mc_return_mapping(d, s, prm, parallel=True), d, s, prm)
    check_mc_exact(o, d, s, prm)


def test_mc_core_edge_semantics(hc):
    d = np.array([[0.0, 0, 0, 0], [0.0, 0, 0, 0], [1e-3, -1e-3, 0, 0]])
    s = np.array([[-1.0, -1.2, -0.9, 0.1], [-1.0, -1, -1, 0], [-1.0, -1, -1, 0]])
    o, r = _mc(hc, d, s, MC_PRM), native.mc_return_mapping(d, s, MC_PRM)
    assert list(o["niter"]) == [0, 0, 0] == list(r["niter"])
    assert np.array_equal(o["sigma"], r["sigma"])
    assert np.array_equal(np.isnan(o["C_tang"]), np.isnan(r["C_tang"]))


@pytest.mark.parametrize("batch", ["mixed", "elastic", "plastic"])
@pytest.mark.parametrize("exact", [True, False])
def test_vm_core_and_factored_tangent(hc, golden_dir, exact, batch):
    """vm_core.cuh on the host: both forms reproduce the golden made by the reference's own Numba kernel (flags exactly,
    values to 1e-12), the exact statement sequence is bit-identical to the C oracle; the tangent's factors (v, cn, cd) rebuild
    C_t = C_elas - cn v v^T - cd dev, and vm_factored_apply(e) == C_t e (what the factored tangent action computes)."""
    g = np.load(os.path.join(golden_dir, "vm_seed0_n1026.npz"))
    deps, sn, p = (np.ascontiguousarray(g[f"{batch}_{k}"], dtype=np.float64) for k in ("deps", "sigma_n", "p"))
    n = p.size
    prm = oc.VonMisesParams()
    rng = np.random.default_rng(3)
    e = rng.normal(size=(n, 4))
    out = {k: np.empty(s) for k, s in (("C", (n, 4, 4)), ("sig", (n, 4)), ("dp", n), ("T6", (n, 6)), ("tau", (n, 4)))}
    ptr = lambda a: a.ctypes.data_as(C.c_void_p)  # noqa: E731
    hc.hostcheck_vm(C.c_double(prm.lmbda), C.c_double(prm.mu), C.c_double(prm.H), C.c_double(prm.sigma_0), ptr(deps), ptr(sn),
                    ptr(p), C.c_int64(n), C.c_int(int(exact)), ptr(out["C"]), ptr(out["sig"]), ptr(out["dp"]), ptr(out["T6"]),
                    ptr(e), ptr(out["tau"]))
    gC, gs, gdp = g[f"{batch}_C_tang"].reshape(n, 4, 4), g[f"{batch}_sigma"].reshape(n, 4), g[f"{batch}_dp"].reshape(n)
    assert np.array_equal(out["dp"] > 0, gdp > 0)
    if batch == "mixed":
        assert 0.2 < (gdp > 0).mean() < 0.8
    for a, b in ((out["C"], gC), (out["sig"], gs), (out["dp"], gdp)):  # the golden itself: rounding (Numba / NumPy order)
        assert np.abs(a - b).max() <= 1e-12 * max(np.abs(b).max(), 1e-300)
    if exact:  # same statement order as the non-contracting C oracle: identical bits (what the GPU kernel is held to)
        rC, rs, rdp = native.vm_return_mapping(deps, sn, p, prm)
        assert np.array_equal(out["C"], rC.reshape(n, 4, 4)) and np.array_equal(out["sig"], rs.reshape(n, 4))
        assert np.array_equal(out["dp"], np.asarray(rdp).reshape(n))
    # the factors
    v, cn, cd = out["T6"][:, :4], out["T6"][:, 4], out["T6"][:, 5]
    el = gdp == 0
    assert np.all(v[el] == 0.0) and np.all(cd[el] == 0.0)
    l, m = prm.lmbda, prm.mu
    Cel = np.array([[l + 2 * m, l, l, 0], [l, l + 2 * m, l, 0], [l, l, l + 2 * m, 0], [0, 0, 0, 2 * m]])
    dev = np.eye(4) - np.outer([1, 1, 1, 0], [1, 1, 1, 0]) / 3.0
    Cr = Cel[None] - cn[:, None, None] * v[:, :, None] * v[:, None, :] - cd[:, None, None] * dev[None]
    assert np.abs(Cr - out["C"]).max() <= 1e-13 * np.abs(out["C"]).max()
    tau_ref = np.einsum("nij,nj->ni", out["C"], e)
    assert np.abs(out["tau"] - tau_ref).max() <= 1e-12 * np.abs(tau_ref).max()
