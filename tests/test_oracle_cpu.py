"""CPU: the oracle (oracle/) against the golden vectors produced by the reference's own code
(oracle/gen_golden.py), and - when /root/reference is present - against the reference executed live."""

import os

import numpy as np
import pytest

from oracle import constitutive as oc
from oracle import inputs, native, ref_exec

RTOL = 1e-12  # closed-form models: north_star tolerance


def _close(a, b, rtol=RTOL):
    # absolute floor = rtol x the field's scale: components that cancel to ~0 carry the rounding
    # error of the O(scale) terms they were formed from
    np.testing.assert_allclose(a, b, rtol=rtol, atol=rtol * np.abs(b).max())


@pytest.mark.parametrize("kind", ["mixed", "elastic", "plastic"])
@pytest.mark.parametrize("impl", ["numpy", "c"])
def test_vm_oracle_matches_reference_golden(golden_dir, kind, impl):
    g = np.load(os.path.join(golden_dir, "vm_seed0_n1026.npz"))
    prm = oc.VonMisesParams()
    np.testing.assert_allclose(g["params"][4:], [prm.H, prm.lmbda, prm.mu], rtol=0, atol=0)
    fn = oc.vm_return_mapping if impl == "numpy" else native.vm_return_mapping
    Ct, sig, dp = fn(g[f"{kind}_deps"], g[f"{kind}_sigma_n"], g[f"{kind}_p"], prm)
    assert np.array_equal(dp > 0, g[f"{kind}_dp"] > 0)  # plastic flags bit-exact
    _close(Ct, g[f"{kind}_C_tang"])
    _close(sig, g[f"{kind}_sigma"])
    _close(dp, g[f"{kind}_dp"])
    if kind == "elastic":  # elastic points: C_t == C_elas and dp == 0 exactly (demo_vm:313-324)
        assert np.all(dp == 0.0)
        assert np.array_equal(Ct, np.broadcast_to(oc.elastic_stiffness(prm.lmbda, prm.mu), Ct.shape))


def test_vm_plastic_fraction_of_the_synthetic_batch():
    deps, sn, p = inputs.vm_batch(200_000, seed=0)
    _, _, dp = native.vm_return_mapping(deps, sn, p, oc.VonMisesParams(), parallel=True)
    assert 0.50 < (dp > 0).mean() < 0.56  # ~53 % plastic


def test_vm_parallel_equals_serial():
    deps, sn, p = inputs.vm_batch(10_001, seed=3)
    a = native.vm_return_mapping(deps, sn, p, oc.VonMisesParams(), parallel=False)
    b = native.vm_return_mapping(deps, sn, p, oc.VonMisesParams(), parallel=True)
    for x, y in zip(a, b):
        assert np.array_equal(x, y)


@pytest.mark.parametrize("which", ["k", "dk", "q", "dqdT", "dqdsigma"])
def test_heat_oracle_matches_reference_golden(golden_dir, which):
    g = np.load(os.path.join(golden_dir, "heat_seed0_n4098.npz"))
    out = native.heat(which, g["T"], g["sigma"])
    assert np.array_equal(out, g[which])  # bit-exact
    T2, s2 = g["T"].reshape(-1, 3), g["sigma"].reshape(-1, 6)
    py = {"k": lambda: oc.heat_k(T2).reshape(-1), "dk": lambda: oc.heat_dkdT(T2).reshape(-1),
          "q": lambda: oc.heat_q(T2, s2), "dqdT": lambda: oc.heat_dqdT(T2, s2),
          "dqdsigma": lambda: oc.heat_dqdsigma(T2, s2)}[which]()
    assert np.array_equal(py, g[which])


def test_heat_analytic_derivative():
    # part1.py:442-448: dk/dT = -B k^2 checked by central differences
    T = np.linspace(0.1, 1.9, 50)
    h = 1e-6
    fd = (oc.heat_k(T + h) - oc.heat_k(T - h)) / (2 * h)
    np.testing.assert_allclose(oc.heat_dkdT(T), fd, rtol=1e-8)


@pytest.mark.skipif(not ref_exec.reference_available(), reason="/root/reference absent (GPU box)")
def test_vm_oracle_against_live_reference():
    ns = ref_exec.load_von_mises(3)
    deps, sn, p = inputs.vm_batch(999, seed=7)
    rC, rs, rdp = ns["return_mapping"](deps.reshape(-1, 3, 4), sn.reshape(-1, 3, 4), p.reshape(-1, 3))
    Ct, sig, dp = native.vm_return_mapping(deps, sn, p, oc.VonMisesParams())
    assert np.array_equal(dp > 0, rdp.reshape(-1) > 0)
    _close(Ct, rC.reshape(-1, 4, 4))
    _close(sig, rs.reshape(-1, 4))
    _close(dp, rdp.reshape(-1))


# ----------------------------------------------------------------------------- Mohr-Coulomb
MC_RTOL = 1e-10  # local-Newton models: north_star tolerance
MC_PRM = oc.MohrCoulombParams()


def _mc_check(o, g, rtol=MC_RTOL):
    assert np.array_equal(o["niter"], g["niter"])
    assert np.array_equal(o["yielding"] > 0, g["yielding"] > 0)  # plastic flags bit-exact
    _close(o["C_tang"], g["C_tang"], rtol)
    _close(o["sigma"], g["sigma"], rtol)
    _close(o["dlambda"], g["dlambda"], rtol)
    _close(o["yielding"], g["yielding"], rtol)
    # ||res|| at exit is a converged residual (<= 1e-8 ||res0||): rounding-level quantity, absolute check
    np.testing.assert_allclose(o["norm_res"], g["norm_res"], rtol=0, atol=1e-10 * np.abs(g["sigma"]).max())


@pytest.mark.parametrize("name", ["mc_path_10x9.npz", "mc_rand_seed0_n96.npz"])
def test_mc_oracle_matches_reference_golden(golden_dir, name):
    """The C++ dual-number restatement against the reference's own source (demo_mc:282-555) executed
    over the torch.func shim of the JAX API (oracle/gen_golden.py gen_mc)."""
    g = np.load(os.path.join(golden_dir, name))
    o = native.mc_return_mapping(g["deps"], g["sigma_n"], MC_PRM)
    _mc_check(o, g)
    el = g["yielding"] <= 0
    C = oc.elastic_stiffness(MC_PRM.lmbda, MC_PRM.mu)
    assert np.array_equal(o["C_tang"][el], np.broadcast_to(C, (el.sum(), 4, 4)))  # demo_mc:442-443
    assert np.all(o["niter"][el] == 1) and np.all(o["dlambda"][el] == 0.0)
    # returned plastic stresses lie on the yield surface: |f| <= tol * ||res0||
    f = native.mc_yield(o["sigma"][~el], MC_PRM)
    assert np.abs(f).max() < 1e-6


def test_mc_demo_tracing_histogram():
    """the demo's 50 x 9 tracing driver (demo_mc:853-930) gives 276 elastic points
    and 99/57/15/3 plastic points with 2/3/4/5 local Newton iterations."""
    step = lambda d, s: native.mc_stress(d, s, MC_PRM, parallel=True)[0]  # noqa: E731
    d, s = inputs.mc_demo_path(50, 9, stepper=step)
    o = native.mc_return_mapping(d, s, MC_PRM, parallel=True)
    it, cnt = np.unique(o["niter"], return_counts=True)
    assert dict(zip(it.tolist(), cnt.tolist())) == {1: 276, 2: 99, 3: 57, 4: 15, 5: 3}


def test_mc_synthetic_batch_is_isotropic():
    """The in-plane rotation that populates the shear component must not change invariants."""
    step = lambda d, s: native.mc_stress(d, s, MC_PRM, parallel=True)[0]  # noqa: E731
    d0, s0 = inputs.mc_batch(500, 1, step, rotate=False)
    d1, s1 = inputs.mc_batch(500, 1, step, rotate=True)
    o0, o1 = native.mc_return_mapping(d0, s0, MC_PRM), native.mc_return_mapping(d1, s1, MC_PRM)
    assert np.array_equal(o0["niter"], o1["niter"])
    np.testing.assert_allclose(o0["yielding"], o1["yielding"], atol=1e-12)
    assert np.abs(d1[:, 3]).max() > 0


@pytest.mark.skipif(not ref_exec.reference_available(), reason="/root/reference absent (GPU box)")
def test_mc_oracle_against_live_reference():
    import torch

    ns = ref_exec.load_mohr_coulomb()
    de = np.array([1.1e-4, -2.3e-4, 0.0, 0.7e-4])
    sn = np.array([-1.0, -1.2, -0.9, 0.1])
    Ct, aux = ns["dsigma_ddeps"](torch.as_tensor(de), torch.as_tensor(sn))
    o = native.mc_return_mapping(de[None], sn[None], MC_PRM)
    assert int(aux[1]) == int(o["niter"][0])
    _close(o["C_tang"][0], Ct.numpy(), MC_RTOL)
    _close(o["sigma"][0], aux[0].numpy(), MC_RTOL)


def test_numba_restatement_is_bit_identical_to_the_reference_golden(golden_dir):
    """oracle/numba_vm.py (the serial-Numba execution model bench.py times as the reference's own CPU callable) against
    the golden made by the reference's Numba kernel."""
    numba = pytest.importorskip("numba")  # noqa: F841
    from oracle import numba_vm

    g = np.load(os.path.join(golden_dir, "vm_seed0_n1026.npz"))
    f = numba_vm.make_return_mapping()
    for kind in ("mixed", "elastic", "plastic"):
        Ct, s, dp = f(*(np.ascontiguousarray(g[f"{kind}_{k}"]) for k in ("deps", "sigma_n", "p")))
        assert np.array_equal(Ct.reshape(-1), g[f"{kind}_C_tang"].reshape(-1))
        assert np.array_equal(s.reshape(-1), g[f"{kind}_sigma"].reshape(-1)) and np.array_equal(dp, g[f"{kind}_dp"])


def test_isihara_torch_restatement_against_the_reference_golden(golden_dir):
    """oracle/isihara_torch.py (functional torch restatement, the CPU baseline of the Isihara bench leg) against the
    golden made by the reference's own module: same float32 network, so agreement is far below the 2e-6 parity bound."""
    from oracle import isihara_torch as it

    g = np.load(os.path.join(golden_dir, "isihara_seed0_n2049.npz"))
    sd = {k[3:]: g[k] for k in g.files if k.startswith("sd/")}
    dP, P = it.dP_dF(g["F"][:300], sd)
    assert np.abs(dP - g["dP"][:300]).max() <= 1e-8 * np.abs(g["dP"]).max()
    assert np.abs(P - g["P"][:300]).max() <= 1e-8 * np.abs(g["P"]).max()
    np.testing.assert_allclose(it.stress_correction(it._sd_tensors(sd)).numpy(), g["H_flat"], rtol=0, atol=1e-12)
