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
