"""CPU tests of the run-time compiled (NVRTC) path: compilation for sm_100a works without a GPU, compile
errors surface with NVRTC's log, and the model texts + the dual-number algebra (include/eo_dual.h, built with
g++ by tests/jit_util.py) reproduce the reference's hand-derived tangents stored in the golden vectors."""

import os

import numpy as np
import pytest

from dolfinx_external_operator_b200 import jit_models as jm
from dolfinx_external_operator_b200._lib import EOError
from dolfinx_external_operator_b200.jit import JitModel
from jit_util import host_eval


def _close(a, b, rtol):
    b = np.asarray(b).reshape(-1)
    np.testing.assert_allclose(np.asarray(a).reshape(-1), b, rtol=rtol, atol=rtol * np.abs(b).max())


def test_compiles_for_sm100a_without_gpu():
    m = jm.von_mises(compile_only=True)
    for d in [(0,), (1,), (2,)]:
        assert m.compile(d) > 1000
    assert m.out_width((0,)) == 4 and m.out_width((1,)) == 16 and m.out_width((2,)) == 64
    cubin = m.cubin((1,))
    assert cubin[:4] == b"\x7fELF"
    h = jm.heat_flux(compile_only=True)
    assert [h.out_width(d) for d in [(0, 0), (1, 0), (0, 1), (1, 1), (2, 0), (0, 2)]] == [2, 2, 4, 4, 2, 8]
    for d in [(0, 0), (1, 0), (0, 1), (1, 1)]:
        assert h.compile(d) > 1000


def test_toolkit_nvrtc_is_preferred_over_the_one_torch_loads():
    """PyTorch brings its own libnvrtc.so.12 (12.8) into the process; its ptxas cannot assemble the 256-bit
    global accesses of sm_100, so the library must pick the toolkit's NVRTC by path (and degrade to 128-bit
    accesses when only an older one exists)."""
    import torch  # noqa: F401  - loads the bundled NVRTC first

    from dolfinx_external_operator_b200 import _lib

    v = _lib.load().eo_jit_nvrtc_version()
    assert v >= 12000
    m = jm.von_mises(compile_only=True)
    assert m.compile((1,)) > 1000


def test_fused_tabulation_variants_compile_without_gpu():
    """eo_jit_compile_tabulated: the model kernel with the operand tabulation fused in (P2 vector triangle -> Mandel
    strain -> von Mises; P2 scalar triangle -> (T, grad T) -> heat flux; P1 tetrahedron -> grad)."""
    import ctypes as C

    ints = lambda *a: (C.c_int * len(a))(*a)  # noqa: E731
    nbytes = C.c_size_t(0)
    m = jm.von_mises(compile_only=True)
    assert m.lib.eo_jit_compile_tabulated(m._h, ints(1), 3, ints(2), ints(2), ints(6), ints(2), C.byref(nbytes)) == 0
    assert nbytes.value > 1000
    h = jm.heat_flux(compile_only=True)
    for d in [(0, 0), (1, 0), (0, 1)]:
        assert h.lib.eo_jit_compile_tabulated(h._h, ints(*d), 3, ints(2, 2), ints(1, 1), ints(6, 6), ints(0, 1), C.byref(nbytes)) == 0
    # operand size 2 does not match a Mandel strain (4 components): rejected
    assert h.lib.eo_jit_compile_tabulated(h._h, ints(0, 0), 3, ints(2, 2), ints(1, 2), ints(6, 6), ints(0, 2), C.byref(nbytes)) == -1


def test_points_per_thread_variant_compiles_for_scalar_sized_models():
    import ctypes as C

    ints = lambda *a: (C.c_int * len(a))(*a)  # noqa: E731
    ppt, nbytes = C.c_int(0), C.c_size_t(0)
    k = jm.heat_conductivity(compile_only=True)
    assert k.lib.eo_jit_compile_ppt(k._h, ints(1), C.byref(ppt), C.byref(nbytes)) == 0 and ppt.value == 4
    q = jm.heat_flux(compile_only=True)
    assert q.lib.eo_jit_compile_ppt(q._h, ints(0, 1), C.byref(ppt), C.byref(nbytes)) == 0 and ppt.value == 2
    vm = jm.von_mises(compile_only=True)  # 9 doubles read per point: one point per thread
    assert vm.lib.eo_jit_compile_ppt(vm._h, ints(1), C.byref(ppt), C.byref(nbytes)) == -4 and ppt.value == 1


def test_cubin_cache_on_disk(tmp_path):
    """EO_JIT_CACHE_DIR: the second process-lifetime of a model loads its CUBIN instead of compiling it."""
    import subprocess
    import sys

    code = ("import sys; sys.path.insert(0, %r)\n"
            "from dolfinx_external_operator_b200 import jit_models as jm\n"
            "m = jm.von_mises(compile_only=True); n = m.compile((1,)); print(n, m.log[:11])\n") % str(__import__('conftest').ROOT)
    env = dict(__import__('os').environ, EO_JIT_CACHE_DIR=str(tmp_path))
    a = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, check=True).stdout.split()
    b = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, check=True).stdout.split()
    assert int(a[0]) == int(b[0]) > 1000
    assert a[1:] != ["loaded", "from"] and b[1:] == ["loaded", "from"]
    assert len(list(tmp_path.glob("eo_jit_*.cubin"))) == 1


def test_compile_error_carries_the_nvrtc_log():
    bad = "template <class T> __device__ void f(const T* x, const double*, const double*, T* y, T*) { y[0] = x[0] +; }"
    m = JitModel(bad, "f", [()], (), compile_only=True)
    with pytest.raises(EOError) as e:
        m.compile((0,))
    assert "model.cu(1): error" in str(e.value)
    with pytest.raises(EOError):  # order 3 is not implemented
        jm.heat_conductivity(compile_only=True).compile((3,))
    with pytest.raises(EOError):  # evaluation needs a context
        jm.heat_conductivity(compile_only=True)._evaluate((0,), None, 1, [np.zeros(3)])


def test_protocol_errors():
    m = jm.von_mises(compile_only=True, supported=[(1,)])
    with pytest.raises(NotImplementedError):  # demo_vm:364-368
        m((0,))
    with pytest.raises(ValueError):
        m((1, 0))
    with pytest.raises(ValueError):
        JitModel("", "f", [()] * 9, (), compile_only=True)


@pytest.mark.parametrize("kind", ["mixed", "elastic", "plastic"])
def test_von_mises_ad_tangent_matches_reference_golden(golden_dir, kind):
    """d sigma / d deps by dual numbers == the reference's closed-form consistent tangent (demo_vm:322-326)."""
    g = np.load(os.path.join(golden_dir, "vm_seed0_n1026.npz"))
    m = jm.von_mises(compile_only=True)
    deps, sn, p = g[f"{kind}_deps"], g[f"{kind}_sigma_n"], g[f"{kind}_p"]
    Ct, sig, (dp,) = host_eval(m, (1,), [deps], [sn, p])
    assert np.array_equal(dp > 0, g[f"{kind}_dp"] > 0)
    _close(Ct, g[f"{kind}_C_tang"], 1e-12)
    _close(sig, g[f"{kind}_sigma"], 1e-12)
    _close(dp, g[f"{kind}_dp"], 1e-12)
    sig0, _, (dp0,) = host_eval(m, (0,), [deps], [sn, p])
    assert np.array_equal(sig0, sig) and np.array_equal(dp0, dp)


def neo_hookean_reference(F, mu=1.0, lam=2.0):
    """Closed forms for W = mu/2 (I1 - 3) - mu ln J + lam/2 ln^2 J:  P = mu (F - F^-T) + lam ln J F^-T,
    A_ijkl = mu d_ik d_jl + (mu - lam ln J) F^-T_il F^-T_kj + lam F^-T_ij F^-T_kl."""
    F = F.reshape(-1, 3, 3)
    Fit = np.linalg.inv(F).transpose(0, 2, 1)
    lnJ = np.log(np.linalg.det(F))
    W = 0.5 * mu * (np.einsum("nij,nij->n", F, F) - 3.0) - mu * lnJ + 0.5 * lam * lnJ**2
    P = mu * (F - Fit) + lam * lnJ[:, None, None] * Fit
    I = np.eye(3)
    A = (mu * np.einsum("ik,jl->ijkl", I, I)[None] + (mu - lam * lnJ)[:, None, None, None, None] * np.einsum("nil,nkj->nijkl", Fit, Fit)
         + lam * np.einsum("nij,nkl->nijkl", Fit, Fit))
    return W, P.reshape(-1, 9), A.reshape(-1, 81)


def neo_hookean_batch(n, seed):
    return np.eye(3).reshape(1, 9) + np.random.default_rng(seed).normal(0.0, 0.08, (n, 9))


def test_energy_only_hyperelastic_model_stress_and_tangent_by_ad():
    """P = dW/dF and dP/dF = d2W/dF2 from the energy text alone (nested duals, 9 x 9 directions) vs closed forms."""
    F = neo_hookean_batch(300, 0)
    m = jm.neo_hookean_3d(compile_only=True)
    W, P, A = neo_hookean_reference(F)
    _close(host_eval(m, (0,), [F])[0], W, 1e-13)
    _close(host_eval(m, (1,), [F])[0], P, 1e-12)
    _close(host_eval(m, (2,), [F])[0], A, 1e-11)
    assert m.out_width((1,)) == 9 and m.out_width((2,)) == 81
    import ctypes as C

    tile, nbytes = C.c_int(0), C.c_size_t(0)  # the staged (TMA bulk copy) variant compiles without a GPU
    d = (C.c_int * 1)(1)
    assert m.lib.eo_jit_compile_staged(m._h, d, C.byref(tile), C.byref(nbytes)) == 0
    assert tile.value in (32, 64, 128, 256, 512, 1024) and nbytes.value > 1000


def _vm3d_batch(n, seed):
    rng = np.random.default_rng(seed)
    deps = rng.normal(0.0, 2e-3, (n, 6))
    sn = rng.normal(0.0, 100.0, (n, 6))
    p = np.abs(rng.normal(0.0, 1e-3, n))
    return deps, sn, p


def test_von_mises_3d_extension_against_its_oracle():
    """EXTENSION (BASELINE config 5 '3D'): 6-component model text + dual-number tangent vs the NumPy restatement."""
    from oracle import constitutive as oc

    deps, sn, p = _vm3d_batch(2000, 0)
    m = jm.von_mises_3d(compile_only=True)
    assert m.compile((1,)) > 1000 and m.out_width((1,)) == 36
    Ct, sig, (dp,) = host_eval(m, (1,), [deps], [sn, p])
    rC, rs, rdp = oc.vm3d_return_mapping(deps, sn, p)
    assert np.array_equal(dp > 0, rdp > 0) and 0.2 < (rdp > 0).mean() < 0.9
    _close(Ct, rC, 1e-11)
    _close(sig, rs, 1e-12)
    _close(dp, rdp, 1e-12)
    # plane-strain consistency with the reference's 4-component model: embed [xx, yy, zz, sqrt2 xy]
    g4 = np.zeros((500, 6)), np.zeros((500, 6))
    d4, s4, p4 = (a[:500] for a in _vm3d_batch(500, 1))
    idx = [0, 1, 2, 5]
    g4[0][:, idx], g4[1][:, idx] = d4[:, :4], s4[:, :4]
    C6, s6, dp6 = oc.vm3d_return_mapping(g4[0], g4[1], p4)
    C4, sg4, dpp4 = oc.vm_return_mapping(d4[:, :4], s4[:, :4], p4)
    _close(s6[:, idx], sg4, 1e-12)
    _close(dp6, dpp4, 1e-12)
    _close(C6[:, idx][:, :, idx], C4, 1e-12)


def test_heat_ad_derivatives_match_reference_golden(golden_dir):
    g = np.load(os.path.join(golden_dir, "heat_seed0_n4098.npz"))
    T, s = g["T"], g["sigma"]
    m = jm.heat_flux(compile_only=True)
    _close(host_eval(m, (0, 0), [T, s])[0], g["q"], 1e-14)
    _close(host_eval(m, (1, 0), [T, s])[0], g["dqdT"], 1e-14)
    _close(host_eval(m, (0, 1), [T, s])[0], g["dqdsigma"], 1e-14)
    k = jm.heat_conductivity(compile_only=True)
    _close(host_eval(k, (0,), [T])[0], g["k"], 1e-14)
    _close(host_eval(k, (1,), [T])[0], g["dk"], 1e-14)
    # second derivatives (no golden in the reference): analytic, k = 1/(A+BT), q = -k sigma, A = B = 1
    kk = 1.0 / (1.0 + T)
    _close(host_eval(k, (2,), [T])[0], 2 * kk**3, 1e-13)
    d2 = host_eval(m, (1, 1), [T, s])[0].reshape(-1, 2, 1, 2)  # d2 q_i / dT dsigma_j = k^2 delta_ij
    ref = np.einsum("n,ij->nij", kk**2, np.eye(2))
    _close(d2[:, :, 0, :], ref, 1e-13)
    _close(host_eval(m, (2, 0), [T, s])[0].reshape(-1, 2), -2 * kk[:, None] ** 3 * s, 1e-13)
    assert np.all(host_eval(m, (0, 2), [T, s])[0] == 0.0)


ELEMENTARY = r"""
template <class T>
__device__ void elem(const T* x, const double*, const double* prm, T* y, T*) {
  const T a = x[0], b = x[1];
  y[0] = sin(a) * cos(b) + tan(a * 0.25);
  y[1] = exp(a * b) - log(1.0 + a * a) + expm1(b * 0.1) + log1p(a * a);
  y[2] = sqrt(a * a + b * b + 1.0) / (2.0 + a) + cbrt(1.0 + b * b);
  y[3] = atan2(a, 1.5 + b * b) + asin(0.3 * sin(a)) + acos(0.2 * cos(b)) + atan(a - b);
  y[4] = pow(1.0 + a * a, 1.5) + pow(2.0 + b * b, a) + pow(1.7, a * b);
  y[5] = tanh(a) + sinh(0.3 * b) * cosh(0.2 * a) + fabs(a - b) + fmax(a, b) * fmin(a * a, 0.5) - a / b + 3 / (2 + a * a);
  y[6] = eo::select(a > b, a * a * b, b * b * a) + (-a) * prm[0] + abs(a - 2.0 * b) + max(a, 0.1) * min(b, a * 3.0) + hypot(a, b);
}
"""


def test_dual_algebra_against_complex_step_and_finite_differences():
    """Every elementary function of eo_dual.h: first derivatives against central differences of the value
    evaluation, second derivatives against central differences of the first."""
    rng = np.random.default_rng(3)
    n = 200
    x = rng.uniform(0.2, 1.2, (n, 2))
    x[:, 1] += 0.05 * np.sign(x[:, 1] - x[:, 0])  # keep away from the kinks a == b
    m = JitModel(ELEMENTARY, "elem", [(2,)], (7,), params=[0.7], compile_only=True)
    f = lambda z: host_eval(m, (0,), [z])[0].reshape(n, 7)  # noqa: E731
    J = host_eval(m, (1,), [x])[0].reshape(n, 7, 2)
    h = 1e-6
    for j in range(2):
        e = np.zeros(2)
        e[j] = h
        fd = (f(x + e) - f(x - e)) / (2 * h)
        np.testing.assert_allclose(J[:, :, j], fd, rtol=2e-8, atol=2e-8)
    H = host_eval(m, (2,), [x])[0].reshape(n, 7, 2, 2)
    J1 = lambda z: host_eval(m, (1,), [z])[0].reshape(n, 7, 2)  # noqa: E731
    for j in range(2):
        e = np.zeros(2)
        e[j] = h
        fd = (J1(x + e) - J1(x - e)) / (2 * h)
        np.testing.assert_allclose(H[:, :, :, j], fd, rtol=5e-7, atol=5e-7)
    np.testing.assert_allclose(H, H.transpose(0, 1, 3, 2), rtol=1e-12, atol=1e-12)
