"""CUDA C++ sources of ready-made `JitModel`s: the reference demos' constitutive laws written once as
function templates, so that their derivatives come from dual numbers instead of hand derivation.  They double
as the parity cases of the generic path (tests/test_jit_*.py compare them with the hard-wired kernels and the
golden vectors)."""

from __future__ import annotations


# von Mises radial return, plane-strain Mandel 4-vectors: sigma(deps; sigma_n, p) and aux = dp.
# Statements follow doc/demo/demo_plasticity_von_mises.py:307-320; the tangent of :322-326 is NOT written
# down - d sigma / d deps by forward-mode AD reproduces it.  prm = [lambda, mu, H, sigma_0].
VON_MISES = r"""
template <class T>
__device__ void vm_sigma(const T* deps, const double* st, const double* prm, T* sig, T* aux) {
  const double l = prm[0], m = prm[1], H = prm[2], sig0 = prm[3];
  const double l2m = l + 2.0 * m, third = 1.0 / 3.0;
  const double p = st[4];
  T se[4];                                            // sigma_n + C_elas @ deps        (:308)
  se[0] = st[0] + (l2m * deps[0] + l * deps[1] + l * deps[2]);
  se[1] = st[1] + (l * deps[0] + l2m * deps[1] + l * deps[2]);
  se[2] = st[2] + (l * deps[0] + l * deps[1] + l2m * deps[2]);
  se[3] = st[3] + 2.0 * m * deps[3];
  const T tr = third * (se[0] + se[1] + se[2]);
  T s[4] = {se[0] - tr, se[1] - tr, se[2] - tr, se[3]};   // deviatoric @ sigma_elastic   (:309)
  const T seq = sqrt(1.5 * (s[0] * s[0] + s[1] * s[1] + s[2] * s[2] + s[3] * s[3]));   // (:310)
  const T f = seq - sig0 - H * p;                      // (:312)
  if (f > 0.0) {
    const T dp = f / (3.0 * m + H);                    // (:315)
    const T beta = 3.0 * m * dp / seq;                 // (:318)
    for (int i = 0; i < 4; ++i) sig[i] = se[i] - beta * s[i];   // (:320)
    aux[0] = dp;
  } else {
    for (int i = 0; i < 4; ++i) sig[i] = se[i];
    aux[0] = T(0.0);
  }
}
"""

# EXTENSION (not in the reference, whose demos are plane strain): the same radial return for full 3-D stress states,
# 6-component Mandel vectors [xx, yy, zz, sqrt2 yz, sqrt2 xz, sqrt2 xy]; tangent 6x6 by dual numbers.
VON_MISES_3D = r"""
template <class T>
__device__ void vm3d_sigma(const T* deps, const double* st, const double* prm, T* sig, T* aux) {
  const double l = prm[0], m = prm[1], H = prm[2], sig0 = prm[3];
  const double third = 1.0 / 3.0;
  const double p = st[6];
  const T trd = deps[0] + deps[1] + deps[2];
  T se[6];
  for (int i = 0; i < 3; ++i) se[i] = st[i] + (l * trd + 2.0 * m * deps[i]);
  for (int i = 3; i < 6; ++i) se[i] = st[i] + 2.0 * m * deps[i];
  const T tr = third * (se[0] + se[1] + se[2]);
  T s[6] = {se[0] - tr, se[1] - tr, se[2] - tr, se[3], se[4], se[5]};
  T ss = s[0] * s[0];
  for (int i = 1; i < 6; ++i) ss = ss + s[i] * s[i];
  const T seq = sqrt(1.5 * ss);
  const T f = seq - sig0 - H * p;
  if (f > 0.0) {
    const T dp = f / (3.0 * m + H);
    const T beta = 3.0 * m * dp / seq;
    for (int i = 0; i < 6; ++i) sig[i] = se[i] - beta * s[i];
    aux[0] = dp;
  } else {
    for (int i = 0; i < 6; ++i) sig[i] = se[i];
    aux[0] = T(0.0);
  }
}
"""

# Compressible neo-Hookean energy W(F) = mu/2 (tr(F^T F) - 3) - mu ln J + lambda/2 (ln J)^2, F row-major 3x3.
# Only the ENERGY is written down: the first Piola-Kirchhoff stress P = dW/dF is the (1,) derivative and the
# tangent dP/dF = d2W/dF2 the (2,) derivative (nested dual numbers) - the "torch.func.grad + jacfwd" recipe of
# doc/demo/demo_hyperelasticity.py:429-456 for a closed-form energy.  prm = [mu, lambda].  Arrays of 9 and 81 doubles
# per point have odd component counts: this model runs through the staged (TMA bulk copy) variant.
NEO_HOOKEAN_3D = r"""
template <class T>
__device__ void neo_hookean_W(const T* F, const double*, const double* prm, T* W, T*) {
  const double mu = prm[0], lam = prm[1];
  T I1 = F[0] * F[0];
  for (int i = 1; i < 9; ++i) I1 = I1 + F[i] * F[i];
  const T J = F[0] * (F[4] * F[8] - F[5] * F[7]) - F[1] * (F[3] * F[8] - F[5] * F[6]) + F[2] * (F[3] * F[7] - F[4] * F[6]);
  const T lnJ = log(J);
  W[0] = 0.5 * mu * (I1 - 3.0) - mu * lnJ + 0.5 * lam * lnJ * lnJ;
}
"""

# nonlinear heat flux q(T, sigma) = -k(T) sigma with k = 1 / (A + B T)
# (doc/demo/demo_nonlinear_heat_equation_part2.py:215-261; derivatives :228-261 come from AD here).
HEAT_FLUX = r"""
template <class T>
__device__ void heat_q(const T* x, const double*, const double* prm, T* q, T*) {
  const T k = 1.0 / (prm[0] + prm[1] * x[0]);
  q[0] = -k * x[1];
  q[1] = -k * x[2];
}
"""

# conductivity k(T) = 1 / (A + B T)  (doc/demo/demo_nonlinear_heat_equation_part1.py:252-272)
HEAT_K = r"""
template <class T>
__device__ void heat_k(const T* x, const double*, const double* prm, T* k, T*) {
  k[0] = 1.0 / (prm[0] + prm[1] * x[0]);
}
"""


def von_mises(E=70e3, nu=0.3, E_tangent=None, sigma_0=250.0, **kw):
    """`JitModel` equivalent of `constitutive.VonMises`: `(1,)` -> (C_tang, sigma, dp) like demo_vm:352,
    `(0,)` -> (sigma, dp).  State 0 = sigma_n [qp][4], state 1 = p [qp]."""
    from .jit import JitModel

    E_tangent = E / 100.0 if E_tangent is None else E_tangent
    H = E * E_tangent / (E - E_tangent)
    lmbda = E * nu / (1.0 + nu) / (1.0 - 2.0 * nu)
    mu = E / 2.0 / (1.0 + nu)
    kw.setdefault("returns", ("out", "value", "aux0"))
    return JitModel(VON_MISES, "vm_sigma", [(4,)], (4,), state_shapes=[(4,), ()], aux_shapes=[()],
                    params=[lmbda, mu, H, sigma_0], **kw)


def von_mises_3d(E=70e3, nu=0.3, E_tangent=None, sigma_0=250.0, **kw):
    """EXTENSION: von Mises for 6-component Mandel vectors; `(1,)` -> (C_tang [qp][6][6], sigma [qp][6], dp [qp])."""
    from .jit import JitModel

    E_tangent = E / 100.0 if E_tangent is None else E_tangent
    H = E * E_tangent / (E - E_tangent)
    lmbda = E * nu / (1.0 + nu) / (1.0 - 2.0 * nu)
    mu = E / 2.0 / (1.0 + nu)
    kw.setdefault("returns", ("out", "value", "aux0"))
    return JitModel(VON_MISES_3D, "vm3d_sigma", [(6,)], (6,), state_shapes=[(6,), ()], aux_shapes=[()],
                    params=[lmbda, mu, H, sigma_0], **kw)


def neo_hookean_3d(mu=1.0, lmbda=2.0, **kw):
    """Hyperelastic model from its energy alone: `(0,)` -> W, `(1,)` -> P = dW/dF [qp][9], `(2,)` -> dP/dF [qp][9][9]."""
    from .jit import JitModel

    return JitModel(NEO_HOOKEAN_3D, "neo_hookean_W", [(3, 3)], (), params=[mu, lmbda], **kw)


def heat_flux(A=1.0, B=1.0, **kw):
    """`JitModel` equivalent of `constitutive.HeatFlux`: operands (T, sigma) -> q, with (1,0), (0,1) and
    the second derivatives from dual numbers."""
    from .jit import JitModel

    return JitModel(HEAT_FLUX, "heat_q", [(), (2,)], (2,), params=[A, B], **kw)


def heat_conductivity(A=1.0, B=1.0, **kw):
    from .jit import JitModel

    return JitModel(HEAT_K, "heat_k", [()], (), params=[A, B], **kw)
