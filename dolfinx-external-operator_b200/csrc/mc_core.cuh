// mc_core.cuh - per-quadrature-point arithmetic of the Mohr-Coulomb return mapping with apex
// smoothing (reference: doc/demo/demo_plasticity_mohr_coulomb.py:282-555; citations ":NNN" below are
// to that file).  Host/device header: the sm_100a kernel in mc.cu is built from it, and the CPU test
// harness (tests/hostcheck/) compiles the very same source with g++ so that the algebra can be checked
// against the oracle without a GPU.  The harness is test-only; the product has no CPU path.
//
// What the reference computes, and how it is restated here
// --------------------------------------------------------
// The reference obtains dg/dsigma, dr/dy and the consistent tangent by three nested levels of JAX
// forward-mode AD, the outermost one THROUGH the Newton `lax.while_loop` (:555).  Written out, the
// outer level carries Y_k = d y_k / d deps (5x4, Y_0 = 0) through every update
//     y_{k+1} = y_k + delta_k,            J_k delta_k = -r_k                            (:511-513)
//     Y_{k+1} = Y_k + d(delta_k) = J_k^{-1} ( [C;0] - (D J_k[Y_k]) delta_k )
// (the total derivative of r_k is J_k Y_k + dr/d(deps) = J_k Y_k - [C;0]; the J_k Y_k part cancels
// against Y_k).  D J_k[Y] is the directional derivative of the local Jacobian along the columns of
// Y and needs THIRD derivatives of the plastic potential g, contracted with two vectors.  It vanishes
// at convergence (delta -> 0), where Y becomes the implicit-function tangent, but the loop stops at a
// relative residual of 1e-8 (:469) and the remaining term is what separates the two tangents by up
// to ~5e-9 - so it is kept, and parity with the AD tangent holds to <1e-10.
//
// The surface is a function of three invariants, h = I1/3 sin(a) + G(J2, J3) - c cos(a)   (:364-374),
// so all derivatives of h follow from the partials of the bivariate function G up to order 3 (a
// 10-coefficient Taylor jet propagated through  arg -> asin -> K(theta) -> sqrt), combined with the
// polynomial derivatives of J2 and J3 (grad J2 = s, grad J3 = dev m(s), Hess J3 = dev M(s) dev,
// third derivative of J3 constant).  This replaces 150-component nested dual numbers per scalar
// by ~10 and is what makes a register-resident per-thread Newton solve possible.
#pragma once

#include <cmath>
#include <cstdint>

#if defined(__CUDACC__)
#define EO_HD __host__ __device__ __forceinline__
#else
#define EO_HD inline
#endif

// ------------------------------------------------------------------------------------------------
// constants (built on the host by mc_make_consts, passed to the kernel by value)
// ------------------------------------------------------------------------------------------------
struct mc_angle {
  double sa;        // sin(angle)
  double kappa;     // sin(angle)/sqrt(3)                                   (:336,303,307)
  double A2;        // (a_g(angle) sin(angle))^2, a_g = a tan(phi)/tan(angle)  (:348-349,371)
  double ccos;      // c cos(angle)                                         (:373)
  double Ar[2], Br[2], Cr[2];  // Abbo-Sloan rounding coefficients for sign(theta) = +1 / -1  (:313-331)
};

struct mc_consts {
  double lam, mu2, l2m;  // lmbda, 2 mu, lmbda + 2 mu                        (:405-415)
  double s_d, s_tr;      // S_elas = C_elas^{-1}: S v = s_d v - s_tr tr(v) [1,1,1,0]  (:416)
  double theta_T;        //                                                  (:115)
  double x_T;            // sin(3 theta_T): |theta| > theta_T  <=>  |sin 3 theta| > x_T   (:334, theta_T < 30 deg)
  double tol;            //                                                  (:469)
  int32_t nitermax;      //                                                  (:469)
  int32_t assoc;         // phi == psi: f and g coincide, evaluate once
  mc_angle f, g;         // yield function (phi) and plastic potential (psi) (:383-388)
};

struct mc_params_in {  // mirrors eo_mc_params of include/eo_b200.h
  double E, nu, c, phi, psi, theta_T, a, tol;
  int32_t nitermax;
};

inline void mc_make_angle(const mc_params_in& p, double angle, mc_angle& o) {
  const double tT = p.theta_T;
  const double isq3 = 1.0 / std::sqrt(3.0);
  o.sa = std::sin(angle);
  o.kappa = isq3 * o.sa;
  const double ag = p.a * std::tan(p.phi) / std::tan(angle);
  o.A2 = ag * ag * o.sa * o.sa;
  o.ccos = p.c * std::cos(angle);
  for (int k = 0; k < 2; ++k) {
    const double sg = k == 0 ? 1.0 : -1.0;
    const double c1 = std::cos(tT) - isq3 * o.sa * std::sin(tT);                                      // :302-303
    const double c2 = sg * std::sin(tT) + isq3 * o.sa * std::cos(tT);                                 // :306-307
    const double c3 = 18.0 * std::cos(3.0 * tT) * std::cos(3.0 * tT) * std::cos(3.0 * tT);            // :310
    o.Cr[k] = (-std::cos(3.0 * tT) * c1 - 3.0 * sg * std::sin(3.0 * tT) * c2) / c3;                   // :313-316
    o.Br[k] = (sg * std::sin(6.0 * tT) * c1 - 6.0 * std::cos(6.0 * tT) * c2) / c3;                    // :319-322
    o.Ar[k] = -isq3 * o.sa * sg * std::sin(tT) - o.Br[k] * sg * std::sin(3 * tT) -
              o.Cr[k] * std::sin(3.0 * tT) * std::sin(3.0 * tT) + std::cos(tT);                       // :325-331
  }
}

inline void mc_make_consts(const mc_params_in& p, mc_consts& k) {
  const double lmbda = p.E * p.nu / ((1.0 + p.nu) * (1.0 - 2.0 * p.nu));  // :405
  const double mu = p.E / (2.0 * (1.0 + p.nu));                           // :406
  k.lam = lmbda;
  k.mu2 = 2 * mu;
  k.l2m = lmbda + 2 * mu;
  k.s_d = 1.0 / (2 * mu);
  k.s_tr = lmbda / (2 * mu * (3 * lmbda + 2 * mu));
  k.theta_T = p.theta_T;
  k.x_T = std::sin(3.0 * p.theta_T);
  k.tol = p.tol;
  k.nitermax = p.nitermax;
  k.assoc = (p.phi == p.psi) ? 1 : 0;
  mc_make_angle(p, p.phi, k.f);
  mc_make_angle(p, p.psi, k.g);
}

// ------------------------------------------------------------------------------------------------
// small fixed-size helpers
// ------------------------------------------------------------------------------------------------
EO_HD void mc_sincos(double x, double& s, double& c) {
#if defined(__CUDA_ARCH__)
  sincos(x, &s, &c);
#else
  s = std::sin(x);
  c = std::cos(x);
#endif
}

EO_HD double mc_rsqrt(double x) {
#if defined(__CUDA_ARCH__)
  return rsqrt(x);
#else
  return 1.0 / std::sqrt(x);
#endif
}

// sin(theta) for theta = asin(x) / 3 (:294) WITHOUT asin / sin in double precision: sin(theta) is the root of
// 4 s^3 - 3 s + x = 0 (sin 3 theta = 3 s - 4 s^3) in [-1/2, 1/2].  Single-precision start (FP32 pipe), one Newton step and
// one chord step with the same reciprocal: error 1e-7 -> 1e-13 -> below rounding; |h'| = 3 cos(3 theta) / cos(theta) >=
// 0.69 on the unrounded range |theta| <= theta_T = 26 deg.  Agrees with sin(asin(x) / 3) to 1 ulp over that range
// (tests/test_hostcheck_cpu.py).  NaN propagates.
EO_HD double mc_sin_third_asin(double x) {
  const float t0 = asinf((float)x) * (1.0f / 3.0f);
  const double s0 = (double)sinf(t0);
  const double r = 1.0 / (12.0 * s0 * s0 - 3.0);
  const double s1 = s0 - ((4.0 * s0 * s0 - 3.0) * s0 + x) * r;
  return s1 - ((4.0 * s1 * s1 - 3.0) * s1 + x) * r;
}

// C_elas @ v                                                                    (:407-415)
EO_HD void mc_Cmul(const mc_consts& k, const double v[4], double out[4]) {
  const double tr = k.lam * (v[0] + v[1] + v[2]);
  out[0] = tr + k.mu2 * v[0];
  out[1] = tr + k.mu2 * v[1];
  out[2] = tr + k.mu2 * v[2];
  out[3] = k.mu2 * v[3];
}

// dev @ v                                                                       (:352-360)
EO_HD void mc_dev(const double v[4], double out[4]) {
  const double m = (v[0] + v[1] + v[2]) * (1.0 / 3.0);
  out[0] = v[0] - m;
  out[1] = v[1] - m;
  out[2] = v[2] - m;
  out[3] = v[3];
}

// M(a) b with M(s) = d m / d s, m = dJ3/ds = [s2 s1, s2 s0, s0 s1 - s3^2/2, -s2 s3]   (from :282-283)
EO_HD void mc_Mmul(const double a[4], const double b[4], double out[4]) {
  out[0] = a[2] * b[1] + a[1] * b[2];
  out[1] = a[2] * b[0] + a[0] * b[2];
  out[2] = a[1] * b[0] + a[0] * b[1] - a[3] * b[3];
  out[3] = -a[3] * b[2] - a[2] * b[3];
}

EO_HD double mc_dot4(const double a[4], const double b[4]) {
  return a[0] * b[0] + a[1] * b[1] + a[2] * b[2] + a[3] * b[3];
}

// ------------------------------------------------------------------------------------------------
// surface h(sigma, angle) and the partials of G(J2, J3) = sqrt(J2 K(theta)^2 + A2)
// ------------------------------------------------------------------------------------------------
// 3rd-order bivariate Taylor jet in (x = J2, y = J3)
struct mc_jet {
  double v, x, y, xx, xy, yy, xxx, xxy, xyy, yyy;
};

// f(u) for a univariate f with derivatives f0..f3 at u.v
template <int ORD>
EO_HD void mc_compose(const mc_jet& u, double f0, double f1, double f2, double f3, mc_jet& o) {
  o.v = f0;
  o.x = f1 * u.x;
  o.y = f1 * u.y;
  if (ORD >= 2) {
    o.xx = f2 * u.x * u.x + f1 * u.xx;
    o.xy = f2 * u.x * u.y + f1 * u.xy;
    o.yy = f2 * u.y * u.y + f1 * u.yy;
  }
  if (ORD >= 3) {
    o.xxx = f3 * u.x * u.x * u.x + 3.0 * f2 * u.x * u.xx + f1 * u.xxx;
    o.xxy = f3 * u.x * u.x * u.y + f2 * (2.0 * u.x * u.xy + u.y * u.xx) + f1 * u.xxy;
    o.xyy = f3 * u.x * u.y * u.y + f2 * (2.0 * u.y * u.xy + u.x * u.yy) + f1 * u.xyy;
    o.yyy = f3 * u.y * u.y * u.y + 3.0 * f2 * u.y * u.yy + f1 * u.yyy;
  }
}

struct mc_surf {
  double s[4];  // dev sigma
  double t[4];  // grad J3 = dev m(s)
  double I1, J2, J3;
  double h;     // surface value
  mc_jet G;     // partials of G wrt (J2, J3), filled up to the requested order
};

// Evaluate the surface at `sig` for angle constants `ac` with G-partials up to order ORD (0..3).
// Statement order of the value follows :282-295, :334-345, :364-374.
template <int ORD>
EO_HD void mc_surface(const mc_consts& k, const mc_angle& ac, const double sig[4], mc_surf& o) {
  mc_dev(sig, o.s);
  const double s0 = o.s[0], s1 = o.s[1], s2 = o.s[2], s3 = o.s[3];
  o.I1 = sig[0] + sig[1] + sig[2];
  const double J2 = 0.5 * (s0 * s0 + s1 * s1 + s2 * s2 + s3 * s3);  // :286-287
  const double J3 = s2 * (s0 * s1 - s3 * s3 / 2.0);                 // :282-283
  o.J2 = J2;
  o.J3 = J3;
  if (ORD >= 1) {
    const double m[4] = {s2 * s1, s2 * s0, s0 * s1 - 0.5 * s3 * s3, -s2 * s3};
    mc_dev(m, o.t);
  }
  // arg = -(3 sqrt3 J3) / (2 sqrt(J2^3)), clipped to [-1, 1]                     (:292-293); one reciprocal square
  // root serves 1 / J2 and J2^(-3/2)
  const double c33 = 5.196152422706632;  // 3 sqrt(3)
  const double rJ = mc_rsqrt(J2);
  const double iJ2 = rJ * rJ;
  const double hh = -0.5 * c33 * (iJ2 * rJ);
  mc_jet a;
  a.v = hh * J3;
  a.x = -1.5 * a.v * iJ2;
  a.y = hh;
  a.xx = 3.75 * a.v * iJ2 * iJ2;
  a.xy = -1.5 * hh * iJ2;
  a.yy = 0.0;
  a.xxx = -13.125 * a.v * iJ2 * iJ2 * iJ2;
  a.xxy = 3.75 * hh * iJ2 * iJ2;
  a.xyy = 0.0;
  a.yyy = 0.0;
  if (a.v < -1.0 || a.v > 1.0) {  // jnp.clip: value pinned, tangent zero outside the interval
    a.v = a.v < 0.0 ? -1.0 : 1.0;
    a.x = a.y = a.xx = a.xy = a.xxx = a.xxy = 0.0;
  }
  // theta = asin(arg) / 3 (:294) enters only through sin / cos of theta (unrounded) or of 3 theta (rounded), and
  // sin 3 theta = arg: neither asin nor sincos is evaluated in double precision
  const double w2 = 1.0 - a.v * a.v;  // cos^2(3 theta)
  // K(theta) and its derivatives wrt theta                                      (:334-345)
  double K0, K1 = 0, K2 = 0, K3 = 0;
  {
    const bool rounded = fabs(a.v) > k.x_T;  // |theta| > theta_T
    if (rounded) {
      const double S = a.v;
      const int sgi = a.v < 0.0 ? 1 : 0;  // sign(0) = +1 (:298-299)
      const double Ac = ac.Ar[sgi], Bc = ac.Br[sgi], Cc = ac.Cr[sgi];
      K0 = Ac + Bc * S + Cc * S * S;  // :338-343
      if (ORD >= 1) {
        const double Cs = sqrt(w2);
        const double S1 = 3.0 * Cs, S2 = -9.0 * S, S3 = -27.0 * Cs;
        const double b2 = Bc + 2.0 * Cc * S;
        K1 = b2 * S1;
        K2 = 2.0 * Cc * S1 * S1 + b2 * S2;
        K3 = 6.0 * Cc * S1 * S2 + b2 * S3;
      }
    } else {
      const double S = mc_sin_third_asin(a.v);
      const double Cs = sqrt(1.0 - S * S);
      K0 = Cs - ac.kappa * S;  // :335-336
      K1 = -S - ac.kappa * Cs;
      K2 = -K0;
      K3 = -K1;
    }
  }
  const double Q0 = J2 * K0 * K0 + ac.A2;
  const double G0 = sqrt(Q0);
  o.G.v = G0;
  o.h = (o.I1 / 3.0 * ac.sa) + G0 - ac.ccos;  // :368-374
  if (ORD >= 1) {
    // theta as a function of arg, then K as a function of arg
    const double iw = mc_rsqrt(w2);
    const double t1 = (1.0 / 3.0) * iw;
    double k1 = K1 * t1, k2 = 0, k3 = 0;
    if (ORD >= 2) {
      const double iw2 = iw * iw;
      const double t2 = t1 * a.v * iw2;
      k2 = K2 * t1 * t1 + K1 * t2;
      if (ORD >= 3) {
        const double t3 = t1 * (1.0 + 2.0 * a.v * a.v) * iw2 * iw2;
        k3 = K3 * t1 * t1 * t1 + 3.0 * K2 * t1 * t2 + K1 * t3;
      }
    }
    mc_jet Kj;
    mc_compose<ORD>(a, K0, k1, k2, k3, Kj);
    // Q = x K^2 + A2
    mc_jet Q;
    const double M0 = K0 * K0;
    const double Mx = 2.0 * K0 * Kj.x, My = 2.0 * K0 * Kj.y;
    Q.v = Q0;
    Q.x = M0 + J2 * Mx;
    Q.y = J2 * My;
    double Mxx = 0, Mxy = 0, Myy = 0;
    if (ORD >= 2) {
      Mxx = 2.0 * (Kj.x * Kj.x + K0 * Kj.xx);
      Mxy = 2.0 * (Kj.x * Kj.y + K0 * Kj.xy);
      Myy = 2.0 * (Kj.y * Kj.y + K0 * Kj.yy);
      Q.xx = 2.0 * Mx + J2 * Mxx;
      Q.xy = My + J2 * Mxy;
      Q.yy = J2 * Myy;
    }
    if (ORD >= 3) {
      const double Mxxx = 2.0 * (3.0 * Kj.x * Kj.xx + K0 * Kj.xxx);
      const double Mxxy = 2.0 * (2.0 * Kj.x * Kj.xy + Kj.y * Kj.xx + K0 * Kj.xxy);
      const double Mxyy = 2.0 * (2.0 * Kj.y * Kj.xy + Kj.x * Kj.yy + K0 * Kj.xyy);
      const double Myyy = 2.0 * (3.0 * Kj.y * Kj.yy + K0 * Kj.yyy);
      Q.xxx = 3.0 * Mxx + J2 * Mxxx;
      Q.xxy = 2.0 * Mxy + J2 * Mxxy;
      Q.xyy = Myy + J2 * Mxyy;
      Q.yyy = J2 * Myyy;
    }
    const double iG = 1.0 / G0;
    const double g1 = 0.5 * iG, g2 = -0.25 * iG * iG * iG, g3 = 0.375 * iG * iG * iG * iG * iG;
    mc_compose<ORD>(Q, G0, g1, g2, g3, o.G);
  }
}

// grad h = sa/3 tr + G_x s + G_y t
EO_HD void mc_grad(const mc_angle& ac, const mc_surf& u, double n[4]) {
  const double h0 = ac.sa * (1.0 / 3.0);
  n[0] = h0 + u.G.x * u.s[0] + u.G.y * u.t[0];
  n[1] = h0 + u.G.x * u.s[1] + u.G.y * u.t[1];
  n[2] = h0 + u.G.x * u.s[2] + u.G.y * u.t[2];
  n[3] = u.G.x * u.s[3] + u.G.y * u.t[3];
}

// (Hess h) a = (G_xx J2_a + G_xy J3_a) s + G_x dev a + (G_xy J2_a + G_yy J3_a) t + G_y N a,  N = dev M(s) dev
EO_HD void mc_hess_mul(const mc_surf& u, const double a[4], double out[4]) {
  double da[4], Ma[4], Na[4];
  mc_dev(a, da);
  mc_Mmul(u.s, da, Ma);
  mc_dev(Ma, Na);
  const double J2a = mc_dot4(u.s, a), J3a = mc_dot4(u.t, a);
  const double al = u.G.xx * J2a + u.G.xy * J3a, be = u.G.xy * J2a + u.G.yy * J3a;
#pragma unroll
  for (int i = 0; i < 4; ++i) out[i] = al * u.s[i] + u.G.x * da[i] + be * u.t[i] + u.G.y * Na[i];
}

// third derivative of h contracted with a and b (symmetric in a, b)
EO_HD void mc_third_mul(const mc_surf& u, const double a[4], const double b[4], double out[4]) {
  double da[4], db[4], Ma[4], Mb[4], Na[4], Nb[4], Mab[4], Nab[4];
  mc_dev(a, da);
  mc_dev(b, db);
  mc_Mmul(u.s, da, Ma);
  mc_Mmul(u.s, db, Mb);
  mc_dev(Ma, Na);
  mc_dev(Mb, Nb);
  mc_Mmul(da, db, Mab);
  mc_dev(Mab, Nab);
  const mc_jet& G = u.G;
  const double J2a = mc_dot4(u.s, a), J3a = mc_dot4(u.t, a), J2b = mc_dot4(u.s, b), J3b = mc_dot4(u.t, b);
  const double J2ab = mc_dot4(a, db), J3ab = mc_dot4(a, Nb);
  const double cs = G.xxx * J2a * J2b + G.xxy * (J2a * J3b + J3a * J2b) + G.xyy * J3a * J3b + G.xx * J2ab + G.xy * J3ab;
  const double ct = G.xxy * J2a * J2b + G.xyy * (J2a * J3b + J3a * J2b) + G.yyy * J3a * J3b + G.xy * J2ab + G.yy * J3ab;
  const double ala = G.xx * J2a + G.xy * J3a, alb = G.xx * J2b + G.xy * J3b;
  const double bea = G.xy * J2a + G.yy * J3a, beb = G.xy * J2b + G.yy * J3b;
#pragma unroll
  for (int i = 0; i < 4; ++i)
    out[i] = cs * u.s[i] + ala * db[i] + alb * da[i] + ct * u.t[i] + bea * Nb[i] + beb * Na[i] + G.y * Nab[i];
}

// ------------------------------------------------------------------------------------------------
// linear algebra of one Newton update
// ------------------------------------------------------------------------------------------------
// The reference solves the 5x5 system  J [dsig; dlam] = rhs,  J = [[I + dl C Hg, C n], [df^T, 0]]  with
// LAPACK's pivoted LU (jnp.linalg.solve, :512).  Multiplying the first block row by S = C^{-1} turns it into
// the bordered SYMMETRIC system
//     M dsig + n dlam = S rhs_sig,      df . dsig = rhs_lam,        M = S + dl Hg   (4x4, SPD: g is convex),
// which is solved by one LDL^T factorisation of M and a scalar Schur complement:
//     z = M^{-1} n,  q = M^{-1} df,  dlam = (q . w - rhs_lam) / (df . z),  dsig = M^{-1} w - z dlam,  w = S rhs_sig.
// Same solution as the reference's LU up to rounding (differences ~ eps * cond, far below the 1e-10 parity
// tolerance), but ~100 flops and no pivot bookkeeping instead of ~600 instructions - the kernel is
// instruction-fetch sensitive - and the tangent right-hand sides [C e_j; 0] become plain unit vectors.

// symmetric 4x4 stored as 10: index of (i,j), i <= j
#define MC_SYM(i, j) ((i) <= (j) ? ((i) * (7 - (i)) / 2 + (j)) : ((j) * (7 - (j)) / 2 + (i)))

struct mc_ldl {
  double l10, l20, l21, l30, l31, l32;  // unit lower factor
  double id0, id1, id2, id3;            // reciprocals of D
};

EO_HD void mc_ldl_factor(const double M[10], mc_ldl& f) {
  f.id0 = 1.0 / M[MC_SYM(0, 0)];
  f.l10 = M[MC_SYM(1, 0)] * f.id0;
  f.l20 = M[MC_SYM(2, 0)] * f.id0;
  f.l30 = M[MC_SYM(3, 0)] * f.id0;
  const double d1 = M[MC_SYM(1, 1)] - f.l10 * M[MC_SYM(1, 0)];
  f.id1 = 1.0 / d1;
  const double m21 = M[MC_SYM(2, 1)] - f.l20 * M[MC_SYM(1, 0)];
  const double m31 = M[MC_SYM(3, 1)] - f.l30 * M[MC_SYM(1, 0)];
  f.l21 = m21 * f.id1;
  f.l31 = m31 * f.id1;
  const double d2 = M[MC_SYM(2, 2)] - f.l20 * M[MC_SYM(2, 0)] - f.l21 * m21;
  f.id2 = 1.0 / d2;
  const double m32 = M[MC_SYM(3, 2)] - f.l30 * M[MC_SYM(2, 0)] - f.l31 * m21;
  f.l32 = m32 * f.id2;
  const double d3 = M[MC_SYM(3, 3)] - f.l30 * M[MC_SYM(3, 0)] - f.l31 * m31 - f.l32 * m32;
  f.id3 = 1.0 / d3;
}

EO_HD void mc_ldl_solve(const mc_ldl& f, double b[4]) {
  b[1] -= f.l10 * b[0];
  b[2] -= f.l20 * b[0] + f.l21 * b[1];
  b[3] -= f.l30 * b[0] + f.l31 * b[1] + f.l32 * b[2];
  b[0] *= f.id0, b[1] *= f.id1, b[2] *= f.id2, b[3] *= f.id3;
  b[2] -= f.l32 * b[3];
  b[1] -= f.l21 * b[2] + f.l31 * b[3];
  b[0] -= f.l10 * b[1] + f.l20 * b[2] + f.l30 * b[3];
}

// S_elas @ v = C_elas^{-1} v                                                      (:416)
EO_HD void mc_Smul(const mc_consts& k, const double v[4], double out[4]) {
  const double tr = k.s_tr * (v[0] + v[1] + v[2]);
  out[0] = k.s_d * v[0] - tr;
  out[1] = k.s_d * v[1] - tr;
  out[2] = k.s_d * v[2] - tr;
  out[3] = k.s_d * v[3];
}

// ------------------------------------------------------------------------------------------------
// per-point state of the local Newton solve ("slot") and the stage function
// ------------------------------------------------------------------------------------------------
// A plastic point lives in a slot of MC_NF doubles between stages.  The kernel keeps slots in shared
// memory as structure-of-arrays (field f of slot s at base[f * stride + s]); the host harness and the
// one-thread-per-point kernel use a private array (stride 1).
enum {
  MC_F_Y = 0,        // y = [sigma(4), dlambda]                                   (:496-498)
  MC_F_TRIAL = 5,    // sigma_n + C deps
  MC_F_NORM0 = 9,    // ||res0||                                                  (:501)
  MC_F_YIELD = 10,   // f(trial)                                                  (:529-530)
  MC_F_NRM = 11,     // ||res|| at the current iterate                            (:516)
  MC_F_G = 12,       // partials of G at y: x, y, xx, xy, yy, xxx, xxy, xyy, yyy
  MC_F_R = 21,       // residual at y (5)                                         (:515)
  MC_F_YY = 26,      // Y = dy/d deps, row-major 5x4, carried through the loop    (:555)
  MC_NF_ASSOC = 46,
  MC_F_F = 46,       // partials of F (yield surface) at y: x, y, xx, xy, yy  - non-associative only
  MC_NF = 51
};

struct mc_slot {
  double* base;
  int stride;
  EO_HD double& operator[](int f) const { return base[f * stride]; }
};

EO_HD double mc_norm5(const double r[5]) {
  return sqrt(r[0] * r[0] + r[1] * r[1] + r[2] * r[2] + r[3] * r[3] + r[4] * r[4]);
}

// trial stress sigma_n + C deps and the yield predicate (:421-422).  Returns f(trial).
EO_HD double mc_trial(const mc_consts& k, const double de[4], const double sn[4], double Cde[4]) {
  mc_Cmul(k, de, Cde);
  const double tr[4] = {sn[0] + Cde[0], sn[1] + Cde[1], sn[2] + Cde[2], sn[3] + Cde[3]};
  mc_surf u;
  mc_surface<0>(k, k.f, tr, u);
  return u.h;
}

// Elastic branch (:424-425, :442-443): r = [sigma - sigma_n - C deps, dlambda], J = I.
// Writes sigma, C_tang (row-major 4x4), and the aux outputs; returns niter.
EO_HD int32_t mc_elastic(const mc_consts& k, const double sn[4], const double Cde[4], double sig[4], double Ct[16],
                         double& norm_res, double& dlambda) {
  double y[5] = {sn[0], sn[1], sn[2], sn[3], 0.0};
  double r[5];
#pragma unroll
  for (int i = 0; i < 4; ++i) r[i] = (y[i] - sn[i]) - Cde[i];
  r[4] = y[4];
  const double norm0 = mc_norm5(r);
  double nrm = norm0;
  int32_t niter = 0;
  while ((nrm / norm0 > k.tol) && niter < k.nitermax) {  // :503-505 (NaN compares false)
#pragma unroll
    for (int i = 0; i < 5; ++i) y[i] = y[i] + (-r[i]);
#pragma unroll
    for (int i = 0; i < 4; ++i) r[i] = (y[i] - sn[i]) - Cde[i];
    r[4] = y[4];
    nrm = mc_norm5(r);
    niter += 1;
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) sig[i] = y[i];
  // Y_1 = -I^{-1} R_0 = C and stays there; zero iterations leave Y_0 = 0
  const double on = niter > 0 ? 1.0 : 0.0;
#pragma unroll
  for (int i = 0; i < 16; ++i) Ct[i] = 0.0;
  Ct[0] = Ct[5] = Ct[10] = on * k.l2m;
  Ct[1] = Ct[2] = Ct[4] = Ct[6] = Ct[8] = Ct[9] = on * k.lam;
  Ct[15] = on * k.mu2;
  norm_res = nrm;
  dlambda = y[4];
  return niter;
}

// Start a plastic point: y0 = [sigma_n, 0] (:496-498).  Y0 = 0 is implicit (the first update overwrites Y).
EO_HD void mc_slot_init(const mc_slot& sl, const double sn[4], const double Cde[4], double yielding) {
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    sl[MC_F_Y + i] = sn[i];
    sl[MC_F_TRIAL + i] = sn[i] + Cde[i];
  }
  sl[MC_F_Y + 4] = 0.0;
  sl[MC_F_YIELD] = yielding;
}

EO_HD void mc_symmul(const double H[10], const double v[4], double out[4]) {
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    double acc = 0.0;
#pragma unroll
    for (int j = 0; j < 4; ++j) acc += H[MC_SYM(i, j)] * v[j];
    out[i] = acc;
  }
}

// Hessian of G(J2(sigma), J3(sigma)) as a symmetric matrix:
//   Gx dev + Gy N + Gxx s s^T + Gxy (s t^T + t s^T) + Gyy t t^T,   N = dev M(s) dev
EO_HD void mc_hess(const double s[4], const double t[4], double Gx, double Gy, double Gxx, double Gxy, double Gyy,
                   double H[10]) {
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    double e[4] = {0, 0, 0, 0}, de[4], Me[4], Ne[4];
    e[j] = 1.0;
    mc_dev(e, de);
    mc_Mmul(s, de, Me);
    mc_dev(Me, Ne);
#pragma unroll
    for (int i = 0; i <= j; ++i)
      H[MC_SYM(i, j)] = Gx * de[i] + Gy * Ne[i] + Gxx * s[i] * s[j] + Gxy * (s[i] * t[j] + t[i] * s[j]) + Gyy * t[i] * t[j];
  }
}

// One stage of the plastic Newton iteration on a slot.
//   kind 0: first residual only                     r0 = r(y0), ||res0||                    (:500-501)
//   kind 1: first update (Y0 = 0, so no DJ term), then residual at the new iterate          (:511-516)
//   kind 2: update with the full tangent recursion, then residual at the new iterate
// after the residual the loop test (:503-505) is applied: returns true when the point has LEFT the loop
// (y, Y, ||res|| in the slot are final), false when another update is needed (the slot then holds the
// G-partials and the residual the next stage starts from).  `niter` is updated in place.
template <bool ASSOC>
EO_HD bool mc_stage(const mc_consts& k, int kind, const mc_slot& sl, int32_t& niter) {
  double y[5];
#pragma unroll
  for (int i = 0; i < 5; ++i) y[i] = sl[MC_F_Y + i];

  if (kind >= 1) {
    double s[4], t[4];
    mc_dev(y, s);
    {
      const double m[4] = {s[2] * s[1], s[2] * s[0], s[0] * s[1] - 0.5 * s[3] * s[3], -s[2] * s[3]};
      mc_dev(m, t);
    }
    const double Gx = sl[MC_F_G + 0], Gy = sl[MC_F_G + 1];
    double n[4], df[4];
    {
      const double h0 = k.g.sa * (1.0 / 3.0);
#pragma unroll
      for (int i = 0; i < 4; ++i) n[i] = (i < 3 ? h0 : 0.0) + Gx * s[i] + Gy * t[i];
    }
    if (ASSOC) {
#pragma unroll
      for (int i = 0; i < 4; ++i) df[i] = n[i];
    } else {
      const double Fx = sl[MC_F_F + 0], Fy = sl[MC_F_F + 1];
      const double h0 = k.f.sa * (1.0 / 3.0);
#pragma unroll
      for (int i = 0; i < 4; ++i) df[i] = (i < 3 ? h0 : 0.0) + Fx * s[i] + Fy * t[i];
    }
    double d[5];
    if (kind == 1) {
      // FIRST update: y0 = [sigma_n, 0], so dlambda = 0 and M = S + 0 Hess g = S EXACTLY (0 x finite = 0; where the
      // Hessian is not finite - J2 = 0, clipped Lode argument - the gradient n is not finite either and poisons the
      // update just like the reference's Jacobian does), M^{-1} = C_elas: no Hessian, no factorisation.  And Y0 = 0: the
      // tangent right-hand sides are the unit vectors.
      //   z = C n,  q = C df,  dlam = (df . rs + r_f) / (df . z),  dsig = rs - z dlam        (rs = -r_sigma, C S = I)
      //   Y_sig[:, j] = C e_j - z yl_j,  Y_lam[j] = yl_j = q_j / (df . z)
      double z[4], q_[4];
      mc_Cmul(k, n, z);
      if (ASSOC) {
#pragma unroll
        for (int i = 0; i < 4; ++i) q_[i] = z[i];
      } else {
        mc_Cmul(k, df, q_);
      }
      const double idenom = 1.0 / mc_dot4(df, z);
      const double rs[4] = {-sl[MC_F_R + 0], -sl[MC_F_R + 1], -sl[MC_F_R + 2], -sl[MC_F_R + 3]};
      d[4] = (mc_dot4(df, rs) + sl[MC_F_R + 4]) * idenom;
#pragma unroll
      for (int i = 0; i < 4; ++i) d[i] = rs[i] - z[i] * d[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const double yl = q_[j] * idenom;
        double e[4] = {0.0, 0.0, 0.0, 0.0}, Ce[4];
        e[j] = 1.0;
        mc_Cmul(k, e, Ce);
#pragma unroll
        for (int i = 0; i < 4; ++i) sl[MC_F_YY + 4 * i + j] = Ce[i] - z[i] * yl;
        sl[MC_F_YY + 16 + j] = yl;
      }
    } else {
    const double Gxx = sl[MC_F_G + 2], Gxy = sl[MC_F_G + 3], Gyy = sl[MC_F_G + 4];
    double H[10];
    mc_hess(s, t, Gx, Gy, Gxx, Gxy, Gyy, H);
    double Hf[10];
    if (!ASSOC) mc_hess(s, t, sl[MC_F_F + 0], sl[MC_F_F + 1], sl[MC_F_F + 2], sl[MC_F_F + 3], sl[MC_F_F + 4], Hf);
    // M = S + dl Hg (symmetric 4x4), z = M^{-1} n, q = M^{-1} df, denom = df . z     (see "linear algebra")
    const double dl = y[4];
    mc_ldl lf;
    {
      double M[10];
#pragma unroll
      for (int i = 0; i < 10; ++i) M[i] = dl * H[i];
      M[MC_SYM(0, 0)] += k.s_d - k.s_tr, M[MC_SYM(1, 1)] += k.s_d - k.s_tr, M[MC_SYM(2, 2)] += k.s_d - k.s_tr;
      M[MC_SYM(3, 3)] += k.s_d;
      M[MC_SYM(0, 1)] -= k.s_tr, M[MC_SYM(0, 2)] -= k.s_tr, M[MC_SYM(1, 2)] -= k.s_tr;
      mc_ldl_factor(M, lf);
    }
    double z[4], q_[4];  // z = M^{-1} n, q_ = M^{-1} df
#pragma unroll
    for (int i = 0; i < 4; ++i) z[i] = n[i];
    mc_ldl_solve(lf, z);
    if (ASSOC) {
#pragma unroll
      for (int i = 0; i < 4; ++i) q_[i] = z[i];
    } else {
#pragma unroll
      for (int i = 0; i < 4; ++i) q_[i] = df[i];
      mc_ldl_solve(lf, q_);
    }
    const double idenom = 1.0 / mc_dot4(df, z);
    // Newton step  J delta = -r                                                   (:511-512)
    {
      const double rs[4] = {-sl[MC_F_R + 0], -sl[MC_F_R + 1], -sl[MC_F_R + 2], -sl[MC_F_R + 3]};
      double w[4];
      mc_Smul(k, rs, w);
      d[4] = (mc_dot4(q_, w) + sl[MC_F_R + 4]) * idenom;
      mc_ldl_solve(lf, w);
#pragma unroll
      for (int i = 0; i < 4; ++i) d[i] = w[i] - z[i] * d[4];
    }

    // tangent recursion  Y[:,j] <- J^{-1} ( [C e_j; 0] - (DJ[Y[:,j]]) delta ), column by column:
    // in the symmetric form the right-hand side is  w = e_j - (al Hd + Nm a),  rho = -(Hf a) . delta_sig  with
    // a = Y_sig[:,j], al = Y_lam[j] and the SYMMETRIC matrix  Nm = dl D(Hess g)[delta_sig] + delta_lam Hess g:
    // the third derivative of g contracted with delta_sig is assembled ONCE per update (10 entries) instead of being
    // contracted with every column.  With b = delta_sig, db = dev b, Nb = N(s) b, J2b = s.b, J3b = t.b:
    //   D(Hess g)[b] = alb dev + beb N(s) + Gy N(db) + (s p^T + p s^T) + (t q^T + q t^T)
    //   p = c1/2 s + Gxx db + c2 t + Gxy Nb,   q = Gxy db + c3/2 t + Gyy Nb,
    //   c1 = Gxxx J2b + Gxxy J3b,  c2 = Gxxy J2b + Gxyy J3b,  c3 = Gxyy J2b + Gyyy J3b
    double Hd[4], Hfd[4], Nm[10];
    mc_symmul(H, d, Hd);
    if (ASSOC) {
#pragma unroll
      for (int i = 0; i < 4; ++i) Hfd[i] = Hd[i];
    } else {
      mc_symmul(Hf, d, Hfd);
    }
    {
      const double Gxxx = sl[MC_F_G + 5], Gxxy = sl[MC_F_G + 6], Gxyy = sl[MC_F_G + 7], Gyyy = sl[MC_F_G + 8];
      double db[4], Nb[4], p[4], q[4];
      mc_dev(d, db);
      {
        double Mb[4];
        mc_Mmul(s, db, Mb);
        mc_dev(Mb, Nb);
      }
      const double J2b = mc_dot4(s, d), J3b = mc_dot4(t, d);
      const double alb = Gxx * J2b + Gxy * J3b, beb = Gxy * J2b + Gyy * J3b;
      const double c1 = 0.5 * (Gxxx * J2b + Gxxy * J3b), c2 = Gxxy * J2b + Gxyy * J3b, c3 = 0.5 * (Gxyy * J2b + Gyyy * J3b);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        p[i] = c1 * s[i] + Gxx * db[i] + c2 * t[i] + Gxy * Nb[i];
        q[i] = Gxy * db[i] + c3 * t[i] + Gyy * Nb[i];
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        double e[4] = {0, 0, 0, 0}, de[4], Ms[4], Ns[4], Md[4], Nd[4];
        e[j] = 1.0;
        mc_dev(e, de);
        mc_Mmul(s, de, Ms);
        mc_dev(Ms, Ns);
        mc_Mmul(db, de, Md);
        mc_dev(Md, Nd);
#pragma unroll
        for (int i = 0; i <= j; ++i) {
          const double DH = alb * de[i] + beb * Ns[i] + Gy * Nd[i] + (s[i] * p[j] + p[i] * s[j]) + (t[i] * q[j] + q[i] * t[j]);
          Nm[MC_SYM(i, j)] = dl * DH + d[4] * H[MC_SYM(i, j)];
        }
      }
    }
#if defined(__CUDA_ARCH__)
#pragma unroll 1
#endif
    for (int j = 0; j < 4; ++j) {
      double w[4] = {0.0, 0.0, 0.0, 0.0};
      double rho = 0.0;
      {
        double a[4], Na[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) a[i] = sl[MC_F_YY + 4 * i + j];
        const double al = sl[MC_F_YY + 16 + j];
        mc_symmul(Nm, a, Na);
#pragma unroll
        for (int i = 0; i < 4; ++i) w[i] = -(al * Hd[i] + Na[i]);
        rho = -mc_dot4(a, Hfd);
      }
      w[0] += (j == 0 ? 1.0 : 0.0), w[1] += (j == 1 ? 1.0 : 0.0), w[2] += (j == 2 ? 1.0 : 0.0), w[3] += (j == 3 ? 1.0 : 0.0);
      const double yl = (mc_dot4(q_, w) - rho) * idenom;
      mc_ldl_solve(lf, w);
#pragma unroll
      for (int i = 0; i < 4; ++i) sl[MC_F_YY + 4 * i + j] = w[i] - z[i] * yl;
      sl[MC_F_YY + 16 + j] = yl;
    }
    }
#pragma unroll
    for (int i = 0; i < 5; ++i) {
      y[i] += d[i];  // :513
      sl[MC_F_Y + i] = y[i];
    }
    niter += 1;  // :519
  }

  // --- residual at the (new) iterate and the loop test
  // the update that follows the FIRST residual (kind 0 -> kind 1) needs the gradients only (dlambda = 0, Y0 = 0): first-
  // order partials suffice there; every later update needs the partials up to order 3 (g) / 2 (f)
  mc_surf ug;
  if (kind == 0) {
    mc_surface<1>(k, k.g, y, ug);
    ug.G.xx = ug.G.xy = ug.G.yy = ug.G.xxx = ug.G.xxy = ug.G.xyy = ug.G.yyy = 0.0;
  } else {
    mc_surface<3>(k, k.g, y, ug);
  }
  double n[4];
  mc_grad(k.g, ug, n);
  double fval;
  if (ASSOC) {
    fval = ug.h;
  } else {
    mc_surf uf;
    if (kind == 0) {
      mc_surface<1>(k, k.f, y, uf);
      uf.G.xx = uf.G.xy = uf.G.yy = 0.0;
    } else {
      mc_surface<2>(k, k.f, y, uf);
    }
    fval = uf.h;
    sl[MC_F_F + 0] = uf.G.x, sl[MC_F_F + 1] = uf.G.y, sl[MC_F_F + 2] = uf.G.xx, sl[MC_F_F + 3] = uf.G.xy,
                sl[MC_F_F + 4] = uf.G.yy;
  }
  double r[5];
  {
    double Cn[4];
    mc_Cmul(k, n, Cn);
#pragma unroll
    for (int i = 0; i < 4; ++i) r[i] = (y[i] - sl[MC_F_TRIAL + i]) + y[4] * Cn[i];  // :435
    r[4] = fval;                                                                  // :446
  }
  const double nrm = mc_norm5(r);
  sl[MC_F_NRM] = nrm;
  double norm0;
  if (kind == 0) {
    norm0 = nrm;
    sl[MC_F_NORM0] = nrm;
  } else {
    norm0 = sl[MC_F_NORM0];
  }
  // nrm / norm0 > tol (:505) without the division; identical on NaN / zero / inf operands
  if (!((nrm > k.tol * norm0) && (niter < k.nitermax))) return true;
  sl[MC_F_G + 0] = ug.G.x, sl[MC_F_G + 1] = ug.G.y, sl[MC_F_G + 2] = ug.G.xx, sl[MC_F_G + 3] = ug.G.xy,
              sl[MC_F_G + 4] = ug.G.yy, sl[MC_F_G + 5] = ug.G.xxx, sl[MC_F_G + 6] = ug.G.xxy, sl[MC_F_G + 7] = ug.G.xyy,
              sl[MC_F_G + 8] = ug.G.yyy;
#pragma unroll
  for (int i = 0; i < 5; ++i) sl[MC_F_R + i] = r[i];
  return false;
}

// outputs of a finished plastic point: C_tang = d sigma / d deps = first four rows of Y (zero when the
// loop never ran: Y0 = 0), sigma = y[:4], dlambda = y[4]
EO_HD void mc_slot_result(const mc_slot& sl, int32_t niter, double Ct[16], double sig[4], double& norm_res,
                          double& dlambda) {
#pragma unroll
  for (int i = 0; i < 16; ++i) Ct[i] = niter > 0 ? sl[MC_F_YY + i] : 0.0;
#pragma unroll
  for (int i = 0; i < 4; ++i) sig[i] = sl[MC_F_Y + i];
  dlambda = sl[MC_F_Y + 4];
  norm_res = sl[MC_F_NRM];
}

// Whole point, serial form (host harness and the simple one-thread-per-point kernel).
EO_HD void mc_point(const mc_consts& k, const double de[4], const double sn[4], double Ct[16], double sig[4],
                    int32_t& niter, double& yielding, double& norm_res, double& dlambda) {
  double Cde[4];
  yielding = mc_trial(k, de, sn, Cde);
  if (yielding <= 0.0) {  // :430,448
    niter = mc_elastic(k, sn, Cde, sig, Ct, norm_res, dlambda);
    return;
  }
  double store[MC_NF];
  const mc_slot sl{store, 1};
  mc_slot_init(sl, sn, Cde, yielding);
  niter = 0;
  int kind = 0;
  if (k.assoc) {
    while (!mc_stage<true>(k, kind, sl, niter)) kind = kind == 0 ? 1 : 2;
  } else {
    while (!mc_stage<false>(k, kind, sl, niter)) kind = kind == 0 ? 1 : 2;
  }
  mc_slot_result(sl, niter, Ct, sig, norm_res, dlambda);
}
