// vm_core.cuh - per-quadrature-point arithmetic of the von Mises radial return
// (reference: doc/demo/demo_plasticity_von_mises.py:307-326, statement by statement).  Shared by the
// streaming kernel (vm_heat.cu) and the kernel fused with the strain tabulation (tab.cu); both are
// compiled with -fmad=false so that the results are the plain IEEE sequence of the reference's statements.
#pragma once

// host/device header: the CUDA kernels and the CPU test harness (tests/hostcheck/) compile the same source
#if defined(__CUDACC__)
#define EO_VM_HD __device__ __forceinline__
#else
#include <cmath>
#define EO_VM_HD inline
#endif

struct vm_consts {
  double l, m, H, s0;
};

struct vm_point_out {
  double C[16];  // consistent tangent, row-major 4x4
  double g[4];   // new stress
  double dp;     // plastic multiplier increment
  // the tangent in FACTORED form (6 numbers): C_t = C_elas - cn v v^T - cd dev, with v = n_elas f+ / f (:317),
  // cn = 3 mu (3 mu / (3 mu + H) - beta) (:323), cd = 2 mu beta (:324).  C above is vm_tangent_*(v, cn, cd).
  double v[4], cn, cd;
};

// C_t from its factors, the reference's statement sequence (:323-326 with C_elas, dev written out)
EO_VM_HD void vm_tangent_exact(const vm_consts& q, const double v[4], double cn, double cd, double C[16]) {
  const double l = q.l, m = q.m;
  const double l2m = l + 2.0 * m;
  const double third = 1.0 / 3.0;
  const double v0 = v[0], v1 = v[1], v2 = v[2], v3 = v[3];
  const double Dd = 1.0 - third, Do = 0.0 - third;  // deviatoric diagonal / off-diagonal (3x3 block)
  C[0] = l2m - cn * (v0 * v0) - cd * Dd, C[1] = l - cn * (v0 * v1) - cd * Do, C[2] = l - cn * (v0 * v2) - cd * Do,
  C[3] = 0.0 - cn * (v0 * v3) - cd * 0.0;
  C[4] = l - cn * (v1 * v0) - cd * Do, C[5] = l2m - cn * (v1 * v1) - cd * Dd, C[6] = l - cn * (v1 * v2) - cd * Do,
  C[7] = 0.0 - cn * (v1 * v3) - cd * 0.0;
  C[8] = l - cn * (v2 * v0) - cd * Do, C[9] = l - cn * (v2 * v1) - cd * Do, C[10] = l2m - cn * (v2 * v2) - cd * Dd,
  C[11] = 0.0 - cn * (v2 * v3) - cd * 0.0;
  C[12] = 0.0 - cn * (v3 * v0) - cd * 0.0, C[13] = 0.0 - cn * (v3 * v1) - cd * 0.0,
  C[14] = 0.0 - cn * (v3 * v2) - cd * 0.0, C[15] = 2.0 * m - cn * (v3 * v3) - cd * 1.0;
}

// the same with the symmetry used and explicit FMAs (vm_point_fast)
EO_VM_HD void vm_tangent_fast(const vm_consts& q, const double v[4], double cn, double cd, double C[16]) {
  const double l = q.l, m = q.m;
  const double l2m = l + 2.0 * m;
  const double third = 1.0 / 3.0;
  const double tt = 1.0 - third;
  const double v0 = v[0], v1 = v[1], v2 = v[2], v3 = v[3];
  const double w0 = cn * v0, w1 = cn * v1, w2 = cn * v2, w3 = cn * v3;
  const double dD = l2m - cd * tt, dO = l + cd * third;  // C - cd * dev on the 3x3 block
  const double c00 = fma(-w0, v0, dD), c01 = fma(-w0, v1, dO), c02 = fma(-w0, v2, dO), c03 = -(w0 * v3);
  const double c11 = fma(-w1, v1, dD), c12 = fma(-w1, v2, dO), c13 = -(w1 * v3);
  const double c22 = fma(-w2, v2, dD), c23 = -(w2 * v3);
  const double c33 = fma(-w3, v3, 2.0 * m - cd);
  C[0] = c00, C[1] = c01, C[2] = c02, C[3] = c03;
  C[4] = c01, C[5] = c11, C[6] = c12, C[7] = c13;
  C[8] = c02, C[9] = c12, C[10] = c22, C[11] = c23;
  C[12] = c03, C[13] = c13, C[14] = c23, C[15] = c33;
}

// tau = C_t e straight from the factors (device-side consumers: 48 instead of 128 bytes per point)
EO_VM_HD void vm_factored_apply(const vm_consts& q, const double v[4], double cn, double cd,
                                                  const double e[4], double tau[4]) {
  const double l = q.l, m = q.m;
  const double dD = (l + 2.0 * m) - cd * (2.0 / 3.0), dO = l + cd * (1.0 / 3.0);
  const double ve = cn * fma(v[3], e[3], fma(v[2], e[2], fma(v[1], e[1], v[0] * e[0])));
  tau[0] = fma(-ve, v[0], fma(dD, e[0], dO * (e[1] + e[2])));
  tau[1] = fma(-ve, v[1], fma(dD, e[1], dO * (e[0] + e[2])));
  tau[2] = fma(-ve, v[2], fma(dD, e[2], dO * (e[0] + e[1])));
  tau[3] = fma(-ve, v[3], (2.0 * m - cd) * e[3]);
}

EO_VM_HD void vm_point(const vm_consts& q, double e0, double e1, double e2, double e3, double n0,
                                         double n1, double n2, double n3, double pi, vm_point_out& o) {
  const double l = q.l, m = q.m, H = q.H;
  const double l2m = l + 2.0 * m;
  // sigma_elastic = sigma_n + C_elas @ deps                                   (:308)
  const double se0 = n0 + (l2m * e0 + l * e1 + l * e2);
  const double se1 = n1 + (l * e0 + l2m * e1 + l * e2);
  const double se2 = n2 + (l * e0 + l * e1 + l2m * e2);
  const double se3 = n3 + 2.0 * m * e3;
  // s = deviatoric @ sigma_elastic                                             (:309)
  const double third = 1.0 / 3.0;
  const double tt = 1.0 - third;
  const double s0 = tt * se0 - third * se1 - third * se2;
  const double s1 = -third * se0 + tt * se1 - third * se2;
  const double s2 = -third * se0 - third * se1 + tt * se2;
  const double s3 = se3;
  const double seq = sqrt(3.0 / 2.0 * (s0 * s0 + s1 * s1 + s2 * s2 + s3 * s3));  // (:310)
  const double f = seq - q.s0 - H * pi;                                        // (:312)
  const double fp = (f + sqrt(f * f)) / 2.0;                                   // (:313)
  const double dp = fp / (3 * m + H);                                          // (:315)
  const double v0 = s0 / seq * fp / f, v1 = s1 / seq * fp / f, v2 = s2 / seq * fp / f,
               v3 = s3 / seq * fp / f;                                         // (:317)
  const double beta = 3 * m * dp / seq;                                        // (:318)
  const double g0 = se0 - beta * s0, g1 = se1 - beta * s1, g2 = se2 - beta * s2, g3 = se3 - beta * s3;  // (:320)
  const double cn = 3 * m * (3 * m / (3 * m + H) - beta);                      // (:323)
  const double cd = 2 * m * beta;
  o.dp = dp;
  o.g[0] = g0, o.g[1] = g1, o.g[2] = g2, o.g[3] = g3;
  o.v[0] = v0, o.v[1] = v1, o.v[2] = v2, o.v[3] = v3, o.cn = cn, o.cd = cd;
  vm_tangent_exact(q, o.v, cn, cd, o.C);
}

// Same update with fewer instructions, for kernels that are issue-bound rather than HBM-bound (the fused
// tabulate + von Mises kernel).  The DECISION path - trial stress, deviator, sigma_eq, f - is the reference's
// exact statement sequence, so the plastic/elastic flag (dp > 0 <=> f > 0) is bit-identical to vm_point; the
// downstream algebra uses two divisions instead of nine (n = s * (f+ / (sigma_eq f))), a precomputed
// 1/(3 mu + H), the symmetry of the tangent and explicit FMAs.  Results agree with vm_point to a few ulp
// (tests: rtol 1e-12); elastic points still give dp = 0, sigma = sigma_trial and C_t = C_elas exactly.
EO_VM_HD void vm_point_fast(const vm_consts& q, double e0, double e1, double e2, double e3, double n0,
                                              double n1, double n2, double n3, double pi, vm_point_out& o) {
  const double l = q.l, m = q.m, H = q.H;
  const double l2m = l + 2.0 * m;
  const double se0 = n0 + (l2m * e0 + l * e1 + l * e2);  // (:308)
  const double se1 = n1 + (l * e0 + l2m * e1 + l * e2);
  const double se2 = n2 + (l * e0 + l * e1 + l2m * e2);
  const double se3 = n3 + 2.0 * m * e3;
  const double third = 1.0 / 3.0;
  const double tt = 1.0 - third;
  const double s0 = tt * se0 - third * se1 - third * se2;  // (:309)
  const double s1 = -third * se0 + tt * se1 - third * se2;
  const double s2 = -third * se0 - third * se1 + tt * se2;
  const double s3 = se3;
  const double seq = sqrt(3.0 / 2.0 * (s0 * s0 + s1 * s1 + s2 * s2 + s3 * s3));  // (:310)
  const double f = seq - q.s0 - H * pi;                                        // (:312)
  const double fp = (f + fabs(f)) * 0.5;  // == (f + sqrt(f*f)) / 2             (:313)
  const double m3 = 3 * m;
  const double i3mH = 1.0 / (m3 + H);
  const double dp = fp * i3mH;                    // (:315)
  const double r = fp / (seq * f);                // n = s / sigma_eq * f+ / f    (:317)
  const double v0 = s0 * r, v1 = s1 * r, v2 = s2 * r, v3 = s3 * r;
  const double beta = m3 * dp / seq;              // (:318)
  o.dp = dp;
  o.g[0] = fma(-beta, s0, se0), o.g[1] = fma(-beta, s1, se1), o.g[2] = fma(-beta, s2, se2), o.g[3] = fma(-beta, s3, se3);
  const double cn = m3 * (m3 * i3mH - beta);      // (:323)
  const double cd = 2 * m * beta;
  o.v[0] = v0, o.v[1] = v1, o.v[2] = v2, o.v[3] = v3, o.cn = cn, o.cd = cd;
  vm_tangent_fast(q, o.v, cn, cd, o.C);
}
