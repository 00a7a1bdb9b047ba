// vm_core.cuh - per-quadrature-point arithmetic of the von Mises radial return
// (reference: doc/demo/demo_plasticity_von_mises.py:307-326, statement by statement).  Shared by the
// streaming kernel (vm_heat.cu) and the kernel fused with the strain tabulation (tab.cu); both are
// compiled with -fmad=false so that the results are the plain IEEE sequence of the reference's statements.
#pragma once

struct vm_consts {
  double l, m, H, s0;
};

struct vm_point_out {
  double C[16];  // consistent tangent, row-major 4x4
  double g[4];   // new stress
  double dp;     // plastic multiplier increment
};

__device__ __forceinline__ void vm_point(const vm_consts& q, double e0, double e1, double e2, double e3, double n0,
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
  const double Dd = 1.0 - third, Do = 0.0 - third;  // deviatoric diagonal / off-diagonal (3x3 block)
  o.dp = dp;
  o.g[0] = g0, o.g[1] = g1, o.g[2] = g2, o.g[3] = g3;
  o.C[0] = l2m - cn * (v0 * v0) - cd * Dd, o.C[1] = l - cn * (v0 * v1) - cd * Do, o.C[2] = l - cn * (v0 * v2) - cd * Do,
  o.C[3] = 0.0 - cn * (v0 * v3) - cd * 0.0;
  o.C[4] = l - cn * (v1 * v0) - cd * Do, o.C[5] = l2m - cn * (v1 * v1) - cd * Dd, o.C[6] = l - cn * (v1 * v2) - cd * Do,
  o.C[7] = 0.0 - cn * (v1 * v3) - cd * 0.0;
  o.C[8] = l - cn * (v2 * v0) - cd * Do, o.C[9] = l - cn * (v2 * v1) - cd * Do, o.C[10] = l2m - cn * (v2 * v2) - cd * Dd,
  o.C[11] = 0.0 - cn * (v2 * v3) - cd * 0.0;
  o.C[12] = 0.0 - cn * (v3 * v0) - cd * 0.0, o.C[13] = 0.0 - cn * (v3 * v1) - cd * 0.0,
  o.C[14] = 0.0 - cn * (v3 * v2) - cd * 0.0, o.C[15] = 2.0 * m - cn * (v3 * v3) - cd * 1.0;
}
