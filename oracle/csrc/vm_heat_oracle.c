/* TEST INFRASTRUCTURE - plain-C restatement (the oracle / CPU baseline port) of
 * the reference's closed-form constitutive callables.  Never linked into the
 * product library.  Citations are relative to /root/reference.
 *
 *   von Mises : doc/demo/demo_plasticity_von_mises.py:298-332 (`return_mapping`)
 *   heat      : doc/demo/demo_nonlinear_heat_equation_part1.py:252-272,
 *               doc/demo/demo_nonlinear_heat_equation_part2.py:215-261
 *
 * Built with -ffp-contract=off so that the arithmetic is the plain IEEE
 * sequence of the source statements (no fused multiply-adds).
 */
#include <math.h>
#include <stdint.h>

#ifdef _OPENMP
#include <omp.h>
#endif

typedef struct {
  double lmbda, mu, H, sigma_0;
} oracle_vm_params;

int oracle_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

/* torchrun exports OMP_NUM_THREADS=1 to every rank; the CPU-baseline legs of bench.py run on rank 0 only and
 * may use all host cores. */
void oracle_set_num_threads(int n) {
#ifdef _OPENMP
  if (n > 0) omp_set_num_threads(n);
#else
  (void)n;
#endif
}

/* One quadrature point; mirrors `_kernel` demo_vm:307-326 statement by statement. */
static void vm_point(const oracle_vm_params* q, const double* deps, const double* sn, double p, double* Ct,
                     double* sig, double* dp_out) {
  const double l = q->lmbda, m = q->mu, H = q->H;
  /* sigma_elastic = sigma_n + C_elas @ deps      (:308, C_elas :193-201) */
  const double tr_e = deps[0] + deps[1] + deps[2];
  double se[4];
  se[0] = sn[0] + ((l + 2.0 * m) * deps[0] + l * deps[1] + l * deps[2]);
  se[1] = sn[1] + (l * deps[0] + (l + 2.0 * m) * deps[1] + l * deps[2]);
  se[2] = sn[2] + (l * deps[0] + l * deps[1] + (l + 2.0 * m) * deps[2]);
  se[3] = sn[3] + 2.0 * m * deps[3];
  (void)tr_e;
  /* s = deviatoric @ sigma_elastic               (:309, deviatoric :203-204) */
  const double third = 1.0 / 3.0;
  double s[4];
  s[0] = (1.0 - third) * se[0] - third * se[1] - third * se[2];
  s[1] = -third * se[0] + (1.0 - third) * se[1] - third * se[2];
  s[2] = -third * se[0] - third * se[1] + (1.0 - third) * se[2];
  s[3] = se[3];
  /* sigma_eq = sqrt(3/2 s.s)                     (:310) */
  const double seq = sqrt(3.0 / 2.0 * (s[0] * s[0] + s[1] * s[1] + s[2] * s[2] + s[3] * s[3]));
  const double f = seq - q->sigma_0 - H * p;             /* :312 */
  const double fp = (f + sqrt(f * f)) / 2.0;             /* :313 */
  const double dp = fp / (3 * m + H);                    /* :315 */
  double n[4];
  for (int i = 0; i < 4; ++i) n[i] = s[i] / seq * fp / f; /* :317 */
  const double beta = 3 * m * dp / seq;                  /* :318 */
  for (int i = 0; i < 4; ++i) sig[i] = se[i] - beta * s[i]; /* :320 */
  const double cn = 3 * m * (3 * m / (3 * m + H) - beta); /* :323 */
  const double cd = 2 * m * beta;
  for (int i = 0; i < 4; ++i)
    for (int j = 0; j < 4; ++j) {
      double Cij = 0.0;
      if (i < 3 && j < 3) Cij = (i == j) ? l + 2.0 * m : l;
      if (i == 3 && j == 3) Cij = 2.0 * m;
      double Dij = (i == j) ? 1.0 : 0.0;
      if (i < 3 && j < 3) Dij -= third;
      Ct[4 * i + j] = Cij - cn * (n[i] * n[j]) - cd * Dij;
    }
  *dp_out = dp;
}

/* the same point update for the cell loop of forms_oracle.c */
void oracle_vm_point(const oracle_vm_params* q, const double* deps, const double* sn, double p, double* Ct, double* sig,
                     double* dp_out) {
  vm_point(q, deps, sn, p, Ct, sig, dp_out);
}

/* AoS layouts exactly as the reference returns them (demo_vm:352):
 * deps/sigma_n/sigma [n][4], p/dp [n], C_tang [n][4][4]. */
void oracle_vm_return_mapping(const oracle_vm_params* q, const double* deps, const double* sigma_n, const double* p,
                              double* C_tang, double* sigma, double* dp, int64_t n, int parallel) {
#pragma omp parallel for schedule(static) if (parallel)
  for (int64_t i = 0; i < n; ++i) vm_point(q, deps + 4 * i, sigma_n + 4 * i, p[i], C_tang + 16 * i, sigma + 4 * i, dp + i);
}

/* heat, part1.py:252-272 and part2.py:219-261; gdim = 2.
 * which: 0 k, 1 dk/dT, 2 q, 3 dq/dT, 4 dq/dsigma.  Output flat as the reference. */
void oracle_heat(int which, double A, double B, const double* T, const double* sigma, double* out, int64_t n,
                 int parallel) {
#pragma omp parallel for schedule(static) if (parallel)
  for (int64_t i = 0; i < n; ++i) {
    const double k = 1.0 / (A + B * T[i]);
    switch (which) {
      case 0: out[i] = k; break;
      case 1: out[i] = -B * (k * k); break;
      case 2:
        out[2 * i] = -k * sigma[2 * i];
        out[2 * i + 1] = -k * sigma[2 * i + 1];
        break;
      case 3:
        out[2 * i] = B * (k * k) * sigma[2 * i];
        out[2 * i + 1] = B * (k * k) * sigma[2 * i + 1];
        break;
      default:
        out[4 * i] = -k * 1.0;
        out[4 * i + 1] = -k * 0.0;
        out[4 * i + 2] = -k * 0.0;
        out[4 * i + 3] = -k * 1.0;
    }
  }
}
