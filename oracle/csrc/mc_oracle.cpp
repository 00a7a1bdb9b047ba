// TEST INFRASTRUCTURE - C++ restatement (the oracle / CPU baseline port) of the reference's
// Mohr-Coulomb return mapping with apex smoothing, doc/demo/demo_plasticity_mohr_coulomb.py
// (citations below are to that file, relative to /root/reference).  Never linked into the product.
//
// The reference obtains every derivative by JAX forward-mode AD:
//     dgdsigma     = jax.jacfwd(g)                                   (:391)
//     drdy         = jax.jacfwd(r)            [r calls dgdsigma]     (:465)
//     dsigma_ddeps = jax.jacfwd(return_mapping, has_aux=True)        (:555)  <- THROUGH lax.while_loop
// This file restates that structure literally with nested forward-mode dual numbers
// (Dual<Dual<Dual<double,4>,5>,4> at the innermost level), so the tangent is the derivative of
// the Newton ITERATION (tangents carried through every update from Y0 = 0), not the implicit-
// function tangent at the converged point.  JAX itself is not installable in the build container;
// this restatement is pinned against the reference's own source executed over a torch.func shim
// of the JAX API (oracle/jax_on_torch.py, tests/golden/mc_*.npz).
#include <cmath>
#include <cstdint>

// The arithmetic type of the restatement.  Default double = the reference's float64 program.  The library is built a
// second time with -DORACLE_REAL="long double" -DORACLE_SUFFIX=_ld (x87 extended precision, eps = 1.1e-19): the same
// program evaluated ~2000x more accurately, used as the "exact" value when the rounding sensitivity of the reference
// itself is audited near the corners of the Mohr-Coulomb hexagon (tests/test_mc_precision_cpu.py).
#ifndef ORACLE_REAL
#define ORACLE_REAL double
#endif
#ifndef ORACLE_SUFFIX
#define ORACLE_SUFFIX
#endif
#define ORACLE_CAT2(a, b) a##b
#define ORACLE_CAT(a, b) ORACLE_CAT2(a, b)
#define ORACLE_NAME(base) ORACLE_CAT(base, ORACLE_SUFFIX)
typedef ORACLE_REAL real;

#ifdef _OPENMP
#include <omp.h>
#endif

namespace {

// ------------------------------------------------------------------------------------------
// forward-mode dual numbers, nestable
// ------------------------------------------------------------------------------------------
template <class T, int N>
struct Dual {
  T v;
  T d[N];
  Dual() : v(0.0) {
    for (int i = 0; i < N; ++i) d[i] = T(0.0);
  }
  Dual(real c) : v(c) {  // NOLINT: implicit lift of a constant
    for (int i = 0; i < N; ++i) d[i] = T(0.0);
  }
  explicit Dual(const T& val, int) : v(val) {
    for (int i = 0; i < N; ++i) d[i] = T(0.0);
  }
};

inline real primal(real x) { return x; }
template <class T, int N>
inline real primal(const Dual<T, N>& x) {
  return primal(x.v);
}

#define AD_BIN(OP, BODY_DD, BODY_DS, BODY_SD)                                   \
  template <class T, int N>                                                     \
  inline Dual<T, N> operator OP(const Dual<T, N>& a, const Dual<T, N>& b) {     \
    Dual<T, N> r;                                                               \
    BODY_DD return r;                                                           \
  }                                                                             \
  template <class T, int N>                                                     \
  inline Dual<T, N> operator OP(const Dual<T, N>& a, real b) {                \
    Dual<T, N> r;                                                               \
    BODY_DS return r;                                                           \
  }                                                                             \
  template <class T, int N>                                                     \
  inline Dual<T, N> operator OP(real a, const Dual<T, N>& b) {                \
    Dual<T, N> r;                                                               \
    BODY_SD return r;                                                           \
  }

AD_BIN(+, r.v = a.v + b.v; for (int i = 0; i < N; ++i) r.d[i] = a.d[i] + b.d[i];
       , r.v = a.v + b; for (int i = 0; i < N; ++i) r.d[i] = a.d[i];
       , r.v = a + b.v; for (int i = 0; i < N; ++i) r.d[i] = b.d[i];)
AD_BIN(-, r.v = a.v - b.v; for (int i = 0; i < N; ++i) r.d[i] = a.d[i] - b.d[i];
       , r.v = a.v - b; for (int i = 0; i < N; ++i) r.d[i] = a.d[i];
       , r.v = a - b.v; for (int i = 0; i < N; ++i) r.d[i] = -b.d[i];)
AD_BIN(*, r.v = a.v * b.v; for (int i = 0; i < N; ++i) r.d[i] = a.d[i] * b.v + a.v * b.d[i];
       , r.v = a.v * b; for (int i = 0; i < N; ++i) r.d[i] = a.d[i] * b;
       , r.v = a * b.v; for (int i = 0; i < N; ++i) r.d[i] = a * b.d[i];)
AD_BIN(/, r.v = a.v / b.v; for (int i = 0; i < N; ++i) r.d[i] = (a.d[i] - r.v * b.d[i]) / b.v;
       , r.v = a.v / b; for (int i = 0; i < N; ++i) r.d[i] = a.d[i] / b;
       , r.v = a / b.v; for (int i = 0; i < N; ++i) r.d[i] = -(r.v * b.d[i]) / b.v;)
#undef AD_BIN

template <class T, int N>
inline Dual<T, N> operator-(const Dual<T, N>& a) {
  Dual<T, N> r;
  r.v = -a.v;
  for (int i = 0; i < N; ++i) r.d[i] = -a.d[i];
  return r;
}

inline real ad_sqrt(real x) { return std::sqrt(x); }
inline real ad_sin(real x) { return std::sin(x); }
inline real ad_cos(real x) { return std::cos(x); }
inline real ad_asin(real x) { return std::asin(x); }
inline real ad_abs(real x) { return std::fabs(x); }
inline real ad_clip(real x, real lo, real hi) { return x < lo ? lo : (x > hi ? hi : x); }

template <class T, int N>
inline Dual<T, N> ad_sqrt(const Dual<T, N>& a) {
  Dual<T, N> r;
  r.v = ad_sqrt(a.v);
  for (int i = 0; i < N; ++i) r.d[i] = a.d[i] / (2.0 * r.v);
  return r;
}
template <class T, int N>
inline Dual<T, N> ad_sin(const Dual<T, N>& a) {
  Dual<T, N> r;
  r.v = ad_sin(a.v);
  const T c = ad_cos(a.v);
  for (int i = 0; i < N; ++i) r.d[i] = a.d[i] * c;
  return r;
}
template <class T, int N>
inline Dual<T, N> ad_cos(const Dual<T, N>& a) {
  Dual<T, N> r;
  r.v = ad_cos(a.v);
  const T s = ad_sin(a.v);
  for (int i = 0; i < N; ++i) r.d[i] = -(a.d[i] * s);
  return r;
}
template <class T, int N>
inline Dual<T, N> ad_asin(const Dual<T, N>& a) {
  Dual<T, N> r;
  r.v = ad_asin(a.v);
  const T g = 1.0 / ad_sqrt(1.0 - a.v * a.v);
  for (int i = 0; i < N; ++i) r.d[i] = a.d[i] * g;
  return r;
}
// clip(x, lo, hi): tangent passes inside the interval, is zero outside (jnp.clip, :293)
template <class T, int N>
inline Dual<T, N> ad_clip(const Dual<T, N>& a, real lo, real hi) {
  const real p = primal(a);
  if (p < lo) return Dual<T, N>(lo);
  if (p > hi) return Dual<T, N>(hi);
  return a;
}

// ------------------------------------------------------------------------------------------
// model
// ------------------------------------------------------------------------------------------
struct McParamsIO {  // as passed through the C interface
  double E, nu, c, phi, psi, theta_T, a;  // :110-116
  double tol;                             // :469
  int32_t Nitermax;                       // :469
};
struct McParams {
  real E, nu, c, phi, psi, theta_T, a, tol;
  int32_t Nitermax;
  McParams() {}
  McParams(const McParamsIO& q)  // NOLINT: the double parameter values are taken as exact
      : E(q.E), nu(q.nu), c(q.c), phi(q.phi), psi(q.psi), theta_T(q.theta_T), a(q.a), tol(q.tol), Nitermax(q.Nitermax) {}
};

struct Consts {
  real C[4][4];  // C_elas :407-415
  real dev[4][4];
  McParams p;
};

inline void make_consts(const McParams& p, Consts& k) {
  const real lmbda = p.E * p.nu / ((1.0 + p.nu) * (1.0 - 2.0 * p.nu));  // :405
  const real mu = p.E / (2.0 * (1.0 + p.nu));                           // :406
  for (int i = 0; i < 4; ++i)
    for (int j = 0; j < 4; ++j) {
      k.C[i][j] = 0.0;
      k.dev[i][j] = 0.0;
    }
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) {
      k.C[i][j] = (i == j) ? lmbda + 2 * mu : lmbda;
      k.dev[i][j] = (i == j) ? real(2.0) / real(3.0) : real(-1.0) / real(3.0);  // :352-360
    }
  k.C[3][3] = 2 * mu;
  k.dev[3][3] = 1.0;
  k.p = p;
}

// K(theta, angle) with the Abbo-Sloan rounding, :298-345.  sign(0) = +1 (:298-299).
template <class T>
inline T K_fun(const T& th, real angle, const McParams& p) {
  const real tT = p.theta_T;
  const real sa = std::sin(angle);
  const real isq3 = 1.0 / std::sqrt(real(3.0));
  if (std::fabs(primal(th)) > tT) {  // K_true :338-343
    const real sg = primal(th) < 0.0 ? -1.0 : 1.0;
    const real c1 = std::cos(tT) - isq3 * sa * std::sin(tT);                     // :302-303
    const real c2 = sg * std::sin(tT) + isq3 * sa * std::cos(tT);                // :306-307
    const real c3 = 18.0 * std::cos(3.0 * tT) * std::cos(3.0 * tT) * std::cos(3.0 * tT);  // :310
    const real Cc = (-std::cos(3.0 * tT) * c1 - 3.0 * sg * std::sin(3.0 * tT) * c2) / c3;  // :313-316
    const real Bc = (sg * std::sin(6.0 * tT) * c1 - 6.0 * std::cos(6.0 * tT) * c2) / c3;   // :319-322
    const real Ac = -isq3 * sa * sg * std::sin(tT) - Bc * sg * std::sin(3 * tT) -
                      Cc * std::sin(3.0 * tT) * std::sin(3.0 * tT) + std::cos(tT);          // :325-331
    const T s3 = ad_sin(3.0 * th);
    return Ac + Bc * s3 + Cc * s3 * s3;
  }
  return ad_cos(th) - isq3 * sa * ad_sin(th);  // K_false :335-336
}

// surface(sigma, angle), :364-374
template <class T>
inline T surface(const T sig[4], real angle, const Consts& k) {
  const McParams& p = k.p;
  T s[4];
  for (int i = 0; i < 4; ++i) {
    T acc = k.dev[i][0] * sig[0];
    for (int j = 1; j < 4; ++j) acc = acc + k.dev[i][j] * sig[j];
    s[i] = acc;
  }
  const T I1 = sig[0] + sig[1] + sig[2];                                   // tr @ sigma :361,366
  const T J2 = 0.5 * (s[0] * s[0] + s[1] * s[1] + s[2] * s[2] + s[3] * s[3]);  // :286-287
  const T J3 = s[2] * (s[0] * s[1] - s[3] * s[3] / 2.0);                   // :282-283
  T arg = -(3.0 * std::sqrt(real(3.0)) * J3) / (2.0 * ad_sqrt(J2 * J2 * J2));    // :292
  arg = ad_clip(arg, -1.0, 1.0);                                           // :293
  const T th = real(1.0) / real(3.0) * ad_asin(arg);                                   // :294
  const T K = K_fun(th, angle, p);
  const real ag = p.a * std::tan(p.phi) / std::tan(angle);               // :348-349
  const real sa = std::sin(angle);
  return (I1 / 3.0 * sa) + ad_sqrt(J2 * K * K + ag * ag * sa * sa) - p.c * std::cos(angle);  // :368-374
}

// dgdsigma = jacfwd(g), :391
template <class T>
inline void dgdsigma(const T sig[4], const Consts& k, T out[4]) {
  typedef Dual<T, 4> D;
  D x[4];
  for (int i = 0; i < 4; ++i) {
    x[i] = D(sig[i], 0);
    x[i].d[i] = T(1.0);
  }
  const D g = surface(x, k.p.psi, k);
  for (int i = 0; i < 4; ++i) out[i] = g.d[i];
}

// r(y, deps, sigma_n), :420-463.  `plastic` is the (piecewise constant) trial-state predicate
// `yielding <= 0` of :422-430/:440-448 evaluated by the caller on primal values.
template <class T>
inline void residual(const T y[5], const T deps[4], const T sn[4], bool plastic, const Consts& k, T res[5]) {
  T eps_e[4];
  if (plastic) {
    T dg[4];
    dgdsigma(y, k, dg);
    for (int i = 0; i < 4; ++i) eps_e[i] = deps[i] - y[4] * dg[i];  // deps - dlambda * dgdsigma :427,435
  } else {
    for (int i = 0; i < 4; ++i) eps_e[i] = deps[i] - T(0.0);
  }
  for (int i = 0; i < 4; ++i) {
    T acc = k.C[i][0] * eps_e[0];
    for (int j = 1; j < 4; ++j) acc = acc + k.C[i][j] * eps_e[j];
    res[i] = y[i] - sn[i] - acc;  // :435
  }
  res[4] = plastic ? surface(y, k.p.phi, k) : y[4];  // :442-446
}

// drdy = jacfwd(r), :465
template <class T>
inline void residual_jacobian(const T y[5], const T deps[4], const T sn[4], bool plastic, const Consts& k,
                              T J[5][5]) {
  typedef Dual<T, 5> D;
  D yy[5], dd[4], ss[4], rr[5];
  for (int i = 0; i < 5; ++i) {
    yy[i] = D(y[i], 0);
    yy[i].d[i] = T(1.0);
  }
  for (int i = 0; i < 4; ++i) {
    dd[i] = D(deps[i], 0);
    ss[i] = D(sn[i], 0);
  }
  residual(yy, dd, ss, plastic, k, rr);
  for (int i = 0; i < 5; ++i)
    for (int j = 0; j < 5; ++j) J[i][j] = rr[i].d[j];
}

// jnp.linalg.solve: LU with partial pivoting (pivot chosen on primal magnitudes), :512
template <class T>
inline void solve5(T A[5][5], T b[5], T x[5]) {
  int piv[5];
  for (int i = 0; i < 5; ++i) piv[i] = i;
  for (int c = 0; c < 5; ++c) {
    int best = c;
    real bv = std::fabs(primal(A[piv[c]][c]));
    for (int r = c + 1; r < 5; ++r) {
      const real v = std::fabs(primal(A[piv[r]][c]));
      if (v > bv) {
        bv = v;
        best = r;
      }
    }
    const int t = piv[c];
    piv[c] = piv[best];
    piv[best] = t;
    const int pc = piv[c];
    for (int r = c + 1; r < 5; ++r) {
      const int pr = piv[r];
      const T m = A[pr][c] / A[pc][c];
      for (int j = c + 1; j < 5; ++j) A[pr][j] = A[pr][j] - m * A[pc][j];
      b[pr] = b[pr] - m * b[pc];
    }
  }
  for (int c = 4; c >= 0; --c) {
    const int pc = piv[c];
    T acc = b[pc];
    for (int j = c + 1; j < 5; ++j) acc = acc - A[pc][j] * x[j];
    x[c] = acc / A[pc][c];
  }
}

template <class T>
inline T norm5(const T r[5]) {
  return ad_sqrt(r[0] * r[0] + r[1] * r[1] + r[2] * r[2] + r[3] * r[3] + r[4] * r[4]);
}

// return_mapping, :474-533, on a generic scalar T (real: values only; Dual<real,4>: values
// and d/d(deps) carried through the loop).
template <class T>
inline void return_mapping(const T deps[4], const T sn[4], const Consts& k, T sigma[4], int32_t& niter_out,
                           real& yielding_out, real& norm_res_out, T& dlambda_out) {
  // trial-state predicate :421-422 (value only: lax.cond predicates carry no tangent)
  real trial[4];
  for (int i = 0; i < 4; ++i) {
    real acc = k.C[i][0] * primal(deps[0]);
    for (int j = 1; j < 4; ++j) acc += k.C[i][j] * primal(deps[j]);
    trial[i] = primal(sn[i]) + acc;
  }
  const real yielding = surface(trial, k.p.phi, k);
  const bool plastic = !(yielding <= 0.0);

  T y[5], res[5];
  for (int i = 0; i < 4; ++i) y[i] = sn[i];  // :496-498
  y[4] = T(0.0);
  residual(y, deps, sn, plastic, k, res);    // :500
  const real norm0 = primal(norm5(res));   // :501
  real nrm = norm0;
  int32_t niter = 0;
  while ((nrm / norm0 > k.p.tol) && (niter < k.p.Nitermax)) {  // :503-505
    T J[5][5], rhs[5], dy[5];
    residual_jacobian(y, deps, sn, plastic, k, J);  // :511
    for (int i = 0; i < 5; ++i) rhs[i] = -res[i];
    solve5(J, rhs, dy);                             // :512
    for (int i = 0; i < 5; ++i) y[i] = y[i] + dy[i];  // :513
    residual(y, deps, sn, plastic, k, res);         // :515
    nrm = primal(norm5(res));                       // :516
    niter += 1;                                     // :519
  }
  for (int i = 0; i < 4; ++i) sigma[i] = y[i];  // :527
  dlambda_out = y[4];                           // :528
  niter_out = niter;
  yielding_out = yielding;                      // :529-530
  norm_res_out = nrm;
}

void mc_point(const Consts& k, const double* deps, const double* sn, double* Ct, double* sig, int32_t* niter,
              double* yielding, double* norm_res, double* dlambda) {
  typedef Dual<real, 4> D;  // jacfwd over deps, :555
  D de[4], s0[4], so[4], dl;
  for (int i = 0; i < 4; ++i) {
    de[i] = D(real(deps[i]));
    de[i].d[i] = 1.0;
    s0[i] = D(real(sn[i]));
  }
  int32_t it;
  real yl, nr;
  return_mapping(de, s0, k, so, it, yl, nr, dl);
  for (int i = 0; i < 4; ++i) {
    sig[i] = (double)so[i].v;
    for (int j = 0; j < 4; ++j) Ct[4 * i + j] = (double)so[i].d[j];
  }
  *niter = it;
  *yielding = (double)yl;
  *norm_res = (double)nr;
  *dlambda = (double)dl.v;
}

}  // namespace

extern "C" {

typedef McParamsIO oracle_mc_params;

// Layouts as the reference returns them (:593): deps/sigma_n/sigma [n][4], C_tang [n][4][4];
// aux outputs per point as in the aux tuple of :533.  Exported as oracle_mc_return_mapping (real = double) and
// oracle_mc_return_mapping_ld (real = long double, results rounded to double).
void ORACLE_NAME(oracle_mc_return_mapping)(const oracle_mc_params* prm, const double* deps, const double* sigma_n,
                                           double* C_tang, double* sigma, int32_t* niter, double* yielding,
                                           double* norm_res, double* dlambda, int64_t n, int parallel) {
  Consts k;
  make_consts(McParams(*prm), k);
#pragma omp parallel for schedule(dynamic, 64) if (parallel)
  for (int64_t i = 0; i < n; ++i)
    mc_point(k, deps + 4 * i, sigma_n + 4 * i, C_tang + 16 * i, sigma + 4 * i, niter + i, yielding + i, norm_res + i,
             dlambda + i);
}

// Values only (no tangent): stress update used to generate stress paths.
void ORACLE_NAME(oracle_mc_stress)(const oracle_mc_params* prm, const double* deps, const double* sigma_n, double* sigma,
                                   int32_t* niter, double* yielding, int64_t n, int parallel) {
  Consts k;
  make_consts(McParams(*prm), k);
#pragma omp parallel for schedule(dynamic, 64) if (parallel)
  for (int64_t i = 0; i < n; ++i) {
    real de[4], sn[4], sg[4], nr, dl, yl;
    for (int j = 0; j < 4; ++j) de[j] = deps[4 * i + j], sn[j] = sigma_n[4 * i + j];
    return_mapping<real>(de, sn, k, sg, niter[i], yl, nr, dl);
    for (int j = 0; j < 4; ++j) sigma[4 * i + j] = (double)sg[j];
    yielding[i] = (double)yl;
  }
}

// f(sigma) and g-gradient, exposed for known-answer tests (f(sigma_returned) ~ 0 etc.)
void ORACLE_NAME(oracle_mc_yield)(const oracle_mc_params* prm, const double* sigma, double* f, int64_t n) {
  Consts k;
  make_consts(McParams(*prm), k);
  for (int64_t i = 0; i < n; ++i) {
    real sg[4];
    for (int j = 0; j < 4; ++j) sg[j] = sigma[4 * i + j];
    f[i] = (double)surface(sg, k.p.phi, k);
  }
}

}  // extern "C"
