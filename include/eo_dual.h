/* eo_dual.h - forward-mode dual numbers for user-written per-quadrature-point models.
 *
 * replaces: the "any array library with automatic differentiation" promise of the reference
 * (README.md:16-25; JAX `jacfwd` in doc/demo/demo_plasticity_mohr_coulomb.py:555, torch.func in
 * doc/demo/demo_hyperelasticity.py:429-456).  A user writes ONE function template
 *
 *     template <class T>
 *     __device__ void model(const T* x, const double* state, const double* prm, T* y, T* aux);
 *
 * and libeo_b200.so (eo_jit_*, see eo_b200.h) compiles it with NVRTC for sm_100a once per derivative
 * multi-index: T = double for the value, T = eo::dual<N> for a first derivative with respect to an operand
 * with N components, T = eo::dual<N, eo::dual<M>> for a second derivative.  Everything is plain C++14,
 * needs no standard header (NVRTC has none) and also compiles on the host (g++) - the CPU tests of the
 * algebra in tests/hostcheck use exactly this file.
 *
 * Conventions: comparisons and the branch of fabs/fmax/fmin look at the primal value only (a model may
 * branch on `f > 0`, as the reference models do through `lax.cond`/`np.where`); `fabs'(0) = +1`,
 * `fmax(a, b)` picks `a` on ties.
 */
#ifndef EO_DUAL_H
#define EO_DUAL_H

#if defined(__CUDACC__) || defined(__CUDACC_RTC__)
#define EO_DUAL_HD __host__ __device__ __forceinline__
#else
#include <math.h>
#define EO_DUAL_HD inline
#endif

namespace eo {

/* scalar base cases: inside namespace eo an unqualified call must find something for double */
EO_DUAL_HD double sqrt(double a) { return ::sqrt(a); }
EO_DUAL_HD double cbrt(double a) { return ::cbrt(a); }
EO_DUAL_HD double exp(double a) { return ::exp(a); }
EO_DUAL_HD double expm1(double a) { return ::expm1(a); }
EO_DUAL_HD double log(double a) { return ::log(a); }
EO_DUAL_HD double log1p(double a) { return ::log1p(a); }
EO_DUAL_HD double sin(double a) { return ::sin(a); }
EO_DUAL_HD double cos(double a) { return ::cos(a); }
EO_DUAL_HD double tan(double a) { return ::tan(a); }
EO_DUAL_HD double asin(double a) { return ::asin(a); }
EO_DUAL_HD double acos(double a) { return ::acos(a); }
EO_DUAL_HD double atan(double a) { return ::atan(a); }
EO_DUAL_HD double atan2(double a, double b) { return ::atan2(a, b); }
EO_DUAL_HD double sinh(double a) { return ::sinh(a); }
EO_DUAL_HD double cosh(double a) { return ::cosh(a); }
EO_DUAL_HD double tanh(double a) { return ::tanh(a); }
EO_DUAL_HD double pow(double a, double b) { return ::pow(a, b); }
EO_DUAL_HD double fabs(double a) { return ::fabs(a); }
EO_DUAL_HD double fmax(double a, double b) { return a >= b ? a : b; }
EO_DUAL_HD double fmin(double a, double b) { return a <= b ? a : b; }
EO_DUAL_HD double value(double a) { return a; }

template <int N, class V = double>
struct dual {
  V v;     /* primal */
  V d[N];  /* tangent, one entry per seeded direction */

  EO_DUAL_HD dual() {}
  EO_DUAL_HD dual(double x) : v(x) {
    for (int i = 0; i < N; ++i) d[i] = V(0.0);
  }
  EO_DUAL_HD dual(int x) : v(double(x)) {
    for (int i = 0; i < N; ++i) d[i] = V(0.0);
  }
  /* r = s with tangent ds * a.d : the chain rule for a unary function */
  EO_DUAL_HD static dual chain(const dual& a, const V& s, const V& ds) {
    dual r;
    r.v = s;
    for (int i = 0; i < N; ++i) r.d[i] = ds * a.d[i];
    return r;
  }

  EO_DUAL_HD dual operator-() const {
    dual r;
    r.v = -v;
    for (int i = 0; i < N; ++i) r.d[i] = -d[i];
    return r;
  }
  EO_DUAL_HD dual operator+() const { return *this; }

  EO_DUAL_HD dual& operator+=(const dual& b) { return *this = *this + b; }
  EO_DUAL_HD dual& operator-=(const dual& b) { return *this = *this - b; }
  EO_DUAL_HD dual& operator*=(const dual& b) { return *this = *this * b; }
  EO_DUAL_HD dual& operator/=(const dual& b) { return *this = *this / b; }
  EO_DUAL_HD dual& operator+=(double b) { v = v + b; return *this; }
  EO_DUAL_HD dual& operator-=(double b) { v = v - b; return *this; }
  EO_DUAL_HD dual& operator*=(double b) { return *this = *this * b; }
  EO_DUAL_HD dual& operator/=(double b) { return *this = *this / b; }

  /* dual (op) dual */
  friend EO_DUAL_HD dual operator+(const dual& a, const dual& b) {
    dual r;
    r.v = a.v + b.v;
    for (int i = 0; i < N; ++i) r.d[i] = a.d[i] + b.d[i];
    return r;
  }
  friend EO_DUAL_HD dual operator-(const dual& a, const dual& b) {
    dual r;
    r.v = a.v - b.v;
    for (int i = 0; i < N; ++i) r.d[i] = a.d[i] - b.d[i];
    return r;
  }
  friend EO_DUAL_HD dual operator*(const dual& a, const dual& b) {
    dual r;
    r.v = a.v * b.v;
    for (int i = 0; i < N; ++i) r.d[i] = a.d[i] * b.v + a.v * b.d[i];
    return r;
  }
  friend EO_DUAL_HD dual operator/(const dual& a, const dual& b) {
    dual r;
    const V ib = 1.0 / b.v;
    r.v = a.v / b.v; /* the primal is the same IEEE quotient as in the T = double instantiation */
    for (int i = 0; i < N; ++i) r.d[i] = (a.d[i] - r.v * b.d[i]) * ib;
    return r;
  }
  /* dual (op) double, double (op) dual */
  friend EO_DUAL_HD dual operator+(const dual& a, double b) {
    dual r = a;
    r.v = a.v + b;
    return r;
  }
  friend EO_DUAL_HD dual operator+(double a, const dual& b) { return b + a; }
  friend EO_DUAL_HD dual operator-(const dual& a, double b) {
    dual r = a;
    r.v = a.v - b;
    return r;
  }
  friend EO_DUAL_HD dual operator-(double a, const dual& b) {
    dual r;
    r.v = a - b.v;
    for (int i = 0; i < N; ++i) r.d[i] = -b.d[i];
    return r;
  }
  friend EO_DUAL_HD dual operator*(const dual& a, double b) {
    dual r;
    r.v = a.v * b;
    for (int i = 0; i < N; ++i) r.d[i] = a.d[i] * b;
    return r;
  }
  friend EO_DUAL_HD dual operator*(double a, const dual& b) { return b * a; }
  friend EO_DUAL_HD dual operator/(const dual& a, double b) {
    dual r;
    const double ib = 1.0 / b;
    r.v = a.v / b;
    for (int i = 0; i < N; ++i) r.d[i] = a.d[i] * ib;
    return r;
  }
  friend EO_DUAL_HD dual operator/(double a, const dual& b) {
    dual r;
    const V ib = 1.0 / b.v;
    r.v = a / b.v;
    const V m = -(r.v * ib);
    for (int i = 0; i < N; ++i) r.d[i] = m * b.d[i];
    return r;
  }

  /* comparisons: primal only */
  friend EO_DUAL_HD bool operator<(const dual& a, const dual& b) { return value(a.v) < value(b.v); }
  friend EO_DUAL_HD bool operator>(const dual& a, const dual& b) { return value(a.v) > value(b.v); }
  friend EO_DUAL_HD bool operator<=(const dual& a, const dual& b) { return value(a.v) <= value(b.v); }
  friend EO_DUAL_HD bool operator>=(const dual& a, const dual& b) { return value(a.v) >= value(b.v); }
  friend EO_DUAL_HD bool operator==(const dual& a, const dual& b) { return value(a.v) == value(b.v); }
  friend EO_DUAL_HD bool operator!=(const dual& a, const dual& b) { return value(a.v) != value(b.v); }
  friend EO_DUAL_HD bool operator<(const dual& a, double b) { return value(a.v) < b; }
  friend EO_DUAL_HD bool operator>(const dual& a, double b) { return value(a.v) > b; }
  friend EO_DUAL_HD bool operator<=(const dual& a, double b) { return value(a.v) <= b; }
  friend EO_DUAL_HD bool operator>=(const dual& a, double b) { return value(a.v) >= b; }
  friend EO_DUAL_HD bool operator==(const dual& a, double b) { return value(a.v) == b; }
  friend EO_DUAL_HD bool operator!=(const dual& a, double b) { return value(a.v) != b; }
  friend EO_DUAL_HD bool operator<(double a, const dual& b) { return a < value(b.v); }
  friend EO_DUAL_HD bool operator>(double a, const dual& b) { return a > value(b.v); }
  friend EO_DUAL_HD bool operator<=(double a, const dual& b) { return a <= value(b.v); }
  friend EO_DUAL_HD bool operator>=(double a, const dual& b) { return a >= value(b.v); }
};

/* innermost primal of a (possibly nested) dual */
template <int N, class V>
EO_DUAL_HD double value(const dual<N, V>& a) {
  return value(a.v);
}

/* ---- elementary functions: r = f(a.v), r.d = f'(a.v) a.d ------------------------------------ */
template <int N, class V>
EO_DUAL_HD dual<N, V> sqrt(const dual<N, V>& a) {
  const V s = sqrt(a.v);
  return dual<N, V>::chain(a, s, 0.5 / s);
}
template <int N, class V>
EO_DUAL_HD dual<N, V> cbrt(const dual<N, V>& a) {
  const V s = cbrt(a.v);
  return dual<N, V>::chain(a, s, 1.0 / (3.0 * (s * s)));
}
template <int N, class V>
EO_DUAL_HD dual<N, V> exp(const dual<N, V>& a) {
  const V s = exp(a.v);
  return dual<N, V>::chain(a, s, s);
}
template <int N, class V>
EO_DUAL_HD dual<N, V> expm1(const dual<N, V>& a) {
  return dual<N, V>::chain(a, expm1(a.v), exp(a.v));
}
template <int N, class V>
EO_DUAL_HD dual<N, V> log(const dual<N, V>& a) {
  return dual<N, V>::chain(a, log(a.v), 1.0 / a.v);
}
template <int N, class V>
EO_DUAL_HD dual<N, V> log1p(const dual<N, V>& a) {
  return dual<N, V>::chain(a, log1p(a.v), 1.0 / (1.0 + a.v));
}
template <int N, class V>
EO_DUAL_HD dual<N, V> sin(const dual<N, V>& a) {
  return dual<N, V>::chain(a, sin(a.v), cos(a.v));
}
template <int N, class V>
EO_DUAL_HD dual<N, V> cos(const dual<N, V>& a) {
  return dual<N, V>::chain(a, cos(a.v), -sin(a.v));
}
template <int N, class V>
EO_DUAL_HD dual<N, V> tan(const dual<N, V>& a) {
  const V t = tan(a.v);
  return dual<N, V>::chain(a, t, 1.0 + t * t);
}
template <int N, class V>
EO_DUAL_HD dual<N, V> asin(const dual<N, V>& a) {
  return dual<N, V>::chain(a, asin(a.v), 1.0 / sqrt(1.0 - a.v * a.v));
}
template <int N, class V>
EO_DUAL_HD dual<N, V> acos(const dual<N, V>& a) {
  return dual<N, V>::chain(a, acos(a.v), -1.0 / sqrt(1.0 - a.v * a.v));
}
template <int N, class V>
EO_DUAL_HD dual<N, V> atan(const dual<N, V>& a) {
  return dual<N, V>::chain(a, atan(a.v), 1.0 / (1.0 + a.v * a.v));
}
template <int N, class V>
EO_DUAL_HD dual<N, V> sinh(const dual<N, V>& a) {
  return dual<N, V>::chain(a, sinh(a.v), cosh(a.v));
}
template <int N, class V>
EO_DUAL_HD dual<N, V> cosh(const dual<N, V>& a) {
  return dual<N, V>::chain(a, cosh(a.v), sinh(a.v));
}
template <int N, class V>
EO_DUAL_HD dual<N, V> tanh(const dual<N, V>& a) {
  const V t = tanh(a.v);
  return dual<N, V>::chain(a, t, 1.0 - t * t);
}
/* atan2(y, x): d = (x dy - y dx) / (x^2 + y^2) */
template <int N, class V>
EO_DUAL_HD dual<N, V> atan2(const dual<N, V>& y, const dual<N, V>& x) {
  dual<N, V> r;
  r.v = atan2(y.v, x.v);
  const V ir2 = 1.0 / (x.v * x.v + y.v * y.v);
  for (int i = 0; i < N; ++i) r.d[i] = (x.v * y.d[i] - y.v * x.d[i]) * ir2;
  return r;
}
/* a^b, constant exponent */
template <int N, class V>
EO_DUAL_HD dual<N, V> pow(const dual<N, V>& a, double b) {
  return dual<N, V>::chain(a, pow(a.v, b), b * pow(a.v, b - 1.0));
}
/* a^b, both varying (a > 0) */
template <int N, class V>
EO_DUAL_HD dual<N, V> pow(const dual<N, V>& a, const dual<N, V>& b) {
  dual<N, V> r;
  r.v = pow(a.v, b.v);
  const V da = b.v * pow(a.v, b.v - 1.0), db = r.v * log(a.v);
  for (int i = 0; i < N; ++i) r.d[i] = da * a.d[i] + db * b.d[i];
  return r;
}
template <int N, class V>
EO_DUAL_HD dual<N, V> pow(double a, const dual<N, V>& b) {
  const V s = pow(a, b.v);
  return dual<N, V>::chain(b, s, s * ::log(a));
}
template <int N, class V>
EO_DUAL_HD dual<N, V> fabs(const dual<N, V>& a) {
  return value(a.v) < 0.0 ? -a : a;
}
template <int N, class V>
EO_DUAL_HD dual<N, V> fmax(const dual<N, V>& a, const dual<N, V>& b) {
  return value(a.v) >= value(b.v) ? a : b;
}
template <int N, class V>
EO_DUAL_HD dual<N, V> fmin(const dual<N, V>& a, const dual<N, V>& b) {
  return value(a.v) <= value(b.v) ? a : b;
}
template <int N, class V>
EO_DUAL_HD dual<N, V> fmax(const dual<N, V>& a, double b) {
  return value(a.v) >= b ? a : dual<N, V>(b);
}
template <int N, class V>
EO_DUAL_HD dual<N, V> fmin(const dual<N, V>& a, double b) {
  return value(a.v) <= b ? a : dual<N, V>(b);
}
template <int N, class V>
EO_DUAL_HD dual<N, V> fmax(double a, const dual<N, V>& b) {
  return a >= value(b.v) ? dual<N, V>(a) : b;
}
template <int N, class V>
EO_DUAL_HD dual<N, V> fmin(double a, const dual<N, V>& b) {
  return a <= value(b.v) ? dual<N, V>(a) : b;
}

/* spellings users reach for: abs / max / min (CUDA defines ::abs, ::max, ::min for double) */
template <int N, class V>
EO_DUAL_HD dual<N, V> abs(const dual<N, V>& a) {
  return fabs(a);
}
template <int N, class V>
EO_DUAL_HD dual<N, V> max(const dual<N, V>& a, const dual<N, V>& b) {
  return fmax(a, b);
}
template <int N, class V>
EO_DUAL_HD dual<N, V> min(const dual<N, V>& a, const dual<N, V>& b) {
  return fmin(a, b);
}
template <int N, class V>
EO_DUAL_HD dual<N, V> max(const dual<N, V>& a, double b) {
  return fmax(a, b);
}
template <int N, class V>
EO_DUAL_HD dual<N, V> min(const dual<N, V>& a, double b) {
  return fmin(a, b);
}
template <int N, class V>
EO_DUAL_HD dual<N, V> hypot(const dual<N, V>& a, const dual<N, V>& b) {
  return sqrt(a * a + b * b);
}

/* `select(c, a, b)`: c ? a : b for any T (the model-side spelling of lax.cond / np.where) */
template <class T>
EO_DUAL_HD T select(bool c, const T& a, const T& b) {
  return c ? a : b;
}

}  // namespace eo

#endif /* EO_DUAL_H */
