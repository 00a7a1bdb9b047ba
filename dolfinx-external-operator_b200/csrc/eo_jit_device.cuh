// eo_jit_device.cuh - device side of the generic (NVRTC) quadrature-point path.  This text is embedded in
// libeo_b200.so and handed to NVRTC as the header "eo_jit_device.cuh"; jit.cu generates the few lines that
// instantiate eo_jit_run<Spec> for one model and one derivative multi-index.
//
// One thread = one quadrature point (the same mapping as vm_kernel): the point's operand / state components
// are contiguous in the reference's flat layout [point][component] (external_operator.py:396,
// demo_vm:344-352), so a thread reads them with the widest access its alignment allows (256-bit when the
// component count is a multiple of 4, 128-bit for a multiple of 2) through the read-only, no-L1-allocate
// path, seeds the dual numbers, calls the user's model and streams the results back the same way.
#pragma once
#include "eo_dual.h"
#ifdef EO_JIT_FUSED
#include "tab_core.cuh"
#endif

#define EO_JIT_MAX_ARGS 8
#define EO_JIT_MAX_PARAMS 32

struct eo_jit_args {
  const double* operand[EO_JIT_MAX_ARGS];
  const double* state[EO_JIT_MAX_ARGS];
  double* out;    // value (order 0) or derivative (order 1, 2)
  double* value;  // order >= 1: the operator's value, optional
  double* aux[EO_JIT_MAX_ARGS];
  long long n;
  double prm[EO_JIT_MAX_PARAMS];
};

// Fused variant (operands tabulated in the kernel from DOF coefficients, never stored): one tabulation source
// per operand.  `operand[]` of the base is unused.
struct tab_tables;
struct eo_jit_fused_args : eo_jit_args {
  const tab_tables* T[EO_JIT_MAX_ARGS];  // device copies of the element tables (staged into shared memory)
  const int* dofmap[EO_JIT_MAX_ARGS];
  const int* x_dofmap[EO_JIT_MAX_ARGS];
  const double* x[EO_JIT_MAX_ARGS];
  const double* u[EO_JIT_MAX_ARGS];
  long long point_offset;  // global index of this launch's first point (cell = point / NQ); chunked host pipeline
};

// widest vector access in doubles: 4 (256-bit, needs the CUDA >= 12.9 ptxas) or 2 (jit.cu passes
// -DEO_JIT_MAX_VEC=2 when the NVRTC it found is older)
#ifndef EO_JIT_MAX_VEC
#define EO_JIT_MAX_VEC 4
#endif

namespace eo_jitd {

#if EO_JIT_MAX_VEC >= 4
__device__ __forceinline__ void ld(const double* p, double& a, double& b, double& c, double& d) {
  asm volatile("ld.global.nc.L1::no_allocate.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(a), "=d"(b), "=d"(c), "=d"(d) : "l"(p));
}
__device__ __forceinline__ void st(double* p, double a, double b, double c, double d) {
  asm volatile("st.global.L1::no_allocate.v4.f64 [%0], {%1,%2,%3,%4};" ::"l"(p), "d"(a), "d"(b), "d"(c), "d"(d) : "memory");
}
#else
__device__ __forceinline__ void ld(const double* p, double& a, double& b, double& c, double& d) {
  asm volatile("ld.global.nc.L1::no_allocate.v2.f64 {%0,%1}, [%2];" : "=d"(a), "=d"(b) : "l"(p));
  asm volatile("ld.global.nc.L1::no_allocate.v2.f64 {%0,%1}, [%2];" : "=d"(c), "=d"(d) : "l"(p + 2));
}
__device__ __forceinline__ void st(double* p, double a, double b, double c, double d) {
  asm volatile("st.global.L1::no_allocate.v2.f64 [%0], {%1,%2};" ::"l"(p), "d"(a), "d"(b) : "memory");
  asm volatile("st.global.L1::no_allocate.v2.f64 [%0], {%1,%2};" ::"l"(p + 2), "d"(c), "d"(d) : "memory");
}
#endif
__device__ __forceinline__ void ld(const double* p, double& a, double& b) {
  asm volatile("ld.global.nc.L1::no_allocate.v2.f64 {%0,%1}, [%2];" : "=d"(a), "=d"(b) : "l"(p));
}
__device__ __forceinline__ void ld(const double* p, double& a) {
  asm volatile("ld.global.nc.L1::no_allocate.f64 %0, [%1];" : "=d"(a) : "l"(p));
}
__device__ __forceinline__ void st(double* p, double a, double b) {
  asm volatile("st.global.L1::no_allocate.v2.f64 [%0], {%1,%2};" ::"l"(p), "d"(a), "d"(b) : "memory");
}
__device__ __forceinline__ void st(double* p, double a) {
  asm volatile("st.global.L1::no_allocate.f64 [%0], %1;" ::"l"(p), "d"(a) : "memory");
}

#ifdef EO_JIT_STAGED
// staged variant: a point's components are read from / written to the CTA's shared-memory tile
__device__ __forceinline__ void lds(unsigned a, double& x, double& y) {
  asm volatile("ld.shared.v2.f64 {%0,%1}, [%2];" : "=d"(x), "=d"(y) : "r"(a));
}
__device__ __forceinline__ void lds(unsigned a, double& x) { asm volatile("ld.shared.f64 %0, [%1];" : "=d"(x) : "r"(a)); }
__device__ __forceinline__ void sts(unsigned a, double x, double y) {
  asm volatile("st.shared.v2.f64 [%0], {%1,%2};" ::"r"(a), "d"(x), "d"(y) : "memory");
}
__device__ __forceinline__ void sts(unsigned a, double x) { asm volatile("st.shared.f64 [%0], %1;" ::"r"(a), "d"(x) : "memory"); }
#endif

// S contiguous doubles of one point in GLOBAL memory; the base pointer is aligned for the widest access used
template <int S>
__device__ __forceinline__ void gload_point(const double* __restrict__ base, long long i, double* r) {
  const double* p = base + i * S;
  if constexpr (S % 4 == 0 && EO_JIT_MAX_VEC >= 4) {
#pragma unroll
    for (int k = 0; k < S; k += 4) ld(p + k, r[k], r[k + 1], r[k + 2], r[k + 3]);
  } else if constexpr (S % 2 == 0) {
#pragma unroll
    for (int k = 0; k < S; k += 2) ld(p + k, r[k], r[k + 1]);
  } else {
#pragma unroll
    for (int k = 0; k < S; ++k) ld(p + k, r[k]);
  }
}
template <int S>
__device__ __forceinline__ void gstore_point(double* __restrict__ base, long long i, const double* r) {
  double* p = base + i * S;
  if constexpr (S % 4 == 0 && EO_JIT_MAX_VEC >= 4) {
#pragma unroll
    for (int k = 0; k < S; k += 4) st(p + k, r[k], r[k + 1], r[k + 2], r[k + 3]);
  } else if constexpr (S % 2 == 0) {
#pragma unroll
    for (int k = 0; k < S; k += 2) st(p + k, r[k], r[k + 1]);
  } else {
#pragma unroll
    for (int k = 0; k < S; ++k) st(p + k, r[k]);
  }
}

// Where the per-point code (run0/1/2) finds a point's components: global memory (direct variant), the CTA's
// shared-memory tile (EO_JIT_STAGED) or the thread's own registers (EO_JIT_PPT: several points per thread, fetched
// together with wide accesses by run_ppt).
template <int S>
__device__ __forceinline__ void load_point(const double* __restrict__ base, long long i, double* r) {
#if defined(EO_JIT_STAGED)
  const unsigned a = (unsigned)__cvta_generic_to_shared(base + i * S);
  if constexpr (S % 2 == 0) {
#pragma unroll
    for (int k = 0; k < S; k += 2) lds(a + 8 * k, r[k], r[k + 1]);
  } else {
#pragma unroll
    for (int k = 0; k < S; ++k) lds(a + 8 * k, r[k]);
  }
#elif defined(EO_JIT_PPT)
#pragma unroll
  for (int k = 0; k < S; ++k) r[k] = base[i * S + k];
#else
  gload_point<S>(base, i, r);
#endif
}
template <int S>
__device__ __forceinline__ void store_point(double* __restrict__ base, long long i, const double* r) {
#if defined(EO_JIT_STAGED)
  const unsigned a = (unsigned)__cvta_generic_to_shared(base + i * S);
  if constexpr (S % 2 == 0) {
#pragma unroll
    for (int k = 0; k < S; k += 2) sts(a + 8 * k, r[k], r[k + 1]);
  } else {
#pragma unroll
    for (int k = 0; k < S; ++k) sts(a + 8 * k, r[k]);
  }
#elif defined(EO_JIT_PPT)
#pragma unroll
  for (int k = 0; k < S; ++k) base[i * S + k] = r[k];
#else
  gstore_point<S>(base, i, r);
#endif
}

// compile-time table helpers over the Spec's static arrays
template <class Spec, int K>
struct op_offset {
  static constexpr int value = op_offset<Spec, K - 1>::value + Spec::op_size(K - 1);
};
template <class Spec>
struct op_offset<Spec, 0> {
  static constexpr int value = 0;
};
template <class Spec, int K>
struct st_offset {
  static constexpr int value = st_offset<Spec, K - 1>::value + Spec::st_size(K - 1);
};
template <class Spec>
struct st_offset<Spec, 0> {
  static constexpr int value = 0;
};
template <class Spec, int K>
struct aux_offset {
  static constexpr int value = aux_offset<Spec, K - 1>::value + Spec::aux_size(K - 1);
};
template <class Spec>
struct aux_offset<Spec, 0> {
  static constexpr int value = 0;
};

template <class Spec, int K, int NK>
struct loader {
  static __device__ __forceinline__ void operands(const eo_jit_args& a, long long i, double* x) {
    load_point<Spec::op_size(K)>(a.operand[K], i, x + op_offset<Spec, K>::value);
    loader<Spec, K + 1, NK>::operands(a, i, x);
  }
  static __device__ __forceinline__ void states(const eo_jit_args& a, long long i, double* s) {
    load_point<Spec::st_size(K)>(a.state[K], i, s + st_offset<Spec, K>::value);
    loader<Spec, K + 1, NK>::states(a, i, s);
  }
  static __device__ __forceinline__ void aux(const eo_jit_args& a, long long i, const double* v) {
    if (a.aux[K]) store_point<Spec::aux_size(K)>(a.aux[K], i, v + aux_offset<Spec, K>::value);
    loader<Spec, K + 1, NK>::aux(a, i, v);
  }
};
template <class Spec, int NK>
struct loader<Spec, NK, NK> {
  static __device__ __forceinline__ void operands(const eo_jit_args&, long long, double*) {}
  static __device__ __forceinline__ void states(const eo_jit_args&, long long, double*) {}
  static __device__ __forceinline__ void aux(const eo_jit_args&, long long, const double*) {}
};

constexpr int at_least_1(int n) { return n > 0 ? n : 1; }

#ifdef EO_JIT_PPT
// wide global accesses of PPT consecutive points per array (run_ppt)
template <class Spec, int K, int NK>
struct ppt_io {
  static __device__ __forceinline__ void load_operands(const eo_jit_args& a, const eo_jit_args& b, long long g) {
    gload_point<Spec::PPT * Spec::op_size(K)>(a.operand[K], g, const_cast<double*>(b.operand[K]));
    ppt_io<Spec, K + 1, NK>::load_operands(a, b, g);
  }
  static __device__ __forceinline__ void load_states(const eo_jit_args& a, const eo_jit_args& b, long long g) {
    gload_point<Spec::PPT * Spec::st_size(K)>(a.state[K], g, const_cast<double*>(b.state[K]));
    ppt_io<Spec, K + 1, NK>::load_states(a, b, g);
  }
  static __device__ __forceinline__ void store_aux(const eo_jit_args& a, const eo_jit_args& b, long long g) {
    if (a.aux[K]) gstore_point<Spec::PPT * Spec::aux_size(K)>(a.aux[K], g, b.aux[K]);
    ppt_io<Spec, K + 1, NK>::store_aux(a, b, g);
  }
};
template <class Spec, int NK>
struct ppt_io<Spec, NK, NK> {
  static __device__ __forceinline__ void load_operands(const eo_jit_args&, const eo_jit_args&, long long) {}
  static __device__ __forceinline__ void load_states(const eo_jit_args&, const eo_jit_args&, long long) {}
  static __device__ __forceinline__ void store_aux(const eo_jit_args&, const eo_jit_args&, long long) {}
};
#endif

// operands of point i: streamed from HBM ...
template <class Spec>
__device__ __forceinline__ void fetch_operands(const eo_jit_args& a, long long i, double* x) {
  loader<Spec, 0, Spec::N_OPERANDS>::operands(a, i, x);
}

#ifdef EO_JIT_FUSED
// ... or tabulated on the fly from the cell's DOF coefficients (tab_core.cuh: the same arithmetic as tab_kernel)
template <class Spec, int K, int NK>
struct tabber {
  static __device__ __forceinline__ void run(const eo_jit_fused_args& f, const tab_tables* sT, long long cell, int q, double* x) {
    constexpr int G = Spec::tab_gdim(K), B = Spec::tab_bs(K), NB = Spec::tab_nb(K), KIND = Spec::tab_kind(K);
    const tab_tables& T = sT[K];
    double w[NB][B], Kinv[G][G];
    tab_load_cell<G, B, NB>(T, f.dofmap[K], f.x_dofmap[K], f.x[K], f.u[K], cell, w, Kinv);
    double val[B], grad[B][G], r[B * G > 4 ? B * G : 4];
    tab_point<G, B, NB>(T, w, Kinv, q, KIND == 0, KIND != 0, val, grad);
    tab_operand<G, B>(KIND, val, grad, r);
#pragma unroll
    for (int c = 0; c < Spec::op_size(K); ++c) x[op_offset<Spec, K>::value + c] = r[c];
    tabber<Spec, K + 1, NK>::run(f, sT, cell, q, x);
  }
};
template <class Spec, int NK>
struct tabber<Spec, NK, NK> {
  static __device__ __forceinline__ void run(const eo_jit_fused_args&, const tab_tables*, long long, int, double*) {}
};

__shared__ tab_tables eo_jit_s_tables[EO_JIT_N_TABLES];

template <class Spec>
__device__ __forceinline__ void fetch_operands(const eo_jit_fused_args& f, long long i, double* x) {
  const long long gp = f.point_offset + i;
  const long long cell = gp / Spec::NQ;
  const int q = int(gp - cell * Spec::NQ);
  tabber<Spec, 0, Spec::N_OPERANDS>::run(f, eo_jit_s_tables, cell, q, x);
}
#endif

// ---- order 0: the value ----------------------------------------------------------------------
template <class Spec, class ARGS>
__device__ __forceinline__ void run0(const ARGS& a, long long i) {
  double x[at_least_1(Spec::NIN)], s[at_least_1(Spec::NST)], y[at_least_1(Spec::NOUT)], w[at_least_1(Spec::NAUX)];
  fetch_operands<Spec>(a, i, x);
  loader<Spec, 0, Spec::N_STATE>::states(a, i, s);
  Spec::template call<double>(x, s, a.prm, y, w);
  store_point<Spec::NOUT>(a.out, i, y);
  loader<Spec, 0, Spec::N_AUX>::aux(a, i, w);
}

// ---- order 1: d y / d operand[A], layout [point][NOUT][size(A)] --------------------------------
template <class Spec, class ARGS>
__device__ __forceinline__ void run1(const ARGS& a, long long i) {
  constexpr int A = Spec::DA, NA = Spec::op_size(A), OA = op_offset<Spec, A>::value;
  using T = eo::dual<NA>;
  double xr[at_least_1(Spec::NIN)], s[at_least_1(Spec::NST)];
  fetch_operands<Spec>(a, i, xr);
  loader<Spec, 0, Spec::N_STATE>::states(a, i, s);
  T x[at_least_1(Spec::NIN)], y[at_least_1(Spec::NOUT)], w[at_least_1(Spec::NAUX)];
#pragma unroll
  for (int k = 0; k < Spec::NIN; ++k) {
    x[k].v = xr[k];
#pragma unroll
    for (int j = 0; j < NA; ++j) x[k].d[j] = (k == OA + j) ? 1.0 : 0.0;
  }
  Spec::template call<T>(x, s, a.prm, y, w);
  double D[Spec::NOUT * NA];
#pragma unroll
  for (int o = 0; o < Spec::NOUT; ++o)
#pragma unroll
    for (int j = 0; j < NA; ++j) D[o * NA + j] = y[o].d[j];
  store_point<Spec::NOUT * NA>(a.out, i, D);
  if (a.value) {
    double yv[at_least_1(Spec::NOUT)];
#pragma unroll
    for (int o = 0; o < Spec::NOUT; ++o) yv[o] = y[o].v;
    store_point<Spec::NOUT>(a.value, i, yv);
  }
  double wv[at_least_1(Spec::NAUX)];
#pragma unroll
  for (int o = 0; o < Spec::NAUX; ++o) wv[o] = w[o].v;
  loader<Spec, 0, Spec::N_AUX>::aux(a, i, wv);
}

// ---- order 2: d2 y / d operand[A] d operand[B] (A <= B), layout [point][NOUT][size(A)][size(B)] --
template <class Spec, class ARGS>
__device__ __forceinline__ void run2(const ARGS& a, long long i) {
  constexpr int A = Spec::DA, B = Spec::DB;
  constexpr int NA = Spec::op_size(A), OA = op_offset<Spec, A>::value;
  constexpr int NB = Spec::op_size(B), OB = op_offset<Spec, B>::value;
  using V = eo::dual<NB>;
  using T = eo::dual<NA, V>;
  double xr[at_least_1(Spec::NIN)], s[at_least_1(Spec::NST)];
  fetch_operands<Spec>(a, i, xr);
  loader<Spec, 0, Spec::N_STATE>::states(a, i, s);
  T x[at_least_1(Spec::NIN)], y[at_least_1(Spec::NOUT)], w[at_least_1(Spec::NAUX)];
#pragma unroll
  for (int k = 0; k < Spec::NIN; ++k) {
    x[k].v.v = xr[k];
#pragma unroll
    for (int j = 0; j < NB; ++j) x[k].v.d[j] = (k == OB + j) ? 1.0 : 0.0;
#pragma unroll
    for (int j = 0; j < NA; ++j) x[k].d[j] = V((k == OA + j) ? 1.0 : 0.0);
  }
  Spec::template call<T>(x, s, a.prm, y, w);
  double D[Spec::NOUT * NA * NB];
#pragma unroll
  for (int o = 0; o < Spec::NOUT; ++o)
#pragma unroll
    for (int ja = 0; ja < NA; ++ja)
#pragma unroll
      for (int jb = 0; jb < NB; ++jb) D[(o * NA + ja) * NB + jb] = y[o].d[ja].d[jb];
  store_point<Spec::NOUT * NA * NB>(a.out, i, D);
  if (a.value) {
    double yv[at_least_1(Spec::NOUT)];
#pragma unroll
    for (int o = 0; o < Spec::NOUT; ++o) yv[o] = y[o].v.v;
    store_point<Spec::NOUT>(a.value, i, yv);
  }
  double wv[at_least_1(Spec::NAUX)];
#pragma unroll
  for (int o = 0; o < Spec::NAUX; ++o) wv[o] = w[o].v.v;
  loader<Spec, 0, Spec::N_AUX>::aux(a, i, wv);
}

template <class Spec, class ARGS>
__device__ __forceinline__ void run(const ARGS& a) {
#ifdef EO_JIT_FUSED
  {  // stage the element tables (a few KB per operand) in shared memory: the point index q is not warp uniform
    const int nw = int(sizeof(tab_tables) / 8);
    for (int k = 0; k < EO_JIT_N_TABLES; ++k) {
      const double* src = reinterpret_cast<const double*>(a.T[k]);
      double* dst = reinterpret_cast<double*>(&eo_jit_s_tables[k]);
      for (int t = threadIdx.x; t < nw; t += blockDim.x) dst[t] = __ldg(src + t);
    }
    __syncthreads();
  }
#endif
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= a.n) return;
  if constexpr (Spec::ORDER == 0)
    run0<Spec, ARGS>(a, i);
  else if constexpr (Spec::ORDER == 1)
    run1<Spec, ARGS>(a, i);
  else
    run2<Spec, ARGS>(a, i);
}

#ifdef EO_JIT_PPT
// ------------------------------------------------------------------------------------------------------------
// Several points per thread, for models with very few bytes per point (scalar fields): a thread fetches PPT
// consecutive points of every array with one wide access each (more bytes in flight per thread - a scalar model at one
// point per thread reaches only 0.77 of the HBM roofline), evaluates them one after the other out of registers and
// writes the PPT results back the same way.  a.n is a multiple of PPT (the host sends the remainder to the direct kernel).
// ------------------------------------------------------------------------------------------------------------
template <class Spec>
__device__ __forceinline__ void run_ppt(const eo_jit_args& a) {
  constexpr int PPT = Spec::PPT;
  constexpr int OUTW = Spec::ORDER == 0 ? Spec::NOUT
                                        : (Spec::ORDER == 1 ? Spec::NOUT * Spec::op_size(Spec::DA)
                                                            : Spec::NOUT * Spec::op_size(Spec::DA) * Spec::op_size(Spec::DB));
  const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (g * PPT >= a.n) return;
  double in[PPT * at_least_1(Spec::NIN)], st_[PPT * at_least_1(Spec::NST)], out[PPT * OUTW], val[PPT * Spec::NOUT],
      aux[PPT * at_least_1(Spec::NAUX)];
  eo_jit_args b = a;
  int off = 0;
#pragma unroll
  for (int k = 0; k < Spec::N_OPERANDS; ++k) {
    b.operand[k] = in + off;
    off += PPT * Spec::op_size(k);
  }
  off = 0;
#pragma unroll
  for (int k = 0; k < Spec::N_STATE; ++k) {
    b.state[k] = st_ + off;
    off += PPT * Spec::st_size(k);
  }
  b.out = out;
  b.value = val;  // always evaluated into registers; written back only when the caller asked for it
  off = 0;
#pragma unroll
  for (int k = 0; k < Spec::N_AUX; ++k) {
    b.aux[k] = aux + off;
    off += PPT * Spec::aux_size(k);
  }
  ppt_io<Spec, 0, Spec::N_OPERANDS>::load_operands(a, b, g);
  ppt_io<Spec, 0, Spec::N_STATE>::load_states(a, b, g);
#pragma unroll
  for (int p = 0; p < PPT; ++p) {
    if constexpr (Spec::ORDER == 0)
      run0<Spec, eo_jit_args>(b, p);
    else if constexpr (Spec::ORDER == 1)
      run1<Spec, eo_jit_args>(b, p);
    else
      run2<Spec, eo_jit_args>(b, p);
  }
  gstore_point<PPT * OUTW>(a.out, g, out);
  if (a.value) gstore_point<PPT * Spec::NOUT>(a.value, g, val);
  ppt_io<Spec, 0, Spec::N_AUX>::store_aux(a, b, g);
}
#endif

#ifdef EO_JIT_STAGED
// ------------------------------------------------------------------------------------------------------------
// Staged variant: one CTA = one tile of TILE points.  Every per-point array of the tile is ONE contiguous block of
// TILE * S * 8 bytes, so it is moved by the TMA unit as a 1-D bulk copy (cp.async.bulk, completion on an mbarrier) -
// fully coalesced whatever S is (odd component counts make per-thread accesses strided: 0.14-0.67 of the HBM
// roofline measured for S = 3, 5, 9, 25).  Threads then read their point from shared memory, and the results go
// back the same way (st.shared -> fence.proxy.async -> cp.async.bulk shared -> global).
// ------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void bulk_g2s(unsigned dst, const void* src, unsigned bytes, unsigned mbar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src),
               "r"(bytes), "r"(mbar)
               : "memory");
}
__device__ __forceinline__ void bulk_s2g(void* dst, unsigned src, unsigned bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(src), "r"(bytes) : "memory");
}

template <class Spec>
__device__ __forceinline__ void run_staged(const eo_jit_args& a) {
  constexpr int P = Spec::TILE;
  constexpr int OUTW = Spec::ORDER == 0 ? Spec::NOUT
                                        : (Spec::ORDER == 1 ? Spec::NOUT * Spec::op_size(Spec::DA)
                                                            : Spec::NOUT * Spec::op_size(Spec::DA) * Spec::op_size(Spec::DB));
  extern __shared__ __align__(128) unsigned char eo_smem[];
  // layout: [mbarrier (128 B)] [operands | state | out | value | aux], each block TILE * S doubles
  const unsigned mbar = (unsigned)__cvta_generic_to_shared(eo_smem);
  double* tile = reinterpret_cast<double*>(eo_smem + 128);
  const long long p0 = (long long)blockIdx.x * P;  // first point of this tile
  eo_jit_args b = a;                               // the same argument block, pointing into shared memory
  b.n = P;
  int off = 0;
  unsigned in_bytes = 0;
#pragma unroll
  for (int k = 0; k < Spec::N_OPERANDS; ++k) {
    b.operand[k] = tile + off;
    off += P * Spec::op_size(k);
    in_bytes += P * Spec::op_size(k) * 8;
  }
#pragma unroll
  for (int k = 0; k < Spec::N_STATE; ++k) {
    b.state[k] = tile + off;
    off += P * Spec::st_size(k);
    in_bytes += P * Spec::st_size(k) * 8;
  }
  b.out = tile + off;
  off += P * OUTW;
  if (a.value) b.value = tile + off, off += P * Spec::NOUT;
#pragma unroll
  for (int k = 0; k < Spec::N_AUX; ++k)
    if (a.aux[k]) b.aux[k] = tile + off, off += P * Spec::aux_size(k);

  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(mbar) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mbar), "r"(in_bytes) : "memory");
#pragma unroll
    for (int k = 0; k < Spec::N_OPERANDS; ++k)
      bulk_g2s((unsigned)__cvta_generic_to_shared(b.operand[k]), a.operand[k] + p0 * Spec::op_size(k), P * Spec::op_size(k) * 8, mbar);
#pragma unroll
    for (int k = 0; k < Spec::N_STATE; ++k)
      bulk_g2s((unsigned)__cvta_generic_to_shared(b.state[k]), a.state[k] + p0 * Spec::st_size(k), P * Spec::st_size(k) * 8, mbar);
  }
  {  // every thread waits for the tile (phase 0 of the barrier)
    unsigned done = 0;
    while (!done)
      asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0; selp.u32 %0, 1, 0, p; }" : "=r"(done) : "r"(mbar) : "memory");
  }
  const long long i = threadIdx.x;
  if constexpr (Spec::ORDER == 0)
    run0<Spec, eo_jit_args>(b, i);
  else if constexpr (Spec::ORDER == 1)
    run1<Spec, eo_jit_args>(b, i);
  else
    run2<Spec, eo_jit_args>(b, i);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy writes -> visible to the bulk-copy engine
  __syncthreads();
  if (threadIdx.x == 0) {
    bulk_s2g(a.out + p0 * OUTW, (unsigned)__cvta_generic_to_shared(b.out), P * OUTW * 8);
    if (a.value) bulk_s2g(a.value + p0 * Spec::NOUT, (unsigned)__cvta_generic_to_shared(b.value), P * Spec::NOUT * 8);
#pragma unroll
    for (int k = 0; k < Spec::N_AUX; ++k)
      if (a.aux[k]) bulk_s2g(a.aux[k] + p0 * Spec::aux_size(k), (unsigned)__cvta_generic_to_shared(b.aux[k]), P * Spec::aux_size(k) * 8);
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");  // shared memory must outlive the copies
  }
}
#endif

}  // namespace eo_jitd
