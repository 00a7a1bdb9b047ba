// tab_core.cuh - per-cell arithmetic of the operand tabulation (hot path (a)): what
// `fem.Expression(operand, eval_points).eval(mesh, entities)` computes at
// src/dolfinx_external_operator/external_operator.py:393-402 for the operands of the reference demos.
// The arithmetic itself lives in un-vendored third-party code (DOLFINx 0.10 tabulate_expression + the
// FFCx-generated kernel + basix tables); this is the standard affine-simplex algorithm (SURVEY.md
// appendix B), with the basis tables as INPUTS (never hard-coded):
//   gather   w[a][c] = u[bs * dofmap[cell][a] + c]                      (blocked dof layout, :18-26)
//   geometry J[i][j] = sum_v x[x_dofmap[cell][v]][i] * dpsi[j][v]       (affine: constant per cell), K = J^-1
//   value    f[c]    = sum_a w[a][c] * phi[q][a]
//   gradient G[c][k] = sum_a w[a][c] * dphi[k][q][a],   grad[c][j] = sum_k G[c][k] * K[k][j]
// followed by the UFL expression on top (eo_operand_kind).  Host/device header: the CUDA kernels in
// tab.cu and the CPU test harness (tests/hostcheck/) compile the same source.
#pragma once

#ifdef __CUDACC_RTC__  // NVRTC has no standard headers (this file is also handed to it for the fused generic path)
typedef int int32_t;
typedef long long int64_t;
#else
#include <cmath>
#include <cstdint>
#endif

#if defined(__CUDACC__)
#define EO_TAB_HD __host__ __device__ __forceinline__
#else
#define EO_TAB_HD inline
#endif

#define EO_TAB_MAX_NB 10  // basis functions per cell (P3 triangle = 10, P2 tetrahedron = 10)
#define EO_TAB_MAX_NQ 16  // evaluation points per cell
#define EO_TAB_MAX_BS 3

// element tables, in constant memory on the device
struct tab_tables {
  int32_t nb, nq, bs, gdim, nv;                          // gdim == tdim (affine simplex), nv = gdim + 1
  double phi[EO_TAB_MAX_NQ][EO_TAB_MAX_NB];              // basix tabulate(0, X)[0]        (nq, nb)
  double dphi[3][EO_TAB_MAX_NQ][EO_TAB_MAX_NB];          // basix tabulate(1, X)[1 + k]    (tdim, nq, nb)
  double dpsi[3][4];                                     // P1 geometry element derivatives (tdim, nv)
};

// inverse of the affine Jacobian
template <int GDIM>
EO_TAB_HD void tab_inverse(const double J[GDIM][GDIM], double K[GDIM][GDIM]) {
  if constexpr (GDIM == 2) {
    const double idet = 1.0 / (J[0][0] * J[1][1] - J[0][1] * J[1][0]);
    K[0][0] = J[1][1] * idet;
    K[0][1] = -J[0][1] * idet;
    K[1][0] = -J[1][0] * idet;
    K[1][1] = J[0][0] * idet;
  } else {
    const double c00 = J[1][1] * J[2][2] - J[1][2] * J[2][1];
    const double c10 = J[1][2] * J[2][0] - J[1][0] * J[2][2];
    const double c20 = J[1][0] * J[2][1] - J[1][1] * J[2][0];
    const double idet = 1.0 / (J[0][0] * c00 + J[0][1] * c10 + J[0][2] * c20);
    K[0][0] = c00 * idet;
    K[1][0] = c10 * idet;
    K[2][0] = c20 * idet;
    K[0][1] = (J[0][2] * J[2][1] - J[0][1] * J[2][2]) * idet;
    K[1][1] = (J[0][0] * J[2][2] - J[0][2] * J[2][0]) * idet;
    K[2][1] = (J[0][1] * J[2][0] - J[0][0] * J[2][1]) * idet;
    K[0][2] = (J[0][1] * J[1][2] - J[0][2] * J[1][1]) * idet;
    K[1][2] = (J[0][2] * J[1][0] - J[0][0] * J[1][2]) * idet;
    K[2][2] = (J[0][0] * J[1][1] - J[0][1] * J[1][0]) * idet;
  }
}

// K = J^-1 of one cell from its vertex coordinates xv[v][i]
template <int GDIM>
EO_TAB_HD void tab_geometry(const tab_tables& T, const double xv[GDIM + 1][GDIM], double K[GDIM][GDIM]) {
  double J[GDIM][GDIM];
#pragma unroll
  for (int i = 0; i < GDIM; ++i)
#pragma unroll
    for (int j = 0; j < GDIM; ++j) {
      double acc = 0.0;
#pragma unroll
      for (int v = 0; v < GDIM + 1; ++v) acc += xv[v][i] * T.dpsi[j][v];
      J[i][j] = acc;
    }
  tab_inverse<GDIM>(J, K);
}

// acc + a * b, as two roundings (the files that include this header are compiled with -fmad=false so that the
// tabulation kernels agree bit for bit among themselves) or as one fused operation (FMA = true: consumers that are
// compared at a tolerance only - the form kernels' action / vector / matrix integrals)
template <bool FMA>
EO_TAB_HD double tab_madd(double a, double b, double acc) {
  if constexpr (FMA)
    return fma(a, b, acc);
  else
    return acc + a * b;
}

// function value and physical gradient at evaluation point q:  w[a][c] gathered coefficients
template <int GDIM, int BS, int NB, bool FMA = false>
EO_TAB_HD void tab_point(const tab_tables& T, const double w[NB][BS], const double K[GDIM][GDIM], int q, bool want_value,
                         bool want_grad, double val[BS], double grad[BS][GDIM]) {
  if (want_value) {
#pragma unroll
    for (int c = 0; c < BS; ++c) {
      double acc = 0.0;
#pragma unroll
      for (int a = 0; a < NB; ++a) acc = tab_madd<FMA>(w[a][c], T.phi[q][a], acc);
      val[c] = acc;
    }
  }
  if (want_grad) {
    double G[BS][GDIM];
#pragma unroll
    for (int c = 0; c < BS; ++c)
#pragma unroll
      for (int k = 0; k < GDIM; ++k) {
        double acc = 0.0;
#pragma unroll
        for (int a = 0; a < NB; ++a) acc = tab_madd<FMA>(w[a][c], T.dphi[k][q][a], acc);
        G[c][k] = acc;
      }
#pragma unroll
    for (int c = 0; c < BS; ++c)
#pragma unroll
      for (int j = 0; j < GDIM; ++j) {
        double acc = 0.0;
#pragma unroll
        for (int k = 0; k < GDIM; ++k) acc = tab_madd<FMA>(G[c][k], K[k][j], acc);
        grad[c][j] = acc;
      }
  }
}

// the UFL expression on top.  kinds (include/eo_b200.h eo_operand_kind):
//   0 VALUE          f                                  (T,                      part1.py:210, part2.py:136)
//   1 GRAD           grad f, row-major (bs, gdim)       (sigma = grad T,         part2.py:137,167)
//   2 MANDEL_STRAIN  [g00, g11, 0, sqrt(2)/2 (g01+g10)] (epsilon(Du), demo_vm:225-227, demo_mc:148-157)
//   3 DEF_GRAD       I + grad u, row-major              (F,                      demo_hyperelasticity.py:479)
EO_TAB_HD int tab_ncomp(int kind, int bs, int gdim) {
  return kind == 0 ? bs : (kind == 2 ? 4 : bs * gdim);
}

template <int GDIM, int BS>
EO_TAB_HD void tab_operand(int kind, const double val[BS], const double grad[BS][GDIM], double* out) {
  if (kind == 0) {
#pragma unroll
    for (int c = 0; c < BS; ++c) out[c] = val[c];
  } else if (kind == 2) {
    if constexpr (GDIM == 2 && BS == 2) {
      out[0] = grad[0][0];
      out[1] = grad[1][1];
      out[2] = 0.0;
      out[3] = 1.4142135623730951 * 0.5 * (grad[0][1] + grad[1][0]);  // np.sqrt(2.0) * 0.5 * (...)
    }
  } else {
#pragma unroll
    for (int c = 0; c < BS; ++c)
#pragma unroll
      for (int j = 0; j < GDIM; ++j) out[c * GDIM + j] = grad[c][j] + ((kind == 3 && c == j) ? 1.0 : 0.0);
  }
}

#if defined(__CUDACC__) || defined(__CUDACC_RTC__)
// the cell's row of the dofmap (NB even: rows are 8-byte aligned, two indices per load)
template <int NB>
__device__ __forceinline__ void tab_load_idx(const int32_t* __restrict__ dofmap, int64_t c, int32_t idx[NB]) {
  if constexpr (NB % 2 == 0) {
    const int2* row = reinterpret_cast<const int2*>(dofmap + c * NB);
#pragma unroll
    for (int a = 0; a < NB / 2; ++a) {
      const int2 v = __ldg(row + a);
      idx[2 * a] = v.x, idx[2 * a + 1] = v.y;
    }
  } else {
#pragma unroll
    for (int a = 0; a < NB; ++a) idx[a] = __ldg(dofmap + c * NB + a);
  }
}

// the coefficients of the dofs idx[] (blocked layout)
template <int BS, int NB>
__device__ __forceinline__ void tab_gather_idx(const double* __restrict__ u, const int32_t idx[NB], double w[NB][BS]) {
#pragma unroll
  for (int a = 0; a < NB; ++a) {
    if constexpr (BS == 2) {
      const double2 v = __ldg(reinterpret_cast<const double2*>(u) + idx[a]);
      w[a][0] = v.x, w[a][1] = v.y;
    } else {
#pragma unroll
      for (int k = 0; k < BS; ++k) w[a][k] = __ldg(u + int64_t(BS) * idx[a] + k);
    }
  }
}

template <int BS, int NB>
__device__ __forceinline__ void tab_gather(const int32_t* __restrict__ dofmap, const double* __restrict__ u, int64_t c,
                                           double w[NB][BS]) {
  int32_t idx[NB];
  tab_load_idx<NB>(dofmap, c, idx);
  tab_gather_idx<BS, NB>(u, idx, w);
}

// gather the cell's coefficients and inverse Jacobian (read-only path; neighbouring cells share nodes: L1/L2 hits)
template <int GDIM, int BS, int NB>
__device__ __forceinline__ void tab_load_cell(const tab_tables& T, const int32_t* __restrict__ dofmap,
                                              const int32_t* __restrict__ x_dofmap, const double* __restrict__ x,
                                              const double* __restrict__ u, int64_t c, double w[NB][BS],
                                              double K[GDIM][GDIM]) {
  tab_gather<BS, NB>(dofmap, u, c, w);
  double xv[GDIM + 1][GDIM];
#pragma unroll
  for (int v = 0; v < GDIM + 1; ++v) {
    const int32_t node = __ldg(x_dofmap + c * (GDIM + 1) + v);
#pragma unroll
    for (int i = 0; i < GDIM; ++i) xv[v][i] = __ldg(x + 3 * int64_t(node) + i);
  }
  tab_geometry<GDIM>(T, xv, K);
}
#endif
