// form_core.cuh - per-cell arithmetic of the device-side consumers (csrc/form.cu): the weighted TRANSPOSE of the operand
// tabulation of tab_core.cuh.  With B the tabulation operator (u -> OP(u) at the points) and W = diag(w_q |det J|),
//   residual        b = B^T W s          assemble_vector(inner(s, OP(v)) dx)            petsc/petsc.py:64, demo_vm:253
//   tangent action  y = B^T W D B x      the action of assemble_matrix(inner(D OP(u_hat), OP(v)) dx)   petsc/petsc.py:88
// Host/device header like tab_core.cuh: the CUDA kernels in form.cu and the CPU test harness (tests/hostcheck/) compile
// the same source.
#pragma once
#include "tab_core.cuh"

// inverse Jacobian AND |det J| of one affine cell from its vertex coordinates xv[v][i]
template <int GDIM>
EO_TAB_HD double form_geometry_xv(const tab_tables& T, const double xv[GDIM + 1][GDIM], double K[GDIM][GDIM]) {
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
  double det;
  if constexpr (GDIM == 2)
    det = J[0][0] * J[1][1] - J[0][1] * J[1][0];
  else
    det = J[0][0] * (J[1][1] * J[2][2] - J[1][2] * J[2][1]) + J[0][1] * (J[1][2] * J[2][0] - J[1][0] * J[2][2]) +
          J[0][2] * (J[1][0] * J[2][1] - J[1][1] * J[2][0]);
  return det < 0.0 ? -det : det;
}

// transpose of tab_operand: the point value s (ncomp of `kind`) as a cotangent of (value, gradient)
template <int GDIM, int BS>
EO_TAB_HD void form_cotangent(int kind, const double* s, double Vs[BS], double Gs[BS][GDIM]) {
#pragma unroll
  for (int c = 0; c < BS; ++c) {
    Vs[c] = 0.0;
#pragma unroll
    for (int j = 0; j < GDIM; ++j) Gs[c][j] = 0.0;
  }
  if (kind == 0) {
#pragma unroll
    for (int c = 0; c < BS; ++c) Vs[c] = s[c];
  } else if (kind == 2) {
    if constexpr (GDIM == 2 && BS == 2) {
      const double h = 1.4142135623730951 * 0.5 * s[3];
      Gs[0][0] = s[0], Gs[1][1] = s[1], Gs[0][1] = h, Gs[1][0] = h;
    }
  } else {
#pragma unroll
    for (int c = 0; c < BS; ++c)
#pragma unroll
      for (int j = 0; j < GDIM; ++j) Gs[c][j] = s[c * GDIM + j];
  }
}

// table access for either home of the tables (constant bank: warp-uniform q; shared memory: per-thread q)
EO_TAB_HD double form_phi(const tab_tables& T, int q, int a) { return T.phi[q][a]; }
EO_TAB_HD double form_dphi(const tab_tables& T, int k, int q, int a) { return T.dphi[k][q][a]; }
template <int GDIM, int NB>
struct form_tabs;
template <int GDIM, int NB>
EO_TAB_HD double form_phi(const form_tabs<GDIM, NB>& S, int q, int a);
template <int GDIM, int NB>
EO_TAB_HD double form_dphi(const form_tabs<GDIM, NB>& S, int k, int q, int a);

// fe[a][c] += scale * ( Vs[c] phi[q][a] + sum_k (sum_j Gs[c][j] K[k][j]) dphi[k][q][a] )   - transpose of tab_point
template <int GDIM, int BS, int NB, bool FMA = false, class Tables>
EO_TAB_HD void form_accumulate(const Tables& T, int kind, int q, double scale, const double Vs[BS],
                                                const double Gs[BS][GDIM], const double K[GDIM][GDIM],
                                                double fe[NB][BS]) {
  if (kind == 0) {
#pragma unroll
    for (int a = 0; a < NB; ++a) {
      const double ph = scale * form_phi(T, q, a);
#pragma unroll
      for (int c = 0; c < BS; ++c) fe[a][c] = tab_madd<FMA>(Vs[c], ph, fe[a][c]);
    }
    return;
  }
  double H[BS][GDIM];
#pragma unroll
  for (int c = 0; c < BS; ++c)
#pragma unroll
    for (int k = 0; k < GDIM; ++k) {
      double acc = 0.0;
#pragma unroll
      for (int j = 0; j < GDIM; ++j) acc = tab_madd<FMA>(Gs[c][j], K[k][j], acc);
      H[c][k] = scale * acc;
    }
#pragma unroll
  for (int a = 0; a < NB; ++a)
#pragma unroll
    for (int c = 0; c < BS; ++c) {
      double acc = 0.0;
#pragma unroll
      for (int k = 0; k < GDIM; ++k) acc = tab_madd<FMA>(H[c][k], form_dphi(T, k, q, a), acc);
      fe[a][c] += acc;
    }
}

