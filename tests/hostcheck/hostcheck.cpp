// TEST-ONLY: compiles the host/device "core" headers of the CUDA kernels with g++ so that their
// algebra can be checked against the oracle on a machine without a GPU.  Never loaded by the product.
#include <cstdint>

#include "mc_core.cuh"
#include "tab_core.cuh"
#include "form_core.cuh"
#include "isihara_core.cuh"
#include "vm_core.cuh"

template <int GDIM, int BS, int NB>
static void tab_cells(const tab_tables& T, int kind, const int32_t* dofmap, const int32_t* x_dofmap, const double* x,
                      const double* u, int64_t n_cells, double* out) {
  const int ncomp = tab_ncomp(kind, BS, GDIM);
  for (int64_t c = 0; c < n_cells; ++c) {
    double w[NB][BS], xv[GDIM + 1][GDIM], K[GDIM][GDIM];
    for (int a = 0; a < NB; ++a)
      for (int k = 0; k < BS; ++k) w[a][k] = u[int64_t(BS) * dofmap[c * NB + a] + k];
    for (int v = 0; v < GDIM + 1; ++v)
      for (int i = 0; i < GDIM; ++i) xv[v][i] = x[3 * int64_t(x_dofmap[c * (GDIM + 1) + v]) + i];
    tab_geometry<GDIM>(T, xv, K);
    for (int q = 0; q < T.nq; ++q) {
      double val[BS], grad[BS][GDIM], r[16];
      tab_point<GDIM, BS, NB>(T, w, K, q, kind == 0, kind != 0, val, grad);
      tab_operand<GDIM, BS>(kind, val, grad, r);
      for (int k = 0; k < ncomp; ++k) out[(c * T.nq + q) * ncomp + k] = r[k];
    }
  }
}

// serial restatement of the form kernels' cell loop on top of form_core.cuh:
//   x == nullptr: y += B^T W D            (eo_form_vector; D = point values, ncomp(kind_test) per point)
//   x != nullptr: y += B^T W D (B x)      (eo_form_action; D row-major (ncomp_test, ncomp_trial) per point)
template <int GDIM, int BS, int NB>
static void form_cells(const tab_tables& T, const double* wq, int kind_test, int kind_trial, const int32_t* dofmap,
                       const int32_t* x_dofmap, const double* xg, const double* D, const double* x, int64_t n_cells,
                       double* y) {
  const int nt = tab_ncomp(kind_test, BS, GDIM), ni = x ? tab_ncomp(kind_trial, BS, GDIM) : 1;
  for (int64_t c = 0; c < n_cells; ++c) {
    double w[NB][BS], xv[GDIM + 1][GDIM], K[GDIM][GDIM], fe[NB][BS];
    for (int a = 0; a < NB; ++a)
      for (int k = 0; k < BS; ++k) {
        w[a][k] = x ? x[int64_t(BS) * dofmap[c * NB + a] + k] : 0.0;
        fe[a][k] = 0.0;
      }
    for (int v = 0; v < GDIM + 1; ++v)
      for (int i = 0; i < GDIM; ++i) xv[v][i] = xg[3 * int64_t(x_dofmap[c * (GDIM + 1) + v]) + i];
    const double adet = form_geometry_xv<GDIM>(T, xv, K);
    for (int q = 0; q < T.nq; ++q) {
      double tau[16];
      const double* Dq = D + (c * T.nq + q) * int64_t(nt * ni);
      if (x) {
        double val[BS], grad[BS][GDIM], e[16];
        tab_point<GDIM, BS, NB>(T, w, K, q, kind_trial == 0, kind_trial != 0, val, grad);
        tab_operand<GDIM, BS>(kind_trial, val, grad, e);
        for (int r = 0; r < nt; ++r) {
          double acc = 0.0;
          for (int l = 0; l < ni; ++l) acc += Dq[r * ni + l] * e[l];
          tau[r] = acc;
        }
      } else {
        for (int r = 0; r < nt; ++r) tau[r] = Dq[r];
      }
      double Vs[BS], Gs[BS][GDIM];
      form_cotangent<GDIM, BS>(kind_test, tau, Vs, Gs);
      form_accumulate<GDIM, BS, NB>(T, kind_test, q, wq[q] * adet, Vs, Gs, K, fe);
    }
    for (int a = 0; a < NB; ++a)
      for (int k = 0; k < BS; ++k) y[int64_t(BS) * dofmap[c * NB + a] + k] += fe[a][k];
  }
}

static void fill_tables(tab_tables& T, int gdim, int bs, int nb, int nq, const double* phi, const double* dphi,
                        const double* dpsi) {
  T.nb = nb, T.nq = nq, T.bs = bs, T.gdim = gdim, T.nv = gdim + 1;
  for (int q = 0; q < nq; ++q)
    for (int a = 0; a < nb; ++a) {
      T.phi[q][a] = phi[q * nb + a];
      for (int k = 0; k < gdim; ++k) T.dphi[k][q][a] = dphi[(k * nq + q) * nb + a];
    }
  for (int k = 0; k < gdim; ++k)
    for (int v = 0; v < gdim + 1; ++v) T.dpsi[k][v] = dpsi[k * (gdim + 1) + v];
}

extern "C" {

// y (bs * n_dofs, zeroed by the caller) += form vector (x == NULL) or form action; kinds as eo_operand_kind with DEF_GRAD
// already mapped to GRAD; returns 0 or -1 (unsupported element)
int hostcheck_form(int gdim, int bs, int nb, int nq, int kind_test, int kind_trial, const double* phi, const double* dphi,
                   const double* dpsi, const double* weights, const int32_t* dofmap, const int32_t* x_dofmap,
                   const double* xg, const double* D, const double* x, int64_t n_cells, double* y) {
  tab_tables T{};
  fill_tables(T, gdim, bs, nb, nq, phi, dphi, dpsi);
  if (gdim == 2 && bs == 2 && nb == 6) form_cells<2, 2, 6>(T, weights, kind_test, kind_trial, dofmap, x_dofmap, xg, D, x, n_cells, y);
  else if (gdim == 2 && bs == 1 && nb == 3) form_cells<2, 1, 3>(T, weights, kind_test, kind_trial, dofmap, x_dofmap, xg, D, x, n_cells, y);
  else if (gdim == 2 && bs == 1 && nb == 6) form_cells<2, 1, 6>(T, weights, kind_test, kind_trial, dofmap, x_dofmap, xg, D, x, n_cells, y);
  else if (gdim == 2 && bs == 2 && nb == 10) form_cells<2, 2, 10>(T, weights, kind_test, kind_trial, dofmap, x_dofmap, xg, D, x, n_cells, y);
  else if (gdim == 2 && bs == 1 && nb == 10) form_cells<2, 1, 10>(T, weights, kind_test, kind_trial, dofmap, x_dofmap, xg, D, x, n_cells, y);
  else if (gdim == 3 && bs == 3 && nb == 4) form_cells<3, 3, 4>(T, weights, kind_test, kind_trial, dofmap, x_dofmap, xg, D, x, n_cells, y);
  else if (gdim == 3 && bs == 3 && nb == 10) form_cells<3, 3, 10>(T, weights, kind_test, kind_trial, dofmap, x_dofmap, xg, D, x, n_cells, y);
  else if (gdim == 3 && bs == 1 && nb == 10) form_cells<3, 1, 10>(T, weights, kind_test, kind_trial, dofmap, x_dofmap, xg, D, x, n_cells, y);
  else return -1;
  return 0;
}

// von Mises radial return of vm_core.cuh: exact != 0 = the reference's statement sequence (vm_point), else the
// few-division / FMA form (vm_point_fast).  T6[n][6] = the tangent's factors (v, cn, cd); tau[n][4] = vm_factored_apply of
// those factors on e_probe[n][4]
void hostcheck_vm(double lmbda, double mu, double H, double sigma_0, const double* deps, const double* sigma_n, const double* p,
                  int64_t n, int exact, double* C_tang, double* sigma, double* dp, double* T6, const double* e_probe,
                  double* tau) {
  const vm_consts q{lmbda, mu, H, sigma_0};
  for (int64_t i = 0; i < n; ++i) {
    vm_point_out o;
    const double *e = deps + 4 * i, *s = sigma_n + 4 * i;
    if (exact)
      vm_point(q, e[0], e[1], e[2], e[3], s[0], s[1], s[2], s[3], p[i], o);
    else
      vm_point_fast(q, e[0], e[1], e[2], e[3], s[0], s[1], s[2], s[3], p[i], o);
    for (int k = 0; k < 16; ++k) C_tang[16 * i + k] = o.C[k];
    for (int k = 0; k < 4; ++k) sigma[4 * i + k] = o.g[k], T6[6 * i + k] = o.v[k];
    dp[i] = o.dp, T6[6 * i + 4] = o.cn, T6[6 * i + 5] = o.cd;
    vm_factored_apply(q, o.v, o.cn, o.cd, e_probe + 4 * i, tau + 4 * i);
  }
}

void hostcheck_isihara(const isi_weights* w, const double* F, double* dP, double* P, int64_t n) {
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < n; ++i) {
    alignas(8) float zs[2 * ISI_NH];
    isi_point(*w, F + 4 * i, P + 4 * i, dP + 16 * i, zs, 2);
  }
}

// tables: phi [nq][nb], dphi [gdim][nq][nb], dpsi [gdim][gdim+1]; returns 0 or -1 (unsupported element)
int hostcheck_tab(int gdim, int bs, int nb, int nq, int kind, const double* phi, const double* dphi, const double* dpsi,
                  const int32_t* dofmap, const int32_t* x_dofmap, const double* x, const double* u, int64_t n_cells,
                  double* out) {
  tab_tables T{};
  T.nb = nb, T.nq = nq, T.bs = bs, T.gdim = gdim, T.nv = gdim + 1;
  for (int q = 0; q < nq; ++q)
    for (int a = 0; a < nb; ++a) {
      T.phi[q][a] = phi[q * nb + a];
      for (int k = 0; k < gdim; ++k) T.dphi[k][q][a] = dphi[(k * nq + q) * nb + a];
    }
  for (int k = 0; k < gdim; ++k)
    for (int v = 0; v < gdim + 1; ++v) T.dpsi[k][v] = dpsi[k * (gdim + 1) + v];
  if (gdim == 2 && bs == 2 && nb == 6) tab_cells<2, 2, 6>(T, kind, dofmap, x_dofmap, x, u, n_cells, out);
  else if (gdim == 2 && bs == 1 && nb == 3) tab_cells<2, 1, 3>(T, kind, dofmap, x_dofmap, x, u, n_cells, out);
  else if (gdim == 2 && bs == 2 && nb == 3) tab_cells<2, 2, 3>(T, kind, dofmap, x_dofmap, x, u, n_cells, out);
  else if (gdim == 2 && bs == 2 && nb == 10) tab_cells<2, 2, 10>(T, kind, dofmap, x_dofmap, x, u, n_cells, out);
  else if (gdim == 2 && bs == 1 && nb == 10) tab_cells<2, 1, 10>(T, kind, dofmap, x_dofmap, x, u, n_cells, out);
  else if (gdim == 3 && bs == 3 && nb == 4) tab_cells<3, 3, 4>(T, kind, dofmap, x_dofmap, x, u, n_cells, out);
  else if (gdim == 3 && bs == 3 && nb == 10) tab_cells<3, 3, 10>(T, kind, dofmap, x_dofmap, x, u, n_cells, out);
  else if (gdim == 3 && bs == 1 && nb == 10) tab_cells<3, 1, 10>(T, kind, dofmap, x_dofmap, x, u, n_cells, out);
  else return -1;
  return 0;
}

void hostcheck_mc(const mc_params_in* prm, const double* deps, const double* sigma_n, double* C_tang, double* sigma,
                  int32_t* niter, double* yielding, double* norm_res, double* dlambda, int64_t n) {
  mc_consts k;
  mc_make_consts(*prm, k);
#pragma omp parallel for schedule(dynamic, 64)
  for (int64_t i = 0; i < n; ++i)
    mc_point(k, deps + 4 * i, sigma_n + 4 * i, C_tang + 16 * i, sigma + 4 * i, niter[i], yielding[i], norm_res[i],
             dlambda[i]);
}
}
