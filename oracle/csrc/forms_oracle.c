/* TEST INFRASTRUCTURE (CPU baseline of bench.py --model tab|fused|step|action, checked against oracle/tabulation.py and
 * oracle/forms.py in tests/test_forms_cpu.py) - the cell loop the reference runs through DOLFINx for the von Mises demo,
 * restated in C with OpenMP over cells so that the CPU figures beside the GPU kernels use every host core:
 *   tabulation of the Mandel strain   fem.Expression.eval, external_operator.py:393-402 (textbook affine-simplex algorithm,
 *                                     SURVEY.md appendix B; parity unpinned like oracle/tabulation.py)
 *   radial return                     demo_plasticity_von_mises.py:298-332 (the same statement sequence as vm_heat_oracle.c)
 *   residual                          assemble_vector(inner(sigma, epsilon(v)) dx), demo_vm:253, petsc/petsc.py:64
 *   tangent action                    assemble_matrix(inner(C_tang epsilon(u_hat), epsilon(v)) dx) applied to a vector
 * 2-d vector Lagrange element with nb <= 10 basis functions, nq <= 16 points per cell.
 * mode: 0 strain only, 1 + radial return (C_tang, sigma, dp), 2 + residual b, 3 tangent action y = A xin (C_tang given). */
#include <math.h>
#include <stdint.h>
#include <string.h>

typedef struct {
  double lmbda, mu, H, sigma_0;
} forms_vm_params;

void oracle_vm_point(const forms_vm_params* q, const double* deps, const double* sn, double p, double* Ct, double* sig,
                     double* dp_out);

static void cell_geometry(const double* x, const int32_t* xd, const double* dpsi, double K[2][2], double* adet) {
  double J[2][2];
  for (int i = 0; i < 2; ++i)
    for (int j = 0; j < 2; ++j) {
      double acc = 0.0;
      for (int v = 0; v < 3; ++v) acc += x[3 * (int64_t)xd[v] + i] * dpsi[j * 3 + v];
      J[i][j] = acc;
    }
  const double det = J[0][0] * J[1][1] - J[0][1] * J[1][0];
  K[0][0] = J[1][1] / det, K[0][1] = -J[0][1] / det, K[1][0] = -J[1][0] / det, K[1][1] = J[0][0] / det;
  *adet = fabs(det);
}

void oracle_forms_p2_cells(int mode, int nb, int nq, const double* dphi /* [2][nq][nb] */, const double* dpsi /* [2][3] */,
                           const double* weights, const int32_t* dofmap, const int32_t* x_dofmap, const double* x,
                           const double* u, const forms_vm_params* prm, const double* sigma_n, const double* p,
                           double* strain, double* C_tang, double* sigma, double* dp, double* out_vec, int64_t n_cells) {
  const double r2 = sqrt(2.0) * 0.5;
#pragma omp parallel for schedule(static)
  for (int64_t c = 0; c < n_cells; ++c) {
    double K[2][2], adet, w[10][2], g[10][2], fe[10][2];
    cell_geometry(x, x_dofmap + 3 * c, dpsi, K, &adet);
    for (int a = 0; a < nb; ++a) {
      const int64_t d = dofmap[c * nb + a];
      w[a][0] = u[2 * d], w[a][1] = u[2 * d + 1];
      fe[a][0] = fe[a][1] = 0.0;
    }
    for (int q = 0; q < nq; ++q) {
      const int64_t i = c * nq + q;
      double G[2][2] = {{0, 0}, {0, 0}};
      for (int a = 0; a < nb; ++a) { /* physical gradient of phi_a, and grad u */
        const double d0 = dphi[(0 * nq + q) * nb + a], d1 = dphi[(1 * nq + q) * nb + a];
        g[a][0] = d0 * K[0][0] + d1 * K[1][0];
        g[a][1] = d0 * K[0][1] + d1 * K[1][1];
        for (int k = 0; k < 2; ++k) G[k][0] += w[a][k] * g[a][0], G[k][1] += w[a][k] * g[a][1];
      }
      double e[4] = {G[0][0], G[1][1], 0.0, r2 * (G[0][1] + G[1][0])}; /* demo_vm:225-227 */
      double tau[4];
      if (mode == 0) {
        memcpy(strain + 4 * i, e, sizeof(e));
        continue;
      }
      if (mode == 3) {
        const double* Ct = C_tang + 16 * i;
        for (int r = 0; r < 4; ++r) tau[r] = Ct[4 * r] * e[0] + Ct[4 * r + 1] * e[1] + Ct[4 * r + 2] * e[2] + Ct[4 * r + 3] * e[3];
      } else {
        oracle_vm_point(prm, e, sigma_n + 4 * i, p[i], C_tang + 16 * i, sigma + 4 * i, dp + i);
        if (mode == 1) continue;
        memcpy(tau, sigma + 4 * i, sizeof(tau));
      }
      const double s = weights[q] * adet;
      for (int a = 0; a < nb; ++a) { /* tau . epsilon(phi_a e_k) */
        fe[a][0] += s * (tau[0] * g[a][0] + r2 * tau[3] * g[a][1]);
        fe[a][1] += s * (tau[1] * g[a][1] + r2 * tau[3] * g[a][0]);
      }
    }
    if (mode >= 2)
      for (int a = 0; a < nb; ++a) {
        const int64_t d = dofmap[c * nb + a];
#pragma omp atomic
        out_vec[2 * d] += fe[a][0];
#pragma omp atomic
        out_vec[2 * d + 1] += fe[a][1];
      }
  }
}
