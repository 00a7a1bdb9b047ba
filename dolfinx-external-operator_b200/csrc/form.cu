// form.cu - device-side CONSUMERS of the external-operator values: the two integrals the
// reference hands to DOLFINx right after `evaluate_external_operators`, evaluated where the stress and the
// tangent already live (HBM), so that only DOF vectors cross the PCIe link instead of 168 B per point.
// replaces: `assemble_vector(b, F)` with F = inner(sigma, epsilon(v)) dx         (demo_plasticity_von_mises.py:253,
//             petsc/petsc.py:64; heat: inner(q, grad(v)) dx, part2.py:181)
//           the ACTION of `assemble_matrix(A, J)` with J = derivative(F, Du, u_hat) =
//             inner(C_tang epsilon(u_hat), epsilon(v)) dx                         (demo_vm:390-398, petsc/petsc.py:88)
//           on a vector (what a Krylov method needs from A), and A itself as CSR values.
// The test/trial "operand kinds" are the ones of the tabulation (tab_core.cuh): the residual is the weighted
// TRANSPOSE of the operand tabulation, b = B^T W s, and the tangent action is y = B^T W D B x.
//
// Kernels (P2 vector triangle, 3 points per cell, measured at 1e8 points on B200 - profiles/r1_forms_ncu_summary.md):
//   form_vector_cell_kernel                    one thread per CELL, nb*bs fp64 RED.ADDs                            (0.68)
//   form_vm_step_kernel                        one thread per POINT, per-CTA shared-memory reduction of the nq
//                                              contributions, one fp64 RED.ADD per element-vector entry   (0.68)
//   form_action_tma_kernel                     one thread per CELL, its nq 4x4 tangents staged by a per-thread TMA bulk copy
//                                              into a padded shared row                                   (0.81-0.82)
//   form_action_cell_kernel                    general fallback of the action (any kinds / nq / element)
//   form_matrix_kernel                         element matrices column by column into CSR values (bisection in the row)
// Each DOF of a P2 triangle mesh is touched by 2-6 cells: the REDs see low contention and resolve in L2.  Summation
// order over cells is not fixed: results agree with the oracle to rounding (tests: 1e-12 of the vector scale), not bit
// for bit.  What bounds the register-path kernels is the LSU wavefront rate, not HBM (see the table in the summary).
// This file is compiled with -fmad=false for the step kernel (its per-point results are bit-identical to tab.cu's fused
// kernel); the vector / action / matrix integrals, compared at a tolerance only, contract with explicit FMAs.
// The per-cell arithmetic lives in form_core.cuh (host/device; also compiled by the CPU test harness).
#include "eo_common.cuh"
#include "tab_core.cuh"
#include "form_core.cuh"
#include "tab_handle.cuh"
#include "vm_core.cuh"

#include <cstdlib>

struct eo_form {
  eo_tab* tab = nullptr;
  eo_ctx* ctx = nullptr;     // = tab->ctx, kept so that destroying the form never touches the (possibly gone) eo_tab
  double w[EO_TAB_MAX_NQ];   // quadrature weights on the reference cell
  double* x_stage = nullptr; // device copy of a host input vector
  double* y_stage = nullptr; // device result when the caller's vector is host memory
  // CSR pattern of the assembled matrix (scalar rows/cols = bs * node + comp), built once on request
  int64_t nnz = 0;
  int32_t* row_ptr = nullptr;  // device [bs*n_dofs + 1]
  int32_t* col = nullptr;      // device [nnz], sorted within a row
  // position of every element-matrix entry in the CSR values, [nd*nd][n_cells] (coalesced over cells), filled by
  // bisection at the first eo_form_matrix after eo_form_set_pattern when memory allows; -1 = not in the pattern
  int32_t* pos = nullptr;
  unsigned* missing = nullptr;  // device counter of element entries the pattern does not hold
  struct form_pipe* pipe = nullptr;  // chunk pipeline of the host-vector path (built at first use)
};

struct form_weights {
  double w[EO_TAB_MAX_NQ];
};

// inverse Jacobian AND |det J| of one affine cell (vertex coordinates through the read-only path)
template <int GDIM>
__device__ __forceinline__ double form_geometry(const tab_tables& T, const int32_t* __restrict__ x_dofmap,
                                                const double* __restrict__ x, int64_t c, double K[GDIM][GDIM]) {
  double xv[GDIM + 1][GDIM];
#pragma unroll
  for (int v = 0; v < GDIM + 1; ++v) {
    const int32_t node = __ldg(x_dofmap + c * (GDIM + 1) + v);
#pragma unroll
    for (int i = 0; i < GDIM; ++i) xv[v][i] = __ldg(x + 3 * int64_t(node) + i);
  }
  return form_geometry_xv<GDIM>(T, xv, K);
}

// the same from the per-cell cache of the handle when it exists (eo_tab_geometry; warp-uniform branch): one 256-bit and one
// 64-bit load instead of 3 + 6 dependent ones and the division
__device__ __forceinline__ double form_cell_geometry(const tab_tables& T, const int32_t* __restrict__ x_dofmap,
                                                     const double* __restrict__ x, const double* __restrict__ geoK,
                                                     const double* __restrict__ geoD, int64_t c, double K[2][2]) {
  if (geoK) {
    const eo_d4 k = eo_ld256(geoK + 4 * c);
    K[0][0] = k.x, K[0][1] = k.y, K[1][0] = k.z, K[1][1] = k.w;
    return eo_ld64(geoD + c);
  }
  return form_geometry<2>(T, x_dofmap, x, c, K);
}

// the cell's coefficients through the dofmap (read-only path; neighbouring cells / points share nodes: L1/L2 hits)
template <int BS, int NB>
__device__ __forceinline__ void form_gather(const int32_t* __restrict__ dofmap, const double* __restrict__ u, int64_t c,
                                            double w[NB][BS]) {
  tab_gather<BS, NB>(dofmap, u, c, w);
}

// Scatter of one cell's element vector per thread (block size 2) by lane PAIRS: the even lane adds component 0 and the odd
// lane component 1 of the SAME node in the same instruction - first for the even lane's cell, then for the odd lane's - so
// that a RED instruction touches 16 aligned 16-byte chunks instead of 32 scattered 8-byte words.  With one RED per lane
// and entry, a Z-order numbered mesh spent 2.5x the L1 wavefronts on the scatter that a row-major numbered one does
// (ncu: 14.7 M against 5.8 M per 2e7 points in the per-cell residual step, which lost 15 % there; with pairs it is level).
// All 32 lanes must call; `active` = this lane has a cell.
template <int NB>
__device__ __forceinline__ void form_scatter_pairs(double* __restrict__ y, const int32_t idx[NB], const double fe[NB][2],
                                                   bool active) {
  const bool odd = threadIdx.x & 1u;
  const bool pact = __shfl_xor_sync(0xffffffffu, int(active), 1) != 0;
#pragma unroll
  for (int a = 0; a < NB; ++a) {
    const double recv = __shfl_xor_sync(0xffffffffu, odd ? fe[a][0] : fe[a][1], 1);  // the partner's share for MY component
    const int32_t pidx = __shfl_xor_sync(0xffffffffu, idx[a], 1);
    if (odd ? pact : active) atomicAdd(y + 2 * int64_t(odd ? pidx : idx[a]) + (odd ? 1 : 0), odd ? recv : fe[a][0]);
    if (odd ? active : pact) atomicAdd(y + 2 * int64_t(odd ? idx[a] : pidx) + (odd ? 1 : 0), odd ? fe[a][1] : recv);
  }
}

// ------------------------------------------------------------------------------------------------------------
// Mapping of the fused residual step: one thread per QUADRATURE POINT, a CTA of FORM_THREADS threads covers
// FORM_THREADS / nq consecutive cells (a cell never straddles CTAs).  The per-point streams (stress 32 B, tangent
// 128 B, history, outputs) are then read and written exactly like in the streaming kernels - consecutive threads,
// consecutive records, two dependent memory round trips per thread - and the nq threads of a cell gather the same
// coefficients / geometry (same sectors: one L1 request).  Each thread turns its point's cotangent into its
// contribution to the cell's nb*bs element-vector entries, parks it in shared memory (row stride nb*bs + 1: no bank
// conflicts), and after a barrier the CTA's threads sum the nq contributions of each entry and issue ONE fp64
// RED.ADD per element-vector entry (each DOF of a P2 triangle mesh is touched by 2-6 cells: low contention,
// resolved in L2).  The CTA grid-strides over tiles of cells.
// ------------------------------------------------------------------------------------------------------------
#define FORM_THREADS 256

// Element tables staged in shared memory: the point index differs between the threads of a warp, which the
// constant bank would serialise (one pass per distinct address).
template <int GDIM, int NB>
struct form_tabs {
  double phi[EO_TAB_MAX_NQ][NB];
  double dphi[EO_TAB_MAX_NQ][GDIM][NB];
};

template <int GDIM, int NB>
__device__ __forceinline__ void form_stage_tables(const tab_tables& T, form_tabs<GDIM, NB>& S) {
  for (int t = threadIdx.x; t < T.nq * NB; t += blockDim.x) {
    const int q = t / NB, a = t - q * NB;
    S.phi[q][a] = T.phi[q][a];
#pragma unroll
    for (int k = 0; k < GDIM; ++k) S.dphi[q][k][a] = T.dphi[k][q][a];
  }
  __syncthreads();
}

template <int GDIM, int NB>
EO_TAB_HD double form_phi(const form_tabs<GDIM, NB>& S, int q, int a) { return S.phi[q][a]; }
template <int GDIM, int NB>
EO_TAB_HD double form_dphi(const form_tabs<GDIM, NB>& S, int k, int q, int a) { return S.dphi[q][k][a]; }

// tab_point with the tables in shared memory (same statement order: identical results)
template <int GDIM, int BS, int NB>
__device__ __forceinline__ void form_point(const form_tabs<GDIM, NB>& S, const double w[NB][BS], const double K[GDIM][GDIM],
                                           int q, bool want_value, bool want_grad, double val[BS], double grad[BS][GDIM]) {
  if (want_value) {
#pragma unroll
    for (int c = 0; c < BS; ++c) {
      double acc = 0.0;
#pragma unroll
      for (int a = 0; a < NB; ++a) acc += w[a][c] * S.phi[q][a];
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
        for (int a = 0; a < NB; ++a) acc += w[a][c] * S.dphi[q][k][a];
        G[c][k] = acc;
      }
#pragma unroll
    for (int c = 0; c < BS; ++c)
#pragma unroll
      for (int j = 0; j < GDIM; ++j) {
        double acc = 0.0;
#pragma unroll
        for (int k = 0; k < GDIM; ++k) acc += G[c][k] * K[k][j];
        grad[c][j] = acc;
      }
  }
}

struct form_tile {
  int64_t c;    // this thread's cell
  int q;        // its point within the cell
  bool active;  // false for the <= nq - 1 spare threads of the CTA and past the last cell
};

__device__ __forceinline__ form_tile form_locate(int nq, int cpb, int64_t tile, int64_t n_cells) {
  const unsigned l = unsigned(threadIdx.x) / unsigned(nq);
  form_tile t;
  t.q = int(unsigned(threadIdx.x) - l * unsigned(nq));
  t.c = tile * cpb + l;
  t.active = int(l) < cpb && t.c < n_cells;
  return t;
}

// this thread's contribution (cotangent tau of `kind` at point q, scaled) -> shared -> one RED per element entry
template <int GDIM, int BS, int NB>
__device__ __forceinline__ void form_reduce_scatter(const form_tabs<GDIM, NB>& S, int nq, int kind, const form_tile& t, double scale,
                                                    const double* tau, const double K[GDIM][GDIM],
                                                    const int32_t* __restrict__ dofmap, int64_t tile, int cpb,
                                                    int64_t n_cells, double* __restrict__ b, double* s_fe) {
  constexpr int ND = NB * BS;
  if (t.active) {
    double Vs[BS], Gs[BS][GDIM], fe[NB][BS];
#pragma unroll
    for (int a = 0; a < NB; ++a)
#pragma unroll
      for (int k = 0; k < BS; ++k) fe[a][k] = 0.0;
    form_cotangent<GDIM, BS>(kind, tau, Vs, Gs);
    form_accumulate<GDIM, BS, NB>(S, kind, t.q, scale, Vs, Gs, K, fe);
    double* row = s_fe + threadIdx.x * (ND + 1);
#pragma unroll
    for (int a = 0; a < NB; ++a)
#pragma unroll
      for (int k = 0; k < BS; ++k) row[a * BS + k] = fe[a][k];
  }
  __syncthreads();
  const int64_t c0 = tile * cpb;
  const int cells_here = int((n_cells - c0) < cpb ? (n_cells - c0) : cpb);
  for (int j = threadIdx.x; j < cells_here * ND; j += FORM_THREADS) {
    const int l = j / ND, e = j - l * ND;
    const double* row = s_fe + (l * nq) * (ND + 1) + e;
    double acc = row[0];
    for (int q = 1; q < nq; ++q) acc += row[q * (ND + 1)];
    const int32_t node = __ldg(dofmap + (c0 + l) * NB + e / BS);
    atomicAdd(b + int64_t(BS) * node + (e % BS), acc);
  }
  __syncthreads();  // the rows are rewritten by the next tile
}

// b += sum_q w_q |det J| B_q^T coef[c][q].  One thread per CELL: geometry once, nq point records (contiguous per cell),
// nb*bs REDs.  For this light kernel (32 B of stream per point) the per-cell mapping beats the per-point mapping of the
// step kernel, whose nq-fold geometry and shared-memory reduction would dominate: 1.57 vs 2.94 ms per 1e8 points.
template <int GDIM, int BS, int NB>
__global__ void __launch_bounds__(128) form_vector_cell_kernel(const __grid_constant__ tab_tables T,
                                                               const __grid_constant__ form_weights W, int kind,
                                                               const int32_t* __restrict__ dofmap,
                                                               const int32_t* __restrict__ x_dofmap,
                                                               const double* __restrict__ x,
                                                               const double* __restrict__ geoK,
                                                               const double* __restrict__ geoD,
                                                               const double* __restrict__ coef, int64_t n_cells,
                                                               double* __restrict__ b) {
  const int64_t c = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
  const bool active = c < n_cells;  // block size 2: every lane stays for the pairwise scatter
  if (BS != 2 && !active) return;
  const int ncomp = tab_ncomp(kind, BS, GDIM);
  const bool vec4 = ncomp == 4 && (reinterpret_cast<uintptr_t>(coef) % 32) == 0;
  const double* s_ptr = coef + c * int64_t(T.nq) * ncomp;
  double K[GDIM][GDIM], adet = 0.0;
  int32_t idx[NB];
  double fe[NB][BS];
#pragma unroll
  for (int a = 0; a < NB; ++a) {
    idx[a] = 0;
#pragma unroll
    for (int k = 0; k < BS; ++k) fe[a][k] = 0.0;
  }
  if (active) {
    if constexpr (GDIM == 2)
      adet = form_cell_geometry(T, x_dofmap, x, geoK, geoD, c, K);
    else
      adet = form_geometry<GDIM>(T, x_dofmap, x, c, K);
#pragma unroll
    for (int a = 0; a < NB; ++a) idx[a] = __ldg(dofmap + c * NB + a);
  }
  for (int q = 0; active && q < T.nq; ++q) {
    double s[BS * GDIM > 4 ? BS * GDIM : 4];
    if (vec4) {
      const eo_d4 v = eo_ld256(s_ptr + 4 * q);
      s[0] = v.x, s[1] = v.y, s[2] = v.z, s[3] = v.w;
    } else {
      for (int k = 0; k < ncomp; ++k) s[k] = eo_ld64(s_ptr + q * ncomp + k);
    }
    double Vs[BS], Gs[BS][GDIM];
    form_cotangent<GDIM, BS>(kind, s, Vs, Gs);
    form_accumulate<GDIM, BS, NB, true>(T, kind, q, W.w[q] * adet, Vs, Gs, K, fe);
  }
  if constexpr (BS == 2) {
    form_scatter_pairs<NB>(b, idx, fe, active);
  } else {
#pragma unroll
    for (int a = 0; a < NB; ++a)
#pragma unroll
      for (int k = 0; k < BS; ++k) atomicAdd(b + int64_t(BS) * idx[a] + k, fe[a][k]);
  }
}

// y += sum_q w_q |det J| B_test,q^T ( D[c][q] (B_trial,q x) ),   D row-major (ncomp_test, ncomp_trial) per point.
// One thread per CELL: it gathers once and walks its nq points (NQ > 0: unrolled, all tangent loads of the cell are
// issued up front).  Measured against the per-point mapping of the other two integrals (which gathers nq times and
// pays the shared-memory reduction): 3.8 ms vs 4.8 ms per 1e8 points - both are bound by the LSU wavefront rate
// (~0.7 / clk / SM), not by HBM.  General fallback of the TMA-staged kernel below.
template <int GDIM, int BS, int NB, int NQ>
__global__ void __launch_bounds__(128) form_action_cell_kernel(const __grid_constant__ tab_tables T,
                                                               const __grid_constant__ form_weights W, int kind_test,
                                                               int kind_trial, const int32_t* __restrict__ dofmap,
                                                               const int32_t* __restrict__ x_dofmap,
                                                               const double* __restrict__ x, const double* __restrict__ D,
                                                               const double* __restrict__ xin, int64_t n_cells,
                                                               double* __restrict__ y) {
  const int64_t c = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
  if (c >= n_cells) return;
  constexpr int MAXC = BS * GDIM > 4 ? BS * GDIM : 4;
  const int nq = NQ > 0 ? NQ : T.nq;
  const int nt = tab_ncomp(kind_test, BS, GDIM), ni = tab_ncomp(kind_trial, BS, GDIM);
  const bool vec44 = nt == 4 && ni == 4 && (reinterpret_cast<uintptr_t>(D) % 32) == 0;
  const double* D_ptr = D + c * int64_t(nq) * nt * ni;
  eo_d4 d[NQ > 0 ? NQ : 1][4];
  if (NQ > 0 && vec44) {
#pragma unroll
    for (int q = 0; q < NQ; ++q)
#pragma unroll
      for (int r = 0; r < 4; ++r) d[q][r] = eo_ld256(D_ptr + 16 * q + 4 * r);
  }
  double K[GDIM][GDIM];
  const double adet = form_geometry<GDIM>(T, x_dofmap, x, c, K);
  int32_t idx[NB];
#pragma unroll
  for (int a = 0; a < NB; ++a) idx[a] = __ldg(dofmap + c * NB + a);
  double w[NB][BS];
#pragma unroll
  for (int a = 0; a < NB; ++a) {
    if constexpr (BS == 2) {
      const double2 v = __ldg(reinterpret_cast<const double2*>(xin) + idx[a]);
      w[a][0] = v.x, w[a][1] = v.y;
    } else {
#pragma unroll
      for (int k = 0; k < BS; ++k) w[a][k] = __ldg(xin + int64_t(BS) * idx[a] + k);
    }
  }
  double fe[NB][BS];
#pragma unroll
  for (int a = 0; a < NB; ++a)
#pragma unroll
    for (int k = 0; k < BS; ++k) fe[a][k] = 0.0;
#pragma unroll
  for (int q = 0; q < nq; ++q) {
    double val[BS], grad[BS][GDIM], e[MAXC], tau[MAXC];
    tab_point<GDIM, BS, NB, true>(T, w, K, q, kind_trial == 0, kind_trial != 0, val, grad);
    tab_operand<GDIM, BS>(kind_trial, val, grad, e);
    if (vec44) {
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        const eo_d4 dd = NQ > 0 ? d[NQ > 0 ? q : 0][r] : eo_ld256(D_ptr + 16 * q + 4 * r);
        tau[r] = fma(dd.w, e[3], fma(dd.z, e[2], fma(dd.y, e[1], dd.x * e[0])));
      }
    } else {
      const double* Dq = D_ptr + int64_t(q) * nt * ni;
      for (int r = 0; r < nt; ++r) {
        double acc = 0.0;
        for (int l = 0; l < ni; ++l) acc = fma(eo_ld64(Dq + r * ni + l), e[l], acc);
        tau[r] = acc;
      }
    }
    double Vs[BS], Gs[BS][GDIM];
    form_cotangent<GDIM, BS>(kind_test, tau, Vs, Gs);
    form_accumulate<GDIM, BS, NB, true>(T, kind_test, q, W.w[q] * adet, Vs, Gs, K, fe);
  }
#pragma unroll
  for (int a = 0; a < NB; ++a)
#pragma unroll
    for (int k = 0; k < BS; ++k) atomicAdd(y + int64_t(BS) * idx[a] + k, fe[a][k]);
}

// Per-cell action with the tangent staged by the TMA unit: every thread asks for ITS cell's NQ contiguous 4x4 tangents
// (NQ * 128 B) with one 1-D bulk copy into a padded shared-memory row (row stride NQ * 128 + 16 B: the 16-byte reads
// of 8 consecutive lanes then cover all 32 banks), completion on one mbarrier per CTA.  The copies bypass the LSU - whose
// wavefront rate, not HBM, bounds the register-path kernels: a warp-wide 256-bit load of 128-byte records touches 32
// lines per instruction - and are in flight while the thread gathers its coefficients and geometry.
// (A persistent variant with two row buffers per CTA and the next tile's gathers software-pipelined was measured
// slower - 4.4 vs 3.3 ms per 1e8 points at 8 warps per SM - and is not kept.)
template <int NB, int NQ>
__global__ void __launch_bounds__(128) form_action_tma_kernel(const __grid_constant__ tab_tables T,
                                                              const __grid_constant__ form_weights W, int kind_test,
                                                              int kind_trial, const int32_t* __restrict__ dofmap,
                                                              const int32_t* __restrict__ x_dofmap,
                                                              const double* __restrict__ x, const double* __restrict__ geoK,
                                                              const double* __restrict__ geoD, const double* __restrict__ D,
                                                              const double* __restrict__ xin, int64_t n_cells,
                                                              double* __restrict__ y) {
  extern __shared__ __align__(128) unsigned char form_smem_raw[];
  constexpr unsigned CELL_B = NQ * 128, ROW_B = CELL_B + 16;
  const unsigned bar = (unsigned)__cvta_generic_to_shared(form_smem_raw);
  const unsigned rows = bar + 128;
  const int64_t c0 = blockIdx.x * int64_t(128);
  const int64_t c = c0 + threadIdx.x;
  const unsigned cells_here = (unsigned)((n_cells - c0) < 128 ? (n_cells - c0) : 128);
  if (threadIdx.x == 0) {  // the transaction count is armed before any copy can complete
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(cells_here * CELL_B) : "memory");
  }
  __syncthreads();
  if (c < n_cells)
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     rows + threadIdx.x * ROW_B),
                 "l"(D + c * int64_t(NQ) * 16), "r"(CELL_B), "r"(bar)
                 : "memory");
  double K[2][2], w[NB][2], adet = 0.0;
  int32_t idx[NB];
#pragma unroll
  for (int a = 0; a < NB; ++a) idx[a] = 0;
  if (c < n_cells) {
    adet = form_cell_geometry(T, x_dofmap, x, geoK, geoD, c, K);
#pragma unroll
    for (int a = 0; a < NB; ++a) idx[a] = __ldg(dofmap + c * NB + a);
#pragma unroll
    for (int a = 0; a < NB; ++a) {
      const double2 v = __ldg(reinterpret_cast<const double2*>(xin) + idx[a]);
      w[a][0] = v.x, w[a][1] = v.y;
    }
  }
  unsigned done = 0;
  while (!done)
    asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0; selp.u32 %0, 1, 0, p; }"
                 : "=r"(done)
                 : "r"(bar)
                 : "memory");
  const bool active = c < n_cells;  // every lane stays for the pairwise scatter
  double fe[NB][2];
#pragma unroll
  for (int a = 0; a < NB; ++a) fe[a][0] = 0.0, fe[a][1] = 0.0;
  const unsigned my = rows + threadIdx.x * ROW_B;
  if (active)
#pragma unroll
  for (int q = 0; q < NQ; ++q) {
    double val[2] = {0.0, 0.0}, grad[2][2], e[4], tau[4];
    tab_point<2, 2, NB, true>(T, w, K, q, false, true, val, grad);
    tab_operand<2, 2>(kind_trial, val, grad, e);
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      double d0, d1, d2, d3;
      asm volatile("ld.shared.v2.f64 {%0,%1}, [%2];" : "=d"(d0), "=d"(d1) : "r"(my + q * 128 + r * 32));
      asm volatile("ld.shared.v2.f64 {%0,%1}, [%2];" : "=d"(d2), "=d"(d3) : "r"(my + q * 128 + r * 32 + 16));
      tau[r] = fma(d3, e[3], fma(d2, e[2], fma(d1, e[1], d0 * e[0])));
    }
    double Vs[2], Gs[2][2];
    form_cotangent<2, 2>(kind_test, tau, Vs, Gs);
    form_accumulate<2, 2, NB, true>(T, kind_test, q, W.w[q] * adet, Vs, Gs, K, fe);
  }
  form_scatter_pairs<NB>(y, idx, fe, active);
}

// Tangent action from the FACTORED von Mises tangent (6 doubles per point: v[4], cn, cd - vm_core.cuh): the shape of
// form_action_tma_kernel with 144 instead of 384 bytes per cell through the TMA unit (row stride 144 B = 9 x 16: the 16-byte
// reads of 8 consecutive lanes cover all 32 banks without padding) and tau = C_t e rebuilt in registers
// (vm_factored_apply).  Mandel strain on both sides, three points per cell.  96 instead of 176 algorithmic bytes per point.
// 78 registers, 6 CTAs per SM; forcing 8 / 7 CTAs (64 / 72 registers, spills) or letting ptxas use 106 registers: 2.31 / 2.10 /
// 2.11 ms against 1.90 ms per 1e8 points.
template <int NB>
__global__ void __launch_bounds__(128) form_action_vm6_kernel(const __grid_constant__ tab_tables T,
                                                              const __grid_constant__ form_weights W, const vm_consts vq,
                                                              const int32_t* __restrict__ dofmap,
                                                              const int32_t* __restrict__ x_dofmap,
                                                              const double* __restrict__ x, const double* __restrict__ geoK,
                                                              const double* __restrict__ geoD, const double* __restrict__ T6,
                                                              const double* __restrict__ xin, int64_t n_cells,
                                                              double* __restrict__ y) {
  extern __shared__ __align__(128) unsigned char form_smem_raw[];
  constexpr unsigned CELL_B = 3 * 48, ROW_B = CELL_B;
  const unsigned bar = (unsigned)__cvta_generic_to_shared(form_smem_raw);
  const unsigned rows = bar + 128;
  const int64_t c0 = blockIdx.x * int64_t(128);
  const int64_t c = c0 + threadIdx.x;
  const unsigned cells_here = (unsigned)((n_cells - c0) < 128 ? (n_cells - c0) : 128);
  if (threadIdx.x == 0) {  // the transaction count is armed before any copy can complete
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(cells_here * CELL_B) : "memory");
  }
  __syncthreads();
  if (c < n_cells)
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     rows + threadIdx.x * ROW_B),
                 "l"(T6 + c * int64_t(18)), "r"(CELL_B), "r"(bar)
                 : "memory");
  double K[2][2], w[NB][2], adet = 0.0;
  int32_t idx[NB];
#pragma unroll
  for (int a = 0; a < NB; ++a) idx[a] = 0;
  if (c < n_cells) {
    adet = form_cell_geometry(T, x_dofmap, x, geoK, geoD, c, K);
    tab_load_idx<NB>(dofmap, c, idx);
    tab_gather_idx<2, NB>(xin, idx, w);
  }
  unsigned done = 0;
  while (!done)
    asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0; selp.u32 %0, 1, 0, p; }"
                 : "=r"(done)
                 : "r"(bar)
                 : "memory");
  const bool active = c < n_cells;  // every lane stays for the pairwise scatter
  double fe[NB][2];
#pragma unroll
  for (int a = 0; a < NB; ++a) fe[a][0] = 0.0, fe[a][1] = 0.0;
  const unsigned my = rows + threadIdx.x * ROW_B;
  if (active)
#pragma unroll
  for (int q = 0; q < 3; ++q) {
    double val[2] = {0.0, 0.0}, grad[2][2], e[4], tau[4], f[6];
    tab_point<2, 2, NB, true>(T, w, K, q, false, true, val, grad);
    tab_operand<2, 2>(2, val, grad, e);
#pragma unroll
    for (int r = 0; r < 3; ++r)
      asm volatile("ld.shared.v2.f64 {%0,%1}, [%2];" : "=d"(f[2 * r]), "=d"(f[2 * r + 1]) : "r"(my + q * 48 + r * 16));
    vm_factored_apply(vq, f, f[4], f[5], e, tau);
    double Vs[2], Gs[2][2];
    form_cotangent<2, 2>(2, tau, Vs, Gs);
    form_accumulate<2, 2, NB, true>(T, 2, q, W.w[q] * adet, Vs, Gs, K, fe);
  }
  form_scatter_pairs<NB>(y, idx, fe, active);
}

// factored tangent -> the reference's (n, 4, 4) layout, by the statements of vm_point / vm_point_fast (bit-identical to what
// the un-factored kernels store)
template <bool EXACT>
__global__ void __launch_bounds__(256) vm_expand_kernel(const vm_consts vq, const double* __restrict__ T6, int64_t n,
                                                        double* __restrict__ C_tang) {
  const int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
  if (i >= n) return;
  const double2 a = eo_ld128(T6 + 6 * i), b2 = eo_ld128(T6 + 6 * i + 2), cc = eo_ld128(T6 + 6 * i + 4);
  const double v[4] = {a.x, a.y, b2.x, b2.y};
  double C[16];
  if (EXACT)
    vm_tangent_exact(vq, v, cc.x, cc.y, C);
  else
    vm_tangent_fast(vq, v, cc.x, cc.y, C);
  double* Ct = C_tang + 16 * i;
  eo_st256(Ct + 0, C[0], C[1], C[2], C[3]);
  eo_st256(Ct + 4, C[4], C[5], C[6], C[7]);
  eo_st256(Ct + 8, C[8], C[9], C[10], C[11]);
  eo_st256(Ct + 12, C[12], C[13], C[14], C[15]);
}

// One Newton residual evaluation of the von Mises problem without leaving the device:
// Mandel strain of u -> radial return (C_tang, sigma, dp stored for the tangent action / the history commit)
// -> b += int sigma . epsilon(v) dx.  The per-point arithmetic and its results are those of eo_tab_vm_fused.
// (A per-cell variant with TMA bulk stores of tangent and stress - the analogue of form_action_tma_kernel - was measured
// slower, 6.6 vs 5.8 ms per 1e8 points: 160 registers leave 12 warps per SM for the division-heavy radial return.)
// FACT: the tangent is stored in its factored form (v[4], cn, cd: 48 instead of 128 bytes per point, vm_core.cuh) for
// eo_form_action_vm_factored; `C_tang` then points at that [n_points][6] array.
template <int NB, bool EXACT, bool FACT>
__global__ void __launch_bounds__(FORM_THREADS, 3) form_vm_step_kernel(
    const __grid_constant__ tab_tables T, const __grid_constant__ form_weights W, const vm_consts vq,
    const int32_t* __restrict__ dofmap, const int32_t* __restrict__ x_dofmap, const double* __restrict__ x,
    const double* __restrict__ u, const double* __restrict__ geoK, const double* __restrict__ geoD, int64_t n_cells,
    const double* __restrict__ sigma_n, const double* __restrict__ p, double* __restrict__ C_tang,
    double* __restrict__ sigma, double* __restrict__ dp_out, double* __restrict__ b, eo_stats* stats) {
  extern __shared__ double s_fe[];
  __shared__ form_tabs<2, NB> S;
  form_stage_tables<2, NB>(T, S);
  const int cpb = FORM_THREADS / T.nq;
  const int64_t tiles = (n_cells + cpb - 1) / cpb;
  int plastic = 0;
  for (int64_t tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
    const form_tile t = form_locate(T.nq, cpb, tile, n_cells);
    double K[2][2], g[4], scale = 0.0;
    if (t.active) {
      const int64_t i = t.c * T.nq + t.q;
      // history first: these loads are in flight while the gather and the contraction run
      const eo_d4 s = eo_ld256(sigma_n + 4 * i);
      const double pi = eo_ld64(p + i);
      double w[NB][2];
      form_gather<2, NB>(dofmap, u, t.c, w);
      scale = W.w[t.q] * form_cell_geometry(T, x_dofmap, x, geoK, geoD, t.c, K);
      double val[2] = {0.0, 0.0}, grad[2][2], e[4];
      form_point<2, 2, NB>(S, w, K, t.q, false, true, val, grad);
      tab_operand<2, 2>(2, val, grad, e);
      vm_point_out o;
      if (EXACT)
        vm_point(vq, e[0], e[1], e[2], e[3], s.x, s.y, s.z, s.w, pi, o);
      else
        vm_point_fast(vq, e[0], e[1], e[2], e[3], s.x, s.y, s.z, s.w, pi, o);
      plastic += o.dp > 0.0;
      if (FACT) {
        double* Tf = C_tang + 6 * i;
        eo_st128(Tf + 0, o.v[0], o.v[1]);
        eo_st128(Tf + 2, o.v[2], o.v[3]);
        eo_st128(Tf + 4, o.cn, o.cd);
      } else {
        double* Ct = C_tang + 16 * i;
        eo_st256(Ct + 0, o.C[0], o.C[1], o.C[2], o.C[3]);
        eo_st256(Ct + 4, o.C[4], o.C[5], o.C[6], o.C[7]);
        eo_st256(Ct + 8, o.C[8], o.C[9], o.C[10], o.C[11]);
        eo_st256(Ct + 12, o.C[12], o.C[13], o.C[14], o.C[15]);
      }
      eo_st256(sigma + 4 * i, o.g[0], o.g[1], o.g[2], o.g[3]);
      eo_st64(dp_out + i, o.dp);
      g[0] = o.g[0], g[1] = o.g[1], g[2] = o.g[2], g[3] = o.g[3];
    }
    form_reduce_scatter<2, 2, NB>(S, T.nq, 2, t, scale, g, K, dofmap, tile, cpb, n_cells, b, s_fe);
  }
  eo_block_sum_add(reinterpret_cast<unsigned long long*>(&stats->n_plastic), plastic);
  if (blockIdx.x == 0 && threadIdx.x == 0)
    atomicAdd(reinterpret_cast<unsigned long long*>(&stats->n_points), (unsigned long long)(n_cells * T.nq));
}

// The same step with ONE THREAD PER CELL for three points per cell (the demos' degree-2 rule, P1 / P2 triangles): dofmap
// row, cached geometry, coefficient gather and the history of the cell's three points are requested up front; the points
// are then processed one after the other with the element tables as constant-bank operands (the point index is a
// compile-time constant), each point's tangent / stress / dp leaving at once (three consecutive records per thread: whole
// sectors), and the element vector accumulates in registers - no shared memory, no barriers, nb*bs REDs per cell, 35 %
// fewer instructions than the per-point mapping.  168 registers, 12 warps per SM, persistent grid-stride loop (the
// block-wide sum of the plastic counts, one barrier, is paid once per CTA).  Per-point arithmetic is that of the kernel
// above, statement for statement (bit-identical per-point results).  Measured per 1e8 points: 5.03 ms against 5.60 for the
// per-point kernel on the row-major numbered mesh, 5.50 against 5.52 on the Z-order numbered one (2 % ahead on RCM, 1 %
// behind on a randomly shuffled numbering) - with the element vector scattered by lane pairs (form_scatter_pairs); with
// one RED per lane and entry this mapping lost 15 % on the Z-order numbering.  Variants measured slower: 128 registers /
// 16 warps (5.71, spills), L2 prefetch of the next cell's streams (5.29), a register-free cp.async pipeline of the next
// cell's inputs through shared memory (5.40) - profiles/r2_rejected_variants.md.
#define FORM_CELL_THREADS 128
template <int NB, bool EXACT, bool FACT>  // FACT: the tangent stored as its 6 factors (see form_vm_step_kernel)
__global__ void __launch_bounds__(FORM_CELL_THREADS, 3) form_vm_step_cell_kernel(
    const __grid_constant__ tab_tables T, const __grid_constant__ form_weights W, const vm_consts vq,
    const int32_t* __restrict__ dofmap, const int32_t* __restrict__ x_dofmap, const double* __restrict__ x,
    const double* __restrict__ u, const double* __restrict__ geoK, const double* __restrict__ geoD, int64_t n_cells,
    const double* __restrict__ sigma_n, const double* __restrict__ p, double* __restrict__ C_tang,
    double* __restrict__ sigma, double* __restrict__ dp_out, double* __restrict__ b, eo_stats* stats) {
  int plastic = 0;
  // the loop condition is warp uniform (the lanes of a pair exchange contributions in the scatter); lanes past the end idle
  for (int64_t cw = blockIdx.x * int64_t(blockDim.x) + (threadIdx.x & ~31u); cw < n_cells; cw += int64_t(gridDim.x) * blockDim.x) {
    const int64_t c = cw + (threadIdx.x & 31u);
    const bool active = c < n_cells;
    int32_t idx[NB];
    double fe[NB][2];
#pragma unroll
    for (int a = 0; a < NB; ++a) idx[a] = 0, fe[a][0] = 0.0, fe[a][1] = 0.0;
    if (active) {
    tab_load_idx<NB>(dofmap, c, idx);
    const int64_t i0 = 3 * c;
    eo_d4 sn[3];
    double pn[3];
#pragma unroll
    for (int q = 0; q < 3; ++q) sn[q] = eo_ld256(sigma_n + 4 * (i0 + q)), pn[q] = eo_ld64(p + i0 + q);
    double w[NB][2], K[2][2];
    const double adet = form_cell_geometry(T, x_dofmap, x, geoK, geoD, c, K);
    tab_gather_idx<2, NB>(u, idx, w);
#pragma unroll
    for (int q = 0; q < 3; ++q) {
      double val[2] = {0.0, 0.0}, grad[2][2], e[4];
      tab_point<2, 2, NB>(T, w, K, q, false, true, val, grad);
      tab_operand<2, 2>(2, val, grad, e);
      vm_point_out o;
      if (EXACT)
        vm_point(vq, e[0], e[1], e[2], e[3], sn[q].x, sn[q].y, sn[q].z, sn[q].w, pn[q], o);
      else
        vm_point_fast(vq, e[0], e[1], e[2], e[3], sn[q].x, sn[q].y, sn[q].z, sn[q].w, pn[q], o);
      plastic += o.dp > 0.0;
      if (FACT) {
        double* Tf = C_tang + 6 * (i0 + q);
        eo_st128(Tf + 0, o.v[0], o.v[1]);
        eo_st128(Tf + 2, o.v[2], o.v[3]);
        eo_st128(Tf + 4, o.cn, o.cd);
      } else {
        double* Ct = C_tang + 16 * (i0 + q);
        eo_st256(Ct + 0, o.C[0], o.C[1], o.C[2], o.C[3]);
        eo_st256(Ct + 4, o.C[4], o.C[5], o.C[6], o.C[7]);
        eo_st256(Ct + 8, o.C[8], o.C[9], o.C[10], o.C[11]);
        eo_st256(Ct + 12, o.C[12], o.C[13], o.C[14], o.C[15]);
      }
      eo_st256(sigma + 4 * (i0 + q), o.g[0], o.g[1], o.g[2], o.g[3]);
      eo_st64(dp_out + i0 + q, o.dp);
      double Vs[2], Gs[2][2];
      form_cotangent<2, 2>(2, o.g, Vs, Gs);
      form_accumulate<2, 2, NB>(T, 2, q, W.w[q] * adet, Vs, Gs, K, fe);
    }
    }
    form_scatter_pairs<NB>(b, idx, fe, active);
  }
  eo_block_sum_add(reinterpret_cast<unsigned long long*>(&stats->n_plastic), plastic);
  if (blockIdx.x == 0 && threadIdx.x == 0)
    atomicAdd(reinterpret_cast<unsigned long long*>(&stats->n_points), (unsigned long long)(n_cells * 3));
}

// position of (row, col) in the CSR values by bisection in the sorted row, or -1
__device__ __forceinline__ int32_t form_csr_find(const int32_t* __restrict__ row_ptr, const int32_t* __restrict__ col,
                                                 int32_t grow, int32_t gcol) {
  int32_t lo = __ldg(row_ptr + grow);
  const int32_t end = __ldg(row_ptr + grow + 1);
  int32_t hi = end - 1;
  if (hi < lo) return -1;
  while (lo < hi) {
    const int32_t mid = (lo + hi) >> 1;
    if (__ldg(col + mid) < gcol) lo = mid + 1; else hi = mid;
  }
  return __ldg(col + lo) == gcol ? lo : -1;
}

// the positions of all (nb*bs)^2 element-matrix entries of every cell, once per pattern
template <int BS, int NB>
__global__ void __launch_bounds__(128) form_positions_kernel(const int32_t* __restrict__ dofmap, int64_t n_cells,
                                                             const int32_t* __restrict__ row_ptr,
                                                             const int32_t* __restrict__ col, int32_t* __restrict__ pos) {
  constexpr int ND = NB * BS;
  const int64_t c = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
  if (c >= n_cells) return;
  for (int r = 0; r < ND; ++r) {
    const int32_t grow = BS * __ldg(dofmap + c * NB + r / BS) + r % BS;
    for (int cc = 0; cc < ND; ++cc) {
      const int32_t gcol = BS * __ldg(dofmap + c * NB + cc / BS) + cc % BS;
      pos[(int64_t(r) * ND + cc) * n_cells + c] = form_csr_find(row_ptr, col, grow, gcol);
    }
  }
}

// A[pos] += element matrix entries; positions from the cache above, or found by bisection in the sorted CSR row
// (rows of a P2 triangle mesh hold <= ~40 entries: <= 6 probes, all L1/L2 hits on the pattern); entries the pattern does
// not hold are dropped and counted.
template <int GDIM, int BS, int NB>
__global__ void __launch_bounds__(128) form_matrix_kernel(const __grid_constant__ tab_tables T,
                                                          const __grid_constant__ form_weights W, int kind_test,
                                                          int kind_trial, const int32_t* __restrict__ dofmap,
                                                          const int32_t* __restrict__ x_dofmap,
                                                          const double* __restrict__ x, const double* __restrict__ D,
                                                          int64_t n_cells, const int32_t* __restrict__ row_ptr,
                                                          const int32_t* __restrict__ col, double* __restrict__ vals,
                                                          const int32_t* __restrict__ pos, int64_t pos_stride,
                                                          unsigned* __restrict__ missing) {
  const int64_t c = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
  if (c >= n_cells) return;
  constexpr int ND = NB * BS;
  double K[GDIM][GDIM];
  const double adet = form_geometry<GDIM>(T, x_dofmap, x, c, K);
  int32_t idx[NB];
#pragma unroll
  for (int a = 0; a < NB; ++a) idx[a] = __ldg(dofmap + c * NB + a);
  constexpr int MAXC = BS * GDIM > 4 ? BS * GDIM : 4;
  const int nt = tab_ncomp(kind_test, BS, GDIM), ni = tab_ncomp(kind_trial, BS, GDIM);
  const double* D_ptr = D + c * int64_t(T.nq) * nt * ni;
  unsigned miss = 0;
  // one trial basis function (bj, cj) at a time: its element-matrix COLUMN is the action on a unit vector
  for (int bj = 0; bj < NB; ++bj)
    for (int cj = 0; cj < BS; ++cj) {
      double fe[NB][BS];
#pragma unroll
      for (int a = 0; a < NB; ++a)
#pragma unroll
        for (int k = 0; k < BS; ++k) fe[a][k] = 0.0;
      for (int q = 0; q < T.nq; ++q) {
        double val[BS], grad[BS][GDIM], e[MAXC], tau[MAXC];
#pragma unroll
        for (int k = 0; k < BS; ++k) {
          val[k] = (k == cj) ? T.phi[q][bj] : 0.0;
#pragma unroll
          for (int j = 0; j < GDIM; ++j) {
            double acc = 0.0;
#pragma unroll
            for (int kk = 0; kk < GDIM; ++kk) acc += T.dphi[kk][q][bj] * K[kk][j];
            grad[k][j] = (k == cj) ? acc : 0.0;
          }
        }
        tab_operand<GDIM, BS>(kind_trial, val, grad, e);
        const double* Dq = D_ptr + int64_t(q) * nt * ni;
        for (int r = 0; r < nt; ++r) {
          double acc = 0.0;
          for (int l = 0; l < ni; ++l) acc = fma(__ldg(Dq + r * ni + l), e[l], acc);
          tau[r] = acc;
        }
        double Vs[BS], Gs[BS][GDIM];
        form_cotangent<GDIM, BS>(kind_test, tau, Vs, Gs);
        form_accumulate<GDIM, BS, NB, true>(T, kind_test, q, W.w[q] * adet, Vs, Gs, K, fe);
      }
      const int32_t gcol = BS * __ldg(dofmap + c * NB + bj) + cj;
#pragma unroll
      for (int a = 0; a < NB; ++a)
#pragma unroll
        for (int k = 0; k < BS; ++k) {
          const int32_t at = pos ? __ldg(pos + (int64_t(a * BS + k) * ND + (bj * BS + cj)) * pos_stride + c)
                                 : form_csr_find(row_ptr, col, BS * idx[a] + k, gcol);
          if (at >= 0)
            atomicAdd(vals + at, fe[a][k]);
          else
            ++miss;
        }
    }
  if (miss) atomicAdd(missing, miss);
}

static int form_kind(int kind) { return kind == EO_OPERAND_DEF_GRAD ? EO_OPERAND_GRAD : kind; }  // d(I + grad u) = grad du

#define EO_FORM_CASES(X) \
  X(2, 1, 3) X(2, 1, 6) X(2, 2, 3) X(2, 2, 6) X(2, 1, 10) X(2, 2, 10) X(3, 1, 4) X(3, 3, 4) X(3, 1, 10) X(3, 3, 10)

// tuning switches (A/B measurements; defaults are the measured best)
static int form_env(const char* name, int dflt) {
  const char* v = getenv(name);
  return v ? atoi(v) : dflt;
}

#ifndef FORM_WAVES
#define FORM_WAVES 8
#endif
// persistent-ish grid: a few waves of the resident CTAs, each grid-striding over tiles of FORM_THREADS / nq cells
static unsigned form_grid(eo_ctx* ctx, const eo_tab* t, int64_t n_cells) {
  const int cpb = FORM_THREADS / t->T.nq;
  const int64_t tiles = (n_cells + cpb - 1) / cpb;
  const int64_t cap = int64_t(ctx->sm_count) * 4 * FORM_WAVES;
  return (unsigned)(tiles < cap ? tiles : cap);
}

// dynamic shared memory of the element-vector rows (FORM_THREADS rows of nd + 1 doubles); opt in above 48 KB
template <class Kernel>
static size_t form_smem(eo_ctx* ctx, Kernel k, int nd) {
  const size_t bytes = size_t(FORM_THREADS) * (nd + 1) * sizeof(double);
  if (bytes > 48 * 1024) cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, int(bytes));
  return bytes;
}

// result vector on the device side: the caller's (device) or the staging copy (host); zeroed unless accumulating
static int form_result(eo_form* f, double* y, int accumulate, double** d_y) {
  eo_tab* t = f->tab;
  eo_ctx* ctx = t->ctx;
  const size_t bytes = size_t(t->n_dofs) * t->T.bs * sizeof(double);
  if (eo_is_device_ptr(y)) {
    *d_y = y;
  } else {
    if (!f->y_stage) EO_CUDA(ctx, cudaMalloc(&f->y_stage, bytes ? bytes : 8));
    *d_y = f->y_stage;
    if (accumulate) EO_CUDA(ctx, cudaMemcpyAsync(f->y_stage, y, bytes, cudaMemcpyHostToDevice, ctx->s_cmp));
  }
  if (!accumulate) EO_CUDA(ctx, cudaMemsetAsync(*d_y, 0, bytes, ctx->s_cmp));
  return EO_OK;
}

static int form_finish(eo_form* f, double* y, double* d_y) {
  eo_ctx* ctx = f->tab->ctx;
  EO_CUDA(ctx, cudaGetLastError());
  if (d_y != y) {
    const size_t bytes = size_t(f->tab->n_dofs) * f->tab->T.bs * sizeof(double);
    EO_CUDA(ctx, cudaMemcpyAsync(y, d_y, bytes, cudaMemcpyDeviceToHost, ctx->s_cmp));
    EO_CUDA(ctx, cudaStreamSynchronize(ctx->s_cmp));
  }
  return EO_OK;
}

static int form_stage_x(eo_form* f, const double* x, const double** d_x) {
  eo_ctx* ctx = f->tab->ctx;
  if (eo_is_device_ptr(x)) {
    *d_x = x;
    return EO_OK;
  }
  const size_t bytes = size_t(f->tab->n_dofs) * f->tab->T.bs * sizeof(double);
  if (!f->x_stage) EO_CUDA(ctx, cudaMalloc(&f->x_stage, bytes ? bytes : 8));
  EO_CUDA(ctx, cudaMemcpyAsync(f->x_stage, x, bytes, cudaMemcpyHostToDevice, ctx->s_cmp));
  *d_x = f->x_stage;
  return EO_OK;
}

// ------------------------------------------------------------------------------------------------------------
// Host-vector path (u / x in host memory, b / y back to host memory): dependency-aware chunk pipeline.
// The DOF vector is the only thing that crosses PCIe (about 11 B per point each way for P2 triangles), but a serial
// H2D -> kernel -> D2H spends 88 % of its time in the two copies.  The cells are cut into K consecutive chunks and the
// vector into P pieces; a small kernel records once per mesh which pieces every chunk reads / accumulates into.  Then
//   * the pieces of u go up in the order of their FIRST use, chunk k starts as soon as the pieces it needs have landed,
//   * a piece of b comes down as soon as the LAST chunk that adds to it has finished,
// on three streams, so that the upload of later pieces, the kernels and the download of finished pieces overlap (the
// link is full duplex).  With a numbering that has locality (structured, RCM, anything DOLFINx produces) this hides
// the kernel and one of the two copies; with a random numbering every chunk touches every piece and the schedule
// degenerates into the serial one - same results either way (to the rounding of the atomic scatter).
// ------------------------------------------------------------------------------------------------------------
#define FORM_PIPE_CHUNKS 16
#define FORM_PIPE_PIECES 64

struct form_pipe {
  int K = 0, P = 0;
  int64_t piece = 0;                    // doubles per piece
  int64_t n_cells = 0;
  std::vector<int64_t> cell_lo;         // K + 1 chunk boundaries
  std::vector<int> h2d_order, d2h_order;
  std::vector<int> h2d_upto, d2h_upto;  // per chunk: length of the prefix of *_order that belongs to chunks <= k
  std::vector<cudaEvent_t> ev_in, ev_k;
};

__global__ void __launch_bounds__(256) form_touch_kernel(const int32_t* __restrict__ dofmap, int64_t n_cells, int nb, int bs,
                                                         int64_t chunk_cells, int64_t piece, int words,
                                                         unsigned int* __restrict__ bits) {
  const int64_t j = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
  if (j >= n_cells * nb) return;
  const int64_t c = j / nb;
  const int k = int(c / chunk_cells);
  const int64_t d0 = int64_t(bs) * dofmap[j];
  const int p0 = int(d0 / piece), p1 = int((d0 + bs - 1) / piece);
  atomicOr(bits + size_t(k) * words + (p0 >> 5), 1u << (p0 & 31));
  if (p1 != p0) atomicOr(bits + size_t(k) * words + (p1 >> 5), 1u << (p1 & 31));
}

static void form_pipe_free(form_pipe* pp) {
  if (!pp) return;
  for (cudaEvent_t e : pp->ev_in) cudaEventDestroy(e);
  for (cudaEvent_t e : pp->ev_k) cudaEventDestroy(e);
  delete pp;
}

static int form_pipe_build(eo_form* f, int64_t n_cells) {
  eo_tab* t = f->tab;
  eo_ctx* ctx = t->ctx;
  form_pipe_free(f->pipe);
  f->pipe = nullptr;
  form_pipe* pp = new form_pipe();
  const int64_t nd = int64_t(t->n_dofs) * t->T.bs;
  pp->K = form_env("EO_FORM_PIPE_CHUNKS", FORM_PIPE_CHUNKS), pp->P = form_env("EO_FORM_PIPE_PIECES", FORM_PIPE_PIECES), pp->n_cells = n_cells;
  if (pp->K < 1) pp->K = 1;
  if (pp->P < 1) pp->P = 1;
  pp->piece = ((nd + pp->P - 1) / pp->P + 3) / 4 * 4;  // 32-byte multiples
  pp->P = int((nd + pp->piece - 1) / pp->piece);
  const int64_t chunk_cells = (n_cells + pp->K - 1) / pp->K;
  pp->K = int((n_cells + chunk_cells - 1) / chunk_cells);
  for (int k = 0; k <= pp->K; ++k) pp->cell_lo.push_back(k * chunk_cells < n_cells ? k * chunk_cells : n_cells);
  const int words = (pp->P + 31) / 32;
  unsigned int* d_bits = nullptr;
  std::vector<unsigned int> bits(size_t(pp->K) * words, 0u);
  cudaError_t e = cudaMalloc(&d_bits, bits.size() * 4);
  if (e == cudaSuccess) e = cudaMemsetAsync(d_bits, 0, bits.size() * 4, ctx->s_cmp);
  if (e == cudaSuccess) {
    const int64_t n = n_cells * t->T.nb;
    form_touch_kernel<<<unsigned((n + 255) / 256), 256, 0, ctx->s_cmp>>>(t->dofmap, n_cells, t->T.nb, t->T.bs, chunk_cells,
                                                                         pp->piece, words, d_bits);
    e = cudaMemcpyAsync(bits.data(), d_bits, bits.size() * 4, cudaMemcpyDeviceToHost, ctx->s_cmp);
  }
  if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->s_cmp);
  if (d_bits) cudaFree(d_bits);
  if (e != cudaSuccess) {
    delete pp;
    return eo_fail(ctx, EO_ERR_CUDA, "eo_form: building the chunk pipeline: %s", cudaGetErrorString(e));
  }
  std::vector<int> first(pp->P, pp->K), last(pp->P, -1);
  for (int k = 0; k < pp->K; ++k)
    for (int q = 0; q < pp->P; ++q)
      if (bits[size_t(k) * words + (q >> 5)] >> (q & 31) & 1u) {
        if (first[q] > k) first[q] = k;
        last[q] = k;
      }
  pp->h2d_upto.assign(pp->K, 0), pp->d2h_upto.assign(pp->K, 0);
  for (int k = 0; k <= pp->K; ++k) {  // k == K: pieces no cell reads (uploaded last, nothing waits for them)
    for (int q = 0; q < pp->P; ++q)
      if (first[q] == k) pp->h2d_order.push_back(q);
    if (k < pp->K) pp->h2d_upto[k] = int(pp->h2d_order.size());
  }
  for (int k = 0; k < pp->K; ++k) {
    for (int q = 0; q < pp->P; ++q)
      if (last[q] == k || (k == 0 && last[q] < 0)) pp->d2h_order.push_back(q);  // untouched pieces: zeros, after chunk 0
    pp->d2h_upto[k] = int(pp->d2h_order.size());
  }
  pp->ev_in.resize(pp->K), pp->ev_k.resize(pp->K);
  for (int k = 0; k < pp->K; ++k) {
    cudaEventCreateWithFlags(&pp->ev_in[k], cudaEventDisableTiming);
    cudaEventCreateWithFlags(&pp->ev_k[k], cudaEventDisableTiming);
  }
  f->pipe = pp;
  return EO_OK;
}

static bool form_pipe_applies(eo_form* f, const void* in, const void* out, int accumulate, int64_t n_cells) {
  static const bool enabled = form_env("EO_FORM_PIPE", 1) != 0;
  return enabled && !accumulate && n_cells == f->tab->n_cells && n_cells >= (int64_t(1) << 17) && !eo_is_device_ptr(in) &&
         !eo_is_device_ptr(out);
}

// launch(c0, c1, d_in, d_out) enqueues the kernel(s) for the cells [c0, c1) on ctx->s_cmp
template <class Launch>
static int form_pipelined(eo_form* f, const double* in_host, double* out_host, int64_t n_cells, Launch launch) {
  eo_tab* t = f->tab;
  eo_ctx* ctx = t->ctx;
  if (!f->pipe || f->pipe->n_cells != n_cells) {
    const int rc = form_pipe_build(f, n_cells);
    if (rc != EO_OK) return rc;
  }
  form_pipe& pp = *f->pipe;
  const int64_t nd = int64_t(t->n_dofs) * t->T.bs;
  const size_t bytes = size_t(nd) * sizeof(double);
  if (!t->u_stage[0]) EO_CUDA(ctx, cudaMalloc(&t->u_stage[0], bytes ? bytes : 8));
  if (!f->y_stage) EO_CUDA(ctx, cudaMalloc(&f->y_stage, bytes ? bytes : 8));
  double* d_in = t->u_stage[0];
  double* d_out = f->y_stage;
  auto piece_bytes = [&](int q) { return size_t((q + 1) * pp.piece <= nd ? pp.piece : nd - q * pp.piece) * sizeof(double); };
  // the copy streams see everything queued on the compute stream so far
  EO_CUDA(ctx, cudaEventRecord(ctx->ev_cmp[0], ctx->s_cmp));
  EO_CUDA(ctx, cudaStreamWaitEvent(ctx->s_in, ctx->ev_cmp[0], 0));
  EO_CUDA(ctx, cudaStreamWaitEvent(ctx->s_out, ctx->ev_cmp[0], 0));
  EO_CUDA(ctx, cudaMemsetAsync(d_out, 0, bytes, ctx->s_cmp));
  int up = 0, down = 0;
  for (int k = 0; k < pp.K; ++k) {
    for (; up < pp.h2d_upto[k]; ++up) {
      const int q = pp.h2d_order[up];
      EO_CUDA(ctx, cudaMemcpyAsync(d_in + q * pp.piece, in_host + q * pp.piece, piece_bytes(q), cudaMemcpyHostToDevice, ctx->s_in));
    }
    EO_CUDA(ctx, cudaEventRecord(pp.ev_in[k], ctx->s_in));
    EO_CUDA(ctx, cudaStreamWaitEvent(ctx->s_cmp, pp.ev_in[k], 0));
    if (pp.cell_lo[k + 1] > pp.cell_lo[k]) {
      const int rc = launch(pp.cell_lo[k], pp.cell_lo[k + 1], (const double*)d_in, d_out);
      if (rc != EO_OK) return rc;
      EO_CUDA(ctx, cudaGetLastError());
    }
    EO_CUDA(ctx, cudaEventRecord(pp.ev_k[k], ctx->s_cmp));
    EO_CUDA(ctx, cudaStreamWaitEvent(ctx->s_out, pp.ev_k[k], 0));
    for (; down < pp.d2h_upto[k]; ++down) {
      const int q = pp.d2h_order[down];
      EO_CUDA(ctx, cudaMemcpyAsync(out_host + q * pp.piece, d_out + q * pp.piece, piece_bytes(q), cudaMemcpyDeviceToHost, ctx->s_out));
    }
  }
  for (; up < int(pp.h2d_order.size()); ++up) {  // pieces no cell reads: keep the staged copy complete
    const int q = pp.h2d_order[up];
    EO_CUDA(ctx, cudaMemcpyAsync(d_in + q * pp.piece, in_host + q * pp.piece, piece_bytes(q), cudaMemcpyHostToDevice, ctx->s_in));
  }
  EO_CUDA(ctx, cudaStreamSynchronize(ctx->s_in));
  EO_CUDA(ctx, cudaStreamSynchronize(ctx->s_cmp));
  EO_CUDA(ctx, cudaStreamSynchronize(ctx->s_out));
  return EO_OK;
}

static int form_check_kind(eo_form* f, int kind, const char* what) {
  if (kind < 0 || kind > 3 || eo_tab_ncomp(f->tab, kind) <= 0)
    return eo_fail(f->tab->ctx, EO_ERR_INVALID, "%s: operand kind %d does not fit this element", what, kind);
  return EO_OK;
}

extern "C" {

int eo_form_create(eo_tab* tab, const double* weights, eo_form** out) {
  if (!tab) return EO_ERR_INVALID;
  eo_ctx* ctx = tab->ctx;
  EO_REQUIRE(ctx, weights && out, "eo_form_create: NULL argument");
  eo_form* f = new eo_form();
  f->tab = tab;
  f->ctx = ctx;
  memset(f->w, 0, sizeof(f->w));
  for (int q = 0; q < tab->T.nq; ++q) f->w[q] = weights[q];
  *out = f;
  return EO_OK;
}

int eo_form_destroy(eo_form* f) {
  if (f) form_pipe_free(f->pipe), f->pipe = nullptr;
  if (!f) return EO_OK;
  cudaSetDevice(f->ctx->device);
  cudaStreamSynchronize(f->ctx->s_cmp);
  if (f->x_stage) cudaFree(f->x_stage);
  if (f->y_stage) cudaFree(f->y_stage);
  if (f->row_ptr) cudaFree(f->row_ptr);
  if (f->col) cudaFree(f->col);
  if (f->pos) cudaFree(f->pos);
  if (f->missing) cudaFree(f->missing);
  delete f;
  return EO_OK;
}

int eo_form_vector(eo_form* f, int kind_test, const double* coef, int64_t n_cells, double* b, int accumulate) {
  if (!f) return EO_ERR_INVALID;
  eo_tab* t = f->tab;
  eo_ctx* ctx = t->ctx;
  int rc = form_check_kind(f, kind_test, "eo_form_vector");
  if (rc != EO_OK) return rc;
  if (n_cells < 0) n_cells = t->n_cells;
  EO_REQUIRE(ctx, n_cells <= t->n_cells, "eo_form_vector: more cells requested than the mesh has");
  EO_REQUIRE(ctx, b != nullptr, "eo_form_vector: b is NULL");
  EO_REQUIRE(ctx, n_cells == 0 || (coef && eo_is_device_ptr(coef)), "eo_form_vector: the point values must be device memory");
  EO_CUDA(ctx, cudaSetDevice(ctx->device));
  double* d_b = nullptr;
  rc = form_result(f, b, accumulate, &d_b);
  if (rc != EO_OK) return rc;
  const int kind = form_kind(kind_test);
  if (n_cells > 0) {
    const int grc = eo_tab_geometry(t);
    if (grc != EO_OK) return grc;
    form_weights W;
    memcpy(W.w, f->w, sizeof(W.w));
    bool done = false;
#define X(G, B, N)                                                                                                     \
  if (!done && t->T.gdim == G && t->T.bs == B && t->T.nb == N) {                                                       \
    form_vector_cell_kernel<G, B, N><<<(unsigned)((n_cells + 127) / 128), 128, 0, ctx->s_cmp>>>(                       \
        t->T, W, kind, t->dofmap, t->x_dofmap, t->x, t->geoK, t->geoD, coef, n_cells, d_b);                            \
    done = true;                                                                                                       \
  }
    EO_FORM_CASES(X)
#undef X
    if (!done) return eo_fail(ctx, EO_ERR_UNSUPPORTED, "eo_form_vector: no kernel for this element");
    ctx->launches += 1;
  }
  return form_finish(f, b, d_b);
}

int eo_form_action(eo_form* f, int kind_test, int kind_trial, const double* D, const double* x, int64_t n_cells,
                   double* y, int accumulate) {
  if (!f) return EO_ERR_INVALID;
  eo_tab* t = f->tab;
  eo_ctx* ctx = t->ctx;
  int rc = form_check_kind(f, kind_test, "eo_form_action");
  if (rc != EO_OK) return rc;
  rc = form_check_kind(f, kind_trial, "eo_form_action");
  if (rc != EO_OK) return rc;
  if (n_cells < 0) n_cells = t->n_cells;
  EO_REQUIRE(ctx, n_cells <= t->n_cells, "eo_form_action: more cells requested than the mesh has");
  EO_REQUIRE(ctx, x && y, "eo_form_action: NULL vector");
  EO_REQUIRE(ctx, n_cells == 0 || (D && eo_is_device_ptr(D)), "eo_form_action: the point values must be device memory");
  EO_CUDA(ctx, cudaSetDevice(ctx->device));
  const int kt = form_kind(kind_test), ki = form_kind(kind_trial);
  form_weights W;
  memcpy(W.w, f->w, sizeof(W.w));
  const int64_t d_per_cell = int64_t(t->T.nq) * eo_tab_ncomp(t, kind_test) * eo_tab_ncomp(t, kind_trial);
  // TMA-staged kernel for 4x4 tangents on 2-d vector fields: C_tang of the plasticity demos (Mandel strain both sides),
  // dP/dF of the hyperelasticity demo (gradient both sides)   (EO_FORM_ACTION_TMA=0: register-path kernel, for A/B)
  const bool tma = form_env("EO_FORM_ACTION_TMA", 1) && t->T.gdim == 2 && t->T.bs == 2 && t->T.nq == 3 && kt != 0 && ki != 0 &&
                   eo_aligned(D, 32);
  if (tma && n_cells > 0) {
    const int grc = eo_tab_geometry(t);
    if (grc != EO_OK) return grc;
  }
  // the cells [c0, c1): every per-cell / per-point array is addressed relative to c0
  auto launch = [&](int64_t c0, int64_t c1, const double* dx, double* dy) -> int {
    const int64_t m = c1 - c0;
    const double* Dc = D + c0 * d_per_cell;
    bool done = false;
    const unsigned gc = (unsigned)((m + 127) / 128);
#define X(G, B, N)                                                                                                       \
  if (!done && t->T.gdim == G && t->T.bs == B && t->T.nb == N) {                                                         \
    const int32_t* dm = t->dofmap + c0 * N;                                                                              \
    const int32_t* xd = t->x_dofmap + c0 * (G + 1);                                                                      \
    if (tma) {                                                                                                           \
      const size_t sm = 128 + 128 * (3 * 128 + 16);                                                                      \
      auto kfn = form_action_tma_kernel<(G == 2 && B == 2 ? N : 3), 3>;                                                  \
      cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, int(sm));                                   \
      kfn<<<gc, 128, sm, ctx->s_cmp>>>(t->T, W, kt, ki, dm, xd, t->x, t->geoK ? t->geoK + 4 * c0 : nullptr,              \
                                       t->geoK ? t->geoD + c0 : nullptr, Dc, dx, m, dy);                                 \
    } else if (t->T.nq == 3) {                                                                                           \
      form_action_cell_kernel<G, B, N, 3><<<gc, 128, 0, ctx->s_cmp>>>(t->T, W, kt, ki, dm, xd, t->x, Dc, dx, m, dy);     \
    } else {                                                                                                             \
      form_action_cell_kernel<G, B, N, 0><<<gc, 128, 0, ctx->s_cmp>>>(t->T, W, kt, ki, dm, xd, t->x, Dc, dx, m, dy);     \
    }                                                                                                                    \
    done = true;                                                                                                         \
  }
    EO_FORM_CASES(X)
#undef X
    if (!done) return eo_fail(ctx, EO_ERR_UNSUPPORTED, "eo_form_action: no kernel for this element");
    ctx->launches += 1;
    return EO_OK;
  };
  if (n_cells > 0 && form_pipe_applies(f, x, y, accumulate, n_cells)) return form_pipelined(f, x, y, n_cells, launch);
  const double* d_x = nullptr;
  rc = form_stage_x(f, x, &d_x);
  if (rc != EO_OK) return rc;
  double* d_y = nullptr;
  rc = form_result(f, y, accumulate, &d_y);
  if (rc != EO_OK) return rc;
  EO_REQUIRE(ctx, d_x != d_y, "eo_form_action: x and y must not alias");
  if (n_cells > 0) {
    rc = launch(0, n_cells, d_x, d_y);
    if (rc != EO_OK) return rc;
  }
  return form_finish(f, y, d_y);
}

// fact == 0: C_tang is the (n, 4, 4) tangent; fact == 1: the factored tangent, 6 doubles per point
static int form_vm_step_impl(eo_form* f, const eo_vm_params* prm, const double* u, const double* sigma_n, const double* p,
                             double* C_tang, double* sigma, double* dp, int64_t n_cells, double* b, int accumulate,
                             int exact, int fact) {
  if (!f) return EO_ERR_INVALID;
  eo_tab* t = f->tab;
  eo_ctx* ctx = t->ctx;
  EO_REQUIRE(ctx, prm != nullptr, "eo_form_vm_step: prm is NULL");
  EO_REQUIRE(ctx, t->T.gdim == 2 && t->T.bs == 2, "eo_form_vm_step: needs a 2-d vector field (plane-strain Mandel strain)");
  EO_REQUIRE(ctx, t->T.nb == 3 || t->T.nb == 6 || t->T.nb == 10, "eo_form_vm_step: P1/P2/P3 triangles only");
  if (n_cells < 0) n_cells = t->n_cells;
  EO_REQUIRE(ctx, n_cells <= t->n_cells, "eo_form_vm_step: more cells requested than the mesh has");
  EO_REQUIRE(ctx, u && b, "eo_form_vm_step: NULL vector");
  EO_REQUIRE(ctx, n_cells == 0 || (sigma_n && p && C_tang && sigma && dp), "eo_form_vm_step: NULL array");
  EO_REQUIRE(ctx, n_cells == 0 || (eo_is_device_ptr(sigma_n) && eo_is_device_ptr(p) && eo_is_device_ptr(C_tang) &&
                                   eo_is_device_ptr(sigma) && eo_is_device_ptr(dp)),
             "eo_form_vm_step: history and point outputs must be device memory (u and b may be host memory)");
  EO_REQUIRE(ctx, eo_aligned(sigma_n, 32) && eo_aligned(C_tang, fact ? 16 : 32) && eo_aligned(sigma, 32),
             "eo_form_vm_step: arrays must be 32-byte aligned (the factored tangent: 16)");
  EO_CUDA(ctx, cudaSetDevice(ctx->device));
  const vm_consts q{prm->lmbda, prm->mu, prm->H, prm->sigma_0};
  form_weights W;
  memcpy(W.w, f->w, sizeof(W.w));
  const int tw = fact ? 6 : 16;  // doubles per point of the tangent array
  if (n_cells > 0) {
    const int grc = eo_tab_geometry(t);
    if (grc != EO_OK) return grc;
  }
  // three points per cell on P1 / P2 triangles: one thread per cell (EO_STEP_CELL=0: the per-point kernel, for A/B)
  static const bool cell_env = [] { const char* e = getenv("EO_STEP_CELL"); return !(e && *e == '0'); }();
  const bool cellwise = cell_env && t->T.nq == 3 && (t->T.nb == 3 || t->T.nb == 6);
  // the cells [c0, c1): every per-cell / per-point array is addressed relative to c0
  auto launch = [&](int64_t c0, int64_t c1, const double* du, double* db) -> int {
    const int64_t m = c1 - c0, o = c0 * t->T.nq;
    const unsigned grid = form_grid(ctx, t, m);
#define EO_STEP_ARGS(N)                                                                                              \
  t->T, W, q, t->dofmap + c0 * N, t->x_dofmap + c0 * 3, t->x, du, t->geoK ? t->geoK + 4 * c0 : nullptr,                \
      t->geoK ? t->geoD + c0 : nullptr, m, sigma_n + 4 * o, p + o, C_tang + tw * o, sigma + 4 * o, dp + o, db, ctx->stats
#define EO_STEP_X(N, X, F)                                                                                           \
  if (bool(exact) == X && bool(fact) == F) {                                                                         \
    const size_t sm = form_smem(ctx, form_vm_step_kernel<N, X, F>, N * 2);                                           \
    form_vm_step_kernel<N, X, F><<<grid, FORM_THREADS, sm, ctx->s_cmp>>>(EO_STEP_ARGS(N));                           \
  }
#define EO_STEP(N)                                                                                                   \
  if (t->T.nb == N) {                                                                                                \
    EO_STEP_X(N, true, false) else EO_STEP_X(N, false, false) else EO_STEP_X(N, true, true) else EO_STEP_X(N, false, true) \
  }
    if (cellwise) {
      const int64_t tiles_c = (m + FORM_CELL_THREADS - 1) / FORM_CELL_THREADS, cap_c = int64_t(ctx->sm_count) * 3 * FORM_WAVES;
      const unsigned gridc = (unsigned)(tiles_c < cap_c ? tiles_c : cap_c);
#define EO_STEP_CELL(N, X)                                                                                              \
  if (t->T.nb == N && bool(exact) == X) {                                                                               \
    if (fact) form_vm_step_cell_kernel<N, X, true><<<gridc, FORM_CELL_THREADS, 0, ctx->s_cmp>>>(EO_STEP_ARGS(N));       \
    else form_vm_step_cell_kernel<N, X, false><<<gridc, FORM_CELL_THREADS, 0, ctx->s_cmp>>>(EO_STEP_ARGS(N));           \
  }
      EO_STEP_CELL(3, true) EO_STEP_CELL(3, false) EO_STEP_CELL(6, true) EO_STEP_CELL(6, false)
#undef EO_STEP_CELL
      ctx->launches += 1;
      return EO_OK;
    }
    EO_STEP(3)
    EO_STEP(6)
    EO_STEP(10)
#undef EO_STEP
#undef EO_STEP_X
#undef EO_STEP_ARGS
    ctx->launches += 1;
    return EO_OK;
  };
  if (n_cells > 0 && form_pipe_applies(f, u, b, accumulate, n_cells)) return form_pipelined(f, u, b, n_cells, launch);
  const double* d_u = nullptr;
  int rc = eo_tab_stage_u(t, u, &d_u);
  if (rc != EO_OK) return rc;
  double* d_b = nullptr;
  rc = form_result(f, b, accumulate, &d_b);
  if (rc != EO_OK) return rc;
  if (n_cells > 0) launch(0, n_cells, d_u, d_b);
  return form_finish(f, b, d_b);
}

int eo_form_vm_step(eo_form* f, const eo_vm_params* prm, const double* u, const double* sigma_n, const double* p,
                    double* C_tang, double* sigma, double* dp, int64_t n_cells, double* b, int accumulate, int exact) {
  return form_vm_step_impl(f, prm, u, sigma_n, p, C_tang, sigma, dp, n_cells, b, accumulate, exact, 0);
}

int eo_form_vm_step_factored(eo_form* f, const eo_vm_params* prm, const double* u, const double* sigma_n, const double* p,
                             double* T6, double* sigma, double* dp, int64_t n_cells, double* b, int accumulate, int exact) {
  return form_vm_step_impl(f, prm, u, sigma_n, p, T6, sigma, dp, n_cells, b, accumulate, exact, 1);
}

int eo_form_action_vm_factored(eo_form* f, const eo_vm_params* prm, const double* T6, const double* x, int64_t n_cells,
                               double* y, int accumulate) {
  if (!f) return EO_ERR_INVALID;
  eo_tab* t = f->tab;
  eo_ctx* ctx = t->ctx;
  EO_REQUIRE(ctx, prm != nullptr, "eo_form_action_vm_factored: prm is NULL");
  EO_REQUIRE(ctx, t->T.gdim == 2 && t->T.bs == 2 && t->T.nq == 3 && (t->T.nb == 3 || t->T.nb == 6 || t->T.nb == 10),
             "eo_form_action_vm_factored: P1/P2/P3 vector triangles with three points per cell only");
  if (n_cells < 0) n_cells = t->n_cells;
  EO_REQUIRE(ctx, n_cells <= t->n_cells, "eo_form_action_vm_factored: more cells requested than the mesh has");
  EO_REQUIRE(ctx, x && y, "eo_form_action_vm_factored: NULL vector");
  EO_REQUIRE(ctx, n_cells == 0 || (T6 && eo_is_device_ptr(T6) && eo_aligned(T6, 16)),
             "eo_form_action_vm_factored: the factored tangent must be 16-byte aligned device memory");
  EO_CUDA(ctx, cudaSetDevice(ctx->device));
  const vm_consts q{prm->lmbda, prm->mu, prm->H, prm->sigma_0};
  form_weights W;
  memcpy(W.w, f->w, sizeof(W.w));
  if (n_cells > 0) {
    const int grc = eo_tab_geometry(t);
    if (grc != EO_OK) return grc;
  }
  auto launch = [&](int64_t c0, int64_t c1, const double* dx, double* dy) -> int {
    const int64_t m = c1 - c0;
    const unsigned gc = (unsigned)((m + 127) / 128);
    const size_t sm = 128 + 128 * (3 * 48);
#define EO_ACT6(N)                                                                                                      \
  if (t->T.nb == N)                                                                                                     \
    form_action_vm6_kernel<N><<<gc, 128, sm, ctx->s_cmp>>>(t->T, W, q, t->dofmap + c0 * N, t->x_dofmap + c0 * 3, t->x,   \
                                                           t->geoK ? t->geoK + 4 * c0 : nullptr,                        \
                                                           t->geoK ? t->geoD + c0 : nullptr, T6 + 18 * c0, dx, m, dy);
    EO_ACT6(3) EO_ACT6(6) EO_ACT6(10)
#undef EO_ACT6
    ctx->launches += 1;
    return EO_OK;
  };
  if (n_cells > 0 && form_pipe_applies(f, x, y, accumulate, n_cells)) return form_pipelined(f, x, y, n_cells, launch);
  const double* d_x = nullptr;
  int rc = form_stage_x(f, x, &d_x);
  if (rc != EO_OK) return rc;
  double* d_y = nullptr;
  rc = form_result(f, y, accumulate, &d_y);
  if (rc != EO_OK) return rc;
  EO_REQUIRE(ctx, d_x != d_y, "eo_form_action_vm_factored: x and y must not alias");
  if (n_cells > 0) {
    rc = launch(0, n_cells, d_x, d_y);
    if (rc != EO_OK) return rc;
  }
  return form_finish(f, y, d_y);
}

int eo_vm_expand_tangent(eo_ctx* ctx, const eo_vm_params* prm, const double* T6, double* C_tang, int64_t n, int exact) {
  if (!ctx) return EO_ERR_INVALID;
  EO_REQUIRE(ctx, prm != nullptr, "eo_vm_expand_tangent: prm is NULL");
  EO_REQUIRE(ctx, n >= 0, "eo_vm_expand_tangent: negative n");
  if (n == 0) return EO_OK;
  EO_REQUIRE(ctx, T6 && C_tang && eo_is_device_ptr(T6) && eo_is_device_ptr(C_tang) && eo_aligned(T6, 16) && eo_aligned(C_tang, 32),
             "eo_vm_expand_tangent: device arrays, 16- / 32-byte aligned");
  EO_CUDA(ctx, cudaSetDevice(ctx->device));
  const vm_consts q{prm->lmbda, prm->mu, prm->H, prm->sigma_0};
  const unsigned grid = (unsigned)((n + 255) / 256);
  if (exact) vm_expand_kernel<true><<<grid, 256, 0, ctx->s_cmp>>>(q, T6, n, C_tang);
  else vm_expand_kernel<false><<<grid, 256, 0, ctx->s_cmp>>>(q, T6, n, C_tang);
  ctx->launches += 1;
  EO_CUDA(ctx, cudaGetLastError());
  return EO_OK;
}

int eo_form_set_pattern(eo_form* f, const int32_t* row_ptr, const int32_t* col, int64_t nnz) {
  if (!f) return EO_ERR_INVALID;
  eo_tab* t = f->tab;
  eo_ctx* ctx = t->ctx;
  const int64_t n_rows = t->n_dofs * t->T.bs;
  EO_REQUIRE(ctx, row_ptr && (col || nnz == 0) && nnz >= 0, "eo_form_set_pattern: bad argument");
  EO_REQUIRE(ctx, nnz < 2147483647LL, "eo_form_set_pattern: more than 2^31 - 1 entries (int32 CSR)");
  EO_REQUIRE(ctx, row_ptr[0] == 0 && row_ptr[n_rows] == nnz, "eo_form_set_pattern: row_ptr does not span [0, nnz]");
  for (int64_t r = 0; r < n_rows; ++r) {
    if (row_ptr[r + 1] < row_ptr[r]) return eo_fail(ctx, EO_ERR_INVALID, "eo_form_set_pattern: row_ptr not monotone");
    for (int32_t k = row_ptr[r]; k < row_ptr[r + 1]; ++k) {
      if (col[k] < 0 || col[k] >= n_rows) return eo_fail(ctx, EO_ERR_INVALID, "eo_form_set_pattern: column out of range");
      if (k > row_ptr[r] && col[k] <= col[k - 1])
        return eo_fail(ctx, EO_ERR_INVALID, "eo_form_set_pattern: columns must be strictly increasing within a row");
    }
  }
  EO_CUDA(ctx, cudaSetDevice(ctx->device));
  EO_CUDA(ctx, cudaStreamSynchronize(ctx->s_cmp));
  if (f->row_ptr) cudaFree(f->row_ptr);
  if (f->col) cudaFree(f->col);
  if (f->pos) cudaFree(f->pos);
  f->row_ptr = nullptr, f->col = nullptr, f->pos = nullptr, f->nnz = 0;
  EO_CUDA(ctx, cudaMalloc(&f->row_ptr, size_t(n_rows + 1) * 4));
  EO_CUDA(ctx, cudaMalloc(&f->col, size_t(nnz ? nnz : 1) * 4));
  EO_CUDA(ctx, cudaMemcpyAsync(f->row_ptr, row_ptr, size_t(n_rows + 1) * 4, cudaMemcpyHostToDevice, ctx->s_cmp));
  EO_CUDA(ctx, cudaMemcpyAsync(f->col, col, size_t(nnz) * 4, cudaMemcpyHostToDevice, ctx->s_cmp));
  EO_CUDA(ctx, cudaStreamSynchronize(ctx->s_cmp));
  f->nnz = nnz;
  return EO_OK;
}

int eo_form_matrix(eo_form* f, int kind_test, int kind_trial, const double* D, int64_t n_cells, double* vals,
                   int accumulate) {
  if (!f) return EO_ERR_INVALID;
  eo_tab* t = f->tab;
  eo_ctx* ctx = t->ctx;
  int rc = form_check_kind(f, kind_test, "eo_form_matrix");
  if (rc != EO_OK) return rc;
  rc = form_check_kind(f, kind_trial, "eo_form_matrix");
  if (rc != EO_OK) return rc;
  EO_REQUIRE(ctx, f->row_ptr != nullptr, "eo_form_matrix: no sparsity pattern (call eo_form_set_pattern first)");
  if (n_cells < 0) n_cells = t->n_cells;
  EO_REQUIRE(ctx, n_cells <= t->n_cells, "eo_form_matrix: more cells requested than the mesh has");
  EO_REQUIRE(ctx, vals && eo_is_device_ptr(vals), "eo_form_matrix: the CSR values must be device memory");
  EO_REQUIRE(ctx, n_cells == 0 || (D && eo_is_device_ptr(D)), "eo_form_matrix: the point values must be device memory");
  EO_CUDA(ctx, cudaSetDevice(ctx->device));
  if (!accumulate) EO_CUDA(ctx, cudaMemsetAsync(vals, 0, size_t(f->nnz) * sizeof(double), ctx->s_cmp));
  const int kt = form_kind(kind_test), ki = form_kind(kind_trial);
  if (n_cells > 0) {
    const unsigned grid = (unsigned)((n_cells + 127) / 128);
    form_weights W;
    memcpy(W.w, f->w, sizeof(W.w));
    if (!f->missing) {
      EO_CUDA(ctx, cudaMalloc(&f->missing, sizeof(unsigned)));
      EO_CUDA(ctx, cudaMemsetAsync(f->missing, 0, sizeof(unsigned), ctx->s_cmp));
    }
    // position cache, built for ALL cells at the first assembly after eo_form_set_pattern when it fits in a quarter of the
    // free memory (P2 vector triangle: 576 B per cell); EO_FORM_MATRIX_POS=0: bisection in every assembly
    const int nd = t->T.nb * t->T.bs;
    bool done = false;
    if (!f->pos && form_env("EO_FORM_MATRIX_POS", 1)) {
      size_t free_b = 0, total_b = 0;
      const size_t need = size_t(nd) * nd * size_t(t->n_cells) * sizeof(int32_t);
      EO_CUDA(ctx, cudaMemGetInfo(&free_b, &total_b));
      if (need <= free_b / 4 && cudaMalloc(&f->pos, need) == cudaSuccess) {
        const unsigned gp = (unsigned)((t->n_cells + 127) / 128);
#define X(G, B, N)                                                                                                   \
  if (!done && t->T.gdim == G && t->T.bs == B && t->T.nb == N) {                                                     \
    form_positions_kernel<B, N><<<gp, 128, 0, ctx->s_cmp>>>(t->dofmap, t->n_cells, f->row_ptr, f->col, f->pos);             \
    done = true;                                                                                                     \
  }
        EO_FORM_CASES(X)
#undef X
        ctx->launches += 1;
      } else {
        cudaGetLastError();  // a failed cudaMalloc is not an error of this call: fall back to bisection
        f->pos = nullptr;
      }
    }
    done = false;
#define X(G, B, N)                                                                                                  \
  if (!done && t->T.gdim == G && t->T.bs == B && t->T.nb == N) {                                                    \
    form_matrix_kernel<G, B, N><<<grid, 128, 0, ctx->s_cmp>>>(t->T, W, kt, ki, t->dofmap, t->x_dofmap, t->x, D, n_cells, \
                                                              f->row_ptr, f->col, vals, f->pos, t->n_cells, f->missing); \
    done = true;                                                                                                    \
  }
    EO_FORM_CASES(X)
#undef X
    if (!done) return eo_fail(ctx, EO_ERR_UNSUPPORTED, "eo_form_matrix: no kernel for this element");
    ctx->launches += 1;
    // element entries outside the pattern are a caller error (a pattern that is not the cell-coupling pattern of this mesh)
    unsigned miss = 0;
    EO_CUDA(ctx, cudaMemcpyAsync(&miss, f->missing, sizeof(unsigned), cudaMemcpyDeviceToHost, ctx->s_cmp));
    EO_CUDA(ctx, cudaStreamSynchronize(ctx->s_cmp));
    if (miss) {
      EO_CUDA(ctx, cudaMemsetAsync(f->missing, 0, sizeof(unsigned), ctx->s_cmp));
      if (f->pos) cudaFree(f->pos);
      f->pos = nullptr;
      return eo_fail(ctx, EO_ERR_INVALID, "eo_form_matrix: %u element-matrix entries are not in the sparsity pattern", miss);
    }
  }
  EO_CUDA(ctx, cudaGetLastError());
  return EO_OK;
}

int64_t eo_form_nnz(const eo_form* f) { return f ? f->nnz : -1; }

}  // extern "C"
