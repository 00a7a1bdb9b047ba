// tab.cu - operand tabulation on sm_100a (hot path (a)) and its fusion with the von Mises update.
// replaces: `expr.eval(operand_mesh, entities)` of `evaluate_operands`,
//           src/dolfinx_external_operator/external_operator.py:393-402 (DOLFINx tabulate_expression + the
//           FFCx kernel: per-cell gather through the dofmap, affine Jacobian, contraction with the basix
//           tables), for the operand expressions of the reference demos (tab_core.cuh: eo_operand_kind).
//
// Mapping: one thread per cell.  The cell's dofmap row (nb int32) and geometry row are read with
// coalesced loads, the nb*bs DOF coefficients are gathered through the read-only path (neighbouring cells
// share nodes: L1/L2 hits), the tables sit in the constant bank (the evaluation-point loop index is warp
// uniform: broadcast reads), the contraction runs in registers and each thread writes its nq*ncomp
// contiguous doubles (96 B for the 3-point Mandel strain) with 256-bit stores where the row is 32 B wide.
// HBM-bound / gather-limited: ~100 FMA versus >= 176 B of compulsory traffic per cell.
//
// Fused kernel: the strain of each point goes straight from registers into the von Mises radial return
// (vm_core.cuh), so it never touches HBM: 235 B per point instead of 240 + 2 x 32.
// Compiled with -fmad=false like vm_heat.cu (the fused and the two-step paths then agree bit for bit).
#include "eo_common.cuh"
#include "tab_core.cuh"
#include "tab_handle.cuh"
#include "vm_core.cuh"
#include "form_core.cuh"

#include <cstdlib>

#ifndef TAB_FUSED_WAVES
#define TAB_FUSED_WAVES 8
#endif


template <int GDIM, int BS, int NB>
__global__ void __launch_bounds__(128) tab_kernel(const __grid_constant__ tab_tables T, int kind,
                                                  const int32_t* __restrict__ dofmap,
                                                  const int32_t* __restrict__ x_dofmap, const double* __restrict__ x,
                                                  const double* __restrict__ u, const double* __restrict__ geoK,
                                                  const int32_t* __restrict__ cells, int64_t n_cells,
                                                  double* __restrict__ out) {
  const int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
  if (i >= n_cells) return;
  const int64_t c = cells ? int64_t(cells[i]) : i;
  double w[NB][BS], K[GDIM][GDIM];
  bool cached = false;
  if constexpr (GDIM == 2) {
    if (geoK) {  // the handle's per-cell inverse Jacobians (eo_tab_geometry): the same values, one load
      const eo_d4 k = eo_ld256(geoK + 4 * c);
      K[0][0] = k.x, K[0][1] = k.y, K[1][0] = k.z, K[1][1] = k.w;
      tab_gather<BS, NB>(dofmap, u, c, w);
      cached = true;
    }
  }
  if (!cached) tab_load_cell<GDIM, BS, NB>(T, dofmap, x_dofmap, x, u, c, w, K);
  const int ncomp = tab_ncomp(kind, BS, GDIM);
  const bool vec4 = ncomp == 4 && (reinterpret_cast<uintptr_t>(out) % 32) == 0;
  double* o = out + i * int64_t(T.nq) * ncomp;
  for (int q = 0; q < T.nq; ++q) {
    double val[BS], grad[BS][GDIM], r[BS * GDIM > 4 ? BS * GDIM : 4];
    tab_point<GDIM, BS, NB>(T, w, K, q, kind == 0, kind != 0, val, grad);
    tab_operand<GDIM, BS>(kind, val, grad, r);
    if (vec4) {
      eo_st256(o + 4 * q, r[0], r[1], r[2], r[3]);
    } else {
      for (int k = 0; k < ncomp; ++k) eo_st64(o + q * ncomp + k, r[k]);
    }
  }
}

// K = J^-1 and |det J| of every cell, once per handle: exactly the values tab_geometry / form_geometry_xv give the
// kernels that compute them per point (same statements, this file and form.cu are compiled with -fmad=false)
__global__ void __launch_bounds__(256) tab_geometry_kernel(const __grid_constant__ tab_tables T,
                                                           const int32_t* __restrict__ x_dofmap,
                                                           const double* __restrict__ x, int64_t n_cells,
                                                           double* __restrict__ geoK, double* __restrict__ geoD) {
  const int64_t c = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
  if (c >= n_cells) return;
  double xv[3][2], K[2][2];
#pragma unroll
  for (int v = 0; v < 3; ++v) {
    const int32_t node = __ldg(x_dofmap + c * 3 + v);
#pragma unroll
    for (int i = 0; i < 2; ++i) xv[v][i] = __ldg(x + 3 * int64_t(node) + i);
  }
  const double adet = form_geometry_xv<2>(T, xv, K);
  eo_st256(geoK + 4 * c, K[0][0], K[0][1], K[1][0], K[1][1]);
  eo_st64(geoD + c, adet);
}

int eo_tab_geometry(eo_tab* t) {
  static const bool on = [] { const char* e = getenv("EO_GEOM_CACHE"); return !(e && *e == '0'); }();
  if (!on || t->geoK || t->geo_failed || t->T.gdim != 2 || t->n_cells == 0) return EO_OK;
  eo_ctx* ctx = t->ctx;
  EO_CUDA(ctx, cudaSetDevice(ctx->device));
  double *k = nullptr, *d = nullptr;
  if (cudaMalloc(&k, size_t(t->n_cells) * 32) != cudaSuccess || cudaMalloc(&d, size_t(t->n_cells) * 8) != cudaSuccess) {
    cudaGetLastError();  // no room for the cache: the kernels compute the geometry per point (not tried again)
    if (k) cudaFree(k);
    t->geo_failed = true;
    return EO_OK;
  }
  tab_geometry_kernel<<<unsigned((t->n_cells + 255) / 256), 256, 0, ctx->s_cmp>>>(t->T, t->x_dofmap, t->x, t->n_cells, k, d);
  EO_CUDA(ctx, cudaGetLastError());
  t->geoK = k, t->geoD = d;
  return EO_OK;
}

// tabulate the Mandel strain and feed it to the von Mises radial return.
// Mapping: one thread per QUADRATURE POINT (cell = i / nq), so that the per-point streams (history in, tangent /
// stress / dp out) are accessed exactly like in vm_kernel - consecutive threads, consecutive 32-byte records.
// The nq threads of a cell gather the same coefficients (same sectors: one L1 request) and each contracts them
// with the derivative-table row of its own point, staged in shared memory (the row index is not warp uniform,
// which the constant bank would serialise).
// QUAD: the tangent leaves as whole 128-byte lines (eo_st_tangent_quad); all 32 lanes stay in the loop, lanes past the
// end recompute the last point and store nothing.
template <int NB, int NQ, bool EXACT, bool QUAD>  // NQ > 0: evaluation points per cell known at compile time (cheap index split)
__global__ void __launch_bounds__(256, 4) tab_vm_kernel(const __grid_constant__ tab_tables T, const vm_consts vq,
                                                     const int32_t* __restrict__ dofmap,
                                                     const int32_t* __restrict__ x_dofmap, const double* __restrict__ x,
                                                     const double* __restrict__ u, const double* __restrict__ geoK,
                                                     int64_t n_points, const double* __restrict__ sigma_n,
                                                     const double* __restrict__ p,
                                                     double* __restrict__ C_tang, double* __restrict__ sigma,
                                                     double* __restrict__ dp_out, double* __restrict__ strain_out,
                                                     eo_stats* stats) {
  __shared__ double s_dphi[EO_TAB_MAX_NQ][2][NB];
  for (int t = threadIdx.x; t < T.nq * 2 * NB; t += blockDim.x) {
    const int q = t / (2 * NB), k = (t / NB) % 2, a = t % NB;
    s_dphi[q][k][a] = T.dphi[k][q][a];
  }
  __syncthreads();
  // grid-stride over 256-point tiles: the table staging and its barrier are paid once per CTA, not once per tile
  int plastic = 0;
  for (int64_t i0 = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; (QUAD ? (i0 & ~int64_t(31)) : i0) < n_points;
       i0 += int64_t(gridDim.x) * blockDim.x) {
    const bool live = !QUAD || i0 < n_points;
    const int64_t i = live ? i0 : n_points - 1;
    int64_t c;
    int q;
    if (NQ > 0) {
      c = i / NQ;
      q = int(i - c * NQ);
    } else if (i < 2147483647LL) {
      const unsigned ci = unsigned(i) / unsigned(T.nq);
      c = ci;
      q = int(unsigned(i) - ci * unsigned(T.nq));
    } else {
      c = i / T.nq;
      q = int(i - c * T.nq);
    }
    // history first: these loads are in flight while the gather and the contraction run
    const eo_d4 s = eo_ld256(sigma_n + 4 * i);
    const double pi = eo_ld64(p + i);
    double w[NB][2], K[2][2];
    if (geoK) {  // warp uniform
      const eo_d4 k = eo_ld256(geoK + 4 * c);
      K[0][0] = k.x, K[0][1] = k.y, K[1][0] = k.z, K[1][1] = k.w;
      tab_gather<2, NB>(dofmap, u, c, w);
    } else {
      tab_load_cell<2, 2, NB>(T, dofmap, x_dofmap, x, u, c, w, K);
    }
    double G[2][2], grad[2][2], val[2] = {0.0, 0.0}, e[4];
#pragma unroll
    for (int cc = 0; cc < 2; ++cc)
#pragma unroll
      for (int k = 0; k < 2; ++k) {
        double acc = 0.0;
#pragma unroll
        for (int a = 0; a < NB; ++a) acc += w[a][cc] * s_dphi[q][k][a];
        G[cc][k] = acc;
      }
#pragma unroll
    for (int cc = 0; cc < 2; ++cc)
#pragma unroll
      for (int j = 0; j < 2; ++j) grad[cc][j] = G[cc][0] * K[0][j] + G[cc][1] * K[1][j];
    tab_operand<2, 2>(2, val, grad, e);
    vm_point_out o;
    if (EXACT)
      vm_point(vq, e[0], e[1], e[2], e[3], s.x, s.y, s.z, s.w, pi, o);
    else
      vm_point_fast(vq, e[0], e[1], e[2], e[3], s.x, s.y, s.z, s.w, pi, o);
    if (QUAD) {
      const int64_t iq = i0 & ~int64_t(3);  // first point of the lane quad
      const int64_t left = n_points - iq;
      eo_st_tangent_quad(C_tang + 16 * iq, left >= 4 ? 4 : (left > 0 ? int(left) : 0), o.C);
      if (!live) continue;
    } else {
      double* Ct = C_tang + 16 * i;
      eo_st256(Ct + 0, o.C[0], o.C[1], o.C[2], o.C[3]);
      eo_st256(Ct + 4, o.C[4], o.C[5], o.C[6], o.C[7]);
      eo_st256(Ct + 8, o.C[8], o.C[9], o.C[10], o.C[11]);
      eo_st256(Ct + 12, o.C[12], o.C[13], o.C[14], o.C[15]);
    }
    plastic += o.dp > 0.0;
    eo_st256(sigma + 4 * i, o.g[0], o.g[1], o.g[2], o.g[3]);
    eo_st64(dp_out + i, o.dp);
    if (strain_out) eo_st256(strain_out + 4 * i, e[0], e[1], e[2], e[3]);
  }
  eo_block_sum_add(reinterpret_cast<unsigned long long*>(&stats->n_plastic), plastic);
  if (blockIdx.x == 0 && threadIdx.x == 0)
    atomicAdd(reinterpret_cast<unsigned long long*>(&stats->n_points), (unsigned long long)n_points);
}

template <int GDIM, int BS, int NB>
static void tab_launch_t(eo_tab* t, int kind, const double* u, const int32_t* cells, int64_t n, double* out) {
  const unsigned grid = (unsigned)((n + 127) / 128);
  tab_kernel<GDIM, BS, NB><<<grid, 128, 0, t->ctx->s_cmp>>>(t->T, kind, t->dofmap, t->x_dofmap, t->x, u, t->geoK, cells, n, out);
}

static int tab_launch(eo_tab* t, int kind, const double* u, const int32_t* cells, int64_t n, double* out) {
  const int g = t->T.gdim, b = t->T.bs, nb = t->T.nb;
  if (kind != 0 && n > 0) {  // gradients need K: cache it on the handle at the first use
    const int grc = eo_tab_geometry(t);
    if (grc != EO_OK) return grc;
  }
#define EO_TAB_CASE(G, B, N)                      \
  if (g == G && b == B && nb == N) {              \
    tab_launch_t<G, B, N>(t, kind, u, cells, n, out); \
    t->ctx->launches += 1;                        \
    return EO_OK;                                 \
  }
  EO_TAB_CASE(2, 1, 3)   // P1 scalar triangle (heat demos)
  EO_TAB_CASE(2, 1, 6)   // P2 scalar triangle
  EO_TAB_CASE(2, 2, 3)   // P1 vector triangle
  EO_TAB_CASE(2, 2, 6)   // P2 vector triangle (von Mises, Mohr-Coulomb, hyperelasticity demos)
  EO_TAB_CASE(2, 1, 10)  // P3 scalar triangle
  EO_TAB_CASE(2, 2, 10)  // P3 vector triangle
  EO_TAB_CASE(3, 1, 4)   // P1 scalar tetrahedron
  EO_TAB_CASE(3, 3, 4)   // P1 vector tetrahedron
  EO_TAB_CASE(3, 1, 10)  // P2 scalar tetrahedron
  EO_TAB_CASE(3, 3, 10)  // P2 vector tetrahedron
#undef EO_TAB_CASE
  return eo_fail(t->ctx, EO_ERR_UNSUPPORTED, "eo_tabulate: no kernel for gdim=%d bs=%d nb=%d", g, b, nb);
}

// device pointer for a coefficient vector given on either side
int eo_tab_stage_u(eo_tab* t, const double* u, const double** d_u, int slot) {
  eo_ctx* ctx = t->ctx;
  EO_REQUIRE(ctx, slot >= 0 && slot < EO_JIT_MAX_ARGS, "eo_tab_stage_u: staging slot out of range");
  if (eo_is_device_ptr(u)) {
    *d_u = u;
    return EO_OK;
  }
  const size_t bytes = size_t(t->n_dofs) * t->T.bs * sizeof(double);
  if (!t->u_stage[slot]) EO_CUDA(ctx, cudaMalloc(&t->u_stage[slot], bytes ? bytes : 8));
  EO_CUDA(ctx, cudaMemcpyAsync(t->u_stage[slot], u, bytes, cudaMemcpyHostToDevice, ctx->s_cmp));
  *d_u = t->u_stage[slot];
  return EO_OK;
}

// internal: what the fused generic path (jit.cu) needs from a tabulation handle; stages a host coefficient vector
int eo_tab_view_get(eo_tab* t, const double* u, eo_tab_view* v, int slot) {
  v->ctx = t->ctx;
  v->T_host = &t->T;
  v->T_dev = t->d_T;
  v->dofmap = t->dofmap, v->x_dofmap = t->x_dofmap, v->x = t->x;
  v->n_cells = t->n_cells, v->n_dofs = t->n_dofs;
  v->u = nullptr;
  if (!u) return EO_OK;
  return eo_tab_stage_u(t, u, &v->u, slot);
}

extern "C" {

int eo_tab_create(eo_ctx* ctx, const eo_tab_desc* d, eo_tab** out) {
  EO_REQUIRE(ctx, ctx != nullptr, "eo_tab_create: ctx is NULL");
  EO_REQUIRE(ctx, d && out, "eo_tab_create: NULL argument");
  *out = nullptr;
  EO_REQUIRE(ctx, d->gdim == 2 || d->gdim == 3, "eo_tab_create: gdim must be 2 or 3 (affine simplex cells)");
  EO_REQUIRE(ctx, d->bs >= 1 && d->bs <= EO_TAB_MAX_BS, "eo_tab_create: block size out of range");
  EO_REQUIRE(ctx, d->nb >= 1 && d->nb <= EO_TAB_MAX_NB, "eo_tab_create: too many basis functions per cell");
  EO_REQUIRE(ctx, d->nq >= 1 && d->nq <= EO_TAB_MAX_NQ, "eo_tab_create: too many evaluation points per cell");
  EO_REQUIRE(ctx, d->n_cells >= 0 && d->n_dofs >= 0 && d->n_nodes >= 0, "eo_tab_create: negative size");
  EO_REQUIRE(ctx, d->dofmap && d->x_dofmap && d->x && d->phi && d->dphi && d->dpsi, "eo_tab_create: NULL array");
  EO_CUDA(ctx, cudaSetDevice(ctx->device));
  const int nv = d->gdim + 1;
  // validate the index arrays once, on the host: the kernels trust them
  for (int64_t i = 0; i < d->n_cells * d->nb; ++i)
    if (d->dofmap[i] < 0 || d->dofmap[i] >= d->n_dofs) return eo_fail(ctx, EO_ERR_INVALID, "eo_tab_create: dofmap entry out of range");
  for (int64_t i = 0; i < d->n_cells * nv; ++i)
    if (d->x_dofmap[i] < 0 || d->x_dofmap[i] >= d->n_nodes) return eo_fail(ctx, EO_ERR_INVALID, "eo_tab_create: x_dofmap entry out of range");
  eo_tab* t = new eo_tab();
  t->ctx = ctx;
  t->n_cells = d->n_cells, t->n_dofs = d->n_dofs, t->n_nodes = d->n_nodes;
  memset(&t->T, 0, sizeof(t->T));
  t->T.nb = d->nb, t->T.nq = d->nq, t->T.bs = d->bs, t->T.gdim = d->gdim, t->T.nv = nv;
  for (int q = 0; q < d->nq; ++q)
    for (int a = 0; a < d->nb; ++a) {
      t->T.phi[q][a] = d->phi[q * d->nb + a];
      for (int k = 0; k < d->gdim; ++k) t->T.dphi[k][q][a] = d->dphi[(k * d->nq + q) * d->nb + a];
    }
  for (int k = 0; k < d->gdim; ++k)
    for (int v = 0; v < nv; ++v) t->T.dpsi[k][v] = d->dpsi[k * nv + v];
  auto fail = [&](cudaError_t e, const char* what) {
    eo_fail(ctx, e == cudaErrorMemoryAllocation ? EO_ERR_NOMEM : EO_ERR_CUDA, "eo_tab_create: %s: %s", what, cudaGetErrorString(e));
    if (t->dofmap) cudaFree(t->dofmap);
    if (t->x_dofmap) cudaFree(t->x_dofmap);
    if (t->x) cudaFree(t->x);
    if (t->d_T) cudaFree(t->d_T);
    const int rc = e == cudaErrorMemoryAllocation ? EO_ERR_NOMEM : EO_ERR_CUDA;
    delete t;
    return rc;
  };
  cudaError_t e;
  const size_t b_dm = size_t(d->n_cells) * d->nb * 4, b_xd = size_t(d->n_cells) * nv * 4, b_x = size_t(d->n_nodes) * 3 * 8;
  if ((e = cudaMalloc(&t->dofmap, b_dm ? b_dm : 4)) != cudaSuccess) return fail(e, "cudaMalloc(dofmap)");
  if ((e = cudaMalloc(&t->x_dofmap, b_xd ? b_xd : 4)) != cudaSuccess) return fail(e, "cudaMalloc(x_dofmap)");
  if ((e = cudaMalloc(&t->x, b_x ? b_x : 8)) != cudaSuccess) return fail(e, "cudaMalloc(x)");
  if ((e = cudaMemcpyAsync(t->dofmap, d->dofmap, b_dm, cudaMemcpyDefault, ctx->s_cmp)) != cudaSuccess) return fail(e, "copy dofmap");
  if ((e = cudaMemcpyAsync(t->x_dofmap, d->x_dofmap, b_xd, cudaMemcpyDefault, ctx->s_cmp)) != cudaSuccess) return fail(e, "copy x_dofmap");
  if ((e = cudaMemcpyAsync(t->x, d->x, b_x, cudaMemcpyDefault, ctx->s_cmp)) != cudaSuccess) return fail(e, "copy x");
  if ((e = cudaMalloc(&t->d_T, sizeof(tab_tables))) != cudaSuccess) return fail(e, "cudaMalloc(tables)");
  if ((e = cudaMemcpyAsync(t->d_T, &t->T, sizeof(tab_tables), cudaMemcpyHostToDevice, ctx->s_cmp)) != cudaSuccess) return fail(e, "copy tables");
  if ((e = cudaStreamSynchronize(ctx->s_cmp)) != cudaSuccess) return fail(e, "sync");
  *out = t;
  return EO_OK;
}

int eo_tab_destroy(eo_tab* t) {
  if (!t) return EO_OK;
  cudaSetDevice(t->ctx->device);
  cudaStreamSynchronize(t->ctx->s_cmp);
  if (t->dofmap) cudaFree(t->dofmap);
  if (t->x_dofmap) cudaFree(t->x_dofmap);
  if (t->x) cudaFree(t->x);
  if (t->d_T) cudaFree(t->d_T);
  for (double* s : t->u_stage)
    if (s) cudaFree(s);
  if (t->cells_stage) cudaFree(t->cells_stage);
  if (t->geoK) cudaFree(t->geoK);
  if (t->geoD) cudaFree(t->geoD);
  delete t;
  return EO_OK;
}

int eo_tab_ncomp(const eo_tab* t, int kind) {
  if (!t || kind < 0 || kind > 3) return EO_ERR_INVALID;
  if (kind == EO_OPERAND_MANDEL_STRAIN && !(t->T.gdim == 2 && t->T.bs == 2)) return EO_ERR_INVALID;
  if (kind == EO_OPERAND_DEF_GRAD && t->T.bs != t->T.gdim) return EO_ERR_INVALID;
  return tab_ncomp(kind, t->T.bs, t->T.gdim);
}

int eo_tabulate(eo_tab* t, int kind, const double* u, const int32_t* cells, int64_t n_cells, double* out) {
  if (!t) return EO_ERR_INVALID;
  eo_ctx* ctx = t->ctx;
  EO_REQUIRE(ctx, kind >= 0 && kind <= 3, "eo_tabulate: unknown operand kind");
  EO_REQUIRE(ctx, eo_tab_ncomp(t, kind) > 0, "eo_tabulate: operand kind does not fit this element (Mandel strain needs a 2-d vector field, F a square gradient)");
  EO_REQUIRE(ctx, n_cells >= 0, "eo_tabulate: n_cells < 0");
  EO_REQUIRE(ctx, cells != nullptr || n_cells <= t->n_cells, "eo_tabulate: more cells requested than the mesh has");
  if (n_cells == 0) return EO_OK;
  EO_REQUIRE(ctx, u && out, "eo_tabulate: NULL array");
  EO_CUDA(ctx, cudaSetDevice(ctx->device));
  const double* d_u = nullptr;
  int rc = eo_tab_stage_u(t, u, &d_u);
  if (rc != EO_OK) return rc;
  const int32_t* d_cells = cells;
  if (cells && !eo_is_device_ptr(cells)) {
    for (int64_t i = 0; i < n_cells; ++i)
      if (cells[i] < 0 || cells[i] >= t->n_cells) return eo_fail(ctx, EO_ERR_INVALID, "eo_tabulate: entity index out of range");
    if (t->cells_stage_n < size_t(n_cells)) {
      EO_CUDA(ctx, cudaStreamSynchronize(ctx->s_cmp));
      if (t->cells_stage) cudaFree(t->cells_stage);
      t->cells_stage = nullptr, t->cells_stage_n = 0;
      EO_CUDA(ctx, cudaMalloc(&t->cells_stage, size_t(n_cells) * 4));
      t->cells_stage_n = size_t(n_cells);
    }
    EO_CUDA(ctx, cudaMemcpyAsync(t->cells_stage, cells, size_t(n_cells) * 4, cudaMemcpyHostToDevice, ctx->s_cmp));
    d_cells = t->cells_stage;
  }
  const size_t out_bytes = size_t(n_cells) * t->T.nq * eo_tab_ncomp(t, kind) * sizeof(double);
  if (eo_is_device_ptr(out)) {
    rc = tab_launch(t, kind, d_u, d_cells, n_cells, out);
    if (rc != EO_OK) return rc;
    EO_CUDA(ctx, cudaGetLastError());
    return EO_OK;
  }
  // host result: tabulate into the staging arena, then one D2H
  if (out_bytes > ctx->arena_bytes) {
    EO_CUDA(ctx, cudaStreamSynchronize(ctx->s_cmp));
    if (ctx->arena) cudaFree(ctx->arena);
    ctx->arena = nullptr, ctx->arena_bytes = 0;
    EO_CUDA(ctx, cudaMalloc(&ctx->arena, out_bytes));
    ctx->arena_bytes = out_bytes;
  }
  rc = tab_launch(t, kind, d_u, d_cells, n_cells, reinterpret_cast<double*>(ctx->arena));
  if (rc != EO_OK) return rc;
  EO_CUDA(ctx, cudaGetLastError());
  EO_CUDA(ctx, cudaMemcpyAsync(out, ctx->arena, out_bytes, cudaMemcpyDeviceToHost, ctx->s_cmp));
  EO_CUDA(ctx, cudaStreamSynchronize(ctx->s_cmp));
  return EO_OK;
}

int eo_tab_vm_fused(eo_tab* t, const eo_vm_params* prm, const double* u, const double* sigma_n, const double* p,
                    double* C_tang, double* sigma, double* dp, double* strain, int exact) {
  if (!t) return EO_ERR_INVALID;
  eo_ctx* ctx = t->ctx;
  EO_REQUIRE(ctx, prm != nullptr, "eo_tab_vm_fused: prm is NULL");
  EO_REQUIRE(ctx, t->T.gdim == 2 && t->T.bs == 2, "eo_tab_vm_fused: needs a 2-d vector field (plane-strain Mandel strain)");
  EO_REQUIRE(ctx, t->T.nb == 3 || t->T.nb == 6 || t->T.nb == 10, "eo_tab_vm_fused: P1/P2/P3 triangles only");
  if (t->n_cells == 0) return EO_OK;
  EO_REQUIRE(ctx, u && sigma_n && p && C_tang && sigma && dp, "eo_tab_vm_fused: NULL array");
  EO_REQUIRE(ctx, eo_is_device_ptr(sigma_n) && eo_is_device_ptr(p) && eo_is_device_ptr(C_tang) && eo_is_device_ptr(sigma) &&
                      eo_is_device_ptr(dp) && (!strain || eo_is_device_ptr(strain)),
             "eo_tab_vm_fused: history and outputs must be device memory (the coefficient vector may be host memory)");
  EO_REQUIRE(ctx, eo_aligned(sigma_n, 32) && eo_aligned(C_tang, 32) && eo_aligned(sigma, 32) && (!strain || eo_aligned(strain, 32)),
             "eo_tab_vm_fused: arrays must be 32-byte aligned");
  EO_CUDA(ctx, cudaSetDevice(ctx->device));
  const double* d_u = nullptr;
  int rc = eo_tab_stage_u(t, u, &d_u);
  if (rc != EO_OK) return rc;
  rc = eo_tab_geometry(t);
  if (rc != EO_OK) return rc;
  const vm_consts q{prm->lmbda, prm->mu, prm->H, prm->sigma_0};
  const int64_t n_points = t->n_cells * t->T.nq;
  const int64_t tiles = (n_points + 255) / 256;
  const int64_t cap = int64_t(ctx->sm_count) * 4 * TAB_FUSED_WAVES;  // 4 resident CTAs per SM x a few waves each
  const unsigned grid = (unsigned)(tiles < cap ? tiles : cap);
  // EO_QUAD_STORE=0: per-thread 128-byte tangent records (A/B; same values)
  static const bool quad_env = [] { const char* e = getenv("EO_QUAD_STORE"); return !(e && *e == '0'); }();
  const bool quad = quad_env && eo_aligned(C_tang, 128);
#define EO_FUSED_LAUNCH(N, Q, X)                                                                                          \
  do {                                                                                                                    \
    if (quad)                                                                                                             \
      tab_vm_kernel<N, Q, X, true><<<grid, 256, 0, ctx->s_cmp>>>(t->T, q, t->dofmap, t->x_dofmap, t->x, d_u, t->geoK,     \
                                                                 n_points, sigma_n, p, C_tang, sigma, dp, strain,         \
                                                                 ctx->stats);                                             \
    else                                                                                                                  \
      tab_vm_kernel<N, Q, X, false><<<grid, 256, 0, ctx->s_cmp>>>(t->T, q, t->dofmap, t->x_dofmap, t->x, d_u, t->geoK,    \
                                                                  n_points, sigma_n, p, C_tang, sigma, dp, strain,        \
                                                                  ctx->stats);                                            \
  } while (0)
#define EO_FUSED_CASE(N)                          \
  if (t->T.nb == N) {                             \
    if (t->T.nq == 3 && exact) EO_FUSED_LAUNCH(N, 3, true);   \
    else if (t->T.nq == 3) EO_FUSED_LAUNCH(N, 3, false);      \
    else if (exact) EO_FUSED_LAUNCH(N, 0, true);  \
    else EO_FUSED_LAUNCH(N, 0, false);            \
  }
  EO_FUSED_CASE(3)
  EO_FUSED_CASE(6)
  EO_FUSED_CASE(10)
#undef EO_FUSED_LAUNCH
#undef EO_FUSED_CASE
  ctx->launches += 1;
  EO_CUDA(ctx, cudaGetLastError());
  return EO_OK;
}

}  // extern "C"
