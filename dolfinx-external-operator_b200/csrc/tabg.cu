// tabg.cu - general operand tabulation on sm_100a: any Lagrange-type element given by its tables (P1..P4
// simplices, Q1/Q2 quadrilaterals and hexahedra, ...), non-affine geometry (Jacobian per evaluation point from the
// geometry element's derivative tables) and codimension-1 entities ((cell, local facet) pairs, one table set per
// local facet).
// replaces: `expr.eval(operand_mesh, entities)`, src/dolfinx_external_operator/external_operator.py:365-402, beyond
//           the affine-simplex fast path of tab.cu: `entities` of shape (n,) or (n, 2)
//           (test/test_codim_external_operator.py:75-109,167), any element family/degree/cell type
//           (test/test_external_operators_evaluation.py, test/test_nested_ex_op.py:93-104).
//
// Mapping: one thread per (entity, evaluation point); sizes are run-time values (only gdim and the block size
// are template parameters), so nothing but the accumulators lives in registers:
//   J[i][j]  = sum_v x[x_dofmap[cell][v]][i] * dgeo[set][j][q][v]      K = J^-1        (per point: non-affine cells)
//   val[c]   = sum_a u[bs * dofmap[cell][a] + c] * phi[set][q][a]
//   G[c][k]  = sum_a u[...]                      * dphi[set][k][q][a]   grad = G K
// The nq threads of an entity gather the same coefficients (same sectors, one L1 request each); the table sets are
// staged in shared memory when they fit (48 KB), otherwise read through the read-only path.
#include "eo_common.cuh"
#include "tab_core.cuh"

struct eo_gtab {
  eo_ctx* ctx = nullptr;
  int gdim = 0, bs = 0, nb = 0, nq = 0, ng = 0, n_sets = 0;
  int64_t n_cells = 0, n_dofs = 0, n_nodes = 0;
  int32_t* dofmap = nullptr;    // device [n_cells][nb]
  int32_t* x_dofmap = nullptr;  // device [n_cells][ng]
  double* x = nullptr;          // device [n_nodes][3]
  double* tables = nullptr;     // device: phi [n_sets][nq][nb] | dphi [n_sets][gdim][nq][nb] | dgeo [n_sets][gdim][nq][ng]
  size_t table_doubles = 0;
  double* u_stage = nullptr;
  int32_t* ent_stage = nullptr;
  size_t ent_stage_n = 0;
};

struct gtab_dims {
  int nb, nq, ng, n_sets, width;  // width: 0 = all cells, 1 = cell list, 2 = (cell, local entity) pairs
};

template <int GDIM, int BS, bool SMEM>
__global__ void __launch_bounds__(256) gtab_kernel(const gtab_dims D, int kind, const int32_t* __restrict__ dofmap,
                                                   const int32_t* __restrict__ x_dofmap, const double* __restrict__ x,
                                                   const double* __restrict__ tables, int table_doubles,
                                                   const double* __restrict__ u, const int32_t* __restrict__ entities,
                                                   int64_t n_points, double* __restrict__ out) {
  extern __shared__ double s_tab[];
  const double* tab = tables;
  if (SMEM) {
    for (int t = threadIdx.x; t < table_doubles; t += blockDim.x) s_tab[t] = __ldg(tables + t);
    __syncthreads();
    tab = s_tab;
  }
  const int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
  if (i >= n_points) return;
  const int64_t e = i / D.nq;
  const int q = int(i - e * D.nq);
  int64_t cell = e;
  int set = 0;
  if (D.width == 1) cell = __ldg(entities + e);
  if (D.width == 2) {
    cell = __ldg(entities + 2 * e);
    set = __ldg(entities + 2 * e + 1);
  }
  const double* phi = tab + (size_t(set) * D.nq + q) * D.nb;                                  // [a]
  const double* dphi = tab + size_t(D.n_sets) * D.nq * D.nb + size_t(set) * GDIM * D.nq * D.nb;  // [k][q][a]
  const double* dgeo = tab + size_t(D.n_sets) * D.nq * D.nb * (1 + GDIM) + size_t(set) * GDIM * D.nq * D.ng;  // [k][q][v]

  // Jacobian at this point
  double J[GDIM][GDIM], K[GDIM][GDIM];
#pragma unroll
  for (int a = 0; a < GDIM; ++a)
#pragma unroll
    for (int b = 0; b < GDIM; ++b) J[a][b] = 0.0;
  for (int v = 0; v < D.ng; ++v) {
    const int64_t node = __ldg(x_dofmap + cell * D.ng + v);
    double xv[GDIM];
#pragma unroll
    for (int a = 0; a < GDIM; ++a) xv[a] = __ldg(x + 3 * node + a);
#pragma unroll
    for (int b = 0; b < GDIM; ++b) {
      const double d = SMEM ? dgeo[(b * D.nq + q) * D.ng + v] : __ldg(dgeo + (b * D.nq + q) * D.ng + v);
#pragma unroll
      for (int a = 0; a < GDIM; ++a) J[a][b] += xv[a] * d;
    }
  }
  tab_inverse<GDIM>(J, K);

  double val[BS], G[BS][GDIM];
#pragma unroll
  for (int c = 0; c < BS; ++c) {
    val[c] = 0.0;
#pragma unroll
    for (int k = 0; k < GDIM; ++k) G[c][k] = 0.0;
  }
  const bool want_value = kind == 0;
  for (int a = 0; a < D.nb; ++a) {
    const int64_t dof = __ldg(dofmap + cell * D.nb + a);
    double w[BS];
#pragma unroll
    for (int c = 0; c < BS; ++c) w[c] = __ldg(u + BS * dof + c);
    if (want_value) {
      const double p = SMEM ? phi[a] : __ldg(phi + a);
#pragma unroll
      for (int c = 0; c < BS; ++c) val[c] += w[c] * p;
    } else {
#pragma unroll
      for (int k = 0; k < GDIM; ++k) {
        const double d = SMEM ? dphi[(k * D.nq + q) * D.nb + a] : __ldg(dphi + (k * D.nq + q) * D.nb + a);
#pragma unroll
        for (int c = 0; c < BS; ++c) G[c][k] += w[c] * d;
      }
    }
  }
  double grad[BS][GDIM];
#pragma unroll
  for (int c = 0; c < BS; ++c)
#pragma unroll
    for (int j = 0; j < GDIM; ++j) {
      double acc = 0.0;
#pragma unroll
      for (int k = 0; k < GDIM; ++k) acc += G[c][k] * K[k][j];
      grad[c][j] = acc;
    }
  double r[BS * GDIM > 4 ? BS * GDIM : 4];
  tab_operand<GDIM, BS>(kind, val, grad, r);
  const int ncomp = tab_ncomp(kind, BS, GDIM);
  double* o = out + i * ncomp;
  if (ncomp == 4 && (reinterpret_cast<uintptr_t>(out) & 31) == 0) {
    eo_st256(o, r[0], r[1], r[2], r[3]);
  } else if (ncomp % 2 == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0) {
    for (int k = 0; k < ncomp; k += 2) eo_st128(o + k, r[k], r[k + 1]);
  } else {
    for (int k = 0; k < ncomp; ++k) eo_st64(o + k, r[k]);
  }
}

template <int GDIM, int BS>
static int gtab_launch_t(eo_gtab* t, int kind, const double* u, const int32_t* ent, int width, int64_t n_entities, double* out) {
  eo_ctx* ctx = t->ctx;
  const gtab_dims D{t->nb, t->nq, t->ng, t->n_sets, width};
  const int64_t n_points = n_entities * t->nq;
  const int64_t grid64 = (n_points + 255) / 256;
  if (grid64 > 2147483647LL) return eo_fail(ctx, EO_ERR_INVALID, "eo_gtab_tabulate: too many points for one launch");
  const size_t smem = t->table_doubles * sizeof(double);
  if (smem <= 48 * 1024)
    gtab_kernel<GDIM, BS, true><<<unsigned(grid64), 256, smem, ctx->s_cmp>>>(D, kind, t->dofmap, t->x_dofmap, t->x, t->tables,
                                                                            int(t->table_doubles), u, ent, n_points, out);
  else
    gtab_kernel<GDIM, BS, false><<<unsigned(grid64), 256, 0, ctx->s_cmp>>>(D, kind, t->dofmap, t->x_dofmap, t->x, t->tables,
                                                                          int(t->table_doubles), u, ent, n_points, out);
  ctx->launches += 1;
  return EO_OK;
}

static int gtab_launch(eo_gtab* t, int kind, const double* u, const int32_t* ent, int width, int64_t n, double* out) {
#define EO_GTAB_CASE(G, B) \
  if (t->gdim == G && t->bs == B) return gtab_launch_t<G, B>(t, kind, u, ent, width, n, out);
  EO_GTAB_CASE(2, 1) EO_GTAB_CASE(2, 2) EO_GTAB_CASE(2, 3) EO_GTAB_CASE(3, 1) EO_GTAB_CASE(3, 2) EO_GTAB_CASE(3, 3)
#undef EO_GTAB_CASE
  return eo_fail(t->ctx, EO_ERR_UNSUPPORTED, "eo_gtab_tabulate: gdim %d with block size %d is not instantiated", t->gdim, t->bs);
}

extern "C" {

int eo_gtab_create(eo_ctx* ctx, const eo_gtab_desc* d, eo_gtab** out) {
  EO_REQUIRE(ctx, ctx != nullptr, "eo_gtab_create: ctx is NULL");
  EO_REQUIRE(ctx, d && out, "eo_gtab_create: NULL argument");
  *out = nullptr;
  EO_REQUIRE(ctx, d->gdim == 2 || d->gdim == 3, "eo_gtab_create: gdim must be 2 or 3");
  EO_REQUIRE(ctx, d->bs >= 1 && d->bs <= 3, "eo_gtab_create: block size must be 1..3");
  EO_REQUIRE(ctx, d->nb >= 1 && d->nb <= 125 && d->nq >= 1 && d->nq <= 125 && d->ng >= d->gdim + 1 && d->ng <= 27,
             "eo_gtab_create: nb / nq must be 1..125, ng gdim+1..27");
  EO_REQUIRE(ctx, d->n_sets >= 1 && d->n_sets <= 6, "eo_gtab_create: n_sets must be 1 (cell points) .. 6 (facets of a hexahedron)");
  EO_REQUIRE(ctx, d->n_cells >= 0 && d->n_dofs >= 0 && d->n_nodes >= 0, "eo_gtab_create: negative size");
  EO_REQUIRE(ctx, d->dofmap && d->x_dofmap && d->x && d->phi && d->dphi && d->dgeo, "eo_gtab_create: NULL array");
  EO_CUDA(ctx, cudaSetDevice(ctx->device));
  for (int64_t i = 0; i < d->n_cells * d->nb; ++i)
    if (d->dofmap[i] < 0 || d->dofmap[i] >= d->n_dofs) return eo_fail(ctx, EO_ERR_INVALID, "eo_gtab_create: dofmap entry out of range");
  for (int64_t i = 0; i < d->n_cells * d->ng; ++i)
    if (d->x_dofmap[i] < 0 || d->x_dofmap[i] >= d->n_nodes) return eo_fail(ctx, EO_ERR_INVALID, "eo_gtab_create: x_dofmap entry out of range");
  eo_gtab* t = new eo_gtab();
  t->ctx = ctx;
  t->gdim = d->gdim, t->bs = d->bs, t->nb = d->nb, t->nq = d->nq, t->ng = d->ng, t->n_sets = d->n_sets;
  t->n_cells = d->n_cells, t->n_dofs = d->n_dofs, t->n_nodes = d->n_nodes;
  const size_t n_phi = size_t(d->n_sets) * d->nq * d->nb, n_dphi = n_phi * d->gdim, n_dgeo = size_t(d->n_sets) * d->gdim * d->nq * d->ng;
  t->table_doubles = n_phi + n_dphi + n_dgeo;
  std::vector<double> h(t->table_doubles);
  memcpy(h.data(), d->phi, n_phi * 8);
  memcpy(h.data() + n_phi, d->dphi, n_dphi * 8);
  memcpy(h.data() + n_phi + n_dphi, d->dgeo, n_dgeo * 8);
  auto fail = [&](cudaError_t e, const char* what) {
    const int rc = eo_fail(ctx, e == cudaErrorMemoryAllocation ? EO_ERR_NOMEM : EO_ERR_CUDA, "eo_gtab_create: %s: %s", what, cudaGetErrorString(e));
    eo_gtab_destroy(t);
    return rc;
  };
  cudaError_t e;
  const size_t b_dm = size_t(d->n_cells) * d->nb * 4, b_xd = size_t(d->n_cells) * d->ng * 4, b_x = size_t(d->n_nodes) * 3 * 8;
  if ((e = cudaMalloc(&t->dofmap, b_dm ? b_dm : 4)) != cudaSuccess) return fail(e, "cudaMalloc(dofmap)");
  if ((e = cudaMalloc(&t->x_dofmap, b_xd ? b_xd : 4)) != cudaSuccess) return fail(e, "cudaMalloc(x_dofmap)");
  if ((e = cudaMalloc(&t->x, b_x ? b_x : 8)) != cudaSuccess) return fail(e, "cudaMalloc(x)");
  if ((e = cudaMalloc(&t->tables, t->table_doubles * 8)) != cudaSuccess) return fail(e, "cudaMalloc(tables)");
  if ((e = cudaMemcpyAsync(t->dofmap, d->dofmap, b_dm, cudaMemcpyDefault, ctx->s_cmp)) != cudaSuccess) return fail(e, "copy dofmap");
  if ((e = cudaMemcpyAsync(t->x_dofmap, d->x_dofmap, b_xd, cudaMemcpyDefault, ctx->s_cmp)) != cudaSuccess) return fail(e, "copy x_dofmap");
  if ((e = cudaMemcpyAsync(t->x, d->x, b_x, cudaMemcpyDefault, ctx->s_cmp)) != cudaSuccess) return fail(e, "copy x");
  if ((e = cudaMemcpyAsync(t->tables, h.data(), t->table_doubles * 8, cudaMemcpyHostToDevice, ctx->s_cmp)) != cudaSuccess) return fail(e, "copy tables");
  if ((e = cudaStreamSynchronize(ctx->s_cmp)) != cudaSuccess) return fail(e, "sync");
  *out = t;
  return EO_OK;
}

int eo_gtab_destroy(eo_gtab* t) {
  if (!t) return EO_OK;
  cudaSetDevice(t->ctx->device);
  cudaStreamSynchronize(t->ctx->s_cmp);
  if (t->dofmap) cudaFree(t->dofmap);
  if (t->x_dofmap) cudaFree(t->x_dofmap);
  if (t->x) cudaFree(t->x);
  if (t->tables) cudaFree(t->tables);
  if (t->u_stage) cudaFree(t->u_stage);
  if (t->ent_stage) cudaFree(t->ent_stage);
  delete t;
  return EO_OK;
}

int eo_gtab_ncomp(const eo_gtab* t, int kind) {
  if (!t || kind < 0 || kind > 3) return EO_ERR_INVALID;
  if (kind == EO_OPERAND_MANDEL_STRAIN && !(t->gdim == 2 && t->bs == 2)) return EO_ERR_INVALID;
  if (kind == EO_OPERAND_DEF_GRAD && t->bs != t->gdim) return EO_ERR_INVALID;
  return tab_ncomp(kind, t->bs, t->gdim);
}

int eo_gtab_tabulate(eo_gtab* t, int kind, const double* u, const int32_t* entities, int entity_width, int64_t n_entities,
                     double* out) {
  if (!t) return EO_ERR_INVALID;
  eo_ctx* ctx = t->ctx;
  EO_REQUIRE(ctx, eo_gtab_ncomp(t, kind) > 0, "eo_gtab_tabulate: operand kind unknown or unfit for this element");
  EO_REQUIRE(ctx, n_entities >= 0, "eo_gtab_tabulate: n_entities < 0");
  EO_REQUIRE(ctx, entity_width >= 0 && entity_width <= 2, "eo_gtab_tabulate: entity_width must be 0 (all cells), 1 (cells) or 2 ((cell, local entity) pairs)");
  EO_REQUIRE(ctx, (entity_width == 0) == (entities == nullptr), "eo_gtab_tabulate: entities must be NULL exactly when entity_width is 0");
  EO_REQUIRE(ctx, entity_width != 0 || n_entities <= t->n_cells, "eo_gtab_tabulate: more cells requested than the mesh has");
  EO_REQUIRE(ctx, entity_width == 2 || t->n_sets == 1, "eo_gtab_tabulate: tables were built per local facet: entities must be (cell, local facet) pairs");
  if (n_entities == 0) return EO_OK;
  EO_REQUIRE(ctx, u && out, "eo_gtab_tabulate: NULL array");
  EO_CUDA(ctx, cudaSetDevice(ctx->device));
  const double* d_u = u;
  if (!eo_is_device_ptr(u)) {
    const size_t bytes = size_t(t->n_dofs) * t->bs * sizeof(double);
    if (!t->u_stage) EO_CUDA(ctx, cudaMalloc(&t->u_stage, bytes ? bytes : 8));
    EO_CUDA(ctx, cudaMemcpyAsync(t->u_stage, u, bytes, cudaMemcpyHostToDevice, ctx->s_cmp));
    d_u = t->u_stage;
  }
  const int32_t* d_ent = entities;
  if (entities && !eo_is_device_ptr(entities)) {
    const size_t cnt = size_t(n_entities) * entity_width;
    for (int64_t i = 0; i < n_entities; ++i) {
      const int32_t c = entities[i * entity_width];
      if (c < 0 || c >= t->n_cells) return eo_fail(ctx, EO_ERR_INVALID, "eo_gtab_tabulate: entity index out of range");
      if (entity_width == 2 && (entities[2 * i + 1] < 0 || entities[2 * i + 1] >= t->n_sets))
        return eo_fail(ctx, EO_ERR_INVALID, "eo_gtab_tabulate: local entity index out of range");
    }
    if (t->ent_stage_n < cnt) {
      EO_CUDA(ctx, cudaStreamSynchronize(ctx->s_cmp));
      if (t->ent_stage) cudaFree(t->ent_stage);
      t->ent_stage = nullptr, t->ent_stage_n = 0;
      EO_CUDA(ctx, cudaMalloc(&t->ent_stage, cnt * 4));
      t->ent_stage_n = cnt;
    }
    EO_CUDA(ctx, cudaMemcpyAsync(t->ent_stage, entities, cnt * 4, cudaMemcpyHostToDevice, ctx->s_cmp));
    d_ent = t->ent_stage;
  }
  const size_t out_bytes = size_t(n_entities) * t->nq * eo_gtab_ncomp(t, kind) * sizeof(double);
  if (eo_is_device_ptr(out)) {
    int rc = gtab_launch(t, kind, d_u, d_ent, entity_width, n_entities, out);
    if (rc != EO_OK) return rc;
    EO_CUDA(ctx, cudaGetLastError());
    return EO_OK;
  }
  if (out_bytes > ctx->arena_bytes) {
    EO_CUDA(ctx, cudaStreamSynchronize(ctx->s_cmp));
    if (ctx->arena) cudaFree(ctx->arena);
    ctx->arena = nullptr, ctx->arena_bytes = 0;
    EO_CUDA(ctx, cudaMalloc(&ctx->arena, out_bytes));
    ctx->arena_bytes = out_bytes;
  }
  int rc = gtab_launch(t, kind, d_u, d_ent, entity_width, n_entities, reinterpret_cast<double*>(ctx->arena));
  if (rc != EO_OK) return rc;
  EO_CUDA(ctx, cudaGetLastError());
  EO_CUDA(ctx, cudaMemcpyAsync(out, ctx->arena, out_bytes, cudaMemcpyDeviceToHost, ctx->s_cmp));
  EO_CUDA(ctx, cudaStreamSynchronize(ctx->s_cmp));
  return EO_OK;
}

}  // extern "C"
