// tab_handle.cuh - the tabulation handle (device-resident mesh arrays + element tables), shared by tab.cu and
// form.cu (the device-side consumers integrate over the same cells with the same tables).
#pragma once
#include "eo_common.cuh"
#include "tab_core.cuh"

struct eo_tab {
  eo_ctx* ctx = nullptr;
  tab_tables T;
  int64_t n_cells = 0, n_dofs = 0, n_nodes = 0;  // n_dofs counts blocked dofs (x bs scalars)
  int32_t* dofmap = nullptr;                     // device [n_cells][nb]
  int32_t* x_dofmap = nullptr;                   // device [n_cells][nv]
  double* x = nullptr;                           // device [n_nodes][3]
  tab_tables* d_T = nullptr;                     // device copy of T (the fused generic kernels stage it in shared memory)
  double* u_stage[EO_JIT_MAX_ARGS] = {};         // device staging copies of host coefficient vectors, one per operand slot
  int32_t* cells_stage = nullptr;                // device staging copy of a host entity list
  size_t cells_stage_n = 0;
  // per-cell geometry, computed once per handle (the coordinates are fixed at creation): inverse Jacobian, row-major
  // [n_cells][4] (2-d meshes), and |det J| [n_cells]; built at the first fused / residual-step launch
  double* geoK = nullptr;
  double* geoD = nullptr;
  bool geo_failed = false;  // the allocation did not fit once: the kernels compute the geometry per point
};

// builds t->geoK / t->geoD if the handle qualifies (2-d affine cells) and EO_GEOM_CACHE is not 0; leaves them NULL
// otherwise (not an error: the kernels then compute the geometry per point from x_dofmap / x)
int eo_tab_geometry(eo_tab* t);

// device pointer for a coefficient vector given on either side (host vectors are copied into t->u_stage[slot];
// calls that tabulate several operands on one handle in ONE launch give every operand its own slot)
int eo_tab_stage_u(eo_tab* t, const double* u, const double** d_u, int slot = 0);
