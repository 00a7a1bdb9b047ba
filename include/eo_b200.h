/* eo_b200.h - C ABI of libeo_b200.so: the B200 (sm_100a) implementation of the
 * quadrature-point hot path of dolfinx-external-operator.
 *
 * The reference has no FFI of its own: its plug-in boundary is the Python
 * callable protocol `external_function(derivatives)(*operand_arrays)` used at
 * src/dolfinx_external_operator/external_operator.py:432, and the operand
 * tabulation `fem.Expression.eval` at external_operator.py:393-402.  Each entry
 * point below names the reference code it stands in for ("replaces:").  The
 * ctypes binding a maintainer would add is shown in INTEGRATION.md.
 *
 * Conventions
 *  - plain C, no C++/torch types; every function returns 0 (EO_OK) or a negative
 *    eo_status; eo_last_error(ctx) gives the text; nothing throws.
 *  - one eo_ctx = one GPU = one compute stream (+2 copy streams).  Calls on one
 *    ctx must come from one thread at a time (the reference is single threaded
 *    per MPI rank, petsc/petsc.py:60).
 *  - array arguments are "any-side" pointers unless stated: they may point to
 *    device memory (used in place, zero copy), to pinned/registered host memory
 *    or to pageable host memory (both staged through a chunked, double-buffered
 *    H2D -> kernel -> D2H pipeline).  The side is detected per pointer with
 *    cudaPointerGetAttributes.  Host-side calls are complete (data landed) on
 *    return; all-device calls are asynchronous on the ctx stream - use eo_sync.
 *  - layouts are the reference's flat C-order arrays: a field with value shape S
 *    at n quadrature points is [n][prod(S)] ("AoS", e.g. demo_vm:352).  Entry
 *    points with an explicit `layout` argument also accept EO_LAYOUT_SOA
 *    ([prod(S)][n]) for fields that stay resident in HBM.
 *  - all floating point data is IEEE binary64 ("f64"); dofmaps are int32.
 */
#ifndef EO_B200_H
#define EO_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define EO_B200_VERSION 100 /* 0.1.0 */

typedef enum eo_status {
  EO_OK = 0,
  EO_ERR_INVALID = -1,     /* bad argument (NULL, negative size, misaligned, unknown enum) */
  EO_ERR_CUDA = -2,        /* a CUDA runtime call failed; text in eo_last_error */
  EO_ERR_NOMEM = -3,       /* device or pinned-host allocation failed */
  EO_ERR_UNSUPPORTED = -4, /* valid request this build does not implement */
  EO_ERR_NO_DEVICE = -5    /* no CUDA device / not an sm_100 device */
} eo_status;

typedef enum eo_layout { EO_LAYOUT_AOS = 0, EO_LAYOUT_SOA = 1 } eo_layout;

typedef struct eo_ctx eo_ctx;

/* ---------------------------------------------------------------- lifetime */
int eo_version(void);
/* Number of visible CUDA devices, or a negative eo_status. */
int eo_device_count(void);
/* PCI bus id ("0000:1b:00.0") of CUDA device `device` - the key that identifies the same GPU to NVML (CPU / NUMA
 * affinity of a rank, parallel.bind_to_gpu_numa) whatever CUDA_VISIBLE_DEVICES / device ordering is in force. */
int eo_device_pci_bus_id(int device, char* buf, int len);
/* Create a context on `device`.  Fails with EO_ERR_NO_DEVICE when there is no
 * GPU: there is NO CPU fallback in this library. */
int eo_create(int device, eo_ctx** out);
int eo_destroy(eo_ctx* ctx);
/* Text of the last error on this ctx ("" if none).  ctx == NULL returns the
 * text of the last eo_create failure. */
const char* eo_last_error(const eo_ctx* ctx);
/* Block until all work queued on the ctx streams has finished. */
int eo_sync(eo_ctx* ctx);
/* The ctx compute stream as a cudaStream_t cast to void* (for interop). */
void* eo_stream(eo_ctx* ctx);
/* Quadrature points per pipeline chunk for host-side calls (default 1<<20). */
int eo_set_chunk(eo_ctx* ctx, int64_t n_qp_per_chunk);
/* Count of kernel launches issued by this ctx since creation. */
int64_t eo_launch_count(const eo_ctx* ctx);

/* ---------------------------------------------------------------- memory */
int eo_dev_alloc(eo_ctx* ctx, size_t bytes, void** dptr);
int eo_dev_free(eo_ctx* ctx, void* dptr);
int eo_dev_memset(eo_ctx* ctx, void* dptr, int value, size_t bytes);
/* Any-side copy (host<->device, device<->device) ordered on the ctx stream;
 * synchronous on return when either side is host memory. */
int eo_copy(eo_ctx* ctx, void* dst, const void* src, size_t bytes);
/* Pinned host memory (the callee-owned result buffers of the callable protocol). */
int eo_host_alloc(eo_ctx* ctx, size_t bytes, void** hptr);
int eo_host_free(eo_ctx* ctx, void* hptr);
/* Page-lock an existing host array in place (e.g. `ref_coefficient.x.array`,
 * external_operator.py:289-290) so the D2H lands in it at full PCIe rate. */
int eo_host_register(eo_ctx* ctx, void* hptr, size_t bytes);
int eo_host_unregister(eo_ctx* ctx, void* hptr);

/* ---------------------------------------------------------------- non-contiguous coefficient assignment
 * replaces: `_assign_non_mixed` (`x.array[unrolled_dofmap] = values`), `_assign_mixed_2d`, `_assign_mixed_3d`,
 *           external_operator.py:286-335 - continuous and mixed coefficient spaces, where several evaluation
 *           points write the same degree of freedom and NumPy's fancy assignment keeps the LAST one.
 * The host side turns the scatter inside out once per operator (for every dof the index of the value that
 * wins: `src_index`, device resident); the kernel is then the race-free gather out[k] = values[src_index[k]],
 * bit-identical to the reference's result.  values : device [n_values]; src_index : device int64 [n_out],
 * entries in [0, n_values) (checked by the host side when the plan is built); out : any-side [n_out] - e.g.
 * the host `ref_coefficient.x.array` itself, so only the compact dof array crosses PCIe. */
int eo_assign_gather(eo_ctx* ctx, const double* values, int64_t n_values, const int64_t* src_index, int64_t n_out,
                     double* out);

/* ---------------------------------------------------------------- timing
 * CUDA events on the ctx compute stream (bench.py times kernels with these). */
int eo_event_create(eo_ctx* ctx, void** ev);
int eo_event_destroy(eo_ctx* ctx, void* ev);
int eo_event_record(eo_ctx* ctx, void* ev);
int eo_event_elapsed_ms(eo_ctx* ctx, void* ev_start, void* ev_stop, float* ms);
/* Write `bytes` of device scratch (> L2) on the ctx stream: L2 flush between
 * timed iterations. */
int eo_flush_l2(eo_ctx* ctx, size_t bytes);

/* Debug: copy the 64 uint32 scheduler counters of the last persistent-kernel launch to `out` (only
 * meaningful when the library was built with -DMC_DEBUG_COUNTERS; [0] is the tile counter). */
int eo_debug_counters(eo_ctx* ctx, uint32_t* out);

/* FP64 roofline denominator for the Newton-bound kernels: runs a register-only DFMA micro-benchmark
 * (8 independent chains per thread, `iters` trips, all SMs) and returns the best of 3 timed launches
 * in TFLOP/s (2 flops per DFMA). */
int eo_fp64_peak(eo_ctx* ctx, int iters, double* tflops);
/* The same for FP32 FFMA: roofline denominator of the Isihara network kernel (float32 CUDA cores). */
int eo_fp32_peak(eo_ctx* ctx, int iters, double* tflops);
/* FP32 FMA throughput by instruction form: variant 0 = scalar FFMA with uniform multiplier / addend (what eo_fp32_peak
 * runs), 1 = scalar FFMA with three per-chain register operands, 2 = packed FFMA2 (fma.rn.f32x2, two FMAs per
 * instruction) with register-pair operands, 3 = scalar FFMA with ONE uniform operand and two registers (the shape of a
 * matrix-vector product with warp-uniform weights). */
int eo_fp32_peak_variant(eo_ctx* ctx, int iters, int variant, double* tflops);

/* ---------------------------------------------------------------- statistics
 * Device-resident record that every constitutive kernel accumulates into in its
 * epilogue (one atomic per CTA).  replaces: the host-side `jnp.unique(niter,
 * return_counts=True)`, `jnp.max(yielding)`, `jnp.max(norm_res)` of
 * demo_plasticity_mohr_coulomb.py:584-591, and is the payload of the one
 * scalar all-reduce of a multi-GPU run (eo_allreduce_stats: sums over the int64
 * block, maxima over the f64 block). */
#define EO_NITER_BINS 208 /* local-Newton iteration histogram, bins 0..Nitermax(200); padded */
typedef struct eo_stats {
  /* SUM block: 4 + EO_NITER_BINS int64 */
  int64_t n_points;       /* quadrature points evaluated */
  int64_t n_plastic;      /* vm: dp > 0 ; mc: yielding > 0 */
  int64_t n_nonconverged; /* mc: left the Newton loop by niter == Nitermax */
  int64_t n_nonfinite;    /* points whose stress is NaN/Inf */
  int64_t niter_hist[EO_NITER_BINS];
  /* MAX block: 4 f64 */
  double niter_max;
  double f_max;   /* mc: max yielding (demo_mc:590); -inf after a reset */
  double res_max; /* mc: max ||res|| (demo_mc:591) */
  double reserved;
} eo_stats;
int eo_stats_reset(eo_ctx* ctx);
/* Copies the record to host (synchronises the ctx stream). */
int eo_stats_read(eo_ctx* ctx, eo_stats* host_out);
/* Device address of the local record. */
void* eo_stats_device_ptr(eo_ctx* ctx);
/* The one collective of the hot path (quadrature points shard along the cell partition, one rank per GPU, no halo):
 * the records of all ranks of the caller's `ncclComm_t` (passed as void*) are combined - SUM over the int64 block, MAX
 * over the f64 block - into a separate GLOBAL record; the local record is left untouched, so the call can follow every
 * evaluation of a record that keeps accumulating without counting anything twice.  ONE ncclAllGather of the 1.7 KB
 * record plus a combine kernel, on the ctx's collective stream: ordered after everything queued on the compute stream
 * so far, overlapping whatever is queued afterwards; eo_stats_read_global / eo_sync wait for it.
 * replaces: the rank-local prints of demo_plasticity_mohr_coulomb.py:584-591 turned into global figures.  NCCL is
 * resolved at run time (libnccl.so.2 of the process; EO_NCCL_LIB overrides). */
int eo_allreduce_stats(eo_ctx* ctx, void* nccl_comm);
/* Copies the global record to host (synchronises the collective stream). */
int eo_stats_read_global(eo_ctx* ctx, eo_stats* host_out);
/* The same collective for callers whose communicator is not an ncclComm_t (torch.distributed process groups):
 * _begin snapshots the local record (or uploads `host_record` when it is not NULL) and returns the send buffer (one
 * record), the receive buffer (`world` records) and the collective stream (cudaStream_t as void*); the caller all-gathers
 * send -> recv on that stream; _end launches the combine kernel there. */
int eo_stats_collective_begin(eo_ctx* ctx, int world, const eo_stats* host_record, void** send, void** recv, void** stream);
int eo_stats_collective_end(eo_ctx* ctx, int world);

/* ---------------------------------------------------------------- von Mises
 * replaces: `return_mapping`/`_kernel`, doc/demo/demo_plasticity_von_mises.py:298-332,
 *           and the body of `C_tang_impl` :343-352.
 * Constants: demo_vm:185-191 (lmbda, mu from E, nu; H = E*Et/(E-Et); sigma_0). */
typedef struct eo_vm_params {
  double lmbda, mu, H, sigma_0;
} eo_vm_params;

/* deps [n][4], sigma_n [n][4], p [n]  ->  C_tang [n][4][4], sigma [n][4], dp [n].
 * Branch-free like the source; returns, like the reference, NaN where
 * sigma_eq == 0 or f == 0 exactly.  Algorithmic HBM traffic: 72 B read + 168 B
 * written per quadrature point.  The number of points with dp > 0 is added to
 * the ctx statistics record (eo_stats.n_plastic). */
int eo_vm_eval(eo_ctx* ctx, const eo_vm_params* prm, const double* deps, const double* sigma_n, const double* p,
               double* C_tang, double* sigma, double* dp, int64_t n);

/* Same computation with the history resident in HBM in the given layout
 * (device pointers only).  sigma_n/p are the committed history, sigma/dp the
 * candidates of this Newton iteration; deps and C_tang are AoS. */
int eo_vm_eval_resident(eo_ctx* ctx, const eo_vm_params* prm, const double* deps, const double* sigma_n,
                        const double* p, double* C_tang, double* sigma, double* dp, int64_t n, int state_layout);

/* History commit after a converged load step, on device.
 * replaces: `p.x.petsc_vec.axpy(1.0, dp.x.petsc_vec)` and
 *           `sigma_n.x.array[:] = sigma.ref_coefficient.x.array`, demo_vm:564-565
 *           (sigma_n <- sigma alone: demo_mc:728; pass p = dp = NULL). */
int eo_commit_history(eo_ctx* ctx, double* sigma_n, const double* sigma, double* p, const double* dp, int64_t n,
                      int ncomp);

/* ---------------------------------------------------------------- heat
 * replaces: `k_impl`/`dkdT_impl`, demo_nonlinear_heat_equation_part1.py:252-272;
 *           `q_impl`/`dqdT_impl`/`dqdsigma_impl`, part2.py:219-261 (gdim = 2).
 * T [n], sigma [n][2] (may be NULL when no q-type output is requested).
 * Outputs (each may be NULL = not requested; all requested ones are produced by
 * ONE fused kernel): k [n], dk [n], q [n][2], dqdT [n][2], dqdsigma [n][2][2]. */
int eo_heat_eval(eo_ctx* ctx, double A, double B, const double* T, const double* sigma, double* k, double* dk,
                 double* q, double* dqdT, double* dqdsigma, int64_t n);

/* ---------------------------------------------------------------- operand tabulation
 * replaces: `fem.Expression(operand, eval_points, dtype).eval(operand_mesh, entities)` inside
 *           `evaluate_operands`, src/dolfinx_external_operator/external_operator.py:386-402 (DOLFINx
 *           tabulate_expression + FFCx kernel + basix tables), for affine simplex cells (triangles,
 *           tetrahedra) and the operand expressions of the reference demos.
 * The basis tables are INPUTS (what basix `element.tabulate(1, eval_points)` returns on the host). */
typedef enum eo_operand_kind {
  EO_OPERAND_VALUE = 0,         /* f                      (T: part1.py:210, part2.py:136)               [bs]       */
  EO_OPERAND_GRAD = 1,          /* grad f, row-major      (sigma = grad T: part2.py:137,167)            [bs][gdim] */
  EO_OPERAND_MANDEL_STRAIN = 2, /* [g00, g11, 0, sqrt2/2 (g01+g10)]  (epsilon(Du): demo_vm:225-227,
                                   demo_mc:148-157); needs gdim = bs = 2                               [4]        */
  EO_OPERAND_DEF_GRAD = 3       /* I + grad u, row-major  (F: demo_hyperelasticity.py:479)              [bs][gdim] */
} eo_operand_kind;

typedef struct eo_tab_desc {
  int32_t gdim;            /* 2 (triangles) or 3 (tetrahedra); geometry is the affine P1 map              */
  int32_t bs;              /* block size of the coefficient: dof index = bs * node + comp (:18-26)        */
  int32_t nb;              /* scalar basis functions per cell                                            */
  int32_t nq;              /* evaluation points per cell (= element.interpolation_points, :200)          */
  int64_t n_cells;         /* local + ghost cells (:368-370)                                             */
  int64_t n_dofs;          /* blocked dofs of the coefficient (its array has bs * n_dofs scalars)         */
  int64_t n_nodes;         /* geometry nodes                                                             */
  const int32_t* dofmap;   /* [n_cells][nb]      V.dofmap.list                                           */
  const int32_t* x_dofmap; /* [n_cells][gdim+1]  mesh.geometry.dofmap                                    */
  const double* x;         /* [n_nodes][3]       mesh.geometry.x                                         */
  const double* phi;       /* [nq][nb]           basix tabulate(1, X)[0]                                 */
  const double* dphi;      /* [gdim][nq][nb]     basix tabulate(1, X)[1:]                                */
  const double* dpsi;      /* [gdim][gdim+1]     derivatives of the P1 geometry element (constant)       */
} eo_tab_desc;

typedef struct eo_tab eo_tab;
/* Uploads the mesh arrays and tables (they stay resident in HBM); index arrays are range-checked here. */
int eo_tab_create(eo_ctx* ctx, const eo_tab_desc* desc, eo_tab** out);
int eo_tab_destroy(eo_tab* tab);
/* Components per evaluation point of an operand kind on this element, or EO_ERR_INVALID. */
int eo_tab_ncomp(const eo_tab* tab, int kind);
/* out[n_cells][nq][ncomp] = operand at the evaluation points of `cells` (NULL: cells 0..n_cells-1, the
 * default entity list of :365-371; otherwise n_cells int32 cell indices = the `entities` argument).
 * u (bs*n_dofs), cells and out are any-side pointers; with a host `out` the call is complete on return. */
int eo_tabulate(eo_tab* tab, int kind, const double* u, const int32_t* cells, int64_t n_cells, double* out);
/* Fused hot path for all cells: Mandel strain of u at every evaluation point -> von Mises radial return
 * (eo_vm_eval) without the strain ever touching HBM.  Point index = cell * nq + q, as the reference's
 * flat layout.  sigma_n, p, C_tang, sigma, dp (and `strain`, optional, may be NULL) are device memory;
 * u is any-side.  Asynchronous on the ctx stream.
 * exact != 0: the von Mises arithmetic is the reference's statement sequence (bit-identical to eo_vm_eval on the
 * tabulated strain); exact == 0: same decision path (identical plastic/elastic flags) but two divisions instead
 * of nine and explicit FMAs downstream - a few ulp from the exact variant, ~1.3x faster (this kernel is
 * issue-bound, not HBM-bound). */
int eo_tab_vm_fused(eo_tab* tab, const eo_vm_params* prm, const double* u, const double* sigma_n, const double* p,
                    double* C_tang, double* sigma, double* dp, double* strain, int exact);

/* ---------------------------------------------------------------- device-side consumers of the operator values
 *  - what the reference hands to DOLFINx right after `evaluate_external_operators`:
 * replaces: `assemble_vector(b, F)`, petsc/petsc.py:64, for F = inner(N, OP(v)) dx with N the external operator's
 *             quadrature coefficient - inner(sigma, epsilon(v)) dx (demo_plasticity_von_mises.py:253),
 *             inner(q, grad(v)) dx (demo_nonlinear_heat_equation_part2.py:181);
 *           `assemble_matrix(A, J)`, petsc/petsc.py:88, for J = derivative(F, u, u_hat) after
 *             `replace_external_operators` = inner(dN OP_trial(u_hat), OP_test(v)) dx (demo_vm:390-398), as the
 *             ACTION on a vector (matrix-free Krylov) and as CSR values.
 * The point values stay in HBM; only DOF vectors (about 11 B per point for P2 triangles) cross the PCIe link
 * instead of the 168 B per point of the tangent.  OP_test / OP_trial are eo_operand_kind values of the element the
 * eo_tab was built for (test space = trial space = the coefficient's space; DEF_GRAD acts as GRAD).  Cells 0..n_cells-1
 * are integrated (n_cells < 0: all; DOLFINx integrates the owned cells, which come first, :368-370 lists owned +
 * ghosts).  Scatter by fp64 atomics: cell order of the sums is not fixed, results are reproducible to rounding only.
 * Vectors (bs * n_dofs doubles) are any-side pointers - with a host result the call is complete on return;
 * point-value arrays must be device memory, laid out exactly as the operator kernels write them
 * ([cell][point][comp], tangents row-major (ncomp_test, ncomp_trial)). */
typedef struct eo_form eo_form;
/* weights[nq]: quadrature weights of the reference cell at the eo_tab's evaluation points
 * (basix.make_quadrature(cell, degree)[1]); the eo_tab must outlive the form. */
int eo_form_create(eo_tab* tab, const double* weights, eo_form** out);
int eo_form_destroy(eo_form* form);
/* b (+)= sum_cells sum_q w_q |det J| coef[cell][q] . OP_test(phi_i)(x_q)      accumulate == 0: b is zeroed first */
int eo_form_vector(eo_form* form, int kind_test, const double* coef, int64_t n_cells, double* b, int accumulate);
/* y (+)= A x with A_ij = sum_cells sum_q w_q |det J| OP_test(phi_i) . D[cell][q] OP_trial(phi_j);  x, y distinct */
int eo_form_action(eo_form* form, int kind_test, int kind_trial, const double* D, const double* x, int64_t n_cells,
                   double* y, int accumulate);
/* One Newton residual evaluation of the von Mises demo on the device (demo_vm:500-513: external_callback +
 * assemble_vector): Mandel strain of u -> radial return (results identical to eo_tab_vm_fused, stored for
 * eo_form_action / eo_commit_history) -> b (+)= int sigma . epsilon(v) dx.  One kernel; u, b any-side. */
int eo_form_vm_step(eo_form* form, const eo_vm_params* prm, const double* u, const double* sigma_n, const double* p,
                    double* C_tang, double* sigma, double* dp, int64_t n_cells, double* b, int accumulate, int exact);
/* The same step with the consistent tangent kept in FACTORED form for the device-side consumers: T6[n_points][6] =
 * (v0, v1, v2, v3, cn, cd) with C_t = C_elas - cn v v^T - cd dev, v = n_elas f+ / f (demo_vm:317),
 * cn = 3 mu (3 mu / (3 mu + H) - beta) (:323), cd = 2 mu beta (:324): 48 instead of 128 bytes per point written here and
 * read by every tangent action (the reference's (n, 4, 4) array is what DOLFINx's assembler needs, not what a matrix-free
 * Krylov method needs).  Stress, dp, statistics and b are those of eo_form_vm_step.  T6: 16-byte aligned device memory. */
int eo_form_vm_step_factored(eo_form* form, const eo_vm_params* prm, const double* u, const double* sigma_n, const double* p,
                             double* T6, double* sigma, double* dp, int64_t n_cells, double* b, int accumulate, int exact);
/* y (+)= A x, A = assemble_matrix(inner(C_t epsilon(u_hat), epsilon(v)) dx) (demo_vm:390-398) with C_t rebuilt from T6 in
 * registers (three points per cell, P1/P2/P3 vector triangles).  Agrees with eo_form_action on the expanded tangent to
 * rounding. */
int eo_form_action_vm_factored(eo_form* form, const eo_vm_params* prm, const double* T6, const double* x, int64_t n_cells,
                               double* y, int accumulate);
/* T6 -> C_tang[n][4][4], by the statements of the un-factored kernels (exact != 0: the reference's statement sequence,
 * else the symmetric / FMA form of the default fused kernels): bit-identical to what eo_form_vm_step stores. */
int eo_vm_expand_tangent(eo_ctx* ctx, const eo_vm_params* prm, const double* T6, double* C_tang, int64_t n, int exact);
/* CSR pattern of the assembled matrix over scalar dofs (row = bs * node + comp): host arrays row_ptr[bs*n_dofs + 1],
 * col[nnz] (strictly increasing within a row - what `fem.create_sparsity_pattern` + finalize gives); validated and
 * uploaded once.  eo_form_matrix then adds every element matrix into vals[nnz] (device memory). */
int eo_form_set_pattern(eo_form* form, const int32_t* row_ptr, const int32_t* col, int64_t nnz);
/* eo_form_matrix is complete on return (it reads back a miss counter): element entries of the integrated cells that the
 * pattern does not hold are an error (EO_ERR_INVALID).  The CSR position of every element entry is cached on the device
 * at the first assembly after eo_form_set_pattern when it fits in a quarter of the free memory ((nb*bs)^2 int32 per
 * cell); later assemblies then neither search nor touch the pattern. */
int64_t eo_form_nnz(const eo_form* form);
int eo_form_matrix(eo_form* form, int kind_test, int kind_trial, const double* D, int64_t n_cells, double* vals,
                   int accumulate);

/* ---------------------------------------------------------------- general operand tabulation
 * replaces: `expr.eval(operand_mesh, entities)`, external_operator.py:365-402, for everything the affine-simplex
 *           fast path above does not cover: any element given by its tables (higher-degree simplices,
 *           quadrilaterals, hexahedra), non-affine geometry (the Jacobian is evaluated per point from the geometry
 *           element's derivative tables) and codimension-1 `entities` of shape (n, 2) = (cell, local facet)
 *           (test/test_codim_external_operator.py:75-109): one table set per local facet, tabulated by basix on
 *           the host at the facet's points mapped into the reference cell.
 * One thread per (entity, point); tables in shared memory when they fit. */
typedef struct eo_gtab_desc {
  int32_t gdim;            /* geometric = topological dimension, 2 or 3                                     */
  int32_t bs;              /* block size of the coefficient, 1..3                                           */
  int32_t nb;              /* scalar basis functions per cell, 1..125                                       */
  int32_t nq;              /* evaluation points per entity, 1..125                                          */
  int32_t ng;              /* geometry nodes per cell (3 triangle, 4 quadrilateral/tetrahedron, 8 hexahedron, 6 P2 triangle ...) */
  int32_t n_sets;          /* table sets: 1 = points inside the cell; number of local facets for codim-1    */
  int64_t n_cells, n_dofs, n_nodes;
  const int32_t* dofmap;   /* [n_cells][nb]                                                                 */
  const int32_t* x_dofmap; /* [n_cells][ng]                                                                 */
  const double* x;         /* [n_nodes][3]                                                                  */
  const double* phi;       /* [n_sets][nq][nb]        basix tabulate(1, X_set)[0]                           */
  const double* dphi;      /* [n_sets][gdim][nq][nb]  basix tabulate(1, X_set)[1:]                          */
  const double* dgeo;      /* [n_sets][gdim][nq][ng]  the geometry element's tabulate(1, X_set)[1:]         */
} eo_gtab_desc;

typedef struct eo_gtab eo_gtab;
int eo_gtab_create(eo_ctx* ctx, const eo_gtab_desc* desc, eo_gtab** out);
int eo_gtab_destroy(eo_gtab* tab);
int eo_gtab_ncomp(const eo_gtab* tab, int kind);
/* out[n_entities][nq][ncomp].  entity_width 0: entities == NULL, cells 0..n_entities-1 (:365-371);
 * 1: int32 cell indices; 2: int32 (cell, local facet) pairs.  u, entities and out are any-side pointers. */
int eo_gtab_tabulate(eo_gtab* tab, int kind, const double* u, const int32_t* entities, int entity_width,
                     int64_t n_entities, double* out);

/* ---------------------------------------------------------------- Mohr-Coulomb
 * replaces: `dsigma_ddeps_vec = jit(vmap(jacfwd(return_mapping, has_aux=True)))` and the body of
 *           `C_tang_impl`, doc/demo/demo_plasticity_mohr_coulomb.py:474-533, :555, :574-593
 *           (surface :282-374, residual/Jacobian :405-465).
 * Constants: demo_mc:110-116 (E, nu, c, phi, psi, theta_T, a) and :469 (tol, Nitermax). */
typedef struct eo_mc_params {
  double E, nu, c, phi, psi, theta_T, a, tol;
  int32_t Nitermax; /* <= 200 */
} eo_mc_params;

/* deps [n][4], sigma_n [n][4]  ->  C_tang [n][4][4] (the derivative of the Newton ITERATION w.r.t.
 * deps, i.e. what jacfwd-through-while_loop returns, not the implicit-function tangent), sigma [n][4],
 * and the aux tuple of demo_mc:533 per point: niter [n] (int32), yielding [n] = f(trial stress),
 * norm_res [n], dlambda [n] - each of the four aux arrays may be NULL.
 * Elastic points (yielding <= 0) take one Newton step with J = I: C_tang = C_elas exactly, niter = 1.
 * IEEE hazards are the reference's: J2 == 0 at a plastic point or ||res0|| == 0 give NaN / zero
 * iterations, clipped asin arguments give NaN tangents.  The histogram of niter, max yielding,
 * max norm_res (the per-call prints of demo_mc:584-591) and the counts of plastic / non-converged /
 * non-finite points are accumulated into the ctx statistics record.
 * Algorithmic HBM traffic: 64 B read + 160 B (+ 28 B aux) written per point; the kernel is bound by
 * the FP64 pipe (local Newton + tangent recursion), not by HBM.
 * deps, sigma_n, C_tang, sigma must be 32-byte aligned when they are device pointers. */
int eo_mc_eval(eo_ctx* ctx, const eo_mc_params* prm, const double* deps, const double* sigma_n, double* C_tang,
               double* sigma, int32_t* niter, double* yielding, double* norm_res, double* dlambda, int64_t n);
/* Same with an explicit execution scheme: 0 = two passes (the default of eo_mc_eval): yield test for every point
 * at full occupancy, elastic points finished there, plastic points listed; then persistent CTAs run the Newton
 * iterations of the listed points out of shared-memory state slots that are partitioned by lane ("lane classes":
 * conflict-free slot accesses, one bit mask per class instead of queues), a warp always executing one kind of stage.
 * 1 = one thread per point (divergent baseline; does not update the statistics record); 2 / 3 = one pass with the
 * ring-queue scheduler of round 1 (the yield test is a stage of the scheduler) without / with sub-partition affinity;
 * 4 = two passes with that ring-queue scheduler - all kept for A/B measurements.  All schemes give the same results
 * (2, 3, 4 bit for bit; 0 and 1 to rounding: the same per-point source inlined into different kernels). */
int eo_mc_eval_scheme(eo_ctx* ctx, const eo_mc_params* prm, const double* deps, const double* sigma_n,
                      double* C_tang, double* sigma, int32_t* niter, double* yielding, double* norm_res,
                      double* dlambda, int64_t n, int scheme);

/* Fused hot path of the Mohr-Coulomb demo for all cells of `tab` (a P1/P2/P3 vector field on triangles): the Mandel
 * strain of the coefficient vector u (any-side) is tabulated INSIDE pass 1 of the two-pass scheme and kept only for the
 * plastic points (next to the list entry that pass 2 works off): the strain array of the mesh is never written.
 * replaces: `evaluate_operands` + `C_tang_impl` of demo_plasticity_mohr_coulomb.py:679-688.  Point index = cell * nq + q;
 * sigma_n, C_tang, sigma and the optional aux arrays are device memory; asynchronous on the ctx stream.  Same results as
 * eo_tabulate -> eo_mc_eval up to the rounding of the strain (this translation unit contracts FMAs, tab.cu does not). */
int eo_mc_eval_tabulated(eo_ctx* ctx, const eo_mc_params* prm, eo_tab* tab, const double* u, const double* sigma_n,
                         double* C_tang, double* sigma, int32_t* niter, double* yielding, double* norm_res, double* dlambda);

/* ---------------------------------------------------------------- Isihara ICNN hyperelasticity
 * replaces: `vectorized_stress_and_tangent` / `dP_dF_impl`, doc/demo/demo_hyperelasticity.py:429-456
 *           (network :242-307 with the state dict Isihara_noise=high.pth, corrections :362-381).
 * The network constants are passed PREPROCESSED (what the Python host side computes once from the state dict:
 * softplus applied to the convex weights, :238; layer 1 collapsed onto the affine layer 0):
 *   a1 = A1 x + c1           A1 = softplus(layers.1.weights) @ layers.0.weight + skip_layers.1.weight   [64][3]
 *                            c1 = softplus(layers.1.weights) @ layers.0.bias   + skip_layers.1.bias     [64]
 *   a2 = W2 phi(a1) + S2 x + b2,  y = w3 . phi(a2) + s3 . x,  phi(a) = softplus(a)^2 / 12   (float32, :286-300) */
typedef struct eo_isihara_weights {
  float A1[64][4];  /* columns 0..2 = A1, column 3 = c1 */
  float S2[64][4];  /* columns 0..2 = skip_layers.2.weight, column 3 = skip_layers.2.bias */
  float W2[64][64]; /* softplus(layers.2.weights), [out][in] */
  float W2T[64][64];
  float w3[64];     /* softplus(layers.3.weights) */
  float s3[4];      /* softplus(skip_layers.3.weights), padded */
  double H[4];      /* H_flat = -dW_NN/dF at F = I (:362-367); see eo_isihara_set_correction */
} eo_isihara_weights;

typedef struct eo_isihara eo_isihara;
int eo_isihara_create(eo_ctx* ctx, const eo_isihara_weights* w, eo_isihara** out);
int eo_isihara_destroy(eo_isihara* m);
/* Replace the stress correction H_flat (e.g. by -P_NN(F = I) evaluated with this library and H = 0). */
int eo_isihara_set_correction(eo_isihara* m, const double H_flat[4]);
/* F [n][4] = [F11, F12, F21, F22] (:263-266)  ->  dP [n][4][4] (dP_i/dF_j), P [n][4]; float64 in and out,
 * network arithmetic in float32 like the reference (results are float32-accurate).  Any-side pointers. */
int eo_isihara_eval(eo_isihara* m, const double* F, double* dP, double* P, int64_t n);
/* The same for device arrays on the CALLER's stream (cudaStream_t passed as void*; NULL = the legacy default stream):
 * asynchronous, ordered only by that stream - what the torch custom op `eo::isihara_dP_dF` uses with torch's current
 * stream, so that the op composes with the surrounding torch work without any synchronisation. */
int eo_isihara_eval_on_stream(eo_isihara* m, const double* F, double* dP, double* P, int64_t n, void* stream);

/* ---------------------------------------------------------------- generic run-time compiled models
 * replaces: any user `external_function(derivatives)(*operands)` (external_operator.py:432) that is not one
 *           of the hard-wired kernels, together with the automatic differentiation the reference demos take
 *           from their array library (README.md:16-25; jax.jacfwd demo_mc:555; torch.func
 *           demo_hyperelasticity.py:429-456).
 * `source` is CUDA C++ text defining
 *     template <class T> __device__ void ENTRY(const T* x, const double* state, const double* prm, T* y, T* aux);
 * x = all operands concatenated (operand_size[0] + ... components), state = all per-point state fields
 * concatenated (read only, not differentiated), prm = n_params doubles, y = out_size results, aux = all
 * auxiliary outputs concatenated (e.g. the plastic multiplier increment of demo_vm:352).  T is double for the
 * value and eo::dual<...> (include/eo_dual.h) for derivatives.  One NVRTC compilation (sm_100a) per derivative
 * multi-index, cached in the handle - and on disk when EO_JIT_CACHE_DIR names a directory (CUBIN files keyed by
 * the translation unit, the embedded headers, the options and the NVRTC version).  `ctx` may be NULL for a
 * compile-only handle (works without a GPU). */
#define EO_JIT_MAX_ARGS 8
#define EO_JIT_MAX_PARAMS 32
typedef struct eo_jit_desc {
  const char* source;
  const char* entry;
  int32_t n_operands, operand_size[EO_JIT_MAX_ARGS]; /* components per point, 1..16 */
  int32_t n_state, state_size[EO_JIT_MAX_ARGS];
  int32_t out_size;                                  /* components of the operator's value */
  int32_t n_aux, aux_size[EO_JIT_MAX_ARGS];
  int32_t n_params;
  int32_t fmad;                                      /* 0: plain IEEE sequence (--fmad=false), 1: contract */
} eo_jit_desc;

typedef struct eo_jit eo_jit;
int eo_jit_create(eo_ctx* ctx, const eo_jit_desc* desc, eo_jit** out);
int eo_jit_destroy(eo_jit* m);
/* Compile (if not cached) the kernel for `derivatives` (one int per operand, sum <= 2; NULL = value). */
int eo_jit_compile(eo_jit* m, const int* derivatives, size_t* cubin_bytes);
/* The staged variant of the kernel for `derivatives`: one CTA per tile of `*tile_points` points, every per-point
 * array of the tile moved by 1-D TMA bulk copies (cp.async.bulk + mbarrier) through shared memory - coalesced for
 * any number of components per point.  eo_jit_eval uses it automatically when some array has an odd component count
 * (EO_JIT_STAGED=0/1 forces never/always); this entry point only compiles it (works without a GPU). */
int eo_jit_compile_staged(eo_jit* m, const int* derivatives, int* tile_points, size_t* cubin_bytes);
/* The several-points-per-thread variant used for scalar-sized models (<= 4 doubles read, <= 8 written per point):
 * each thread fetches 2 or 4 consecutive points of every array with one wide access (EO_JIT_PPT=0 disables it).
 * Compile-only entry point, like eo_jit_compile_staged. */
int eo_jit_compile_ppt(eo_jit* m, const int* derivatives, int* points_per_thread, size_t* cubin_bytes);
/* Version of the NVRTC that was found (major * 1000 + minor * 10, e.g. 12090), or EO_ERR_UNSUPPORTED.  The
 * toolkit's /usr/local/cuda/lib64/libnvrtc.so.12 is preferred (EO_NVRTC_LIB overrides); NVRTC older than 12.9
 * cannot assemble 256-bit global accesses, models compiled with it use 128-bit ones. */
int eo_jit_nvrtc_version(void);
/* Copy the compiled sm_100a CUBIN for `derivatives` (size from eo_jit_compile) - for inspection with
 * cuobjdump / caching by the caller. */
int eo_jit_cubin(eo_jit* m, const int* derivatives, void* buf, size_t buf_bytes);
/* NVRTC log of the last compilation / text of the last error on this handle. */
const char* eo_jit_log(const eo_jit* m);
const char* eo_jit_last_error(const eo_jit* m);
/* doubles per point of `out` for `derivatives`: out_size x size(operand a) [x size(operand b)], or < 0. */
int eo_jit_out_width(eo_jit* m, const int* derivatives);
/* Evaluate at n points.  operands[i] : [n][operand_size[i]], state[i] : [n][state_size[i]] (any-side).
 * out : [n][out_width]  - the value, or the derivative laid out [point][out][operand a][operand b] like the
 *       reference's `space_shape + operand_shape` convention (external_operator.py:117-121);
 * value (optional, derivative orders >= 1) : [n][out_size];  aux[i] (each optional) : [n][aux_size[i]]. */
int eo_jit_eval(eo_jit* m, const int* derivatives, const double* params, const double* const* operands,
                const double* const* state, double* out, double* value, double* const* aux, int64_t n);
/* Fused hot path for a run-time compiled model (the general form of eo_tab_vm_fused; SURVEY.md 8b `eo_fused_step`):
 * operand i is tabulated INSIDE the kernel from coefficient vector coefficients[i] (any-side, bs * n_dofs doubles)
 * with operand kind kinds[i] on tabs[i] (all tabs: same cells and evaluation points), fed to the model from
 * registers and never stored.  n = n_cells * nq points, point index = cell * nq + q.  state / out / value / aux as in
 * eo_jit_eval.  One NVRTC compilation per (derivatives, element signature), cached. */
int eo_jit_eval_tabulated(eo_jit* m, const int* derivatives, const double* params, eo_tab* const* tabs, const int* kinds,
                          const double* const* coefficients, const double* const* state, double* out, double* value,
                          double* const* aux);
/* Compile such a kernel without a GPU: element signature (gdim, block size, basis functions, kind) per operand. */
int eo_jit_compile_tabulated(eo_jit* m, const int* derivatives, int nq, const int* gdim, const int* bs, const int* nb,
                             const int* kinds, size_t* cubin_bytes);

#ifdef __cplusplus
}
#endif
#endif /* EO_B200_H */
