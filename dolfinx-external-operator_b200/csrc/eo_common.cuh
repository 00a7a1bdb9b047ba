// eo_common.cuh - internals shared by the translation units of libeo_b200.so.
#pragma once

#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <vector>

#include "eo_b200.h"

#define EO_NSLOT 3  // pipeline depth of the host-side (staged) path

struct eo_ctx {
  int device = 0;
  int sm_count = 0;
  cudaStream_t s_cmp = nullptr;  // compute stream (== eo_stream)
  cudaStream_t s_in = nullptr;   // H2D copies
  cudaStream_t s_out = nullptr;  // D2H copies
  cudaEvent_t ev_in[EO_NSLOT] = {};
  cudaEvent_t ev_cmp[EO_NSLOT] = {};
  cudaEvent_t ev_out[EO_NSLOT] = {};
  int64_t chunk = int64_t(1) << 20;
  // staging arena for the host-side path (grown on demand, never shrunk)
  char* arena = nullptr;
  size_t arena_bytes = 0;
  // L2 flush scratch
  char* flush = nullptr;
  size_t flush_bytes = 0;
  // general device scratch (plastic-point list of the two-pass Mohr-Coulomb scheme), grown on demand
  char* scratch = nullptr;
  size_t scratch_bytes = 0;
  eo_stats* stats = nullptr;  // device: the LOCAL record the kernels accumulate into
  // the one collective (eo_allreduce_stats): snapshot of the local record, the gathered records of all ranks and
  // their combination; the collective runs on its own stream so that it overlaps the next evaluation
  cudaStream_t s_coll = nullptr;
  cudaEvent_t ev_coll_ready = nullptr, ev_coll_done = nullptr;
  eo_stats* stats_send = nullptr;    // device [1]
  eo_stats* stats_recv = nullptr;    // device [stats_recv_world]
  int stats_recv_world = 0;
  eo_stats* stats_global = nullptr;  // device [1]
  unsigned int* work_ctr = nullptr;  // device: tile counter of the persistent kernels
  int64_t launches = 0;
  char err[512] = {0};
};

extern char g_eo_create_error[512];

int eo_fail(eo_ctx* ctx, int code, const char* fmt, ...);

#define EO_CUDA(ctx, call)                                                                         \
  do {                                                                                             \
    cudaError_t e__ = (call);                                                                      \
    if (e__ != cudaSuccess)                                                                        \
      return eo_fail(ctx, e__ == cudaErrorMemoryAllocation ? EO_ERR_NOMEM : EO_ERR_CUDA, "%s: %s", \
                     #call, cudaGetErrorString(e__));                                              \
  } while (0)

#define EO_REQUIRE(ctx, cond, msg) \
  do {                             \
    if (!(cond)) return eo_fail(ctx, EO_ERR_INVALID, "%s", msg); \
  } while (0)

// device scratch of at least `bytes` (contents undefined; valid until the next eo_scratch call on this ctx)
int eo_scratch(eo_ctx* ctx, size_t bytes, void** out);

// true when `p` is memory the GPU kernels can dereference in place
bool eo_is_device_ptr(const void* p);

// what jit.cu needs from an eo_tab (defined in tab.cu)
struct tab_tables;
struct eo_tab_view {
  eo_ctx* ctx;
  const tab_tables* T_host;
  const tab_tables* T_dev;
  const int32_t* dofmap;
  const int32_t* x_dofmap;
  const double* x;
  const double* u;  // device pointer of the (possibly staged) coefficient vector
  int64_t n_cells, n_dofs;
};
// `slot`: staging buffer of the handle used when `u` is a host vector (one per operand of a fused launch)
int eo_tab_view_get(eo_tab* t, const double* u, eo_tab_view* v, int slot);

// ------------------------------------------------------------------------------------
// Any-side argument of a per-quadrature-point ("streamed") operation.
// ------------------------------------------------------------------------------------
struct eo_arg {
  const void* ptr;  // host or device, may be nullptr (= absent)
  size_t bpq;       // bytes per quadrature point
  bool is_out;
  bool on_host = false;  // filled by eo_run_streamed
};

// `launch(ptrs, n_chunk, qp_offset)` must enqueue the kernel(s) on ctx->s_cmp; ptrs[i] is the device
// address of argument i for this chunk (nullptr when the argument is absent).
// All-device arguments: one launch over [0, n), asynchronous.  Otherwise chunks of ctx->chunk points
// flow through EO_NSLOT staging slots: H2D on s_in, kernel on s_cmp, D2H on s_out, chained by events.
template <class Launch>
int eo_run_streamed(eo_ctx* ctx, eo_arg* args, int nargs, int64_t n, Launch launch) {
  if (n == 0) return EO_OK;
  bool any_host = false;
  size_t host_bpq = 0;
  for (int i = 0; i < nargs; ++i) {
    args[i].on_host = args[i].ptr != nullptr && !eo_is_device_ptr(args[i].ptr);
    if (args[i].on_host) {
      any_host = true;
      host_bpq += (args[i].bpq + 31) / 32 * 32;
    }
  }
  std::vector<void*> ptrs(nargs);
  if (!any_host) {
    for (int i = 0; i < nargs; ++i) ptrs[i] = const_cast<void*>(args[i].ptr);
    int rc = launch(ptrs.data(), n, int64_t(0));
    if (rc != EO_OK) return rc;
    EO_CUDA(ctx, cudaGetLastError());
    return EO_OK;
  }
  const int64_t chunk = n < ctx->chunk ? n : ctx->chunk;
  // every staged array starts on a 256 B boundary inside its slot
  size_t slot_bytes = 0;
  for (int i = 0; i < nargs; ++i)
    if (args[i].on_host) slot_bytes += (size_t(chunk) * args[i].bpq + 255) / 256 * 256;
  const size_t need = slot_bytes * EO_NSLOT;
  if (need > ctx->arena_bytes) {
    EO_CUDA(ctx, cudaStreamSynchronize(ctx->s_cmp));
    if (ctx->arena) cudaFree(ctx->arena);
    ctx->arena = nullptr;
    ctx->arena_bytes = 0;
    EO_CUDA(ctx, cudaMalloc(&ctx->arena, need));
    ctx->arena_bytes = need;
  }
  // make the copy streams see everything queued on the compute stream so far
  EO_CUDA(ctx, cudaEventRecord(ctx->ev_cmp[0], ctx->s_cmp));
  EO_CUDA(ctx, cudaStreamWaitEvent(ctx->s_in, ctx->ev_cmp[0], 0));
  EO_CUDA(ctx, cudaStreamWaitEvent(ctx->s_out, ctx->ev_cmp[0], 0));
  int64_t c = 0;
  for (int64_t off = 0; off < n; off += chunk, ++c) {
    const int slot = int(c % EO_NSLOT);
    const int64_t m = (n - off) < chunk ? (n - off) : chunk;
    char* base = ctx->arena + size_t(slot) * slot_bytes;
    if (c >= EO_NSLOT) {
      // slot reuse: its previous kernel must have read its inputs, its previous D2H must have drained
      EO_CUDA(ctx, cudaStreamWaitEvent(ctx->s_in, ctx->ev_cmp[slot], 0));
      EO_CUDA(ctx, cudaStreamWaitEvent(ctx->s_cmp, ctx->ev_out[slot], 0));
    }
    size_t cur = 0;
    for (int i = 0; i < nargs; ++i) {
      if (!args[i].ptr) {
        ptrs[i] = nullptr;
      } else if (args[i].on_host) {
        ptrs[i] = base + cur;
        cur += (size_t(chunk) * args[i].bpq + 255) / 256 * 256;
        if (!args[i].is_out)
          EO_CUDA(ctx, cudaMemcpyAsync(ptrs[i], (const char*)args[i].ptr + size_t(off) * args[i].bpq,
                                       size_t(m) * args[i].bpq, cudaMemcpyHostToDevice, ctx->s_in));
      } else {
        ptrs[i] = (char*)args[i].ptr + size_t(off) * args[i].bpq;
      }
    }
    EO_CUDA(ctx, cudaEventRecord(ctx->ev_in[slot], ctx->s_in));
    EO_CUDA(ctx, cudaStreamWaitEvent(ctx->s_cmp, ctx->ev_in[slot], 0));
    int rc = launch(ptrs.data(), m, off);
    if (rc != EO_OK) return rc;
    EO_CUDA(ctx, cudaGetLastError());
    EO_CUDA(ctx, cudaEventRecord(ctx->ev_cmp[slot], ctx->s_cmp));
    EO_CUDA(ctx, cudaStreamWaitEvent(ctx->s_out, ctx->ev_cmp[slot], 0));
    for (int i = 0; i < nargs; ++i)
      if (args[i].ptr && args[i].on_host && args[i].is_out)
        EO_CUDA(ctx, cudaMemcpyAsync((char*)args[i].ptr + size_t(off) * args[i].bpq, ptrs[i],
                                     size_t(m) * args[i].bpq, cudaMemcpyDeviceToHost, ctx->s_out));
    EO_CUDA(ctx, cudaEventRecord(ctx->ev_out[slot], ctx->s_out));
  }
  EO_CUDA(ctx, cudaStreamSynchronize(ctx->s_in));
  EO_CUDA(ctx, cudaStreamSynchronize(ctx->s_cmp));
  EO_CUDA(ctx, cudaStreamSynchronize(ctx->s_out));
  return EO_OK;
}

// ------------------------------------------------------------------------------------
// device helpers: 256-bit streaming loads/stores (sm_100: LDG.E.256 / STG.E.256)
// ------------------------------------------------------------------------------------
#ifdef __CUDACC__
struct __align__(32) eo_d4 {
  double x, y, z, w;
};

// read-once data: bypass L1 allocation
__device__ __forceinline__ eo_d4 eo_ld256(const double* p) {
  eo_d4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f64 {%0,%1,%2,%3}, [%4];"
               : "=d"(r.x), "=d"(r.y), "=d"(r.z), "=d"(r.w)
               : "l"(p));
  return r;
}
__device__ __forceinline__ void eo_st256(double* p, double x, double y, double z, double w) {
  asm volatile("st.global.L1::no_allocate.v4.f64 [%0], {%1,%2,%3,%4};" ::"l"(p), "d"(x), "d"(y), "d"(z), "d"(w)
               : "memory");
}
__device__ __forceinline__ double2 eo_ld128(const double* p) {
  double2 r;
  asm volatile("ld.global.nc.L1::no_allocate.v2.f64 {%0,%1}, [%2];" : "=d"(r.x), "=d"(r.y) : "l"(p));
  return r;
}
__device__ __forceinline__ void eo_st128(double* p, double x, double y) {
  asm volatile("st.global.L1::no_allocate.v2.f64 [%0], {%1,%2};" ::"l"(p), "d"(x), "d"(y) : "memory");
}
__device__ __forceinline__ double eo_ld64(const double* p) {
  double r;
  asm volatile("ld.global.nc.L1::no_allocate.f64 %0, [%1];" : "=d"(r) : "l"(p));
  return r;
}
__device__ __forceinline__ void eo_st64(double* p, double x) {
  asm volatile("st.global.L1::no_allocate.f64 [%0], %1;" ::"l"(p), "d"(x) : "memory");
}

// 4x4 tangents of the four consecutive points of a lane quad (lanes 4k..4k+3), stored as WHOLE 128-byte lines:
// a warp-wide 256-bit store of per-thread 128-byte records touches 32 different lines per instruction (32 L1
// wavefronts; four instructions per tangent), which is what bounds the fused kernels, not HBM.  The quad swaps rows
// with a two-stage butterfly (2 x 16 SHFL.32) so that lane r holds row r of all four points; instruction j then
// writes point j's full line from the four lanes of the quad: 8 lines per instruction.  The values are untouched.
// Must be called by all 32 lanes; `C_quad` = address of the tangent of the quad's FIRST point, `n_valid` = how many
// of the quad's points exist (0..4).
__device__ __forceinline__ void eo_st_tangent_quad(double* C_quad, int n_valid, double C[16]) {
  unsigned lane;
  asm("mov.u32 %0, %%laneid;" : "=r"(lane));
  const bool b2 = lane & 2u, b1 = lane & 1u;
#pragma unroll
  for (int s = 0; s < 2; ++s)  // rows (s, s + 2) <-> lanes (q, q ^ 2)
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const double got = __shfl_xor_sync(0xffffffffu, b2 ? C[4 * s + k] : C[4 * (s + 2) + k], 2);
      if (b2) C[4 * s + k] = got; else C[4 * (s + 2) + k] = got;
    }
#pragma unroll
  for (int s = 0; s < 4; s += 2)  // rows (s, s + 1) <-> lanes (q, q ^ 1)
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const double got = __shfl_xor_sync(0xffffffffu, b1 ? C[4 * s + k] : C[4 * (s + 1) + k], 1);
      if (b1) C[4 * s + k] = got; else C[4 * (s + 1) + k] = got;
    }
  double* row = C_quad + 4 * (lane & 3u);
#pragma unroll
  for (int j = 0; j < 4; ++j)
    if (j < n_valid) eo_st256(row + 16 * j, C[4 * j + 0], C[4 * j + 1], C[4 * j + 2], C[4 * j + 3]);
}

// block-wide sum of per-thread counts (grid-stride kernels: a thread may have counted several points)
__device__ __forceinline__ void eo_block_sum_add(unsigned long long* dst, int count) {
  __shared__ int s_sum;
  if (threadIdx.x == 0) s_sum = 0;
  __syncthreads();
  const int w = __reduce_add_sync(0xffffffffu, count);
  if ((threadIdx.x & 31) == 0 && w) atomicAdd(&s_sum, w);
  __syncthreads();
  if (threadIdx.x == 0 && s_sum) atomicAdd(dst, (unsigned long long)s_sum);
}

// block-wide sum of a per-thread FLAG (0/1), one atomicAdd per CTA
__device__ __forceinline__ void eo_block_count_add(unsigned long long* dst, int flag) {
  __shared__ int s_cnt;
  if (threadIdx.x == 0) s_cnt = 0;
  __syncthreads();
  unsigned m = __ballot_sync(0xffffffffu, flag);
  if ((threadIdx.x & 31) == 0 && m) atomicAdd(&s_cnt, __popc(m));
  __syncthreads();
  if (threadIdx.x == 0 && s_cnt) atomicAdd(dst, (unsigned long long)s_cnt);
}
#endif

static inline bool eo_aligned(const void* p, size_t a) { return (reinterpret_cast<uintptr_t>(p) % a) == 0; }
