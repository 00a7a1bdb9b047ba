// eo_runtime.cu - context, memory, events and statistics of libeo_b200.so.
#include "eo_common.cuh"

#include <dlfcn.h>

#include <cmath>
#include <cstddef>

char g_eo_create_error[512] = {0};

int eo_fail(eo_ctx* ctx, int code, const char* fmt, ...) {
  char* dst = ctx ? ctx->err : g_eo_create_error;
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(dst, 512, fmt, ap);
  va_end(ap);
  return code;
}

int eo_scratch(eo_ctx* ctx, size_t bytes, void** out) {
  if (bytes > ctx->scratch_bytes) {
    EO_CUDA(ctx, cudaStreamSynchronize(ctx->s_cmp));
    if (ctx->scratch) cudaFree(ctx->scratch);
    ctx->scratch = nullptr, ctx->scratch_bytes = 0;
    EO_CUDA(ctx, cudaMalloc(&ctx->scratch, bytes));
    ctx->scratch_bytes = bytes;
  }
  *out = ctx->scratch;
  return EO_OK;
}

bool eo_is_device_ptr(const void* p) {
  cudaPointerAttributes a;
  cudaError_t e = cudaPointerGetAttributes(&a, p);
  if (e != cudaSuccess) {
    cudaGetLastError();  // clear
    return false;
  }
  return a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged;
}

__global__ void eo_flush_kernel(float4* buf, size_t n4) {
  size_t i = blockIdx.x * size_t(blockDim.x) + threadIdx.x;
  const size_t stride = size_t(gridDim.x) * blockDim.x;
  for (; i < n4; i += stride) buf[i] = make_float4(float(i), 0.f, 0.f, 0.f);
}

// FP64 roofline denominator: 8 independent DFMA chains per thread, no memory traffic
__global__ void __launch_bounds__(256) eo_fp64_peak_kernel(double* out, int iters, double b, double c) {
  double a0 = threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
  for (int i = 0; i < iters; ++i) {
    a0 = fma(a0, b, c), a1 = fma(a1, b, c), a2 = fma(a2, b, c), a3 = fma(a3, b, c);
    a4 = fma(a4, b, c), a5 = fma(a5, b, c), a6 = fma(a6, b, c), a7 = fma(a7, b, c);
  }
  out[blockIdx.x * size_t(blockDim.x) + threadIdx.x] = ((a0 + a1) + (a2 + a3)) + ((a4 + a5) + (a6 + a7));
}

// FP32 roofline denominator (Isihara network): 8 independent FFMA chains per thread
__global__ void __launch_bounds__(256) eo_fp32_peak_kernel(float* out, int iters, float b, float c) {
  float a0 = threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
  for (int i = 0; i < iters; ++i) {
    a0 = fmaf(a0, b, c), a1 = fmaf(a1, b, c), a2 = fmaf(a2, b, c), a3 = fmaf(a3, b, c);
    a4 = fmaf(a4, b, c), a5 = fmaf(a5, b, c), a6 = fmaf(a6, b, c), a7 = fmaf(a7, b, c);
  }
  out[blockIdx.x * size_t(blockDim.x) + threadIdx.x] = ((a0 + a1) + (a2 + a3)) + ((a4 + a5) + (a6 + a7));
}

// variant 1: every operand in a register that changes per chain (no constant-bank / immediate operand form)
__global__ void __launch_bounds__(256) eo_fp32_peak3_kernel(float* out, int iters, float b0, float c0) {
  float a[8], b[8], c[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) a[k] = threadIdx.x + k, b[k] = b0 + 1e-7f * (threadIdx.x + k), c[k] = c0 * (k + 1);
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int k = 0; k < 8; ++k) a[k] = fmaf(a[k], b[k], c[k]);
  }
  float s = 0.f;
#pragma unroll
  for (int k = 0; k < 8; ++k) s += a[k];
  out[blockIdx.x * size_t(blockDim.x) + threadIdx.x] = s;
}

// variant 3: one uniform operand (kernel parameter -> constant bank / uniform register), two per-chain registers:
// acc = fma(z, w_uniform, acc) - the shape of a matrix-vector product whose weights are warp uniform
__global__ void __launch_bounds__(256) eo_fp32_peak_u1_kernel(float* out, int iters, float b0, float c0) {
  float a[8], z[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) a[k] = threadIdx.x + k, z[k] = c0 * (k + 1) + 1e-7f * threadIdx.x;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int k = 0; k < 8; ++k) a[k] = fmaf(z[k], b0, a[k]);
  }
  float s = 0.f;
#pragma unroll
  for (int k = 0; k < 8; ++k) s += a[k];
  out[blockIdx.x * size_t(blockDim.x) + threadIdx.x] = s;
}

// variant 2: packed FFMA2 (fma.rn.f32x2), 8 independent chains of register pairs = 16 FMAs per trip
__global__ void __launch_bounds__(256) eo_fp32_peak2x_kernel(float* out, int iters, float b0, float c0) {
  float2 a[8], b[8], c[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    a[k] = make_float2(threadIdx.x + k, threadIdx.x + k + 0.5f);
    b[k] = make_float2(b0 + 1e-7f * (threadIdx.x + k), b0 - 1e-7f * (threadIdx.x + k));
    c[k] = make_float2(c0 * (k + 1), c0 * (k + 2));
  }
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int k = 0; k < 8; ++k) a[k] = __ffma2_rn(a[k], b[k], c[k]);
  }
  float s = 0.f;
#pragma unroll
  for (int k = 0; k < 8; ++k) s += a[k].x + a[k].y;
  out[blockIdx.x * size_t(blockDim.x) + threadIdx.x] = s;
}

// out[k] = values[src[k]]: the non-contiguous coefficient assignment turned inside out (see eo_assign_gather)
__global__ void __launch_bounds__(256) eo_gather_kernel(const double* __restrict__ values, const int64_t* __restrict__ src,
                                                        double* __restrict__ out, int64_t n_out) {
  const int64_t k = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
  if (k < n_out) out[k] = __ldg(values + __ldg(src + k));
}

// global record = SUM over ranks of the int64 block, MAX over ranks of the f64 block of the gathered records
__global__ void __launch_bounds__(256) eo_stats_combine_kernel(const eo_stats* __restrict__ recv, int world,
                                                               eo_stats* __restrict__ out) {
  constexpr int NS = 4 + EO_NITER_BINS, NM = 4;
  const int j = threadIdx.x;
  if (j < NS) {
    long long acc = 0;
    for (int r = 0; r < world; ++r) acc += reinterpret_cast<const long long*>(recv + r)[j];
    reinterpret_cast<long long*>(out)[j] = acc;
  } else if (j < NS + NM) {
    const int k = j - NS;
    double m = (&recv[0].niter_max)[k];
    for (int r = 1; r < world; ++r) {
      const double v = (&recv[r].niter_max)[k];
      m = (v > m || m != m) ? v : m;  // NaN-free maximum unless every rank holds NaN
    }
    (&out->niter_max)[k] = m;
  }
}

extern "C" {

int eo_version(void) { return EO_B200_VERSION; }

int eo_device_count(void) {
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess) {
    cudaGetLastError();
    eo_fail(nullptr, EO_ERR_NO_DEVICE, "cudaGetDeviceCount: %s", cudaGetErrorString(e));
    return EO_ERR_NO_DEVICE;
  }
  return n;
}

int eo_device_pci_bus_id(int device, char* buf, int len) {
  if (!buf || len < 16) return EO_ERR_INVALID;
  cudaError_t e = cudaDeviceGetPCIBusId(buf, len, device);
  if (e != cudaSuccess) {
    cudaGetLastError();
    return eo_fail(nullptr, EO_ERR_CUDA, "cudaDeviceGetPCIBusId: %s", cudaGetErrorString(e));
  }
  return EO_OK;
}

int eo_create(int device, eo_ctx** out) {
  if (!out) return eo_fail(nullptr, EO_ERR_INVALID, "eo_create: out is NULL");
  *out = nullptr;
  int n = eo_device_count();
  if (n <= 0) return eo_fail(nullptr, EO_ERR_NO_DEVICE, "eo_create: no CUDA device visible (this library has no CPU path)");
  if (device < 0 || device >= n) return eo_fail(nullptr, EO_ERR_INVALID, "eo_create: device %d out of range [0,%d)", device, n);
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) != cudaSuccess)
    return eo_fail(nullptr, EO_ERR_CUDA, "eo_create: cudaGetDeviceProperties failed");
  if (prop.major != 10)
    return eo_fail(nullptr, EO_ERR_NO_DEVICE, "eo_create: device %d is sm_%d%d; this library is built for sm_100a only",
                   device, prop.major, prop.minor);
  eo_ctx* ctx = new eo_ctx();
  ctx->device = device;
  ctx->sm_count = prop.multiProcessorCount;
#define EO_CREATE_CUDA(call)                                                                      \
  do {                                                                                            \
    cudaError_t e__ = (call);                                                                     \
    if (e__ != cudaSuccess) {                                                                     \
      eo_fail(nullptr, EO_ERR_CUDA, "eo_create: %s: %s", #call, cudaGetErrorString(e__));         \
      delete ctx;                                                                                 \
      return EO_ERR_CUDA;                                                                         \
    }                                                                                             \
  } while (0)
  EO_CREATE_CUDA(cudaSetDevice(device));
  EO_CREATE_CUDA(cudaStreamCreateWithFlags(&ctx->s_cmp, cudaStreamNonBlocking));
  EO_CREATE_CUDA(cudaStreamCreateWithFlags(&ctx->s_in, cudaStreamNonBlocking));
  EO_CREATE_CUDA(cudaStreamCreateWithFlags(&ctx->s_out, cudaStreamNonBlocking));
  for (int i = 0; i < EO_NSLOT; ++i) {
    EO_CREATE_CUDA(cudaEventCreateWithFlags(&ctx->ev_in[i], cudaEventDisableTiming));
    EO_CREATE_CUDA(cudaEventCreateWithFlags(&ctx->ev_cmp[i], cudaEventDisableTiming));
    EO_CREATE_CUDA(cudaEventCreateWithFlags(&ctx->ev_out[i], cudaEventDisableTiming));
  }
  EO_CREATE_CUDA(cudaStreamCreateWithFlags(&ctx->s_coll, cudaStreamNonBlocking));
  EO_CREATE_CUDA(cudaEventCreateWithFlags(&ctx->ev_coll_ready, cudaEventDisableTiming));
  EO_CREATE_CUDA(cudaEventCreateWithFlags(&ctx->ev_coll_done, cudaEventDisableTiming));
  EO_CREATE_CUDA(cudaMalloc(&ctx->stats, sizeof(eo_stats)));
  EO_CREATE_CUDA(cudaMemset(ctx->stats, 0, sizeof(eo_stats)));
  EO_CREATE_CUDA(cudaMalloc(&ctx->stats_send, sizeof(eo_stats)));
  EO_CREATE_CUDA(cudaMalloc(&ctx->stats_global, sizeof(eo_stats)));
  EO_CREATE_CUDA(cudaMemset(ctx->stats_global, 0, sizeof(eo_stats)));
  EO_CREATE_CUDA(cudaMalloc(&ctx->work_ctr, 256));
  EO_CREATE_CUDA(cudaMemset(ctx->work_ctr, 0, 256));
#undef EO_CREATE_CUDA
  eo_stats_reset(ctx);
  *out = ctx;
  return EO_OK;
}

int eo_destroy(eo_ctx* ctx) {
  if (!ctx) return EO_OK;
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->s_in);
  cudaStreamSynchronize(ctx->s_cmp);
  cudaStreamSynchronize(ctx->s_out);
  if (ctx->s_coll) cudaStreamSynchronize(ctx->s_coll);
  if (ctx->ev_coll_ready) cudaEventDestroy(ctx->ev_coll_ready);
  if (ctx->ev_coll_done) cudaEventDestroy(ctx->ev_coll_done);
  if (ctx->stats_send) cudaFree(ctx->stats_send);
  if (ctx->stats_recv) cudaFree(ctx->stats_recv);
  if (ctx->stats_global) cudaFree(ctx->stats_global);
  if (ctx->s_coll) cudaStreamDestroy(ctx->s_coll);
  for (int i = 0; i < EO_NSLOT; ++i) {
    if (ctx->ev_in[i]) cudaEventDestroy(ctx->ev_in[i]);
    if (ctx->ev_cmp[i]) cudaEventDestroy(ctx->ev_cmp[i]);
    if (ctx->ev_out[i]) cudaEventDestroy(ctx->ev_out[i]);
  }
  if (ctx->arena) cudaFree(ctx->arena);
  if (ctx->flush) cudaFree(ctx->flush);
  if (ctx->scratch) cudaFree(ctx->scratch);
  if (ctx->stats) cudaFree(ctx->stats);
  if (ctx->work_ctr) cudaFree(ctx->work_ctr);
  if (ctx->s_in) cudaStreamDestroy(ctx->s_in);
  if (ctx->s_cmp) cudaStreamDestroy(ctx->s_cmp);
  if (ctx->s_out) cudaStreamDestroy(ctx->s_out);
  delete ctx;
  return EO_OK;
}

const char* eo_last_error(const eo_ctx* ctx) { return ctx ? ctx->err : g_eo_create_error; }

int eo_sync(eo_ctx* ctx) {
  EO_REQUIRE(ctx, ctx != nullptr, "eo_sync: ctx is NULL");
  EO_CUDA(ctx, cudaStreamSynchronize(ctx->s_in));
  EO_CUDA(ctx, cudaStreamSynchronize(ctx->s_cmp));
  EO_CUDA(ctx, cudaStreamSynchronize(ctx->s_out));
  EO_CUDA(ctx, cudaStreamSynchronize(ctx->s_coll));
  return EO_OK;
}

void* eo_stream(eo_ctx* ctx) { return ctx ? (void*)ctx->s_cmp : nullptr; }

int eo_set_chunk(eo_ctx* ctx, int64_t n) {
  EO_REQUIRE(ctx, ctx != nullptr, "eo_set_chunk: ctx is NULL");
  EO_REQUIRE(ctx, n >= 32, "eo_set_chunk: chunk must be >= 32 quadrature points");
  ctx->chunk = (n + 31) / 32 * 32;  // keeps every chunk start 32 B aligned for every f64 field
  return EO_OK;
}

int64_t eo_launch_count(const eo_ctx* ctx) { return ctx ? ctx->launches : 0; }

// ------------------------------------------------------------------ memory
int eo_dev_alloc(eo_ctx* ctx, size_t bytes, void** dptr) {
  EO_REQUIRE(ctx, ctx && dptr, "eo_dev_alloc: NULL argument");
  *dptr = nullptr;
  if (bytes == 0) return EO_OK;
  EO_CUDA(ctx, cudaSetDevice(ctx->device));
  EO_CUDA(ctx, cudaMalloc(dptr, bytes));
  return EO_OK;
}

int eo_dev_free(eo_ctx* ctx, void* dptr) {
  EO_REQUIRE(ctx, ctx != nullptr, "eo_dev_free: ctx is NULL");
  if (!dptr) return EO_OK;
  EO_CUDA(ctx, cudaStreamSynchronize(ctx->s_cmp));
  EO_CUDA(ctx, cudaFree(dptr));
  return EO_OK;
}

int eo_dev_memset(eo_ctx* ctx, void* dptr, int value, size_t bytes) {
  EO_REQUIRE(ctx, ctx != nullptr, "eo_dev_memset: ctx is NULL");
  if (bytes == 0) return EO_OK;
  EO_REQUIRE(ctx, dptr != nullptr, "eo_dev_memset: dptr is NULL");
  EO_CUDA(ctx, cudaMemsetAsync(dptr, value, bytes, ctx->s_cmp));
  return EO_OK;
}

int eo_copy(eo_ctx* ctx, void* dst, const void* src, size_t bytes) {
  EO_REQUIRE(ctx, ctx != nullptr, "eo_copy: ctx is NULL");
  if (bytes == 0) return EO_OK;
  EO_REQUIRE(ctx, dst && src, "eo_copy: NULL pointer");
  const bool dd = eo_is_device_ptr(dst), sd = eo_is_device_ptr(src);
  EO_CUDA(ctx, cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDefault, ctx->s_cmp));
  if (!(dd && sd)) EO_CUDA(ctx, cudaStreamSynchronize(ctx->s_cmp));
  return EO_OK;
}

int eo_host_alloc(eo_ctx* ctx, size_t bytes, void** hptr) {
  EO_REQUIRE(ctx, ctx && hptr, "eo_host_alloc: NULL argument");
  *hptr = nullptr;
  if (bytes == 0) return EO_OK;
  EO_CUDA(ctx, cudaHostAlloc(hptr, bytes, cudaHostAllocDefault));
  return EO_OK;
}

int eo_host_free(eo_ctx* ctx, void* hptr) {
  EO_REQUIRE(ctx, ctx != nullptr, "eo_host_free: ctx is NULL");
  if (!hptr) return EO_OK;
  EO_CUDA(ctx, cudaFreeHost(hptr));
  return EO_OK;
}

int eo_host_register(eo_ctx* ctx, void* hptr, size_t bytes) {
  EO_REQUIRE(ctx, ctx && hptr, "eo_host_register: NULL argument");
  if (bytes == 0) return EO_OK;
  cudaError_t e = cudaHostRegister(hptr, bytes, cudaHostRegisterDefault);
  if (e == cudaErrorHostMemoryAlreadyRegistered) {
    cudaGetLastError();
    return EO_OK;
  }
  EO_CUDA(ctx, e);
  return EO_OK;
}

int eo_host_unregister(eo_ctx* ctx, void* hptr) {
  EO_REQUIRE(ctx, ctx && hptr, "eo_host_unregister: NULL argument");
  cudaError_t e = cudaHostUnregister(hptr);
  if (e == cudaErrorHostMemoryNotRegistered) {
    cudaGetLastError();
    return EO_OK;
  }
  EO_CUDA(ctx, e);
  return EO_OK;
}

// ------------------------------------------------------------------ timing
int eo_event_create(eo_ctx* ctx, void** ev) {
  EO_REQUIRE(ctx, ctx && ev, "eo_event_create: NULL argument");
  cudaEvent_t e;
  EO_CUDA(ctx, cudaEventCreate(&e));
  *ev = (void*)e;
  return EO_OK;
}
int eo_event_destroy(eo_ctx* ctx, void* ev) {
  EO_REQUIRE(ctx, ctx != nullptr, "eo_event_destroy: ctx is NULL");
  if (ev) EO_CUDA(ctx, cudaEventDestroy((cudaEvent_t)ev));
  return EO_OK;
}
int eo_event_record(eo_ctx* ctx, void* ev) {
  EO_REQUIRE(ctx, ctx && ev, "eo_event_record: NULL argument");
  EO_CUDA(ctx, cudaEventRecord((cudaEvent_t)ev, ctx->s_cmp));
  return EO_OK;
}
int eo_event_elapsed_ms(eo_ctx* ctx, void* a, void* b, float* ms) {
  EO_REQUIRE(ctx, ctx && a && b && ms, "eo_event_elapsed_ms: NULL argument");
  EO_CUDA(ctx, cudaEventSynchronize((cudaEvent_t)b));
  EO_CUDA(ctx, cudaEventElapsedTime(ms, (cudaEvent_t)a, (cudaEvent_t)b));
  return EO_OK;
}

int eo_flush_l2(eo_ctx* ctx, size_t bytes) {
  EO_REQUIRE(ctx, ctx != nullptr, "eo_flush_l2: ctx is NULL");
  if (bytes == 0) return EO_OK;
  bytes = (bytes + 15) / 16 * 16;
  if (bytes > ctx->flush_bytes) {
    EO_CUDA(ctx, cudaStreamSynchronize(ctx->s_cmp));
    if (ctx->flush) cudaFree(ctx->flush);
    ctx->flush = nullptr;
    ctx->flush_bytes = 0;
    EO_CUDA(ctx, cudaMalloc(&ctx->flush, bytes));
    ctx->flush_bytes = bytes;
  }
  eo_flush_kernel<<<ctx->sm_count * 8, 256, 0, ctx->s_cmp>>>((float4*)ctx->flush, bytes / 16);
  EO_CUDA(ctx, cudaGetLastError());
  return EO_OK;
}

int eo_assign_gather(eo_ctx* ctx, const double* values, int64_t n_values, const int64_t* src_index, int64_t n_out,
                     double* out) {
  EO_REQUIRE(ctx, ctx != nullptr, "eo_assign_gather: ctx is NULL");
  EO_REQUIRE(ctx, n_out >= 0 && n_values >= 0, "eo_assign_gather: negative size");
  if (n_out == 0) return EO_OK;
  EO_REQUIRE(ctx, values && src_index && out, "eo_assign_gather: NULL array");
  EO_REQUIRE(ctx, eo_is_device_ptr(values) && eo_is_device_ptr(src_index),
             "eo_assign_gather: values and src_index must be device arrays (the operator's result and the plan)");
  EO_CUDA(ctx, cudaSetDevice(ctx->device));
  // streamed over the OUTPUT (dof) index: `out` may be the host coefficient array, src_index stays on the device
  eo_arg args[2] = {{src_index, 8, false}, {out, 8, true}};
  return eo_run_streamed(ctx, args, 2, n_out, [&](void** a, int64_t m, int64_t) {
    eo_gather_kernel<<<unsigned((m + 255) / 256), 256, 0, ctx->s_cmp>>>(values, (const int64_t*)a[0], (double*)a[1], m);
    ctx->launches += 1;
    return EO_OK;
  });
}

// The one collective of the path.  The local record is never modified: it is snapshot on the compute
// stream, the snapshots of all ranks are gathered with ONE ncclAllGather on the collective stream, and a combine kernel
// writes the global record (SUM over the int64 block, MAX over the f64 block).  Calling it after every evaluation on a
// record that keeps accumulating therefore never counts anything twice, and the next evaluation overlaps it.
int eo_stats_collective_begin(eo_ctx* ctx, int world, const eo_stats* host_record, void** send, void** recv, void** stream) {
  EO_REQUIRE(ctx, ctx != nullptr, "eo_stats_collective_begin: ctx is NULL");
  EO_REQUIRE(ctx, world >= 1 && world <= 4096, "eo_stats_collective_begin: world size out of range");
  EO_CUDA(ctx, cudaSetDevice(ctx->device));
  if (world > ctx->stats_recv_world) {
    EO_CUDA(ctx, cudaStreamSynchronize(ctx->s_coll));
    if (ctx->stats_recv) cudaFree(ctx->stats_recv);
    ctx->stats_recv = nullptr, ctx->stats_recv_world = 0;
    EO_CUDA(ctx, cudaMalloc(&ctx->stats_recv, size_t(world) * sizeof(eo_stats)));
    ctx->stats_recv_world = world;
  }
  // the previous collective must have read the snapshot before it is overwritten
  EO_CUDA(ctx, cudaStreamWaitEvent(ctx->s_cmp, ctx->ev_coll_done, 0));
  if (host_record)
    EO_CUDA(ctx, cudaMemcpyAsync(ctx->stats_send, host_record, sizeof(eo_stats), cudaMemcpyHostToDevice, ctx->s_cmp));
  else
    EO_CUDA(ctx, cudaMemcpyAsync(ctx->stats_send, ctx->stats, sizeof(eo_stats), cudaMemcpyDeviceToDevice, ctx->s_cmp));
  EO_CUDA(ctx, cudaEventRecord(ctx->ev_coll_ready, ctx->s_cmp));
  EO_CUDA(ctx, cudaStreamWaitEvent(ctx->s_coll, ctx->ev_coll_ready, 0));
  if (send) *send = ctx->stats_send;
  if (recv) *recv = ctx->stats_recv;
  if (stream) *stream = (void*)ctx->s_coll;
  return EO_OK;
}

int eo_stats_collective_end(eo_ctx* ctx, int world) {
  EO_REQUIRE(ctx, ctx != nullptr, "eo_stats_collective_end: ctx is NULL");
  EO_REQUIRE(ctx, world >= 1 && world <= ctx->stats_recv_world, "eo_stats_collective_end: no matching eo_stats_collective_begin");
  EO_CUDA(ctx, cudaSetDevice(ctx->device));
  eo_stats_combine_kernel<<<1, 256, 0, ctx->s_coll>>>(ctx->stats_recv, world, ctx->stats_global);
  EO_CUDA(ctx, cudaGetLastError());
  EO_CUDA(ctx, cudaEventRecord(ctx->ev_coll_done, ctx->s_coll));
  ctx->launches += 1;
  return EO_OK;
}

// NCCL is resolved at run time from whatever libnccl.so.2 the process has (PyTorch's bundled one, or the system
// library): no link-time dependency.
int eo_allreduce_stats(eo_ctx* ctx, void* nccl_comm) {
  EO_REQUIRE(ctx, ctx != nullptr, "eo_allreduce_stats: ctx is NULL");
  EO_REQUIRE(ctx, nccl_comm != nullptr, "eo_allreduce_stats: communicator is NULL");
  typedef int (*allgather_fn)(const void*, void*, size_t, int, void*, cudaStream_t);
  typedef int (*count_fn)(void*, int*);
  typedef const char* (*errstr_fn)(int);
  static allgather_fn allgather = nullptr;
  static count_fn comm_count = nullptr;
  static errstr_fn errstr = nullptr;
  if (!allgather) {
    void* h = nullptr;
    const char* names[] = {getenv("EO_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
    for (const char* nm : names) {
      if (!nm || !*nm) continue;
      if ((h = dlopen(nm, RTLD_NOW | RTLD_GLOBAL))) break;
    }
    if (!h) return eo_fail(ctx, EO_ERR_UNSUPPORTED, "eo_allreduce_stats: libnccl.so.2 not found (set EO_NCCL_LIB)");
    *(void**)(&comm_count) = dlsym(h, "ncclCommCount");
    *(void**)(&errstr) = dlsym(h, "ncclGetErrorString");
    *(void**)(&allgather) = dlsym(h, "ncclAllGather");
    if (!allgather || !comm_count) {
      allgather = nullptr;
      return eo_fail(ctx, EO_ERR_UNSUPPORTED, "eo_allreduce_stats: ncclAllGather / ncclCommCount not found in libnccl");
    }
  }
  int world = 0;
  int rc = comm_count(nccl_comm, &world);
  if (rc != 0) return eo_fail(ctx, EO_ERR_CUDA, "eo_allreduce_stats: ncclCommCount: %s", errstr ? errstr(rc) : "error");
  rc = eo_stats_collective_begin(ctx, world, nullptr, nullptr, nullptr, nullptr);
  if (rc != EO_OK) return rc;
  static_assert(sizeof(eo_stats) % 8 == 0 && offsetof(eo_stats, niter_max) == (4 + EO_NITER_BINS) * 8,
                "SUM block first, MAX block after it, whole record a multiple of 8 bytes");
  const int NCCL_INT64 = 4;
  rc = allgather(ctx->stats_send, ctx->stats_recv, sizeof(eo_stats) / 8, NCCL_INT64, nccl_comm, ctx->s_coll);
  if (rc != 0) return eo_fail(ctx, EO_ERR_CUDA, "eo_allreduce_stats: ncclAllGather: %s", errstr ? errstr(rc) : "error");
  return eo_stats_collective_end(ctx, world);
}

int eo_stats_read_global(eo_ctx* ctx, eo_stats* out) {
  EO_REQUIRE(ctx, ctx && out, "eo_stats_read_global: NULL argument");
  EO_CUDA(ctx, cudaMemcpyAsync(out, ctx->stats_global, sizeof(eo_stats), cudaMemcpyDeviceToHost, ctx->s_coll));
  EO_CUDA(ctx, cudaStreamSynchronize(ctx->s_coll));
  return EO_OK;
}

int eo_debug_counters(eo_ctx* ctx, uint32_t* out) {
  EO_REQUIRE(ctx, ctx && out, "eo_debug_counters: NULL argument");
  EO_CUDA(ctx, cudaMemcpyAsync(out, ctx->work_ctr, 256, cudaMemcpyDeviceToHost, ctx->s_cmp));
  EO_CUDA(ctx, cudaStreamSynchronize(ctx->s_cmp));
  return EO_OK;
}

int eo_fp64_peak(eo_ctx* ctx, int iters, double* tflops) {
  EO_REQUIRE(ctx, ctx && tflops && iters > 0, "eo_fp64_peak: bad argument");
  EO_CUDA(ctx, cudaSetDevice(ctx->device));
  const int grid = ctx->sm_count * 8, block = 256;
  double* out = nullptr;
  EO_CUDA(ctx, cudaMalloc(&out, size_t(grid) * block * sizeof(double)));
  cudaEvent_t e0, e1;
  EO_CUDA(ctx, cudaEventCreate(&e0));
  EO_CUDA(ctx, cudaEventCreate(&e1));
  float best = 1e30f;
  for (int rep = 0; rep < 4; ++rep) {  // first repetition is the warm-up
    EO_CUDA(ctx, cudaEventRecord(e0, ctx->s_cmp));
    eo_fp64_peak_kernel<<<grid, block, 0, ctx->s_cmp>>>(out, iters, 0.999999, 1e-6);
    EO_CUDA(ctx, cudaEventRecord(e1, ctx->s_cmp));
    EO_CUDA(ctx, cudaEventSynchronize(e1));
    float ms = 0.f;
    EO_CUDA(ctx, cudaEventElapsedTime(&ms, e0, e1));
    if (rep > 0 && ms < best) best = ms;
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(out);
  *tflops = 2.0 * 8.0 * double(iters) * double(grid) * block / (double(best) * 1e-3) / 1e12;
  return EO_OK;
}

int eo_fp32_peak_variant(eo_ctx* ctx, int iters, int variant, double* tflops) {
  EO_REQUIRE(ctx, ctx && tflops && iters > 0 && variant >= 0 && variant <= 3, "eo_fp32_peak_variant: bad argument");
  EO_CUDA(ctx, cudaSetDevice(ctx->device));
  const int grid = ctx->sm_count * 8, block = 256;
  float* out = nullptr;
  EO_CUDA(ctx, cudaMalloc(&out, size_t(grid) * block * sizeof(float)));
  cudaEvent_t e0, e1;
  EO_CUDA(ctx, cudaEventCreate(&e0));
  EO_CUDA(ctx, cudaEventCreate(&e1));
  float best = 1e30f;
  for (int rep = 0; rep < 4; ++rep) {  // first repetition is the warm-up
    EO_CUDA(ctx, cudaEventRecord(e0, ctx->s_cmp));
    if (variant == 0) eo_fp32_peak_kernel<<<grid, block, 0, ctx->s_cmp>>>(out, iters, 0.999999f, 1e-6f);
    else if (variant == 1) eo_fp32_peak3_kernel<<<grid, block, 0, ctx->s_cmp>>>(out, iters, 0.999999f, 1e-6f);
    else if (variant == 3) eo_fp32_peak_u1_kernel<<<grid, block, 0, ctx->s_cmp>>>(out, iters, 0.999999f, 1e-6f);
    else eo_fp32_peak2x_kernel<<<grid, block, 0, ctx->s_cmp>>>(out, iters, 0.999999f, 1e-6f);
    EO_CUDA(ctx, cudaEventRecord(e1, ctx->s_cmp));
    EO_CUDA(ctx, cudaEventSynchronize(e1));
    float ms = 0.f;
    EO_CUDA(ctx, cudaEventElapsedTime(&ms, e0, e1));
    if (rep > 0 && ms < best) best = ms;
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(out);
  const double fma_per_trip = variant == 2 ? 16.0 : 8.0;
  *tflops = 2.0 * fma_per_trip * double(iters) * double(grid) * block / (double(best) * 1e-3) / 1e12;
  return EO_OK;
}

int eo_fp32_peak(eo_ctx* ctx, int iters, double* tflops) {
  EO_REQUIRE(ctx, ctx && tflops && iters > 0, "eo_fp32_peak: bad argument");
  EO_CUDA(ctx, cudaSetDevice(ctx->device));
  const int grid = ctx->sm_count * 8, block = 256;
  float* out = nullptr;
  EO_CUDA(ctx, cudaMalloc(&out, size_t(grid) * block * sizeof(float)));
  cudaEvent_t e0, e1;
  EO_CUDA(ctx, cudaEventCreate(&e0));
  EO_CUDA(ctx, cudaEventCreate(&e1));
  float best = 1e30f;
  for (int rep = 0; rep < 4; ++rep) {  // first repetition is the warm-up
    EO_CUDA(ctx, cudaEventRecord(e0, ctx->s_cmp));
    eo_fp32_peak_kernel<<<grid, block, 0, ctx->s_cmp>>>(out, iters, 0.999999f, 1e-6f);
    EO_CUDA(ctx, cudaEventRecord(e1, ctx->s_cmp));
    EO_CUDA(ctx, cudaEventSynchronize(e1));
    float ms = 0.f;
    EO_CUDA(ctx, cudaEventElapsedTime(&ms, e0, e1));
    if (rep > 0 && ms < best) best = ms;
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(out);
  *tflops = 2.0 * 8.0 * double(iters) * double(grid) * block / (double(best) * 1e-3) / 1e12;
  return EO_OK;
}

// ------------------------------------------------------------------ statistics
int eo_stats_reset(eo_ctx* ctx) {
  EO_REQUIRE(ctx, ctx != nullptr, "eo_stats_reset: ctx is NULL");
  EO_CUDA(ctx, cudaMemsetAsync(ctx->stats, 0, sizeof(eo_stats), ctx->s_cmp));
  static const double neg_inf = -INFINITY;  // max yielding may be negative (all points elastic)
  EO_CUDA(ctx, cudaMemcpyAsync(&ctx->stats->f_max, &neg_inf, sizeof(double), cudaMemcpyHostToDevice, ctx->s_cmp));
  return EO_OK;
}
int eo_stats_read(eo_ctx* ctx, eo_stats* out) {
  EO_REQUIRE(ctx, ctx && out, "eo_stats_read: NULL argument");
  EO_CUDA(ctx, cudaMemcpyAsync(out, ctx->stats, sizeof(eo_stats), cudaMemcpyDeviceToHost, ctx->s_cmp));
  EO_CUDA(ctx, cudaStreamSynchronize(ctx->s_cmp));
  return EO_OK;
}
void* eo_stats_device_ptr(eo_ctx* ctx) { return ctx ? (void*)ctx->stats : nullptr; }

}  // extern "C"
