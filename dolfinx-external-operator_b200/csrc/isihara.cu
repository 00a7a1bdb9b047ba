// isihara.cu - Isihara ICNN hyperelastic model on sm_100a.
// replaces: `vectorized_stress_and_tangent = vmap(jacfwd(compute_stress_local, has_aux=True))` and
//           `dP_dF_impl`, doc/demo/demo_hyperelasticity.py:429-456 (network :242-307, corrections :362-381).
// The per-point arithmetic is in isihara_core.cuh.  One thread per quadrature point; the preprocessed network
// (isi_weights, 35 KB: the one remaining 64x64 layer in both orientations plus the collapsed layer 1) is staged
// once per CTA in shared memory and read as warp-wide broadcasts (every lane needs the same weight at the
// same time), so the inner loops are float32 FMA bound: ~21 k FMA per point in five 64x64 matrix-vector
// products, the invariants and the chain rule back to F in float64.  Too small and irregular per point for
// tensor cores (north_star) - and TF32 inputs would not keep the float32 parity of the reference network.
#include "eo_common.cuh"
#include "isihara_core.cuh"

struct eo_isihara {
  eo_ctx* ctx = nullptr;
  isi_weights_c* d_w = nullptr;  // device copy (compact: without W2)
  isi_weights h_w;
};

static void isi_compact(const isi_weights& w, isi_weights_c& c) {
  memcpy(c.A1, w.A1, sizeof c.A1);
  memcpy(c.S2, w.S2, sizeof c.S2);
  memcpy(c.W2T, w.W2T, sizeof c.W2T);
  memcpy(c.w3, w.w3, sizeof c.w3);
  memcpy(c.s3, w.s3, sizeof c.s3);
  memcpy(c.H, w.H, sizeof c.H);
}

#define ISI_THREADS 384  // one CTA per SM (12 warps, 168 registers): 19 KB of weights + 192 KB of activation scratch
#define ISI_SMEM (sizeof(isi_weights_c) + size_t(ISI_NH) * ISI_THREADS * 2 * sizeof(float))

__global__ void __launch_bounds__(ISI_THREADS, 1) isihara_kernel(const isi_weights_c* __restrict__ gw,
                                                              const double* __restrict__ F, double* __restrict__ dP,
                                                              double* __restrict__ P, int64_t n) {
  extern __shared__ __align__(16) unsigned char s_raw[];
  isi_weights_c& W = *reinterpret_cast<isi_weights_c*>(s_raw);
  // layer-1 activation scratch: [unit][thread][u, w] - consecutive threads are 2 words apart: one 64-bit access per unit
  float* zs = reinterpret_cast<float*>(s_raw + sizeof(isi_weights_c)) + 2 * threadIdx.x;
  {
    const int nw = int(sizeof(isi_weights_c) / 16);
    const float4* src = reinterpret_cast<const float4*>(gw);
    float4* dst = reinterpret_cast<float4*>(s_raw);
    for (int t = threadIdx.x; t < nw; t += blockDim.x) dst[t] = src[t];
  }
  __syncthreads();
  const int64_t stride = int64_t(gridDim.x) * blockDim.x;
  for (int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; i < n; i += stride) {
    const eo_d4 f = eo_ld256(F + 4 * i);
    const double Fv[4] = {f.x, f.y, f.z, f.w};
    double Pv[4], T[16];
    isi_point(W, Fv, Pv, T, zs, 2 * ISI_THREADS);
    double* o = dP + 16 * i;
    eo_st256(o + 0, T[0], T[1], T[2], T[3]);
    eo_st256(o + 4, T[4], T[5], T[6], T[7]);
    eo_st256(o + 8, T[8], T[9], T[10], T[11]);
    eo_st256(o + 12, T[12], T[13], T[14], T[15]);
    eo_st256(P + 4 * i, Pv[0], Pv[1], Pv[2], Pv[3]);
  }
}

static bool g_isi_attr = false;

extern "C" {

int eo_isihara_create(eo_ctx* ctx, const eo_isihara_weights* w, eo_isihara** out) {
  EO_REQUIRE(ctx, ctx != nullptr, "eo_isihara_create: ctx is NULL");
  EO_REQUIRE(ctx, w && out, "eo_isihara_create: NULL argument");
  static_assert(sizeof(eo_isihara_weights) == sizeof(isi_weights), "eo_isihara_weights must mirror isi_weights");
  static_assert(sizeof(isi_weights_c) % 16 == 0, "isi_weights_c is copied in 16-byte pieces");
  *out = nullptr;
  EO_CUDA(ctx, cudaSetDevice(ctx->device));
  eo_isihara* m = new eo_isihara();
  m->ctx = ctx;
  memcpy(&m->h_w, w, sizeof(isi_weights));
  isi_weights_c hc;
  isi_compact(m->h_w, hc);
  cudaError_t e = cudaMalloc(&m->d_w, sizeof(isi_weights_c));
  if (e == cudaSuccess) e = cudaMemcpyAsync(m->d_w, &hc, sizeof(isi_weights_c), cudaMemcpyHostToDevice, ctx->s_cmp);
  if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->s_cmp);
  if (e == cudaSuccess && !g_isi_attr) {
    e = cudaFuncSetAttribute(isihara_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ISI_SMEM);
    g_isi_attr = e == cudaSuccess;
  }
  if (e != cudaSuccess) {
    if (m->d_w) cudaFree(m->d_w);
    delete m;
    return eo_fail(ctx, e == cudaErrorMemoryAllocation ? EO_ERR_NOMEM : EO_ERR_CUDA, "eo_isihara_create: %s",
                   cudaGetErrorString(e));
  }
  *out = m;
  return EO_OK;
}

int eo_isihara_destroy(eo_isihara* m) {
  if (!m) return EO_OK;
  cudaSetDevice(m->ctx->device);
  cudaStreamSynchronize(m->ctx->s_cmp);
  if (m->d_w) cudaFree(m->d_w);
  delete m;
  return EO_OK;
}

int eo_isihara_set_correction(eo_isihara* m, const double H_flat[4]) {
  if (!m || !H_flat) return EO_ERR_INVALID;
  eo_ctx* ctx = m->ctx;
  for (int i = 0; i < 4; ++i) m->h_w.H[i] = H_flat[i];
  isi_weights_c hc;
  isi_compact(m->h_w, hc);
  EO_CUDA(ctx, cudaMemcpyAsync(m->d_w, &hc, sizeof(isi_weights_c), cudaMemcpyHostToDevice, ctx->s_cmp));
  EO_CUDA(ctx, cudaStreamSynchronize(ctx->s_cmp));
  return EO_OK;
}

static void isi_launch(eo_isihara* m, const double* F, double* dP, double* P, int64_t cnt, cudaStream_t stream) {
  int64_t grid = (cnt + ISI_THREADS - 1) / ISI_THREADS;
  const int64_t cap = int64_t(m->ctx->sm_count) * 2;  // persistent: one resident CTA per SM, grid-stride over the points
  if (grid > cap) grid = cap;
  isihara_kernel<<<(unsigned)grid, ISI_THREADS, ISI_SMEM, stream>>>(m->d_w, F, dP, P, cnt);
  m->ctx->launches += 1;
}

int eo_isihara_eval_on_stream(eo_isihara* m, const double* F, double* dP, double* P, int64_t n, void* stream) {
  if (!m) return EO_ERR_INVALID;
  eo_ctx* ctx = m->ctx;
  EO_REQUIRE(ctx, n >= 0, "eo_isihara_eval_on_stream: n < 0");
  if (n == 0) return EO_OK;
  EO_REQUIRE(ctx, F && dP && P, "eo_isihara_eval_on_stream: NULL array");
  EO_REQUIRE(ctx, eo_is_device_ptr(F) && eo_is_device_ptr(dP) && eo_is_device_ptr(P),
             "eo_isihara_eval_on_stream: arrays must be device memory (the caller's stream orders them)");
  EO_REQUIRE(ctx, eo_aligned(F, 32) && eo_aligned(dP, 32) && eo_aligned(P, 32), "eo_isihara_eval_on_stream: arrays must be 32-byte aligned");
  EO_CUDA(ctx, cudaSetDevice(ctx->device));
  isi_launch(m, F, dP, P, n, (cudaStream_t)stream);
  EO_CUDA(ctx, cudaGetLastError());
  return EO_OK;
}

int eo_isihara_eval(eo_isihara* m, const double* F, double* dP, double* P, int64_t n) {
  if (!m) return EO_ERR_INVALID;
  eo_ctx* ctx = m->ctx;
  EO_REQUIRE(ctx, n >= 0, "eo_isihara_eval: n < 0");
  if (n == 0) return EO_OK;
  EO_REQUIRE(ctx, F && dP && P, "eo_isihara_eval: NULL array");
  EO_CUDA(ctx, cudaSetDevice(ctx->device));
  eo_arg args[3] = {{F, 32, false}, {dP, 128, true}, {P, 32, true}};
  return eo_run_streamed(ctx, args, 3, n, [&](void** a, int64_t cnt, int64_t) {
    for (int i = 0; i < 3; ++i)
      if (!eo_aligned(a[i], 32)) return eo_fail(ctx, EO_ERR_INVALID, "eo_isihara_eval: arrays must be 32-byte aligned");
    isi_launch(m, (const double*)a[0], (double*)a[1], (double*)a[2], cnt, ctx->s_cmp);
    return (int)EO_OK;
  });
}

}  // extern "C"
