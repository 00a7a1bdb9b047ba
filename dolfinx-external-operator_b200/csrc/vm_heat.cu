// vm_heat.cu - closed-form constitutive kernels (HBM-bound): von Mises radial return and the
// nonlinear heat conductivity family.  Compiled with -fmad=false so that the f64 arithmetic is
// the plain IEEE sequence of the reference's statements (bit-comparable with a non-contracting
// CPU evaluation); these kernels are bandwidth bound, the un-fused multiplies are free.
//
// Layout: one quadrature point per thread.  In the reference's AoS layout a point owns 32 B of
// strain, 32 B of old stress, 128 B of tangent and 32 B of new stress, all contiguous and 32 B
// aligned, so every access is a single LDG.E.256 / STG.E.256 per thread and a warp instruction
// covers 1 KB of consecutive, fully used sectors.
#include "eo_common.cuh"
#include "vm_core.cuh"

#include <cstdlib>

// ------------------------------------------------------------------------------------------
// von Mises                                   (reference: demo_plasticity_von_mises.py:307-326)
// ------------------------------------------------------------------------------------------
template <bool VEC, int STATE_LAYOUT>
__global__ void __launch_bounds__(256) vm_kernel(vm_consts q, const double* __restrict__ deps,
                                                 const double* __restrict__ sigma_n, const double* __restrict__ p,
                                                 double* __restrict__ C_tang, double* __restrict__ sigma,
                                                 double* __restrict__ dp_out, int64_t n, eo_stats* stats) {
  const int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
  int plastic = 0;
  if (i < n) {
    double e0, e1, e2, e3, n0, n1, n2, n3;
    if (VEC) {
      eo_d4 e = eo_ld256(deps + 4 * i);
      e0 = e.x, e1 = e.y, e2 = e.z, e3 = e.w;
    } else {
      e0 = eo_ld64(deps + 4 * i), e1 = eo_ld64(deps + 4 * i + 1), e2 = eo_ld64(deps + 4 * i + 2),
      e3 = eo_ld64(deps + 4 * i + 3);
    }
    if (STATE_LAYOUT == EO_LAYOUT_SOA) {
      n0 = eo_ld64(sigma_n + i), n1 = eo_ld64(sigma_n + n + i), n2 = eo_ld64(sigma_n + 2 * n + i),
      n3 = eo_ld64(sigma_n + 3 * n + i);
    } else if (VEC) {
      eo_d4 s = eo_ld256(sigma_n + 4 * i);
      n0 = s.x, n1 = s.y, n2 = s.z, n3 = s.w;
    } else {
      n0 = eo_ld64(sigma_n + 4 * i), n1 = eo_ld64(sigma_n + 4 * i + 1), n2 = eo_ld64(sigma_n + 4 * i + 2),
      n3 = eo_ld64(sigma_n + 4 * i + 3);
    }
    const double pi = eo_ld64(p + i);

    vm_point_out o;
    vm_point(q, e0, e1, e2, e3, n0, n1, n2, n3, pi, o);
    plastic = o.dp > 0.0;
    const double g0 = o.g[0], g1 = o.g[1], g2 = o.g[2], g3 = o.g[3], dp = o.dp;

    double* Ct = C_tang + 16 * i;
    if (VEC) {
      eo_st256(Ct + 0, o.C[0], o.C[1], o.C[2], o.C[3]);
      eo_st256(Ct + 4, o.C[4], o.C[5], o.C[6], o.C[7]);
      eo_st256(Ct + 8, o.C[8], o.C[9], o.C[10], o.C[11]);
      eo_st256(Ct + 12, o.C[12], o.C[13], o.C[14], o.C[15]);
    } else {
#pragma unroll
      for (int a = 0; a < 16; ++a) eo_st64(Ct + a, o.C[a]);
    }
    if (STATE_LAYOUT == EO_LAYOUT_SOA) {
      eo_st64(sigma + i, g0), eo_st64(sigma + n + i, g1), eo_st64(sigma + 2 * n + i, g2),
          eo_st64(sigma + 3 * n + i, g3);
    } else if (VEC) {
      eo_st256(sigma + 4 * i, g0, g1, g2, g3);
    } else {
      eo_st64(sigma + 4 * i, g0), eo_st64(sigma + 4 * i + 1, g1), eo_st64(sigma + 4 * i + 2, g2),
          eo_st64(sigma + 4 * i + 3, g3);
    }
    eo_st64(dp_out + i, dp);
  }
  eo_block_count_add(reinterpret_cast<unsigned long long*>(&stats->n_plastic), plastic);
  if (blockIdx.x == 0 && threadIdx.x == 0)
    atomicAdd(reinterpret_cast<unsigned long long*>(&stats->n_points), (unsigned long long)n);
}

// SoA history with 256-bit accesses: a thread owns FOUR consecutive points, so component c of its points is one
// 32-byte load from the SoA plane c (a warp instruction covers 1 KB of consecutive sectors, like the AoS kernel) and
// the register file does the transposition for free.  Strain and tangent stay in the reference's AoS layout (one
// 256-bit access per point and row).  Points are processed one after the other, the tangent leaves at once; only the
// new stress and dp of the four points (20 doubles) are held for the SoA stores.  Needs n % 4 == 0 (32-byte aligned
// planes); otherwise the 8-byte SoA kernel above runs.
__global__ void __launch_bounds__(128) vm_soa4_kernel(vm_consts q, const double* __restrict__ deps,
                                                      const double* __restrict__ sigma_n, const double* __restrict__ p,
                                                      double* __restrict__ C_tang, double* __restrict__ sigma,
                                                      double* __restrict__ dp_out, int64_t n, eo_stats* stats) {
  const int64_t g = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;  // group of four points
  int plastic = 0;
  if (4 * g < n) {
    const int64_t i0 = 4 * g;
    const eo_d4 s0 = eo_ld256(sigma_n + i0), s1 = eo_ld256(sigma_n + n + i0), s2 = eo_ld256(sigma_n + 2 * n + i0),
                s3 = eo_ld256(sigma_n + 3 * n + i0), pp = eo_ld256(p + i0);
    const double sn[4][4] = {{s0.x, s1.x, s2.x, s3.x}, {s0.y, s1.y, s2.y, s3.y}, {s0.z, s1.z, s2.z, s3.z}, {s0.w, s1.w, s2.w, s3.w}};
    const double pj[4] = {pp.x, pp.y, pp.z, pp.w};
    double gs[4][4], dpj[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const eo_d4 e = eo_ld256(deps + 4 * (i0 + j));
      vm_point_out o;
      vm_point(q, e.x, e.y, e.z, e.w, sn[j][0], sn[j][1], sn[j][2], sn[j][3], pj[j], o);
      plastic += o.dp > 0.0;
      double* Ct = C_tang + 16 * (i0 + j);
      eo_st256(Ct + 0, o.C[0], o.C[1], o.C[2], o.C[3]);
      eo_st256(Ct + 4, o.C[4], o.C[5], o.C[6], o.C[7]);
      eo_st256(Ct + 8, o.C[8], o.C[9], o.C[10], o.C[11]);
      eo_st256(Ct + 12, o.C[12], o.C[13], o.C[14], o.C[15]);
      gs[j][0] = o.g[0], gs[j][1] = o.g[1], gs[j][2] = o.g[2], gs[j][3] = o.g[3], dpj[j] = o.dp;
    }
#pragma unroll
    for (int c = 0; c < 4; ++c) eo_st256(sigma + c * n + i0, gs[0][c], gs[1][c], gs[2][c], gs[3][c]);
    eo_st256(dp_out + i0, dpj[0], dpj[1], dpj[2], dpj[3]);
  }
  eo_block_sum_add(reinterpret_cast<unsigned long long*>(&stats->n_plastic), plastic);
  if (blockIdx.x == 0 && threadIdx.x == 0)
    atomicAdd(reinterpret_cast<unsigned long long*>(&stats->n_points), (unsigned long long)n);
}

static int vm_launch(eo_ctx* ctx, const eo_vm_params* prm, const double* deps, const double* sigma_n, const double* p,
                     double* C_tang, double* sigma, double* dp, int64_t n, int layout) {
  vm_consts q{prm->lmbda, prm->mu, prm->H, prm->sigma_0};
  const bool vec = eo_aligned(deps, 32) && eo_aligned(C_tang, 32) &&
                   (layout == EO_LAYOUT_SOA || (eo_aligned(sigma_n, 32) && eo_aligned(sigma, 32)));
  const int block = 256;
  const int64_t grid64 = (n + block - 1) / block;
  if (grid64 > 2147483647LL) return eo_fail(ctx, EO_ERR_INVALID, "eo_vm_eval: n too large for one launch");
  const unsigned grid = (unsigned)grid64;
  static const bool soa4 = [] { const char* e = getenv("EO_VM_SOA4"); return !(e && *e == '0'); }();
  if (layout == EO_LAYOUT_SOA && soa4 && vec && n % 4 == 0 && eo_aligned(sigma_n, 32) && eo_aligned(sigma, 32) &&
      eo_aligned(p, 32) && eo_aligned(dp, 32)) {
    const int64_t groups = n / 4;
    vm_soa4_kernel<<<(unsigned)((groups + 127) / 128), 128, 0, ctx->s_cmp>>>(q, deps, sigma_n, p, C_tang, sigma, dp, n, ctx->stats);
  } else if (layout == EO_LAYOUT_SOA) {
    if (vec)
      vm_kernel<true, EO_LAYOUT_SOA><<<grid, block, 0, ctx->s_cmp>>>(q, deps, sigma_n, p, C_tang, sigma, dp, n, ctx->stats);
    else
      vm_kernel<false, EO_LAYOUT_SOA><<<grid, block, 0, ctx->s_cmp>>>(q, deps, sigma_n, p, C_tang, sigma, dp, n, ctx->stats);
  } else {
    if (vec)
      vm_kernel<true, EO_LAYOUT_AOS><<<grid, block, 0, ctx->s_cmp>>>(q, deps, sigma_n, p, C_tang, sigma, dp, n, ctx->stats);
    else
      vm_kernel<false, EO_LAYOUT_AOS><<<grid, block, 0, ctx->s_cmp>>>(q, deps, sigma_n, p, C_tang, sigma, dp, n, ctx->stats);
  }
  ctx->launches += 1;
  return EO_OK;
}

// ------------------------------------------------------------------------------------------
// history commit                                          (reference: demo_vm:564-565, demo_mc:728)
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) commit_kernel(double* __restrict__ sigma_n, const double* __restrict__ sigma,
                                                     int64_t n_sig, double* __restrict__ p,
                                                     const double* __restrict__ dp, int64_t n_p) {
  const int64_t stride = int64_t(gridDim.x) * blockDim.x;
  for (int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; i < n_sig; i += stride) sigma_n[i] = sigma[i];
  for (int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; i < n_p; i += stride) p[i] = p[i] + 1.0 * dp[i];
}

// ------------------------------------------------------------------------------------------
// heat                       (reference: part1.py:252-272, part2.py:219-261; gdim = 2, A, B runtime)
// Two points per thread so that sigma / q / dq/dT move as 256-bit and T / k as 128-bit accesses.
// ------------------------------------------------------------------------------------------
template <bool VEC>
__global__ void __launch_bounds__(256) heat_kernel(double A, double B, const double* __restrict__ T,
                                                   const double* __restrict__ sigma, double* __restrict__ k_out,
                                                   double* __restrict__ dk_out, double* __restrict__ q_out,
                                                   double* __restrict__ dqdT_out, double* __restrict__ dqds_out,
                                                   int64_t n) {
  const int64_t j = (blockIdx.x * int64_t(blockDim.x) + threadIdx.x) * 2;
  if (j >= n) return;
  if (VEC && j + 1 < n) {
    const double2 t = eo_ld128(T + j);
    const double k0 = 1.0 / (A + B * t.x), k1 = 1.0 / (A + B * t.y);
    if (k_out) eo_st128(k_out + j, k0, k1);
    if (dk_out) eo_st128(dk_out + j, -B * (k0 * k0), -B * (k1 * k1));
    if (q_out || dqdT_out) {
      const eo_d4 s = eo_ld256(sigma + 2 * j);
      if (q_out) eo_st256(q_out + 2 * j, -k0 * s.x, -k0 * s.y, -k1 * s.z, -k1 * s.w);
      if (dqdT_out)
        eo_st256(dqdT_out + 2 * j, B * (k0 * k0) * s.x, B * (k0 * k0) * s.y, B * (k1 * k1) * s.z, B * (k1 * k1) * s.w);
    }
    if (dqds_out) {
      eo_st256(dqds_out + 4 * j, -k0 * 1.0, -k0 * 0.0, -k0 * 0.0, -k0 * 1.0);
      eo_st256(dqds_out + 4 * j + 4, -k1 * 1.0, -k1 * 0.0, -k1 * 0.0, -k1 * 1.0);
    }
  } else {
    for (int64_t i = j; i < n && i < j + 2; ++i) {
      const double k = 1.0 / (A + B * eo_ld64(T + i));
      if (k_out) eo_st64(k_out + i, k);
      if (dk_out) eo_st64(dk_out + i, -B * (k * k));
      if (q_out || dqdT_out) {
        const double sx = eo_ld64(sigma + 2 * i), sy = eo_ld64(sigma + 2 * i + 1);
        if (q_out) eo_st64(q_out + 2 * i, -k * sx), eo_st64(q_out + 2 * i + 1, -k * sy);
        if (dqdT_out) eo_st64(dqdT_out + 2 * i, B * (k * k) * sx), eo_st64(dqdT_out + 2 * i + 1, B * (k * k) * sy);
      }
      if (dqds_out) {
        eo_st64(dqds_out + 4 * i, -k * 1.0), eo_st64(dqds_out + 4 * i + 1, -k * 0.0);
        eo_st64(dqds_out + 4 * i + 2, -k * 0.0), eo_st64(dqds_out + 4 * i + 3, -k * 1.0);
      }
    }
  }
}

extern "C" {

int eo_vm_eval(eo_ctx* ctx, const eo_vm_params* prm, const double* deps, const double* sigma_n, const double* p,
               double* C_tang, double* sigma, double* dp, int64_t n) {
  EO_REQUIRE(ctx, ctx != nullptr, "eo_vm_eval: ctx is NULL");
  EO_REQUIRE(ctx, prm != nullptr, "eo_vm_eval: prm is NULL");
  EO_REQUIRE(ctx, n >= 0, "eo_vm_eval: n < 0");
  if (n == 0) return EO_OK;
  EO_REQUIRE(ctx, deps && sigma_n && p && C_tang && sigma && dp, "eo_vm_eval: NULL array");
  EO_CUDA(ctx, cudaSetDevice(ctx->device));
  eo_arg args[6] = {{deps, 32, false},   {sigma_n, 32, false}, {p, 8, false},
                    {C_tang, 128, true}, {sigma, 32, true},    {dp, 8, true}};
  const eo_vm_params q = *prm;
  return eo_run_streamed(ctx, args, 6, n, [&](void** a, int64_t m, int64_t) {
    return vm_launch(ctx, &q, (const double*)a[0], (const double*)a[1], (const double*)a[2], (double*)a[3],
                     (double*)a[4], (double*)a[5], m, EO_LAYOUT_AOS);
  });
}

int eo_vm_eval_resident(eo_ctx* ctx, const eo_vm_params* prm, const double* deps, const double* sigma_n,
                        const double* p, double* C_tang, double* sigma, double* dp, int64_t n, int state_layout) {
  EO_REQUIRE(ctx, ctx != nullptr, "eo_vm_eval_resident: ctx is NULL");
  EO_REQUIRE(ctx, prm != nullptr, "eo_vm_eval_resident: prm is NULL");
  EO_REQUIRE(ctx, n >= 0, "eo_vm_eval_resident: n < 0");
  EO_REQUIRE(ctx, state_layout == EO_LAYOUT_AOS || state_layout == EO_LAYOUT_SOA,
             "eo_vm_eval_resident: unknown state_layout");
  if (n == 0) return EO_OK;
  EO_REQUIRE(ctx, deps && sigma_n && p && C_tang && sigma && dp, "eo_vm_eval_resident: NULL array");
  EO_REQUIRE(ctx,
             eo_is_device_ptr(deps) && eo_is_device_ptr(sigma_n) && eo_is_device_ptr(p) && eo_is_device_ptr(C_tang) &&
                 eo_is_device_ptr(sigma) && eo_is_device_ptr(dp),
             "eo_vm_eval_resident: all arrays must be device memory");
  EO_CUDA(ctx, cudaSetDevice(ctx->device));
  int rc = vm_launch(ctx, prm, deps, sigma_n, p, C_tang, sigma, dp, n, state_layout);
  if (rc != EO_OK) return rc;
  EO_CUDA(ctx, cudaGetLastError());
  return EO_OK;
}

int eo_commit_history(eo_ctx* ctx, double* sigma_n, const double* sigma, double* p, const double* dp, int64_t n,
                      int ncomp) {
  EO_REQUIRE(ctx, ctx != nullptr, "eo_commit_history: ctx is NULL");
  EO_REQUIRE(ctx, n >= 0 && ncomp >= 0, "eo_commit_history: negative size");
  if (n == 0) return EO_OK;
  EO_REQUIRE(ctx, (sigma_n == nullptr) == (sigma == nullptr), "eo_commit_history: sigma_n/sigma must both be given");
  EO_REQUIRE(ctx, (p == nullptr) == (dp == nullptr), "eo_commit_history: p/dp must both be given");
  EO_REQUIRE(ctx, (!sigma_n || (eo_is_device_ptr(sigma_n) && eo_is_device_ptr(sigma))) &&
                      (!p || (eo_is_device_ptr(p) && eo_is_device_ptr(dp))),
             "eo_commit_history: device memory only");
  EO_CUDA(ctx, cudaSetDevice(ctx->device));
  commit_kernel<<<ctx->sm_count * 8, 256, 0, ctx->s_cmp>>>(sigma_n, sigma, sigma_n ? n * ncomp : 0, p, dp, p ? n : 0);
  ctx->launches += 1;
  EO_CUDA(ctx, cudaGetLastError());
  return EO_OK;
}

int eo_heat_eval(eo_ctx* ctx, double A, double B, const double* T, const double* sigma, double* k, double* dk,
                 double* q, double* dqdT, double* dqdsigma, int64_t n) {
  EO_REQUIRE(ctx, ctx != nullptr, "eo_heat_eval: ctx is NULL");
  EO_REQUIRE(ctx, n >= 0, "eo_heat_eval: n < 0");
  if (n == 0) return EO_OK;
  EO_REQUIRE(ctx, T != nullptr, "eo_heat_eval: T is NULL");
  EO_REQUIRE(ctx, k || dk || q || dqdT || dqdsigma, "eo_heat_eval: no output requested");
  EO_REQUIRE(ctx, sigma || !(q || dqdT), "eo_heat_eval: sigma is required for q / dqdT");
  EO_CUDA(ctx, cudaSetDevice(ctx->device));
  // sigma is only read for q / dqdT: do not stage it otherwise
  const double* sig_used = (q || dqdT) ? sigma : nullptr;
  eo_arg args[7] = {{T, 8, false}, {sig_used, 16, false}, {k, 8, true},       {dk, 8, true},
                    {q, 16, true}, {dqdT, 16, true},      {dqdsigma, 32, true}};
  return eo_run_streamed(ctx, args, 7, n, [&](void** a, int64_t m, int64_t) {
    bool vec = true;
    const size_t al[7] = {16, 32, 16, 16, 32, 32, 32};
    for (int i = 0; i < 7; ++i) vec = vec && (a[i] == nullptr || eo_aligned(a[i], al[i]));
    const int block = 256;
    const int64_t grid64 = ((m + 1) / 2 + block - 1) / block;
    if (grid64 > 2147483647LL) return eo_fail(ctx, EO_ERR_INVALID, "eo_heat_eval: n too large for one launch");
    if (vec)
      heat_kernel<true><<<(unsigned)grid64, block, 0, ctx->s_cmp>>>(A, B, (const double*)a[0], (const double*)a[1],
                                                                    (double*)a[2], (double*)a[3], (double*)a[4],
                                                                    (double*)a[5], (double*)a[6], m);
    else
      heat_kernel<false><<<(unsigned)grid64, block, 0, ctx->s_cmp>>>(A, B, (const double*)a[0], (const double*)a[1],
                                                                     (double*)a[2], (double*)a[3], (double*)a[4],
                                                                     (double*)a[5], (double*)a[6], m);
    ctx->launches += 1;
    return (int)EO_OK;
  });
}

}  // extern "C"
