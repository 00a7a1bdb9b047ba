// isihara_core.cuh - per-quadrature-point arithmetic of the Isihara input-convex neural network
// hyperelastic model (reference: doc/demo/demo_hyperelasticity.py:242-307 network, :362-381 corrections,
// :429-456 `compute_stress_local` / `dP_dF_impl`; citations ":NNN" are to that file).  Host/device header:
// isihara.cu builds the sm_100a kernel from it, tests/hostcheck/ compiles it with g++ for CPU checks.
//
// What the reference computes:  W_NN(F) = ICNN(K1, K2, K3)(C = F^T F) with the network in float32 and the
// invariants in the dtype of F (float64), P = dW_NN/dF + F @ H by `torch.func.grad`, tangent = dP/dF by
// `jacfwd`, batched by `vmap`.  Written out here:
//   * features x = (K1, K2, K3) and their first and second derivatives w.r.t. the four components of F are
//     propagated in float64 with a 2nd-order Taylor jet (15 coefficients) - the reference's AD does the same
//     arithmetic in float64 on that side of the `.float()` cast (:286);
//   * the network  z0 = L0 x + b0 (no activation, :289) ; a1 = sp(W1)^T z0 + Skip1(x) ; z1 = softplus(a1)^2/12 ;
//     a2 = sp(W2)^T z1 + Skip2(x) ; z2 = softplus(a2)^2/12 ; y = sp(W3)^T z2 + sp(Ws3)^T x  (:288-300)
//     runs in float32.  Because z0 is affine in x, layer 1 collapses on the host to a1 = A1 x + c1
//     (A1 = sp(W1)^T L0 + Skip1, 64x3), leaving ONE 64x64 layer.  Gradient and Hessian of y w.r.t. x:
//       g_j = w3_j phi'(a2_j),  v = W2^T g,  d_j = da2_j/dx = sum_i W2_ji phi'(a1_i) A1_i + S2_j
//       dy/dx   = s3 + S2^T g + A1^T (phi'(a1) . v)
//       d2y/dx2 = sum_j w3_j phi''(a2_j) d_j d_j^T + sum_i v_i phi''(a1_i) A1_i A1_i^T
//     i.e. five 64x64 matrix-vector products per point instead of differentiating twice through the graph;
//   * chain rule back to F in float64, plus the constant corrections P += F @ H, tangent += H^T (:436-441).
#pragma once

#include <cmath>
#include <cstdint>

#if defined(__CUDACC__)
#define EO_ISI_HD __host__ __device__ __forceinline__
#else
#define EO_ISI_HD inline
#endif

#define ISI_NH 64  // hidden width (n_hidden = [64, 64, 64], :304)

// network constants after the host-side preprocessing (softplus of the convex weights applied once, :238)
struct isi_weights {
  float A1[ISI_NH][4];      // a1 = A1 x + c1: columns 0..2 = A1, column 3 = c1
  float S2[ISI_NH][4];      // skip of layer 2: columns 0..2 = weight, column 3 = bias
  float W2[ISI_NH][ISI_NH];   // softplus(layers.2.weights): a2_j = sum_i W2[j][i] z1_i + ...
  float W2T[ISI_NH][ISI_NH];  // its transpose (v_i = sum_j W2T[i][j] g_j)
  float w3[ISI_NH];         // softplus(layers.3.weights)
  float s3[4];              // softplus(skip_layers.3.weights), padded
  double H[4];              // H_flat = -dW_NN/dF at F = I (:367)
};

// what the kernel keeps in shared memory: the same constants without W2 (both passes read W2T rows)
struct isi_weights_c {
  float A1[ISI_NH][4];
  float S2[ISI_NH][4];
  float W2T[ISI_NH][ISI_NH];
  float w3[ISI_NH];
  float s3[4];
  double H[4];
};

struct __attribute__((aligned(8))) isi_f2 {
  float x, y;
};

// softplus and the activation phi(a) = softplus(a)^2 / 12 with its first two derivatives, float32.
// torch.nn.functional.softplus is log1p(exp(a)), linear above the threshold 20.  Evaluated here in the symmetric form
//     t = exp(-|a|) in (0, 1],   softplus(a) = max(a, 0) + log1p(t),   sigmoid(a) = (a >= 0 ? 1 : t) / (1 + t)
// which needs no threshold (a > 20: log1p(t) < 2.1e-9 is below half an ulp of a, the sum rounds to a and the sigmoid to 1,
// the reference's values) and keeps every rounding error ABSOLUTE and small: the device path uses the SFU approximations
// ex2 / lg2 / rcp (three MUFU + ~10 FP32 instructions instead of ~45 for expf + log1pf + a division; the 128 evaluations
// per point were 23 % of the kernel's instructions).  Error budget: t carries a relative error |a| 2^-23 from the rounded
// product a log2(e), which moves softplus by at most 0.28 x 2^-23; lg2.approx is within 2^-22 absolute on (1, 2], i.e.
// softplus within 1.7e-7 absolute where it is read from the SFU (t >= 2^-7), and the series t - t^2/2 + t^3/3 (error
// < 1e-9) is used below.  Measured against the reference golden: tests/isi_util.py.
EO_ISI_HD void isi_softplus(float a, float& sp, float& sg) {
#if defined(__CUDA_ARCH__)
  float t, lg, r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(fabsf(a) * -1.4426950408889634f));
  const float u = 1.0f + t;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(lg) : "f"(u));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(u));
  const float series = fmaf(fmaf(t, 1.0f / 3.0f, -0.5f) * t, t, t);
  const float l1p = t < 0.0078125f ? series : lg * 0.6931471805599453f;
#else
  const float t = expf(-fabsf(a));
  const float l1p = log1pf(t), r = 1.0f / (1.0f + t);
#endif
  sp = fmaxf(a, 0.0f) + l1p;
  sg = (a >= 0.0f ? 1.0f : t) * r;  // sigmoid = d softplus / da
}
EO_ISI_HD void isi_phi(float a, float& p0, float& p1, float& p2) {
  float sp, sg;
  isi_softplus(a, sp, sg);
  const float sg1 = sg * (1.0f - sg);  // d sigmoid / da
  p0 = sp * sp * (1.0f / 12.0f);
  p1 = sp * sg * (1.0f / 6.0f);
  p2 = (sg * sg + sp * sg1) * (1.0f / 6.0f);
}
// The layer-1 activations are kept between the passes as TWO floats per unit, u = softplus / sqrt(12) and
// w = sigmoid / sqrt(3) (512 bytes of scratch per point instead of 768: 12 instead of 8 warps per SM), from which
//     phi = u^2,   phi' = u w,   phi'' = w^2 / 2 + phi' (1 - sqrt(3) w)
// cost one or two multiplications where they are used.
#define ISI_C_U 0.28867513459481287f  // 1 / sqrt(12)
#define ISI_C_W 0.57735026918962584f  // 1 / sqrt(3)
#define ISI_SQRT3 1.7320508075688772f

// Two float32 FMAs in one instruction: sm_100's packed FFMA2 (fma.rn.f32x2).  A three-register scalar FFMA issues
// every second cycle per SM sub-partition on Blackwell (B300_MICROARCH.md: rt_SMSP = 2), i.e. 64 FMA/clk/SM; the packed
// form carries two FMAs at the same issue cost and so reaches the 128 FMA/clk/SM of the pipe.  Each half is an IEEE
// fma: bit-identical to two scalar FFMAs.  Host build: plain multiply-add.
EO_ISI_HD isi_f2 isi_fma2(isi_f2 a, isi_f2 b, isi_f2 c) {
#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ >= 1000)
  const float2 r = __ffma2_rn(make_float2(a.x, a.y), make_float2(b.x, b.y), make_float2(c.x, c.y));
  return isi_f2{r.x, r.y};
#else
  return isi_f2{a.x * b.x + c.x, a.y * b.y + c.y};
#endif
}

// y(x), dy/dx (3), d2y/dx2 (6: xx, xy, xz, yy, yz, zz) of the network, float32.
// `W` may live in shared memory (device) or anywhere (host).  `zs` is per-point scratch for the layer-1
// activations ((u, w) of the 64 units, see above, computed ONCE): the 8-byte cell of unit i at zs + i * zstride
// (8-byte aligned) - shared memory on the device (consecutive threads 2 words apart: one 64-bit access per unit, no
// bank conflicts), a local array on the host.  WT: isi_weights or isi_weights_c.
template <class WT>
EO_ISI_HD void isi_network(const WT& W, const float x[3], float* zs, int zstride, float& y, float gx[3],
                           float hx[6]) {
  float g[ISI_NH];  // g_j = w3_j phi'(a2_j)
  float yy = W.s3[0] * x[0] + W.s3[1] * x[1] + W.s3[2] * x[2];
  float G0 = W.s3[0], G1 = W.s3[1], G2 = W.s3[2];
  float h0 = 0.f, h1 = 0.f, h2 = 0.f, h3 = 0.f, h4 = 0.f, h5 = 0.f;
  // ---- layer 1 (collapsed onto the affine layer 0): a1 = A1 x + c1, activations kept for both passes
#pragma unroll 4
  for (int i = 0; i < ISI_NH; ++i) {
    const float a1 = W.A1[i][0] * x[0] + W.A1[i][1] * x[1] + W.A1[i][2] * x[2] + W.A1[i][3];
    float sp, sg;
    isi_softplus(a1, sp, sg);
    *reinterpret_cast<isi_f2*>(zs + i * zstride) = isi_f2{sp * ISI_C_U, sg * ISI_C_W};
  }
  // ---- pass A: a2_j and d_j = da2_j/dx for blocks of 16 outputs (64 accumulators in registers, paired along j for
  //      the packed FMA); weights read as W2T[i][jb .. jb+15]: contiguous, four 128-bit broadcasts per 32 FFMA2
#pragma unroll 1
  for (int jb = 0; jb < ISI_NH; jb += 16) {
    isi_f2 acc[8][4];  // acc[m][k] = (component k of output jb + 2m, of output jb + 2m + 1)
#pragma unroll
    for (int m = 0; m < 8; ++m) {
      const int j0 = jb + 2 * m, j1 = j0 + 1;
      acc[m][0] = isi_f2{W.S2[j0][0] * x[0] + W.S2[j0][1] * x[1] + W.S2[j0][2] * x[2] + W.S2[j0][3],
                         W.S2[j1][0] * x[0] + W.S2[j1][1] * x[1] + W.S2[j1][2] * x[2] + W.S2[j1][3]};
      acc[m][1] = isi_f2{W.S2[j0][0], W.S2[j1][0]};
      acc[m][2] = isi_f2{W.S2[j0][1], W.S2[j1][1]};
      acc[m][3] = isi_f2{W.S2[j0][2], W.S2[j1][2]};
    }
#pragma unroll 2
    for (int i = 0; i < ISI_NH; ++i) {
      const isi_f2 uw = *reinterpret_cast<const isi_f2*>(zs + i * zstride);
      const float p1 = uw.x * uw.y;
      const float z0 = uw.x * uw.x, z1 = p1 * W.A1[i][0], z2 = p1 * W.A1[i][1], z3 = p1 * W.A1[i][2];
      const isi_f2 zz[4] = {{z0, z0}, {z1, z1}, {z2, z2}, {z3, z3}};
#pragma unroll
      for (int m = 0; m < 8; ++m) {
        const isi_f2 w{W.W2T[i][jb + 2 * m], W.W2T[i][jb + 2 * m + 1]};
#pragma unroll
        for (int k = 0; k < 4; ++k) acc[m][k] = isi_fma2(w, zz[k], acc[m][k]);
      }
    }
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      const int m = j >> 1;
      const float a2 = (j & 1) ? acc[m][0].y : acc[m][0].x;
      const float d0 = (j & 1) ? acc[m][1].y : acc[m][1].x, d1 = (j & 1) ? acc[m][2].y : acc[m][2].x,
                  d2 = (j & 1) ? acc[m][3].y : acc[m][3].x;
      float p0, p1, p2;
      isi_phi(a2, p0, p1, p2);
      const float w3 = W.w3[jb + j];
      yy += w3 * p0;
      const float gj = w3 * p1;
      g[jb + j] = gj;
      G0 += W.S2[jb + j][0] * gj, G1 += W.S2[jb + j][1] * gj, G2 += W.S2[jb + j][2] * gj;
      const float c = w3 * p2;
      h0 += c * d0 * d0, h1 += c * d0 * d1, h2 += c * d0 * d2, h3 += c * d1 * d1, h4 += c * d1 * d2, h5 += c * d2 * d2;
    }
  }
  // ---- pass B: v = W2^T g (two packed partial sums per unit: same pairwise order on host and device), then the
  //      layer-1 contributions
#pragma unroll 2
  for (int i = 0; i < ISI_NH; ++i) {
    isi_f2 v2{0.f, 0.f};
#pragma unroll
    for (int j = 0; j < ISI_NH; j += 2) v2 = isi_fma2(isi_f2{W.W2T[i][j], W.W2T[i][j + 1]}, isi_f2{g[j], g[j + 1]}, v2);
    const float v = v2.x + v2.y;
    const isi_f2 uw = *reinterpret_cast<const isi_f2*>(zs + i * zstride);
    const float p1 = uw.x * uw.y, p2 = fmaf(0.5f * uw.y, uw.y, p1 * fmaf(-ISI_SQRT3, uw.y, 1.0f));
    const float A0 = W.A1[i][0], A1 = W.A1[i][1], A2 = W.A1[i][2];
    const float t1 = p1 * v, t2 = p2 * v;
    G0 += A0 * t1, G1 += A1 * t1, G2 += A2 * t1;
    h0 += t2 * A0 * A0, h1 += t2 * A0 * A1, h2 += t2 * A0 * A2, h3 += t2 * A1 * A1, h4 += t2 * A1 * A2, h5 += t2 * A2 * A2;
  }
  y = yy;
  gx[0] = G0, gx[1] = G1, gx[2] = G2;
  hx[0] = h0, hx[1] = h1, hx[2] = h2, hx[3] = h3, hx[4] = h4, hx[5] = h5;
}

// ------------------------------------------------------------------------------------------------
// float64 2nd-order Taylor jets in the four components of F: value, gradient (4), Hessian (10, i <= j)
// ------------------------------------------------------------------------------------------------
#define ISI_SYM(i, j) ((i) <= (j) ? ((i) * (7 - (i)) / 2 + (j)) : ((j) * (7 - (j)) / 2 + (i)))

struct isi_jet {
  double v, g[4], h[10];
};

EO_ISI_HD isi_jet isi_var(double val, int k) {
  isi_jet r;
  r.v = val;
#pragma unroll
  for (int i = 0; i < 4; ++i) r.g[i] = (i == k) ? 1.0 : 0.0;
#pragma unroll
  for (int i = 0; i < 10; ++i) r.h[i] = 0.0;
  return r;
}
EO_ISI_HD isi_jet isi_add(const isi_jet& a, const isi_jet& b, double sb = 1.0) {
  isi_jet r;
  r.v = a.v + sb * b.v;
#pragma unroll
  for (int i = 0; i < 4; ++i) r.g[i] = a.g[i] + sb * b.g[i];
#pragma unroll
  for (int i = 0; i < 10; ++i) r.h[i] = a.h[i] + sb * b.h[i];
  return r;
}
EO_ISI_HD isi_jet isi_addc(const isi_jet& a, double c) {
  isi_jet r = a;
  r.v += c;
  return r;
}
EO_ISI_HD isi_jet isi_mul(const isi_jet& a, const isi_jet& b) {
  isi_jet r;
  r.v = a.v * b.v;
#pragma unroll
  for (int i = 0; i < 4; ++i) r.g[i] = a.g[i] * b.v + a.v * b.g[i];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = i; j < 4; ++j)
      r.h[ISI_SYM(i, j)] = a.h[ISI_SYM(i, j)] * b.v + a.v * b.h[ISI_SYM(i, j)] + a.g[i] * b.g[j] + a.g[j] * b.g[i];
  return r;
}
// f(u) with f0 = f(u.v), f1 = f'(u.v), f2 = f''(u.v)
EO_ISI_HD isi_jet isi_compose(const isi_jet& u, double f0, double f1, double f2) {
  isi_jet r;
  r.v = f0;
#pragma unroll
  for (int i = 0; i < 4; ++i) r.g[i] = f1 * u.g[i];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = i; j < 4; ++j) r.h[ISI_SYM(i, j)] = f2 * u.g[i] * u.g[j] + f1 * u.h[ISI_SYM(i, j)];
  return r;
}
// u^p with the value f0 = pow(u.v, p) supplied by the caller
EO_ISI_HD isi_jet isi_pow(const isi_jet& u, double p, double f0) {
  return isi_compose(u, f0, p * f0 / u.v, p * (p - 1.0) * f0 / (u.v * u.v));
}

// One point: F = [F11, F12, F21, F22] (:263-266)  ->  P (4), tangent dP_i/dF_j (row-major 4x4)
// zs / zstride: scratch for isi_network (2 * ISI_NH floats per point when zstride = 2), 8-byte aligned
template <class WT>
EO_ISI_HD void isi_point(const WT& W, const double F[4], double P[4], double dP[16], float* zs, int zstride) {
  // features (:269-283) as plain values first: the 2nd-order jets (42 doubles) are built AFTER the network so that they
  // do not occupy registers while it runs; the three powers of I3 are computed once and handed on
  double pw[3];
  float x[3];
  {
    const double C11 = F[0] * F[0] + F[2] * F[2], C12 = F[0] * F[1] + F[2] * F[3], C22 = F[1] * F[1] + F[3] * F[3];
    const double C1221 = C12 * C12, C1122 = C11 * C22;
    const double I1 = (C11 + C22) + 1.0, I2 = ((C11 + C22) - C1221) + C1122, I3 = C1122 - C1221;
    pw[0] = pow(I3, -1.0 / 3.0), pw[1] = pow(I3, -2.0 / 3.0), pw[2] = pow(I3, 0.5);
    const double Jm1 = pw[2] - 1.0;
    x[0] = (float)(I1 * pw[0] - 3.0), x[1] = (float)(I2 * pw[1] - 3.0), x[2] = (float)(Jm1 * Jm1);
  }
  // network in float32 (:286)
  float y, gx[3], hx[6];
  isi_network(W, x, zs, zstride, y, gx, hx);
  const isi_jet F11 = isi_var(F[0], 0), F12 = isi_var(F[1], 1), F21 = isi_var(F[2], 2), F22 = isi_var(F[3], 3);
  // right Cauchy-Green tensor and invariants (:269-277)
  const isi_jet C11 = isi_add(isi_mul(F11, F11), isi_mul(F21, F21));
  const isi_jet C12 = isi_add(isi_mul(F11, F12), isi_mul(F21, F22));
  const isi_jet C22 = isi_add(isi_mul(F12, F12), isi_mul(F22, F22));
  const isi_jet C1221 = isi_mul(C12, C12);
  const isi_jet C1122 = isi_mul(C11, C22);
  const isi_jet I1 = isi_addc(isi_add(C11, C22), 1.0);
  const isi_jet I2 = isi_add(isi_add(isi_add(C11, C22), C1221, -1.0), C1122);
  const isi_jet I3 = isi_add(C1122, C1221, -1.0);
  // features (:280-283)
  isi_jet X[3];
  X[0] = isi_addc(isi_mul(I1, isi_pow(I3, -1.0 / 3.0, pw[0])), -3.0);
  X[1] = isi_addc(isi_mul(I2, isi_pow(I3, -2.0 / 3.0, pw[1])), -3.0);
  {
    const isi_jet J = isi_pow(I3, 0.5, pw[2]);
    const isi_jet Jm1 = isi_addc(J, -1.0);
    X[2] = isi_mul(Jm1, Jm1);
  }
  const double Wx[3] = {(double)gx[0], (double)gx[1], (double)gx[2]};
  const double Wxx[3][3] = {{(double)hx[0], (double)hx[1], (double)hx[2]},
                            {(double)hx[1], (double)hx[3], (double)hx[4]},
                            {(double)hx[2], (double)hx[4], (double)hx[5]}};
  // chain rule to F, then the corrections P += F @ H, dP += H^T with H the 4x4 block matrix of :368-376
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    double acc = 0.0;
#pragma unroll
    for (int k = 0; k < 3; ++k) acc += Wx[k] * X[k].g[i];
    P[i] = acc;
  }
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      double acc = 0.0;
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        acc += Wx[k] * X[k].h[ISI_SYM(i, j)];
#pragma unroll
        for (int l = 0; l < 3; ++l) acc += Wxx[k][l] * X[k].g[i] * X[l].g[j];
      }
      dP[4 * i + j] = acc;
    }
  const double h0 = W.H[0], h1 = W.H[1], h2 = W.H[2], h3 = W.H[3];
  // Hm = [[h0,h1,0,0],[h2,h3,0,0],[0,0,h0,h1],[0,0,h2,h3]];  P_j += sum_i F_i Hm[i][j];  dP[j][i] += Hm[i][j]
  P[0] += F[0] * h0 + F[1] * h2;
  P[1] += F[0] * h1 + F[1] * h3;
  P[2] += F[2] * h0 + F[3] * h2;
  P[3] += F[2] * h1 + F[3] * h3;
  dP[4 * 0 + 0] += h0, dP[4 * 0 + 1] += h2;
  dP[4 * 1 + 0] += h1, dP[4 * 1 + 1] += h3;
  dP[4 * 2 + 2] += h0, dP[4 * 2 + 3] += h2;
  dP[4 * 3 + 2] += h1, dP[4 * 3 + 3] += h3;
}
