// Training-step kernels (Pyfer.training_step = forward + GenLoss + backward through net_B + AdamW).
//
// Everything here is HBM-bound elementwise / reduction work in the same c8 layout as the forward pass:
//   - backward of the fused InstanceNorm/BatchNorm(train) + activation + residual unit (two passes: per-(n,c)
//     reductions, then the apply pass), recomputing the forward values from the saved fp32 pre-norm tensor;
//   - batch combination of statistics for train-mode BatchNorm3d (+ running-stat update);
//   - weight / bias gradients of the 1x1x1 convolutions, the 2x transposed convolutions and the dose heads;
//   - data gradient of the 2x transposed convolution;
//   - deep-supervised masked L1 loss (GenLoss) forward + backward, fused AdamW.
// Upstream gradients arrive as a sum of up to three fp32 c8 tensors (one per consumer of the forward tensor)
// and/or one fp16 c8 tensor; gradients handed to tensor-core dgrad/wgrad kernels leave as fp16 c8 tensors
// (static loss scaling keeps them inside the fp16 range).  Small parameter gradients are accumulated with
// fp64 atomics into a scratch arena and converted once (dp_grad_finalize).
#include <algorithm>

#include "common.cuh"
#include "dose_b200.h"

namespace dp {

__device__ __forceinline__ void t_load8(const __half* hi, const __half* lo, size_t off, float (&x)[8]) {
  const uint4 u = *reinterpret_cast<const uint4*>(hi + off);
  const __half2* h = reinterpret_cast<const __half2*>(&u);
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float2 f = __half22float2(h[j]);
    x[2 * j] = f.x;
    x[2 * j + 1] = f.y;
  }
  if (lo != nullptr) {
    const uint4 v = *reinterpret_cast<const uint4*>(lo + off);
    const __half2* l = reinterpret_cast<const __half2*>(&v);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 f = __half22float2(l[j]);
      x[2 * j] += f.x;
      x[2 * j + 1] += f.y;
    }
  }
}
__device__ __forceinline__ void t_load8f(const float* p, size_t off, float (&x)[8]) {
  ld_global_v8f(p + off, x);          // c8 fp32 voxel = one 32-byte sector (tensors are 32-byte aligned)
}
__device__ __forceinline__ void t_store8h(__half* hi, size_t off, const float (&x)[8]) {
  __align__(16) __half h[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) h[j] = __float2half_rn(x[j]);
  *reinterpret_cast<uint4*>(hi + off) = *reinterpret_cast<const uint4*>(h);
}
__device__ __forceinline__ void t_store8f(float* p, size_t off, const float (&y)[8]) {
  st_global_v8f(p + off, y);
}
__device__ __forceinline__ void t_finalize(const double* stats, size_t idx, double inv_count, float& mean, float& rstd) {
  const double s = stats[idx * 2], ss = stats[idx * 2 + 1];
  const double m = s * inv_count;
  double var = ss * inv_count - m * m;
  if (var < 0.0) var = 0.0;
  mean = static_cast<float>(m);
  rstd = rsqrtf(static_cast<float>(var) + 1e-5f);
}

// d act(t) / dt
__device__ __forceinline__ float act_grad(float t, int act) {
  switch (act) {
    case ACT_RELU: return t > 0.f ? 1.f : 0.f;
    case ACT_LRELU: return t > 0.f ? 1.f : 0.01f;
    case ACT_MISH: {
      if (t > 20.f) return 1.f;
      const float e = __expf(t);
      const float u = e * (e + 2.f);
      const float th = u / (u + 2.f);          // tanh(softplus(t))
      const float sg = e / (1.f + e);          // sigmoid(t)
      return th + t * (1.f - th * th) * sg;
    }
    case ACT_GELU: return 0.5f * (1.f + erff(t * 0.70710678118654752f)) + t * 0.3989422804014327f * __expf(-0.5f * t * t);
    default: return 1.f;
  }
}

// upstream gradient = sum of <=3 fp32 c8 tensors and/or one fp16 c8 tensor
struct GradSum {
  const float* f[3]; int f_cbt[3], f_cbo[3]; int nf;
  const __half* h; int h_cbt, h_cbo;
};
__device__ __forceinline__ void load_grad8(const GradSum& g, int n, int cb, long long vox, long long v, float (&y)[8]) {
#pragma unroll
  for (int j = 0; j < 8; ++j) y[j] = 0.f;
  for (int i = 0; i < g.nf; ++i) {
    float t[8];
    t_load8f(g.f[i], ((static_cast<size_t>(n) * g.f_cbt[i] + g.f_cbo[i] + cb) * vox + v) * 8, t);
#pragma unroll
    for (int j = 0; j < 8; ++j) y[j] += t[j];
  }
  if (g.h != nullptr) {
    float t[8];
    t_load8(g.h, nullptr, ((static_cast<size_t>(n) * g.h_cbt + g.h_cbo + cb) * vox + v) * 8, t);
#pragma unroll
    for (int j = 0; j < 8; ++j) y[j] += t[j];
  }
}
static int fill_grad(GradSum& g, int nf, const float* const* f, const int* cbt, const int* cbo, const void* h, int h_cbt,
                     int h_cbo) {
  if (nf < 0 || nf > 3) return 1;
  g.nf = nf;
  for (int i = 0; i < 3; ++i) {
    g.f[i] = i < nf ? f[i] : nullptr;
    g.f_cbt[i] = i < nf ? cbt[i] : 0;
    g.f_cbo[i] = i < nf ? cbo[i] : 0;
  }
  g.h = static_cast<const __half*>(h);
  g.h_cbt = h_cbt;
  g.h_cbo = h_cbo;
  return 0;
}

// block reduction of NV x 8 per-thread partial sums, one fp64 atomic per (value, channel): dst[(idx0 + j) * stride + slot0 + k]
template <int NV>
__device__ __forceinline__ void block_reduce64(const float (&acc)[NV][8], double* dst, size_t idx0, int stride, int slot0,
                                               int cvalid) {
  __shared__ float red[NV][8][8];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
  for (int k = 0; k < NV; ++k)
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float s = warp_sum(acc[k][j]);
      if (lane == 0) red[k][warp][j] = s;
    }
  __syncthreads();
  if (threadIdx.x < NV * 8) {
    const int k = threadIdx.x >> 3, j = threadIdx.x & 7;
    if (j < cvalid) {
      double t = 0.0;
      for (int w = 0; w < (blockDim.x >> 5); ++w) t += static_cast<double>(red[k][w][j]);
      atomicAdd(&dst[(idx0 + j) * stride + slot0 + k], t);
    }
  }
  __syncthreads();
}

// ------------------------------------------------------------------ backward of norm_act_kernel (ops.cu)
// forward:  xh = (x-mean)*rstd;  t = xh*gamma+beta;  a = act(t);  eh = (e-mean_e)*rstd_e;  y = act2(a + eh)
struct NormBwdParams {
  const float* raw_f32; const __half* raw_hi; const __half* raw_lo; int in_cb_total, in_cb_off;
  const double* stats; const float* gamma; const float* beta; int act;
  const __half* res_hi; const __half* res_lo; const float* res_raw; const double* res_stats;
  int res_cb_total, res_cb_off; int act_after_res;
  GradSum dy;
  double* bsum;          // [N][C][6] = {sum gt, sum gt*xh, sum gz, sum gz*eh, dgamma, dbeta}
  int phase;             // 0: reductions into bsum;  1: apply
  __half* dx_hi; float* dx_f32; int dx_cb_total, dx_cb_off;
  __half* dres_hi; float* dres_f32; int dres_cb_total, dres_cb_off;
  int C, ncb; long long vox; double inv_vox;
};
constexpr int NB_IT = 4;

// d act(t)/dt for 8 values with one switch (the per-element switch dominated the instruction count)
__device__ __forceinline__ void act_grad8(const float (&t)[8], int act, float (&d)[8]) {
  switch (act) {
    case ACT_RELU:
#pragma unroll
      for (int j = 0; j < 8; ++j) d[j] = t[j] > 0.f ? 1.f : 0.f;
      break;
    case ACT_LRELU:
#pragma unroll
      for (int j = 0; j < 8; ++j) d[j] = t[j] > 0.f ? 1.f : 0.01f;
      break;
    case ACT_MISH:
#pragma unroll
      for (int j = 0; j < 8; ++j) d[j] = act_grad(t[j], ACT_MISH);
      break;
    case ACT_GELU:
#pragma unroll
      for (int j = 0; j < 8; ++j) d[j] = act_grad(t[j], ACT_GELU);
      break;
    default:
#pragma unroll
      for (int j = 0; j < 8; ++j) d[j] = 1.f;
      break;
  }
}
__device__ __forceinline__ void act_fwd8(float (&y)[8], int act) {
  if (act == ACT_NONE) return;
#pragma unroll
  for (int j = 0; j < 8; ++j) y[j] = act_apply(y[j], act);
}

__global__ void __launch_bounds__(256, 2) norm_act_bwd_kernel(const NormBwdParams p) {
  const int cb = blockIdx.y % p.ncb, n = blockIdx.y / p.ncb;
  const double inv = p.inv_vox;
  __shared__ float s_mean[2][8], s_rstd[2][8], s_sum[4][8], s_g[8], s_b[8];
  if (threadIdx.x < 16) {
    const int which = threadIdx.x >> 3, j = threadIdx.x & 7, c = cb * 8 + j;
    const double* st = which ? p.res_stats : p.stats;
    float m = 0.f, r = 1.f;
    if (st != nullptr && c < p.C) t_finalize(st, static_cast<size_t>(n) * p.C + c, inv, m, r);
    if (c >= p.C) r = 0.f;            // padded channels: every product below becomes 0 (their upstream gradient is 0 too)
    s_mean[which][j] = m;
    s_rstd[which][j] = r;
    if (which == 0) {
      s_g[j] = (p.gamma && c < p.C) ? p.gamma[c] : 1.f;
      s_b[j] = (p.beta && c < p.C) ? p.beta[c] : 0.f;
    }
  }
  if (threadIdx.x >= 32 && threadIdx.x < 64) {
    const int k = (threadIdx.x - 32) >> 3, j = threadIdx.x & 7, c = cb * 8 + j;
    float s = 0.f;
    if (p.phase == 1 && c < p.C) s = static_cast<float>(p.bsum[(static_cast<size_t>(n) * p.C + c) * 6 + k] * inv);
    s_sum[k][j] = s;
  }
  __syncthreads();
  const bool has_res = p.res_hi || p.res_raw;
  const bool affine = p.gamma != nullptr;
  const bool res_norm = p.res_stats != nullptr;
  const bool main_norm = p.stats != nullptr;
  float acc[6][8];
#pragma unroll
  for (int k = 0; k < 6; ++k)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[k][j] = 0.f;
  const size_t in_base = (static_cast<size_t>(n) * p.in_cb_total + p.in_cb_off + cb) * p.vox;
#pragma unroll 1
  for (int it = 0; it < NB_IT; ++it) {
    const long long v = (blockIdx.x * static_cast<long long>(NB_IT) + it) * blockDim.x + threadIdx.x;
    if (v >= p.vox) break;
    float x[8], g[8], e[8], t[8], d[8];
    if (p.raw_f32) t_load8f(p.raw_f32, (in_base + v) * 8, x); else t_load8(p.raw_hi, p.raw_lo, (in_base + v) * 8, x);
    load_grad8(p.dy, n, cb, p.vox, v, g);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      x[j] = (x[j] - s_mean[0][j]) * s_rstd[0][j];                    // xh
      t[j] = affine ? fmaf(x[j], s_g[j], s_b[j]) : x[j];
    }
    if (has_res) {
      const size_t r_off = ((static_cast<size_t>(n) * p.res_cb_total + p.res_cb_off + cb) * p.vox + v) * 8;
      if (p.res_raw) t_load8f(p.res_raw, r_off, e); else t_load8(p.res_hi, p.res_lo, r_off, e);
      float z[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) { e[j] = (e[j] - s_mean[1][j]) * s_rstd[1][j]; z[j] = t[j]; }     // eh
      act_fwd8(z, p.act);
#pragma unroll
      for (int j = 0; j < 8; ++j) z[j] += e[j];
      act_grad8(z, p.act_after_res, d);
#pragma unroll
      for (int j = 0; j < 8; ++j) g[j] *= d[j];                        // gz
    }
    act_grad8(t, p.act, d);
    float gt[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) gt[j] = g[j] * d[j];                   // g0
    if (p.phase == 0) {
      if (affine) {
#pragma unroll
        for (int j = 0; j < 8; ++j) { acc[4][j] = fmaf(gt[j], x[j], acc[4][j]); acc[5][j] += gt[j]; gt[j] *= s_g[j]; }
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) { acc[0][j] += gt[j]; acc[1][j] = fmaf(gt[j], x[j], acc[1][j]); }
      if (res_norm) {
#pragma unroll
        for (int j = 0; j < 8; ++j) { acc[2][j] += g[j]; acc[3][j] = fmaf(g[j], e[j], acc[3][j]); }
      }
    } else {
      if (affine) {
#pragma unroll
        for (int j = 0; j < 8; ++j) gt[j] *= s_g[j];
      }
      float o[8];
      if (p.dx_hi || p.dx_f32) {
#pragma unroll
        for (int j = 0; j < 8; ++j)
          o[j] = main_norm ? s_rstd[0][j] * (gt[j] - s_sum[0][j] - x[j] * s_sum[1][j]) : gt[j];
        const size_t off = ((static_cast<size_t>(n) * p.dx_cb_total + p.dx_cb_off + cb) * p.vox + v) * 8;
        if (p.dx_hi) t_store8h(p.dx_hi, off, o); else t_store8f(p.dx_f32, off, o);
      }
      if (has_res && (p.dres_hi || p.dres_f32)) {
#pragma unroll
        for (int j = 0; j < 8; ++j)
          o[j] = res_norm ? s_rstd[1][j] * (g[j] - s_sum[2][j] - e[j] * s_sum[3][j]) : g[j];
        const size_t off = ((static_cast<size_t>(n) * p.dres_cb_total + p.dres_cb_off + cb) * p.vox + v) * 8;
        if (p.dres_hi) t_store8h(p.dres_hi, off, o); else t_store8f(p.dres_f32, off, o);
      }
    }
  }
  if (p.phase == 0) {
    const int cvalid = min(8, p.C - cb * 8);
    block_reduce64<6>(acc, p.bsum, static_cast<size_t>(n) * p.C + cb * 8, 6, 0, cvalid);
  }
}

// train-mode BatchNorm3d: replace every per-(n,c) slot by the batch mean of the slots, so that the
// per-instance consumers (finalize with 1/vox) see batch statistics; optionally update the running stats.
__global__ void batch_combine_kernel(double* a, int N, int C, int k, long long vox, float* running_mean,
                                     float* running_var, float momentum) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= C * k) return;
  // with running statistics (k == 2: {sum, sumsq}) ONE thread owns both slots of a channel: the running variance
  // needs the original sumsq values, which a neighbouring thread would otherwise be overwriting concurrently
  const bool pair = running_mean != nullptr && k == 2;
  if (pair && (i & 1)) return;
  const int nslots = pair ? 2 : 1;
  double m[2] = {0.0, 0.0};
  for (int q = 0; q < nslots; ++q) {
    double s = 0.0;
    for (int n = 0; n < N; ++n) s += a[static_cast<size_t>(n) * C * k + i + q];
    m[q] = s / N;
  }
  for (int q = 0; q < nslots; ++q)
    for (int n = 0; n < N; ++n) a[static_cast<size_t>(n) * C * k + i + q] = m[q];
  if (pair) {
    const int c = i >> 1;
    const double cnt = static_cast<double>(vox) * N;
    const double mean = m[0] / vox;
    double var = m[1] / vox - mean * mean;
    if (var < 0.0) var = 0.0;
    const double unbiased = cnt > 1.0 ? var * cnt / (cnt - 1.0) : var;
    running_mean[c] = static_cast<float>((1.0 - momentum) * running_mean[c] + momentum * mean);
    running_var[c] = static_cast<float>((1.0 - momentum) * running_var[c] + momentum * unbiased);
  }
}

// dgamma[c], dbeta[c] = sum_n bsum[n][c][4|5]
__global__ void affine_grad_kernel(const double* bsum, int N, int C, float* dgamma, float* dbeta, float scale) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  double g = 0.0, b = 0.0;
  for (int n = 0; n < N; ++n) {
    g += bsum[(static_cast<size_t>(n) * C + c) * 6 + 4];
    b += bsum[(static_cast<size_t>(n) * C + c) * 6 + 5];
  }
  dgamma[c] = static_cast<float>(g * scale);
  dbeta[c] = static_cast<float>(b * scale);
}

__global__ void grad_finalize_kernel(const double* acc, float* grad, long long n, float scale) {
  const long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (i < n) grad[i] = static_cast<float>(acc[i] * scale);
}

// ------------------------------------------------------------------ dW / dbias of 1x1x1 convs and 2x deconvs
// dW[co][ci] (+)= sum_{n,v} g[n][co][up(v,o)] * x[n][ci][v];  up = identity (pointwise) or the o-th of the 8 children
// of v in the 2x finer grid (transposed conv, o = blockIdx.z % 8).  Block: 16 co x 32 ci outputs, 8 voxel lanes.
struct SmallWgradParams {
  GradSum g; int g_C;
  const __half* x_hi; const __half* x_lo; int x_cb_total, x_cb_off, x_C;
  const __half* x_tok; long long tok_nstride; int tok_C;      // alternative x: [N][T][C] fp16 tokens
  int N; long long vox; int D, H, W;      // geometry of x
  int deconv;
  double* dw;          // fp64 accumulation target
  long long dw_co_stride, dw_ci_stride, dw_o_stride;
  double* dbias;       // [co] or null (pointwise only; added once, by ci-tile 0)
  int vsplit;
};
constexpr int SW_CO = 16, SW_CI = 32, SW_V = 64;

__global__ void __launch_bounds__(256) small_wgrad_kernel(const SmallWgradParams p) {
  __shared__ float gs[SW_V][SW_CO + 1];
  __shared__ float xs[SW_V][SW_CI + 1];
  __shared__ float red[8][SW_CO][SW_CI + 1];
  const int ci0 = blockIdx.x * SW_CI, co0 = blockIdx.y * SW_CO;
  const int o = p.deconv ? blockIdx.z % 8 : 0;
  const int split = p.deconv ? blockIdx.z / 8 : blockIdx.z;
  const int og = threadIdx.x & 31, vg = threadIdx.x >> 5;
  const int co4 = (og & 3) * 4, ci4 = (og >> 2) * 4;
  float acc[4][4];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) acc[a][b] = 0.f;
  float bsum = 0.f;
  const long long total = static_cast<long long>(p.N) * p.vox;
  const long long per = (total + p.vsplit - 1) / p.vsplit;
  const long long v_begin = split * per, v_end = min(total, v_begin + per);
  const long long gvox = p.deconv ? p.vox * 8 : p.vox;
  // (image, voxel) of the tile's first voxel: one 64-bit division per tile, then increments (the staging loop used
  // to spend more instructions on 64-bit div / mod per item than the FMA loop on arithmetic)
  int n_base = static_cast<int>(v_begin / p.vox);
  long long v_base = v_begin - n_base * p.vox;
  for (long long base = v_begin; base < v_end; base += SW_V) {
    // stage: 64 voxels x (2 co blocks + 4 ci blocks) of 8 channels
    for (int i = threadIdx.x; i < SW_V * 6; i += blockDim.x) {
      const int vi = i % SW_V, blk = i / SW_V;
      const long long gv = base + vi;
      float t[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) t[j] = 0.f;
      if (gv < v_end) {
        int n = n_base;
        long long v = v_base + vi;
        while (v >= p.vox) { v -= p.vox; ++n; }
        if (blk < 2) {
          const int cb = (co0 >> 3) + blk;
          if (cb * 8 < p.g_C) {
            long long vo = v;
            if (p.deconv) {
              const unsigned v32 = static_cast<unsigned>(v), hw = static_cast<unsigned>(p.W) * p.H;
              const int d = static_cast<int>(v32 / hw);
              const unsigned r2 = v32 - d * hw;
              const int h = static_cast<int>(r2 / p.W), w = static_cast<int>(r2 - h * p.W);
              vo = (static_cast<long long>(2 * d + (o >> 2)) * (2 * p.H) + 2 * h + ((o >> 1) & 1)) * (2 * p.W) + 2 * w + (o & 1);
            }
            load_grad8(p.g, n, cb, gvox, vo, t);
          }
        } else {
          const int cb = (ci0 >> 3) + blk - 2;
          if (cb * 8 < p.x_C) {
            if (p.x_tok) {
              const uint4 u = *reinterpret_cast<const uint4*>(p.x_tok + n * p.tok_nstride + v * p.tok_C + cb * 8);
              const __half2* hh = reinterpret_cast<const __half2*>(&u);
#pragma unroll
              for (int j = 0; j < 4; ++j) { const float2 f = __half22float2(hh[j]); t[2 * j] = f.x; t[2 * j + 1] = f.y; }
            } else {
              t_load8(p.x_hi, p.x_lo, ((static_cast<size_t>(n) * p.x_cb_total + p.x_cb_off + cb) * p.vox + v) * 8, t);
            }
          }
        }
      }
      if (blk < 2) {
#pragma unroll
        for (int j = 0; j < 8; ++j) gs[vi][blk * 8 + j] = t[j];
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) xs[vi][(blk - 2) * 8 + j] = t[j];
      }
    }
    __syncthreads();
#pragma unroll
    for (int s = 0; s < SW_V / 8; ++s) {
      const int vi = s * 8 + vg;
      float ga[4], xb[4];
#pragma unroll
      for (int a = 0; a < 4; ++a) ga[a] = gs[vi][co4 + a];
#pragma unroll
      for (int b = 0; b < 4; ++b) xb[b] = xs[vi][ci4 + b];
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) acc[a][b] = fmaf(ga[a], xb[b], acc[a][b]);
    }
    if (p.dbias && blockIdx.x == 0 && threadIdx.x < SW_CO) {
      for (int vi = 0; vi < SW_V; ++vi) bsum += gs[vi][threadIdx.x];
    }
    __syncthreads();
    v_base += SW_V;
    while (v_base >= p.vox) { v_base -= p.vox; ++n_base; }
  }
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) red[vg][co4 + a][ci4 + b] = acc[a][b];
  __syncthreads();
  for (int i = threadIdx.x; i < SW_CO * SW_CI; i += blockDim.x) {
    const int a = i / SW_CI, b = i % SW_CI;
    const int co = co0 + a, ci = ci0 + b;
    if (co < p.g_C && ci < p.x_C) {
      double t = 0.0;
#pragma unroll
      for (int w = 0; w < 8; ++w) t += static_cast<double>(red[w][a][b]);
      atomicAdd(&p.dw[co * p.dw_co_stride + ci * p.dw_ci_stride + o * p.dw_o_stride], t);
    }
  }
  if (p.dbias && blockIdx.x == 0 && threadIdx.x < SW_CO && co0 + threadIdx.x < p.g_C)
    atomicAdd(&p.dbias[co0 + threadIdx.x], static_cast<double>(bsum));
}

// ------------------------------------------------------------------ data gradient of ConvTranspose3d k2 s2
// dx[n][ci][v] = sum_{o,co} g[n][co][child(v,o)] * W[ci][co][o]; one thread = one voxel x 8 input channels.
struct DeconvBwdParams {
  GradSum g; int Co;
  const float* w;        // [8][Co][Ci] fp32 (re-laid-out on the host side)
  int Ci, N, D, H, W;
  float* dx_c8; int dx_cb_total, dx_cb_off;    // fp32 c8 output ...
  float* dx_tok; long long tok_nstride; int tok_C;   // ... or fp32 tokens [N][T][C]
};
__global__ void __launch_bounds__(128) deconv2x_bwd_data_kernel(const DeconvBwdParams p) {
  extern __shared__ float wsm[];            // [8][Co][8 ci of this block]
  const int cib = blockIdx.y, n = blockIdx.z;
  for (int i = threadIdx.x; i < 8 * p.Co * 8; i += blockDim.x) {
    const int j = i & 7, co = (i >> 3) % p.Co, o = i / (8 * p.Co);
    const int ci = cib * 8 + j;
    wsm[i] = ci < p.Ci ? p.w[(static_cast<size_t>(o) * p.Co + co) * p.Ci + ci] : 0.f;
  }
  __syncthreads();
  const long long vox = static_cast<long long>(p.D) * p.H * p.W;
  const long long v = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (v >= vox) return;
  const int w = static_cast<int>(v % p.W), h = static_cast<int>((v / p.W) % p.H);
  const int d = static_cast<int>(v / (static_cast<long long>(p.W) * p.H));
  float acc[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[j] = 0.f;
  const int ncbo = (p.Co + 7) / 8;
  for (int o = 0; o < 8; ++o) {
    const long long vo = (static_cast<long long>(2 * d + (o >> 2)) * (2 * p.H) + 2 * h + ((o >> 1) & 1)) * (2 * p.W) + 2 * w + (o & 1);
    for (int cb = 0; cb < ncbo; ++cb) {
      float g[8];
      load_grad8(p.g, n, cb, vox * 8, vo, g);
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        if (cb * 8 + q >= p.Co) break;
        const float4* wr = reinterpret_cast<const float4*>(&wsm[((o * p.Co) + cb * 8 + q) * 8]);
        const float4 w0 = wr[0], w1 = wr[1];
        acc[0] = fmaf(g[q], w0.x, acc[0]); acc[1] = fmaf(g[q], w0.y, acc[1]);
        acc[2] = fmaf(g[q], w0.z, acc[2]); acc[3] = fmaf(g[q], w0.w, acc[3]);
        acc[4] = fmaf(g[q], w1.x, acc[4]); acc[5] = fmaf(g[q], w1.y, acc[5]);
        acc[6] = fmaf(g[q], w1.z, acc[6]); acc[7] = fmaf(g[q], w1.w, acc[7]);
      }
    }
  }
  if (p.dx_c8) {
    t_store8f(p.dx_c8, ((static_cast<size_t>(n) * p.dx_cb_total + p.dx_cb_off + cib) * vox + v) * 8, acc);
  } else {
    float* dst = p.dx_tok + n * p.tok_nstride + v * p.tok_C + cib * 8;
    *reinterpret_cast<float4*>(dst) = make_float4(acc[0], acc[1], acc[2], acc[3]);
    *reinterpret_cast<float4*>(dst + 4) = make_float4(acc[4], acc[5], acc[6], acc[7]);
  }
}

// ------------------------------------------------------------------ dose heads (1x1x1 conv C -> 1, planar output)
// g: planar fp32 [N][vox];  dX (fp32 c8) = g * w[c];  dW[c] += sum g*x[c];  db += sum g
struct HeadBwdParams {
  const float* g; const __half* x_hi; const __half* x_lo; int x_cb_total, x_cb_off, C;
  const float* w; long long vox;
  float* dx; int dx_cb_total, dx_cb_off;
  double* dw; double* db;
};
__global__ void __launch_bounds__(256) head_bwd_kernel(const HeadBwdParams p) {
  const int ncb = (p.C + 7) / 8;
  const int cb = blockIdx.y % ncb, n = blockIdx.y / ncb;
  const long long v = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  float acc[2][8];
#pragma unroll
  for (int j = 0; j < 8; ++j) { acc[0][j] = 0.f; acc[1][j] = 0.f; }
  if (v < p.vox) {
    const float g = p.g[static_cast<size_t>(n) * p.vox + v];
    float x[8], o[8];
    t_load8(p.x_hi, p.x_lo, ((static_cast<size_t>(n) * p.x_cb_total + p.x_cb_off + cb) * p.vox + v) * 8, x);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int c = cb * 8 + j;
      o[j] = c < p.C ? g * __ldg(&p.w[c]) : 0.f;
      acc[0][j] = g * x[j];
      acc[1][j] = (cb == 0 && j == 0) ? g : 0.f;
    }
    t_store8f(p.dx, ((static_cast<size_t>(n) * p.dx_cb_total + p.dx_cb_off + cb) * p.vox + v) * 8, o);
  }
  __shared__ float red[2][8][8];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
  for (int k = 0; k < 2; ++k)
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float s = warp_sum(acc[k][j]);
      if (lane == 0) red[k][warp][j] = s;
    }
  __syncthreads();
  if (threadIdx.x < 16) {
    const int k = threadIdx.x >> 3, j = threadIdx.x & 7;
    double t = 0.0;
    for (int w = 0; w < 8; ++w) t += static_cast<double>(red[k][w][j]);
    if (k == 0 && cb * 8 + j < p.C) atomicAdd(&p.dw[cb * 8 + j], t);
    if (k == 1 && cb == 0 && j == 0) atomicAdd(p.db, t);
  }
}

// ------------------------------------------------------------------ GenLoss (loss.py:69-119): masked L1 at one scale
// target = trilinear(align_corners=True) resample of the GT dose to s^3, mask = nearest-exact resample (loss.py:57-67)
__device__ __forceinline__ float resample_gt(const float* gt, int S, int s, int d, int h, int w) {
  if (s == S) return gt[(static_cast<size_t>(d) * S + h) * S + w];
  const float sc = static_cast<float>(S - 1) / static_cast<float>(s - 1);
  const float fd = d * sc, fh = h * sc, fw = w * sc;
  const int d0 = static_cast<int>(fd), h0 = static_cast<int>(fh), w0 = static_cast<int>(fw);
  const int d1 = min(d0 + 1, S - 1), h1 = min(h0 + 1, S - 1), w1 = min(w0 + 1, S - 1);
  const float ld = fd - d0, lh = fh - h0, lw = fw - w0;
  auto at = [&](int a, int b, int c) { return gt[(static_cast<size_t>(a) * S + b) * S + c]; };
  // same association order as ATen's upsample_trilinear3d: w, then h, then d
  const float c00 = at(d0, h0, w0) * (1.f - lw) + at(d0, h0, w1) * lw;
  const float c01 = at(d0, h1, w0) * (1.f - lw) + at(d0, h1, w1) * lw;
  const float c10 = at(d1, h0, w0) * (1.f - lw) + at(d1, h0, w1) * lw;
  const float c11 = at(d1, h1, w0) * (1.f - lw) + at(d1, h1, w1) * lw;
  const float c0 = c00 * (1.f - lh) + c01 * lh, c1 = c10 * (1.f - lh) + c11 * lh;
  return c0 * (1.f - ld) + c1 * ld;
}
__device__ __forceinline__ float resample_mask(const float* m, int S, int s, int d, int h, int w) {
  if (s == S) return m[(static_cast<size_t>(d) * S + h) * S + w];
  const float sc = static_cast<float>(S) / static_cast<float>(s);
  const int ds = min(static_cast<int>(floorf((d + 0.5f) * sc)), S - 1);
  const int hs = min(static_cast<int>(floorf((h + 0.5f) * sc)), S - 1);
  const int ws = min(static_cast<int>(floorf((w + 0.5f) * sc)), S - 1);
  return m[(static_cast<size_t>(ds) * S + hs) * S + ws];
}
// gt: [N][2][S^3] (channel 0 dose, channel 1 possible-dose mask); pred: [N][1][s^3]
// phase 0: acc[0] += sum |p - t| over mask>0, acc[1] += count;  phase 1: dpred = coef * sign(p - t) / count
__global__ void __launch_bounds__(256) masked_l1_kernel(const float* pred, const float* gt, int N, int S, int s, double* acc,
                                                        int phase, float coef, float* dpred) {
  const long long vox = static_cast<long long>(s) * s * s;
  const long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  float a = 0.f, c = 0.f;
  if (i < vox * N) {
    const int n = static_cast<int>(i / vox);
    const long long v = i % vox;
    const int w = static_cast<int>(v % s), h = static_cast<int>((v / s) % s), d = static_cast<int>(v / (static_cast<long long>(s) * s));
    const size_t S3 = static_cast<size_t>(S) * S * S;
    const float* dose = gt + static_cast<size_t>(n) * 2 * S3;
    const float m = resample_mask(dose + S3, S, s, d, h, w);
    float gout = 0.f;
    if (m > 0.f) {
      const float diff = pred[i] - resample_gt(dose, S, s, d, h, w);
      a = fabsf(diff);
      c = 1.f;
      if (phase == 1) gout = (diff > 0.f ? coef : (diff < 0.f ? -coef : 0.f)) / static_cast<float>(acc[1]);
    }
    if (phase == 1) dpred[i] = gout;
  }
  if (phase == 0) {
    __shared__ float red[2][8];
    const float sa = warp_sum(a), sc = warp_sum(c);
    if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = sa; red[1][threadIdx.x >> 5] = sc; }
    __syncthreads();
    if (threadIdx.x < 2) {
      double t = 0.0;
      for (int w = 0; w < 8; ++w) t += static_cast<double>(red[threadIdx.x][w]);
      atomicAdd(&acc[threadIdx.x], t);
    }
  }
}
// loss = delta1 * L1(full) + delta2 * mean_i L1(scale i)   (loss.py:96-112)
__global__ void genloss_finalize_kernel(const double* acc, int n_scales, float delta1, float delta2, const double* acc_a,
                                        float weight_a, float* loss) {
  double l = delta1 * acc[0] / acc[1];
  if (acc_a != nullptr) l += weight_a * acc_a[0] / acc_a[1];          // casecade and not freez: + 0.5 * L1(pred_A) (loss.py:114-115)
  double ds = 0.0;
  for (int i = 1; i < n_scales; ++i) ds += acc[2 * i] / acc[2 * i + 1];
  if (n_scales > 1) l += delta2 * ds / (n_scales - 1);
  *loss = static_cast<float>(l);
}

// ------------------------------------------------------------------ adjoint of the 2x linear interpolation along one axis
// F.interpolate(scale_factor=2, mode='trilinear', align_corners=True) (c3d.py:36) is separable; its backward is three of
// these passes (D, H, W).  g [outer][2L][inner][8] fp32 -> out [outer][L][inner][8]:
//   out[i] = sum_o g[o] * ((i0(o) == i) * (1 - l(o)) + (i1(o) == i) * l(o)),  src(o) = o (L-1)/(2L-1), i0 = floor(src),
//   i1 = min(i0 + 1, L - 1), l = src - i0 — the forward kernel's float arithmetic, so this is its exact transpose.
// Gather form (deterministic): only o in [2i-2, 2i+3] can reach input i.
__global__ void __launch_bounds__(256) lerp2x_bwd_kernel(const float* __restrict__ g, long long outer, int L, long long inner,
                                                         float* __restrict__ out) {
  const long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (idx >= outer * L * inner) return;
  const long long in_ = idx % inner;
  const int i = static_cast<int>((idx / inner) % L);
  const long long o_ = idx / (inner * L);
  const int Lo = 2 * L;
  const float s = Lo > 1 ? static_cast<float>(L - 1) / static_cast<float>(Lo - 1) : 0.f;
  float acc[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[j] = 0.f;
  const int o_lo = max(0, 2 * i - 2), o_hi = min(Lo - 1, 2 * i + 3);
  for (int o = o_lo; o <= o_hi; ++o) {
    const float f = s * o;
    const int i0 = static_cast<int>(f), i1 = min(i0 + 1, L - 1);
    const float l = f - i0;
    const float wgt = (i0 == i ? 1.f - l : 0.f) + (i1 == i ? l : 0.f);
    if (wgt != 0.f) {
      float x[8];
      ld_global_v8f(g + ((o_ * Lo + o) * inner + in_) * 8, x);
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] = fmaf(wgt, x[j], acc[j]);
    }
  }
  st_global_v8f(out + idx * 8, acc);
}

// ------------------------------------------------------------------ fused AdamW over a flat parameter buffer
// torch.optim.AdamW semantics (decoupled decay first, then the Adam update); grads carry 1/inv_scale loss scaling.
// found_inf (device int): set by dp_grad_check; when non-zero the step is skipped (dynamic loss scaling contract).
__global__ void __launch_bounds__(256) adamw_kernel(float* p, const float* g, float* m, float* v, long long n, float lr,
                                                    float b1, float b2, float eps, float wd, float bc1, float bc2_sqrt,
                                                    float inv_scale, const int* found_inf) {
  if (found_inf != nullptr && *found_inf != 0) return;
  const long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (i >= n) return;
  const float gi = g[i] * inv_scale;
  const float mi = b1 * m[i] + (1.f - b1) * gi;
  const float vi = b2 * v[i] + (1.f - b2) * gi * gi;
  m[i] = mi;
  v[i] = vi;
  float pi = p[i] * (1.f - lr * wd);
  const float denom = sqrtf(vi) / bc2_sqrt + eps;
  pi -= (lr / bc1) * (mi / denom);
  p[i] = pi;
}
// Same update with the optimizer step counter kept ON THE DEVICE (state[0] = completed steps, state[1] = consecutive
// skipped steps, state[2] = total skipped steps): a step skipped for non-finite gradients must not advance the bias
// correction, and the host never has to read found_inf back to know the step number.
__global__ void __launch_bounds__(256) adamw_dev_kernel(float* p, const float* g, float* m, float* v, long long n, float lr,
                                                        float b1, float b2, float eps, float wd, float inv_scale,
                                                        const int* found_inf, const int* state) {
  if (found_inf != nullptr && *found_inf != 0) return;
  __shared__ float s_bc[2];
  if (threadIdx.x == 0) {
    const float step = static_cast<float>(state[0] + 1);
    s_bc[0] = 1.f - powf(b1, step);
    s_bc[1] = sqrtf(1.f - powf(b2, step));
  }
  __syncthreads();
  const long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (i >= n) return;
  const float gi = g[i] * inv_scale;
  const float mi = b1 * m[i] + (1.f - b1) * gi;
  const float vi = b2 * v[i] + (1.f - b2) * gi * gi;
  m[i] = mi;
  v[i] = vi;
  float pi = p[i] * (1.f - lr * wd);
  const float denom = sqrtf(vi) / s_bc[1] + eps;
  pi -= (lr / s_bc[0]) * (mi / denom);
  p[i] = pi;
}
__global__ void adam_state_kernel(const int* found_inf, int* state) {
  if (found_inf != nullptr && *found_inf != 0) { state[1] += 1; state[2] += 1; }
  else { state[0] += 1; state[1] = 0; }
}
__global__ void __launch_bounds__(256) grad_check_kernel(const float* g, long long n, int* found_inf) {
  const long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  bool bad = false;
  if (i < n) { const float x = g[i]; bad = !(fabsf(x) <= 3.0e38f); }
  if (__syncthreads_or(bad) && threadIdx.x == 0) *found_inf = 1;
}


// ------------------------------------------------------------------ conv weight gradient, CUDA-core version
// dW[co][ci][kd][kh][kw] = sum_{n,d,h,w} g[n][co][d][h][w] * x[n][ci][d+kd-p][h+kh-p][w+kw-p]
// Block = one (kd, kh) pair x 16 input channels x 16 output channels x one slice of the (n,d,h) rows; every warp
// owns one row at a time (x row and g row staged in its private shared-memory slice), lane = (ci, 8 output
// channels), K x 8 accumulators per lane.  Partials go to ws[split][...] (reduced by dp_splitk_reduce).
struct ConvWgradParams {
  const __half* x; int x_cb_total; uint8_t chunk_cb[64]; int16_t chunk_ci0[64]; uint8_t chunk_nci[64]; int n_chunks;
  const __half* g; int g_cb_total, g_cb_off; int Co, Ci;
  int N, D, H, W, dil;
  float* ws; long long wsize; int splits;
};
template <int K>
__global__ void __launch_bounds__(256) conv_wgrad_kernel(const ConvWgradParams p) {
  extern __shared__ float sm[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int P = (K / 2) * p.dil;
  const int xw = p.W + 2 * P;                       // staged x row width
  float* xs = sm + static_cast<size_t>(warp) * (static_cast<size_t>(xw) + p.W) * 16;     // [xw][16 ci]
  float* gs = xs + static_cast<size_t>(xw) * 16;                                            // [W][16 co]
  const int kd = blockIdx.x / K, kh = blockIdx.x % K;
  const int n_cot = p.Co / 16;
  const int chunk = blockIdx.y / n_cot, cot = blockIdx.y % n_cot;
  const int split = blockIdx.z;
  const int ci = lane & 15, cog = lane >> 4;
  float acc[K][8];
#pragma unroll
  for (int a = 0; a < K; ++a)
#pragma unroll
    for (int b = 0; b < 8; ++b) acc[a][b] = 0.f;
  const long long rows = static_cast<long long>(p.N) * p.D * p.H;
  const long long per = (rows + p.splits - 1) / p.splits;
  const long long r_end = min(rows, (split + 1) * per);
  const size_t vox = static_cast<size_t>(p.D) * p.H * p.W;
  const int xcb = p.chunk_cb[chunk];
  for (long long r = split * per + warp; r < r_end; r += 8) {
    const int h = static_cast<int>(r % p.H), d = static_cast<int>((r / p.H) % p.D), n = static_cast<int>(r / (static_cast<long long>(p.H) * p.D));
    const int dx = d + kd * p.dil - P, hx = h + kh * p.dil - P;
    if (dx < 0 || dx >= p.D || hx < 0 || hx >= p.H) continue;      // warp-uniform
    // stage the x row (zero padded) and the g row as fp32
    const __half* xrow = p.x + ((static_cast<size_t>(n) * p.x_cb_total + xcb) * vox + (static_cast<size_t>(dx) * p.H + hx) * p.W) * 8;
    for (int i = lane; i < xw * 2; i += 32) {
      const int blk = i / xw, wv = i % xw - P;
      float t[8];
      if (wv >= 0 && wv < p.W) t_load8(xrow + static_cast<size_t>(blk) * vox * 8, nullptr, static_cast<size_t>(wv) * 8, t);
      else {
#pragma unroll
        for (int j = 0; j < 8; ++j) t[j] = 0.f;
      }
      float* dst = xs + (wv + P) * 16 + blk * 8;
      *reinterpret_cast<float4*>(dst) = make_float4(t[0], t[1], t[2], t[3]);
      *reinterpret_cast<float4*>(dst + 4) = make_float4(t[4], t[5], t[6], t[7]);
    }
    const __half* grow = p.g + ((static_cast<size_t>(n) * p.g_cb_total + p.g_cb_off + cot * 2) * vox + (static_cast<size_t>(d) * p.H + h) * p.W) * 8;
    for (int i = lane; i < p.W * 2; i += 32) {
      const int blk = i / p.W, wv = i % p.W;
      float t[8];
      t_load8(grow + static_cast<size_t>(blk) * vox * 8, nullptr, static_cast<size_t>(wv) * 8, t);
      float* dst = gs + wv * 16 + blk * 8;
      *reinterpret_cast<float4*>(dst) = make_float4(t[0], t[1], t[2], t[3]);
      *reinterpret_cast<float4*>(dst + 4) = make_float4(t[4], t[5], t[6], t[7]);
    }
    __syncwarp();
    for (int w = 0; w < p.W; ++w) {
      const float4 g0 = *reinterpret_cast<const float4*>(gs + w * 16 + cog * 8);
      const float4 g1 = *reinterpret_cast<const float4*>(gs + w * 16 + cog * 8 + 4);
#pragma unroll
      for (int a = 0; a < K; ++a) {
        const float xv = xs[(w + a * p.dil) * 16 + ci];
        acc[a][0] = fmaf(xv, g0.x, acc[a][0]); acc[a][1] = fmaf(xv, g0.y, acc[a][1]);
        acc[a][2] = fmaf(xv, g0.z, acc[a][2]); acc[a][3] = fmaf(xv, g0.w, acc[a][3]);
        acc[a][4] = fmaf(xv, g1.x, acc[a][4]); acc[a][5] = fmaf(xv, g1.y, acc[a][5]);
        acc[a][6] = fmaf(xv, g1.z, acc[a][6]); acc[a][7] = fmaf(xv, g1.w, acc[a][7]);
      }
    }
    __syncwarp();
  }
  // reduce the 8 warps through shared memory (reusing the staging area), then write the partial
  __syncthreads();
  float* red = sm;                                  // [8 warps][32 lanes][K*8]
#pragma unroll
  for (int a = 0; a < K; ++a)
#pragma unroll
    for (int b = 0; b < 8; ++b) red[(warp * 32 + lane) * (K * 8) + a * 8 + b] = acc[a][b];
  __syncthreads();
  float* out = p.ws + static_cast<size_t>(split) * p.wsize;
  const int nci = p.chunk_nci[chunk], ci0 = p.chunk_ci0[chunk];
  for (int i = threadIdx.x; i < 32 * K * 8; i += blockDim.x) {
    const int ln = i / (K * 8), a = (i / 8) % K, b = i % 8;
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) t += red[(w * 32 + ln) * (K * 8) + a * 8 + b];
    const int c_in = ln & 15, co = cot * 16 + (ln >> 4) * 8 + b;
    if (c_in < nci)
      out[(((static_cast<size_t>(co) * p.Ci + ci0 + c_in) * K + kd) * K + kh) * K + a] = t;
  }
}

// ------------------------------------------------------------------ token-side (ViT) backward helpers
// LayerNorm backward over rows of x [rows][cols]; dx = add + rstd*(g - mean(g) - xh*mean(g*xh)), g = dy*gamma
constexpr int LN_ROWS = 4;
__global__ void __launch_bounds__(256) layernorm_bwd_kernel(const float* x, const float* gamma, const float* dy,
                                                            const float* add, int rows, int cols, float* dx,
                                                            double* dgamma, double* dbeta) {
  __shared__ float red[4][8];
  __shared__ float bc[4];
  float pg[4], pb[4];            // per-thread column partials (cols <= 1024)
#pragma unroll
  for (int q = 0; q < 4; ++q) { pg[q] = 0.f; pb[q] = 0.f; }
  for (int rr = 0; rr < LN_ROWS; ++rr) {
    const int r = blockIdx.x * LN_ROWS + rr;
    if (r >= rows) break;
    const float* xr = x + static_cast<size_t>(r) * cols;
    const float* gr = dy + static_cast<size_t>(r) * cols;
    float s1 = 0.f, s2 = 0.f;
    for (int c = threadIdx.x; c < cols; c += 256) { const float v = xr[c]; s1 += v; s2 = fmaf(v, v, s2); }
    s1 = warp_sum(s1); s2 = warp_sum(s2);
    if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = s1; red[1][threadIdx.x >> 5] = s2; }
    __syncthreads();
    if (threadIdx.x == 0) {
      float a = 0.f, b = 0.f;
      for (int w = 0; w < 8; ++w) { a += red[0][w]; b += red[1][w]; }
      const float mean = a / cols;
      bc[0] = mean;
      bc[1] = rsqrtf(fmaxf(b / cols - mean * mean, 0.f) + 1e-5f);
    }
    __syncthreads();
    const float mean = bc[0], rstd = bc[1];
    float t1 = 0.f, t2 = 0.f;
    for (int c = threadIdx.x; c < cols; c += 256) {
      const float xh = (xr[c] - mean) * rstd, g = gr[c] * gamma[c];
      t1 += g; t2 = fmaf(g, xh, t2);
    }
    t1 = warp_sum(t1); t2 = warp_sum(t2);
    if ((threadIdx.x & 31) == 0) { red[2][threadIdx.x >> 5] = t1; red[3][threadIdx.x >> 5] = t2; }
    __syncthreads();
    if (threadIdx.x == 0) {
      float a = 0.f, b = 0.f;
      for (int w = 0; w < 8; ++w) { a += red[2][w]; b += red[3][w]; }
      bc[2] = a / cols; bc[3] = b / cols;
    }
    __syncthreads();
    const float m1 = bc[2], m2 = bc[3];
    int q = 0;
    for (int c = threadIdx.x; c < cols; c += 256, ++q) {
      const float xh = (xr[c] - mean) * rstd, go = gr[c];
      const float v = rstd * (go * gamma[c] - m1 - xh * m2);
      dx[static_cast<size_t>(r) * cols + c] = (add ? add[static_cast<size_t>(r) * cols + c] : 0.f) + v;
      pg[q] = fmaf(go, xh, pg[q]);
      pb[q] += go;
    }
    __syncthreads();
  }
  int q = 0;
  for (int c = threadIdx.x; c < cols; c += 256, ++q) {
    atomicAdd(&dgamma[c], static_cast<double>(pg[q]));
    atomicAdd(&dbeta[c], static_cast<double>(pb[q]));
  }
}

// softmax backward: dS = P * (dP - sum_j dP*P), rows of length cols; P fp16 [rows][ld_p], dP fp32 [rows][ld_dp]
__global__ void __launch_bounds__(256) softmax_bwd_kernel(const __half* P, int ld_p, const float* dP, int ld_dp, int rows,
                                                          int cols, __half* dS, int ld_ds) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int r = blockIdx.x * 8 + warp;
  if (r >= rows) return;
  const __half* pr = P + static_cast<size_t>(r) * ld_p;
  const float* gr = dP + static_cast<size_t>(r) * ld_dp;
  float s = 0.f;
  for (int c = lane; c < cols; c += 32) s = fmaf(__half2float(pr[c]), gr[c], s);
  s = warp_sum(s);
  for (int c = lane; c < cols; c += 32) dS[static_cast<size_t>(r) * ld_ds + c] = __float2half_rn(__half2float(pr[c]) * (gr[c] - s));
}

// y16 = act(u)   /   du = dh * act'(u)   (elementwise over fp32 u)
__global__ void __launch_bounds__(256) act_fwd_kernel(const float* u, long long n, int act, __half* y) {
  const long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (i < n) y[i] = __float2half_rn(act_apply(u[i], act));
}
__global__ void __launch_bounds__(256) act_bwd_kernel(const float* u, const float* dh, long long n, int act, float* du32,
                                                      __half* du16) {
  const long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (i >= n) return;
  const float v = dh[i] * act_grad(u[i], act);
  if (du32) du32[i] = v;
  if (du16) du16[i] = __float2half_rn(v);
}

// dst[b][c][r] = scale * src[b][r][c]  (fp32 or fp16 source, fp16 destination), 32x32 tiles through shared memory
__global__ void __launch_bounds__(256) transpose_kernel(const void* src, int src_f32, long long src_bs, int ld_src, int R,
                                                        int C, __half* dst, long long dst_bs, int ld_dst, float scale) {
  __shared__ float tile[32][33];
  const int b = blockIdx.z;
  const int r0 = blockIdx.y * 32, c0 = blockIdx.x * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int i = ty; i < 32; i += 8) {
    const int r = r0 + i, c = c0 + tx;
    float v = 0.f;
    if (r < R && c < C) {
      const size_t off = static_cast<size_t>(b) * src_bs + static_cast<size_t>(r) * ld_src + c;
      v = src_f32 ? static_cast<const float*>(src)[off] : __half2float(static_cast<const __half*>(src)[off]);
    }
    tile[i][tx] = v * scale;
  }
  __syncthreads();
  for (int i = ty; i < 32; i += 8) {
    const int c = c0 + i, r = r0 + tx;
    if (r < R && c < C) dst[static_cast<size_t>(b) * dst_bs + static_cast<size_t>(c) * ld_dst + r] = __float2half_rn(tile[tx][i]);
  }
}

// head split / merge between [B][T][ld] rows (column offset col0 + head*hd + d) and [B*heads][T][hd]
__global__ void __launch_bounds__(256) heads_kernel(const void* src, int src_f32, void* dst, int dst_f32, int B, int T, int heads,
                                                    int hd, int ld, int col0, int merge, float scale) {
  const long long total = static_cast<long long>(B) * heads * T * hd;
  const long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (i >= total) return;
  const int d = static_cast<int>(i % hd);
  const int t = static_cast<int>((i / hd) % T);
  const int h = static_cast<int>((i / (static_cast<long long>(hd) * T)) % heads);
  const int b = static_cast<int>(i / (static_cast<long long>(hd) * T * heads));
  const size_t row_idx = (static_cast<size_t>(b) * T + t) * ld + col0 + h * hd + d;
  const size_t head_idx = static_cast<size_t>(i);
  const size_t si = merge ? head_idx : row_idx, di = merge ? row_idx : head_idx;
  const float v = (src_f32 ? static_cast<const float*>(src)[si] : __half2float(static_cast<const __half*>(src)[si])) * scale;
  if (dst_f32) static_cast<float*>(dst)[di] = v; else static_cast<__half*>(dst)[di] = __float2half_rn(v);
}

// out64[c] += sum_r a[r][c]   (bias gradients; also the position-embedding gradient with rows = batch)
__global__ void __launch_bounds__(256) colsum_kernel(const float* a, int rows, long long cols, double* out) {
  const long long c = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (c >= cols) return;
  const int r0 = blockIdx.y * 64, r1 = min(rows, r0 + 64);
  float s = 0.f;
  for (int r = r0; r < r1; ++r) s += a[static_cast<size_t>(r) * cols + c];
  atomicAdd(&out[c], static_cast<double>(s));
}

// y = a + b (fp32), optional fp16 copy
__global__ void __launch_bounds__(256) add_kernel(const float* a, const float* b, long long n, float* y32, __half* y16) {
  const long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (i >= n) return;
  const float v = a[i] + (b ? b[i] : 0.f);
  if (y32) y32[i] = v;
  if (y16) y16[i] = __float2half_rn(v);
}


// ------------------------------------------------------------------ conv weight packing for the tensor-core kernels
// fp32 [Co][Ci][k][k][k] (optionally read transposed + flipped: the dgrad weights) -> the fp16 operand layouts of
// conv_tc.cu ([kd][chunk][kh][kw][2][Co][8]) and conv_stack.cu ([rot][chunk][kh][kw][2][slot][Co][8], slot s of
// rotation r holding depth tap k-1-j, j = (s-r) mod (k+1), zeros for j == k).  One thread per output element.
struct PackParams {
  const float* w; int Co, Ci, k; long long s_co, s_ci; int flip;
  int16_t chunk_ci0[192]; uint8_t chunk_nci[192]; int n_chunks;
  int stacked; __half* out; long long total;
};
__global__ void __launch_bounds__(256) pack_conv_weight_kernel(const PackParams p) {
  const long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (i >= p.total) return;
  const int k = p.k;
  long long r = i;
  const int e = static_cast<int>(r % 8); r /= 8;
  const int co = static_cast<int>(r % p.Co); r /= p.Co;
  int kd;
  bool zero = false;
  int khalf, kw, kh, chunk;
  if (p.stacked) {
    const int G = k + 1;
    const int slot = static_cast<int>(r % G); r /= G;
    khalf = static_cast<int>(r % 2); r /= 2;
    kw = static_cast<int>(r % k); r /= k;
    kh = static_cast<int>(r % k); r /= k;
    chunk = static_cast<int>(r % p.n_chunks); r /= p.n_chunks;
    const int rot = static_cast<int>(r);
    const int j = (slot - rot + G) % G;
    zero = j == k;
    kd = k - 1 - j;
  } else {
    khalf = static_cast<int>(r % 2); r /= 2;
    kw = static_cast<int>(r % k); r /= k;
    kh = static_cast<int>(r % k); r /= k;
    chunk = static_cast<int>(r % p.n_chunks); r /= p.n_chunks;
    kd = static_cast<int>(r);
  }
  const int cl = khalf * 8 + e;
  float v = 0.f;
  if (!zero && cl < p.chunk_nci[chunk]) {
    const int ci = p.chunk_ci0[chunk] + cl;
    const int a = p.flip ? k - 1 - kd : kd, b = p.flip ? k - 1 - kh : kh, c = p.flip ? k - 1 - kw : kw;
    v = p.w[co * p.s_co + ci * p.s_ci + (static_cast<long long>(a) * k + b) * k + c];
  }
  p.out[i] = __float2half_rn(v);
}

// dst = fp16(src) (optionally transposed [R][C] -> [C][R]) for the GEMM weight operands
__global__ void __launch_bounds__(256) cast_f16_kernel(const float* src, long long n, __half* dst) {
  const long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (i < n) dst[i] = __float2half_rn(src[i]);
}


// ------------------------------------------------------------------ DiceCELoss (monai 0.7.0, to_onehot_y, softmax)
// logits: c8 fp32, C <= 8 classes in channel block 0; label: fp32 class index per voxel [N][vox].
// acc: double[N][8][3] = {sum p*t, sum t, sum p} followed by one double = sum of -log p[label].
// phase 0 accumulates acc; phase 1 writes coef * dLoss/dlogit as fp16 c8 (block 0), where
// Loss = mean_{n,c} (1 - (2 I + eps) / (G + P + eps)) + mean_voxels CE.
__global__ void __launch_bounds__(256) dice_ce_kernel(const float* logits, int cb_total, const float* label, int N, int C,
                                                      long long vox, double* acc, int phase, float coef, __half* g16,
                                                      int g_cb_total) {
  const int n = blockIdx.y;
  const long long v = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  const bool valid = v < vox;
  float p[8], t[8];
  float ce = 0.f;
#pragma unroll
  for (int j = 0; j < 8; ++j) { p[j] = 0.f; t[j] = 0.f; }
  if (valid) {
    float z[8];
    t_load8f(logits, (static_cast<size_t>(n) * cb_total * vox + v) * 8, z);
    const int lab = static_cast<int>(label[static_cast<size_t>(n) * vox + v]);
    float m = -3.0e38f;
#pragma unroll
    for (int j = 0; j < 8; ++j) if (j < C) m = fmaxf(m, z[j]);
    float ssum = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) { p[j] = j < C ? __expf(z[j] - m) : 0.f; ssum += p[j]; }
    const float inv = 1.f / ssum;
#pragma unroll
    for (int j = 0; j < 8; ++j) { p[j] *= inv; t[j] = (j == lab) ? 1.f : 0.f; }
    float zl = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) if (j == lab) zl = z[j];
    ce = -(zl - m - __logf(ssum));
  }
  if (phase == 0) {
    float a[3][8];
#pragma unroll
    for (int j = 0; j < 8; ++j) { a[0][j] = p[j] * t[j]; a[1][j] = t[j]; a[2][j] = p[j]; }
    block_reduce64<3>(a, acc, static_cast<size_t>(n) * 8, 3, 0, C);
    __shared__ float red[8];
    const float s = warp_sum(ce);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
      double tsum = 0.0;
      for (int w = 0; w < 8; ++w) tsum += static_cast<double>(red[w]);
      atomicAdd(&acc[static_cast<size_t>(N) * 24], tsum);
    }
    return;
  }
  if (!valid) return;
  float a[8], dot = 0.f;
  const float w_dice = 1.f / (static_cast<float>(N) * C);
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    a[j] = 0.f;
    if (j < C) {
      const double* q = acc + (static_cast<size_t>(n) * 8 + j) * 3;
      const float I = static_cast<float>(q[0]), D = static_cast<float>(q[1] + q[2]) + 1e-5f;
      a[j] = -w_dice * (2.f * t[j] * D - (2.f * I + 1e-5f)) / (D * D);
      dot = fmaf(a[j], p[j], dot);
    }
  }
  const float w_ce = 1.f / (static_cast<float>(N) * static_cast<float>(vox));
  float g[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) g[j] = j < C ? coef * (p[j] * (a[j] - dot) + w_ce * (p[j] - t[j])) : 0.f;
  t_store8h(g16, (static_cast<size_t>(n) * g_cb_total * vox + v) * 8, g);
}
__global__ void dice_ce_finalize_kernel(const double* acc, int N, int C, long long vox, float* loss) {
  double dice = 0.0;
  for (int n = 0; n < N; ++n)
    for (int c = 0; c < C; ++c) {
      const double* q = acc + (static_cast<size_t>(n) * 8 + c) * 3;
      dice += 1.0 - (2.0 * q[0] + 1e-5) / (q[1] + q[2] + 1e-5);
    }
  *loss = static_cast<float>(dice / (static_cast<double>(N) * C) + acc[static_cast<size_t>(N) * 24] / (static_cast<double>(N) * vox));
}

static inline unsigned nblk(long long n, int threads) { return static_cast<unsigned>((n + threads - 1) / threads); }

}  // namespace dp

using namespace dp;

extern "C" int dp_norm_act_bwd(const float* raw_f32, const void* raw_hi, const void* raw_lo, int in_cb_total, int in_cb_off,
                               const double* stats, const float* gamma, const float* beta, int act, const void* res_hi,
                               const void* res_lo, const float* res_raw, const double* res_stats, int res_cb_total,
                               int res_cb_off, int act_after_res, int n_dy, const float* const* dy_f32, const int* dy_cb_total,
                               const int* dy_cb_off, const void* dy_f16, int dy16_cb_total, int dy16_cb_off, double* bsum,
                               int phase, void* dx_hi, float* dx_f32, int dx_cb_total, int dx_cb_off, void* dres_hi,
                               float* dres_f32, int dres_cb_total, int dres_cb_off, int N, int C, long long vox,
                               cudaStream_t stream) {
  DP_REQUIRE((raw_f32 != nullptr) != (raw_hi != nullptr), "norm_act_bwd: exactly one forward source");
  DP_REQUIRE(bsum != nullptr && (phase == 0 || phase == 1), "norm_act_bwd: bsum / phase");
  NormBwdParams p{};
  p.raw_f32 = raw_f32; p.raw_hi = static_cast<const __half*>(raw_hi); p.raw_lo = static_cast<const __half*>(raw_lo);
  p.in_cb_total = in_cb_total; p.in_cb_off = in_cb_off; p.stats = stats; p.gamma = gamma; p.beta = beta; p.act = act;
  p.res_hi = static_cast<const __half*>(res_hi); p.res_lo = static_cast<const __half*>(res_lo); p.res_raw = res_raw;
  p.res_stats = res_stats; p.res_cb_total = res_cb_total; p.res_cb_off = res_cb_off; p.act_after_res = act_after_res;
  DP_REQUIRE(fill_grad(p.dy, n_dy, dy_f32, dy_cb_total, dy_cb_off, dy_f16, dy16_cb_total, dy16_cb_off) == 0,
             "norm_act_bwd: at most 3 fp32 gradient sources");
  p.bsum = bsum; p.phase = phase;
  p.dx_hi = static_cast<__half*>(dx_hi); p.dx_f32 = dx_f32; p.dx_cb_total = dx_cb_total; p.dx_cb_off = dx_cb_off;
  p.dres_hi = static_cast<__half*>(dres_hi); p.dres_f32 = dres_f32; p.dres_cb_total = dres_cb_total; p.dres_cb_off = dres_cb_off;
  p.C = C; p.ncb = (C + 7) / 8; p.vox = vox; p.inv_vox = 1.0 / static_cast<double>(vox);
  dim3 grid(nblk(vox, 256 * NB_IT), static_cast<unsigned>(N * p.ncb));
  norm_act_bwd_kernel<<<grid, 256, 0, stream>>>(p);
  return check_cuda(cudaGetLastError(), "norm_act_bwd");
}

extern "C" int dp_batch_combine(double* slots, int N, int C, int k, long long vox, float* running_mean, float* running_var,
                                float momentum, cudaStream_t stream) {
  batch_combine_kernel<<<nblk(C * k, 128), 128, 0, stream>>>(slots, N, C, k, vox, running_mean, running_var, momentum);
  return check_cuda(cudaGetLastError(), "batch_combine");
}

extern "C" int dp_affine_grad(const double* bsum, int N, int C, float* dgamma, float* dbeta, float scale,
                              cudaStream_t stream) {
  affine_grad_kernel<<<nblk(C, 128), 128, 0, stream>>>(bsum, N, C, dgamma, dbeta, scale);
  return check_cuda(cudaGetLastError(), "affine_grad");
}

extern "C" int dp_grad_finalize(const double* acc, float* grad, long long n, float scale, cudaStream_t stream) {
  grad_finalize_kernel<<<nblk(n, 256), 256, 0, stream>>>(acc, grad, n, scale);
  return check_cuda(cudaGetLastError(), "grad_finalize");
}

extern "C" int dp_small_wgrad(int n_g, const float* const* g_f32, const int* g_cb_total, const int* g_cb_off,
                              const void* g_f16, int g16_cb_total, int g16_cb_off, int g_C, const void* x_hi,
                              const void* x_lo, int x_cb_total, int x_cb_off, int x_C, const void* x_tok, int N, int D,
                              int H, int W, int deconv, double* dw, long long dw_co_stride, long long dw_ci_stride,
                              long long dw_o_stride, double* dbias, cudaStream_t stream) {
  SmallWgradParams p{};
  DP_REQUIRE(fill_grad(p.g, n_g, g_f32, g_cb_total, g_cb_off, g_f16, g16_cb_total, g16_cb_off) == 0,
             "small_wgrad: at most 3 fp32 gradient sources");
  p.g_C = g_C;
  p.x_hi = static_cast<const __half*>(x_hi); p.x_lo = static_cast<const __half*>(x_lo);
  p.x_cb_total = x_cb_total; p.x_cb_off = x_cb_off; p.x_C = x_C;
  p.x_tok = static_cast<const __half*>(x_tok);
  p.N = N; p.D = D; p.H = H; p.W = W; p.vox = static_cast<long long>(D) * H * W;
  p.tok_C = x_C; p.tok_nstride = p.vox * x_C;
  DP_REQUIRE(x_tok == nullptr || x_C % 8 == 0, "small_wgrad: token sources need C % 8 == 0");
  p.deconv = deconv; p.dw = dw; p.dw_co_stride = dw_co_stride; p.dw_ci_stride = dw_ci_stride; p.dw_o_stride = dw_o_stride;
  p.dbias = dbias;
  const long long total = p.vox * N;
  const int tiles = ((x_C + SW_CI - 1) / SW_CI) * ((g_C + SW_CO - 1) / SW_CO) * (deconv ? 8 : 1);
  int vsplit = static_cast<int>(std::min<long long>((total + SW_V * 4 - 1) / (SW_V * 4), std::max(1, (8 * sm_count()) / tiles)));
  p.vsplit = std::max(1, vsplit);
  dim3 grid((x_C + SW_CI - 1) / SW_CI, (g_C + SW_CO - 1) / SW_CO, p.vsplit * (deconv ? 8 : 1));
  small_wgrad_kernel<<<grid, 256, 0, stream>>>(p);
  return check_cuda(cudaGetLastError(), "small_wgrad");
}

extern "C" int dp_deconv2x_bwd_data(int n_g, const float* const* g_f32, const int* g_cb_total, const int* g_cb_off,
                                    const void* g_f16, int g16_cb_total, int g16_cb_off, int Co, const float* w, int Ci,
                                    int N, int D, int H, int W, float* dx_c8, int dx_cb_total, int dx_cb_off,
                                    float* dx_tok, cudaStream_t stream) {
  DeconvBwdParams p{};
  DP_REQUIRE(fill_grad(p.g, n_g, g_f32, g_cb_total, g_cb_off, g_f16, g16_cb_total, g16_cb_off) == 0,
             "deconv2x_bwd_data: at most 3 fp32 gradient sources");
  DP_REQUIRE((dx_c8 != nullptr) != (dx_tok != nullptr), "deconv2x_bwd_data: exactly one output");
  DP_REQUIRE(dx_tok == nullptr || Ci % 8 == 0, "deconv2x_bwd_data: token output needs C_in % 8 == 0");
  p.Co = Co; p.w = w; p.Ci = Ci; p.N = N; p.D = D; p.H = H; p.W = W;
  p.dx_c8 = dx_c8; p.dx_cb_total = dx_cb_total; p.dx_cb_off = dx_cb_off;
  p.dx_tok = dx_tok; p.tok_C = Ci; p.tok_nstride = static_cast<long long>(D) * H * W * Ci;
  const size_t smem = static_cast<size_t>(8) * Co * 8 * sizeof(float);
  DP_REQUIRE(smem <= 200 * 1024, "deconv2x_bwd_data: C_out too large for the shared-memory weight tile");
  if (smem > 48 * 1024)
    DP_CHECK(cudaFuncSetAttribute(deconv2x_bwd_data_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
  dim3 grid(nblk(static_cast<long long>(D) * H * W, 128), (Ci + 7) / 8, N);
  deconv2x_bwd_data_kernel<<<grid, 128, smem, stream>>>(p);
  return check_cuda(cudaGetLastError(), "deconv2x_bwd_data");
}

extern "C" int dp_head_bwd(const float* g, const void* x_hi, const void* x_lo, int x_cb_total, int x_cb_off, int C,
                           const float* w, int N, long long vox, float* dx, int dx_cb_total, int dx_cb_off, double* dw,
                           double* db, cudaStream_t stream) {
  HeadBwdParams p{g, static_cast<const __half*>(x_hi), static_cast<const __half*>(x_lo), x_cb_total, x_cb_off, C, w, vox,
                  dx, dx_cb_total, dx_cb_off, dw, db};
  dim3 grid(nblk(vox, 256), static_cast<unsigned>(N * ((C + 7) / 8)));
  head_bwd_kernel<<<grid, 256, 0, stream>>>(p);
  return check_cuda(cudaGetLastError(), "head_bwd");
}

extern "C" int dp_masked_l1(const float* pred, const float* gt, int N, int S, int s, double* acc, int phase, float coef,
                            float* dpred, cudaStream_t stream) {
  DP_REQUIRE(s >= 2 && S >= s, "masked_l1: sizes");
  DP_REQUIRE(phase == 0 || dpred != nullptr, "masked_l1: backward needs dpred");
  const long long total = static_cast<long long>(s) * s * s * N;
  masked_l1_kernel<<<nblk(total, 256), 256, 0, stream>>>(pred, gt, N, S, s, acc, phase, coef, dpred);
  return check_cuda(cudaGetLastError(), "masked_l1");
}

extern "C" int dp_lerp2x_bwd(const float* g, long long outer, int len_in, long long inner, float* out, cudaStream_t stream) {
  DP_REQUIRE(outer > 0 && len_in > 0 && inner > 0, "lerp2x_bwd: empty tensor");
  lerp2x_bwd_kernel<<<nblk(outer * len_in * inner, 256), 256, 0, stream>>>(g, outer, len_in, inner, out);
  return check_cuda(cudaGetLastError(), "lerp2x_bwd");
}

extern "C" int dp_genloss_finalize(const double* acc, int n_scales, float delta1, float delta2, const double* acc_a,
                                   float weight_a, float* loss, cudaStream_t stream) {
  genloss_finalize_kernel<<<1, 1, 0, stream>>>(acc, n_scales, delta1, delta2, acc_a, weight_a, loss);
  return check_cuda(cudaGetLastError(), "genloss_finalize");
}

extern "C" int dp_adamw(float* p, const float* g, float* m, float* v, long long n, float lr, float beta1, float beta2,
                        float eps, float weight_decay, int step, float inv_scale, const int* found_inf,
                        cudaStream_t stream) {
  DP_REQUIRE(step >= 1, "adamw: step counts from 1");
  const float bc1 = 1.f - powf(beta1, static_cast<float>(step));
  const float bc2 = sqrtf(1.f - powf(beta2, static_cast<float>(step)));
  adamw_kernel<<<nblk(n, 256), 256, 0, stream>>>(p, g, m, v, n, lr, beta1, beta2, eps, weight_decay, bc1, bc2, inv_scale,
                                                  found_inf);
  return check_cuda(cudaGetLastError(), "adamw");
}

extern "C" int dp_adamw_dev(float* p, const float* g, float* m, float* v, long long n, float lr, float beta1, float beta2,
                            float eps, float weight_decay, float inv_scale, const int* found_inf, int* state,
                            cudaStream_t stream) {
  DP_REQUIRE(state != nullptr, "adamw_dev: device optimizer state {steps, consecutive skips, total skips} required");
  adamw_dev_kernel<<<nblk(n, 256), 256, 0, stream>>>(p, g, m, v, n, lr, beta1, beta2, eps, weight_decay, inv_scale, found_inf,
                                                      state);
  DP_CHECK(cudaGetLastError());
  adam_state_kernel<<<1, 1, 0, stream>>>(found_inf, state);
  return check_cuda(cudaGetLastError(), "adamw_dev");
}

extern "C" int dp_grad_check(const float* g, long long n, int* found_inf, cudaStream_t stream) {
  grad_check_kernel<<<nblk(n, 256), 256, 0, stream>>>(g, n, found_inf);
  return check_cuda(cudaGetLastError(), "grad_check");
}

extern "C" int dp_conv3d_wgrad(const void* x_c8, int x_cb_total, const uint8_t* chunk_cb, const int* chunk_ci0,
                               const int* chunk_nci, int n_chunks, const void* g_c8, int g_cb_total, int g_cb_off, int N,
                               int D, int H, int W, int cin, int cout, int k, int dil, float* ws, int splits,
                               cudaStream_t stream) {
  DP_REQUIRE(k == 1 || k == 3 || k == 7, "conv3d_wgrad: k must be 1, 3 or 7 (got %d)", k);
  DP_REQUIRE(cout % 16 == 0 && n_chunks >= 1 && n_chunks <= 64 && splits >= 1, "conv3d_wgrad: C_out %% 16, 1..64 chunks");
  ConvWgradParams p{};
  p.x = static_cast<const __half*>(x_c8); p.x_cb_total = x_cb_total; p.n_chunks = n_chunks;
  for (int i = 0; i < n_chunks; ++i) {
    p.chunk_cb[i] = chunk_cb[i];
    p.chunk_ci0[i] = static_cast<int16_t>(chunk_ci0[i]);
    p.chunk_nci[i] = static_cast<uint8_t>(chunk_nci[i]);
  }
  p.g = static_cast<const __half*>(g_c8); p.g_cb_total = g_cb_total; p.g_cb_off = g_cb_off; p.Co = cout; p.Ci = cin;
  p.N = N; p.D = D; p.H = H; p.W = W; p.dil = dil; p.ws = ws; p.splits = splits;
  p.wsize = static_cast<long long>(cout) * cin * k * k * k;
  const int P = (k / 2) * dil;
  size_t smem = static_cast<size_t>(8) * (2 * W + 2 * P) * 16 * sizeof(float);
  smem = std::max(smem, static_cast<size_t>(8) * 32 * k * 8 * sizeof(float));
  DP_REQUIRE(smem <= 220 * 1024, "conv3d_wgrad: row of %d voxels does not fit the shared-memory staging", W);
  dim3 grid(k * k, n_chunks * (cout / 16), splits);
  auto launch = [&](auto kern) -> int {
    if (smem > 48 * 1024) DP_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    kern<<<grid, 256, smem, stream>>>(p);
    return check_cuda(cudaGetLastError(), "conv3d_wgrad");
  };
  if (k == 7) return launch(conv_wgrad_kernel<7>);
  if (k == 3) return launch(conv_wgrad_kernel<3>);
  return launch(conv_wgrad_kernel<1>);
}

extern "C" int dp_layernorm_bwd(const float* x, const float* gamma, const float* dy, const float* add, int rows, int cols,
                                float* dx, double* dgamma, double* dbeta, cudaStream_t stream) {
  DP_REQUIRE(cols <= 1024, "layernorm_bwd: cols <= 1024");
  layernorm_bwd_kernel<<<(rows + LN_ROWS - 1) / LN_ROWS, 256, 0, stream>>>(x, gamma, dy, add, rows, cols, dx, dgamma, dbeta);
  return check_cuda(cudaGetLastError(), "layernorm_bwd");
}

extern "C" int dp_softmax_bwd(const void* probs, int ld_p, const float* dprobs, int ld_dp, int rows, int cols, void* ds,
                              int ld_ds, cudaStream_t stream) {
  softmax_bwd_kernel<<<(rows + 7) / 8, 256, 0, stream>>>(static_cast<const __half*>(probs), ld_p, dprobs, ld_dp, rows, cols,
                                                        static_cast<__half*>(ds), ld_ds);
  return check_cuda(cudaGetLastError(), "softmax_bwd");
}

extern "C" int dp_act_fwd(const float* u, long long n, int act, void* y_f16, cudaStream_t stream) {
  act_fwd_kernel<<<nblk(n, 256), 256, 0, stream>>>(u, n, act, static_cast<__half*>(y_f16));
  return check_cuda(cudaGetLastError(), "act_fwd");
}

extern "C" int dp_act_bwd(const float* u, const float* dh, long long n, int act, float* du_f32, void* du_f16,
                          cudaStream_t stream) {
  act_bwd_kernel<<<nblk(n, 256), 256, 0, stream>>>(u, dh, n, act, du_f32, static_cast<__half*>(du_f16));
  return check_cuda(cudaGetLastError(), "act_bwd");
}

extern "C" int dp_transpose(const void* src, int src_f32, long long src_batch_stride, int ld_src, int R, int C, void* dst_f16,
                            long long dst_batch_stride, int ld_dst, int batch, float scale, cudaStream_t stream) {
  dim3 grid((C + 31) / 32, (R + 31) / 32, batch);
  transpose_kernel<<<grid, 256, 0, stream>>>(src, src_f32, src_batch_stride, ld_src, R, C, static_cast<__half*>(dst_f16),
                                            dst_batch_stride, ld_dst, scale);
  return check_cuda(cudaGetLastError(), "transpose");
}

extern "C" int dp_heads(const void* src, int src_f32, void* dst, int dst_f32, int B, int T, int heads, int hd, int ld,
                        int col0, int merge, float scale, cudaStream_t stream) {
  const long long total = static_cast<long long>(B) * heads * T * hd;
  heads_kernel<<<nblk(total, 256), 256, 0, stream>>>(src, src_f32, dst, dst_f32, B, T, heads, hd, ld, col0, merge, scale);
  return check_cuda(cudaGetLastError(), "heads");
}

extern "C" int dp_colsum(const float* a, int rows, long long cols, double* out, cudaStream_t stream) {
  dim3 grid(nblk(cols, 256), (rows + 63) / 64);
  colsum_kernel<<<grid, 256, 0, stream>>>(a, rows, cols, out);
  return check_cuda(cudaGetLastError(), "colsum");
}

extern "C" int dp_add(const float* a, const float* b, long long n, float* y_f32, void* y_f16, cudaStream_t stream) {
  add_kernel<<<nblk(n, 256), 256, 0, stream>>>(a, b, n, y_f32, static_cast<__half*>(y_f16));
  return check_cuda(cudaGetLastError(), "add");
}

extern "C" int dp_pack_conv_weight(const float* w, int cout, int cin, int k, int transpose_flip, const int* chunk_ci0,
                                   const int* chunk_nci, int n_chunks, int stacked, void* out, cudaStream_t stream) {
  DP_REQUIRE(n_chunks >= 1 && n_chunks <= 192, "pack_conv_weight: 1..192 chunks");
  PackParams p{};
  p.w = w; p.Co = cout; p.Ci = cin; p.k = k; p.flip = transpose_flip;
  const long long taps = static_cast<long long>(k) * k * k;
  // logical (cout, cin) index into the stored tensor: [cout][cin][taps], or [cin][cout][taps] read transposed
  p.s_co = transpose_flip ? taps : taps * cin;
  p.s_ci = transpose_flip ? taps * cout : taps;
  for (int i = 0; i < n_chunks; ++i) {
    p.chunk_ci0[i] = static_cast<int16_t>(chunk_ci0[i]);
    p.chunk_nci[i] = static_cast<uint8_t>(chunk_nci[i]);
  }
  p.n_chunks = n_chunks; p.stacked = stacked; p.out = static_cast<__half*>(out);
  p.total = static_cast<long long>(stacked ? (k + 1) * (k + 1) : k) * n_chunks * k * k * 2 * cout * 8;
  pack_conv_weight_kernel<<<nblk(p.total, 256), 256, 0, stream>>>(p);
  return check_cuda(cudaGetLastError(), "pack_conv_weight");
}

extern "C" int dp_cast_f16(const float* src, long long n, void* dst, cudaStream_t stream) {
  cast_f16_kernel<<<nblk(n, 256), 256, 0, stream>>>(src, n, static_cast<__half*>(dst));
  return check_cuda(cudaGetLastError(), "cast_f16");
}

extern "C" int dp_dice_ce(const float* logits_c8, int cb_total, const float* label, int N, int C, long long vox, double* acc,
                          int phase, float coef, void* g_f16, int g_cb_total, cudaStream_t stream) {
  DP_REQUIRE(C >= 2 && C <= 8, "dice_ce: 2..8 classes (got %d)", C);
  DP_REQUIRE(phase == 0 || g_f16 != nullptr, "dice_ce: backward needs the gradient tensor");
  dim3 grid(nblk(vox, 256), static_cast<unsigned>(N));
  dice_ce_kernel<<<grid, 256, 0, stream>>>(logits_c8, cb_total, label, N, C, vox, acc, phase, coef, static_cast<__half*>(g_f16),
                                          g_cb_total);
  return check_cuda(cudaGetLastError(), "dice_ce");
}

extern "C" int dp_dice_ce_finalize(const double* acc, int N, int C, long long vox, float* loss, cudaStream_t stream) {
  dice_ce_finalize_kernel<<<1, 1, 0, stream>>>(acc, N, C, vox, loss);
  return check_cuda(cudaGetLastError(), "dice_ce_finalize");
}
