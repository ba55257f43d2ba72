// HBM-bound fused kernels around the tensor-core contractions: layout packing, InstanceNorm apply
// (+affine, +activation, +residual, +chained statistics), 1x1x1 convolutions with normalise-on-load,
// 2x transposed convolution, trilinear 2x upsampling, LayerNorm, softmax, patch flattening, the
// seg-argmax -> dose-input hand-off, and a generic direct convolution (strided convs).
//
// Activation layout "c8": [N][C/8][D][H][W][8]; one thread moves one 16-byte (8 x fp16) or 32-byte
// (8 x fp32) channel vector, threads along W -> fully coalesced.  Per-(n,c) statistics are
// {sum, sum of squares} in fp64, accumulated with one atomicAdd per block per channel.
#include "common.cuh"
#include "dose_b200.h"

namespace dp {

__device__ __forceinline__ void load8(const __half* hi, const __half* lo, size_t off, float (&x)[8]) {
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
__device__ __forceinline__ void load8f(const float* p, size_t off, float (&x)[8]) {
  ld_global_v8f(p + off, x);
}
__device__ __forceinline__ void store8(__half* hi, __half* lo, size_t off, const float (&x)[8]) {
  // hi = fp16(x), lo = fp16(x - hi), with the packed two-at-a-time conversions (same rounding as the scalar ones)
  __align__(16) __half2 h[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) h[j] = __floats2half2_rn(x[2 * j], x[2 * j + 1]);
  *reinterpret_cast<uint4*>(hi + off) = *reinterpret_cast<const uint4*>(h);
  if (lo != nullptr) {
    __align__(16) __half2 l[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 f = __half22float2(h[j]);
      l[j] = __floats2half2_rn(x[2 * j] - f.x, x[2 * j + 1] - f.y);
    }
    *reinterpret_cast<uint4*>(lo + off) = *reinterpret_cast<const uint4*>(l);
  }
}

// one switch per 8 values (the per-element switch of act_apply cost more instructions than the arithmetic)
__device__ __forceinline__ void act8(float (&y)[8], int act) {
  switch (act) {
    case ACT_RELU:
#pragma unroll
      for (int j = 0; j < 8; ++j) y[j] = fmaxf(y[j], 0.f);
      break;
    case ACT_LRELU:
#pragma unroll
      for (int j = 0; j < 8; ++j) y[j] = y[j] > 0.f ? y[j] : 0.01f * y[j];
      break;
    case ACT_MISH:
#pragma unroll
      for (int j = 0; j < 8; ++j) y[j] = act_apply(y[j], ACT_MISH);
      break;
    case ACT_GELU:
#pragma unroll
      for (int j = 0; j < 8; ++j) y[j] = act_apply(y[j], ACT_GELU);
      break;
    default: break;
  }
}
// hi = fp16(x), lo = fp16(x - hi) with the packed two-at-a-time conversions
__device__ __forceinline__ void store8p(__half* hi, __half* lo, size_t off, const float (&x)[8]) {
  __align__(16) __half2 h[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) h[j] = __floats2half2_rn(x[2 * j], x[2 * j + 1]);
  *reinterpret_cast<uint4*>(hi + off) = *reinterpret_cast<const uint4*>(h);
  if (lo != nullptr) {
    __align__(16) __half2 l[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 f = __half22float2(h[j]);
      l[j] = __floats2half2_rn(x[2 * j] - f.x, x[2 * j + 1] - f.y);
    }
    *reinterpret_cast<uint4*>(lo + off) = *reinterpret_cast<const uint4*>(l);
  }
}
// mean / rstd of channel c of image n from fp64 {sum, sumsq}; biased variance, eps 1e-5 (InstanceNorm3d).
__device__ __forceinline__ void finalize_stats(const double* stats, size_t idx, double inv_count, float& mean,
                                               float& rstd) {
  const double s = stats[idx * 2], ss = stats[idx * 2 + 1];
  const double m = s * inv_count;
  double var = ss * inv_count - m * m;
  if (var < 0.0) var = 0.0;
  mean = static_cast<float>(m);
  rstd = rsqrtf(static_cast<float>(var) + 1e-5f);     // variance itself is formed in fp64 (cancellation)
}

// Block-wide reduction of 8 channel sums + 8 sums of squares, then one fp64 atomic per channel.
__device__ __forceinline__ void block_accumulate_sums(const float (&a1)[8], const float (&a2)[8], double* stats, size_t idx0) {
  __shared__ float red[2][8][8];  // [sum|sq][warp][channel]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const float s1 = warp_sum(a1[j]), s2 = warp_sum(a2[j]);
    if (lane == 0) { red[0][warp][j] = s1; red[1][warp][j] = s2; }
  }
  __syncthreads();
  if (threadIdx.x < 16) {
    const int which = threadIdx.x >> 3, j = threadIdx.x & 7;
    float t = 0.f;
    for (int w = 0; w < (blockDim.x >> 5); ++w) t += red[which][w][j];
    atomicAdd(&stats[(idx0 + j) * 2 + which], static_cast<double>(t));
  }
}
__device__ __forceinline__ void block_accumulate_stats(const float (&y)[8], bool valid, double* stats, size_t idx0) {
  __shared__ float red[2][8][8];  // [sum|sq][warp][channel]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const float v = valid ? y[j] : 0.f;
    const float s1 = warp_sum(v), s2 = warp_sum(v * v);
    if (lane == 0) { red[0][warp][j] = s1; red[1][warp][j] = s2; }
  }
  __syncthreads();
  if (threadIdx.x < 16) {
    const int which = threadIdx.x >> 3, j = threadIdx.x & 7;
    float t = 0.f;
    for (int w = 0; w < (blockDim.x >> 5); ++w) t += red[which][w][j];
    atomicAdd(&stats[(idx0 + j) * 2 + which], static_cast<double>(t));
  }
}

// ------------------------------------------------------------------ NCDHW fp32 -> c8 fp16 (hi[/lo])
__global__ void pack_ncdhw_kernel(const float* __restrict__ src, int C, long long vox, __half* hi, __half* lo,
                                  int cb_total, int cb_off, int ncb) {
  const long long v = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  const int cb = blockIdx.y % ncb, n = blockIdx.y / ncb;
  if (v >= vox) return;
  float x[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int c = cb * 8 + j;
    x[j] = c < C ? src[(static_cast<size_t>(n) * C + c) * vox + v] : 0.f;
  }
  store8(hi, lo, ((static_cast<size_t>(n) * cb_total + cb_off + cb) * vox + v) * 8, x);
}

// ------------------------------------------------------------------ c8 fp16 -> NCDHW fp32
__global__ void unpack_c8_kernel(const __half* __restrict__ hi, const __half* __restrict__ lo, int cb_total, int cb_off,
                                 int C, long long vox, float* dst) {
  const long long v = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  const int ncb = (C + 7) / 8;
  const int cb = blockIdx.y % ncb, n = blockIdx.y / ncb;
  if (v >= vox) return;
  float x[8];
  load8(hi, lo, ((static_cast<size_t>(n) * cb_total + cb_off + cb) * vox + v) * 8, x);
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int c = cb * 8 + j;
    if (c < C) dst[(static_cast<size_t>(n) * C + c) * vox + v] = x[j];
  }
}

// ------------------------------------------------------------------ InstanceNorm apply (+residual, +chained stats)
struct NormActParams {
  const float* raw_f32; const __half* raw_hi; const __half* raw_lo; int in_cb_total, in_cb_off;
  const double* stats;            // [N][C][2] of the raw tensor, or null (identity)
  const float* gamma; const float* beta;
  int act;
  // residual: either a plain c8 fp16 tensor, or a raw fp32 tensor with its own instance statistics
  const __half* res_hi; const __half* res_lo; const float* res_raw; const double* res_stats;
  int res_cb_total, res_cb_off;
  int act_after_res;
  __half* out_hi; __half* out_lo; int out_cb_total, out_cb_off;
  double* stats_out;              // statistics of the produced tensor (chained InstanceNorm), or null
  int C, ncb; long long vox; double inv_vox;
  // optional second copy in space-to-depth layout: channel block (parity*ncb + cb) of a half-resolution
  // tensor, parity = (d&1)*4 + (h&1)*2 + (w&1); feeds the stride-2 convs as sparse stride-1 convs
  __half* s2d_hi; __half* s2d_lo; int s2d_cb_total, s2d_cb_off, D, H, W;
  // closed-form residual branch norm3(conv3(x)) of a res block whose input x has ONE channel (seg encoder1): conv3 is
  // w_c * x, so its InstanceNorm is w_c (x - mean_x) / sqrt(w_c^2 var_x + eps): an affine map of x per output channel,
  // evaluated here from the planar fp32 input instead of materialising conv3's 16-channel output
  const float* res_x; const float* res_w; const double* res_xstats;
};

constexpr int NA_IT = 4;      // voxel chunks per block: amortises the statistics prologue
// ncu (profiles/r1_norm_pointwise_ncu.md): the first version of this kernel was INSTRUCTION bound (sm throughput 78 %,
// 400 thread-instructions per 8-channel vector: per-element channel predicates, a per-element activation switch,
// scalar fp16 conversions).  Channels beyond C are neutralised through the per-channel constants instead
// (rstd = 0 -> 0), the activation switch is hoisted out of the element loop, conversions are packed.
__global__ void __launch_bounds__(256, 4) norm_act_kernel(const NormActParams p) {
  const int cb = blockIdx.y % p.ncb, n = blockIdx.y / p.ncb;
  const double inv = p.inv_vox;                    // 1/vox from the host: no fp64 division on the device
  __shared__ float s_mean[2][8], s_rstd[2][8], s_gamma[8], s_beta[8];
  if (threadIdx.x < 16) {
    const int which = threadIdx.x >> 3, j = threadIdx.x & 7, c = cb * 8 + j;
    const double* st = which ? p.res_stats : p.stats;
    float m = 0.f, r = 1.f;
    if (st != nullptr && c < p.C) finalize_stats(st, static_cast<size_t>(n) * p.C + c, inv, m, r);
    if (c >= p.C) r = 0.f;                          // padded channels: (x - 0) * 0 = 0 through every activation
    s_mean[which][j] = m;
    s_rstd[which][j] = r;
    if (which == 0) {
      s_gamma[j] = (p.gamma && c < p.C) ? p.gamma[c] : 1.f;
      s_beta[j] = (p.gamma && c < p.C) ? p.beta[c] : 0.f;
    } else if (p.res_x != nullptr) {
      // residual = a_c * x + b_c, stored as (x - mean_x) * a_c in the (mean, rstd) slots
      const double sx = p.res_xstats[n * 2], sxx = p.res_xstats[n * 2 + 1];
      const double mx = sx * inv;
      double vx = sxx * inv - mx * mx;
      if (vx < 0.0) vx = 0.0;
      const float wc = c < p.C ? p.res_w[c] : 0.f;
      s_mean[1][j] = static_cast<float>(mx);
      s_rstd[1][j] = wc * rsqrtf(wc * wc * static_cast<float>(vx) + 1e-5f);
    }
  }
  __syncthreads();
  float mean0[8], rstd0[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) { mean0[j] = s_mean[0][j]; rstd0[j] = s_rstd[0][j]; }
  const bool has_res = p.res_hi || p.res_raw || p.res_x;
  const bool affine = p.gamma != nullptr;
  const bool want_stats = p.stats_out != nullptr;
  float a1[8], a2[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) { a1[j] = 0.f; a2[j] = 0.f; }
  const size_t in_base = (static_cast<size_t>(n) * p.in_cb_total + p.in_cb_off + cb) * p.vox;
  const size_t out_base = (static_cast<size_t>(n) * p.out_cb_total + p.out_cb_off + cb) * p.vox;
#pragma unroll 1
  for (int it = 0; it < NA_IT; ++it) {
    const long long v = (blockIdx.x * static_cast<long long>(NA_IT) + it) * blockDim.x + threadIdx.x;
    if (v >= p.vox) break;
    float y[8];
    if (p.raw_f32) load8f(p.raw_f32, (in_base + v) * 8, y); else load8(p.raw_hi, p.raw_lo, (in_base + v) * 8, y);
#pragma unroll
    for (int j = 0; j < 8; ++j) y[j] = (y[j] - mean0[j]) * rstd0[j];
    if (affine) {
#pragma unroll
      for (int j = 0; j < 8; ++j) y[j] = fmaf(y[j], s_gamma[j], s_beta[j]);
    }
    act8(y, p.act);
    if (has_res) {
      float r8[8];
      const size_t r_off = ((static_cast<size_t>(n) * p.res_cb_total + p.res_cb_off + cb) * p.vox + v) * 8;
      if (p.res_x) {
        const float xv = __ldg(&p.res_x[static_cast<size_t>(n) * p.vox + v]);
#pragma unroll
        for (int j = 0; j < 8; ++j) r8[j] = xv;
      } else if (p.res_raw) load8f(p.res_raw, r_off, r8); else load8(p.res_hi, p.res_lo, r_off, r8);
#pragma unroll
      for (int j = 0; j < 8; ++j) y[j] += (r8[j] - s_mean[1][j]) * s_rstd[1][j];
      act8(y, p.act_after_res);
    }
    if (p.out_hi) store8p(p.out_hi, p.out_lo, (out_base + v) * 8, y);
    if (p.s2d_hi) {
      const int w = static_cast<int>(v % p.W), h = static_cast<int>((v / p.W) % p.H), d = static_cast<int>(v / (static_cast<long long>(p.W) * p.H));
      const int parity = ((d & 1) << 2) | ((h & 1) << 1) | (w & 1);
      const size_t vh = (static_cast<size_t>(d >> 1) * (p.H >> 1) + (h >> 1)) * (p.W >> 1) + (w >> 1);
      store8p(p.s2d_hi, p.s2d_lo,
              ((static_cast<size_t>(n) * p.s2d_cb_total + p.s2d_cb_off + parity * p.ncb + cb) * (p.vox >> 3) + vh) * 8, y);
    }
    if (want_stats) {
#pragma unroll
      for (int j = 0; j < 8; ++j) { a1[j] += y[j]; a2[j] = fmaf(y[j], y[j], a2[j]); }
    }
  }
  if (want_stats) block_accumulate_sums(a1, a2, p.stats_out, static_cast<size_t>(n) * p.C + cb * 8);
}

// ------------------------------------------------------------------ InstanceNorm apply + 1x1x1 head in one pass
// The decoder outputs feed both the next level and a C -> {1, 8} head (dose_convertors dose_pyfer.py:290-300,316-317,
// conv_out_A :353,359, the seg logits base_blocks.py:151).  As a separate launch the head re-read the whole activation
// (and spent 16 FMA columns on 1 output channel); here the thread that normalises a voxel keeps all its channels in
// registers and adds the head's dot products, so the head costs one planar fp32 store.
struct NormHeadParams {
  const float* raw_f32; int in_cb_total;
  const double* stats; const float* gamma; const float* beta; int act;
  __half* out_hi; __half* out_lo; int out_cb_total, out_cb_off;
  const float* head_w; const float* head_b; int head_co;       // [head_co][C] fp32, [head_co]
  float* head_out;                                              // NCDHW fp32 [N][head_co][vox]
  int C, ncb; long long vox; double inv_vox;
};
constexpr int NH_MAXC = 128, NH_MAXCO = 8;

template <int CO>
__global__ void __launch_bounds__(256) norm_act_head_kernel(const NormHeadParams p) {
  __shared__ float s_mean[NH_MAXC], s_rstd[NH_MAXC], s_gamma[NH_MAXC], s_beta[NH_MAXC];   // same arithmetic as norm_act_kernel
  __shared__ float s_w[CO][NH_MAXC];
  const int n = blockIdx.y;
  const int cpad = p.ncb * 8;
  const bool affine = p.gamma != nullptr;
  for (int c = threadIdx.x; c < cpad; c += blockDim.x) {
    float m = 0.f, r = 0.f;                              // padded channels: (x - 0) * 0 = 0 through every activation
    if (c < p.C) finalize_stats(p.stats, static_cast<size_t>(n) * p.C + c, p.inv_vox, m, r);
    s_mean[c] = m;
    s_rstd[c] = r;
    s_gamma[c] = (affine && c < p.C) ? p.gamma[c] : 1.f;
    s_beta[c] = (affine && c < p.C) ? p.beta[c] : 0.f;
  }
  for (int i = threadIdx.x; i < CO * cpad; i += blockDim.x) {
    const int co = i / cpad, c = i % cpad;
    s_w[co][c] = (c < p.C && co < p.head_co) ? p.head_w[co * p.C + c] : 0.f;
  }
  __syncthreads();
  const long long v = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (v >= p.vox) return;
  float acc[CO];
#pragma unroll
  for (int co = 0; co < CO; ++co) acc[co] = (p.head_b && co < p.head_co) ? __ldg(&p.head_b[co]) : 0.f;
  for (int cb = 0; cb < p.ncb; ++cb) {
    float y[8];
    load8f(p.raw_f32, ((static_cast<size_t>(n) * p.in_cb_total + cb) * p.vox + v) * 8, y);
#pragma unroll
    for (int j = 0; j < 8; ++j) y[j] = (y[j] - s_mean[cb * 8 + j]) * s_rstd[cb * 8 + j];
    if (affine) {
#pragma unroll
      for (int j = 0; j < 8; ++j) y[j] = fmaf(y[j], s_gamma[cb * 8 + j], s_beta[cb * 8 + j]);
    }
    act8(y, p.act);
    if (p.out_hi) {
      store8p(p.out_hi, p.out_lo, ((static_cast<size_t>(n) * p.out_cb_total + p.out_cb_off + cb) * p.vox + v) * 8, y);
      if (p.out_lo == nullptr) {                           // the head must see what every other consumer of this tensor sees
#pragma unroll
        for (int j = 0; j < 8; ++j) y[j] = __half2float(__float2half_rn(y[j]));
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float h = __half2float(__float2half_rn(y[j]));
          y[j] = h + __half2float(__float2half_rn(y[j] - h));
        }
      }
    }
#pragma unroll
    for (int co = 0; co < CO; ++co) {
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[co] = fmaf(y[j], s_w[co][cb * 8 + j], acc[co]);
    }
  }
#pragma unroll
  for (int co = 0; co < CO; ++co)
    if (co < p.head_co) p.head_out[(static_cast<size_t>(n) * p.head_co + co) * p.vox + v] = acc[co];
}

// ------------------------------------------------------------------ 1x1x1 convolution, normalise-on-load, <=3 sources
struct PwSource {
  const __half* hi; const __half* lo; const float* raw; int cb_total, cb_off, C;
  const double* stats; int act;          // x <- act((x-mean)*rstd) when stats != null, else act(x)
};
struct PointwiseParams {
  PwSource src[3]; int nsrc;
  const float* w;       // [C_out][C_in_total] fp32 (C_in_total = sum of source C)
  const float* bias;    // [C_out] or null
  int cin_total, cout;
  long long vox; double inv_vox;
  float* out_raw; int out_cb_total, out_cb_off;     // c8 fp32 (+stats) ...
  __half* out_hi; __half* out_lo;                   // ... or c8 fp16
  float* out_planar;                                // ... or NCDHW fp32 [N][C_out][vox]
  double* stats_out;
  int out_act;
};
constexpr int PW_CO = 16;   // output channels per block pass
constexpr int PW_IT = 4;    // voxel chunks per block: amortises the weight / statistics prologue and the stats atomics

// Shared-memory weights are indexed by PADDED input channel (every source rounded up to whole 8-channel blocks, zero
// rows for the padding), so the inner loop needs no channel predicates; the activation switch runs once per 8-vector.
__global__ void __launch_bounds__(128, 6) pointwise_kernel(const PointwiseParams p) {
  extern __shared__ float wsm[];           // [cin_pad][PW_CO] weights, then [cin_pad] mean, [cin_pad] rstd
  int cin_pad = 0;
  for (int s = 0; s < p.nsrc; ++s) cin_pad += ((p.src[s].C + 7) / 8) * 8;
  float* s_mean = wsm + cin_pad * PW_CO;
  float* s_rstd = s_mean + cin_pad;
  const int co0 = blockIdx.y * PW_CO;
  const int n = blockIdx.z;
  const double inv = p.inv_vox;
  {
    int base = 0, lbase = 0;               // padded / logical first channel of the source
    for (int s = 0; s < p.nsrc; ++s) {
      const int Cs = p.src[s].C, Cp = ((Cs + 7) / 8) * 8;
      for (int i = threadIdx.x; i < Cp * PW_CO; i += blockDim.x) {
        const int c = i / PW_CO, j = i % PW_CO;
        wsm[(base + c) * PW_CO + j] =
            (c < Cs && co0 + j < p.cout) ? p.w[static_cast<size_t>(co0 + j) * p.cin_total + lbase + c] : 0.f;
      }
      for (int c = threadIdx.x; c < Cp; c += blockDim.x) {
        float m = 0.f, r = 1.f;
        if (p.src[s].stats && c < Cs) finalize_stats(p.src[s].stats, static_cast<size_t>(n) * Cs + c, inv, m, r);
        s_mean[base + c] = m;
        s_rstd[base + c] = r;
      }
      base += Cp;
      lbase += Cs;
    }
  }
  __syncthreads();
  float st1[PW_CO], st2[PW_CO];
#pragma unroll
  for (int j = 0; j < PW_CO; ++j) { st1[j] = 0.f; st2[j] = 0.f; }
  const bool want_stats = p.stats_out != nullptr;
#pragma unroll 1
  for (int it = 0; it < PW_IT; ++it) {
    const long long v = (blockIdx.x * static_cast<long long>(PW_IT) + it) * blockDim.x + threadIdx.x;
    if (v >= p.vox) break;
    float acc[PW_CO];
#pragma unroll
    for (int j = 0; j < PW_CO; ++j) acc[j] = (p.bias && co0 + j < p.cout) ? __ldg(&p.bias[co0 + j]) : 0.f;
    int base = 0;
    for (int s = 0; s < p.nsrc; ++s) {
      const PwSource& S = p.src[s];
      const int ncb = (S.C + 7) / 8;
      const bool norm = S.stats != nullptr;
      for (int cb = 0; cb < ncb; ++cb) {
        float x[8];
        const size_t off = ((static_cast<size_t>(n) * S.cb_total + S.cb_off + cb) * p.vox + v) * 8;
        if (S.raw) load8f(S.raw, off, x); else load8(S.hi, S.lo, off, x);
        const float* mr = s_mean + base + cb * 8;
        if (norm) {
          const float* rr = s_rstd + base + cb * 8;
#pragma unroll
          for (int j = 0; j < 8; ++j) x[j] = (x[j] - mr[j]) * rr[j];
        }
        act8(x, S.act);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float4* wr = reinterpret_cast<const float4*>(&wsm[(base + cb * 8 + j) * PW_CO]);
#pragma unroll
          for (int q = 0; q < PW_CO / 4; ++q) {
            const float4 w4 = wr[q];
            acc[4 * q + 0] = fmaf(x[j], w4.x, acc[4 * q + 0]);
            acc[4 * q + 1] = fmaf(x[j], w4.y, acc[4 * q + 1]);
            acc[4 * q + 2] = fmaf(x[j], w4.z, acc[4 * q + 2]);
            acc[4 * q + 3] = fmaf(x[j], w4.w, acc[4 * q + 3]);
          }
        }
      }
      base += ncb * 8;
    }
    if (p.out_act != ACT_NONE) {
#pragma unroll
      for (int j = 0; j < PW_CO; ++j) acc[j] = act_apply(acc[j], p.out_act);
    }
    if (p.out_planar) {
#pragma unroll
      for (int j = 0; j < PW_CO; ++j)
        if (co0 + j < p.cout) p.out_planar[(static_cast<size_t>(n) * p.cout + co0 + j) * p.vox + v] = acc[j];
    }
#pragma unroll
    for (int b = 0; b < PW_CO / 8; ++b) {
      if (co0 + b * 8 >= p.cout) break;
      const size_t off = ((static_cast<size_t>(n) * p.out_cb_total + p.out_cb_off + (co0 >> 3) + b) * p.vox + v) * 8;
      float y[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) y[j] = acc[b * 8 + j];
      if (p.out_raw) {
        st_global_v8f(p.out_raw + off, y);
      }
      if (p.out_hi) store8p(p.out_hi, p.out_lo, off, y);
    }
    if (want_stats) {
#pragma unroll
      for (int j = 0; j < PW_CO; ++j) { st1[j] += acc[j]; st2[j] = fmaf(acc[j], acc[j], st2[j]); }
    }
  }
  if (want_stats) {
#pragma unroll
    for (int b = 0; b < PW_CO / 8; ++b) {
      if (co0 + b * 8 >= p.cout) break;          // block-uniform
      float y1[8], y2[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) { y1[j] = st1[b * 8 + j]; y2[j] = st2[b * 8 + j]; }
      block_accumulate_sums(y1, y2, p.stats_out, static_cast<size_t>(n) * p.cout + co0 + b * 8);
      __syncthreads();
    }
  }
}

// ------------------------------------------------------------------ 1x1x1 convolution, weights in the constant bank
// ncu on pointwise_kernel: the 128 LDS.128 weight fetches per voxel saturate the shared-memory return path (512 B
// per warp instruction even when every lane reads the same address), 3-4x more time than the FMAs need.  For the
// inference plans the weights are known on the host, so this variant receives the 16 output channels' weights of
// one pass as KERNEL PARAMETERS: they sit in the constant bank and are consumed directly as FFMA operands (no load
// instruction at all).  NCB = padded input channel blocks (compile time, so every weight offset is an immediate).
struct PwBlock {
  const __half* hi; const __half* lo; const float* raw;   // already offset to this channel block of image 0
  long long n_stride;                                      // elements between images
  const double* stats; int stat_c0, stat_C;                // instance statistics of the source (or null), channel of j = 0
  int act;
};
template <int NCB>
struct PwConstParams {
  PwBlock blk[NCB];
  float w[NCB * 8][PW_CO];
  float bias[PW_CO];
  int cout, co0;
  long long vox; double inv_vox;
  float* out_raw; int out_cb_total, out_cb_off;
  __half* out_hi; __half* out_lo;
  float* out_planar;
  double* stats_out;
  int out_act;
};

template <int NCB>
__global__ void __launch_bounds__(128, 6) pointwise_cw_kernel(const __grid_constant__ PwConstParams<NCB> p) {
  __shared__ __align__(16) float s_mean[NCB * 8], s_rstd[NCB * 8];
  const int n = blockIdx.z;
  const int co0 = p.co0;
  for (int i = threadIdx.x; i < NCB * 8; i += blockDim.x) {
    const PwBlock& B = p.blk[i >> 3];
    const int c = B.stat_c0 + (i & 7);
    float m = 0.f, r = 1.f;
    if (B.stats && c < B.stat_C) finalize_stats(B.stats, static_cast<size_t>(n) * B.stat_C + c, p.inv_vox, m, r);
    s_mean[i] = m;
    s_rstd[i] = r;
  }
  __syncthreads();
  float st1[PW_CO], st2[PW_CO];
#pragma unroll
  for (int j = 0; j < PW_CO; ++j) { st1[j] = 0.f; st2[j] = 0.f; }
  const bool want_stats = p.stats_out != nullptr;
#pragma unroll 1
  for (int it = 0; it < PW_IT; ++it) {
    const long long v = (blockIdx.x * static_cast<long long>(PW_IT) + it) * blockDim.x + threadIdx.x;
    if (v >= p.vox) break;
    // packed accumulators: one FFMA2 (fma.rn.f32x2) updates two output channels -- the kernel is issue bound
    // (ncu: 1250 thread instructions per voxel at 32 input channels, 512 of them FFMA)
    unsigned long long acc2[PW_CO / 2];
#pragma unroll
    for (int j = 0; j < PW_CO / 2; ++j) acc2[j] = pack_f32x2(p.bias[2 * j], p.bias[2 * j + 1]);
#pragma unroll
    for (int b = 0; b < NCB; ++b) {
      const PwBlock& B = p.blk[b];
      float x[8];
      const size_t off = static_cast<size_t>(n) * B.n_stride + static_cast<size_t>(v) * 8;
      if (B.raw) load8f(B.raw, off, x); else load8(B.hi, B.lo, off, x);
      if (B.stats) {
        const float4 m0 = *reinterpret_cast<const float4*>(&s_mean[b * 8]), m1 = *reinterpret_cast<const float4*>(&s_mean[b * 8 + 4]);
        const float4 r0 = *reinterpret_cast<const float4*>(&s_rstd[b * 8]), r1 = *reinterpret_cast<const float4*>(&s_rstd[b * 8 + 4]);
        x[0] = (x[0] - m0.x) * r0.x; x[1] = (x[1] - m0.y) * r0.y; x[2] = (x[2] - m0.z) * r0.z; x[3] = (x[3] - m0.w) * r0.w;
        x[4] = (x[4] - m1.x) * r1.x; x[5] = (x[5] - m1.y) * r1.y; x[6] = (x[6] - m1.z) * r1.z; x[7] = (x[7] - m1.w) * r1.w;
      }
      act8(x, B.act);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const unsigned long long xx = pack_f32x2(x[j], x[j]);
#pragma unroll
        for (int q = 0; q < PW_CO / 2; ++q)
          acc2[q] = fma_f32x2(xx, pack_f32x2(p.w[b * 8 + j][2 * q], p.w[b * 8 + j][2 * q + 1]), acc2[q]);
      }
    }
    float acc[PW_CO];
#pragma unroll
    for (int j = 0; j < PW_CO / 2; ++j) unpack_f32x2(acc2[j], acc[2 * j], acc[2 * j + 1]);
    if (p.out_act != ACT_NONE) {
#pragma unroll
      for (int j = 0; j < PW_CO; ++j) acc[j] = act_apply(acc[j], p.out_act);
    }
    if (p.out_planar) {
#pragma unroll
      for (int j = 0; j < PW_CO; ++j)
        if (co0 + j < p.cout) p.out_planar[(static_cast<size_t>(n) * p.cout + co0 + j) * p.vox + v] = acc[j];
    }
#pragma unroll
    for (int b = 0; b < PW_CO / 8; ++b) {
      if (co0 + b * 8 >= p.cout) break;
      const size_t off = ((static_cast<size_t>(n) * p.out_cb_total + p.out_cb_off + (co0 >> 3) + b) * p.vox + v) * 8;
      float y[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) y[j] = acc[b * 8 + j];
      if (p.out_raw) {
        st_global_v8f(p.out_raw + off, y);
      }
      if (p.out_hi) store8p(p.out_hi, p.out_lo, off, y);
    }
    if (want_stats) {
#pragma unroll
      for (int j = 0; j < PW_CO; ++j) { st1[j] += acc[j]; st2[j] = fmaf(acc[j], acc[j], st2[j]); }
    }
  }
  if (want_stats) {
#pragma unroll
    for (int b = 0; b < PW_CO / 8; ++b) {
      if (co0 + b * 8 >= p.cout) break;          // block-uniform
      float y1[8], y2[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) { y1[j] = st1[b * 8 + j]; y2[j] = st2[b * 8 + j]; }
      block_accumulate_sums(y1, y2, p.stats_out, static_cast<size_t>(n) * p.cout + co0 + b * 8);
      __syncthreads();
    }
  }
}

// ------------------------------------------------------------------ ConvTranspose3d k=2 s=2 (no bias)
// out[n, co, 2d+i, 2h+j, 2w+l] = sum_ci in[n, ci, d, h, w] * W[ci, co, i, j, l]
struct DeconvParams {
  const __half* in_hi; const __half* in_lo;
  long long in_nstride, in_vstride, in_cbstride;     // element strides (c8: vstride 8; token-major: vstride C)
  int cin, cout, D, H, W;                            // input spatial dims
  const float* w;                                    // packed [8 parity][cin][cout] fp32
  __half* out_hi; __half* out_lo; int out_cb_total, out_cb_off;
};
constexpr int DC_CO = 16;
__global__ void __launch_bounds__(128) deconv2x_kernel(const DeconvParams p) {
  extern __shared__ float wsm[];   // [cin][DC_CO] for the current parity
  const int co0 = blockIdx.y * DC_CO, n = blockIdx.z;
  const long long vox_in = static_cast<long long>(p.D) * p.H * p.W;
  const long long v = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  const bool valid = v < vox_in;
  const int w = static_cast<int>(v % p.W), h = static_cast<int>((v / p.W) % p.H), d = static_cast<int>(v / (static_cast<long long>(p.W) * p.H));
  const int ncb = p.cin / 8;
  const long long vox_out = vox_in * 8;
  for (int q = 0; q < 8; ++q) {
    __syncthreads();
    for (int i = threadIdx.x; i < p.cin * DC_CO; i += blockDim.x) {
      const int ci = i / DC_CO, j = i % DC_CO;
      wsm[i] = (co0 + j < p.cout) ? p.w[(static_cast<size_t>(q) * p.cin + ci) * p.cout + co0 + j] : 0.f;
    }
    __syncthreads();
    if (!valid) continue;
    float acc[DC_CO];
#pragma unroll
    for (int j = 0; j < DC_CO; ++j) acc[j] = 0.f;
    for (int cb = 0; cb < ncb; ++cb) {
      float x[8];
      load8(p.in_hi, p.in_lo, static_cast<size_t>(n) * p.in_nstride + v * p.in_vstride + cb * p.in_cbstride, x);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float4* wr = reinterpret_cast<const float4*>(&wsm[(cb * 8 + j) * DC_CO]);
#pragma unroll
        for (int t = 0; t < DC_CO / 4; ++t) {
          const float4 w4 = wr[t];
          acc[4 * t + 0] = fmaf(x[j], w4.x, acc[4 * t + 0]);
          acc[4 * t + 1] = fmaf(x[j], w4.y, acc[4 * t + 1]);
          acc[4 * t + 2] = fmaf(x[j], w4.z, acc[4 * t + 2]);
          acc[4 * t + 3] = fmaf(x[j], w4.w, acc[4 * t + 3]);
        }
      }
    }
    const int i = q >> 2, j2 = (q >> 1) & 1, l = q & 1;
    const long long vo = (static_cast<long long>(2 * d + i) * (2 * p.H) + (2 * h + j2)) * (2 * p.W) + (2 * w + l);
#pragma unroll
    for (int b = 0; b < DC_CO / 8; ++b) {
      if (co0 + b * 8 >= p.cout) break;
      float y[8];
#pragma unroll
      for (int t = 0; t < 8; ++t) y[t] = acc[b * 8 + t];
      store8(p.out_hi, p.out_lo, ((static_cast<size_t>(n) * p.out_cb_total + p.out_cb_off + (co0 >> 3) + b) * vox_out + vo) * 8, y);
    }
  }
}

// ConvTranspose3d k2 s2 with static weights in the constant bank (see pointwise_cw_kernel): the input vector of a
// voxel stays in registers for all parities of the launch, weights are FFMA uniform operands, and a thread writes the
// two W-neighbours (l = 0, 1) back to back so every 32-byte sector is filled by one thread.
template <int NCB, int NPAR>
struct DeconvConstParams {
  const __half* in_hi; const __half* in_lo;
  long long in_nstride, in_cbstride;                 // c8 input: element strides between images / channel blocks
  int D, H, W, cout, co0, q0;                        // q0 = first parity of this launch
  __half* out_hi; __half* out_lo; int out_cb_total, out_cb_off;
  float w[NPAR][NCB * 8][DC_CO];
};
template <int NCB, int NPAR>
__global__ void __launch_bounds__(128) deconv2x_cw_kernel(const __grid_constant__ DeconvConstParams<NCB, NPAR> p) {
  const int n = blockIdx.z;
  const long long vox_in = static_cast<long long>(p.D) * p.H * p.W;
  const long long v = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (v >= vox_in) return;
  const int w = static_cast<int>(v % p.W), h = static_cast<int>((v / p.W) % p.H), d = static_cast<int>(v / (static_cast<long long>(p.W) * p.H));
  float x[NCB][8];
#pragma unroll
  for (int cb = 0; cb < NCB; ++cb)
    load8(p.in_hi, p.in_lo, static_cast<size_t>(n) * p.in_nstride + static_cast<size_t>(v) * 8 + cb * p.in_cbstride, x[cb]);
  const long long vox_out = vox_in * 8;
#pragma unroll
  for (int qq = 0; qq < NPAR; ++qq) {
    float acc[DC_CO];
#pragma unroll
    for (int j = 0; j < DC_CO; ++j) acc[j] = 0.f;
#pragma unroll
    for (int cb = 0; cb < NCB; ++cb)
#pragma unroll
      for (int j = 0; j < 8; ++j)
#pragma unroll
        for (int t = 0; t < DC_CO; ++t) acc[t] = fmaf(x[cb][j], p.w[qq][cb * 8 + j][t], acc[t]);
    const int q = p.q0 + qq;
    const int i = q >> 2, j2 = (q >> 1) & 1, l = q & 1;
    const long long vo = (static_cast<long long>(2 * d + i) * (2 * p.H) + (2 * h + j2)) * (2 * p.W) + (2 * w + l);
#pragma unroll
    for (int b = 0; b < DC_CO / 8; ++b) {
      if (p.co0 + b * 8 >= p.cout) break;
      float y[8];
#pragma unroll
      for (int t = 0; t < 8; ++t) y[t] = acc[b * 8 + t];
      store8p(p.out_hi, p.out_lo, ((static_cast<size_t>(n) * p.out_cb_total + p.out_cb_off + (p.co0 >> 3) + b) * vox_out + vo) * 8, y);
    }
  }
}

// ------------------------------------------------------------------ trilinear 2x, align_corners=True (c3d.py:36)
// A thread owns source position (jd, jh, k) and produces the 2 x 2 x 2 output voxels
// (2jd + pd, 2jh + ph, 2k + o).  With align_corners the source index of output 2j + p lies in [j - 1 + p, j + p], so those four
// output rows need only the 3 x 3 source rows (jd - 1 .. jd + 1) x (jh - 1 .. jh + 1) at column k: 9 gathers for four rows
// instead of 16, 9 independent loads in flight per thread, and the index arithmetic / barrier amortised over 256 output bytes
// per thread.  Separable: the (d, h) blend is done per source column (results parked in shared memory), then the columns
// k-1, k, k+1 are combined into the two W-neighbours wo = 2k, 2k+1 (the source column of output 2k lies in (k-1, k], that of
// 2k+1 in [k, k+1/2)); the pair is written as one 32-byte sector per plane.  Interpolation weights are
// l = s * o - (j - 1 + p) against the clamped neighbours (identical to torch's floor-based form except exactly AT a knot,
// where both give the knot's value).  History: 8 corner gathers per output voxel ran at 24 % of the HBM bandwidth, one output
// row per thread (4 gathers per source column) at 38 %, this version at ~60 % (1.24 -> 0.81 ms per batch-8 step).
// Block = 256 threads = TR source rows x UP_COLS source columns (UP_COLS = min(64, W rounded up to a power of two));
// grid: x = (row group, column chunk), y = jd, z = (image, channel block).
template <int UP_COLS>
__global__ void __launch_bounds__(256, 2)
upsample2x_quad_kernel(const __half* __restrict__ in_hi, const __half* __restrict__ in_lo, int in_cb_total, int in_cb_off,
                       int ncb, int D, int H, int W, __half* out_hi, __half* out_lo, int out_cb_total, int out_cb_off) {
  constexpr int TR = 256 / UP_COLS;
  __shared__ float4 sm[TR][4][2][UP_COLS + 2];          // [thread row][pd * 2 + ph][channel half][column + 1]
  const int Do = 2 * D, Ho = 2 * H, Wo = 2 * W;
  const int kk = threadIdx.x & (UP_COLS - 1), r = threadIdx.x / UP_COLS;
  const int chunks = (W + UP_COLS - 1) / UP_COLS;
  const int k0 = (blockIdx.x % chunks) * UP_COLS, jh = (blockIdx.x / chunks) * TR + r;
  const int k = k0 + kk;
  const int jd = blockIdx.y;
  const int cb = blockIdx.z % ncb, n = blockIdx.z / ncb;
  const bool active = jh < H && k < W;
  const float sd = Do > 1 ? static_cast<float>(D - 1) / static_cast<float>(Do - 1) : 0.f;
  const float sh = Ho > 1 ? static_cast<float>(H - 1) / static_cast<float>(Ho - 1) : 0.f;
  const float sw = Wo > 1 ? static_cast<float>(W - 1) / static_cast<float>(Wo - 1) : 0.f;
  const size_t vox_i = static_cast<size_t>(D) * H * W;
  if (active) {
    float ld[2], lh[2];
#pragma unroll
    for (int pp = 0; pp < 2; ++pp) {
      ld[pp] = sd * (2 * jd + pp) - static_cast<float>(jd - 1 + pp);
      lh[pp] = sh * (2 * jh + pp) - static_cast<float>(jh - 1 + pp);
    }
    const size_t base = (static_cast<size_t>(n) * in_cb_total + in_cb_off + cb) * vox_i;
    size_t rowoff[3][3];
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
      for (int j = 0; j < 3; ++j)
        rowoff[i][j] = base + (static_cast<size_t>(min(max(jd - 1 + i, 0), D - 1)) * H + min(max(jh - 1 + j, 0), H - 1)) * W;
    // corner weights of source slot i for output parity p: slot p carries 1 - l, slot p + 1 carries l
    float wd[2][3], wh[2][3];
#pragma unroll
    for (int pp = 0; pp < 2; ++pp)
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        wd[pp][i] = i == pp ? 1.f - ld[pp] : (i == pp + 1 ? ld[pp] : 0.f);
        wh[pp][i] = i == pp ? 1.f - lh[pp] : (i == pp + 1 ? lh[pp] : 0.f);
      }
    auto blend = [&](int col, int slot) {
      float a[4][8];
#pragma unroll
      for (int q = 0; q < 4; ++q)
#pragma unroll
        for (int c = 0; c < 8; ++c) a[q][c] = 0.f;
#pragma unroll
      for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) {
          float v[8];
          load8(in_hi, in_lo, (rowoff[i][j] + col) * 8, v);
#pragma unroll
          for (int pd = 0; pd < 2; ++pd)
#pragma unroll
            for (int ph = 0; ph < 2; ++ph) {
              if (i < pd || i > pd + 1 || j < ph || j > ph + 1) continue;        // compile-time: this slot is no corner
              const float wgt = wd[pd][i] * wh[ph][j];
#pragma unroll
              for (int c = 0; c < 8; ++c) a[pd * 2 + ph][c] = fmaf(wgt, v[c], a[pd * 2 + ph][c]);
            }
        }
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        sm[r][q][0][slot] = make_float4(a[q][0], a[q][1], a[q][2], a[q][3]);
        sm[r][q][1][slot] = make_float4(a[q][4], a[q][5], a[q][6], a[q][7]);
      }
    };
    blend(k, kk + 1);
    // halo columns of the chunk: a clamped neighbour IS this thread's own column (no second gather)
    if (kk == 0) {
      if (k == 0) {
#pragma unroll
        for (int q = 0; q < 4; ++q) { sm[r][q][0][0] = sm[r][q][0][1]; sm[r][q][1][0] = sm[r][q][1][1]; }
      } else {
        blend(k - 1, 0);
      }
    }
    if (kk == UP_COLS - 1 || k == W - 1) {
      if (k == W - 1) {
#pragma unroll
        for (int q = 0; q < 4; ++q) { sm[r][q][0][kk + 2] = sm[r][q][0][kk + 1]; sm[r][q][1][kk + 2] = sm[r][q][1][kk + 1]; }
      } else {
        blend(k + 1, kk + 2);
      }
    }
  }
  __syncthreads();
  if (!active) return;
  float cw[2][3];
#pragma unroll
  for (int o = 0; o < 2; ++o) {
    const float fw = sw * (2 * k + o);
    const int w0 = static_cast<int>(fw);
    const float lw = fw - w0;
    const int pos = w0 - (k - 1);                       // 0 or 1
    cw[o][0] = pos == 0 ? 1.f - lw : 0.f;
    cw[o][1] = pos == 0 ? lw : 1.f - lw;
    cw[o][2] = pos == 0 ? 0.f : lw;
  }
  const size_t vox_o = vox_i * 8;
  const size_t obase = (static_cast<size_t>(n) * out_cb_total + out_cb_off + cb) * vox_o;
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    float acc[2][8];
#pragma unroll
    for (int o = 0; o < 2; ++o)
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[o][j] = 0.f;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float4 lo4 = sm[r][q][0][kk + c], hi4 = sm[r][q][1][kk + c];
      const float x[8] = {lo4.x, lo4.y, lo4.z, lo4.w, hi4.x, hi4.y, hi4.z, hi4.w};
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        acc[0][j] = fmaf(cw[0][c], x[j], acc[0][j]);
        acc[1][j] = fmaf(cw[1][c], x[j], acc[1][j]);
      }
    }
    const size_t o0 = (obase + (static_cast<size_t>(2 * jd + (q >> 1)) * Ho + (2 * jh + (q & 1))) * Wo + 2 * k) * 8;
    uint32_t hi[8], lo[8];
#pragma unroll
    for (int o = 0; o < 2; ++o)
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const __half2 h2 = __floats2half2_rn(acc[o][2 * j], acc[o][2 * j + 1]);
        const float2 f = __half22float2(h2);
        const __half2 l2 = __floats2half2_rn(acc[o][2 * j] - f.x, acc[o][2 * j + 1] - f.y);
        hi[o * 4 + j] = *reinterpret_cast<const uint32_t*>(&h2);
        lo[o * 4 + j] = *reinterpret_cast<const uint32_t*>(&l2);
      }
    st_global_v8u(out_hi + o0, hi);
    if (out_lo != nullptr) st_global_v8u(out_lo + o0, lo);
  }
}

// ------------------------------------------------------------------ LayerNorm over the last dim (one warp per row)
__global__ void __launch_bounds__(256)
layernorm_kernel(const float* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ beta, int rows,
                 int cols, __half* out_f16, float* out_f32) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const float* xr = x + static_cast<size_t>(row) * cols;
  float s = 0.f;
  for (int c = lane; c < cols; c += 32) s += xr[c];
  const float mean = warp_sum(s) / cols;
  float ss = 0.f;
  for (int c = lane; c < cols; c += 32) { const float d = xr[c] - mean; ss = fmaf(d, d, ss); }
  const float rstd = rsqrtf(warp_sum(ss) / cols + 1e-5f);
  for (int c = lane; c < cols; c += 32) {
    const float y = fmaf((xr[c] - mean) * rstd, gamma[c], beta[c]);
    if (out_f16) out_f16[static_cast<size_t>(row) * cols + c] = __float2half_rn(y);
    if (out_f32) out_f32[static_cast<size_t>(row) * cols + c] = y;
  }
}

// ------------------------------------------------------------------ deterministic split-K finish
__global__ void __launch_bounds__(256)
splitk_reduce_kernel(const float* __restrict__ ws, int splits, long long MN, int N, const float* __restrict__ bias,
                     const float* __restrict__ rowvec, int row_period, float* out) {
  const long long i = (blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x) * 4;
  if (i >= MN) return;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int s = 0; s < splits; ++s) {
    const float4 v = *reinterpret_cast<const float4*>(ws + s * MN + i);
    acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
  }
  const int n = static_cast<int>(i % N);
  const long long m = i / N;
  if (bias) { acc.x += bias[n]; acc.y += bias[n + 1]; acc.z += bias[n + 2]; acc.w += bias[n + 3]; }
  if (rowvec) {
    const float* r = rowvec + (m % row_period) * N + n;
    acc.x += r[0]; acc.y += r[1]; acc.z += r[2]; acc.w += r[3];
  }
  *reinterpret_cast<float4*>(out + i) = acc;
}

// ------------------------------------------------------------------ row softmax fp32 -> fp16 (one warp per row)
__global__ void __launch_bounds__(256)
softmax_kernel(const float* __restrict__ s, int rows, int cols, int ld_in, __half* p, int ld_out) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const float* sr = s + static_cast<size_t>(row) * ld_in;
  float mx = -INFINITY;
  for (int c = lane; c < cols; c += 32) mx = fmaxf(mx, sr[c]);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  float sum = 0.f;
  for (int c = lane; c < cols; c += 32) sum += __expf(sr[c] - mx);
  const float inv = 1.f / warp_sum(sum);
  for (int c = lane; c < cols; c += 32) p[static_cast<size_t>(row) * ld_out + c] = __float2half_rn(__expf(sr[c] - mx) * inv);
}

// ------------------------------------------------------------------ 16^3 patch flattening for the "perceptron" embedding
// einops 'b c (h p1)(w p2)(d p3) -> b (h w d)(p1 p2 p3 c)' restated for the c8 layout: K order is
// (c8-block, p1, p2, p3, c%8); the Linear weight is permuted identically when packed.
__global__ void __launch_bounds__(256)
patchify_kernel(const __half* __restrict__ in, int cb_total, int cb_off, int ncb, int S0, int S1, int S2, __half* out) {
  const int g0 = S0 / 16, g1 = S1 / 16, g2 = S2 / 16;
  // 32-bit index arithmetic (the host checks the range).  The copy is bandwidth bound: 2 x 1.07 GB per batch-8 step of
  // the dose ViT in 0.45 ms = 73 % of the HBM copy bandwidth.
  const unsigned total = static_cast<unsigned>(g0) * g1 * g2 * ncb * 4096u;
  const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int n = blockIdx.y;
  unsigned t = i;
  const int p3 = static_cast<int>(t & 15u); t >>= 4;
  const int p2 = static_cast<int>(t & 15u); t >>= 4;
  const int p1 = static_cast<int>(t & 15u); t >>= 4;
  const int cb = static_cast<int>(t % static_cast<unsigned>(ncb)); t /= static_cast<unsigned>(ncb);
  const int tok = static_cast<int>(t);
  const int gz = tok % g2, gy = (tok / g2) % g1, gx = tok / (g2 * g1);
  const size_t vox = (static_cast<size_t>(gx * 16 + p1) * S1 + (gy * 16 + p2)) * S2 + (gz * 16 + p3);
  const size_t vol = static_cast<size_t>(S0) * S1 * S2;
  const uint4 val = *reinterpret_cast<const uint4*>(in + ((static_cast<size_t>(n) * cb_total + cb_off + cb) * vol + vox) * 8);
  const size_t K = static_cast<size_t>(ncb) * 4096 * 8;
  const size_t ntok = static_cast<size_t>(g0) * g1 * g2;
  *reinterpret_cast<uint4*>(out + (static_cast<size_t>(n) * ntok + tok) * K + ((static_cast<size_t>(cb) * 16 + p1) * 16 + p2) * 128 + p3 * 8) = val;
}

// The same flatten for a ONE-channel planar fp32 volume (the seg net's CT input): K = 4096 = (p1, p2, p3) exactly as the
// Linear weight stores it, instead of the c8 path's K = 8 * 4096 with seven zero channels.  One thread = 8 consecutive p3.
__global__ void __launch_bounds__(256)
patchify_planar_kernel(const float* __restrict__ in, int S0, int S1, int S2, __half* out) {
  const int g1 = S1 / 16, g2 = S2 / 16;
  const unsigned total = static_cast<unsigned>(S0 / 16) * g1 * g2 * 512u;          // 4096 / 8 vectors per token
  const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int n = blockIdx.y;
  unsigned t = i;
  const int p3h = static_cast<int>(t & 1u); t >>= 1;
  const int p2 = static_cast<int>(t & 15u); t >>= 4;
  const int p1 = static_cast<int>(t & 15u); t >>= 4;
  const int tok = static_cast<int>(t);
  const int gz = tok % g2, gy = (tok / g2) % g1, gx = tok / (g2 * g1);
  const size_t vol = static_cast<size_t>(S0) * S1 * S2;
  const size_t vox = (static_cast<size_t>(gx * 16 + p1) * S1 + (gy * 16 + p2)) * S2 + (gz * 16 + p3h * 8);
  float x[8];
  ld_global_v8f(in + static_cast<size_t>(n) * vol + vox, x);
  __align__(16) __half2 h[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) h[j] = __floats2half2_rn(x[2 * j], x[2 * j + 1]);
  const size_t ntok = static_cast<size_t>(S0 / 16) * g1 * g2;
  *reinterpret_cast<uint4*>(out + (static_cast<size_t>(n) * ntok + tok) * 4096 + (p1 * 16 + p2) * 16 + p3h * 8) = *reinterpret_cast<const uint4*>(h);
}

// ------------------------------------------------------------------ cascade hand-off (train_light_linked_model.py:156-167)
// logits [N,8,X,Y,Z] fp32 -> argmax (first maximum wins) -> 7 one-hot OAR masks, spatially transposed
// (x,y,z)->(z,y,x), + PTV (not transposed) + CT (transposed) -> dose-net input, written both as the c8
// fp16 hi/lo tensor the dose net consumes and (optionally) as NCDHW fp32 "structures".
__global__ void __launch_bounds__(256)
handoff_kernel(const float* __restrict__ logits, int ncls, const float* __restrict__ ptv, const float* __restrict__ ct,
               int S, __half* out_hi, __half* out_lo, int out_cb_total, int out_cb_off, float* structures) {
  // tile over (x, z) for fixed y: source index [x][y][z] (z fastest); destination voxel (a=z, b=y, c=x) (c fastest)
  __shared__ uint8_t cls[32][33];
  __shared__ float ctv[32][33];
  const int n = blockIdx.z / S, y = blockIdx.z % S;
  const int x0 = blockIdx.y * 32, z0 = blockIdx.x * 32;
  const size_t vol = static_cast<size_t>(S) * S * S;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;   // 32 x 8
  for (int r = ty; r < 32; r += 8) {
    const int x = x0 + r, z = z0 + tx;
    if (x < S && z < S) {
      const size_t v = (static_cast<size_t>(x) * S + y) * S + z;
      float best = logits[(static_cast<size_t>(n) * ncls) * vol + v];
      int arg = 0;
      for (int c = 1; c < ncls; ++c) {
        const float l = logits[(static_cast<size_t>(n) * ncls + c) * vol + v];
        if (l > best) { best = l; arg = c; }
      }
      cls[r][tx] = static_cast<uint8_t>(arg);
      ctv[r][tx] = ct[static_cast<size_t>(n) * vol + v];
    }
  }
  __syncthreads();
  for (int r = ty; r < 32; r += 8) {
    const int a = z0 + r, c = x0 + tx;          // destination (a, y, c); source (x=c, y, z=a)
    if (a < S && c < S) {
      const size_t vo = (static_cast<size_t>(a) * S + y) * S + c;
      const int k = cls[tx][r];
      float ch[16];
      ch[0] = ptv[static_cast<size_t>(n) * vol + vo];
#pragma unroll
      for (int j = 1; j < 8; ++j) ch[j] = (k == j) ? 1.f : 0.f;
      ch[8] = ctv[tx][r];
#pragma unroll
      for (int j = 9; j < 16; ++j) ch[j] = 0.f;
      float lo8[8], hi8[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) { lo8[j] = ch[j]; hi8[j] = ch[8 + j]; }
      store8(out_hi, out_lo, ((static_cast<size_t>(n) * out_cb_total + out_cb_off) * vol + vo) * 8, lo8);
      store8(out_hi, out_lo, ((static_cast<size_t>(n) * out_cb_total + out_cb_off + 1) * vol + vo) * 8, hi8);
      if (structures) {
#pragma unroll
        for (int j = 0; j < 9; ++j) structures[(static_cast<size_t>(n) * 9 + j) * vol + vo] = ch[j];
      }
    }
  }
}

// ------------------------------------------------------------------ sliding-window inference plumbing
// monai.inferers.sliding_window_inference (constant blending) as called at train_light_linked_model.py:152-154:
// crop ROI windows out of the volume, run the predictor, add every window's logits into the full-size sum in
// window order (deterministic, like the reference's sequential +=), divide by the per-voxel window count.
struct WindowList {
  int n;                      // windows handled by this launch
  int b[64], x0[64], y0[64], z0[64];
};
__global__ void __launch_bounds__(256)
crop_pack_kernel(const float* __restrict__ src, int C, int S0, int S1, int S2, int R, const WindowList wl, __half* hi,
                 __half* lo, int cb_total, int cb_off, int ncb) {
  const long long rv = static_cast<long long>(R) * R * R;
  const long long v = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (v >= rv) return;
  const int cb = blockIdx.y % ncb, wi = blockIdx.y / ncb;
  const int z = static_cast<int>(v % R), y = static_cast<int>((v / R) % R), x = static_cast<int>(v / (static_cast<long long>(R) * R));
  const size_t vol = static_cast<size_t>(S0) * S1 * S2;
  const size_t sv = (static_cast<size_t>(wl.x0[wi] + x) * S1 + (wl.y0[wi] + y)) * S2 + (wl.z0[wi] + z);
  float xv[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int c = cb * 8 + j;
    xv[j] = c < C ? src[(static_cast<size_t>(wl.b[wi]) * C + c) * vol + sv] : 0.f;
  }
  store8(hi, lo, ((static_cast<size_t>(wi) * cb_total + cb_off + cb) * rv + v) * 8, xv);
}
// out[b, c, window region] += win[w, c, :, :, :] for ONE window per volume per launch (no write conflicts)
__global__ void __launch_bounds__(256)
window_add_kernel(const float* __restrict__ win, int ncls, int R, const WindowList wl, int win_stride, float* out, int S0,
                  int S1, int S2, int first) {
  const long long rv = static_cast<long long>(R) * R * R;
  const long long v = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (v >= rv) return;
  const int c = blockIdx.y % ncls, wi = blockIdx.y / ncls;
  const int z = static_cast<int>(v % R), y = static_cast<int>((v / R) % R), x = static_cast<int>(v / (static_cast<long long>(R) * R));
  const size_t vol = static_cast<size_t>(S0) * S1 * S2;
  const size_t ov = (static_cast<size_t>(wl.x0[wi] + x) * S1 + (wl.y0[wi] + y)) * S2 + (wl.z0[wi] + z);
  float* o = out + (static_cast<size_t>(wl.b[wi]) * ncls + c) * vol + ov;
  const float val = win[(static_cast<size_t>(wi) * win_stride + c) * rv + v];
  *o = *o + val;
}
__global__ void __launch_bounds__(256)
div_count_kernel(float* __restrict__ data, const float* __restrict__ count, long long vol, int rows) {
  const long long v = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (v >= vol) return;
  const float c = count[v];
  for (int r = blockIdx.y; r < rows; r += gridDim.y) data[static_cast<size_t>(r) * vol + v] = data[static_cast<size_t>(r) * vol + v] / c;
}

// ------------------------------------------------------------------ generic direct convolution (strided convs)
// out[n,co,do,ho,wo] = bias + sum_{ci,kd,kh,kw} in[n,ci,do*s+kd*dil-pad,...] * w[co,ci,kd,kh,kw]
// One thread = one output voxel x DC2_CO output channels, fp32 math on fp16 hi(+lo) c8 inputs.
struct DirectConvParams {
  const __half* in_hi; const __half* in_lo; int in_cb_total, in_cb_off, cin;
  int D, H, W, Do, Ho, Wo, k, stride, dil, pad;
  const float* w;        // packed [tap][cin][cout] fp32
  const float* scale; const float* shift; int relu; int cout;
  float* out_raw; __half* out_hi; __half* out_lo; int out_cb_total, out_cb_off;
  double* stats_out;
};
constexpr int DC2_CO = 16;
constexpr int DC2_CI = 32;   // input channels staged per shared-memory weight tile
__global__ void __launch_bounds__(128) direct_conv_kernel(const DirectConvParams p) {
  extern __shared__ float wsm[];   // [taps][DC2_CI][DC2_CO]
  const int co0 = blockIdx.y * DC2_CO, n = blockIdx.z;
  const long long vox_o = static_cast<long long>(p.Do) * p.Ho * p.Wo, vox_i = static_cast<long long>(p.D) * p.H * p.W;
  const long long v = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  const bool valid = v < vox_o;
  const int wo = static_cast<int>(v % p.Wo), ho = static_cast<int>((v / p.Wo) % p.Ho), dd = static_cast<int>(v / (static_cast<long long>(p.Wo) * p.Ho));
  const int taps = p.k * p.k * p.k;
  float acc[DC2_CO];
#pragma unroll
  for (int j = 0; j < DC2_CO; ++j) acc[j] = 0.f;
  for (int ci0 = 0; ci0 < p.cin; ci0 += DC2_CI) {
    const int nci = min(DC2_CI, p.cin - ci0);
    __syncthreads();
    for (int i = threadIdx.x; i < taps * nci * DC2_CO; i += blockDim.x) {
      const int j = i % DC2_CO, ci = (i / DC2_CO) % nci, tap = i / (DC2_CO * nci);
      wsm[(tap * DC2_CI + ci) * DC2_CO + j] =
          (co0 + j < p.cout) ? p.w[(static_cast<size_t>(tap) * p.cin + ci0 + ci) * p.cout + co0 + j] : 0.f;
    }
    __syncthreads();
    if (!valid) continue;
    for (int kd = 0; kd < p.k; ++kd) {
      const int z = dd * p.stride + kd * p.dil - p.pad;
      if (z < 0 || z >= p.D) continue;
      for (int kh = 0; kh < p.k; ++kh) {
        const int y = ho * p.stride + kh * p.dil - p.pad;
        if (y < 0 || y >= p.H) continue;
        for (int kw = 0; kw < p.k; ++kw) {
          const int x = wo * p.stride + kw * p.dil - p.pad;
          if (x < 0 || x >= p.W) continue;
          const int tap = (kd * p.k + kh) * p.k + kw;
          const size_t vi = (static_cast<size_t>(z) * p.H + y) * p.W + x;
          for (int cb = 0; cb < nci / 8; ++cb) {
            float xin[8];
            load8(p.in_hi, p.in_lo, ((static_cast<size_t>(n) * p.in_cb_total + p.in_cb_off + (ci0 >> 3) + cb) * vox_i + vi) * 8, xin);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float4* wr = reinterpret_cast<const float4*>(&wsm[(tap * DC2_CI + cb * 8 + j) * DC2_CO]);
#pragma unroll
              for (int t = 0; t < DC2_CO / 4; ++t) {
                const float4 w4 = wr[t];
                acc[4 * t + 0] = fmaf(xin[j], w4.x, acc[4 * t + 0]);
                acc[4 * t + 1] = fmaf(xin[j], w4.y, acc[4 * t + 1]);
                acc[4 * t + 2] = fmaf(xin[j], w4.z, acc[4 * t + 2]);
                acc[4 * t + 3] = fmaf(xin[j], w4.w, acc[4 * t + 3]);
              }
            }
          }
        }
      }
    }
  }
#pragma unroll
  for (int j = 0; j < DC2_CO; ++j) {
    if (co0 + j < p.cout) {
      acc[j] = fmaf(acc[j], __ldg(&p.scale[co0 + j]), __ldg(&p.shift[co0 + j]));
      if (p.relu) acc[j] = fmaxf(acc[j], 0.f);
    }
  }
#pragma unroll
  for (int b = 0; b < DC2_CO / 8; ++b) {
    if (co0 + b * 8 >= p.cout) break;
    float y[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) y[j] = acc[b * 8 + j];
    if (valid) {
      const size_t off = ((static_cast<size_t>(n) * p.out_cb_total + p.out_cb_off + (co0 >> 3) + b) * vox_o + v) * 8;
      if (p.out_raw) {
        st_global_v8f(p.out_raw + off, y);
      }
      if (p.out_hi) store8(p.out_hi, p.out_lo, off, y);
    }
    if (p.stats_out) {
      block_accumulate_stats(y, valid, p.stats_out, static_cast<size_t>(n) * p.cout + co0 + b * 8);
      __syncthreads();
    }
  }
}

}  // namespace dp

// =============================================================================== C ABI wrappers
using namespace dp;

static inline unsigned blocks_for(long long n, int threads) { return static_cast<unsigned>((n + threads - 1) / threads); }

extern "C" int dp_pack_ncdhw(const float* src, int N, int C, long long vox, void* hi, void* lo, int cb_total, int cb_off,
                             cudaStream_t stream) {
  const int ncb = (C + 7) / 8;
  DP_REQUIRE(N > 0 && C > 0 && vox > 0, "dp_pack_ncdhw: empty tensor");
  dim3 grid(blocks_for(vox, 256), N * ncb);
  pack_ncdhw_kernel<<<grid, 256, 0, stream>>>(src, C, vox, static_cast<__half*>(hi), static_cast<__half*>(lo), cb_total, cb_off, ncb);
  DP_CHECK(cudaGetLastError());
  return 0;
}

extern "C" int dp_unpack_c8(const void* hi, const void* lo, int cb_total, int cb_off, int N, int C, long long vox,
                            float* dst, cudaStream_t stream) {
  const int ncb = (C + 7) / 8;
  dim3 grid(blocks_for(vox, 256), N * ncb);
  unpack_c8_kernel<<<grid, 256, 0, stream>>>(static_cast<const __half*>(hi), static_cast<const __half*>(lo), cb_total, cb_off, C, vox, dst);
  DP_CHECK(cudaGetLastError());
  return 0;
}

extern "C" int dp_norm_act(const float* raw_f32, const void* raw_hi, const void* raw_lo, int in_cb_total, int in_cb_off,
                           const double* stats, const float* gamma, const float* beta, int act, const void* res_hi,
                           const void* res_lo, const float* res_raw, const double* res_stats, int res_cb_total,
                           int res_cb_off, int act_after_res, void* out_hi, void* out_lo, int out_cb_total,
                           int out_cb_off, double* stats_out, int N, int C, long long vox, void* s2d_hi, void* s2d_lo,
                           int s2d_cb_total, int s2d_cb_off, int D, int H, int W, cudaStream_t stream) {
  DP_REQUIRE(raw_f32 != nullptr || raw_hi != nullptr, "dp_norm_act: no input");
  DP_REQUIRE(s2d_hi == nullptr || (D % 2 == 0 && H % 2 == 0 && W % 2 == 0 && static_cast<long long>(D) * H * W == vox && C % 8 == 0),
             "dp_norm_act: space-to-depth output needs even dims and C %% 8 == 0");
  NormActParams p{};
  p.raw_f32 = raw_f32; p.raw_hi = static_cast<const __half*>(raw_hi); p.raw_lo = static_cast<const __half*>(raw_lo);
  p.in_cb_total = in_cb_total; p.in_cb_off = in_cb_off; p.stats = stats; p.gamma = gamma; p.beta = beta; p.act = act;
  p.res_hi = static_cast<const __half*>(res_hi); p.res_lo = static_cast<const __half*>(res_lo); p.res_raw = res_raw;
  p.res_stats = res_stats; p.res_cb_total = res_cb_total; p.res_cb_off = res_cb_off; p.act_after_res = act_after_res;
  p.out_hi = static_cast<__half*>(out_hi); p.out_lo = static_cast<__half*>(out_lo); p.out_cb_total = out_cb_total;
  p.out_cb_off = out_cb_off; p.stats_out = stats_out; p.C = C; p.ncb = (C + 7) / 8; p.vox = vox; p.inv_vox = 1.0 / static_cast<double>(vox);
  p.s2d_hi = static_cast<__half*>(s2d_hi); p.s2d_lo = static_cast<__half*>(s2d_lo); p.s2d_cb_total = s2d_cb_total;
  p.s2d_cb_off = s2d_cb_off; p.D = D; p.H = H > 0 ? H : 1; p.W = W > 0 ? W : 1;
  dim3 grid(blocks_for(vox, 256 * NA_IT), N * p.ncb);
  norm_act_kernel<<<grid, 256, 0, stream>>>(p);
  DP_CHECK(cudaGetLastError());
  return 0;
}

extern "C" int dp_norm_act_resx(const float* raw_f32, int in_cb_total, const double* stats, const float* res_x, const float* res_w,
                                const double* res_xstats, int act_after_res, void* out_hi, void* out_lo, int out_cb_total,
                                int out_cb_off, int N, int C, long long vox, cudaStream_t stream) {
  DP_REQUIRE(raw_f32 && stats && res_x && res_w && res_xstats && out_hi, "dp_norm_act_resx: missing operand");
  NormActParams p{};
  p.raw_f32 = raw_f32; p.in_cb_total = in_cb_total; p.stats = stats; p.act = ACT_NONE;
  p.res_x = res_x; p.res_w = res_w; p.res_xstats = res_xstats; p.act_after_res = act_after_res;
  p.out_hi = static_cast<__half*>(out_hi); p.out_lo = static_cast<__half*>(out_lo); p.out_cb_total = out_cb_total;
  p.out_cb_off = out_cb_off; p.C = C; p.ncb = (C + 7) / 8; p.vox = vox; p.inv_vox = 1.0 / static_cast<double>(vox);
  p.H = 1; p.W = 1;
  dim3 grid(blocks_for(vox, 256 * NA_IT), N * p.ncb);
  norm_act_kernel<<<grid, 256, 0, stream>>>(p);
  DP_CHECK(cudaGetLastError());
  return 0;
}

extern "C" int dp_norm_act_head(const float* raw_f32, int in_cb_total, const double* stats, const float* gamma, const float* beta,
                                int act, void* out_hi, void* out_lo, int out_cb_total, int out_cb_off, const float* head_w,
                                const float* head_b, int head_co, float* head_out, int N, int C, long long vox,
                                cudaStream_t stream) {
  DP_REQUIRE(raw_f32 != nullptr && stats != nullptr && head_w != nullptr && head_out != nullptr, "dp_norm_act_head: missing operand");
  DP_REQUIRE(C >= 1 && C <= NH_MAXC && head_co >= 1 && head_co <= NH_MAXCO, "dp_norm_act_head: C=%d (<= %d), head_co=%d (<= %d)",
             C, NH_MAXC, head_co, NH_MAXCO);
  NormHeadParams p{};
  p.raw_f32 = raw_f32; p.in_cb_total = in_cb_total; p.stats = stats; p.gamma = gamma; p.beta = beta; p.act = act;
  p.out_hi = static_cast<__half*>(out_hi); p.out_lo = static_cast<__half*>(out_lo); p.out_cb_total = out_cb_total;
  p.out_cb_off = out_cb_off; p.head_w = head_w; p.head_b = head_b; p.head_co = head_co; p.head_out = head_out;
  p.C = C; p.ncb = (C + 7) / 8; p.vox = vox; p.inv_vox = 1.0 / static_cast<double>(vox);
  dim3 grid(blocks_for(vox, 256), N);
  if (head_co == 1) norm_act_head_kernel<1><<<grid, 256, 0, stream>>>(p);
  else norm_act_head_kernel<NH_MAXCO><<<grid, 256, 0, stream>>>(p);
  DP_CHECK(cudaGetLastError());
  return 0;
}

extern "C" int dp_pointwise_conv(int nsrc, const void* const* src_hi, const void* const* src_lo, const float* const* src_raw,
                                 const int* src_cb_total, const int* src_cb_off, const int* src_C,
                                 const double* const* src_stats, const int* src_act, const float* w, const float* bias,
                                 int cout, int N, long long vox, float* out_raw, void* out_hi, void* out_lo,
                                 int out_cb_total, int out_cb_off, float* out_planar, double* stats_out, int out_act,
                                 cudaStream_t stream) {
  DP_REQUIRE(nsrc >= 1 && nsrc <= 3, "dp_pointwise_conv: 1..3 sources supported, got %d", nsrc);
  PointwiseParams p{};
  p.nsrc = nsrc;
  int cin = 0;
  for (int s = 0; s < nsrc; ++s) {
    p.src[s].hi = static_cast<const __half*>(src_hi[s]);
    p.src[s].lo = static_cast<const __half*>(src_lo ? src_lo[s] : nullptr);
    p.src[s].raw = src_raw ? src_raw[s] : nullptr;
    p.src[s].cb_total = src_cb_total[s]; p.src[s].cb_off = src_cb_off[s]; p.src[s].C = src_C[s];
    p.src[s].stats = src_stats ? src_stats[s] : nullptr; p.src[s].act = src_act ? src_act[s] : 0;
    DP_REQUIRE(p.src[s].hi || p.src[s].raw, "dp_pointwise_conv: source %d has no tensor", s);
    cin += src_C[s];
  }
  p.w = w; p.bias = bias; p.cin_total = cin; p.cout = cout; p.vox = vox; p.inv_vox = 1.0 / static_cast<double>(vox);
  p.out_raw = out_raw; p.out_hi = static_cast<__half*>(out_hi); p.out_lo = static_cast<__half*>(out_lo);
  p.out_cb_total = out_cb_total; p.out_cb_off = out_cb_off; p.out_planar = out_planar; p.stats_out = stats_out;
  p.out_act = out_act;
  int cin_pad = 0;
  for (int s = 0; s < nsrc; ++s) cin_pad += ((src_C[s] + 7) / 8) * 8;
  const size_t smem = static_cast<size_t>(cin_pad) * (PW_CO + 2) * sizeof(float);
  DP_REQUIRE(smem <= 96 * 1024, "dp_pointwise_conv: C_in=%d too large", cin);
  static size_t configured = 48 * 1024;
  if (smem > configured) {
    DP_CHECK(cudaFuncSetAttribute(pointwise_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
    configured = 96 * 1024;
  }
  dim3 grid(blocks_for(vox, 128 * PW_IT), (cout + PW_CO - 1) / PW_CO, N);
  pointwise_kernel<<<grid, 128, smem, stream>>>(p);
  DP_CHECK(cudaGetLastError());
  return 0;
}

template <int NCB>
static int launch_pointwise_cw(int nsrc, const void* const* src_hi, const void* const* src_lo, const float* const* src_raw,
                               const int* src_cb_total, const int* src_cb_off, const int* src_C,
                               const double* const* src_stats, const int* src_act, const float* w_host, const float* bias_host,
                               int cin, int cout, int N, long long vox, float* out_raw, void* out_hi, void* out_lo,
                               int out_cb_total, int out_cb_off, float* out_planar, double* stats_out, int out_act,
                               cudaStream_t stream) {
  PwConstParams<NCB> p{};
  int b = 0;
  int lbase[NCB];
  for (int s = 0; s < nsrc; ++s) {
    const int nb = (src_C[s] + 7) / 8;
    int l0 = 0;
    for (int t = 0; t < s; ++t) l0 += src_C[t];
    for (int cb = 0; cb < nb; ++cb, ++b) {
      PwBlock& B = p.blk[b];
      const size_t blk_off = static_cast<size_t>(src_cb_off[s] + cb) * vox * 8;
      B.hi = src_hi[s] ? static_cast<const __half*>(src_hi[s]) + blk_off : nullptr;
      B.lo = (src_lo && src_lo[s]) ? static_cast<const __half*>(src_lo[s]) + blk_off : nullptr;
      B.raw = (src_raw && src_raw[s]) ? src_raw[s] + blk_off : nullptr;
      B.n_stride = static_cast<long long>(src_cb_total[s]) * vox * 8;
      B.stats = src_stats ? src_stats[s] : nullptr;
      B.stat_c0 = cb * 8; B.stat_C = src_C[s];
      B.act = src_act ? src_act[s] : 0;
      lbase[b] = (cb * 8 < src_C[s]) ? l0 + cb * 8 : -1;
      // logical channels of this block: l0 + cb*8 .. min(l0 + C, ...) ; remember how many are real
      B.stat_C = src_C[s];
    }
  }
  p.cout = cout; p.vox = vox; p.inv_vox = 1.0 / static_cast<double>(vox);
  p.out_raw = out_raw; p.out_hi = static_cast<__half*>(out_hi); p.out_lo = static_cast<__half*>(out_lo);
  p.out_cb_total = out_cb_total; p.out_cb_off = out_cb_off; p.out_planar = out_planar; p.stats_out = stats_out;
  p.out_act = out_act;
  dim3 grid(blocks_for(vox, 128 * PW_IT), 1, N);
  for (int co0 = 0; co0 < cout; co0 += PW_CO) {
    p.co0 = co0;
    for (int bb = 0; bb < NCB; ++bb)
      for (int j = 0; j < 8; ++j) {
        const int valid_c = p.blk[bb].stat_C - p.blk[bb].stat_c0;           // real channels left in this block
        for (int q = 0; q < PW_CO; ++q)
          p.w[bb * 8 + j][q] = (j < valid_c && co0 + q < cout) ? w_host[static_cast<size_t>(co0 + q) * cin + lbase[bb] + j] : 0.f;
      }
    for (int q = 0; q < PW_CO; ++q) p.bias[q] = (bias_host && co0 + q < cout) ? bias_host[co0 + q] : 0.f;
    pointwise_cw_kernel<NCB><<<grid, 128, 0, stream>>>(p);
    DP_CHECK(cudaGetLastError());
  }
  return 0;
}

extern "C" int dp_pointwise_conv_cw(int nsrc, const void* const* src_hi, const void* const* src_lo, const float* const* src_raw,
                                    const int* src_cb_total, const int* src_cb_off, const int* src_C,
                                    const double* const* src_stats, const int* src_act, const float* w_host,
                                    const float* bias_host, int cout, int N, long long vox, float* out_raw, void* out_hi,
                                    void* out_lo, int out_cb_total, int out_cb_off, float* out_planar, double* stats_out,
                                    int out_act, cudaStream_t stream) {
  DP_REQUIRE(nsrc >= 1 && nsrc <= 3, "dp_pointwise_conv_cw: 1..3 sources supported, got %d", nsrc);
  int cin = 0, ncb = 0;
  for (int s = 0; s < nsrc; ++s) {
    DP_REQUIRE(src_hi[s] || (src_raw && src_raw[s]), "dp_pointwise_conv_cw: source %d has no tensor", s);
    cin += src_C[s];
    ncb += (src_C[s] + 7) / 8;
  }
#define DP_PW_CW(NCB_)                                                                                                      \
  return launch_pointwise_cw<NCB_>(nsrc, src_hi, src_lo, src_raw, src_cb_total, src_cb_off, src_C, src_stats, src_act, w_host, \
                                   bias_host, cin, cout, N, vox, out_raw, out_hi, out_lo, out_cb_total, out_cb_off,          \
                                   out_planar, stats_out, out_act, stream)
  switch (ncb) {
    case 1: DP_PW_CW(1);
    case 2: DP_PW_CW(2);
    case 3: DP_PW_CW(3);
    case 4: DP_PW_CW(4);
    case 6: DP_PW_CW(6);
    case 8: DP_PW_CW(8);
    default: break;
  }
#undef DP_PW_CW
  set_error("dp_pointwise_conv_cw: %d input channel blocks not instantiated (1,2,3,4,6,8); use dp_pointwise_conv", ncb);
  return 1;
}

extern "C" int dp_deconv2x(const void* in_hi, const void* in_lo, long long in_nstride, long long in_vstride,
                           long long in_cbstride, int cin, int cout, int N, int D, int H, int W, const float* w_packed,
                           void* out_hi, void* out_lo, int out_cb_total, int out_cb_off, cudaStream_t stream) {
  DP_REQUIRE(cin % 8 == 0, "dp_deconv2x: C_in=%d must be a multiple of 8", cin);
  DeconvParams p{};
  p.in_hi = static_cast<const __half*>(in_hi); p.in_lo = static_cast<const __half*>(in_lo);
  p.in_nstride = in_nstride; p.in_vstride = in_vstride; p.in_cbstride = in_cbstride;
  p.cin = cin; p.cout = cout; p.D = D; p.H = H; p.W = W; p.w = w_packed;
  p.out_hi = static_cast<__half*>(out_hi); p.out_lo = static_cast<__half*>(out_lo);
  p.out_cb_total = out_cb_total; p.out_cb_off = out_cb_off;
  const size_t smem = static_cast<size_t>(cin) * DC_CO * sizeof(float);
  static size_t configured = 48 * 1024;
  if (smem > configured) {
    DP_CHECK(cudaFuncSetAttribute(deconv2x_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
    configured = 96 * 1024;
  }
  DP_REQUIRE(smem <= 96 * 1024, "dp_deconv2x: C_in=%d too large", cin);
  const long long vox = static_cast<long long>(D) * H * W;
  dim3 grid(blocks_for(vox, 128), (cout + DC_CO - 1) / DC_CO, N);
  deconv2x_kernel<<<grid, 128, smem, stream>>>(p);
  DP_CHECK(cudaGetLastError());
  return 0;
}

template <int NCB, int NPAR>
static int launch_deconv_cw(const void* in_hi, const void* in_lo, long long in_nstride, long long in_cbstride, int cout, int N,
                            int D, int H, int W, const float* w_host, void* out_hi, void* out_lo, int out_cb_total,
                            int out_cb_off, cudaStream_t stream) {
  thread_local static DeconvConstParams<NCB, NPAR> p;      // 16 KB: kept off the stack; every launch copies it
  p.in_hi = static_cast<const __half*>(in_hi); p.in_lo = static_cast<const __half*>(in_lo);
  p.in_nstride = in_nstride; p.in_cbstride = in_cbstride; p.D = D; p.H = H; p.W = W; p.cout = cout;
  p.out_hi = static_cast<__half*>(out_hi); p.out_lo = static_cast<__half*>(out_lo);
  p.out_cb_total = out_cb_total; p.out_cb_off = out_cb_off;
  const int cin = NCB * 8;
  dim3 grid(blocks_for(static_cast<long long>(D) * H * W, 128), 1, N);
  for (int co0 = 0; co0 < cout; co0 += DC_CO)
    for (int q0 = 0; q0 < 8; q0 += NPAR) {
      p.co0 = co0; p.q0 = q0;
      for (int qq = 0; qq < NPAR; ++qq)
        for (int ci = 0; ci < cin; ++ci)
          for (int t = 0; t < DC_CO; ++t)
            p.w[qq][ci][t] = (co0 + t < cout) ? w_host[(static_cast<size_t>(q0 + qq) * cin + ci) * cout + co0 + t] : 0.f;
      deconv2x_cw_kernel<NCB, NPAR><<<grid, 128, 0, stream>>>(p);
      DP_CHECK(cudaGetLastError());
    }
  return 0;
}

extern "C" int dp_deconv2x_cw(const void* in_hi, const void* in_lo, long long in_nstride, long long in_cbstride, int cin,
                              int cout, int N, int D, int H, int W, const float* w_host_packed, void* out_hi, void* out_lo,
                              int out_cb_total, int out_cb_off, cudaStream_t stream) {
  if (cin == 32)
    return launch_deconv_cw<4, 8>(in_hi, in_lo, in_nstride, in_cbstride, cout, N, D, H, W, w_host_packed, out_hi, out_lo,
                                  out_cb_total, out_cb_off, stream);
  if (cin == 64)
    return launch_deconv_cw<8, 4>(in_hi, in_lo, in_nstride, in_cbstride, cout, N, D, H, W, w_host_packed, out_hi, out_lo,
                                  out_cb_total, out_cb_off, stream);
  set_error("dp_deconv2x_cw: C_in=%d not instantiated (32, 64); use dp_deconv2x", cin);
  return 1;
}

extern "C" int dp_upsample2x(const void* in_hi, const void* in_lo, int in_cb_total, int in_cb_off, int ncb, int N, int D,
                             int H, int W, void* out_hi, void* out_lo, int out_cb_total, int out_cb_off,
                             cudaStream_t stream) {
  DP_REQUIRE(D <= 65535 && N * ncb <= 65535, "dp_upsample2x: grid limits (D=%d, N*ncb=%d)", D, N * ncb);
  int cols = 8;
  while (cols < 64 && cols < W) cols <<= 1;
  const int rows = 256 / cols, chunks = (W + cols - 1) / cols;
  dim3 qgrid(static_cast<unsigned>(chunks * ((H + rows - 1) / rows)), D, N * ncb);
#define DP_UPQ_LAUNCH(C_) upsample2x_quad_kernel<C_><<<qgrid, 256, 0, stream>>>(static_cast<const __half*>(in_hi), \
      static_cast<const __half*>(in_lo), in_cb_total, in_cb_off, ncb, D, H, W, static_cast<__half*>(out_hi),        \
      static_cast<__half*>(out_lo), out_cb_total, out_cb_off)
  switch (cols) {
    case 8: DP_UPQ_LAUNCH(8); break;
    case 16: DP_UPQ_LAUNCH(16); break;
    case 32: DP_UPQ_LAUNCH(32); break;
    default: DP_UPQ_LAUNCH(64); break;
  }
#undef DP_UPQ_LAUNCH
  DP_CHECK(cudaGetLastError());
  return 0;
}

extern "C" int dp_layernorm(const float* x, const float* gamma, const float* beta, int rows, int cols, void* out_f16,
                            float* out_f32, cudaStream_t stream) {
  layernorm_kernel<<<blocks_for(rows, 8), 256, 0, stream>>>(x, gamma, beta, rows, cols, static_cast<__half*>(out_f16), out_f32);
  DP_CHECK(cudaGetLastError());
  return 0;
}

extern "C" int dp_splitk_reduce(const float* ws, int splits, int M, int N, const float* bias, const float* rowvec,
                                int row_period, float* out, cudaStream_t stream) {
  DP_REQUIRE(N % 4 == 0, "dp_splitk_reduce: N=%d must be a multiple of 4", N);
  const long long MN = static_cast<long long>(M) * N;
  splitk_reduce_kernel<<<blocks_for(MN / 4, 256), 256, 0, stream>>>(ws, splits, MN, N, bias, rowvec, row_period > 0 ? row_period : 1, out);
  DP_CHECK(cudaGetLastError());
  return 0;
}

extern "C" int dp_softmax(const float* s, int rows, int cols, int ld_in, void* p, int ld_out, cudaStream_t stream) {
  softmax_kernel<<<blocks_for(rows, 8), 256, 0, stream>>>(s, rows, cols, ld_in, static_cast<__half*>(p), ld_out);
  DP_CHECK(cudaGetLastError());
  return 0;
}

extern "C" int dp_patchify(const void* in_c8, int cb_total, int cb_off, int ncb, int N, int S0, int S1, int S2, void* out,
                           cudaStream_t stream) {
  DP_REQUIRE(S0 % 16 == 0 && S1 % 16 == 0 && S2 % 16 == 0, "dp_patchify: volume %dx%dx%d not divisible by the 16^3 patch", S0, S1, S2);
  const long long total = static_cast<long long>(S0 / 16) * (S1 / 16) * (S2 / 16) * ncb * 4096;
  DP_REQUIRE(total < (1LL << 31), "dp_patchify: %lld vectors per image exceed the 32-bit index range", total);
  dim3 grid(blocks_for(total, 256), N);
  patchify_kernel<<<grid, 256, 0, stream>>>(static_cast<const __half*>(in_c8), cb_total, cb_off, ncb, S0, S1, S2, static_cast<__half*>(out));
  DP_CHECK(cudaGetLastError());
  return 0;
}

extern "C" int dp_patchify_planar(const float* in_planar, int N, int S0, int S1, int S2, void* out, cudaStream_t stream) {
  DP_REQUIRE(S0 % 16 == 0 && S1 % 16 == 0 && S2 % 16 == 0, "dp_patchify_planar: volume %dx%dx%d is not a multiple of the patch size 16", S0, S1, S2);
  const long long total = static_cast<long long>(S0 / 16) * (S1 / 16) * (S2 / 16) * 512;
  DP_REQUIRE(total < (1LL << 31), "dp_patchify_planar: %lld vectors per image exceed the 32-bit index range", total);
  dim3 grid(blocks_for(total, 256), N);
  patchify_planar_kernel<<<grid, 256, 0, stream>>>(in_planar, S0, S1, S2, static_cast<__half*>(out));
  DP_CHECK(cudaGetLastError());
  return 0;
}

extern "C" int dp_handoff(const float* logits, int ncls, const float* ptv, const float* ct, int N, int S, void* out_hi,
                          void* out_lo, int out_cb_total, int out_cb_off, float* structures, cudaStream_t stream) {
  DP_REQUIRE(ncls == 8, "dp_handoff: expects background + 7 OAR classes, got %d", ncls);
  dim3 grid((S + 31) / 32, (S + 31) / 32, N * S);
  handoff_kernel<<<grid, 256, 0, stream>>>(logits, ncls, ptv, ct, S, static_cast<__half*>(out_hi), static_cast<__half*>(out_lo),
                                           out_cb_total, out_cb_off, structures);
  DP_CHECK(cudaGetLastError());
  return 0;
}

static int fill_windows(WindowList& wl, int n, const int* b, const int* x0, const int* y0, const int* z0) {
  DP_REQUIRE(n >= 1 && n <= 64, "sliding window: 1..64 windows per launch, got %d", n);
  wl.n = n;
  for (int i = 0; i < n; ++i) { wl.b[i] = b[i]; wl.x0[i] = x0[i]; wl.y0[i] = y0[i]; wl.z0[i] = z0[i]; }
  return 0;
}

extern "C" int dp_crop_pack(const float* src, int C, int S0, int S1, int S2, int R, int n_win, const int* win_b,
                            const int* win_x0, const int* win_y0, const int* win_z0, void* hi, void* lo, int cb_total,
                            int cb_off, cudaStream_t stream) {
  WindowList wl;
  if (int rc = fill_windows(wl, n_win, win_b, win_x0, win_y0, win_z0)) return rc;
  const int ncb = (C + 7) / 8;
  dim3 grid(blocks_for(static_cast<long long>(R) * R * R, 256), n_win * ncb);
  crop_pack_kernel<<<grid, 256, 0, stream>>>(src, C, S0, S1, S2, R, wl, static_cast<__half*>(hi), static_cast<__half*>(lo), cb_total, cb_off, ncb);
  DP_CHECK(cudaGetLastError());
  return 0;
}

extern "C" int dp_window_add(const float* win, int ncls, int R, int n_win, const int* win_b, const int* win_x0,
                             const int* win_y0, const int* win_z0, const int* win_slot, float* out, int S0, int S1, int S2,
                             cudaStream_t stream) {
  // one kernel per window, launched in list order: overlapping windows accumulate deterministically in the
  // reference's order.  win_slot[i] = batch entry of the predictor output that holds window i.
  WindowList wl;
  if (int rc = fill_windows(wl, n_win, win_b, win_x0, win_y0, win_z0)) return rc;
  for (int i = 0; i < n_win; ++i) {
    WindowList one;
    one.n = 1; one.b[0] = wl.b[i]; one.x0[0] = wl.x0[i]; one.y0[0] = wl.y0[i]; one.z0[0] = wl.z0[i];
    dim3 grid(blocks_for(static_cast<long long>(R) * R * R, 256), ncls);
    window_add_kernel<<<grid, 256, 0, stream>>>(win + static_cast<size_t>(win_slot[i]) * ncls * R * R * R, ncls, R, one, ncls, out, S0, S1, S2, 0);
  }
  DP_CHECK(cudaGetLastError());
  return 0;
}

extern "C" int dp_div_count(float* data, const float* count, long long vol, int rows, cudaStream_t stream) {
  dim3 grid(blocks_for(vol, 256), rows < 64 ? rows : 64);
  div_count_kernel<<<grid, 256, 0, stream>>>(data, count, vol, rows);
  DP_CHECK(cudaGetLastError());
  return 0;
}

extern "C" int dp_conv3d_direct(const void* in_hi, const void* in_lo, int in_cb_total, int in_cb_off, int cin, int N, int D,
                                int H, int W, int k, int stride, int dil, const float* w_packed, const float* scale,
                                const float* shift, int relu, int cout, float* out_raw, void* out_hi, void* out_lo,
                                int out_cb_total, int out_cb_off, double* stats_out, cudaStream_t stream) {
  DP_REQUIRE(cin % 8 == 0, "dp_conv3d_direct: C_in=%d must be a multiple of 8", cin);
  DirectConvParams p{};
  p.in_hi = static_cast<const __half*>(in_hi); p.in_lo = static_cast<const __half*>(in_lo);
  p.in_cb_total = in_cb_total; p.in_cb_off = in_cb_off; p.cin = cin;
  p.D = D; p.H = H; p.W = W; p.k = k; p.stride = stride; p.dil = dil; p.pad = dil * (k - 1) / 2;
  p.Do = (D + 2 * p.pad - dil * (k - 1) - 1) / stride + 1;
  p.Ho = (H + 2 * p.pad - dil * (k - 1) - 1) / stride + 1;
  p.Wo = (W + 2 * p.pad - dil * (k - 1) - 1) / stride + 1;
  p.w = w_packed; p.scale = scale; p.shift = shift; p.relu = relu; p.cout = cout;
  p.out_raw = out_raw; p.out_hi = static_cast<__half*>(out_hi); p.out_lo = static_cast<__half*>(out_lo);
  p.out_cb_total = out_cb_total; p.out_cb_off = out_cb_off; p.stats_out = stats_out;
  const size_t smem = static_cast<size_t>(k) * k * k * DC2_CI * DC2_CO * sizeof(float);
  DP_REQUIRE(smem <= 200 * 1024, "dp_conv3d_direct: kernel %d too large for the weight tile", k);
  static size_t configured = 48 * 1024;
  if (smem > configured) {
    DP_CHECK(cudaFuncSetAttribute(direct_conv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    configured = smem;
  }
  const long long vox_o = static_cast<long long>(p.Do) * p.Ho * p.Wo;
  dim3 grid(blocks_for(vox_o, 128), (cout + DC2_CO - 1) / DC2_CO, N);
  direct_conv_kernel<<<grid, 128, smem, stream>>>(p);
  DP_CHECK(cudaGetLastError());
  return 0;
}
