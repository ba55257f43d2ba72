// 1x1x1 convolution over the channel concatenation of several sources on tcgen05 tensor cores, with the sources'
// pending InstanceNorm + activation applied ON LOAD by the threads that stage the A operand.
//
// Why: the SIMT kernel (ops.cu pointwise_cw_kernel) is issue bound — 1250-1770 thread instructions per voxel at 32
// input channels, 37 % of the HBM copy bandwidth (profiles/r1_norm_pointwise_ncu.md).  The contraction itself is tiny
// (K <= 128, N <= 64), so here it costs a handful of MMAs per 128-voxel tile and the kernel is left with what is
// irreducible: read every source once, normalise/activate in registers, write the result once.
//
//   thread t of the 128-thread CTA owns voxel row t of the tile, twice:
//     * producer: loads its voxel's channels (8 per 16/32-byte vector, coalesced across the warp), applies
//       (x - mean) * rstd and the activation, splits the fp32 result into an fp16 hi + lo pair and stores both as
//       16-byte rows of the canonical K-major no-swizzle UMMA layout  A[k block][row][8]  (LBO = 128 rows * 16 B,
//       SBO = 128 B);
//     * epilogue: TMEM lane t holds its voxel's outputs; bias, per-(n,c) statistics, fp32 raw / fp16 store.
//   thread 0 issues, per 16-channel K step, D += A_hi * [W_hi | W_lo]  (N = 2*C_out) and D += A_lo * W_hi (N = C_out):
//   the 3-term operand split (~22 mantissa bits), so the result matches the fp32 SIMT kernel it replaces.
// Several CTAs are resident per SM (20 KB of shared memory and 32 TMEM columns at K = 32, C_out = 16); their load /
// MMA / store phases overlap each other, so no intra-CTA pipelining is needed.
#include <atomic>

#include "common.cuh"
#include "dose_b200.h"

namespace dp {

constexpr int kPtcMaxKb = 16;            // K <= 128 input channels (padded to 8-channel blocks)

struct PtcBlock {
  const __half* hi; const __half* lo; const float* raw;   // offset to this channel block of image 0 (all null: zero block)
  long long n_stride;                                      // elements between images
  const double* stats; int stat_c0, stat_C;                // instance statistics of the source (or null), channel of j = 0
  int act;
  const double* stats0; int act0;                          // optional FIRST stage: act0(IN(x; stats0)), then IN(.; stats) + act —
                                                           // conv_block_3's closing IN + ReLU straight from its raw conv output
};
struct PtcParams {
  PtcBlock blk[kPtcMaxKb];
  int nkb;                              // 8-channel blocks (even)
  const __half* wpack;                  // [nkb][2*cout][8] fp16: rows 0..cout-1 = W_hi, cout..2cout-1 = W_lo
  const float* bias;                    // [cout] or null
  long long vox; double inv_vox;
  int tiles;                            // 128-voxel tiles per image
  int tiles_per_cta;                    // CTA x of image y owns tiles [x * tiles_per_cta, ...): a function of vox alone, so the
                                        // summation order of the statistics does not depend on the batch size
  float* out_raw; __half* out_hi; __half* out_lo; int out_cb_total, out_cb_off;
  double* stats_out;
  int* err_flag;
};

// mean / rstd of channel idx from fp64 {sum, sumsq}; biased variance, eps 1e-5 (nn.InstanceNorm3d) — as ops.cu finalize_stats
__device__ __forceinline__ void ptc_finalize_stats(const double* stats, size_t idx, double inv_count, float& mean, float& rstd) {
  const double s = stats[idx * 2], ss = stats[idx * 2 + 1];
  const double m = s * inv_count;
  double var = ss * inv_count - m * m;
  if (var < 0.0) var = 0.0;
  mean = static_cast<float>(m);
  rstd = rsqrtf(static_cast<float>(var) + 1e-5f);
}

__device__ __forceinline__ void ptc_act8(float (&y)[8], int act) {
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
      for (int j = 0; j < 8; ++j) y[j] = mish_fast(y[j]);
      break;
    case ACT_GELU:
#pragma unroll
      for (int j = 0; j < 8; ++j) y[j] = act_apply(y[j], ACT_GELU);
      break;
    default: break;
  }
}

// 16 values per lane -> column sums over the 32 lanes; lane l ends up holding column (l >> 1) & 15 (both lanes of a pair).
__device__ __forceinline__ float ptc_colsum16(const float (&v)[16], int lane) {
  float a[8], b[4], c[2];
  const bool h4 = lane & 16, h3 = lane & 8, h2 = lane & 4, h1 = lane & 2;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const float keep = h4 ? v[j + 8] : v[j], send = h4 ? v[j] : v[j + 8];
    a[j] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float keep = h3 ? a[j + 4] : a[j], send = h3 ? a[j] : a[j + 4];
    b[j] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
  }
#pragma unroll
  for (int j = 0; j < 2; ++j) {
    const float keep = h2 ? b[j + 2] : b[j], send = h2 ? b[j] : b[j + 2];
    c[j] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
  }
  const float keep = h1 ? c[1] : c[0], send = h1 ? c[0] : c[1];
  float r = keep + __shfl_xor_sync(0xffffffffu, send, 2);
  r += __shfl_xor_sync(0xffffffffu, r, 1);
  return r;
}

// One 8-channel block of one voxel as it sits in HBM: fp32 raw = 32 bytes in (a, b); fp16 hi (+ lo) = 16 (+16) bytes.
struct PtcVec { uint4 a, b; };

__device__ __forceinline__ void ptc_load(const PtcBlock& B, int n, long long v, bool valid, PtcVec& r) {
  r.a = make_uint4(0u, 0u, 0u, 0u);
  r.b = make_uint4(0u, 0u, 0u, 0u);
  if (!valid) return;
  const size_t off = static_cast<size_t>(n) * B.n_stride + static_cast<size_t>(v) * 8;
  if (B.raw != nullptr) {
    float y[8];
    ld_global_v8f(B.raw + off, y);
    r.a = make_uint4(__float_as_uint(y[0]), __float_as_uint(y[1]), __float_as_uint(y[2]), __float_as_uint(y[3]));
    r.b = make_uint4(__float_as_uint(y[4]), __float_as_uint(y[5]), __float_as_uint(y[6]), __float_as_uint(y[7]));
  } else if (B.hi != nullptr) {
    r.a = *reinterpret_cast<const uint4*>(B.hi + off);
    if (B.lo != nullptr) r.b = *reinterpret_cast<const uint4*>(B.lo + off);
  }
}

// registers -> fp32 values -> (x - mean) * rstd -> activation -> fp16 hi / lo rows of the A operand
__device__ __forceinline__ void ptc_stage(const PtcBlock& B, const PtcVec& r, const float* mean /* -mean*rstd */, const float* rstd,
                                          const float* mean0, const float* rstd0, bool valid, uint8_t* dst_hi, uint8_t* dst_lo) {
  float x[8];
  if (B.raw != nullptr) {
    x[0] = __uint_as_float(r.a.x); x[1] = __uint_as_float(r.a.y); x[2] = __uint_as_float(r.a.z); x[3] = __uint_as_float(r.a.w);
    x[4] = __uint_as_float(r.b.x); x[5] = __uint_as_float(r.b.y); x[6] = __uint_as_float(r.b.z); x[7] = __uint_as_float(r.b.w);
  } else {
    const __half2* h = reinterpret_cast<const __half2*>(&r.a);
    const __half2* l = reinterpret_cast<const __half2*>(&r.b);       // zeros when the source has no lo part
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 f = __half22float2(h[j]), g = __half22float2(l[j]);
      x[2 * j] = f.x + g.x;
      x[2 * j + 1] = f.y + g.y;
    }
  }
  if (valid) {
    if (B.stats0 != nullptr) {
#pragma unroll
      for (int j = 0; j < 8; ++j) x[j] = fmaf(x[j], rstd0[j], mean0[j]);
      ptc_act8(x, B.act0);
    }
    if (B.stats != nullptr) {
#pragma unroll
      for (int j = 0; j < 8; ++j) x[j] = fmaf(x[j], rstd[j], mean[j]);
    }
    ptc_act8(x, B.act);
  }
  __align__(16) __half2 h[4];
  __align__(16) __half2 l[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    h[j] = __floats2half2_rn(x[2 * j], x[2 * j + 1]);
    const float2 f = __half22float2(h[j]);
    l[j] = __floats2half2_rn(x[2 * j] - f.x, x[2 * j + 1] - f.y);
  }
  *reinterpret_cast<uint4*>(dst_hi) = *reinterpret_cast<const uint4*>(h);
  *reinterpret_cast<uint4*>(dst_lo) = *reinterpret_cast<const uint4*>(l);
}

// NKB > 0: the tile's NKB channel blocks are PREFETCHED into registers one tile ahead (the loads of tile i+1 are in flight
// while tile i is normalised, contracted and stored) and both the A stage and the TMEM accumulator are double buffered,
// so the epilogue of tile i-1 overlaps the MMAs of tile i.  NKB == 0: generic path for K up to 128 (coarse levels, small).
template <int CO, int NKB>
__global__ void __launch_bounds__(128, (NKB > 0 && NKB <= 4) ? 4 : 2) pointwise_tc_kernel(const __grid_constant__ PtcParams p) {
  constexpr int ACC_COLS = 2 * CO;                                // [W_hi | W_lo] halves, added in the epilogue
  constexpr int NBUF = NKB > 0 ? 2 : 1;
  constexpr int TM_COLS = (NBUF * ACC_COLS <= 32) ? 32 : ((NBUF * ACC_COLS <= 64) ? 64 : ((NBUF * ACC_COLS <= 128) ? 128 : 256));
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((128u - (smem_u32(smem_raw) & 127u)) & 127u);
  __shared__ uint64_t mma_bar[2];
  __shared__ uint32_t tmem_base_smem;
  __shared__ float s_mean[kPtcMaxKb * 8], s_rstd[kPtcMaxKb * 8], s_bias[CO];
  __shared__ float s_mean0[kPtcMaxKb * 8], s_rstd0[kPtcMaxKb * 8];       // first-stage constants (PtcBlock::stats0)
  __shared__ float s_stat[4][CO][2];

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int n = blockIdx.y;
  const int nkb = NKB > 0 ? NKB : p.nkb;
  const uint32_t stage_bytes = static_cast<uint32_t>(nkb) * 4096u;      // A_hi + A_lo of one tile
  uint8_t* a_stage = smem;                                              // [NBUF][hi|lo][nkb][128][8] halves
  uint8_t* b_w = smem + static_cast<size_t>(NBUF) * stage_bytes;        // [nkb][2*CO][8] halves

  if (tid == 0) { mbar_init(&mma_bar[0], 1); mbar_init(&mma_bar[1], 1); fence_barrier_init(); }
  if (warp == 0) tmem_alloc<TM_COLS>(&tmem_base_smem);
  for (int i = tid; i < nkb * 8; i += 128) {
    const PtcBlock& B = p.blk[i >> 3];
    const int c = B.stat_c0 + (i & 7);
    float m = 0.f, r = 1.f;
    if (B.stats != nullptr) {
      if (c < B.stat_C) ptc_finalize_stats(B.stats, static_cast<size_t>(n) * B.stat_C + c, p.inv_vox, m, r);
      else r = 0.f;                                              // padded channel: (x - 0) * 0
    }
    s_mean[i] = -m * r;                                           // x * rstd + (-mean * rstd): one FFMA per element
    s_rstd[i] = r;
    float m0 = 0.f, r0 = 1.f;
    if (B.stats0 != nullptr) {
      if (c < B.stat_C) ptc_finalize_stats(B.stats0, static_cast<size_t>(n) * B.stat_C + c, p.inv_vox, m0, r0);
      else r0 = 0.f;
    }
    s_mean0[i] = -m0 * r0;
    s_rstd0[i] = r0;
  }
  for (int i = tid; i < CO; i += 128) s_bias[i] = p.bias ? p.bias[i] : 0.f;
  for (int i = tid; i < 4 * CO * 2; i += 128) (&s_stat[0][0][0])[i] = 0.f;
  {  // weights: nkb * 2*CO rows of 16 bytes, copied as they are packed
    const uint4* src = reinterpret_cast<const uint4*>(p.wpack);
    uint4* dst = reinterpret_cast<uint4*>(b_w);
    for (int i = tid; i < nkb * 2 * CO; i += 128) dst[i] = src[i];
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_smem;
  const uint32_t tlane = tmem_base + (static_cast<uint32_t>(warp * 32) << 16);

  // descriptors: A K-major no swizzle, LBO = 2048 B (next 8-channel block), SBO = 128 B (next 8 rows);
  // B rows = output columns: LBO = 2*CO*16 B, SBO = 128 B
  const uint32_t idesc_full = make_idesc_f16(128, 2 * CO);
  const uint32_t idesc_half = make_idesc_f16(128, CO);
  const uint64_t a_desc0 = make_smem_desc(smem_u32(a_stage), 2048u, 128u, 0);
  const uint64_t b_desc = make_smem_desc(smem_u32(b_w), 2u * CO * 16u, 128u, 0);

  constexpr bool kRegStats = CO <= 16;                           // per-thread running sums; folded across lanes once
  float ts1[kRegStats ? CO : 1], ts2[kRegStats ? CO : 1];
#pragma unroll
  for (int j = 0; j < (kRegStats ? CO : 1); ++j) { ts1[j] = 0.f; ts2[j] = 0.f; }
  const bool want_stats = p.stats_out != nullptr;

  auto issue_mma = [&](int buf) {      // thread 0 only
    tc_fence_after();
    const uint64_t a_hi = a_desc0 + static_cast<uint64_t>((static_cast<uint32_t>(buf) * stage_bytes) >> 4);
    const uint64_t a_lo = a_hi + static_cast<uint64_t>((static_cast<uint32_t>(nkb) * 2048u) >> 4);
    const uint32_t d = tmem_base + static_cast<uint32_t>(buf * ACC_COLS);
    for (int ks = 0; ks < (nkb >> 1); ++ks) {
      const uint64_t adv_a = static_cast<uint64_t>((ks * 4096u) >> 4);            // two 8-channel blocks per K step
      const uint64_t adv_b = static_cast<uint64_t>((ks * 2u * (2u * CO * 16u)) >> 4);
      umma_f16_ss(d, a_hi + adv_a, b_desc + adv_b, idesc_full, ks > 0 ? 1u : 0u);
      umma_f16_ss(d, a_lo + adv_a, b_desc + adv_b, idesc_half, 1u);
    }
    umma_commit(&mma_bar[buf]);
  };
  auto epilogue = [&](int buf, long long v, bool valid) {
#pragma unroll
    for (int c0 = 0; c0 < CO; c0 += 16) {
      uint32_t r[16], r2[16];
      tmem_ld16(tlane + buf * ACC_COLS + c0, r);
      tmem_ld16(tlane + buf * ACC_COLS + CO + c0, r2);
      tmem_ld_wait();
      float y[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) y[j] = __uint_as_float(r[j]) + __uint_as_float(r2[j]) + s_bias[c0 + j];
      if (want_stats) {
        if constexpr (kRegStats) {
          if (valid) {
#pragma unroll
            for (int j = 0; j < 16; ++j) { ts1[c0 + j] += y[j]; ts2[c0 + j] = fmaf(y[j], y[j], ts2[c0 + j]); }
          }
        } else {
          float a[16], b[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) { a[j] = valid ? y[j] : 0.f; b[j] = a[j] * a[j]; }
          const float s1 = ptc_colsum16(a, lane), s2 = ptc_colsum16(b, lane);
          if ((lane & 1) == 0) {
            s_stat[warp][c0 + ((lane >> 1) & 15)][0] += s1;
            s_stat[warp][c0 + ((lane >> 1) & 15)][1] += s2;
          }
        }
      }
      if (valid) {
#pragma unroll
        for (int b = 0; b < 2; ++b) {
          const size_t off = ((static_cast<size_t>(n) * p.out_cb_total + p.out_cb_off + (c0 >> 3) + b) * p.vox + v) * 8;
          if (p.out_raw != nullptr) {
            float y8[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) y8[j] = y[b * 8 + j];
            st_global_v8f(p.out_raw + off, y8);
          }
          if (p.out_hi != nullptr) {
            __align__(16) __half2 h[4];
            __align__(16) __half2 l[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              h[j] = __floats2half2_rn(y[b * 8 + 2 * j], y[b * 8 + 2 * j + 1]);
              const float2 f = __half22float2(h[j]);
              l[j] = __floats2half2_rn(y[b * 8 + 2 * j] - f.x, y[b * 8 + 2 * j + 1] - f.y);
            }
            *reinterpret_cast<uint4*>(p.out_hi + off) = *reinterpret_cast<const uint4*>(h);
            if (p.out_lo != nullptr) *reinterpret_cast<uint4*>(p.out_lo + off) = *reinterpret_cast<const uint4*>(l);
          }
        }
      }
    }
  };

  if constexpr (NKB > 0) {
    // -------------------------------------------------------------- software pipeline over this CTA's tiles
    PtcVec cur[NKB];
    int tile = blockIdx.x * p.tiles_per_cta;
    const int tile_end = min(p.tiles, tile + p.tiles_per_cta);
    {
      const long long v0 = static_cast<long long>(tile) * 128 + tid;
#pragma unroll
      for (int kb = 0; kb < NKB; ++kb) ptc_load(p.blk[kb], n, v0, tile < tile_end && v0 < p.vox, cur[kb]);
    }
    uint32_t phase_bits = 0u;                            // bit b = parity to wait for on mma_bar[b]
    long long v_prev = 0;
    bool valid_prev = false, have_prev = false;
    int it = 0;
#pragma unroll 1
    for (; tile < tile_end; ++tile, ++it) {
      const int buf = it & 1;
      const long long v = static_cast<long long>(tile) * 128 + tid;
      const bool valid = v < p.vox;
      // stage tile `tile` from the prefetched registers, then refill them with the next tile's loads
      uint8_t* sh = a_stage + static_cast<size_t>(buf) * stage_bytes + tid * 16;
#pragma unroll
      for (int kb = 0; kb < NKB; ++kb)
        ptc_stage(p.blk[kb], cur[kb], &s_mean[kb * 8], &s_rstd[kb * 8], &s_mean0[kb * 8], &s_rstd0[kb * 8], valid, sh + kb * 2048, sh + (NKB + kb) * 2048);
      {
        const int nt = tile + 1;
        const long long vn = static_cast<long long>(nt) * 128 + tid;
#pragma unroll
        for (int kb = 0; kb < NKB; ++kb) ptc_load(p.blk[kb], n, vn, nt < tile_end && vn < p.vox, cur[kb]);
      }
      fence_proxy_async();
      tc_fence_before();
      __syncthreads();
      if (tid == 0) {
        issue_mma(buf);
        if (have_prev) mbar_wait(&mma_bar[buf ^ 1], (phase_bits >> (buf ^ 1)) & 1u, p.err_flag);
      }
      if (have_prev) {                                   // epilogue of the previous tile overlaps this tile's MMAs
        __syncthreads();                                 // (its MMAs were committed one iteration ago: thread 0 polled)
        phase_bits ^= 1u << (buf ^ 1);
        tc_fence_after();
        epilogue(buf ^ 1, v_prev, valid_prev);
      }
      v_prev = v; valid_prev = valid; have_prev = true;
    }
    if (have_prev) {
      const int buf = (it - 1) & 1;
      if (tid == 0) mbar_wait(&mma_bar[buf], (phase_bits >> buf) & 1u, p.err_flag);
      __syncthreads();
      tc_fence_after();
      epilogue(buf, v_prev, valid_prev);
    }
  } else {
    uint32_t phase = 0;
    const int tile_end = min(p.tiles, (static_cast<int>(blockIdx.x) + 1) * p.tiles_per_cta);
#pragma unroll 1
    for (int tile = blockIdx.x * p.tiles_per_cta; tile < tile_end; ++tile) {
      const long long v = static_cast<long long>(tile) * 128 + tid;
      const bool valid = v < p.vox;
      uint8_t* sh = a_stage + tid * 16;
#pragma unroll 2
      for (int kb = 0; kb < nkb; ++kb) {
        PtcVec r;
        ptc_load(p.blk[kb], n, v, valid, r);
        ptc_stage(p.blk[kb], r, &s_mean[kb * 8], &s_rstd[kb * 8], &s_mean0[kb * 8], &s_rstd0[kb * 8], valid, sh + kb * 2048, sh + (nkb + kb) * 2048);
      }
      fence_proxy_async();               // generic-proxy smem writes -> visible to the tensor core (async proxy)
      tc_fence_before();                 // this thread's TMEM loads of the previous tile are complete (tcgen05.wait::ld)
      __syncthreads();
      if (tid == 0) { issue_mma(0); mbar_wait(&mma_bar[0], phase, p.err_flag); }
      __syncthreads();
      phase ^= 1;
      tc_fence_after();
      epilogue(0, v, valid);
    }
  }

  // ------------------------------------------------------------------ statistics: lanes -> warp -> fp64 atomics
  if (want_stats) {
    if constexpr (kRegStats) {
#pragma unroll
      for (int c0 = 0; c0 < CO; c0 += 16) {
        float a[16], b[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) { a[j] = ts1[c0 + j]; b[j] = ts2[c0 + j]; }
        const float s1 = ptc_colsum16(a, lane), s2 = ptc_colsum16(b, lane);
        if ((lane & 1) == 0) {
          s_stat[warp][c0 + ((lane >> 1) & 15)][0] = s1;
          s_stat[warp][c0 + ((lane >> 1) & 15)][1] = s2;
        }
      }
    }
    __syncthreads();
    for (int i = tid; i < CO * 2; i += 128) {
      const int c = i >> 1, which = i & 1;
      const float t = s_stat[0][c][which] + s_stat[1][c][which] + s_stat[2][c][which] + s_stat[3][c][which];
      atomicAdd(&p.stats_out[(static_cast<size_t>(n) * CO + c) * 2 + which], static_cast<double>(t));
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc<TM_COLS>(tmem_base);
  }
}

template <int CO, int NKB>
static int launch_ptc2(const PtcParams& p, int N, cudaStream_t stream) {
  constexpr int NBUF = NKB > 0 ? 2 : 1;
  const size_t smem = static_cast<size_t>(NBUF) * p.nkb * 4096 + static_cast<size_t>(p.nkb) * 2 * CO * 16 + 128;
  DP_REQUIRE(smem <= 200 * 1024, "dp_pointwise_tc: K=%d, C_out=%d needs %zu bytes of shared memory", p.nkb * 8, CO, smem);
  auto kern = pointwise_tc_kernel<CO, NKB>;
  static std::atomic<unsigned long long> configured{0};          // bit per device: max dynamic smem set + occupancy cached
  static int resident[64];
  int dev = 0;
  DP_CHECK(cudaGetDevice(&dev));
  if (dev < 0 || dev >= 64) dev = 0;
  if (!(configured.load() >> dev & 1ull)) {
    DP_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    configured.fetch_or(1ull << dev);
  }
  int per_sm = 0;
  (void)resident;
  (void)per_sm;
  PtcParams q = p;
  q.tiles_per_cta = p.tiles / 256;                       // 128^3: 32 tiles (512 CTAs per image), 64^3: 8 (256 CTAs), <= 32^3: 1
  if (q.tiles_per_cta > 32) q.tiles_per_cta = 32;
  if (q.tiles_per_cta < 1) q.tiles_per_cta = 1;
  const int ctas = (p.tiles + q.tiles_per_cta - 1) / q.tiles_per_cta;
  dim3 grid(static_cast<unsigned>(ctas), static_cast<unsigned>(N));
  kern<<<grid, 128, smem, stream>>>(q);
  return check_cuda(cudaGetLastError(), "pointwise_tc");
}

template <int CO>
static int launch_ptc(const PtcParams& p, int N, cudaStream_t stream) {
  if (p.nkb == 2) return launch_ptc2<CO, 2>(p, N, stream);
  if (p.nkb == 4) return launch_ptc2<CO, 4>(p, N, stream);
  if (p.nkb == 8) return launch_ptc2<CO, 8>(p, N, stream);
  return launch_ptc2<CO, 0>(p, N, stream);
}

}  // namespace dp

extern "C" int dp_pointwise_tc(int nsrc, const void* const* src_hi, const void* const* src_lo, const float* const* src_raw,
                               const int* src_cb_total, const int* src_cb_off, const int* src_C,
                               const double* const* src_stats, const int* src_act, const double* const* src_stats0,
                               const int* src_act0, const void* wpack, const float* bias,
                               int cout, int N, long long vox, float* out_raw, void* out_hi, void* out_lo, int out_cb_total,
                               int out_cb_off, double* stats_out, int* err_flag, cudaStream_t stream) {
  using namespace dp;
  DP_REQUIRE(nsrc >= 1 && nsrc <= 8, "dp_pointwise_tc: 1..8 sources, got %d", nsrc);
  DP_REQUIRE(cout == 16 || cout == 32 || cout == 64, "dp_pointwise_tc: C_out=%d unsupported (16, 32, 64)", cout);
  DP_REQUIRE(out_raw != nullptr || out_hi != nullptr, "dp_pointwise_tc: no output tensor given");
  PtcParams p{};
  int kb = 0;
  for (int s = 0; s < nsrc; ++s) {
    const int ncb = (src_C[s] + 7) / 8;
    DP_REQUIRE(kb + ncb <= kPtcMaxKb, "dp_pointwise_tc: more than %d input channels", kPtcMaxKb * 8);
    const __half* hi = static_cast<const __half*>(src_hi ? src_hi[s] : nullptr);
    const __half* lo = static_cast<const __half*>(src_lo ? src_lo[s] : nullptr);
    const float* raw = src_raw ? src_raw[s] : nullptr;
    DP_REQUIRE(hi != nullptr || raw != nullptr, "dp_pointwise_tc: source %d has no tensor", s);
    for (int b = 0; b < ncb; ++b, ++kb) {
      PtcBlock& B = p.blk[kb];
      const size_t boff = static_cast<size_t>(src_cb_off[s] + b) * static_cast<size_t>(vox) * 8;
      B.hi = (raw == nullptr && hi) ? hi + boff : nullptr;
      B.lo = (raw == nullptr && lo) ? lo + boff : nullptr;
      B.raw = raw ? raw + boff : nullptr;
      B.n_stride = static_cast<long long>(src_cb_total[s]) * vox * 8;
      B.stats = src_stats ? src_stats[s] : nullptr;
      B.stat_c0 = b * 8; B.stat_C = src_C[s];
      B.act = src_act ? src_act[s] : 0;
      B.stats0 = src_stats0 ? src_stats0[s] : nullptr;
      B.act0 = (src_act0 && B.stats0) ? src_act0[s] : 0;
    }
  }
  if (kb & 1) ++kb;                    // K steps are 16 channels: an all-null block stages zeros
  p.nkb = kb;
  p.wpack = static_cast<const __half*>(wpack); p.bias = bias; p.vox = vox; p.inv_vox = 1.0 / static_cast<double>(vox);
  p.tiles = static_cast<int>((vox + 127) / 128);
  p.out_raw = out_raw; p.out_hi = static_cast<__half*>(out_hi); p.out_lo = static_cast<__half*>(out_lo);
  p.out_cb_total = out_cb_total; p.out_cb_off = out_cb_off; p.stats_out = stats_out; p.err_flag = err_flag;
  if (cout == 16) return launch_ptc<16>(p, N, stream);
  if (cout == 32) return launch_ptc<32>(p, N, stream);
  return launch_ptc<64>(p, N, stream);
}
