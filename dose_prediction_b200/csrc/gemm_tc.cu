// C[M,N] = A[M,K] * B[N,K]^T on tcgen05 tensor cores (fp16 operands, fp32 accumulation in TMEM).
//
// Used for every nn.Linear on the hot path (monai PatchEmbeddingBlock "perceptron" Linear, SABlock
// qkv/out_proj, MLPBlock linear1/linear2), for the two attention contractions (einsum "blxd,blyd->blxy"
// and "bhxy,bhyd->bhxd" in monai SABlock.forward) and for the 2x transposed convolutions that read ViT
// tokens (deconv = GEMM [voxels, C_in] x [C_in, 8*C_out] with a pixel-shuffle scatter epilogue).
//
// Persistent: grid = min(#tiles, #SMs); every CTA walks tiles blockIdx.x + i*gridDim.x.  Tiles are
// 128 x 128 x 64, or 128 x 256 x 64 for wide-N problems (TMA 2-D boxes, 128-byte swizzle, 6- / 4-stage mbarrier ring
// that runs ahead across tiles);
// two 128-column TMEM accumulators so the epilogue of tile i overlaps the main loop of tile i+1.
// Warp roles (320 threads): warp 0 TMA producer, warp 1 MMA issuer + TMEM owner, warps 2..9 epilogue
// (two warps per TMEM lane quarter, each draining 64 of the 128 accumulator columns).
// Epilogue options: +bias[n], +rowvec[m % period][n] (position embeddings), GELU(erf), *alpha,
// +residual (fp32, may alias out_f32), fp32 / fp16 stores, deterministic split-K partials, the
// q/k/v^T head scatter, or the deconv scatter into a c8 tensor.
#include "common.cuh"
#include "dose_b200.h"

namespace dp {

struct GemmParams {
  int M, N, K;
  int batch, split_k, kb_per_split, total_kb;
  int tiles_m, tiles_n, num_tiles;
  long long split_stride;              // split-K: partial sums of split s go to out_f32 + s*split_stride
  int a_batch_rows, b_batch_rows;      // row offset per batch entry inside the A / B tensor maps
  long long c_batch_stride;            // element offset per batch entry in the outputs ...
  int c_batch_period; long long c_batch_stride2;   // ... or (z / period) * stride + (z % period) * stride2
  int ldc;
  const float* bias;
  const float* rowvec; int row_period;
  const float* resid;
  float alpha;
  int act;
  float* out_f32;
  __half* out_f16;
  // qkv scatter (mode 1): N = 3*heads*hd, tokens per image T
  int mode, heads, hd, T, vt_ld;
  __half* q; __half* kk; __half* vt; float q_scale;
  // deconv scatter (mode 2): rows = (b, d, h, w) of a [Dg,Hg,Wg] grid, cols = parity*cout + co
  int Dg, Hg, Wg, cout;
  __half* dc_hi; __half* dc_lo; int dc_cb_total, dc_cb_off;
  int* err_flag;
  // patch-embedding gather (monai PatchEmbeddingBlock "perceptron"): A is never materialised; the A tile of
  // 128 tokens x 64 K-elements is ONE 5-D TMA box over the c8 activation: rows = (gx pair, gy, gz) tokens 16 voxels apart
  // in every direction, the 128-byte row = 8 W-voxels x 8 channels at the patch offset (p1, p2, p3 half) that K block
  // stands for.  K order (c8 block, p1, p2, p3, c%8) as dp_patchify's.
  int patch_mode, patch_cb_total, patch_cb_off, patch_tokens, patch_D, patch_H;
  int debug;                 // timing experiments only (DP_GEMM_DEBUG): 1 = epilogue hands the accumulator straight back, 2 = no MMAs
};

constexpr int kGemmThreads = 320;        // producer + MMA + 8 epilogue warps (2 per TMEM lane quarter)
constexpr int kGemmEpiWarps = 8;
constexpr int BM = 128, BK = 64;
constexpr uint32_t kStageA = BM * BK * 2;
// N tile: 128 (6 stages) or 256 (4 stages).  The ViT GEMMs are bound by the L2 -> SM operand stream (a 128 x 128
// tile of a K = 768 problem pulls 393 KB for 1.6 us of MMA work); a 128 x 256 tile moves 25 % fewer bytes per FLOP.
template <int BN> struct GemmCfg {
  static constexpr int kStages = BN == 256 ? 4 : 6;
  static constexpr uint32_t kStageB = BN * BK * 2, kStage = kStageA + kStageB;
};

struct GemmTile {
  int m0, n0, bz, sk, kb0, kb1;
};
template <int BN>
__device__ __forceinline__ GemmTile gemm_tile(const GemmParams& p, int tile) {
  GemmTile t;
  const int tn = tile % p.tiles_n; tile /= p.tiles_n;
  const int tm = tile % p.tiles_m; tile /= p.tiles_m;
  t.sk = tile % p.split_k;
  t.bz = tile / p.split_k;
  t.m0 = tm * BM;
  t.n0 = tn * BN;
  t.kb0 = t.sk * p.kb_per_split;
  t.kb1 = min(p.total_kb, t.kb0 + p.kb_per_split);
  return t;
}

// Epilogue of one 128-row x BN-column accumulator (TMEM lanes = rows): the calling warp owns lane quarter `quarter` and the
// column half `chalf`; taddr = this quarter's lanes, column 0 of the accumulator.  Shared by the 1-CTA and the CTA-pair kernel.
template <int BN>
__device__ __forceinline__ void gemm_epilogue_tile(const GemmParams& p, const GemmTile& t, uint32_t taddr, int quarter, int chalf,
                                                   int lane, bool vecN) {
    const int m = t.m0 + quarter * 32 + lane;
    const int n0 = t.n0;
    const bool lead = (t.sk == 0);
    const float* rv = (p.rowvec && lead && m < p.M) ? p.rowvec + static_cast<size_t>(m % p.row_period) * p.N : nullptr;
    const float* bias = lead ? p.bias : nullptr;
    int rb = 0, rt = 0;
    size_t dc_vox = 0;
    if (p.mode == 1) { rb = m / p.T; rt = m % p.T; }
    if (p.mode == 2) {
      const int vg = p.Dg * p.Hg * p.Wg;
      rb = m / vg;
      const int v = m % vg;
      const int w = v % p.Wg, h = (v / p.Wg) % p.Hg, d = v / (p.Wg * p.Hg);
      dc_vox = (static_cast<size_t>(2 * d) * (2 * p.Hg) + 2 * h) * (2 * p.Wg) + 2 * w;
    }
    const size_t boff = p.c_batch_period > 0
        ? static_cast<size_t>(t.bz / p.c_batch_period) * p.c_batch_stride + static_cast<size_t>(t.bz % p.c_batch_period) * p.c_batch_stride2
        : static_cast<size_t>(t.bz) * p.c_batch_stride;
    const size_t row_base = boff + static_cast<size_t>(t.sk) * p.split_stride + static_cast<size_t>(m) * p.ldc;
#pragma unroll 1
    for (int cc = 0; cc < BN / 32; cc += 2) {
      // two 16-column chunks in flight per iteration (ILP across the TMEM loads)
      const int c0 = chalf * (BN / 2) + cc * 16;
      if (n0 + c0 >= p.N) break;                          // warp-uniform
      uint32_t r[2][16];
      tmem_ld16(taddr + c0, r[0]);
      const bool second = (n0 + c0 + 16 < p.N);
      if (second) tmem_ld16(taddr + c0 + 16, r[1]);
      tmem_ld_wait();
      if (m >= p.M) continue;
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        if (u == 1 && !second) break;
        const int cu = c0 + u * 16;
        const int n = n0 + cu;
        const bool full = vecN || (n + 16 <= p.N);
        float v[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(r[u][j]) * p.alpha;
        if (full) {
          if (bias) {
#pragma unroll
            for (int j = 0; j < 16; j += 4) {
              const float4 b4 = __ldg(reinterpret_cast<const float4*>(bias + n + j));
              v[j] += b4.x; v[j + 1] += b4.y; v[j + 2] += b4.z; v[j + 3] += b4.w;
            }
          }
          if (rv) {
#pragma unroll
            for (int j = 0; j < 16; j += 4) {
              const float4 b4 = __ldg(reinterpret_cast<const float4*>(rv + n + j));
              v[j] += b4.x; v[j + 1] += b4.y; v[j + 2] += b4.z; v[j + 3] += b4.w;
            }
          }
        } else {
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            if (n + j < p.N) {
              if (bias) v[j] += __ldg(&bias[n + j]);
              if (rv) v[j] += __ldg(&rv[n + j]);
            }
          }
        }
        if (p.act == ACT_GELU) {
#pragma unroll
          for (int j = 0; j < 16; ++j) v[j] = gelu_fast(v[j]);
        } else if (p.act) {
#pragma unroll
          for (int j = 0; j < 16; ++j) v[j] = act_apply(v[j], p.act);
        }
        if (p.mode == 1) {
          const int hidden = p.heads * p.hd;                // 16 | hd, so the 16 columns share (which, head)
          const int which = n / hidden, rem = n % hidden;
          const int hh = rem / p.hd, dd = rem % p.hd;
          const size_t bh = static_cast<size_t>(rb) * p.heads + hh;
          if (which == 2) {
#pragma unroll
            for (int j = 0; j < 16; ++j) p.vt[(bh * p.hd + dd + j) * p.vt_ld + rt] = __float2half_rn(v[j]);
          } else {
            __half* dst = (which == 0 ? p.q : p.kk) + (bh * p.T + rt) * p.hd + dd;
            const float sc = which == 0 ? p.q_scale : 1.f;
            __align__(16) __half h[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) h[j] = __float2half_rn(v[j] * sc);
            *reinterpret_cast<uint4*>(dst) = *reinterpret_cast<const uint4*>(&h[0]);
            *reinterpret_cast<uint4*>(dst + 8) = *reinterpret_cast<const uint4*>(&h[8]);
          }
          continue;
        }
        if (p.mode == 2) {
          const int qp = n / p.cout, co = n % p.cout;       // 16 | cout, so the 16 columns share the parity
          const size_t vo = dc_vox + (static_cast<size_t>(qp >> 2) * (2 * p.Hg) + ((qp >> 1) & 1)) * (2 * p.Wg) + (qp & 1);
          const size_t vox_out = static_cast<size_t>(8) * p.Dg * p.Hg * p.Wg;
#pragma unroll
          for (int b = 0; b < 2; ++b) {
            const size_t off = ((static_cast<size_t>(rb) * p.dc_cb_total + p.dc_cb_off + (co >> 3) + b) * vox_out + vo) * 8;
            __align__(16) __half hi[8];
            __align__(16) __half lo[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              hi[j] = __float2half_rn(v[b * 8 + j]);
              lo[j] = __float2half_rn(v[b * 8 + j] - __half2float(hi[j]));
            }
            *reinterpret_cast<uint4*>(p.dc_hi + off) = *reinterpret_cast<const uint4*>(hi);
            if (p.dc_lo) *reinterpret_cast<uint4*>(p.dc_lo + off) = *reinterpret_cast<const uint4*>(lo);
          }
          continue;
        }
        const size_t base = row_base + n;
        if (p.resid && lead) {
          if (full && (p.ldc % 4 == 0)) {
#pragma unroll
            for (int j = 0; j < 16; j += 4) {
              const float4 r4 = *reinterpret_cast<const float4*>(&p.resid[base + j]);
              v[j] += r4.x; v[j + 1] += r4.y; v[j + 2] += r4.z; v[j + 3] += r4.w;
            }
          } else {
#pragma unroll
            for (int j = 0; j < 16; ++j)
              if (n + j < p.N) v[j] += p.resid[base + j];
          }
        }
        if (p.out_f32) {
          if (full && (p.ldc % 4 == 0)) {
#pragma unroll
            for (int j = 0; j < 16; j += 4)
              *reinterpret_cast<float4*>(&p.out_f32[base + j]) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
          } else {
#pragma unroll
            for (int j = 0; j < 16; ++j)
              if (n + j < p.N) p.out_f32[base + j] = v[j];
          }
        }
        if (p.out_f16) {
          if (full && (p.ldc % 8 == 0)) {
            __align__(16) __half h[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) h[j] = __float2half_rn(v[j]);
            *reinterpret_cast<uint4*>(&p.out_f16[base]) = *reinterpret_cast<const uint4*>(&h[0]);
            *reinterpret_cast<uint4*>(&p.out_f16[base + 8]) = *reinterpret_cast<const uint4*>(&h[8]);
          } else {
#pragma unroll
            for (int j = 0; j < 16; ++j)
              if (n + j < p.N) p.out_f16[base + j] = __float2half_rn(v[j]);
          }
        }
      }
    }
}

template <int BN>
__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
               const __grid_constant__ GemmParams p) {
  constexpr int kGemmStages = GemmCfg<BN>::kStages;
  constexpr uint32_t kStage = GemmCfg<BN>::kStage;
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint64_t full_bar[kGemmStages];
  __shared__ uint64_t empty_bar[kGemmStages];
  __shared__ uint64_t acc_full[2], acc_empty[2];
  __shared__ uint32_t tmem_base_smem;
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_b);
    for (int i = 0; i < kGemmStages; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&acc_full[i], 1); mbar_init(&acc_empty[i], kGemmEpiWarps); }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<2 * BN>(&tmem_base_smem);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_smem;

  if (warp == 0) {
    // ===================================================================== TMA producer
    int stage = 0;
    uint32_t phase = 0;
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
      const GemmTile t = gemm_tile<BN>(p, tile);
      const int a_row = t.bz * p.a_batch_rows + t.m0, b_row = t.bz * p.b_batch_rows + t.n0;
      for (int kb = t.kb0; kb < t.kb1; ++kb) {
        if (!mbar_wait_relaxed(&empty_bar[stage], phase ^ 1, p.err_flag)) goto teardown;
        if (elect_one()) {
          uint8_t* sa = smem + static_cast<size_t>(stage) * kStage;
          mbar_arrive_expect_tx(&full_bar[stage], kStage);
          if (p.patch_mode) {
            const int p3h = kb & 1, p2 = (kb >> 1) & 15, p1 = (kb >> 5) & 15, cb = kb >> 9;
            const int img = t.m0 / p.patch_tokens, gx0 = (t.m0 % p.patch_tokens) >> 6;         // 64 tokens per gx slab (8 x 8)
            tma_load_5d(sa, &tmap_a, &full_bar[stage], p3h * 64, 0, 0, gx0,
                        ((img * p.patch_cb_total + p.patch_cb_off + cb) * p.patch_D + p1) * p.patch_H + p2);
          } else {
            tma_load_2d(sa, &tmap_a, &full_bar[stage], kb * BK, a_row);
          }
          tma_load_2d(sa + kStageA, &tmap_b, &full_bar[stage], kb * BK, b_row);
        }
        __syncwarp();
        if (++stage == kGemmStages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ===================================================================== MMA issuer
    const uint32_t idesc = make_idesc_f16(BM, BN);
    // 128B-swizzled K-major tiles: 8-row atoms of 1024 B (SBO), 32 B start advance per 16-element K step
    const uint32_t d_hi = (1024u >> 4) | (1u << 14) | (2u << 29);
    const uint32_t d_lo_c = 1u << 16;
    int stage = 0;
    uint32_t phase = 0;
    int iter = 0;
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++iter) {
      const GemmTile t = gemm_tile<BN>(p, tile);
      const int slot = iter & 1;
      if (!mbar_wait(&acc_empty[slot], ((iter >> 1) & 1) ^ 1, p.err_flag)) goto teardown;
      tc_fence_after();
      const uint32_t tmem_d = tmem_base + static_cast<uint32_t>(slot * BN);
      uint32_t accumulate = 0;
      for (int kb = t.kb0; kb < t.kb1; ++kb) {
        if (!mbar_wait(&full_bar[stage], phase, p.err_flag)) goto teardown;
        tc_fence_after();
        if (elect_one()) {
          const uint32_t sa16 = smem_u32(smem + static_cast<size_t>(stage) * kStage) >> 4;
          const uint32_t sb16 = sa16 + (kStageA >> 4);
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
            if (p.debug & 2) break;
            asm volatile(
                "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
                "mov.b64 da, {%1, %2};\n\t"
                "mov.b64 db, {%3, %2};\n\t"
                "setp.ne.b32 p, %5, 0;\n\t"
                "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %4, p;\n\t}"
                ::"r"(tmem_d), "r"(d_lo_c | ((sa16 + 2u * k) & 0x3FFFu)), "r"(d_hi), "r"(d_lo_c | ((sb16 + 2u * k) & 0x3FFFu)),
                "r"(idesc), "r"(accumulate)
                : "memory");
            accumulate = 1;
          }
          umma_commit(&empty_bar[stage]);
        }
        __syncwarp();
        accumulate = 1;
        if (++stage == kGemmStages) { stage = 0; phase ^= 1; }
      }
      if (elect_one()) umma_commit(&acc_full[slot]);
      __syncwarp();
    }
  } else {
    // ===================================================================== epilogue (warps 2..9)
    const int quarter = warp & 3;
    const int chalf = (warp - 2) >> 2;                      // which half of the accumulator's columns
    const bool vecN = (p.N % 16) == 0;
    int iter = 0;
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++iter) {
      const GemmTile t = gemm_tile<BN>(p, tile);
      const int slot = iter & 1;
      if (!mbar_wait_relaxed(&acc_full[slot], (iter >> 1) & 1, p.err_flag)) goto teardown;
      tc_fence_after();
      if (!(p.debug & 1))
        gemm_epilogue_tile<BN>(p, t, tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + static_cast<uint32_t>(slot * BN), quarter,
                               chalf, lane, vecN);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&acc_empty[slot]);
    }
  }

teardown:
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<2 * BN>(tmem_base);
  }
}

// ===================================================================================================================
// CTA-pair variant (tcgen05 cta_group::2): a cluster of two CTAs (one TPC) computes a 256 x BN tile.  CTA r holds rows
// m0 + 128 r of A and columns n0 + (BN/2) r of B per stage; ONE tcgen05.mma.cta_group::2 (M = 256), issued by the leader CTA,
// reads A from each CTA's own shared memory and the two B halves from both, and writes each CTA's 128 accumulator rows into its
// own TMEM.  Per CTA and K block that is 16 KB of A + BN x 64 B of B for 128 x BN x 64 MACs — half the L2 -> SM operand bytes
// per FLOP of the 1-CTA 128 x 128 tile, which is what bounds the K = 768 ViT linears.
// Barriers: both producers' TMA loads complete on the LEADER's full_bar (the leader's producer posts the byte count of both
// halves); tcgen05.commit multicasts to empty_bar / acc_full of both CTAs; the epilogue warps of both CTAs arrive on the
// leader's acc_empty (remote mbarrier arrive for the peer).
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t mapa_shared(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void mbar_arrive_remote(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// TMA load into this CTA's shared memory whose completion bytes are posted on a barrier of the CTA pair (cluster address)
__device__ __forceinline__ void tma_load_2d_pair(void* dst, const CUtensorMap* m, uint32_t bar_cluster_addr, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar_cluster_addr), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void umma_commit_pair(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(smem_u32(bar)), "h"(static_cast<uint16_t>(3)) : "memory");
}

template <int BN> struct GemmPairCfg {
  static constexpr uint32_t kStageB = (BN / 2) * BK * 2, kStage = kStageA + kStageB;
  static constexpr int kStages = BN == 256 ? 6 : 8;
};

template <int BN>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kGemmThreads, 1)
gemm_tc_pair_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                    const __grid_constant__ GemmParams p) {
  constexpr int kStages = GemmPairCfg<BN>::kStages;
  constexpr uint32_t kStage = GemmPairCfg<BN>::kStage;
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint64_t full_bar[kStages];
  __shared__ uint64_t empty_bar[kStages];
  __shared__ uint64_t acc_full[2], acc_empty[2];
  __shared__ uint32_t tmem_base_smem;
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int pair = blockIdx.x >> 1, npairs = gridDim.x >> 1;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_b);
    for (int i = 0; i < kStages; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&acc_full[i], 1); mbar_init(&acc_empty[i], 2 * kGemmEpiWarps); }
    fence_barrier_init();
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_smem)), "n"(2 * BN)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();               // both CTAs' barriers are initialised before any remote arrive / TMA completion
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_smem;

  auto tile_of = [&](int tile) {
    GemmTile t;
    const int tn = tile % p.tiles_n;
    const int tm = tile / p.tiles_n;
    t.sk = 0; t.bz = 0; t.kb0 = 0; t.kb1 = p.total_kb;
    t.m0 = tm * 2 * BM + static_cast<int>(rank) * BM;
    t.n0 = tn * BN;
    return t;
  };

  if (warp == 0) {
    // ===================================================================== TMA producer (both CTAs)
    int stage = 0;
    uint32_t phase = 0;
    for (int tile = pair; tile < p.num_tiles; tile += npairs) {
      const GemmTile t = tile_of(tile);
      for (int kb = 0; kb < p.total_kb; ++kb) {
        if (!mbar_wait_relaxed(&empty_bar[stage], phase ^ 1, p.err_flag)) goto teardown;
        if (elect_one()) {
          uint8_t* sa = smem + static_cast<size_t>(stage) * kStage;
          if (leader) mbar_arrive_expect_tx(&full_bar[stage], 2 * kStage);
          const uint32_t bar = mapa_shared(smem_u32(&full_bar[stage]), 0);
          tma_load_2d_pair(sa, &tmap_a, bar, kb * BK, t.m0);
          tma_load_2d_pair(sa + kStageA, &tmap_b, bar, kb * BK, t.n0 + static_cast<int>(rank) * (BN / 2));
        }
        __syncwarp();
        if (++stage == kStages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ===================================================================== MMA issuer (leader CTA only)
    if (leader) {
      const uint32_t idesc = make_idesc_f16(2 * BM, BN);
      const uint32_t d_hi = (1024u >> 4) | (1u << 14) | (2u << 29);
      const uint32_t d_lo_c = 1u << 16;
      int stage = 0;
      uint32_t phase = 0;
      int iter = 0;
      for (int tile = pair; tile < p.num_tiles; tile += npairs, ++iter) {
        const int slot = iter & 1;
        if (!mbar_wait(&acc_empty[slot], ((iter >> 1) & 1) ^ 1, p.err_flag)) goto teardown;
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + static_cast<uint32_t>(slot * BN);
        uint32_t accumulate = 0;
        for (int kb = 0; kb < p.total_kb; ++kb) {
          if (!mbar_wait(&full_bar[stage], phase, p.err_flag)) goto teardown;
          tc_fence_after();
          if (elect_one()) {
            const uint32_t sa16 = smem_u32(smem + static_cast<size_t>(stage) * kStage) >> 4;
            const uint32_t sb16 = sa16 + (kStageA >> 4);
#pragma unroll
            for (int k = 0; k < BK / 16; ++k) {
              asm volatile(
                  "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
                  "mov.b64 da, {%1, %2};\n\t"
                  "mov.b64 db, {%3, %2};\n\t"
                  "setp.ne.b32 p, %5, 0;\n\t"
                  "tcgen05.mma.cta_group::2.kind::f16 [%0], da, db, %4, p;\n\t}"
                  ::"r"(tmem_d), "r"(d_lo_c | ((sa16 + 2u * k) & 0x3FFFu)), "r"(d_hi), "r"(d_lo_c | ((sb16 + 2u * k) & 0x3FFFu)),
                  "r"(idesc), "r"(accumulate)
                  : "memory");
              accumulate = 1;
            }
            umma_commit_pair(&empty_bar[stage]);
          }
          __syncwarp();
          accumulate = 1;
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
        if (elect_one()) umma_commit_pair(&acc_full[slot]);
        __syncwarp();
      }
    }
  } else {
    // ===================================================================== epilogue (warps 2..9, both CTAs)
    const int quarter = warp & 3;
    const int chalf = (warp - 2) >> 2;
    const bool vecN = (p.N % 16) == 0;
    int iter = 0;
    for (int tile = pair; tile < p.num_tiles; tile += npairs, ++iter) {
      const GemmTile t = tile_of(tile);
      const int slot = iter & 1;
      if (!mbar_wait_relaxed(&acc_full[slot], (iter >> 1) & 1, p.err_flag)) goto teardown;
      tc_fence_after();
      gemm_epilogue_tile<BN>(p, t, tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + static_cast<uint32_t>(slot * BN), quarter,
                             chalf, lane, vecN);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (leader) mbar_arrive(&acc_empty[slot]);
        else mbar_arrive_remote(mapa_shared(smem_u32(&acc_empty[slot]), 0));
      }
    }
  }

teardown:
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();               // the peer may still be reading its accumulator / the leader's MMAs may still read peer smem
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(2 * BN) : "memory");
  }
}

template <int BN>
static int launch_gemm_pair(const void* A, const void* B, dp::GemmParams& p, cudaStream_t stream) {
  p.total_kb = (p.K + BK - 1) / BK;
  p.kb_per_split = p.total_kb;
  p.split_stride = 0;
  p.a_batch_rows = p.b_batch_rows = 0;
  p.tiles_m = (p.M + 2 * BM - 1) / (2 * BM);
  p.tiles_n = p.N / BN;
  p.num_tiles = p.tiles_m * p.tiles_n;
  CUtensorMap ta, tb;
  {
    const uint64_t dims[2] = {static_cast<uint64_t>(p.K), static_cast<uint64_t>(p.M)};
    const uint64_t strides[1] = {static_cast<uint64_t>(p.K) * 2};
    const uint32_t box[2] = {BK, BM};
    if (int rc = encode_tiled(&ta, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, A, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B))
      return rc;
  }
  {
    const uint64_t dims[2] = {static_cast<uint64_t>(p.K), static_cast<uint64_t>(p.N)};
    const uint64_t strides[1] = {static_cast<uint64_t>(p.K) * 2};
    const uint32_t box[2] = {BK, BN / 2};
    if (int rc = encode_tiled(&tb, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, B, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B))
      return rc;
  }
  const size_t smem = static_cast<size_t>(GemmPairCfg<BN>::kStages) * GemmPairCfg<BN>::kStage + 1024;
  if (first_use_on_device(BN == 256 ? KF_GEMM_PAIR256 : KF_GEMM_PAIR128)) {
    DP_CHECK(cudaFuncSetAttribute(gemm_tc_pair_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
  }
  int pairs = sm_count() / 2;
  if (pairs > p.num_tiles) pairs = p.num_tiles;
  gemm_tc_pair_kernel<BN><<<2 * pairs, kGemmThreads, smem, stream>>>(ta, tb, p);
  DP_CHECK(cudaGetLastError());
  return 0;
}

template <int BN>
static int launch_gemm_bn(const void* A, const void* B, dp::GemmParams& p, int a_batch_rows, int b_batch_rows,
                          cudaStream_t stream, const CUtensorMap* patch_map = nullptr) {
  DP_REQUIRE(p.K % 8 == 0, "dp_gemm_tc: K=%d must be a multiple of 8 (TMA 16-byte row pitch)", p.K);
  DP_REQUIRE(p.batch >= 1 && p.split_k >= 1, "dp_gemm_tc: bad batch/split_k");
  p.total_kb = (p.K + BK - 1) / BK;
  p.kb_per_split = (p.total_kb + p.split_k - 1) / p.split_k;
  p.split_k = (p.total_kb + p.kb_per_split - 1) / p.kb_per_split;      // no empty splits
  p.split_stride = static_cast<long long>(p.M) * p.ldc;
  p.a_batch_rows = a_batch_rows; p.b_batch_rows = b_batch_rows;
  p.tiles_m = (p.M + BM - 1) / BM;
  p.tiles_n = (p.N + BN - 1) / BN;
  p.num_tiles = p.tiles_m * p.tiles_n * p.batch * p.split_k;
  const uint64_t a_rows = static_cast<uint64_t>(p.batch > 1 && a_batch_rows ? (p.batch - 1) * static_cast<uint64_t>(a_batch_rows) + p.M : p.M);
  const uint64_t b_rows = static_cast<uint64_t>(p.batch > 1 && b_batch_rows ? (p.batch - 1) * static_cast<uint64_t>(b_batch_rows) + p.N : p.N);
  CUtensorMap ta, tb;
  if (p.patch_mode) {
    ta = *patch_map;
  } else {
    const uint64_t dims[2] = {static_cast<uint64_t>(p.K), a_rows};
    const uint64_t strides[1] = {static_cast<uint64_t>(p.K) * 2};
    const uint32_t box[2] = {BK, BM};
    if (int rc = encode_tiled(&ta, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, A, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B))
      return rc;
  }
  {
    const uint64_t dims[2] = {static_cast<uint64_t>(p.K), b_rows};
    const uint64_t strides[1] = {static_cast<uint64_t>(p.K) * 2};
    const uint32_t box[2] = {BK, BN};
    if (int rc = encode_tiled(&tb, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, B, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B))
      return rc;
  }
  const size_t smem = static_cast<size_t>(GemmCfg<BN>::kStages) * GemmCfg<BN>::kStage + 1024;
  if (first_use_on_device(BN == 256 ? KF_GEMM256 : KF_GEMM128)) {
    DP_CHECK(cudaFuncSetAttribute(gemm_tc_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
  }
  { static const int dbg = [] { const char* e = getenv("DP_GEMM_DEBUG"); return e ? atoi(e) : 0; }(); p.debug = dbg; }
  int grid = sm_count();
  if (grid > p.num_tiles) grid = p.num_tiles;
  gemm_tc_kernel<BN><<<grid, kGemmThreads, smem, stream>>>(ta, tb, p);
  DP_CHECK(cudaGetLastError());
  return 0;
}

static int launch_gemm(const void* A, const void* B, dp::GemmParams& p, int a_batch_rows, int b_batch_rows,
                       cudaStream_t stream, const CUtensorMap* patch_map = nullptr) {
  // CTA-pair tiles (cta_group::2) for the plain 2-D problems whose N the pair tile divides.  OFF by default: on the ViT
  // linears of this path (M = 4096, K = 768 / 3072) the pair kernel halves the L2 -> SM operand bytes and changes nothing
  // (profiles/r3_gemm_pair.md: 21.2 / 24.4 / 45.7 / 35.4 us per launch vs 22.1 / 24.7 / 45.0 / 36.4) — these launches are
  // bound by per-launch latency (one or two tiles per CTA), not by the operand stream.  DP_GEMM_PAIR = 1 (auto) / 128 / 256.
  static const int pair_mode = getenv("DP_GEMM_PAIR") ? atoi(getenv("DP_GEMM_PAIR")) : 0;
  if (pair_mode && !p.patch_mode && p.batch == 1 && p.split_k == 1 && p.M >= 2 * BM && p.K % BK == 0) {
    const bool can256 = p.N % 256 == 0, can128 = p.N % 128 == 0;
    const int pairs = sm_count() / 2;
    const long long t256 = static_cast<long long>((p.M + 255) / 256) * (p.N / 256);
    if (pair_mode == 256 ? can256 : (pair_mode == 128 ? false : (can256 && t256 >= pairs)))
      return launch_gemm_pair<256>(A, B, p, stream);
    if (pair_mode == 128 ? can128 : (pair_mode == 1 && can128))
      return launch_gemm_pair<128>(A, B, p, stream);
  }
  // wide-N problems with enough 256-wide tiles to fill the machine take the 128 x 256 tile
  static const bool wide_ok = getenv("DP_GEMM_BN256") == nullptr || atoi(getenv("DP_GEMM_BN256")) != 0;
  const long long tiles256 = static_cast<long long>((p.M + BM - 1) / BM) * ((p.N + 255) / 256) * p.batch * p.split_k;
  // measured (batch-8 ViT, M = 4096): N = 768, K = 131072: 714 -> 1096 TFLOP/s; N = 3072 / 2304, K = 768: +8 %;
  // N = 768, K = 768 (96 wide tiles, short K): -5 %, so under one wave of wide tiles only long-K problems take them
  static const int min_tiles = getenv("DP_GEMM_BN256_MIN") ? atoi(getenv("DP_GEMM_BN256_MIN")) : (sm_count() * 5) / 8;
  if (wide_ok && p.N % 256 == 0 && tiles256 >= min_tiles && (tiles256 >= sm_count() || p.K >= 2048))
    return launch_gemm_bn<256>(A, B, p, a_batch_rows, b_batch_rows, stream, patch_map);
  return launch_gemm_bn<128>(A, B, p, a_batch_rows, b_batch_rows, stream, patch_map);
}

}  // namespace dp

extern "C" int dp_gemm_tc(const void* A, const void* B, int M, int N, int K, int batch, int a_batch_rows,
                          int b_batch_rows, long long c_batch_stride, int c_batch_period, long long c_batch_stride2, int ldc,
                          int split_k, const float* bias, const float* rowvec, int row_period, const float* resid,
                          float alpha, int act, float* out_f32, int atomic, void* out_f16, int mode_qkv, int heads,
                          int hd, int T, int vt_ld, void* q, void* k, void* vt, float q_scale, int* err_flag, cudaStream_t stream) {
  using namespace dp;
  DP_REQUIRE(!atomic, "dp_gemm_tc: the atomic epilogue was removed (split-K is deterministic: dp_splitk_reduce)");
  DP_REQUIRE(split_k == 1 || (out_f32 && !out_f16 && !mode_qkv && act == 0 && !bias && !rowvec && !resid && batch == 1),
             "dp_gemm_tc: split-K writes plain fp32 partials [split_k][M][ldc] (finish with dp_splitk_reduce)");
  DP_REQUIRE(!mode_qkv || (hd % 16 == 0 && N == 3 * heads * hd), "dp_gemm_tc: qkv scatter needs head_dim %% 16 == 0");
  GemmParams p{};
  p.M = M; p.N = N; p.K = K; p.batch = batch; p.split_k = split_k;
  p.c_batch_stride = c_batch_stride; p.c_batch_period = c_batch_period; p.c_batch_stride2 = c_batch_stride2; p.ldc = ldc;
  p.bias = bias; p.rowvec = rowvec; p.row_period = row_period > 0 ? row_period : 1; p.resid = resid;
  p.alpha = alpha; p.act = act; p.out_f32 = out_f32; p.out_f16 = static_cast<__half*>(out_f16);
  p.mode = mode_qkv ? 1 : 0; p.heads = heads; p.hd = hd; p.T = T > 0 ? T : 1; p.vt_ld = vt_ld > 0 ? vt_ld : p.T;
  p.q = static_cast<__half*>(q); p.kk = static_cast<__half*>(k); p.vt = static_cast<__half*>(vt); p.q_scale = q_scale;
  p.err_flag = err_flag;
  return launch_gemm(A, B, p, a_batch_rows, b_batch_rows, stream);
}

extern "C" int dp_gemm_patch_embed(const void* in_c8, int cb_total, int cb_off, int ncb, int N, int D, int H, int W,
                                   const void* w_nk, int hidden, int split_k, const float* bias, const float* rowvec,
                                   int row_period, float* out_f32, int* err_flag, cudaStream_t stream) {
  using namespace dp;
  DP_REQUIRE(H == 128 && W == 128 && D % 32 == 0, "dp_gemm_patch_embed: gathers 2 x 8 x 8 patch slabs per 128-token tile: "
             "needs H = W = 128 and D %% 32 == 0 (got %dx%dx%d); use dp_patchify + dp_gemm_tc otherwise", D, H, W);
  DP_REQUIRE(split_k >= 1 && out_f32 != nullptr, "dp_gemm_patch_embed: fp32 output required");
  DP_REQUIRE(split_k == 1 || (bias == nullptr && rowvec == nullptr),
             "dp_gemm_patch_embed: split-K writes plain fp32 partials [split_k][M][hidden] (finish with dp_splitk_reduce)");
  GemmParams p{};
  const int tokens = (D / 16) * 64;
  p.M = N * tokens; p.N = hidden; p.K = ncb * 4096 * 8; p.batch = 1; p.split_k = split_k; p.ldc = hidden;
  p.bias = bias; p.rowvec = rowvec; p.row_period = row_period > 0 ? row_period : 1;
  p.alpha = 1.f; p.T = 1; p.out_f32 = out_f32; p.err_flag = err_flag;
  p.patch_mode = 1; p.patch_cb_total = cb_total; p.patch_cb_off = cb_off; p.patch_tokens = tokens;
  CUtensorMap ta;
  // (element strides are limited to 8, so the 16-voxel patch pitch is expressed through the byte strides instead)
  //   d0: the 128 fp16 of a 16-voxel W run (box: the 64 of one p3 half)      d1: gz, 256 B apart
  //   d2: gy, 16 rows apart     d3: gx, 16 planes apart     d4: every (image, block, plane, row) of the tensor, one row
  //   (W * 16 B) apart: its coordinate carries the K block's (n, cb, p1, p2) offset
  const uint64_t dims[5] = {128, static_cast<uint64_t>(W / 16), static_cast<uint64_t>(H / 16), static_cast<uint64_t>(D / 16),
                            static_cast<uint64_t>(N) * cb_total * D * H};
  const uint64_t strides[4] = {256, static_cast<uint64_t>(W) * 16 * 16, static_cast<uint64_t>(H) * W * 16 * 16,
                               static_cast<uint64_t>(W) * 16};
  const uint32_t box[5] = {64, 8, 8, 2, 1};
  if (int rc = encode_tiled(&ta, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 5, in_c8, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B))
    return rc;
  p.patch_D = D; p.patch_H = H;
  return launch_gemm(nullptr, w_nk, p, 0, 0, stream, &ta);
}

extern "C" int dp_deconv2x_gemm(const void* tokens, const void* w_nk, int B, int Dg, int Hg, int Wg, int cin, int cout,
                                void* out_hi, void* out_lo, int out_cb_total, int out_cb_off, int* err_flag,
                                cudaStream_t stream) {
  using namespace dp;
  DP_REQUIRE(cout % 16 == 0 && cin % 8 == 0, "dp_deconv2x_gemm: needs C_out %% 16 == 0 and C_in %% 8 == 0 (got %d, %d)", cout, cin);
  GemmParams p{};
  p.M = B * Dg * Hg * Wg; p.N = 8 * cout; p.K = cin; p.batch = 1; p.split_k = 1; p.ldc = p.N;
  p.row_period = 1; p.alpha = 1.f; p.T = 1;
  p.mode = 2; p.Dg = Dg; p.Hg = Hg; p.Wg = Wg; p.cout = cout;
  p.dc_hi = static_cast<__half*>(out_hi); p.dc_lo = static_cast<__half*>(out_lo);
  p.dc_cb_total = out_cb_total; p.dc_cb_off = out_cb_off; p.err_flag = err_flag;
  return launch_gemm(tokens, w_nk, p, 0, 0, stream);
}
