// softmax(Q K^T) V for one (image, head) and 128 queries per CTA, on tcgen05 tensor cores: the scores and the
// probabilities never leave the SM (monai SABlock.forward: einsum "blxd,blyd->blxy" * scale, softmax(-1),
// einsum "bhxy,bhyd->bhxd", rearrange "b h l d -> b l (h d)").
//
// Exact two-pass softmax instead of an online rescale: pass 1 recomputes S = Q K_j^T per 128-key block and keeps
// only the running row maximum m (thread = query row, so no shuffles and no exponentials); pass 2 recomputes S,
// writes P = exp(S - m) as fp16 into a 128B-swizzled K-major smem tile and accumulates O += P V_j in TMEM.  The
// row sum l is accumulated in fp32 next to it, and O is divided by l in the epilogue.  QK^T is 2 % of the
// ViT's FLOPs, so computing it twice is cheaper than a TMEM round trip of O per key block.
//
// Operands: q, k [BH][T][HD] fp16 (q pre-scaled by hd^-0.5 in the qkv GEMM epilogue), v^T [BH][HD][Tp] fp16 (zero
// padded key axis, 8 | Tp), all read through 3-D TMA maps so that rows / keys past T are zero-filled per head.
// Warp roles (320 threads): warp 0 TMA producer, warp 1 MMA issuer + TMEM owner, warps 2..9 softmax + epilogue: thread <->
// TMEM lane <-> query row, and the TWO warps of a lane quarter split the 128 key columns of a score block (and the columns
// of O) in halves — the softmax warps were the issue-bound part (ncu: 47 % issue-active at four warps); the halves' row
// maxima / row sums are combined once per pass through shared memory.  TMEM: 128 columns S + HD columns O (256 allocated:
// two CTAs per SM).
#include "common.cuh"
#include "dose_b200.h"

namespace dp {

struct AttnParams {
  int T, heads, ld_out;         // tokens per image, heads, row pitch of the merged output (= heads * HD)
  __half* out;                  // [B * T][ld_out]
  int* err_flag;
};

constexpr int kAttnThreads = 320;
constexpr int kAttnSoftmaxWarps = 8;
constexpr int kAttnKStages = 2;

template <int HD> struct AttnCfg {
  static constexpr uint32_t kChunksD = HD / 64;                  // 64-element (128-byte) K chunks of the head dim
  static constexpr uint32_t kQBytes = 128 * HD * 2;
  static constexpr uint32_t kKBytes = 128 * HD * 2;              // one 128-key block of K
  static constexpr uint32_t kVBytes = HD * 128 * 2;              // one 128-key block of V^T: 2 chunks of [HD][64]
  static constexpr uint32_t kPBytes = 128 * 128 * 2;
  static constexpr uint32_t kSmem = kQBytes + kAttnKStages * kKBytes + kVBytes + kPBytes + 1024;
};

__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}

// one K = 16 step of a 128B-swizzled K-major operand pair (see gemm_tc.cu)
__device__ __forceinline__ void attn_umma(uint32_t tmem_d, uint32_t a16, uint32_t b16, uint32_t idesc, uint32_t accumulate) {
  const uint32_t d_hi = (1024u >> 4) | (1u << 14) | (2u << 29);
  const uint32_t d_lo_c = 1u << 16;
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
      "mov.b64 da, {%1, %2};\n\t"
      "mov.b64 db, {%3, %2};\n\t"
      "setp.ne.b32 p, %5, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %4, p;\n\t}"
      ::"r"(tmem_d), "r"(d_lo_c | (a16 & 0x3FFFu)), "r"(d_hi), "r"(d_lo_c | (b16 & 0x3FFFu)), "r"(idesc), "r"(accumulate)
      : "memory");
}

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// keys >= valid of a 32-column score slab -> -inf
__device__ __forceinline__ void mask_keys(uint32_t (&r)[2][16], int c0, int valid) {
#pragma unroll
  for (int u = 0; u < 2; ++u)
#pragma unroll
    for (int i = 0; i < 16; ++i)
      if (c0 + u * 16 + i >= valid) r[u][i] = 0xff800000u;
}

template <int HD>
__global__ void __launch_bounds__(kAttnThreads, 1)
attention_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_k,
                 const __grid_constant__ CUtensorMap tmap_vt, const __grid_constant__ AttnParams p) {
  using Cfg = AttnCfg<HD>;
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint64_t q_full, k_full[kAttnKStages], k_empty[kAttnKStages], v_full, v_empty;
  __shared__ uint64_t s_full, s_empty, p_full, p_empty, o_full;
  __shared__ uint32_t tmem_base_smem;
  __shared__ float xch[2][128];                                     // row maximum / row sum of the other column half
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sQ = smem;
  uint8_t* sK = sQ + Cfg::kQBytes;
  uint8_t* sV = sK + kAttnKStages * Cfg::kKBytes;
  uint8_t* sP = sV + Cfg::kVBytes;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int bh = blockIdx.y, q0 = blockIdx.x * 128;
  const int nkb = (p.T + 127) >> 7;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_q);
    tma_prefetch_desc(&tmap_k);
    tma_prefetch_desc(&tmap_vt);
    mbar_init(&q_full, 1);
    for (int i = 0; i < kAttnKStages; ++i) { mbar_init(&k_full[i], 1); mbar_init(&k_empty[i], 1); }
    mbar_init(&v_full, 1); mbar_init(&v_empty, 1);
    mbar_init(&s_full, 1); mbar_init(&s_empty, kAttnSoftmaxWarps);
    mbar_init(&p_full, kAttnSoftmaxWarps); mbar_init(&p_empty, 1);
    mbar_init(&o_full, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<256>(&tmem_base_smem);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_s = tmem_base_smem, tmem_o = tmem_base_smem + 128;

  if (warp == 0) {
    // ===================================================================== TMA producer
    if (elect_one()) {
      mbar_arrive_expect_tx(&q_full, Cfg::kQBytes);
      for (uint32_t c = 0; c < Cfg::kChunksD; ++c) tma_load_3d(sQ + c * 16384, &tmap_q, &q_full, c * 64, q0, bh);
    }
    __syncwarp();
    for (int it = 0; it < 2 * nkb; ++it) {
      const int j = it < nkb ? it : it - nkb;
      const int stage = it % kAttnKStages;
      if (!mbar_wait_relaxed(&k_empty[stage], ((it / kAttnKStages) & 1) ^ 1, p.err_flag)) goto teardown;
      if (elect_one()) {
        uint8_t* dst = sK + stage * Cfg::kKBytes;
        mbar_arrive_expect_tx(&k_full[stage], Cfg::kKBytes);
        for (uint32_t c = 0; c < Cfg::kChunksD; ++c) tma_load_3d(dst + c * 16384, &tmap_k, &k_full[stage], c * 64, j * 128, bh);
      }
      __syncwarp();
      if (it >= nkb) {
        if (!mbar_wait_relaxed(&v_empty, (j & 1) ^ 1, p.err_flag)) goto teardown;
        if (elect_one()) {
          mbar_arrive_expect_tx(&v_full, Cfg::kVBytes);
          for (int c = 0; c < 2; ++c) tma_load_3d(sV + c * (HD * 128), &tmap_vt, &v_full, j * 128 + c * 64, 0, bh);
        }
        __syncwarp();
      }
    }
  } else if (warp == 1) {
    // ===================================================================== MMA issuer
    const uint32_t idesc_s = make_idesc_f16(128, 128), idesc_o = make_idesc_f16(128, HD);
    const uint32_t q16 = smem_u32(sQ) >> 4, p16 = smem_u32(sP) >> 4, v16 = smem_u32(sV) >> 4;
    // S(it) = Q K_j^T; `it` counts score blocks over both passes
    auto issue_qk = [&](int it) -> bool {
      const int stage = it % kAttnKStages;
      if (!mbar_wait(&s_empty, (it & 1) ^ 1, p.err_flag)) return false;
      if (!mbar_wait(&k_full[stage], (it / kAttnKStages) & 1, p.err_flag)) return false;
      tc_fence_after();
      if (elect_one()) {
        const uint32_t k16 = smem_u32(sK + stage * Cfg::kKBytes) >> 4;
#pragma unroll
        for (uint32_t c = 0; c < Cfg::kChunksD; ++c)
#pragma unroll
          for (uint32_t ks = 0; ks < 4; ++ks)
            attn_umma(tmem_s, q16 + c * 1024 + 2 * ks, k16 + c * 1024 + 2 * ks, idesc_s, (c | ks) ? 1u : 0u);
        umma_commit(&k_empty[stage]);
        umma_commit(&s_full);
      }
      __syncwarp();
      return true;
    };
    if (!mbar_wait(&q_full, 0, p.err_flag)) goto teardown;
    for (int it = 0; it < nkb; ++it)
      if (!issue_qk(it)) goto teardown;
    if (!issue_qk(nkb)) goto teardown;
    for (int j = 0; j < nkb; ++j) {
      if (j + 1 < nkb && !issue_qk(nkb + j + 1)) goto teardown;       // overlaps the softmax warps' exp / P stores of block j
      if (!mbar_wait(&p_full, j & 1, p.err_flag)) goto teardown;
      if (!mbar_wait(&v_full, j & 1, p.err_flag)) goto teardown;
      tc_fence_after();
      if (elect_one()) {
#pragma unroll
        for (uint32_t c = 0; c < 2; ++c)
#pragma unroll
          for (uint32_t ks = 0; ks < 4; ++ks)
            attn_umma(tmem_o, p16 + c * 1024 + 2 * ks, v16 + c * (HD * 128 / 16) + 2 * ks, idesc_o, (j | c | ks) ? 1u : 0u);
        umma_commit(&v_empty);
        umma_commit(&p_empty);
        if (j + 1 == nkb) umma_commit(&o_full);
      }
      __syncwarp();
    }
  } else {
    // ===================================================================== softmax + epilogue (warps 2..9)
    const int quarter = warp & 3;
    const int half = (warp - 2) >> 2;                               // key columns [64 half, 64 half + 64) of every score block
    const int row = quarter * 32 + lane;                            // TMEM lane == query row of the tile
    const uint32_t t_s = tmem_s + (static_cast<uint32_t>(quarter * 32) << 16);
    const uint32_t t_o = tmem_o + (static_cast<uint32_t>(quarter * 32) << 16);
    constexpr float kLog2e = 1.4426950408889634f;
    float m = -INFINITY;
    // ---- pass 1: row maximum
    for (int it = 0; it < nkb; ++it) {
      const int valid = min(128, p.T - it * 128);
      if (!mbar_wait(&s_full, it & 1, p.err_flag)) goto teardown;
      tc_fence_after();
#pragma unroll 1
      for (int c0 = half * 64; c0 < half * 64 + 64; c0 += 32) {
        uint32_t r[2][16];
        tmem_ld16(t_s + c0, r[0]);
        tmem_ld16(t_s + c0 + 16, r[1]);
        tmem_ld_wait();
        if (valid < 128) mask_keys(r, c0, valid);                   // warp-uniform: only the ragged last block
#pragma unroll
        for (int u = 0; u < 2; ++u)
#pragma unroll
          for (int i = 0; i < 16; ++i) m = fmaxf(m, __uint_as_float(r[u][i]));
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&s_empty);
    }
    // the row maximum over both column halves
    xch[half][row] = m;
    asm volatile("bar.sync 1, 256;" ::: "memory");
    m = fmaxf(m, xch[half ^ 1][row]);
    asm volatile("bar.sync 1, 256;" ::: "memory");                  // xch is reused for the row sums
    // ---- pass 2: P = exp(S - m) -> smem, l = sum P
    const float mb = m * kLog2e;
    float l0 = 0.f, l1 = 0.f;
    uint8_t* prow = sP + row * 128;
    for (int j = 0; j < nkb; ++j) {
      const int it = nkb + j;
      const int valid = min(128, p.T - j * 128);
      if (!mbar_wait(&s_full, it & 1, p.err_flag)) goto teardown;
      tc_fence_after();
      if (!mbar_wait(&p_empty, (j & 1) ^ 1, p.err_flag)) goto teardown;     // P V_{j-1} has read the tile
#pragma unroll 1
      for (int c0 = half * 64; c0 < half * 64 + 64; c0 += 32) {
        uint32_t r[2][16];
        tmem_ld16(t_s + c0, r[0]);
        tmem_ld16(t_s + c0 + 16, r[1]);
        tmem_ld_wait();
        if (c0 == half * 64 + 32) {                                 // S is in registers: release it for Q K_{j+1}^T
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&s_empty);
        }
        if (valid < 128) mask_keys(r, c0, valid);                   // exp2(-inf) = 0
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          __align__(16) __half2 h[8];
#pragma unroll
          for (int i = 0; i < 16; i += 2) {
            const float e0 = ex2_approx(fmaf(__uint_as_float(r[u][i]), kLog2e, -mb));
            const float e1 = ex2_approx(fmaf(__uint_as_float(r[u][i + 1]), kLog2e, -mb));
            h[i >> 1] = __floats2half2_rn(e0, e1);
            l0 += e0;                                               // two chains; the fp16 rounding of P averages out
            l1 += e1;                                               // over the row (relative 2^-12 / sqrt(T))
          }
          // keys c0 + u*16 .. +15 = 16-byte units 2*(c0%64/16 ...) of chunk c0/64; 128B swizzle: unit ^= row & 7
          const int chunk = c0 >> 6, unit0 = ((c0 & 63) >> 3) + u * 2;
          uint8_t* base = prow + chunk * 16384;
          *reinterpret_cast<uint4*>(base + (((unit0) ^ (row & 7)) << 4)) = *reinterpret_cast<const uint4*>(&h[0]);
          *reinterpret_cast<uint4*>(base + (((unit0 + 1) ^ (row & 7)) << 4)) = *reinterpret_cast<const uint4*>(&h[4]);
        }
      }
      fence_proxy_async();                                          // generic-proxy smem writes -> visible to the MMA
      __syncwarp();
      if (lane == 0) mbar_arrive(&p_full);
    }
    // ---- epilogue: O / l -> merged heads
    if (!mbar_wait_relaxed(&o_full, 0, p.err_flag)) goto teardown;
    tc_fence_after();
    {
      xch[half][row] = l0 + l1;
      asm volatile("bar.sync 1, 256;" ::: "memory");
      const float inv_l = 1.f / (xch[0][row] + xch[1][row]);
      const int t = q0 + row;
      const int b = bh / p.heads, hh = bh % p.heads;
      __half* dst = p.out + (static_cast<size_t>(b) * p.T + t) * p.ld_out + hh * HD;
#pragma unroll 1
      for (int c0 = half * (HD / 2); c0 < (half + 1) * (HD / 2); c0 += 32) {
        uint32_t r[2][16];
        tmem_ld16(t_o + c0, r[0]);
        tmem_ld16(t_o + c0 + 16, r[1]);
        tmem_ld_wait();
        if (t < p.T) {
#pragma unroll
          for (int u = 0; u < 2; ++u) {
            __align__(16) __half2 h[8];
#pragma unroll
            for (int i = 0; i < 16; i += 2)
              h[i >> 1] = __floats2half2_rn(__uint_as_float(r[u][i]) * inv_l, __uint_as_float(r[u][i + 1]) * inv_l);
            *reinterpret_cast<uint4*>(dst + c0 + u * 16) = *reinterpret_cast<const uint4*>(&h[0]);
            *reinterpret_cast<uint4*>(dst + c0 + u * 16 + 8) = *reinterpret_cast<const uint4*>(&h[4]);
          }
        }
      }
    }
  }

teardown:
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<256>(tmem_base_smem);
  }
}

template <int HD>
static int launch_attention(const void* q, const void* k, const void* vt, int BH, int T, int Tp, const AttnParams& p,
                            cudaStream_t stream) {
  CUtensorMap tq, tk, tv;
  {
    const uint64_t dims[3] = {static_cast<uint64_t>(HD), static_cast<uint64_t>(T), static_cast<uint64_t>(BH)};
    const uint64_t strides[2] = {static_cast<uint64_t>(HD) * 2, static_cast<uint64_t>(T) * HD * 2};
    const uint32_t box[3] = {64, 128, 1};
    if (int rc = encode_tiled(&tq, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, q, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B)) return rc;
    if (int rc = encode_tiled(&tk, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, k, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B)) return rc;
  }
  {
    const uint64_t dims[3] = {static_cast<uint64_t>(Tp), static_cast<uint64_t>(HD), static_cast<uint64_t>(BH)};
    const uint64_t strides[2] = {static_cast<uint64_t>(Tp) * 2, static_cast<uint64_t>(Tp) * HD * 2};
    const uint32_t box[3] = {64, HD, 1};
    if (int rc = encode_tiled(&tv, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, vt, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B)) return rc;
  }
  if (first_use_on_device(HD == 64 ? KF_ATTN64 : KF_ATTN128)) {
    DP_CHECK(cudaFuncSetAttribute(attention_kernel<HD>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  static_cast<int>(AttnCfg<HD>::kSmem)));
  }
  dim3 grid((T + 127) / 128, BH);
  attention_kernel<HD><<<grid, kAttnThreads, AttnCfg<HD>::kSmem, stream>>>(tq, tk, tv, p);
  DP_CHECK(cudaGetLastError());
  return 0;
}

}  // namespace dp

extern "C" int dp_attention(const void* q, const void* k, const void* vt, int batch, int heads, int T, int vt_ld, int hd,
                            void* out, int ld_out, int* err_flag, cudaStream_t stream) {
  using namespace dp;
  DP_REQUIRE(hd == 64 || hd == 128, "dp_attention: head dim %d not instantiated (64, 128); use the dp_gemm_tc + dp_softmax path", hd);
  DP_REQUIRE(batch >= 1 && heads >= 1 && T >= 1 && batch * heads <= 65535, "dp_attention: bad batch/heads/T");
  DP_REQUIRE(vt_ld >= T && vt_ld % 8 == 0, "dp_attention: v^T row pitch %d must be >= T and a multiple of 8", vt_ld);
  DP_REQUIRE(ld_out % 8 == 0 && ld_out >= heads * hd, "dp_attention: output pitch %d", ld_out);
  AttnParams p;
  p.T = T; p.heads = heads; p.ld_out = ld_out; p.out = static_cast<__half*>(out); p.err_flag = err_flag;
  if (hd == 64) return launch_attention<64>(q, k, vt, batch * heads, T, vt_ld, p, stream);
  return launch_attention<128>(q, k, vt, batch * heads, T, vt_ld, p, stream);
}
