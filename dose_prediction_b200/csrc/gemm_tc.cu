// C[M,N] = A[M,K] * B[N,K]^T on tcgen05 tensor cores (fp16 operands, fp32 accumulation in TMEM).
//
// Used for every nn.Linear on the hot path (monai PatchEmbeddingBlock "perceptron" Linear, SABlock
// qkv/out_proj, MLPBlock linear1/linear2) and for the two attention contractions
// (einsum "blxd,blyd->blxy" and "bhxy,bhyd->bhxd" in monai SABlock.forward).
//
// Tiles: 128 x 128 x 64, TMA 2-D boxes with 128-byte swizzle, 6-stage mbarrier ring.
// Warp roles (192 threads): warp 0 TMA producer, warp 1 MMA issuer + TMEM owner, warps 2..5 epilogue.
// grid = (tiles_n, tiles_m, batch * split_k).  Epilogue options: +bias[n], +rowvec[m % period][n]
// (position embeddings), GELU(erf), *alpha, +residual (fp32, may alias out_f32), fp32 store or
// atomicAdd (split-K), fp16 copy, or the q/k/v^T head scatter.
#include "common.cuh"
#include "dose_b200.h"

namespace dp {

struct GemmParams {
  int M, N, K;
  int batch, split_k, kb_per_split;
  long long split_stride;              // split-K: partial sums of split s go to out_f32 + s*split_stride (deterministic)
  int a_batch_rows, b_batch_rows;      // row offset per batch entry inside the A / B tensor maps
  long long c_batch_stride;            // element offset per batch entry in the outputs ...
  int c_batch_period; long long c_batch_stride2;   // ... or (z / period) * stride + (z % period) * stride2
  int ldc;
  const float* bias;
  const float* rowvec; int row_period;
  const float* resid;
  float alpha;
  int act;
  float* out_f32; int atomic;
  __half* out_f16;
  // qkv scatter (mode_qkv): N = 3*heads*hd, tokens per image T
  int mode_qkv, heads, hd, T;
  __half* q; __half* kk; __half* vt; float q_scale;
  int* err_flag;
};

constexpr int kGemmThreads = 192;
constexpr int BM = 128, BN = 128, BK = 64;
constexpr int kGemmStages = 6;
constexpr uint32_t kStageA = BM * BK * 2, kStageB = BN * BK * 2, kStage = kStageA + kStageB;

__device__ __forceinline__ void gemm_producer(const CUtensorMap* ta, const CUtensorMap* tb, const GemmParams& p,
                                              uint8_t* smem, uint64_t* full_bar, uint64_t* empty_bar, int m0, int n0,
                                              int a_row0, int b_row0, int kb0, int kb1) {
  int stage = 0;
  uint32_t phase = 0;
  for (int kb = kb0; kb < kb1; ++kb) {
    if (!mbar_wait(&empty_bar[stage], phase ^ 1, p.err_flag)) return;
    uint8_t* sa = smem + static_cast<size_t>(stage) * kStage;
    mbar_arrive_expect_tx(&full_bar[stage], kStage);
    tma_load_2d(sa, ta, &full_bar[stage], kb * BK, a_row0 + m0);
    tma_load_2d(sa + kStageA, tb, &full_bar[stage], kb * BK, b_row0 + n0);
    if (++stage == kGemmStages) { stage = 0; phase ^= 1; }
  }
}

__device__ __forceinline__ void gemm_mma(const GemmParams& p, uint8_t* smem, uint64_t* full_bar, uint64_t* empty_bar,
                                         uint64_t* acc_bar, uint32_t tmem_d, int kb0, int kb1) {
  const uint32_t idesc = make_idesc_f16(BM, BN);
  int stage = 0;
  uint32_t phase = 0;
  uint32_t accumulate = 0;
  for (int kb = kb0; kb < kb1; ++kb) {
    if (!mbar_wait(&full_bar[stage], phase, p.err_flag)) return;
    tc_fence_after();
    const uint32_t sa = smem_u32(smem + static_cast<size_t>(stage) * kStage);
    const uint32_t sb = sa + kStageA;
#pragma unroll
    for (int k = 0; k < BK / 16; ++k) {
      // 128B-swizzled K-major tile: 8-row atoms of 1024 B (SBO), advance 32 B per 16-element K step
      umma_f16_ss(tmem_d, make_smem_desc(sa + k * 32, 16, 1024, 2), make_smem_desc(sb + k * 32, 16, 1024, 2), idesc,
                  accumulate);
      accumulate = 1;
    }
    umma_commit(&empty_bar[stage]);
    if (++stage == kGemmStages) { stage = 0; phase ^= 1; }
  }
  umma_commit(acc_bar);
}

__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
               const __grid_constant__ GemmParams p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint64_t full_bar[kGemmStages];
  __shared__ uint64_t empty_bar[kGemmStages];
  __shared__ uint64_t acc_bar;
  __shared__ uint32_t tmem_base_smem;
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n0 = blockIdx.x * BN, m0 = blockIdx.y * BM;
  const int z = blockIdx.z;
  const int bz = z / p.split_k, sk = z % p.split_k;
  const int total_kb = (p.K + BK - 1) / BK;
  const int kb0 = sk * p.kb_per_split;
  const int kb1 = min(total_kb, kb0 + p.kb_per_split);

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_b);
    for (int i = 0; i < kGemmStages; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], 1); }
    mbar_init(&acc_bar, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<BN>(&tmem_base_smem);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_smem;

  if (kb0 < kb1) {
    if (warp == 0) {
      if (lane == 0)
        gemm_producer(&tmap_a, &tmap_b, p, smem, full_bar, empty_bar, m0, n0, bz * p.a_batch_rows, bz * p.b_batch_rows,
                      kb0, kb1);
    } else if (warp == 1) {
      if (lane == 0) gemm_mma(p, smem, full_bar, empty_bar, &acc_bar, tmem_base, kb0, kb1);
    } else {
      const int quarter = warp & 3;
      const int m = m0 + quarter * 32 + lane;
      if (mbar_wait(&acc_bar, 0, p.err_flag)) {
        tc_fence_after();
        const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16);
        const bool lead = (sk == 0);
        for (int c0 = 0; c0 < BN; c0 += 16) {
          if (n0 + c0 >= p.N) break;                         // warp-uniform
          uint32_t r[16];
          tmem_ld16(taddr + c0, r);
          tmem_ld_wait();
          if (m >= p.M) continue;
          float v[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const int n = n0 + c0 + j;
            float x = __uint_as_float(r[j]) * p.alpha;
            if (n < p.N) {
              if (lead && p.bias) x += __ldg(&p.bias[n]);
              if (lead && p.rowvec) x += __ldg(&p.rowvec[static_cast<size_t>(m % p.row_period) * p.N + n]);
              x = act_apply(x, p.act);
            }
            v[j] = x;
          }
          if (p.mode_qkv) {
            const int b = m / p.T, t = m % p.T;
            const int hidden = p.heads * p.hd;
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              const int n = n0 + c0 + j;
              if (n >= p.N) break;
              const int which = n / hidden, rem = n % hidden;
              const int hh = rem / p.hd, dd = rem % p.hd;
              const size_t bh = static_cast<size_t>(b) * p.heads + hh;
              if (which == 0) p.q[(bh * p.T + t) * p.hd + dd] = __float2half_rn(v[j] * p.q_scale);
              else if (which == 1) p.kk[(bh * p.T + t) * p.hd + dd] = __float2half_rn(v[j]);
              else p.vt[(bh * p.hd + dd) * p.T + t] = __float2half_rn(v[j]);
            }
            continue;
          }
          const size_t boff = p.c_batch_period > 0
              ? static_cast<size_t>(bz / p.c_batch_period) * p.c_batch_stride + static_cast<size_t>(bz % p.c_batch_period) * p.c_batch_stride2
              : static_cast<size_t>(bz) * p.c_batch_stride;
          const size_t base = boff + static_cast<size_t>(sk) * p.split_stride + static_cast<size_t>(m) * p.ldc + n0 + c0;
          const bool full = (n0 + c0 + 16 <= p.N);
          if (p.resid && lead) {
#pragma unroll
            for (int j = 0; j < 16; ++j)
              if (full || n0 + c0 + j < p.N) v[j] += p.resid[base + j];
          }
          if (p.out_f32) {
            if (p.atomic) {
#pragma unroll
              for (int j = 0; j < 16; ++j)
                if (full || n0 + c0 + j < p.N) atomicAdd(&p.out_f32[base + j], v[j]);
            } else if (full && (p.ldc % 4 == 0)) {
#pragma unroll
              for (int j = 0; j < 16; j += 4)
                *reinterpret_cast<float4*>(&p.out_f32[base + j]) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
            } else {
#pragma unroll
              for (int j = 0; j < 16; ++j)
                if (n0 + c0 + j < p.N) p.out_f32[base + j] = v[j];
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
                if (n0 + c0 + j < p.N) p.out_f16[base + j] = __float2half_rn(v[j]);
            }
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<BN>(tmem_base);
  }
}

}  // namespace dp

extern "C" int dp_gemm_tc(const void* A, const void* B, int M, int N, int K, int batch, int a_batch_rows,
                          int b_batch_rows, long long c_batch_stride, int c_batch_period, long long c_batch_stride2, int ldc,
                          int split_k, const float* bias,
                          const float* rowvec, int row_period, const float* resid, float alpha, int act,
                          float* out_f32, int atomic, void* out_f16, int mode_qkv, int heads, int hd, int T, void* q,
                          void* k, void* vt, float q_scale, int* err_flag, cudaStream_t stream) {
  using namespace dp;
  DP_REQUIRE(K % 8 == 0, "dp_gemm_tc: K=%d must be a multiple of 8 (TMA 16-byte row pitch)", K);
  DP_REQUIRE(batch >= 1 && split_k >= 1, "dp_gemm_tc: bad batch/split_k");
  DP_REQUIRE(split_k == 1 || (out_f32 && !atomic && !out_f16 && !mode_qkv && act == 0 && !bias && !rowvec && !resid && batch == 1),
             "dp_gemm_tc: split-K writes plain fp32 partials [split_k][M][ldc] (finish with dp_splitk_reduce)");
  GemmParams p{};
  p.M = M; p.N = N; p.K = K; p.batch = batch; p.split_k = split_k;
  const int total_kb = (K + BK - 1) / BK;
  p.kb_per_split = (total_kb + split_k - 1) / split_k;
  p.split_stride = static_cast<long long>(M) * ldc;
  p.a_batch_rows = a_batch_rows; p.b_batch_rows = b_batch_rows; p.c_batch_stride = c_batch_stride; p.c_batch_period = c_batch_period; p.c_batch_stride2 = c_batch_stride2; p.ldc = ldc;
  p.bias = bias; p.rowvec = rowvec; p.row_period = row_period > 0 ? row_period : 1; p.resid = resid;
  p.alpha = alpha; p.act = act; p.out_f32 = out_f32; p.atomic = atomic; p.out_f16 = static_cast<__half*>(out_f16);
  p.mode_qkv = mode_qkv; p.heads = heads; p.hd = hd; p.T = T > 0 ? T : 1;
  p.q = static_cast<__half*>(q); p.kk = static_cast<__half*>(k); p.vt = static_cast<__half*>(vt); p.q_scale = q_scale;
  p.err_flag = err_flag;

  const uint64_t a_rows = static_cast<uint64_t>(batch > 1 && a_batch_rows ? (batch - 1) * static_cast<uint64_t>(a_batch_rows) + M : M);
  const uint64_t b_rows = static_cast<uint64_t>(batch > 1 && b_batch_rows ? (batch - 1) * static_cast<uint64_t>(b_batch_rows) + N : N);
  CUtensorMap ta, tb;
  {
    const uint64_t dims[2] = {static_cast<uint64_t>(K), a_rows};
    const uint64_t strides[1] = {static_cast<uint64_t>(K) * 2};
    const uint32_t box[2] = {BK, BM};
    if (int rc = encode_tiled(&ta, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, A, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B))
      return rc;
  }
  {
    const uint64_t dims[2] = {static_cast<uint64_t>(K), b_rows};
    const uint64_t strides[1] = {static_cast<uint64_t>(K) * 2};
    const uint32_t box[2] = {BK, BN};
    if (int rc = encode_tiled(&tb, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, B, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B))
      return rc;
  }
  const size_t smem = static_cast<size_t>(kGemmStages) * kStage + 1024;
  static bool configured = false;
  if (!configured) {
    DP_CHECK(cudaFuncSetAttribute(gemm_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    configured = true;
  }
  dim3 grid((N + BN - 1) / BN, (M + BM - 1) / BM, batch * split_k);
  gemm_tc_kernel<<<grid, kGemmThreads, smem, stream>>>(ta, tb, p);
  DP_CHECK(cudaGetLastError());
  return 0;
}
