// Convolution weight gradient on tcgen05 tensor cores (training step).
//
//   dW[co][ci][kd][kh][kw] = sum_{n,d,h,w} g[n][co][d][h][w] * x[n][ci][d+kd-p][h+kh-p][w+kw-p]
//
// The reduction runs over voxels, so both MMA operands are read "MN-major" straight from the c8 activation
// layout: along W, 8 consecutive voxels x 8 channels (16 B each) are exactly one canonical no-swizzle core
// matrix with K = voxel.  One tcgen05.mma (M=128, N=k*16, K=16 voxels) computes
//   D[(s, c)][(j, co)] += sum_w x[d_x][h_x][w + s - p][c] * g[d_x - kd + p][h_x + j - p][w][co]
//     - M side: 16 W-shifts s of ONE 8-channel block of the x row; the shift is free: the descriptor's
//       M-block stride (SBO) is 16 bytes = one voxel, so the 16 "blocks" are the same shared-memory row read at
//       16 overlapping offsets (rows s >= k are computed and discarded);
//     - N side: the k rows h_g = h_x + j - p of the g plane (kh = k-1-j) x 16 output channels; rows live in a
//       linear block of RB + k - 1 rows in shared memory, so the k rows of any x row are an affine slice.
// A CTA owns one (16-channel chunk of C_in, 16-channel tile of C_out, group of KDP depth taps) and keeps its
// 2*KDP accumulators (ci block x kd) in TMEM for its whole share of the volume; partial results of the `splits`
// CTAs of a type go to ws[split] and are summed by dp_splitk_reduce (deterministic).
// Warp roles: warp 0 TMA producer, warps 1 and 6 MMA issuers (one per 8-channel block of the chunk: single-thread issue
// was the limiter, ncu: 130 cycles per MMA against a 56-cycle tensor floor), warps 2-5 epilogue (TMEM lane quarters).
#include <algorithm>

#include "common.cuh"
#include "dose_b200.h"

namespace dp {

constexpr int kWgThreads = 224;       // producer, MMA issuer A, 4 epilogue warps, MMA issuer B
constexpr int kWgMmaWarpB = 6;
constexpr int kWgXStages = 6;

// tcgen05.mma with the two shared-memory descriptors given as 32-bit halves (only the start address in lo varies)
__device__ __forceinline__ void wg_umma(uint32_t tmem_d, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi,
                                        uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
      "mov.b64 da, {%1, %2};\n\t"
      "mov.b64 db, {%3, %4};\n\t"
      "setp.ne.b32 p, %6, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n\t}"
      ::"r"(tmem_d), "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate)
      : "memory");
}

struct WgradTcParams {
  int N, D, H, W, Ci, Co, k;
  uint8_t chunk_cb[64]; int16_t chunk_ci0[64]; uint8_t chunk_nci[64]; int n_chunks;
  int g_cb_off;
  int w_tile, n_wt, rb, n_hb;     // voxels per W tile, tiles per row, x rows per block, blocks per plane
  int xr;                         // x rows per TMA stage (several for narrow volumes: the per-row barrier round trip
                                  // is what limits the deep decoder levels, not the MMAs)
  int kdp, n_kdg;                 // depth taps per CTA, number of depth-tap groups
  int splits; long long wsize; float* ws; int* err_flag;
  uint32_t g_plane_bytes, g_buf_bytes, x_stage_bytes, x_box_bytes;
};

__global__ void __launch_bounds__(kWgThreads, 1)
conv_wgrad_tc_kernel(const __grid_constant__ CUtensorMap x_map, const __grid_constant__ CUtensorMap g_map,
                     const __grid_constant__ WgradTcParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  __shared__ uint64_t g_full[2], g_empty[2], x_full[kWgXStages], x_empty[kWgXStages], done_bar;
  __shared__ uint32_t tmem_base_smem;
  __shared__ uint32_t started_smem[2];

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int k = p.k, P = k / 2;
  const int ncols = k * 16;
  const int n_cot = p.Co / 16;
  int type = blockIdx.x;
  const int kdg = type % p.n_kdg; type /= p.n_kdg;
  const int cot = type % n_cot;
  const int chunk = type / n_cot;
  const int kd0 = kdg * p.kdp;
  const int n_kd = min(p.kdp, k - kd0);
  const int split = blockIdx.y;
  const long long n_blocks = static_cast<long long>(p.N) * p.D * p.n_hb * p.n_wt;
  const long long per = (n_blocks + p.splits - 1) / p.splits;
  const long long b_begin = split * per, b_end = min(n_blocks, b_begin + per);
  const int g_rows = p.rb + k - 1;

  uint8_t* g_buf = smem;
  uint8_t* x_buf = smem + 2 * static_cast<size_t>(p.g_buf_bytes);

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&x_map);
    tma_prefetch_desc(&g_map);
    for (int i = 0; i < 2; ++i) { mbar_init(&g_full[i], 1); mbar_init(&g_empty[i], 2); }
    for (int i = 0; i < kWgXStages; ++i) { mbar_init(&x_full[i], 1); mbar_init(&x_empty[i], 2); }
    mbar_init(&done_bar, 2);
    started_smem[0] = started_smem[1] = 0;
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<512>(&tmem_base_smem);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_smem;

  auto decode = [&](long long b, int& n, int& d, int& h0, int& w0) {
    w0 = static_cast<int>(b % p.n_wt) * p.w_tile; b /= p.n_wt;
    h0 = static_cast<int>(b % p.n_hb) * p.rb; b /= p.n_hb;
    d = static_cast<int>(b % p.D);
    n = static_cast<int>(b / p.D);
  };

  if (warp == 0) {
    // ===================================================================== TMA producer
    if (lane == 0) {
      int xs = 0; uint32_t xph = 0;
      int it = 0;
      for (long long b = b_begin; b < b_end; ++b, ++it) {
        int n, d, h0, w0;
        decode(b, n, d, h0, w0);
        const int buf = it & 1;
        const uint32_t gph = (it >> 1) & 1;
        if (!mbar_wait_relaxed(&g_empty[buf], gph ^ 1, p.err_flag)) break;
        mbar_arrive_expect_tx(&g_full[buf], static_cast<uint32_t>(n_kd) * p.g_plane_bytes);
        for (int kdi = 0; kdi < n_kd; ++kdi) {
          const int dg = d - (kd0 + kdi) + P;            // out-of-range planes / rows / columns are zero-filled by TMA
          tma_load_5d(g_buf + buf * static_cast<size_t>(p.g_buf_bytes) + kdi * static_cast<size_t>(p.g_plane_bytes), &g_map,
                      &g_full[buf], w0 * 2, p.g_cb_off + cot * 2, h0 - P, dg, n);
        }
        const int rows = min(p.rb, p.H - h0);
        for (int i = 0; i < rows; i += p.xr) {
          if (!mbar_wait_relaxed(&x_empty[xs], xph ^ 1, p.err_flag)) break;
          mbar_arrive_expect_tx(&x_full[xs], p.x_box_bytes);
          tma_load_5d(x_buf + xs * static_cast<size_t>(p.x_stage_bytes), &x_map, &x_full[xs], (w0 - P) * 2,
                      p.chunk_cb[chunk], h0 + i, d, n);
          if (++xs == kWgXStages) { xs = 0; xph ^= 1; }
        }
      }
    }
  } else if (warp == 1 || warp == kWgMmaWarpB) {
    // ===================================================================== MMA issuers (cb = 0 / cb = 1)
    if (elect_one()) {
      const int cb = (warp == 1) ? 0 : 1;
      // both operands MN-major (idesc bits 15, 16), fp16 in, fp32 accumulate
      const uint32_t idesc = make_idesc_f16(128, ncols) | (1u << 15) | (1u << 16);
      const uint32_t x_cb_bytes = static_cast<uint32_t>(p.w_tile + 16) * 16;
      const uint32_t g_cob_bytes = static_cast<uint32_t>(p.w_tile) * 16;
      const uint32_t g_row16 = (2 * g_cob_bytes) >> 4;
      // descriptor halves (see make_smem_desc): lo = start>>4 | LBO>>4 << 16, hi = SBO>>4 | version 1 << 14
      const uint32_t a_lo_c = (128u >> 4) << 16, a_hi = (16u >> 4) | (1u << 14);
      const uint32_t b_lo_c = (128u >> 4) << 16, b_hi = (g_cob_bytes >> 4) | (1u << 14);
      const int ksteps = p.w_tile / 16;
      uint32_t started = 0;
      int xs = 0; uint32_t xph = 0;
      int it = 0;
      bool ok = true;
      for (long long b = b_begin; b < b_end && ok; ++b, ++it) {
        int n, d, h0, w0;
        decode(b, n, d, h0, w0);
        const int buf = it & 1;
        const uint32_t gph = (it >> 1) & 1;
        if (!mbar_wait(&g_full[buf], gph, p.err_flag)) break;
        const uint32_t gb16 = smem_u32(g_buf + buf * static_cast<size_t>(p.g_buf_bytes)) >> 4;
        const int rows = min(p.rb, p.H - h0);
        for (int i0 = 0; i0 < rows; i0 += p.xr) {
          if (!mbar_wait(&x_full[xs], xph, p.err_flag)) { ok = false; break; }
          tc_fence_after();
          const uint32_t xs16 = (smem_u32(x_buf + xs * static_cast<size_t>(p.x_stage_bytes)) + cb * x_cb_bytes) >> 4;
          const int nr = min(p.xr, rows - i0);
          for (int r = 0; r < nr; ++r) {
            const uint32_t xa = a_lo_c | ((xs16 + static_cast<uint32_t>(r) * ((2 * x_cb_bytes) >> 4)) & 0x3FFFu);
            for (int kdi = 0; kdi < n_kd; ++kdi) {
              const int dg = d - (kd0 + kdi) + P;
              if (dg < 0 || dg >= p.D) continue;               // all-zero plane
              const int acc = kdi * 2 + cb;
              const uint32_t tcol = tmem_base + static_cast<uint32_t>(acc * ncols);
              const uint32_t gbl = b_lo_c | ((gb16 + static_cast<uint32_t>(kdi) * (p.g_plane_bytes >> 4) +
                                              static_cast<uint32_t>(i0 + r) * g_row16) & 0x3FFFu);
              uint32_t accum = (started >> acc) & 1u;
#pragma unroll 4
              for (int ks = 0; ks < ksteps; ++ks) {
                wg_umma(tcol, xa + ks * 16, a_hi, gbl + ks * 16, b_hi, idesc, accum);
                accum = 1;
              }
              started |= 1u << acc;
            }
          }
          umma_commit(&x_empty[xs]);
          if (++xs == kWgXStages) { xs = 0; xph ^= 1; }
        }
        umma_commit(&g_empty[buf]);
      }
      *reinterpret_cast<volatile uint32_t*>(&started_smem[cb]) = started;
      __threadfence_block();
      umma_commit(&done_bar);
    }
    __syncwarp();
  } else if (warp >= 2 && warp < 6) {
    // ===================================================================== epilogue: TMEM -> ws[split]
    mbar_wait_relaxed(&done_bar, 0, p.err_flag);
    tc_fence_after();
    __syncwarp();
    const uint32_t started = *reinterpret_cast<volatile uint32_t*>(&started_smem[0]) |
                             *reinterpret_cast<volatile uint32_t*>(&started_smem[1]);
    const int q = warp & 3;
    const int m = q * 32 + lane;             // accumulator row = s*8 + c
    const int s = m >> 3, c = m & 7;
    float* out = p.ws + static_cast<size_t>(split) * p.wsize;
    const int nci = p.chunk_nci[chunk], ci0 = p.chunk_ci0[chunk];
    if (q * 32 < k * 8) {                    // warp-uniform: quarters holding rows s < k
      for (int kdi = 0; kdi < n_kd; ++kdi)
        for (int cb = 0; cb < 2; ++cb) {
          const int acc = kdi * 2 + cb;
          const bool have = (started >> acc) & 1u;
          for (int j = 0; j < k; ++j) {
            uint32_t r[16];
            if (have) {
              tmem_ld16(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(acc * ncols + j * 16), r);
              tmem_ld_wait();
            } else {
#pragma unroll
              for (int t = 0; t < 16; ++t) r[t] = 0u;
            }
            const int ci = cb * 8 + c;
            if (s < k && ci < nci) {
              const int kh = k - 1 - j;
#pragma unroll
              for (int t = 0; t < 16; ++t) {
                const int co = cot * 16 + t;
                out[(((static_cast<size_t>(co) * p.Ci + ci0 + ci) * k + kd0 + kdi) * k + kh) * k + s] = __uint_as_float(r[t]);
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
    tmem_dealloc<512>(tmem_base);
  }
}

}  // namespace dp

using namespace dp;

extern "C" int dp_conv3d_wgrad_tc(const void* x_c8, int x_cb_total, const uint8_t* chunk_cb, const int* chunk_ci0,
                                  const int* chunk_nci, int n_chunks, const void* g_c8, int g_cb_total, int g_cb_off,
                                  int N, int D, int H, int W, int cin, int cout, int k, float* ws, int splits,
                                  int* err_flag, cudaStream_t stream) {
  DP_REQUIRE(k == 3 || k == 7, "conv3d_wgrad_tc: k must be 3 or 7 (got %d)", k);
  DP_REQUIRE(cout % 16 == 0 && n_chunks >= 1 && n_chunks <= 64 && splits >= 1, "conv3d_wgrad_tc: C_out %% 16, 1..64 chunks");
  WgradTcParams p{};
  p.N = N; p.D = D; p.H = H; p.W = W; p.Ci = cin; p.Co = cout; p.k = k; p.n_chunks = n_chunks;
  for (int i = 0; i < n_chunks; ++i) {
    p.chunk_cb[i] = chunk_cb[i];
    p.chunk_ci0[i] = static_cast<int16_t>(chunk_ci0[i]);
    p.chunk_nci[i] = static_cast<uint8_t>(chunk_nci[i]);
  }
  p.g_cb_off = g_cb_off; p.ws = ws; p.splits = splits; p.err_flag = err_flag;
  p.wsize = static_cast<long long>(cout) * cin * k * k * k;
  p.n_wt = (W + 63) / 64;                                           // W tiles of <= 64 voxels, balanced: 96 -> 2 x 48
  p.w_tile = ((W + p.n_wt - 1) / p.n_wt + 15) / 16 * 16;
  p.kdp = k == 7 ? 2 : 3;
  p.n_kdg = (k + p.kdp - 1) / p.kdp;
  // x rows per block: as many as fit two g buffers in ~170 KB of shared memory (and at most H)
  const int row_bytes = 2 * p.w_tile * 16;
  int g_rows = std::min(H + k - 1, (170 * 1024 / 2) / (p.kdp * row_bytes));
  g_rows = std::min(g_rows, 32);
  DP_REQUIRE(g_rows >= k, "conv3d_wgrad_tc: shared memory too small for a %d-row block", k);
  p.rb = g_rows - (k - 1);
  p.n_hb = (H + p.rb - 1) / p.rb;
  p.g_plane_bytes = static_cast<uint32_t>(g_rows) * row_bytes;
  p.g_buf_bytes = (p.g_plane_bytes * p.kdp + 1023u) & ~1023u;
  p.xr = std::max(1, std::min(8, 64 / p.w_tile));
  p.x_box_bytes = static_cast<uint32_t>(p.xr * 2 * (p.w_tile + 16) * 16);     // what one x TMA box transfers
  p.x_stage_bytes = (p.x_box_bytes + 1023u) & ~1023u;

  CUtensorMap x_map, g_map;
  const uint64_t vox16 = static_cast<uint64_t>(D) * H * W * 16;
  {
    const uint64_t dims[5] = {static_cast<uint64_t>(W) * 2, static_cast<uint64_t>(x_cb_total), static_cast<uint64_t>(H),
                              static_cast<uint64_t>(D), static_cast<uint64_t>(N)};
    const uint64_t strides[4] = {vox16, static_cast<uint64_t>(W) * 16, static_cast<uint64_t>(H) * W * 16, vox16 * x_cb_total};
    const uint32_t box[5] = {static_cast<uint32_t>((p.w_tile + 16) * 2), 2, static_cast<uint32_t>(p.xr), 1, 1};
    if (int rc = encode_tiled(&x_map, CU_TENSOR_MAP_DATA_TYPE_UINT64, 5, x_c8, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_NONE))
      return rc;
  }
  {
    const uint64_t dims[5] = {static_cast<uint64_t>(W) * 2, static_cast<uint64_t>(g_cb_total), static_cast<uint64_t>(H),
                              static_cast<uint64_t>(D), static_cast<uint64_t>(N)};
    const uint64_t strides[4] = {vox16, static_cast<uint64_t>(W) * 16, static_cast<uint64_t>(H) * W * 16, vox16 * g_cb_total};
    const uint32_t box[5] = {static_cast<uint32_t>(p.w_tile * 2), 2, static_cast<uint32_t>(g_rows), 1, 1};
    if (int rc = encode_tiled(&g_map, CU_TENSOR_MAP_DATA_TYPE_UINT64, 5, g_c8, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_NONE))
      return rc;
  }
  const size_t smem = 2 * static_cast<size_t>(p.g_buf_bytes) + static_cast<size_t>(kWgXStages) * p.x_stage_bytes + 1024;
  constexpr size_t kMaxSmem = 200 * 1024;      // + static barriers stays under the 227 KB opt-in limit
  DP_REQUIRE(smem <= kMaxSmem, "conv3d_wgrad_tc: %zu bytes of shared memory", smem);
  if (first_use_on_device(KF_WGRAD)) {
    DP_CHECK(cudaFuncSetAttribute(conv_wgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(kMaxSmem)));
  }
  dim3 grid(static_cast<unsigned>(n_chunks * (cout / 16) * p.n_kdg), static_cast<unsigned>(splits));
  conv_wgrad_tc_kernel<<<grid, kWgThreads, smem, stream>>>(x_map, g_map, p);
  return check_cuda(cudaGetLastError(), "conv3d_wgrad_tc");
}
