// 3-D convolution (stride 1, k in {1,3,5,7}, dilation d) as an implicit GEMM on tcgen05 tensor cores.
//
//   M = 128 output voxels  (a 16(H) x 8(W) patch of one D-plane of one image)
//   N = C_out              (16..256, multiple of 16)
//   K = taps x C_in        (consumed 16 channels per tcgen05.mma)
//
// Data layout in HBM ("c8"): activations are [N][C/8][D][H][W][8] fp16, i.e. channel-blocked by 8 so
// that 8 channels of a voxel are one 16-byte vector and voxels along W are 16 bytes apart.  That is
// exactly the canonical K-major / no-swizzle UMMA core-matrix layout (8 rows x 16 B), which lets one
// TMA box load of an input halo patch [2 c8-blocks][16+halo rows][8+halo voxels][8 ch] serve EVERY
// in-plane tap of the kernel: the A operand of tap (kh,kw) is the same shared-memory patch with the
// descriptor start address advanced by ((kh*dil)*PW + kw*dil)*16 bytes (SBO = patch row pitch,
// LBO = c8-block pitch).  Halo zero padding comes from TMA out-of-bounds fill; out-of-range D planes
// are skipped.  Weights are pre-packed per (kd, 16-channel chunk) as [kh][kw][2][C_out][8] fp16 and
// streamed with 1-D bulk copies.  Accumulators live in TMEM (double buffered), the epilogue applies
// the folded bias/BatchNorm affine (+ReLU), accumulates InstanceNorm partial sums from the fp32
// accumulators and writes c8 tensors (fp32 "raw" for a following InstanceNorm, or fp16 hi[/lo]).
//
// Warp roles (320 threads): warp 0 = TMA producer, warp 1 = MMA issuer (+TMEM alloc), warps 2..9 =
// epilogue (TMEM lane quarter = warp_id % 4, two warps per quarter).  Persistent: each CTA walks tiles blockIdx.x + i*gridDim.x.
//
// Reference call sites this replaces (cuDNN via nn.Conv3d): OARSegmentation/Models/Nets/
// blocks_MDUNet.py:68-71,102-105 (conv_block_3 / conv_block_7), DosePrediction/Models/Networks/c3d.py:16,30,
// monai dynunet_block.UnetResBlock.conv1/conv2.
#include "common.cuh"
#include "dose_b200.h"

namespace dp {

struct ConvTcParams {
  int N, D, H, W;
  int k, dil, pad;
  int n_chunks, cb_total_in;
  int cout;
  int kd_s, n_kdg;          // kd planes per stage (1, or k when the whole kernel depth fits one stage)
  int kh_s, n_khg;          // kh rows per stage
  int PW, PHs;
  int stages;
  uint32_t a_bytes, a_bytes_al, stage_bytes;
  int tiles_h, tiles_w, num_tiles;   // num_tiles counts tile GROUPS: T adjacent W tiles share one stage (weights read once)
  int T, groups_w;
  const __half* wpack;
  const float* scale;
  const float* shift;
  int relu;
  float* out_f32;
  __half* out_hi;
  __half* out_lo;
  int cb_total_out, cb_out_off;
  int fold;                  // 1: cout = 2 * C_out MMA columns per tile, output channel c = column c + column c + C_out
                             // (3-term operand split with [W_hi | W_lo] stacked in N: hi chunks are read once)
  int dc_co, dc_q0;          // > 0: ConvTranspose3d k2 s2 as a 1^3 conv with columns (parity q - dc_q0) * dc_co + co, scattered to
                             // the output voxel (2d + (q >> 2), 2h + ((q >> 1) & 1), 2w + (q & 1)) of a [2D, 2H, 2W] tensor
  double* stats;
  int* err_flag;
  uint32_t tmem_cols;
  uint8_t chunk_cb[192];
  uint32_t tap_mask[192];    // per K-chunk bit mask over the k^3 taps (bit (kd*k+kh)*k+kw); zero bits are skipped
  int masked;                // tap masks given (k = 3, one depth plane and all three kernel rows per stage)
  // depth-pair mode (fold == 2): a tile covers output planes (d, d + 1); the MMA columns are [plane d: C_out | plane d + 1:
  // C_out] and the k + 1 "virtual" depth taps v = 0..k carry the weight rows [W[kd = v] | W[kd = v - 1]] (zero where the tap
  // does not exist): input plane d + v - pad is read ONCE for both output planes, and an N = 64 MMA (48 cycles for 64
  // columns, scripts/micro/mma_mix.cu) becomes an N = 128 MMA (64 cycles for 128)
  int dpair, Dt, acc_bufs;   // Dt = d-tiles per image (D, or ceil(D / 2)); accumulator buffers in TMEM (2, or 1 when T * cout fills it)
};

constexpr int kConvThreads = 320;       // producer + MMA + 8 epilogue warps (2 per TMEM lane quarter)
constexpr int kConvEpiWarps = 8;
constexpr int kMaxStages = 8;

// D[tmem] (+)= A * B with descriptors given as {lo, hi} 32-bit halves (only the start address in lo varies).
__device__ __forceinline__ void umma_f16_ss_split(uint32_t tmem_d, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo,
                                                  uint32_t b_hi, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
      "mov.b64 da, {%1, %2};\n\t"
      "mov.b64 db, {%3, %4};\n\t"
      "setp.ne.b32 p, %6, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n\t}"
      ::"r"(tmem_d), "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate)
      : "memory");
}

// Column sums of a 32(lanes = rows) x 16(columns) register tile with 16 shuffles instead of 80:
// every step trades half of the remaining columns with the partner lane.  On return lane L holds the
// sum over all 32 rows of column col16(L) = bits 4..1 of L (lanes L and L^1 hold the same value).
__device__ __forceinline__ float colsum16(const float (&v)[16], int lane) {
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

// Tap-masked (space-to-depth) convs: the valid taps of one depth plane form a rectangle [h0, h1] x [w0, w1] of the 3 x 3
// in-plane taps (the mask is a product set per input parity class), typically 1-4 of 9.  Producer and issuer derive the
// rectangle from the mask; only its weight blocks are copied, packed row by row (a stage was 20 KB of patch + 18 KB of
// weights for 1-4 MMAs per tile: the layers ran at the L2 -> SM bandwidth).
struct TapRect { int h0, h1, w0, w1; };
__device__ __forceinline__ TapRect tap_rect3(uint32_t m9) {
  TapRect r;
  const uint32_t rows = ((m9 & 7u) ? 1u : 0u) | (((m9 >> 3) & 7u) ? 2u : 0u) | (((m9 >> 6) & 7u) ? 4u : 0u);
  const uint32_t cols = (m9 | (m9 >> 3) | (m9 >> 6)) & 7u;
  r.h0 = __ffs(rows) - 1; r.h1 = 31 - __clz(rows);
  r.w0 = __ffs(cols) - 1; r.w1 = 31 - __clz(cols);
  return r;
}

template <int KS, int TG>
__global__ void __launch_bounds__(kConvThreads, 1)
conv3d_tc_kernel(const __grid_constant__ CUtensorMap tmap_in, const __grid_constant__ ConvTcParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  __shared__ uint64_t full_bar[kMaxStages];
  __shared__ uint64_t empty_bar[kMaxStages];
  __shared__ uint64_t tmem_full_bar[2];
  __shared__ uint64_t tmem_empty_bar[2];
  __shared__ uint32_t tmem_base_smem;
  __shared__ float stat_acc[kConvEpiWarps][256][2];
  __shared__ float s_scale[256], s_shift[256];

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_in);
    for (int i = 0; i < p.stages; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tmem_full_bar[i], 1);
      mbar_init(&tmem_empty_bar[i], kConvEpiWarps);
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                     smem_u32(&tmem_base_smem)),
                 "r"(p.tmem_cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  for (int i = threadIdx.x; i < kConvEpiWarps * 256 * 2; i += kConvThreads) (&stat_acc[0][0][0])[i] = 0.f;
  for (int i = threadIdx.x; i < ((p.fold || p.dpair) ? p.cout >> 1 : p.cout); i += kConvThreads) { s_scale[i] = p.scale[i]; s_shift[i] = p.shift[i]; }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_smem;

  const int tiles_per_plane = p.tiles_h * p.groups_w;          // tile groups per D plane
  const uint32_t tap_b_bytes = 32u * static_cast<uint32_t>(p.cout);  // one tap: [2][cout][8] fp16
  const size_t kd_w_halfs = static_cast<size_t>(p.n_chunks) * KS * KS * 16 * p.cout;   // weights of one kd (all chunks)
  const size_t chunk_w_halfs = static_cast<size_t>(KS) * KS * 16 * p.cout;

  if (warp == 0) {
    // ===================================================================== TMA producer (warp-uniform loop)
    int stage = 0;
    uint32_t phase = 0;
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
      int t = tile;
      const int tw = (t % p.groups_w) * p.T; t /= p.groups_w;
      const int th = t % p.tiles_h; t /= p.tiles_h;
      const int d = (t % p.Dt) << p.dpair;
      const int n = t / p.Dt;
      const int h0 = th * 16, w0 = tw * 8;
      for (int kdg = 0; kdg < p.n_kdg; ++kdg) {
        const int kd0 = kdg * p.kd_s;
        const int dz = d + kd0 * p.dil - p.pad;
        if (p.kd_s == 1 && (dz < 0 || dz >= p.D)) continue;
        for (int c = 0; c < p.n_chunks; ++c) {
          if (KS <= 3 && p.kd_s == 1 && ((p.tap_mask[c] >> (kd0 * KS * KS)) & ((1u << (KS * KS)) - 1u)) == 0u) continue;   // no tap of this depth
          const __half* wsrc = p.wpack + static_cast<size_t>(kd0) * kd_w_halfs + static_cast<size_t>(c) * chunk_w_halfs;
          for (int g = 0; g < p.n_khg; ++g) {
            const int kh0 = g * p.kh_s;
            const int cnt = min(p.kh_s, KS - kh0);
            if (!mbar_wait_relaxed(&empty_bar[stage], phase ^ 1, p.err_flag)) goto teardown;
            if (elect_one()) {
              uint8_t* sa = smem + static_cast<size_t>(stage) * p.stage_bytes;
              uint8_t* sb = sa + p.a_bytes_al;
              if (KS == 3 && p.masked) {
                // only the valid taps' weight blocks, packed row by row
                const TapRect tr = tap_rect3((p.tap_mask[c] >> (kd0 * 9)) & 0x1FFu);
                const uint32_t nw = static_cast<uint32_t>(tr.w1 - tr.w0 + 1), nh = static_cast<uint32_t>(tr.h1 - tr.h0 + 1);
                mbar_arrive_expect_tx(&full_bar[stage], p.a_bytes + nh * nw * tap_b_bytes);
                tma_load_5d(sa, &tmap_in, &full_bar[stage], 0, w0 - p.pad, h0 - p.pad, dz, n * p.cb_total_in + p.chunk_cb[c]);
                for (uint32_t r = 0; r < nh; ++r)
                  bulk_load_1d(sb + r * nw * tap_b_bytes, wsrc + static_cast<size_t>((tr.h0 + r) * 3 + tr.w0) * 16 * p.cout,
                               nw * tap_b_bytes, &full_bar[stage]);
              } else {
              const uint32_t b_bytes = static_cast<uint32_t>(cnt * KS) * tap_b_bytes;
              mbar_arrive_expect_tx(&full_bar[stage], p.a_bytes + b_bytes * p.kd_s);
              tma_load_5d(sa, &tmap_in, &full_bar[stage], 0, w0 - p.pad, h0 - p.pad + kh0 * p.dil, dz,
                          n * p.cb_total_in + p.chunk_cb[c]);
              for (int kdl = 0; kdl < p.kd_s; ++kdl)
                bulk_load_1d(sb + kdl * b_bytes, wsrc + kdl * kd_w_halfs + static_cast<size_t>(kh0) * KS * 16 * p.cout,
                             b_bytes, &full_bar[stage]);
              }
            }
            __syncwarp();
            if (++stage == p.stages) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================================================================== MMA issuer (warp-uniform loop, one elected lane issues)
    const uint32_t idesc_full = make_idesc_f16(128, p.cout);
    // folded 3-term split: chunks [n_chunks/2, n_chunks) are the x_lo blocks, which only need the W_hi half of the columns
    // (executed MMA columns are what a power-capped B200 charges for: scripts/micro/mma_mix.cu)
    const uint32_t idesc_lo = p.fold ? make_idesc_f16(128, p.cout >> 1) : idesc_full;
    const int first_lo_chunk = p.fold ? (p.n_chunks >> 1) : p.n_chunks;
    // descriptor halves: lo = start>>4 | LBO>>4 << 16 ; hi = SBO>>4 | version(1) << 14 | layout(0) << 29
    const uint32_t a_lo_c = (static_cast<uint32_t>(p.kd_s * p.PHs * p.PW) & 0x3FFFu) << 16;   // LBO = c8-block pitch
    const uint32_t a_hi = (static_cast<uint32_t>(p.PW) & 0x3FFFu) | (1u << 14);               // SBO = patch row pitch
    const uint32_t b_lo_c = (static_cast<uint32_t>(p.cout) & 0x3FFFu) << 16;                  // LBO = cout*16 B
    const uint32_t b_hi = 8u | (1u << 14);                                                    // SBO = 128 B
    const uint32_t tap_b16 = tap_b_bytes >> 4;
    int stage = 0;
    uint32_t phase = 0;
    int iter = 0;
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++iter) {
      const int d = ((tile / tiles_per_plane) % p.Dt) << p.dpair;
      const int ntile = min(p.T, p.tiles_w - (tile % p.groups_w) * p.T);
      const int slot = p.acc_bufs == 2 ? (iter & 1) : 0;
      const int use = p.acc_bufs == 2 ? (iter >> 1) : iter;
      if (!mbar_wait(&tmem_empty_bar[slot], (use & 1) ^ 1, p.err_flag)) goto teardown;
      tc_fence_after();
      const uint32_t tmem_d = tmem_base + static_cast<uint32_t>(slot * p.T * p.cout);
      uint32_t accumulate = 0;
      for (int kdg = 0; kdg < p.n_kdg; ++kdg) {
        const int dz = d + kdg * p.kd_s * p.dil - p.pad;
        if (p.kd_s == 1 && (dz < 0 || dz >= p.D)) continue;
        for (int c = 0; c < p.n_chunks; ++c) {
          if (KS <= 3 && p.kd_s == 1 && ((p.tap_mask[c] >> (kdg * KS * KS)) & ((1u << (KS * KS)) - 1u)) == 0u) continue;
          for (int g = 0; g < p.n_khg; ++g) {
            const int cnt = min(p.kh_s, KS - g * p.kh_s);
            if (!mbar_wait(&full_bar[stage], phase, p.err_flag)) goto teardown;
            tc_fence_after();
            if (elect_one()) {
              const uint32_t sa16 = smem_u32(smem + static_cast<size_t>(stage) * p.stage_bytes) >> 4;
              uint32_t b16 = sa16 + (p.a_bytes_al >> 4);
              const uint32_t mask = p.tap_mask[c];
              const uint32_t idesc = c < first_lo_chunk ? idesc_full : idesc_lo;
              if (KS == 3 && p.masked) {
                const TapRect tr = tap_rect3((mask >> (kdg * 9)) & 0x1FFu);
                for (int kh = tr.h0; kh <= tr.h1; ++kh)
                  for (int kw = tr.w0; kw <= tr.w1; ++kw) {
                    const uint32_t a16 = sa16 + static_cast<uint32_t>(kh * p.dil * p.PW + kw * p.dil);
#pragma unroll
                    for (int tt = 0; tt < TG; ++tt)
                      if (TG == 1 || tt < ntile)
                        umma_f16_ss_split(tmem_d + static_cast<uint32_t>(tt * p.cout), a_lo_c | ((a16 + tt * 8) & 0x3FFFu), a_hi,
                                          b_lo_c | (b16 & 0x3FFFu), b_hi, idesc, accumulate);
                    accumulate = 1;
                    b16 += tap_b16;
                  }
              } else
              for (int kdl = 0; kdl < p.kd_s; ++kdl) {
                for (int khl = 0; khl < cnt; ++khl) {
                  uint32_t a16 = sa16 + static_cast<uint32_t>((kdl * p.PHs + khl * p.dil) * p.PW);
                  const int tap0 = ((kdg * p.kd_s + kdl) * KS + g * p.kh_s + khl) * KS;
#pragma unroll
                  for (int kw = 0; kw < KS; ++kw) {
                    if (KS > 3 || ((mask >> (tap0 + kw)) & 1u)) {
#pragma unroll
                      for (int tt = 0; tt < TG; ++tt)
                        if (TG == 1 || tt < ntile)
                          umma_f16_ss_split(tmem_d + static_cast<uint32_t>(tt * p.cout), a_lo_c | ((a16 + tt * 8) & 0x3FFFu), a_hi,
                                            b_lo_c | (b16 & 0x3FFFu), b_hi, idesc, accumulate);
                      accumulate = 1;
                    }
                    a16 += static_cast<uint32_t>(p.dil);
                    b16 += tap_b16;
                  }
                }
              }
              umma_commit(&empty_bar[stage]);
            }
            __syncwarp();
            accumulate = __any_sync(0xffffffffu, accumulate != 0) ? 1u : 0u;   // only the elected lane issued
            if (++stage == p.stages) { stage = 0; phase ^= 1; }
          }
        }
      }
      if (elect_one()) umma_commit(&tmem_full_bar[slot]);
      __syncwarp();
    }
  } else {
    // ===================================================================== epilogue (warps 2..9)
    // two warps per TMEM lane quarter; the pair interleaves the 16-column chunks of the accumulator
    const int quarter = warp & 3;
    const int ew = warp - 2;
    const int cgrp = ew >> 2;
    const int row = quarter * 32 + lane;
    const int hl = row >> 3, wl = row & 7;
    const int mycol = (lane >> 1) & 15;
    const int cout_out = (p.fold || p.dpair) ? p.cout >> 1 : p.cout;
    const float relu_floor = p.relu ? 0.f : -INFINITY;      // one FMNMX per value instead of a predicated pair
    int cur_n = -1;
    int iter = 0;
    auto flush_stats = [&](int n) {
      if (p.stats == nullptr || n < 0) return;
      __syncwarp();
      for (int c = lane; c < cout_out; c += 32) {
        atomicAdd(&p.stats[(static_cast<size_t>(n) * cout_out + c) * 2 + 0], static_cast<double>(stat_acc[ew][c][0]));
        atomicAdd(&p.stats[(static_cast<size_t>(n) * cout_out + c) * 2 + 1], static_cast<double>(stat_acc[ew][c][1]));
        stat_acc[ew][c][0] = 0.f;
        stat_acc[ew][c][1] = 0.f;
      }
      __syncwarp();
    };
    const size_t plane = static_cast<size_t>(p.D) * p.H * p.W;
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++iter) {
      int t = tile;
      const int tw0 = (t % p.groups_w) * p.T; t /= p.groups_w;
      const int th = t % p.tiles_h; t /= p.tiles_h;
      const int d = (t % p.Dt) << p.dpair;
      const int n = t / p.Dt;
      if (n != cur_n) { flush_stats(cur_n); cur_n = n; }
      const int ntile = min(p.T, p.tiles_w - tw0);
      const int slot = p.acc_bufs == 2 ? (iter & 1) : 0;
      const int use = p.acc_bufs == 2 ? (iter >> 1) : iter;
      if (!mbar_wait_relaxed(&tmem_full_bar[slot], use & 1, p.err_flag)) goto teardown;
      tc_fence_after();
      for (int tt = 0; tt < ntile; ++tt) {
      const int h = th * 16 + hl, w = (tw0 + tt) * 8 + wl;
      const bool valid = (h < p.H) && (w < p.W);
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + static_cast<uint32_t>((slot * p.T + tt) * p.cout);
      if (p.dc_co > 0) {
        // transposed conv: the W-neighbours l = 0 / 1 of an output voxel pair are the column chunks c and c + dc_co;
        // one thread converts both and writes whole 32-byte sectors (st.global.v8) per channel block
        const int cpq = p.dc_co >> 4;
        const int npairs = p.cout >> 5;
        for (int pi = cgrp; pi < npairs; pi += 2) {
          const int qe = (pi / cpq) * 2, cc = pi % cpq;
          const int c_even = (qe * cpq + cc) * 16;
          uint32_t r[2][16];
          tmem_ld16(taddr + c_even, r[0]);
          tmem_ld16(taddr + c_even + p.dc_co, r[1]);
          tmem_ld_wait();
          if (!valid) continue;
          const int q = p.dc_q0 + qe;
          const size_t vo = (static_cast<size_t>(2 * d + (q >> 2)) * (2 * p.H) + 2 * h + ((q >> 1) & 1)) * (2 * p.W) + 2 * w;
#pragma unroll
          for (int b = 0; b < 2; ++b) {
            const size_t cb = static_cast<size_t>(n) * p.cb_total_out + p.cb_out_off + cc * 2 + b;
            const size_t off = (cb * (8 * plane) + vo) * 8;
            uint32_t hi[8], lo[8];
#pragma unroll
            for (int u = 0; u < 2; ++u)
#pragma unroll
              for (int j = 0; j < 8; j += 2) {
                const float x0 = __uint_as_float(r[u][b * 8 + j]), x1 = __uint_as_float(r[u][b * 8 + j + 1]);
                const __half2 h2 = __floats2half2_rn(x0, x1);
                const float2 back = __half22float2(h2);
                const __half2 l2 = __floats2half2_rn(x0 - back.x, x1 - back.y);
                hi[u * 4 + (j >> 1)] = *reinterpret_cast<const uint32_t*>(&h2);
                lo[u * 4 + (j >> 1)] = *reinterpret_cast<const uint32_t*>(&l2);
              }
            st_global_v8u(p.out_hi + off, hi);
            if (p.out_lo != nullptr) st_global_v8u(p.out_lo + off, lo);
          }
        }
      } else
      for (int half = 0; half <= p.dpair; ++half) {       // depth-pair mode: columns [0, C_out) = plane d, [C_out, 2 C_out) = d + 1
      if (d + half >= p.D) break;
      const size_t vox = (static_cast<size_t>(d + half) * p.H + h) * p.W + w;
      const uint32_t thalf = taddr + static_cast<uint32_t>(half * cout_out);
      for (int c0 = cgrp * 16; c0 < cout_out; c0 += 32) {
        uint32_t r[16];
        tmem_ld16(thalf + c0, r);
        if (p.fold) {
          uint32_t r2[16];
          tmem_ld16(taddr + c0 + cout_out, r2);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 16; ++j) r[j] = __float_as_uint(__uint_as_float(r[j]) + __uint_as_float(r2[j]));
        } else {
          tmem_ld_wait();
        }
        float v[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          float x = fmaf(__uint_as_float(r[j]), s_scale[c0 + j], s_shift[c0 + j]);
          x = fmaxf(x, relu_floor);
          v[j] = x;
        }
        if (!valid) {                                     // edge tiles only: rows outside the volume must not count
#pragma unroll
          for (int j = 0; j < 16; ++j) v[j] = 0.f;
        }
        if (p.stats != nullptr) {
          float sq[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) sq[j] = v[j] * v[j];
          const float s1 = colsum16(v, lane);
          const float s2 = colsum16(sq, lane);
          if ((lane & 1) == 0) {
            stat_acc[ew][c0 + mycol][0] += s1;
            stat_acc[ew][c0 + mycol][1] += s2;
          }
        }
        if (valid) {
#pragma unroll
          for (int b = 0; b < 2; ++b) {
            const size_t cb = static_cast<size_t>(n) * p.cb_total_out + p.cb_out_off + (c0 >> 3) + b;
            const size_t off = (cb * plane + vox) * 8;
            if (p.out_f32 != nullptr) {
              float y8[8];
#pragma unroll
              for (int j = 0; j < 8; ++j) y8[j] = v[b * 8 + j];
              st_global_v8f(p.out_f32 + off, y8);
            }
            if (p.out_hi != nullptr) {
              __align__(16) __half hi[8];
              __align__(16) __half lo[8];
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                hi[j] = __float2half_rn(v[b * 8 + j]);
                lo[j] = __float2half_rn(v[b * 8 + j] - __half2float(hi[j]));
              }
              *reinterpret_cast<uint4*>(p.out_hi + off) = *reinterpret_cast<const uint4*>(hi);
              if (p.out_lo != nullptr) *reinterpret_cast<uint4*>(p.out_lo + off) = *reinterpret_cast<const uint4*>(lo);
            }
          }
        }
      }
      }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty_bar[slot]);
    }
    flush_stats(cur_n);
  }

teardown:
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(p.tmem_cols) : "memory");
  }
}

}  // namespace dp

// =============================================================================== C ABI
namespace dp {
static int launch_conv_tc(const void* in_c8, int cb_total_in, const uint8_t* chunk_cb, int n_chunks,
                          const void* wpack, int N, int D, int H, int W, int cout, int k, int dil,
                          const float* scale, const float* shift, int relu, float* out_f32, void* out_hi,
                          void* out_lo, int cb_total_out, int cb_out_off, double* stats, int* err_flag,
                          int max_ctas, const uint32_t* tap_mask, int dc_co, int dc_q0, int fold, cudaStream_t stream) {
  DP_REQUIRE(fold == 0 || (cout <= 128 && dc_co == 0), "dp_conv3d_tc: fold needs C_out <= 128");
  const bool dpair = (fold == 2);      // depth-pair mode: two output planes per tile, k + 1 virtual depth taps
  DP_REQUIRE(!dpair || (dil == 1 && tap_mask == nullptr && cout <= 64 && D >= 2), "dp_conv3d_tc: depth-pair mode needs dilation 1, "
             "no tap masks, C_out <= 64 and D >= 2");
  if (dpair) fold = 0;
  if (fold || dpair) cout *= 2;        // MMA columns per tile; everything below sizes by the MMA N
  DP_REQUIRE(cout % 16 == 0 && cout >= 16 && cout <= 256, "dp_conv3d_tc: C_out=%d must be a multiple of 16 in [16,256]", cout);
  DP_REQUIRE(k >= 1 && k <= 7 && (k & 1), "dp_conv3d_tc: kernel size %d unsupported", k);
  DP_REQUIRE(n_chunks >= 1 && n_chunks <= 192, "dp_conv3d_tc: n_chunks=%d out of range", n_chunks);
  DP_REQUIRE(out_f32 != nullptr || out_hi != nullptr, "dp_conv3d_tc: no output tensor given");
  ConvTcParams p{};
  p.N = N; p.D = D; p.H = H; p.W = W;
  p.k = k; p.dil = dil; p.pad = dil * (k - 1) / 2;
  p.n_chunks = n_chunks; p.cb_total_in = cb_total_in; p.cout = cout;
  for (int i = 0; i < n_chunks; ++i) { p.chunk_cb[i] = chunk_cb[i]; p.tap_mask[i] = tap_mask ? tap_mask[i] : 0xFFFFFFFFu; }
  DP_REQUIRE(tap_mask == nullptr || k == 3, "dp_conv3d_tc: tap masks are supported for k = 3 only");
  p.masked = tap_mask != nullptr ? 1 : 0;
  // W-tile groups: for k <= 3 the layer is bound by the L2 -> SM weight stream (every 128-voxel tile re-reads all
  // weights), so T adjacent tiles share one stage; bounded by the double-buffered accumulators (2*T*cout <= 512 columns)
  p.tiles_w = (W + 7) / 8;
  int T = 1;
  if (k <= 3) T = cout <= 64 ? 4 : (cout <= 128 ? 2 : 1);
  else if (k == 7 && dpair) T = 4;     // one accumulator buffer of 4 x 128 columns: the 4 KB weight blocks are shared by 4 tiles
  else if (k == 7 && cout <= 128) T = 2;
  while (T > 1 && T > p.tiles_w) T >>= 1;
  p.T = T;
  p.groups_w = (p.tiles_w + T - 1) / T;
  p.PW = 8 * T + (k - 1) * dil;
  // stage sizing: as many kh rows per stage as fit ~56 KB, then as many stages as fit ~200 KB
  const int tap_b = 32 * cout;
  int kh_s = k, kd_s = (dil == 1 && tap_mask == nullptr && !dpair) ? k : 1;   // tap-masked (space-to-depth) convs: per-depth stages, empty ones skipped
  auto stage_bytes_for = [&](int sd, int s) {
    const int phs = 16 + (s - 1) * dil;
    const int a = 2 * sd * phs * p.PW * 16;
    return ((a + 127) / 128) * 128 + sd * s * k * tap_b;
  };
  if (stage_bytes_for(kd_s, kh_s) > 64 * 1024) kd_s = 1;     // whole-depth stages only when they stay small
  int masked_taps = 0;                                        // tap-masked convs: weight area = the largest valid tap rectangle
  if (tap_mask != nullptr) {
    for (int c = 0; c < n_chunks; ++c)
      for (int kd = 0; kd < 3; ++kd) {
        const uint32_t m9 = (tap_mask[c] >> (kd * 9)) & 0x1FFu;
        if (!m9) continue;
        const uint32_t rows = ((m9 & 7u) ? 1u : 0u) | (((m9 >> 3) & 7u) ? 2u : 0u) | (((m9 >> 6) & 7u) ? 4u : 0u);
        const uint32_t cols = (m9 | (m9 >> 3) | (m9 >> 6)) & 7u;
        auto span = [](uint32_t b) { int lo = 0, hi = 2; while (!((b >> lo) & 1u)) ++lo; while (!((b >> hi) & 1u)) --hi; return hi - lo + 1; };
        const int area = span(rows) * span(cols);
        if (area > masked_taps) masked_taps = area;
      }
    DP_REQUIRE(masked_taps >= 1, "dp_conv3d_tc: every tap is masked");
  }
  while (!masked_taps && kh_s > 1 && stage_bytes_for(kd_s, kh_s) > 56 * 1024) --kh_s;
  p.kd_s = kd_s;
  p.n_kdg = dpair ? k + 1 : k / kd_s;
  p.dpair = dpair ? 1 : 0;
  p.Dt = dpair ? (D + 1) / 2 : D;
  p.acc_bufs = (2 * T * cout <= 512) ? 2 : 1;
  p.kh_s = kh_s;
  p.n_khg = (k + kh_s - 1) / kh_s;
  p.PHs = 16 + (kh_s - 1) * dil;
  p.a_bytes = 2u * kd_s * p.PHs * p.PW * 16u;
  p.a_bytes_al = ((p.a_bytes + 127u) / 128u) * 128u;
  p.stage_bytes = ((static_cast<uint32_t>(stage_bytes_for(kd_s, kh_s)) + 1023u) / 1024u) * 1024u;
  if (masked_taps)              // all three kernel rows of the patch, but only the valid rectangle's weight blocks
    p.stage_bytes = ((p.a_bytes_al + static_cast<uint32_t>(masked_taps * tap_b) + 1023u) / 1024u) * 1024u;
  int stages = static_cast<int>((200u * 1024u) / p.stage_bytes);
  if (stages > kMaxStages) stages = kMaxStages;
  DP_REQUIRE(stages >= 2, "dp_conv3d_tc: stage of %u bytes does not fit twice in shared memory", p.stage_bytes);
  p.stages = stages;
  p.tiles_h = (H + 15) / 16;
  p.num_tiles = N * p.Dt * p.tiles_h * p.groups_w;
  p.wpack = static_cast<const __half*>(wpack);
  p.scale = scale; p.shift = shift; p.relu = relu;
  p.out_f32 = out_f32; p.out_hi = static_cast<__half*>(out_hi); p.out_lo = static_cast<__half*>(out_lo);
  p.cb_total_out = cb_total_out; p.cb_out_off = cb_out_off;
  p.stats = stats; p.err_flag = err_flag;
  p.dc_co = dc_co; p.dc_q0 = dc_q0; p.fold = fold;
  uint32_t cols = 32;
  while (cols < static_cast<uint32_t>(p.acc_bufs * T * cout)) cols <<= 1;
  DP_REQUIRE(cols <= 512, "dp_conv3d_tc: %d tiles x %d columns do not fit TMEM", T, cout);
  p.tmem_cols = cols;

  CUtensorMap tmap;
  const uint64_t dims[5] = {8, static_cast<uint64_t>(W), static_cast<uint64_t>(H), static_cast<uint64_t>(D),
                            static_cast<uint64_t>(N) * cb_total_in};
  const uint64_t strides[4] = {16, static_cast<uint64_t>(W) * 16, static_cast<uint64_t>(H) * W * 16,
                               static_cast<uint64_t>(D) * H * W * 16};
  const uint32_t box[5] = {8, static_cast<uint32_t>(p.PW), static_cast<uint32_t>(p.PHs), static_cast<uint32_t>(p.kd_s), 2};
  if (int rc = encode_tiled(&tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 5, in_c8, dims, strides, box,
                            CU_TENSOR_MAP_SWIZZLE_NONE))
    return rc;

  const size_t smem = static_cast<size_t>(p.stages) * p.stage_bytes + 1024;
  int grid = sm_count();
  if (max_ctas > 0 && grid > max_ctas) grid = max_ctas;
  if (grid > p.num_tiles) grid = p.num_tiles;
  if (first_use_on_device(KF_CONV_TC)) {
    const int max_smem = 8 * 25 * 1024 + 1024 + 4096;
#define DP_TC_ATTR(K_, T_) DP_CHECK(cudaFuncSetAttribute(conv3d_tc_kernel<K_, T_>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem))
    DP_TC_ATTR(1, 1); DP_TC_ATTR(1, 2); DP_TC_ATTR(1, 4);
    DP_TC_ATTR(3, 1); DP_TC_ATTR(3, 2); DP_TC_ATTR(3, 4);
    DP_TC_ATTR(5, 1); DP_TC_ATTR(7, 1); DP_TC_ATTR(7, 2); DP_TC_ATTR(7, 4);
#undef DP_TC_ATTR
  }
#define DP_TC_LAUNCH(K_, T_) conv3d_tc_kernel<K_, T_><<<grid, kConvThreads, smem, stream>>>(tmap, p)
  switch (k) {
    case 1: if (T == 4) DP_TC_LAUNCH(1, 4); else if (T == 2) DP_TC_LAUNCH(1, 2); else DP_TC_LAUNCH(1, 1); break;
    case 3: if (T == 4) DP_TC_LAUNCH(3, 4); else if (T == 2) DP_TC_LAUNCH(3, 2); else DP_TC_LAUNCH(3, 1); break;
    case 5: DP_TC_LAUNCH(5, 1); break;
    default: if (T == 4) DP_TC_LAUNCH(7, 4); else if (T == 2) DP_TC_LAUNCH(7, 2); else DP_TC_LAUNCH(7, 1); break;
  }
#undef DP_TC_LAUNCH
  DP_CHECK(cudaGetLastError());
  return 0;
}
}  // namespace dp

extern "C" int dp_conv3d_tc(const void* in_c8, int cb_total_in, const uint8_t* chunk_cb, int n_chunks,
                            const void* wpack, int N, int D, int H, int W, int cout, int k, int dil,
                            const float* scale, const float* shift, int relu, float* out_f32, void* out_hi,
                            void* out_lo, int cb_total_out, int cb_out_off, double* stats, int* err_flag,
                            int max_ctas, const uint32_t* tap_mask, int fold, cudaStream_t stream) {
  return dp::launch_conv_tc(in_c8, cb_total_in, chunk_cb, n_chunks, wpack, N, D, H, W, cout, k, dil, scale, shift, relu,
                            out_f32, out_hi, out_lo, cb_total_out, cb_out_off, stats, err_flag, max_ctas, tap_mask, 0, 0,
                            fold, stream);
}

extern "C" int dp_deconv2x_tc(const void* in_c8, int cb_total_in, const uint8_t* chunk_cb, int n_chunks,
                              const void* wpack, int N, int D, int H, int W, int cout, int q0, int nq,
                              const float* scale, const float* shift, void* out_hi, void* out_lo, int cb_total_out,
                              int cb_out_off, int* err_flag, cudaStream_t stream) {
  using namespace dp;
  DP_REQUIRE(cout % 16 == 0 && nq >= 2 && (nq & 1) == 0 && q0 >= 0 && (q0 & 1) == 0 && q0 + nq <= 8 && nq * cout <= 256,
             "dp_deconv2x_tc: C_out=%d, parities [%d, %d) do not fit one 256-column launch", cout, q0, q0 + nq);
  DP_REQUIRE(out_hi != nullptr, "dp_deconv2x_tc: no output tensor given");
  return launch_conv_tc(in_c8, cb_total_in, chunk_cb, n_chunks, wpack, N, D, H, W, nq * cout, 1, 1, scale, shift, 0,
                        nullptr, out_hi, out_lo, cb_total_out, cb_out_off, nullptr, err_flag, 0, nullptr, cout, q0, 0, stream);
}
