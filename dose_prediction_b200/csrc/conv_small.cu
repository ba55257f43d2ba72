// 3x3x3 convolution of a ONE-channel volume (the first conv of the seg net's res-block encoder, monai UnetResBlock.conv1
// with in_channels = 1: oar_transeg.py:92-100) as an exact fp32 direct convolution.
//
// Why not the tensor cores: with C_in = 1 a 16-channel K chunk is 15/16 padding, and the 3-term operand split doubles it:
// the depth-stacked tcgen05 kernel spent 0.86 ms (17 TFLOP/s algorithmic) on this layer at batch 8 x 128^3.  The layer is
// 27 x 16 = 432 FMA per voxel: FMA-bound at ~0.2 ms on the CUDA cores, reads the planar fp32 input (4 B/voxel instead of
// the 32 B/voxel hi+lo c8 copy) and is exact in fp32.  The weights are kernel parameters (constant bank -> FFMA operands,
// no load instructions); each thread walks a column along D keeping its 3x3x3 window in registers (9 shared-memory loads
// per output voxel).  Epilogue: raw fp32 c8 + per-(n,c) statistics, plus {sum x, sum x^2} of the INPUT per image, which
// lets the block's residual branch norm3(conv3(x)) (a 1x1x1 conv of the same single channel followed by InstanceNorm) be
// evaluated in closed form inside dp_norm_act (see res_x there) instead of being materialised.
#include "common.cuh"
#include "dose_b200.h"

namespace dp {

constexpr int C1_TD = 8, C1_TH = 8, C1_TW = 32;      // output tile per 256-thread block (thread = one (h, w) column)

struct ConvC1Params {
  const float* x;                 // planar fp32 [N][D][H][W]
  float w[27][16];                // [tap = (kd*3+kh)*3+kw][cout]
  float bias[16];
  int D, H, W;
  float* out_raw; int out_cb_total;
  double* stats;                  // [N][16][2] of the output, or null
  double* xstats;                 // [N][2] {sum x, sum x^2} of the input, or null
};

__global__ void __launch_bounds__(256) conv3_c1_kernel(const __grid_constant__ ConvC1Params p) {
  __shared__ float tile[C1_TD + 2][C1_TH + 2][C1_TW + 2];
  __shared__ float red[8][34];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int tiles_w = (p.W + C1_TW - 1) / C1_TW, tiles_h = (p.H + C1_TH - 1) / C1_TH;
  int b = blockIdx.x;
  const int tw = b % tiles_w; b /= tiles_w;
  const int th = b % tiles_h; b /= tiles_h;
  const int td = b;
  const int n = blockIdx.y;
  const int w0 = tw * C1_TW, h0 = th * C1_TH, d0 = td * C1_TD;
  const size_t vol = static_cast<size_t>(p.D) * p.H * p.W;
  const float* xin = p.x + static_cast<size_t>(n) * vol;
  for (int i = threadIdx.x; i < (C1_TD + 2) * (C1_TH + 2) * (C1_TW + 2); i += 256) {
    const int lw = i % (C1_TW + 2), lh = (i / (C1_TW + 2)) % (C1_TH + 2), ld = i / ((C1_TW + 2) * (C1_TH + 2));
    const int d = d0 + ld - 1, h = h0 + lh - 1, w = w0 + lw - 1;
    float v = 0.f;
    if (d >= 0 && d < p.D && h >= 0 && h < p.H && w >= 0 && w < p.W) v = __ldg(&xin[(static_cast<size_t>(d) * p.H + h) * p.W + w]);
    tile[ld][lh][lw] = v;
  }
  __syncthreads();
  const int h = h0 + ty, w = w0 + tx;
  const bool col_valid = h < p.H && w < p.W;
  float win[3][3][3];                                   // [d-1..d+1][h-1..h+1][w-1..w+1]
#pragma unroll
  for (int a = 0; a < 2; ++a)
#pragma unroll
    for (int bb = 0; bb < 3; ++bb)
#pragma unroll
      for (int c = 0; c < 3; ++c) win[a + 1][bb][c] = tile[a][ty + bb][tx + c];
  float s1[16], s2[16];
#pragma unroll
  for (int j = 0; j < 16; ++j) { s1[j] = 0.f; s2[j] = 0.f; }
  float xs1 = 0.f, xs2 = 0.f;
#pragma unroll 1
  for (int ld = 0; ld < C1_TD; ++ld) {
    const int d = d0 + ld;
#pragma unroll
    for (int bb = 0; bb < 3; ++bb)
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        win[0][bb][c] = win[1][bb][c];
        win[1][bb][c] = win[2][bb][c];
        win[2][bb][c] = tile[ld + 2][ty + bb][tx + c];
      }
    if (d >= p.D) break;                                // block-uniform
    float acc[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) acc[j] = p.bias[j];
#pragma unroll
    for (int a = 0; a < 3; ++a)
#pragma unroll
      for (int bb = 0; bb < 3; ++bb)
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          const float xv = win[a][bb][c];
#pragma unroll
          for (int j = 0; j < 16; ++j) acc[j] = fmaf(xv, p.w[(a * 3 + bb) * 3 + c][j], acc[j]);
        }
    if (col_valid) {
      const size_t v = (static_cast<size_t>(d) * p.H + h) * p.W + w;
#pragma unroll
      for (int blk = 0; blk < 2; ++blk) {
        float y8[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) y8[j] = acc[blk * 8 + j];
        st_global_v8f(p.out_raw + ((static_cast<size_t>(n) * p.out_cb_total + blk) * vol + v) * 8, y8);
      }
#pragma unroll
      for (int j = 0; j < 16; ++j) { s1[j] += acc[j]; s2[j] = fmaf(acc[j], acc[j], s2[j]); }
      const float xc = win[1][1][1];
      xs1 += xc;
      xs2 = fmaf(xc, xc, xs2);
    }
  }
  // block reduction of 16 x {sum, sumsq} (+ the input's): warp shuffles, then 8 warps through shared memory
  float vals[34];
#pragma unroll
  for (int j = 0; j < 16; ++j) { vals[2 * j] = s1[j]; vals[2 * j + 1] = s2[j]; }
  vals[32] = xs1; vals[33] = xs2;
#pragma unroll
  for (int j = 0; j < 34; ++j) {
    const float r = warp_sum(vals[j]);
    if (tx == 0) red[ty][j] = r;
  }
  __syncthreads();
  if (threadIdx.x < 34) {
    float t = 0.f;
#pragma unroll
    for (int q = 0; q < 8; ++q) t += red[q][threadIdx.x];
    if (threadIdx.x < 32) {
      if (p.stats != nullptr) atomicAdd(&p.stats[static_cast<size_t>(n) * 32 + threadIdx.x], static_cast<double>(t));
    } else if (p.xstats != nullptr) {
      atomicAdd(&p.xstats[static_cast<size_t>(n) * 2 + (threadIdx.x - 32)], static_cast<double>(t));
    }
  }
}

}  // namespace dp

extern "C" int dp_conv3d_c1(const float* x_planar, const float* w_host, const float* bias_host, int N, int D, int H, int W,
                            float* out_raw, int out_cb_total, double* stats, double* xstats, cudaStream_t stream) {
  using namespace dp;
  DP_REQUIRE(x_planar != nullptr && w_host != nullptr && out_raw != nullptr, "dp_conv3d_c1: missing operand");
  ConvC1Params p{};
  p.x = x_planar;
  for (int co = 0; co < 16; ++co) {                     // w_host: [16][1][3][3][3] (the nn.Conv3d layout)
    for (int t = 0; t < 27; ++t) p.w[t][co] = w_host[co * 27 + t];
    p.bias[co] = bias_host ? bias_host[co] : 0.f;
  }
  p.D = D; p.H = H; p.W = W; p.out_raw = out_raw; p.out_cb_total = out_cb_total; p.stats = stats; p.xstats = xstats;
  const int tiles = ((W + C1_TW - 1) / C1_TW) * ((H + C1_TH - 1) / C1_TH) * ((D + C1_TD - 1) / C1_TD);
  dim3 grid(static_cast<unsigned>(tiles), static_cast<unsigned>(N));
  conv3_c1_kernel<<<grid, 256, 0, stream>>>(p);
  return check_cuda(cudaGetLastError(), "conv3d_c1");
}
