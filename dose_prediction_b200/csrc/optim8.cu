// Block-wise 8-bit optimizer-state (de)quantisation in the layout of bitsandbytes' 8-bit optimizers (the reference trains
// DOSE-PYFER with bnb.optim.Adam8bit, DosePrediction/Train/train_light_pyfer.py:194-197): every 2048-element block of a
// state tensor is stored as uint8 codes into a 256-entry "dynamic" code book (qmap) times the block's absmax.
//   quantise:   absmax[b] = max |x| over the block;  code = argmin_j |qmap[j] - x / absmax[b]|
//   dequantise: x = qmap[code] * absmax[b]
// bitsandbytes itself is an un-vendored dependency that is absent offline: the layout and the code book follow its
// published sources (functional.create_dynamic_map, Optimizer2State.init_state, blocksize 2048); parity vs the real package
// is unpinned.  Our trainers keep fp32 Adam moments while training; these kernels import / export them.
#include "common.cuh"
#include "dose_b200.h"

namespace dp {

constexpr int Q8_BLOCK = 2048;

__global__ void __launch_bounds__(256) quantize_blockwise_kernel(const float* __restrict__ x, long long n, const float* __restrict__ qmap,
                                                                 uint8_t* codes, float* absmax) {
  __shared__ float s_map[256];
  __shared__ float s_red[8];
  const long long b0 = static_cast<long long>(blockIdx.x) * Q8_BLOCK;
  s_map[threadIdx.x] = qmap[threadIdx.x];
  float v[Q8_BLOCK / 256];
  float m = 0.f;
#pragma unroll
  for (int i = 0; i < Q8_BLOCK / 256; ++i) {
    const long long idx = b0 + i * 256 + threadIdx.x;
    v[i] = idx < n ? x[idx] : 0.f;
    m = fmaxf(m, fabsf(v[i]));
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = m;
  __syncthreads();
  m = s_red[0];
#pragma unroll
  for (int w = 1; w < 8; ++w) m = fmaxf(m, s_red[w]);
  if (threadIdx.x == 0) absmax[blockIdx.x] = m;
  const float inv = m > 0.f ? 1.f / m : 0.f;
#pragma unroll
  for (int i = 0; i < Q8_BLOCK / 256; ++i) {
    const long long idx = b0 + i * 256 + threadIdx.x;
    if (idx >= n) continue;
    const float t = v[i] * inv;
    int lo = 0, hi = 255;                       // the code book is sorted ascending: binary search, then the nearer neighbour
    while (hi - lo > 1) {
      const int mid = (lo + hi) >> 1;
      if (s_map[mid] <= t) lo = mid; else hi = mid;
    }
    codes[idx] = static_cast<uint8_t>((t - s_map[lo] <= s_map[hi] - t) ? lo : hi);
  }
}

__global__ void __launch_bounds__(256) dequantize_blockwise_kernel(const uint8_t* __restrict__ codes, const float* __restrict__ absmax,
                                                                   const float* __restrict__ qmap, long long n, float* x) {
  __shared__ float s_map[256];
  s_map[threadIdx.x] = qmap[threadIdx.x];
  __syncthreads();
  const long long b0 = static_cast<long long>(blockIdx.x) * Q8_BLOCK;
  const float a = absmax[blockIdx.x];
#pragma unroll
  for (int i = 0; i < Q8_BLOCK / 256; ++i) {
    const long long idx = b0 + i * 256 + threadIdx.x;
    if (idx < n) x[idx] = s_map[codes[idx]] * a;
  }
}

}  // namespace dp

extern "C" int dp_quantize_blockwise(const float* x, long long n, const float* qmap256, void* codes_u8, float* absmax,
                                     cudaStream_t stream) {
  using namespace dp;
  DP_REQUIRE(n >= 1 && x && qmap256 && codes_u8 && absmax, "dp_quantize_blockwise: missing operand");
  const unsigned blocks = static_cast<unsigned>((n + Q8_BLOCK - 1) / Q8_BLOCK);
  quantize_blockwise_kernel<<<blocks, 256, 0, stream>>>(x, n, qmap256, static_cast<uint8_t*>(codes_u8), absmax);
  return check_cuda(cudaGetLastError(), "quantize_blockwise");
}

extern "C" int dp_dequantize_blockwise(const void* codes_u8, const float* absmax, const float* qmap256, long long n, float* x,
                                       cudaStream_t stream) {
  using namespace dp;
  DP_REQUIRE(n >= 1 && x && qmap256 && codes_u8 && absmax, "dp_dequantize_blockwise: missing operand");
  const unsigned blocks = static_cast<unsigned>((n + Q8_BLOCK - 1) / Q8_BLOCK);
  dequantize_blockwise_kernel<<<blocks, 256, 0, stream>>>(static_cast<const uint8_t*>(codes_u8), absmax, qmap256, n, x);
  return check_cuda(cudaGetLastError(), "dequantize_blockwise");
}
