// Input pipeline on the device: the per-volume numpy / monai transforms of
// DosePrediction/DataLoader/dataloader_OpenKBP_monai.py (prepare_data, :160-243) after the NIfTI files are read:
//   Transposed(indices=[2,1,0]) (:173), NormalizePTVTr (:113-125), MyIntensityNormalTransform (:137-146),
//   NormalizeDoseTr (:128-134), ConcatItemsd -> 'Input' = [PTV, 7 OARs, CT] and 'GT' = [dose/70, dose_mask] (:195-201),
//   RandShiftIntensityd on the CT (:189-193), RandFlipd x3 and RandRotate90d (:214-236).
// One fused kernel builds Input and GT straight from the raw arrays (masks uint8, CT int16 or fp32, dose fp32); one
// gather kernel applies any combination of flips and a 90-degree rotation.  Random decisions are made by the caller.
#include "common.cuh"
#include "dose_b200.h"

namespace dp {

struct PrepareParams {
  const unsigned char* ptv[3];      // PTV70, PTV63, PTV56 (null = not delineated)
  const unsigned char* oar[7];
  const short* ct_i16; const float* ct_f32;
  const float* dose; const unsigned char* dose_mask;
  int A, B, C;                      // raw array dims [A][B][C]; outputs are [C][B][A] (Transposed [2,1,0])
  float a_min, a_max, ct_shift;
  float* input; float* gt;          // [9][C][B][A], [2][C][B][A]
};
// 32x32 tiles of the (A, C) plane through shared memory: reads are contiguous along C, writes along A
__global__ void __launch_bounds__(256) prepare_input_kernel(const PrepareParams p) {
  __shared__ float tile[11][32][33];
  const int b = blockIdx.z;
  const int a0 = blockIdx.y * 32, c0 = blockIdx.x * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int i = ty; i < 32; i += 8) {
    const int a = a0 + i, c = c0 + tx;
    float v[11];
#pragma unroll
    for (int k = 0; k < 11; ++k) v[k] = 0.f;
    if (a < p.A && c < p.C) {
      const size_t idx = (static_cast<size_t>(a) * p.B + b) * p.C + c;
      const double p70 = p.ptv[0] ? p.ptv[0][idx] : 0, p63 = p.ptv[1] ? p.ptv[1][idx] : 0, p56 = p.ptv[2] ? p.ptv[2][idx] : 0;
      v[0] = static_cast<float>(70.0 / 70. * p70 + 63.0 / 70. * p63 + 56.0 / 70. * p56);
#pragma unroll
      for (int k = 0; k < 7; ++k) v[1 + k] = p.oar[k] ? static_cast<float>(p.oar[k][idx]) : 0.f;
      float ct = p.ct_i16 ? static_cast<float>(p.ct_i16[idx]) : p.ct_f32[idx];
      ct = fminf(fmaxf(ct, p.a_min), p.a_max);
      v[8] = ct / 1000.f + p.ct_shift;
      v[9] = p.dose ? p.dose[idx] / 70.0f : 0.f;
      v[10] = p.dose_mask ? static_cast<float>(p.dose_mask[idx]) : 0.f;
    }
#pragma unroll
    for (int k = 0; k < 11; ++k) tile[k][i][tx] = v[k];
  }
  __syncthreads();
  const size_t plane = static_cast<size_t>(p.C) * p.B * p.A;
  for (int i = ty; i < 32; i += 8) {
    const int c = c0 + i, a = a0 + tx;
    if (a < p.A && c < p.C) {
      const size_t o = (static_cast<size_t>(c) * p.B + b) * p.A + a;
#pragma unroll
      for (int k = 0; k < 9; ++k) p.input[k * plane + o] = tile[k][tx][i];
      if (p.gt) { p.gt[o] = tile[9][tx][i]; p.gt[plane + o] = tile[10][tx][i]; }
    }
  }
}

// out[c] = rot90(flip(in[c], axes), k, axes=(0,1)) for in [C][S0][S1][S2] (np.flip per axis, then np.rot90)
__global__ void __launch_bounds__(256) flip_rot90_kernel(const float* in, float* out, int C, int S0, int S1, int S2, int f0,
                                                         int f1, int f2, int k) {
  const int O0 = (k & 1) ? S1 : S0, O1 = (k & 1) ? S0 : S1;
  const long long total = static_cast<long long>(C) * O0 * O1 * S2;
  const long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (idx >= total) return;
  const int l = static_cast<int>(idx % S2);
  const int j = static_cast<int>((idx / S2) % O1);
  const int i = static_cast<int>((idx / (static_cast<long long>(S2) * O1)) % O0);
  const int c = static_cast<int>(idx / (static_cast<long long>(S2) * O1 * O0));
  // invert np.rot90(m, k, (0, 1)): k=1: out[i][j] = m[j][S1-1-i]; k=2: m[S0-1-i][S1-1-j]; k=3: m[S0-1-j][i]
  int a, b;
  switch (k & 3) {
    case 1: a = j; b = S1 - 1 - i; break;
    case 2: a = S0 - 1 - i; b = S1 - 1 - j; break;
    case 3: a = S0 - 1 - j; b = i; break;
    default: a = i; b = j; break;
  }
  if (f0) a = S0 - 1 - a;
  if (f1) b = S1 - 1 - b;
  const int cc = f2 ? S2 - 1 - l : l;
  out[idx] = in[((static_cast<size_t>(c) * S0 + a) * S1 + b) * S2 + cc];
}

}  // namespace dp

using namespace dp;

extern "C" int dp_prepare_input(const void* const* ptv_u8, const void* const* oar_u8, const void* ct_i16, const float* ct_f32,
                                const float* dose, const void* dose_mask_u8, int A, int B, int C, float a_min, float a_max,
                                float ct_shift, float* input, float* gt, cudaStream_t stream) {
  DP_REQUIRE((ct_i16 != nullptr) != (ct_f32 != nullptr), "prepare_input: exactly one CT array (int16 or fp32)");
  DP_REQUIRE(input != nullptr, "prepare_input: no output");
  PrepareParams p{};
  for (int i = 0; i < 3; ++i) p.ptv[i] = static_cast<const unsigned char*>(ptv_u8 ? ptv_u8[i] : nullptr);
  for (int i = 0; i < 7; ++i) p.oar[i] = static_cast<const unsigned char*>(oar_u8 ? oar_u8[i] : nullptr);
  p.ct_i16 = static_cast<const short*>(ct_i16); p.ct_f32 = ct_f32; p.dose = dose;
  p.dose_mask = static_cast<const unsigned char*>(dose_mask_u8);
  p.A = A; p.B = B; p.C = C; p.a_min = a_min; p.a_max = a_max; p.ct_shift = ct_shift; p.input = input; p.gt = gt;
  dim3 grid((C + 31) / 32, (A + 31) / 32, B);
  prepare_input_kernel<<<grid, 256, 0, stream>>>(p);
  return check_cuda(cudaGetLastError(), "prepare_input");
}

extern "C" int dp_flip_rot90(const float* in, float* out, int C, int S0, int S1, int S2, int flip0, int flip1, int flip2,
                             int k, cudaStream_t stream) {
  DP_REQUIRE(in != out, "flip_rot90: in-place is not supported");
  const long long total = static_cast<long long>(C) * S0 * S1 * S2;
  flip_rot90_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, stream>>>(in, out, C, S0, S1, S2, flip0, flip1, flip2, k & 3);
  return check_cuda(cudaGetLastError(), "flip_rot90");
}
