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

// out[c][i0][i1][i2] = in[c][...] with output axis a reading input axis perm[a], reversed when flip[a] (Orientationd:
// nibabel's ornt_transform applied as transposes + flips; monai 0.7.0 Orientation.__call__)
__global__ void __launch_bounds__(256) permute_flip_kernel(const float* in, float* out, int C, int S0, int S1, int S2, int p0, int p1,
                                                           int p2, int f0, int f1, int f2) {
  const int S[3] = {S0, S1, S2};
  const int O0 = S[p0], O1 = S[p1], O2 = S[p2];
  const long long total = static_cast<long long>(C) * O0 * O1 * O2;
  const long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (idx >= total) return;
  int o[3];
  o[2] = static_cast<int>(idx % O2);
  o[1] = static_cast<int>((idx / O2) % O1);
  o[0] = static_cast<int>((idx / (static_cast<long long>(O2) * O1)) % O0);
  const int c = static_cast<int>(idx / (static_cast<long long>(O2) * O1 * O0));
  if (f0) o[0] = O0 - 1 - o[0];
  if (f1) o[1] = O1 - 1 - o[1];
  if (f2) o[2] = O2 - 1 - o[2];
  int i[3];
  i[p0] = o[0]; i[p1] = o[1]; i[p2] = o[2];
  out[idx] = in[((static_cast<size_t>(c) * S0 + i[0]) * S1 + i[1]) * S2 + i[2]];
}

// ---- RandCropByPosNegLabeld (monai 0.7.0), device part.  Foreground = any label channel > 0; background = any image channel
// > threshold and not foreground (map_binary_to_indices).  Pass 1 counts both per 4096-voxel block; the host draws
// (numpy RandomState semantics need the totals) and names, per sample, a block and the rank of the wanted voxel inside it;
// pass 2 finds that voxel, clamps the centre (correct_crop_centers) and writes the crop origin; pass 3 crops.
constexpr int PN_BLOCK = 4096;
__device__ __forceinline__ void posneg_flags(const float* label, int lc, const float* image, int ic, float thr, long long vox, long long v,
                                             bool& fg, bool& bg) {
  fg = false;
  for (int c = 0; c < lc; ++c) fg = fg || (label[c * vox + v] > 0.f);
  bool img = image == nullptr;
  for (int c = 0; c < ic && !img; ++c) img = image[c * vox + v] > thr;
  bg = img && !fg;
}
__global__ void __launch_bounds__(256) posneg_count_kernel(const float* label, int lc, const float* image, int ic, float thr, long long vox,
                                                           int* counts /* [blocks][2] */) {
  __shared__ int s_fg, s_bg;
  if (threadIdx.x == 0) { s_fg = 0; s_bg = 0; }
  __syncthreads();
  int nf = 0, nb = 0;
  for (int i = threadIdx.x; i < PN_BLOCK; i += 256) {
    const long long v = static_cast<long long>(blockIdx.x) * PN_BLOCK + i;
    if (v >= vox) break;
    bool fg, bg;
    posneg_flags(label, lc, image, ic, thr, vox, v, fg, bg);
    nf += fg; nb += bg;
  }
  atomicAdd(&s_fg, nf);
  atomicAdd(&s_bg, nb);
  __syncthreads();
  if (threadIdx.x == 0) { counts[2 * blockIdx.x] = s_fg; counts[2 * blockIdx.x + 1] = s_bg; }
}
// one thread block per sample: walk the named 4096-voxel block in order, find the rank-th voxel of the wanted kind
__global__ void __launch_bounds__(32) posneg_select_kernel(const float* label, int lc, const float* image, int ic, float thr, int S0, int S1,
                                                           int S2, int R, const int* pick /* [ns][3] = {block, rank, want_fg} */,
                                                           int* roi_start /* [ns][3] */) {
  const long long vox = static_cast<long long>(S0) * S1 * S2;
  const int s = blockIdx.x;
  const int blk = pick[3 * s], rank = pick[3 * s + 1], want_fg = pick[3 * s + 2];
  __shared__ long long found;
  if (threadIdx.x == 0) found = -1;
  __syncwarp();
  int seen = 0;
  for (int base = 0; base < PN_BLOCK && found < 0; base += 32) {
    const long long v = static_cast<long long>(blk) * PN_BLOCK + base + threadIdx.x;
    bool hit = false;
    if (v < vox) {
      bool fg, bg;
      posneg_flags(label, lc, image, ic, thr, vox, v, fg, bg);
      hit = want_fg ? fg : bg;
    }
    const unsigned m = __ballot_sync(0xffffffffu, hit);
    const int before = __popc(m & ((1u << threadIdx.x) - 1u));
    if (hit && seen + before == rank) found = v;
    seen += __popc(m);
    __syncwarp();
  }
  __syncwarp();
  if (threadIdx.x == 0) {
    const long long v = found < 0 ? 0 : found;
    int c[3] = {static_cast<int>(v / (static_cast<long long>(S1) * S2)), static_cast<int>((v / S2) % S1), static_cast<int>(v % S2)};
    const int S[3] = {S0, S1, S2};
    for (int a = 0; a < 3; ++a) {                      // correct_crop_centers + SpatialCrop(roi_center, roi_size)
      const int valid_start = R / 2;
      int valid_end = static_cast<int>(static_cast<float>(S[a] + 1) - static_cast<float>(R) / 2.f);   // astype(uint16): truncation
      if (valid_start == valid_end) valid_end += 1;
      int ci = c[a];
      if (ci < valid_start) ci = valid_start;
      if (ci >= valid_end) ci = valid_end - 1;
      int start = ci - R / 2;
      if (start < 0) start = 0;
      roi_start[3 * s + a] = start;
    }
  }
}
__global__ void __launch_bounds__(256) crop_samples_kernel(const float* in, int C, int S0, int S1, int S2, int R, const int* roi_start,
                                                           float* out /* [ns][C][R][R][R] */) {
  const long long rv = static_cast<long long>(R) * R * R;
  const long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (idx >= rv * C) return;
  const int s = blockIdx.y;
  const int c = static_cast<int>(idx / rv);
  const long long r = idx % rv;
  const int z = static_cast<int>(r % R), y = static_cast<int>((r / R) % R), x = static_cast<int>(r / (static_cast<long long>(R) * R));
  const int a = roi_start[3 * s] + x, b = roi_start[3 * s + 1] + y, d = roi_start[3 * s + 2] + z;
  float v = 0.f;
  if (a < S0 && b < S1 && d < S2) v = in[((static_cast<size_t>(c) * S0 + a) * S1 + b) * S2 + d];
  out[(static_cast<size_t>(s) * C + c) * rv + r] = v;
}

}  // namespace dp

using namespace dp;

extern "C" int dp_permute_flip(const float* in, float* out, int C, int S0, int S1, int S2, int perm0, int perm1, int perm2,
                               int flip0, int flip1, int flip2, cudaStream_t stream) {
  DP_REQUIRE(in != out, "permute_flip: in-place is not supported");
  const int seen = (1 << perm0) | (1 << perm1) | (1 << perm2);
  DP_REQUIRE(perm0 >= 0 && perm0 < 3 && perm1 >= 0 && perm1 < 3 && perm2 >= 0 && perm2 < 3 && seen == 7,
             "permute_flip: (%d, %d, %d) is not a permutation of the three spatial axes", perm0, perm1, perm2);
  const long long total = static_cast<long long>(C) * S0 * S1 * S2;
  permute_flip_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, stream>>>(in, out, C, S0, S1, S2, perm0, perm1, perm2,
                                                                                     flip0, flip1, flip2);
  return check_cuda(cudaGetLastError(), "permute_flip");
}

extern "C" int dp_posneg_count(const float* label, int label_channels, const float* image, int image_channels, float image_threshold,
                               long long vox, int* block_counts, cudaStream_t stream) {
  DP_REQUIRE(label != nullptr && block_counts != nullptr && label_channels >= 1, "posneg_count: missing operand");
  const unsigned blocks = static_cast<unsigned>((vox + PN_BLOCK - 1) / PN_BLOCK);
  posneg_count_kernel<<<blocks, 256, 0, stream>>>(label, label_channels, image, image_channels, image_threshold, vox, block_counts);
  return check_cuda(cudaGetLastError(), "posneg_count");
}

extern "C" int dp_posneg_crop(const float* label, int label_channels, const float* image, int image_channels, float image_threshold,
                              int S0, int S1, int S2, int R, int n_samples, const int* pick_dev, int* roi_start_dev, int n_src,
                              const float* const* src, const int* src_channels, float* const* dst, cudaStream_t stream) {
  DP_REQUIRE(R >= 1 && R <= S0 && R <= S1 && R <= S2, "posneg_crop: the crop ROI %d^3 is larger than the image %dx%dx%d", R, S0, S1, S2);
  DP_REQUIRE(n_samples >= 1 && pick_dev != nullptr && roi_start_dev != nullptr, "posneg_crop: missing operand");
  posneg_select_kernel<<<n_samples, 32, 0, stream>>>(label, label_channels, image, image_channels, image_threshold, S0, S1, S2, R,
                                                     pick_dev, roi_start_dev);
  DP_CHECK(cudaGetLastError());
  const long long rv = static_cast<long long>(R) * R * R;
  for (int i = 0; i < n_src; ++i) {
    dim3 grid(static_cast<unsigned>((rv * src_channels[i] + 255) / 256), static_cast<unsigned>(n_samples));
    crop_samples_kernel<<<grid, 256, 0, stream>>>(src[i], src_channels[i], S0, S1, S2, R, roi_start_dev, dst[i]);
    DP_CHECK(cudaGetLastError());
  }
  return 0;
}

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
