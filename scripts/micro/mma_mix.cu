// Micro-benchmark: sustained rate of the depth-stacked conv's MMA stream for different column counts per chunk.
// Mimics conv3d_stack_kernel<3,4>'s issue pattern: two issuing warps x 2 tiles, per "plane" 9 taps x {chunk a: N = Na,
// chunk b: N = Nb} (Nb = 0: one chunk; Na > 128: one 256-column tile per issuer), A descriptors of the real halo patch (SBO = 34 x 16 B, taps shift the start by
// 16 B / one patch row), B advancing one tap block per tap.  Reports SM cycles AND wall time (power capping shows up in
// the latter only).   nvcc -arch=sm_100a -O3 -o mma_mix mma_mix.cu && ./mma_mix
#include <cstdio>
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include "../../dose_prediction_b200/csrc/common.cuh"
using namespace dp;
namespace dp { void set_error(const char*, ...) {} int check_cuda(cudaError_t, const char*) { return 0; } }

__device__ __forceinline__ void mma(uint32_t d, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi, uint32_t idesc) {
  asm volatile("{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\tmov.b64 da, {%1, %2};\n\tmov.b64 db, {%3, %4};\n\tsetp.ne.b32 p, %6, 0;\n\t"
               "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n\t}"
               ::"r"(d), "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(1u) : "memory");
}

__global__ void __launch_bounds__(128, 1) mma_mix(int Na, int Nb, int planes, int aligned, int M, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint64_t bar[2];
  __shared__ uint32_t tbase;
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 196 * 1024 / 2; i += blockDim.x) {       // random operands in [-1, 1): power draw depends on the data
    uint32_t h = (i + 1) * 2654435761u; h ^= h >> 15; h *= 2246822519u; h ^= h >> 13;
    reinterpret_cast<__half*>(smem)[i] = __float2half_rn((h & 0xFFFF) / 32768.0f - 1.0f);
  }
  if (threadIdx.x == 0) { mbar_init(&bar[0], 1); mbar_init(&bar[1], 1); fence_barrier_init(); }
  if (warp == 0) tmem_alloc<512>(&tbase);
  fence_proxy_async();
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tb = tbase;
  if ((warp == 1 || warp == 2) && lane == 0) {
    const int w = warp - 1;
    const uint32_t PW = aligned ? 40u : 34u, PH = 18u;               // patch row pitch in 16-B units
    const uint32_t ia = make_idesc_f16(M, Na), ib = make_idesc_f16(M, Nb > 0 ? Nb : 16);
    const uint32_t sa16 = smem_u32(smem) >> 4, sb16 = (smem_u32(smem) + 100 * 1024) >> 4;
    const uint32_t a_hi = PW | (1u << 14), b_hi = 8u | (1u << 14);
    const uint32_t a_lo_c = ((PH * PW) & 0x3FFFu) << 16, b_lo_c = ((Na > 128 ? 256u : 128u) << 16);
    const uint32_t brows = Na > 128 ? 256u : 128u;
    const uint32_t tap_b16 = (32u * brows) >> 4;                     // [2][rows][8] fp16 per tap
    const uint32_t tile_cols = Na > 128 ? 256u : 128u;
    const uint32_t a_stage16 = (2u * PH * PW * 16u + 1023u) / 1024u * 64u;
    long long t0 = clock64();
    for (int pl = 0; pl < planes; ++pl) {
      const uint32_t a0 = a_lo_c | (sa16 + (pl % 2) * a_stage16 + w * 16);
#pragma unroll
      for (int ch = 0; ch < 2; ++ch) {
        if (ch == 1 && Nb == 0) break;
        const uint32_t idesc = ch ? ib : ia;
        const uint32_t ac = a0 + (ch ? 2 * a_stage16 : 0);
#pragma unroll
        for (int kh = 0; kh < 3; ++kh)
#pragma unroll
          for (int kw = 0; kw < 3; ++kw)
#pragma unroll
            for (int t = 0; t < (Na > 128 ? 1 : 2); ++t)
              mma(tb + (w * (Na > 128 ? 1 : 2) + t) * tile_cols, ac + kh * PW + (aligned ? 0 : kw) + t * 8, a_hi, (b_lo_c | sb16) + (kh * 3 + kw) * tap_b16, b_hi, idesc);
      }
    }
    umma_commit(&bar[w]);
    mbar_wait(&bar[w], 0, nullptr);
    long long t1 = clock64();
    if (w == 0) out[blockIdx.x] = t1 - t0;
  }
  tc_fence_before(); __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc<512>(tb); }
}

int main() {
  long long* d; cudaMalloc(&d, 148 * sizeof(long long));
  const int smem = 200 * 1024;
  cudaFuncSetAttribute(mma_mix, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int cfg[][2] = {{128, 128}, {128, 64}, {256, 0}, {256, 128}, {256, 256}, {192, 0}, {160, 0}, {128, 0}};
  for (int M = 128; M >= 128; M -= 64)
  for (int aligned = 0; aligned < 1; ++aligned)
    for (auto& c : cfg) {
      const int planes = 20000;
      for (int rep = 0; rep < 2; ++rep) {       // second repetition = warm (sustained) number
        cudaEventRecord(e0);
        mma_mix<<<148, 128, smem>>>(c[0], c[1], planes, aligned, M, d);
        cudaEventRecord(e1);
        cudaError_t e = cudaDeviceSynchronize();
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        long long h[148]; cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
        double avg = 0; for (int i = 0; i < 148; ++i) avg += h[i]; avg /= 148;
        const double mmas = planes * 9.0 * (c[0] > 128 ? 2 : 4) * (c[1] ? 2 : 1);     // per CTA (both issuers)
        const double cols = planes * 9.0 * (c[0] > 128 ? 2 : 4) * (c[0] + c[1]);
        if (rep) printf("M=%d aligned=%d Na=%3d Nb=%3d : %6.1f cycles per tap (4 tiles, both chunks), %5.1f per MMA, %6.2f ms wall, %6.0f TF/s executed, clk %.0f MHz  %s\n",
               M, aligned, c[0], c[1], avg / (planes * 9.0), avg / mmas * 1.0, ms, cols * 128 * 16 * 2 * 148 / (ms * 1e-3) / 1e12,
               avg / (ms * 1e-3) / 1e6, cudaGetErrorString(e));
      }
    }
  return 0;
}
