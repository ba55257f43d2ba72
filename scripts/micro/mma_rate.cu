// Micro-benchmark: issue rate / throughput of tcgen05.mma (M=128, K=16, fp16) vs N on sm_100a.
// One CTA per SM, one thread issues ITERS MMAs, smem operands fixed (no-swizzle K-major, garbage data).
#include <cstdio>
#include <cuda_runtime.h>
#include "../../dose_prediction_b200/csrc/common.cuh"
using namespace dp;
namespace dp { void set_error(const char*, ...) {} int check_cuda(cudaError_t, const char*) { return 0; } }

__global__ void __launch_bounds__(128, 1) mma_rate(int N, int iters, int ndst, int a_step, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tbase;
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 64 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
  if (warp == 0) tmem_alloc<512>(&tbase);
  fence_proxy_async();
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tb = tbase;
  if (warp == 1 && lane == 0) {
    const uint32_t idesc = make_idesc_f16(128, N);
    const uint32_t sa16 = smem_u32(smem) >> 4, sb16 = (smem_u32(smem) + 32768) >> 4;
    const uint32_t a_hi = 14u | (1u << 14), b_hi = 8u | (1u << 14);
    const uint32_t a_lo_c = (308u << 16), b_lo_c = (static_cast<uint32_t>(N) << 16);
    long long t0 = clock64();
    uint32_t a = sa16;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        const uint32_t d = tb + (ndst > 1 ? (t % ndst) * N : 0);
        asm volatile("{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\tmov.b64 da, {%1, %2};\n\tmov.b64 db, {%3, %4};\n\tsetp.ne.b32 p, %6, 0;\n\t"
                     "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n\t}"
                     ::"r"(d), "r"(a_lo_c | (a + t * 8)), "r"(a_hi), "r"(b_lo_c | sb16), "r"(b_hi), "r"(idesc), "r"(1u) : "memory");
      }
      a = sa16 + ((a - sa16 + a_step) & 63);
    }
    umma_commit(&bar);
    mbar_wait(&bar, 0, nullptr);
    long long t1 = clock64();
    out[blockIdx.x] = t1 - t0;
  }
  tc_fence_before(); __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc<512>(tb); }
}

int main() {
  long long* d; cudaMalloc(&d, 148 * sizeof(long long));
  cudaFuncSetAttribute(mma_rate, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  const int iters = 4000;
  for (int ndst = 1; ndst <= 4; ndst *= 4)
    for (int N : {16, 32, 64, 112, 128, 256}) {
      if (ndst * N > 512) continue;
      mma_rate<<<148, 128, 100 * 1024>>>(N, iters, ndst, 1, d);
      cudaError_t e = cudaDeviceSynchronize();
      long long h[148]; cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
      double avg = 0; for (int i = 0; i < 148; ++i) avg += h[i]; avg /= 148;
      printf("N=%3d ndst=%d : %.1f cycles/MMA (tensor floor %.0f)  %s\n", N, ndst, avg / (4.0 * iters), N / 2.0, cudaGetErrorString(e));
    }
  return 0;
}
