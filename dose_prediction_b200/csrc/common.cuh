// Shared device helpers for the sm_100a kernels: mbarrier, TMA, tcgen05/TMEM PTX wrappers,
// activation functions and the C-ABI error plumbing.  Everything here is hand-written inline PTX;
// no CUTLASS/CuTe headers are included.
#pragma once
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace dp {

// ----------------------------------------------------------------------------- activations
enum Act : int { ACT_NONE = 0, ACT_RELU = 1, ACT_LRELU = 2, ACT_MISH = 3, ACT_GELU = 4 };

// Mish: x * tanh(softplus(x)) = x * u / (u + 2) with u = e^x (e^x + 2).  Branch-free: the exponent is clamped at torch's
// softplus threshold 20 (beyond it u / (u + 2) rounds to 1 in fp32, i.e. the result is x, exactly what torch returns), and
// the raw ex2.approx / rcp.approx units are used without the range fix-ups of __expf / __fdividef (nothing here can
// overflow or go denormal in a way that matters: ncu showed 9 FMUL + 3 FSETP + 2 branches per element for the old form).
__device__ __forceinline__ float mish_fast(float x) {
  float t, r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(fminf(x, 20.f) * 1.4426950408889634f));
  const float u = t * (t + 2.f);
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(u + 2.f));
  return x * (u * r);
}

// GELU(x) = x Phi(x) with erf from Abramowitz-Stegun 7.1.26 (|error| <= 1.5e-7 absolute, far below the fp16 rounding of the
// stored activation) on the raw rcp / ex2 units: ~14 instructions against ~40 of erff().  Used by the GEMM epilogue of the
// inference MLP (fc1 + GELU), whose epilogue is ALU-issue bound (profiles/r3_gemm_pair.md); the training kernels keep erff.
__device__ __forceinline__ float gelu_fast(float x) {
  const float z = fabsf(x) * 0.70710678118654752f;
  float t, e;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(fmaf(0.3275911f, z, 1.f)));
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(-z * z * 1.4426950408889634f));
  const float poly = t * fmaf(t, fmaf(t, fmaf(t, fmaf(t, 1.061405429f, -1.453152027f), 1.421413741f), -0.284496736f), 0.254829592f);
  const float erf_abs = fmaf(-poly, e, 1.f);
  return 0.5f * x * (1.f + copysignf(erf_abs, x));
}

__device__ __forceinline__ float act_apply(float x, int act) {
  switch (act) {
    case ACT_RELU: return fmaxf(x, 0.f);
    case ACT_LRELU: return x > 0.f ? x : 0.01f * x;
    case ACT_MISH: return mish_fast(x);
    case ACT_GELU: return 0.5f * x * (1.f + erff(x * 0.70710678118654752f));
    default: return x;
  }
}

// ----------------------------------------------------------------------------- smem / mbarrier
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded spin: a protocol bug must surface as an error (flag + __trap -> sticky launch failure), never as a hung GPU
// box and never as silently wrong results.
// mbar_wait      : latency-critical waiter (the MMA-issuing warp) — polls back to back.
// mbar_wait_relaxed : producer / epilogue warps — sleep between polls so their spin loops do not take
//                  issue slots from the MMA warp sharing the SM sub-partition (ncu: 72 % of all warp
//                  samples of the stacked conv kernel sat in these loops).
__device__ __forceinline__ bool mbar_wait(uint64_t* bar, uint32_t parity, int* err_flag) {
#pragma unroll 1
  for (uint32_t it = 0; it < (1u << 22); ++it) {
    if (mbar_try_wait(bar, parity)) return true;
    if (it > 4096) __nanosleep(64);
  }
  if (err_flag) atomicExch(err_flag, 1);
  __threadfence_system();
  __trap();              // a protocol fault must never yield silent garbage: the launch fails, the context reports it
  return false;
}
__device__ __forceinline__ bool mbar_wait_relaxed(uint64_t* bar, uint32_t parity, int* err_flag) {
  if (mbar_try_wait(bar, parity)) return true;
#pragma unroll 1
  for (uint32_t it = 0; it < (1u << 22); ++it) {
    __nanosleep(it < 64 ? 40 : 200);
    if (mbar_try_wait(bar, parity)) return true;
  }
  if (err_flag) atomicExch(err_flag, 1);
  __threadfence_system();
  __trap();
  return false;
}

__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(pred));
  return pred != 0;
}

// ----------------------------------------------------------------------------- TMA
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* m) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}
__device__ __forceinline__ void tma_load_5d(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1,
                                            int c2, int c3, int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.shared::cluster.global.tile.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2),
      "r"(c3), "r"(c4)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void bulk_load_1d(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// ----------------------------------------------------------------------------- tcgen05 / TMEM
template <int kCols>
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_dst) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)),
               "n"(kCols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
template <int kCols>
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(kCols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
// D[tmem] (+)= A[smem] * B[smem]; kind::f16 (fp16/bf16 operands, fp32 accumulate), one CTA.
__device__ __forceinline__ void umma_f16_ss(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                            uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// Arrive on an mbarrier once all previously issued tcgen05.mma of this thread have completed.
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
// 32 lanes x 16 consecutive fp32 columns -> 16 registers per thread (thread i <-> TMEM lane base+i).
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// Shared-memory matrix descriptor (PTX ISA "tcgen05 matrix descriptor"; bit layout as in the
// CUTLASS SmemDescriptor union): start>>4 [0,14), LBO>>4 [16,30), SBO>>4 [32,46), version=1 [46,48),
// layout type [61,64) (0 = no swizzle, 2 = 128B swizzle).
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes,
                                                   uint32_t layout_type) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((saddr >> 4) & 0x3FFF);
  d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(layout_type & 7) << 61;
  return d;
}
// Instruction descriptor for kind::f16: fp16 A/B (K-major), fp32 D, M x N.
__host__ __device__ __forceinline__ uint32_t make_idesc_f16(int m, int n) {
  uint32_t d = 0;
  d |= 1u << 4;                                  // D format: F32
  d |= 0u << 7;                                  // A format: F16
  d |= 0u << 10;                                 // B format: F16
  d |= static_cast<uint32_t>(n >> 3) << 17;      // N / 8
  d |= static_cast<uint32_t>(m >> 4) << 24;      // M / 16
  return d;
}

// 32-byte (one full sector) global accesses: a c8 fp32 voxel is 8 floats = 32 B; written as two 16-byte stores every
// instruction of a warp touches half sectors only (measured on the transposed-conv scatter: 0.32 -> 0.13 ms).
__device__ __forceinline__ void st_global_v8f(float* p, const float (&y)[8]) {
  asm volatile("st.global.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "f"(y[0]), "f"(y[1]), "f"(y[2]), "f"(y[3]),
               "f"(y[4]), "f"(y[5]), "f"(y[6]), "f"(y[7]) : "memory");
}
__device__ __forceinline__ void ld_global_v8f(const float* p, float (&y)[8]) {
  asm volatile("ld.global.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];" : "=f"(y[0]), "=f"(y[1]), "=f"(y[2]), "=f"(y[3]),
               "=f"(y[4]), "=f"(y[5]), "=f"(y[6]), "=f"(y[7]) : "l"(p) : "memory");
}
__device__ __forceinline__ void st_global_v8u(void* p, const uint32_t (&v)[8]) {
  asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]),
               "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]) : "memory");
}

// packed fp32 pairs: fma.rn.f32x2 (FFMA2) does two IEEE fp32 FMAs per issued instruction
__device__ __forceinline__ unsigned long long pack_f32x2(float lo, float hi) {
  unsigned long long r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void unpack_f32x2(unsigned long long v, float& lo, float& hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ unsigned long long fma_f32x2(unsigned long long a, unsigned long long b, unsigned long long c) {
  unsigned long long d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

}  // namespace dp

// ----------------------------------------------------------------------------- host-side error plumbing
namespace dp {
void set_error(const char* fmt, ...);
int check_cuda(cudaError_t e, const char* what);
// cuTensorMapEncodeTiled through the runtime's driver entry point (no -lcuda link dependency).
int encode_tiled(CUtensorMap* map, CUtensorMapDataType dtype, uint32_t rank, const void* base,
                 const uint64_t* dims, const uint64_t* strides_bytes /* rank-1 */, const uint32_t* box,
                 CUtensorMapSwizzle swizzle);
// the same with per-dimension element (traversal) strides: a box may take every n-th element of a dimension
int encode_tiled_es(CUtensorMap* map, CUtensorMapDataType dtype, uint32_t rank, const void* base, const uint64_t* dims,
                    const uint64_t* strides_bytes, const uint32_t* box, const uint32_t* elem_strides, CUtensorMapSwizzle swizzle);
void clear_tensor_map_cache();
int sm_count();
// true exactly once per (current CUDA device, family): function attributes (max dynamic shared memory) are per device,
// so every device a process drives configures its kernels on first use (no process-wide flag).
enum KernelFamily : int { KF_CONV_STACK = 0, KF_CONV_TC, KF_GEMM128, KF_GEMM256, KF_GEMM_PAIR128, KF_GEMM_PAIR256, KF_ATTN64, KF_ATTN128, KF_WGRAD, KF_POINTWISE_TC16, KF_POINTWISE_TC32, KF_POINTWISE_TC64, KF_COUNT };
bool first_use_on_device(int family);
}  // namespace dp

#define DP_CHECK(call)                                   \
  do {                                                   \
    int _rc = dp::check_cuda((call), #call);             \
    if (_rc) return _rc;                                 \
  } while (0)
#define DP_REQUIRE(cond, ...)                            \
  do {                                                   \
    if (!(cond)) {                                       \
      dp::set_error(__VA_ARGS__);                        \
      return 1;                                          \
    }                                                    \
  } while (0)
