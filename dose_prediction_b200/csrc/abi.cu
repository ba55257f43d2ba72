// Host-side plumbing shared by the C-ABI entry points: thread-local error string, CUDA error
// translation, cuTensorMapEncodeTiled via the runtime's driver entry point, device queries.
#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstring>

#include "common.cuh"
#include "dose_b200.h"

namespace dp {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int check_cuda(cudaError_t e, const char* what) {
  if (e == cudaSuccess) return 0;
  set_error("%s failed: %s (%s)", what, cudaGetErrorName(e), cudaGetErrorString(e));
  return 2;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (fn == nullptr) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

int encode_tiled(CUtensorMap* map, CUtensorMapDataType dtype, uint32_t rank, const void* base, const uint64_t* dims,
                 const uint64_t* strides_bytes, const uint32_t* box, CUtensorMapSwizzle swizzle) {
  EncodeTiledFn fn = get_encode_fn();
  if (fn == nullptr) {
    set_error("cuTensorMapEncodeTiled entry point unavailable (no CUDA driver?)");
    return 3;
  }
  cuuint64_t d[5], s[4];
  cuuint32_t b[5], es[5];
  for (uint32_t i = 0; i < rank; ++i) { d[i] = dims[i]; b[i] = box[i]; es[i] = 1; }
  for (uint32_t i = 0; i + 1 < rank; ++i) s[i] = strides_bytes[i];
  const CUresult r = fn(map, dtype, rank, const_cast<void*>(base), d, s, b, es, CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle,
                        CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed with CUresult %d (rank %u, dims %llu %llu %llu.., box %u %u %u..)",
              static_cast<int>(r), rank, (unsigned long long)d[0], (unsigned long long)d[1],
              (unsigned long long)(rank > 2 ? d[2] : 0), b[0], b[1], rank > 2 ? b[2] : 0);
    return 4;
  }
  return 0;
}

constexpr int kMaxDevices = 64;

int sm_count() {
  static std::atomic<int> cache[kMaxDevices];            // per device: a process may drive GPUs of different sizes
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= kMaxDevices) return 148;
  int n = cache[dev].load(std::memory_order_relaxed);
  if (n == 0) {
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    cache[dev].store(n, std::memory_order_relaxed);
  }
  return n;
}

bool first_use_on_device(int family) {
  static std::atomic<unsigned char> done[kMaxDevices][KF_COUNT];
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= kMaxDevices || family < 0 || family >= KF_COUNT) return true;
  // racing threads may both see "first" and both set the (idempotent) attribute: harmless
  return done[dev][family].exchange(1, std::memory_order_acq_rel) == 0;
}

}  // namespace dp

extern "C" const char* dp_last_error(void) { return dp::g_err; }
extern "C" int dp_abi_version(void) { return 1; }
extern "C" int dp_device_sm_count(void) { return dp::sm_count(); }
