// Host-side plumbing shared by the C-ABI entry points: thread-local error string, CUDA error
// translation, cuTensorMapEncodeTiled via the runtime's driver entry point, device queries.
#include <atomic>
#include <cstdarg>
#include <mutex>
#include <unordered_map>
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

// ---- tensor-map cache (one per process, keyed by everything cuTensorMapEncodeTiled sees).  The plans replay the same
// (buffer, shape, box) combinations every step; in eager replays (training, profiling) each launch used to re-encode its
// maps through the driver (~1-2 us each, ~700 per training step).  A CUtensorMap is a plain 128-byte value, so cached
// copies stay valid for as long as the device buffer does; dp_handle_destroy() drops them.
namespace {
struct MapKey {
  uint64_t v[16];
  bool operator==(const MapKey& o) const { return memcmp(v, o.v, sizeof(v)) == 0; }
};
struct MapKeyHash {
  size_t operator()(const MapKey& k) const {
    uint64_t h = 1469598103934665603ull;
    for (uint64_t x : k.v) { h ^= x; h *= 1099511628211ull; }
    return static_cast<size_t>(h);
  }
};
std::mutex g_map_mutex;
std::unordered_map<MapKey, CUtensorMap, MapKeyHash>& map_cache() {
  static std::unordered_map<MapKey, CUtensorMap, MapKeyHash> c;
  return c;
}
}  // namespace

void clear_tensor_map_cache() {
  std::lock_guard<std::mutex> lock(g_map_mutex);
  map_cache().clear();
}

static int encode_tiled_uncached(CUtensorMap* map, CUtensorMapDataType dtype, uint32_t rank, const void* base, const uint64_t* dims,
                                 const uint64_t* strides_bytes, const uint32_t* box, const uint32_t* elem_strides,
                                 CUtensorMapSwizzle swizzle);

int encode_tiled(CUtensorMap* map, CUtensorMapDataType dtype, uint32_t rank, const void* base, const uint64_t* dims,
                 const uint64_t* strides_bytes, const uint32_t* box, CUtensorMapSwizzle swizzle) {
  return encode_tiled_es(map, dtype, rank, base, dims, strides_bytes, box, nullptr, swizzle);
}

int encode_tiled_es(CUtensorMap* map, CUtensorMapDataType dtype, uint32_t rank, const void* base, const uint64_t* dims,
                    const uint64_t* strides_bytes, const uint32_t* box, const uint32_t* elem_strides, CUtensorMapSwizzle swizzle) {
  MapKey key{};
  key.v[0] = reinterpret_cast<uint64_t>(base);
  key.v[1] = (static_cast<uint64_t>(dtype) << 40) | (static_cast<uint64_t>(rank) << 32) | static_cast<uint64_t>(swizzle);
  if (elem_strides != nullptr)
    for (uint32_t i = 0; i < rank && i < 5; ++i) key.v[1] ^= static_cast<uint64_t>(elem_strides[i]) << (44 + 4 * i);
  for (uint32_t i = 0; i < rank && i < 5; ++i) key.v[2 + i] = dims[i];
  for (uint32_t i = 0; i + 1 < rank && i < 4; ++i) key.v[7 + i] = strides_bytes[i];
  for (uint32_t i = 0; i < rank && i < 5; ++i) key.v[11 + i] = box[i];
  {
    std::lock_guard<std::mutex> lock(g_map_mutex);
    auto it = map_cache().find(key);
    if (it != map_cache().end()) { *map = it->second; return 0; }
  }
  if (int rc = encode_tiled_uncached(map, dtype, rank, base, dims, strides_bytes, box, elem_strides, swizzle)) return rc;
  std::lock_guard<std::mutex> lock(g_map_mutex);
  if (map_cache().size() > 65536) map_cache().clear();          // bound the cache (plans rebuilt with fresh buffers)
  map_cache()[key] = *map;
  return 0;
}

static int encode_tiled_uncached(CUtensorMap* map, CUtensorMapDataType dtype, uint32_t rank, const void* base, const uint64_t* dims,
                                 const uint64_t* strides_bytes, const uint32_t* box, const uint32_t* elem_strides,
                                 CUtensorMapSwizzle swizzle) {
  EncodeTiledFn fn = get_encode_fn();
  if (fn == nullptr) {
    set_error("cuTensorMapEncodeTiled entry point unavailable (no CUDA driver?)");
    return 3;
  }
  cuuint64_t d[5], s[4];
  cuuint32_t b[5], es[5];
  for (uint32_t i = 0; i < rank; ++i) { d[i] = dims[i]; b[i] = box[i]; es[i] = elem_strides ? elem_strides[i] : 1; }
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

// ---- dp_handle: the per-device library state a host integration owns explicitly.  The library keeps no other mutable
// state than what a handle stands for: the kernels' per-device function attributes (configured on first use) and the
// tensor-map cache.  Creating a handle selects nothing globally (entry points act on the CALLER's current device and
// stream, like the CUDA runtime); destroying the last handle of a process drops the cached tensor maps.
struct dp_handle { int device; };
static std::atomic<int> g_handles{0};
extern "C" dp_handle* dp_handle_create(int device) {
  int count = 0;
  if (cudaGetDeviceCount(&count) != cudaSuccess || device < 0 || device >= count) {
    dp::set_error("dp_handle_create: device %d not present (%d visible)", device, count);
    return nullptr;
  }
  g_handles.fetch_add(1);
  return new dp_handle{device};
}
extern "C" int dp_handle_device(const dp_handle* h) { return h ? h->device : -1; }
extern "C" void dp_handle_destroy(dp_handle* h) {
  if (h == nullptr) return;
  delete h;
  if (g_handles.fetch_sub(1) == 1) dp::clear_tensor_map_cache();
}
