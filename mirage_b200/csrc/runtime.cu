// mirage_b200/csrc/runtime.cu
//
// Host-side runtime shared by the kernels: last-error string, SM count, and a small cache of TMA
// descriptors (CUtensorMap) keyed by (pointer, dtype, dims, strides, box).
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <unordered_map>

#include "../../include/mirage_b200.h"
#include "common.cuh"

namespace mb200 {

static thread_local char g_last_error[1024] = "";

void set_last_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_last_error, sizeof(g_last_error), fmt, ap);
  va_end(ap);
}

static int g_sm_physical[64] = {};
static int g_sm_reserve = 0;

static int sm_physical() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 148;
  int& cached = g_sm_physical[dev & 63];
  if (cached == 0) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0)
      return 148;
    cached = n;
  }
  return cached;
}

int sm_count() {
  int n = sm_physical() - g_sm_reserve;
  n &= ~1;  // CTA pairs
  return n < 2 ? 2 : n;
}

static int g_dir = +1;
static int g_serpentine = -1;

bool serpentine_enabled() {
  if (g_serpentine < 0) {
    const char* e = getenv("MB_SERPENTINE");
    g_serpentine = (e != nullptr && e[0] == '1') ? 1 : 0;
  }
  return g_serpentine != 0;
}

int take_direction() {
  if (!serpentine_enabled()) return +1;
  g_dir = -g_dir;
  return g_dir;
}

void note_direction(int dir) { g_dir = dir; }

static int g_pdl = -1;

int pdl_mode() {
  if (g_pdl < 0) {
    const char* e = getenv("MB_PDL");
    g_pdl = (e != nullptr && e[0] >= '0' && e[0] <= '3') ? e[0] - '0' : 0;
  }
  return g_pdl;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, []() {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) ==
            cudaSuccess &&
        qres == cudaDriverEntryPointSuccess) {
      fn = reinterpret_cast<EncodeTiledFn>(p);
    }
  });
  return fn;
}

struct MapKey {
  uint64_t w[16];
  bool operator==(const MapKey& o) const { return memcmp(w, o.w, sizeof(w)) == 0; }
};
struct MapKeyHash {
  size_t operator()(const MapKey& k) const {
    uint64_t h = 1469598103934665603ull;
    for (int i = 0; i < 16; ++i) {
      h ^= k.w[i];
      h *= 1099511628211ull;
    }
    return static_cast<size_t>(h);
  }
};

static std::mutex g_map_mu;
static std::unordered_map<MapKey, CUtensorMap, MapKeyHash> g_map_cache;

int make_tensor_map(CUtensorMap* out, const void* base, TmaDtype dtype, int rank,
                    const uint64_t* dims, const uint64_t* strides_bytes, const uint32_t* box,
                    int swizzle_bytes) {
  MB_REQUIRE(rank >= 1 && rank <= 5, "make_tensor_map: rank %d unsupported", rank);
  MB_REQUIRE((reinterpret_cast<uintptr_t>(base) & 15) == 0,
             "make_tensor_map: base pointer %p is not 16-byte aligned", base);
  MapKey key;
  memset(&key, 0, sizeof(key));
  key.w[0] = reinterpret_cast<uint64_t>(base);
  key.w[1] = (static_cast<uint64_t>(swizzle_bytes) << 16) | (static_cast<uint64_t>(dtype) << 8) |
             static_cast<uint64_t>(rank);
  MB_REQUIRE(swizzle_bytes == 128 || swizzle_bytes == 64, "make_tensor_map: swizzle %d unsupported",
             swizzle_bytes);
  for (int i = 0; i < rank; ++i) {
    key.w[2 + i] = dims[i];
    key.w[7 + i] = (i + 1 < rank) ? strides_bytes[i] : 0;
    key.w[12 + (i >> 1)] |= static_cast<uint64_t>(box[i]) << (32 * (i & 1));
  }
  {
    std::lock_guard<std::mutex> lk(g_map_mu);
    auto it = g_map_cache.find(key);
    if (it != g_map_cache.end()) {
      *out = it->second;
      return 0;
    }
  }
  EncodeTiledFn fn = get_encode_fn();
  MB_REQUIRE(fn != nullptr, "cuTensorMapEncodeTiled entry point not available (driver too old?)");
  const size_t esize = (dtype == kTmaBF16) ? 2 : 4;
  for (int i = 0; i + 1 < rank; ++i)
    MB_REQUIRE((strides_bytes[i] & 15) == 0,
               "make_tensor_map: stride[%d]=%llu bytes is not a multiple of 16", i,
               (unsigned long long)strides_bytes[i]);
  MB_REQUIRE(box[0] * esize <= (size_t)swizzle_bytes,
             "make_tensor_map: inner box %u elements exceeds the %d-byte swizzle span", box[0],
             swizzle_bytes);
  cuuint64_t gdims[5];
  cuuint64_t gstr[5];
  cuuint32_t gbox[5];
  cuuint32_t estr[5];
  for (int i = 0; i < rank; ++i) {
    gdims[i] = dims[i];
    gbox[i] = box[i];
    estr[i] = 1;
    MB_REQUIRE(box[i] >= 1 && box[i] <= 256, "make_tensor_map: box[%d]=%u out of range", i, box[i]);
  }
  for (int i = 0; i + 1 < rank; ++i) gstr[i] = strides_bytes[i];
  CUresult r = fn(out,
                  dtype == kTmaBF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16
                                    : CU_TENSOR_MAP_DATA_TYPE_FLOAT32,
                  static_cast<cuuint32_t>(rank), const_cast<void*>(base), gdims, gstr, gbox, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE,
                  swizzle_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  MB_REQUIRE(r == CUDA_SUCCESS,
             "cuTensorMapEncodeTiled failed with CUresult %d (rank %d, dims %llu,%llu box %u,%u)",
             (int)r, rank, (unsigned long long)dims[0], (unsigned long long)(rank > 1 ? dims[1] : 0),
             box[0], rank > 1 ? box[1] : 0);
  {
    std::lock_guard<std::mutex> lk(g_map_mu);
    if (g_map_cache.size() > 4096) g_map_cache.clear();
    g_map_cache.emplace(key, *out);
  }
  return 0;
}

}  // namespace mb200

extern "C" {

const char* mb_last_error(void) { return mb200::g_last_error; }

int mb_version(void) { return MB_VERSION; }

int mb_sm_count(void) { return mb200::sm_count(); }

int mb_set_sm_reserve(int n) {
  const int prev = mb200::g_sm_reserve;
  mb200::g_sm_reserve = n < 0 ? 0 : n;
  return prev;
}

int mb_set_pdl(int on) {
  const int prev = mb200::pdl_mode();
  mb200::g_pdl = on & 3;
  return prev;
}

void mb_clear_tensor_map_cache(void) {
  std::lock_guard<std::mutex> lk(mb200::g_map_mu);
  mb200::g_map_cache.clear();
}

}  // extern "C"
