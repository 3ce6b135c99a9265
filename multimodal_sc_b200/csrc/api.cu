// api.cu — library plumbing: error string, device check, launch counter, TMA descriptor
// encoding through the driver entry point (no link-time dependency on libcuda).
#include <atomic>
#include <cstdlib>
#include <cstring>

#include "common.cuh"

static thread_local char g_err[512] = "";
static std::atomic<long long> g_launches{0};

void sc_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

void sc_count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

// programmatic dependent launch: -1 = not decided yet (SC_PDL=1 turns it on; default off: measured 1-3 % slower on
// the PPO step, profiles/README.md)
static std::atomic<int> g_pdl{-1};
int sc_pdl_enabled() {
  int v = g_pdl.load(std::memory_order_relaxed);
  if (v < 0) {
    const char* e = getenv("SC_PDL");
    v = (e && e[0] == '1') ? 1 : 0;
    g_pdl.store(v, std::memory_order_relaxed);
  }
  return v;
}
void sc_pdl_set(int v) { g_pdl.store(v ? 1 : 0, std::memory_order_relaxed); }

// traversal direction of the streaming kernels.  sc_debug_set("alternate", 1): consecutive launches walk their rows
// in opposite directions, so a consumer starts with the rows its producer touched last (still in the 126 MB L2).
static std::atomic<int> g_alt{0};
static std::atomic<int> g_dir{0};
void sc_alternate_set(int v) { g_alt.store(v ? 1 : 0); g_dir.store(0); }
int sc_take_direction() {
  if (!g_alt.load(std::memory_order_relaxed)) return 0;
  return g_dir.fetch_xor(1, std::memory_order_relaxed) & 1;
}

struct DevInfo {
  int checked = 0;
  int major = 0;
  int sms = 0;
};
static DevInfo g_dev[64];

static DevInfo* dev_info() {
  int d = 0;
  if (cudaGetDevice(&d) != cudaSuccess || d < 0 || d >= 64) return nullptr;
  DevInfo* di = &g_dev[d];
  if (!di->checked) {
    cudaDeviceProp p;
    if (cudaGetDeviceProperties(&p, d) != cudaSuccess) return nullptr;
    di->major = p.major;
    di->sms = p.multiProcessorCount;
    di->checked = 1;
  }
  return di;
}

int sc_check_device() {
  DevInfo* di = dev_info();
  if (!di) {
    sc_set_error("no CUDA device available: %s", cudaGetErrorString(cudaGetLastError()));
    return SC_E_ARCH;
  }
  if (di->major != 10) {
    sc_set_error("libsemcom_sm100 requires an sm_100 (B200) device, found sm_%d0; there is no fallback",
                 di->major);
    return SC_E_ARCH;
  }
  return SC_OK;
}

int sc_device_index() {
  int d = 0;
  if (cudaGetDevice(&d) != cudaSuccess || d < 0 || d >= 64) return -1;
  return d;
}

int sc_num_sms() {
  DevInfo* di = dev_info();
  return di ? di->sms : 148;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                  CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                  CUtensorMapFloatOOBfill);
static EncodeTiledFn g_encode = nullptr;

static int load_encode() {
  if (g_encode) return SC_OK;
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
  if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || !fn) {
    sc_set_error("cuTensorMapEncodeTiled not available from the driver (%s)", cudaGetErrorString(e));
    return SC_E_DRIVER;
  }
  g_encode = reinterpret_cast<EncodeTiledFn>(fn);
  return SC_OK;
}

// Descriptor cache.  cuTensorMapEncodeTiled is a pure function of its arguments (it fills 128 bytes, registers nothing
// with the driver), yet costs ~1 us per call and a GEMM needs up to five: an eager small-batch forward spent more time
// encoding descriptors than its kernels ran (cfg2, B = 256: 1.21 ms eager vs 0.70 ms as a CUDA graph).  Direct-mapped,
// thread-local (no locks, re-entrant), keyed by every encode argument.
struct TmapKey {
  const void* base;
  uint64_t inner, outer, stride;
  uint32_t box_inner, box_outer;
  bool operator==(const TmapKey& o) const {
    return base == o.base && inner == o.inner && outer == o.outer && stride == o.stride && box_inner == o.box_inner &&
           box_outer == o.box_outer;
  }
};
struct TmapEntry {
  TmapKey key;
  CUtensorMap tm;
  bool valid;
};
constexpr int kTmapCacheSize = 4096;
static thread_local TmapEntry* g_tmap_cache = nullptr;
static std::atomic<long long> g_tmap_hits{0}, g_tmap_misses{0};
static std::atomic<int> g_tmap_cache_on{1};
void sc_tmap_cache_set(int on) { g_tmap_cache_on.store(on ? 1 : 0); }
long long sc_tmap_cache_stat(int which) { return which ? g_tmap_misses.load() : g_tmap_hits.load(); }

static inline uint32_t tmap_hash(const TmapKey& k) {
  uint64_t h = reinterpret_cast<uint64_t>(k.base) * 0x9E3779B97F4A7C15ull;
  h ^= (k.inner + 0x7F4A7C15ull) * 0xBF58476D1CE4E5B9ull;
  h ^= (k.outer + 0x1CE4E5B9ull) * 0x94D049BB133111EBull;
  h ^= (k.stride << 7) ^ (static_cast<uint64_t>(k.box_inner) << 40) ^ (static_cast<uint64_t>(k.box_outer) << 52);
  h ^= h >> 29;
  return static_cast<uint32_t>(h) & (kTmapCacheSize - 1);
}

// 2-D bf16 tensor, inner dimension contiguous, SWIZZLE_128B boxes (box_inner * 2 bytes == 128).
int sc_encode_tmap_2d_bf16(CUtensorMap* tm, const void* base, uint64_t inner, uint64_t outer,
                           uint64_t row_stride_bytes, uint32_t box_inner, uint32_t box_outer) {
  int rc = load_encode();
  if (rc != SC_OK) return rc;
  const TmapKey key{base, inner, outer, row_stride_bytes, box_inner, box_outer};
  TmapEntry* slot = nullptr;
  if (g_tmap_cache_on.load(std::memory_order_relaxed)) {
    if (!g_tmap_cache) g_tmap_cache = static_cast<TmapEntry*>(calloc(kTmapCacheSize, sizeof(TmapEntry)));
    if (g_tmap_cache) {
      slot = &g_tmap_cache[tmap_hash(key)];
      if (slot->valid && slot->key == key) {
        memcpy(tm, &slot->tm, sizeof(CUtensorMap));
        g_tmap_hits.fetch_add(1, std::memory_order_relaxed);
        return SC_OK;
      }
    }
  }
  g_tmap_misses.fetch_add(1, std::memory_order_relaxed);
  cuuint64_t dims[2] = {inner, outer};
  cuuint64_t strides[1] = {row_stride_bytes};
  cuuint32_t box[2] = {box_inner, box_outer};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = g_encode(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides,
                        box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                        CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    sc_set_error("cuTensorMapEncodeTiled failed (CUresult %d): base=%p inner=%llu outer=%llu stride=%llu box=%ux%u",
                 (int)r, base, (unsigned long long)inner, (unsigned long long)outer,
                 (unsigned long long)row_stride_bytes, box_inner, box_outer);
    return SC_E_DRIVER;
  }
  if (slot) {
    slot->key = key;
    memcpy(&slot->tm, tm, sizeof(CUtensorMap));
    slot->valid = true;
  }
  return SC_OK;
}

extern "C" {

int sc_version(void) { return SC_ABI_VERSION; }
const char* sc_last_error(void) { return g_err; }
int sc_device_ok(void) { return sc_check_device(); }
int64_t sc_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

}  // extern "C"
