// Host-side TMA descriptor (CUtensorMap) construction with a small cache, device queries and the
// last-error slot for the C ABI.  cuTensorMapEncodeTiled is fetched through
// cudaGetDriverEntryPoint so the library links against cudart only.
#include <cstdlib>
#include <mutex>
#include <string>
#include <unordered_map>
#include <vector>

#include "common.cuh"

namespace lako {

static thread_local std::string g_last_error;

void set_last_error(const char* msg, const char* file, int line) {
  char buf[1024];
  snprintf(buf, sizeof(buf), "%s (%s:%d)", msg, file, line);
  g_last_error = buf;
}
const char* last_error_cstr() { return g_last_error.c_str(); }

static int g_pdl = -1;
bool pdl_enabled() {
  if (g_pdl < 0) {
    const char* e = getenv("LAKO_PDL");
    g_pdl = (e && e[0] == '0') ? 0 : 1;
  }
  return g_pdl != 0;
}
static unsigned long long g_launches = 0;
void count_launch() { ++g_launches; }
void count_launches(long long n) { g_launches += n; }
unsigned long long launch_count_value() { return g_launches; }

// ---------------------------------------------------------------- profiling scopes
struct ProfRec {
  int cat;
  double work;
  cudaEvent_t e0, e1;
  cudaStream_t stream;
};
static bool g_prof_on = false;
bool prof_enabled() { return g_prof_on; }
static std::vector<ProfRec*> g_prof_recs;
static std::vector<ProfRec*> g_prof_pool;

ProfScope::ProfScope(int c, double work, cudaStream_t s) : cat(c), rec(nullptr) {
  if (!g_prof_on) return;
  ProfRec* r;
  if (!g_prof_pool.empty()) {
    r = g_prof_pool.back();
    g_prof_pool.pop_back();
  } else {
    r = new ProfRec();
    cudaEventCreate(&r->e0);
    cudaEventCreate(&r->e1);
  }
  r->cat = c;
  r->work = work;
  r->stream = s;
  cudaEventRecord(r->e0, s);
  rec = r;
}
ProfScope::~ProfScope() {
  if (!rec) return;
  ProfRec* r = static_cast<ProfRec*>(rec);
  cudaEventRecord(r->e1, r->stream);
  g_prof_recs.push_back(r);
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres);
    if (e == cudaSuccess && qres == cudaDriverEntryPointSuccess) fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}

struct MapKey {
  uint64_t v[10];
  bool operator==(const MapKey& o) const {
    for (int i = 0; i < 10; ++i)
      if (v[i] != o.v[i]) return false;
    return true;
  }
};
struct MapKeyHash {
  size_t operator()(const MapKey& k) const {
    uint64_t h = 1469598103934665603ull;
    for (int i = 0; i < 10; ++i) {
      h ^= k.v[i];
      h *= 1099511628211ull;
    }
    return static_cast<size_t>(h);
  }
};

static std::mutex g_map_mu;
static std::unordered_map<MapKey, CUtensorMap, MapKeyHash> g_maps;

static int encode(CUtensorMap* out, const MapKey& key, int rank, const void* base, const cuuint64_t* dims,
                  const cuuint64_t* strides_bytes, const cuuint32_t* box,
                  CUtensorMapDataType dtype = CU_TENSOR_MAP_DATA_TYPE_BFLOAT16) {
  {
    std::lock_guard<std::mutex> lk(g_map_mu);
    auto it = g_maps.find(key);
    if (it != g_maps.end()) {
      *out = it->second;
      return LAKO_OK;
    }
  }
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) {
    set_last_error("cuTensorMapEncodeTiled unavailable (driver too old?)", __FILE__, __LINE__);
    return LAKO_ERR_DRIVER;
  }
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = fn(out, dtype, rank, const_cast<void*>(base), dims, strides_bytes, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    char msg[256];
    snprintf(msg, sizeof(msg),
             "cuTensorMapEncodeTiled failed: CUresult %d (rank %d base %p dims %llu,%llu stride %llu box %u,%u)", (int)r,
             rank, base, (unsigned long long)dims[0], (unsigned long long)dims[1],
             (unsigned long long)strides_bytes[0], box[0], box[1]);
    set_last_error(msg, __FILE__, __LINE__);
    return LAKO_ERR_DRIVER;
  }
  std::lock_guard<std::mutex> lk(g_map_mu);
  if (g_maps.size() > 65536) g_maps.clear();
  g_maps.emplace(key, *out);
  return LAKO_OK;
}

int get_tensor_map_2d(CUtensorMap* out, const void* base, uint64_t inner, uint64_t outer, uint64_t pitch_elems,
                      uint32_t box_inner, uint32_t box_outer) {
  LAKO_REQUIRE((reinterpret_cast<uintptr_t>(base) & 15) == 0, LAKO_ERR_BAD_ARG);
  LAKO_REQUIRE((pitch_elems * 2) % 16 == 0, LAKO_ERR_BAD_ARG);
  LAKO_REQUIRE(box_inner * 2 <= 128 && box_outer <= 256 && box_outer >= 1, LAKO_ERR_BAD_ARG);
  MapKey key{{2, reinterpret_cast<uint64_t>(base), inner, outer, pitch_elems, box_inner, box_outer, 0, 0, 0}};
  cuuint64_t dims[2] = {inner, outer};
  cuuint64_t strides[1] = {pitch_elems * 2};
  cuuint32_t box[2] = {box_inner, box_outer};
  return encode(out, key, 2, base, dims, strides, box);
}

// fp32 [outer][pitch_elems] matrix, box = box_inner (<= 32 floats = 128 B) x box_outer rows, 128B swizzle: the
// destination of the split-K / gradient-accumulation epilogue's TMA reduce-add stores
int get_tensor_map_2d_f32(CUtensorMap* out, const void* base, uint64_t inner, uint64_t outer, uint64_t pitch_elems,
                          uint32_t box_inner, uint32_t box_outer) {
  LAKO_REQUIRE((reinterpret_cast<uintptr_t>(base) & 15) == 0, LAKO_ERR_BAD_ARG);
  LAKO_REQUIRE((pitch_elems * 4) % 16 == 0, LAKO_ERR_BAD_ARG);
  LAKO_REQUIRE(box_inner * 4 <= 128 && box_outer <= 256 && box_outer >= 1, LAKO_ERR_BAD_ARG);
  MapKey key{{4, reinterpret_cast<uint64_t>(base), inner, outer, pitch_elems, box_inner, box_outer, 0, 0, 0}};
  cuuint64_t dims[2] = {inner, outer};
  cuuint64_t strides[1] = {pitch_elems * 4};
  cuuint32_t box[2] = {box_inner, box_outer};
  return encode(out, key, 2, base, dims, strides, box, CU_TENSOR_MAP_DATA_TYPE_FLOAT32);
}

int get_tensor_map_3d(CUtensorMap* out, const void* base, uint64_t inner, uint64_t d1, uint64_t d2,
                      uint64_t stride1_elems, uint64_t stride2_elems, uint32_t box_inner, uint32_t box1,
                      uint32_t box2) {
  LAKO_REQUIRE((reinterpret_cast<uintptr_t>(base) & 15) == 0, LAKO_ERR_BAD_ARG);
  LAKO_REQUIRE((stride1_elems * 2) % 16 == 0 && (stride2_elems * 2) % 16 == 0, LAKO_ERR_BAD_ARG);
  LAKO_REQUIRE(box_inner * 2 <= 128 && box1 <= 256 && box2 <= 256, LAKO_ERR_BAD_ARG);
  MapKey key{{3, reinterpret_cast<uint64_t>(base), inner, d1, d2, stride1_elems, stride2_elems, box_inner, box1, box2}};
  cuuint64_t dims[3] = {inner, d1, d2};
  cuuint64_t strides[2] = {stride1_elems * 2, stride2_elems * 2};
  cuuint32_t box[3] = {box_inner, box1, box2};
  return encode(out, key, 3, base, dims, strides, box);
}

int device_sm_count() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

int check_arch() {
  static int ok = -1;
  if (ok < 0) {
    int dev = 0, major = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return LAKO_ERR_ARCH;
    cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
    ok = (major == 10) ? 1 : 0;
  }
  if (!ok) {
    set_last_error("lako_b200 kernels are sm_100a only; no fallback path exists", __FILE__, __LINE__);
    return LAKO_ERR_ARCH;
  }
  return LAKO_OK;
}

}  // namespace lako

extern "C" void lako_set_pdl(int on) { lako::g_pdl = on ? 1 : 0; }
extern "C" unsigned long long lako_launch_count() { return lako::g_launches; }
extern "C" void lako_prof_enable(int on) { lako::g_prof_on = on != 0; }
// Synchronises, then fills ms[cat], work[cat], count[cat] (arrays of LAKO_PROF_NCAT) and clears the records.
extern "C" int lako_prof_collect(double* ms, double* work, long long* count) {
  cudaDeviceSynchronize();
  for (int i = 0; i < lako::PROF_NCAT; ++i) {
    ms[i] = 0;
    work[i] = 0;
    count[i] = 0;
  }
  for (lako::ProfRec* r : lako::g_prof_recs) {
    float t = 0.f;
    cudaEventElapsedTime(&t, r->e0, r->e1);
    ms[r->cat] += t;
    work[r->cat] += r->work;
    count[r->cat] += 1;
    lako::g_prof_pool.push_back(r);
  }
  lako::g_prof_recs.clear();
  return lako::PROF_NCAT;
}

extern "C" const char* lako_last_error() { return lako::last_error_cstr(); }
extern "C" int lako_check_device() { return lako::check_arch(); }
