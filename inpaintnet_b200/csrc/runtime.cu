#include "runtime.h"

#include <stdarg.h>

#include <algorithm>
#include <atomic>
#include <mutex>
#include <map>
#include <string>
#include <unordered_map>
#include <vector>

namespace ipn {

static thread_local char g_err[1024] = "";
static std::atomic<long long> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

int ensure_device() {
  static std::mutex mu;
  static int cached[64];
  static bool init = false;
  int dev = -1;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) {
    set_error("no CUDA device available: this library has no CPU fallback");
    cudaGetLastError();
    return IPN_ERR_ARCH;
  }
  std::lock_guard<std::mutex> lk(mu);
  if (!init) { for (int i = 0; i < 64; ++i) cached[i] = -1; init = true; }
  if (cached[dev] < 0) {
    int major = 0, minor = 0;
    cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
    cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev);
    cached[dev] = (major == 10) ? 1 : 0;
    if (!cached[dev]) set_error("device %d is sm_%d%d; this library is built for sm_100a only", dev, major, minor);
  }
  if (!cached[dev]) {
    set_error("device %d is not sm_100; this library is built for sm_100a only", dev);
    return IPN_ERR_ARCH;
  }
  return IPN_OK;
}

// ---------------------------------------------------------------------------------------------
// per-kernel-class event timing
// ---------------------------------------------------------------------------------------------
struct ProfRec {
  const char* tag;
  double flops, bytes;
  cudaEvent_t e0, e1;
};
static bool g_prof_on = false;
static std::vector<ProfRec> g_prof;
static std::vector<cudaEvent_t> g_event_pool;
static std::mutex g_prof_mu;

static cudaEvent_t prof_event() {
  if (!g_event_pool.empty()) {
    cudaEvent_t e = g_event_pool.back();
    g_event_pool.pop_back();
    return e;
  }
  cudaEvent_t e;
  cudaEventCreate(&e);
  return e;
}

void prof_begin(const char* tag, double flops, double bytes, cudaStream_t stream) {
  if (!g_prof_on) return;
  std::lock_guard<std::mutex> lk(g_prof_mu);
  ProfRec r{tag, flops, bytes, prof_event(), prof_event()};
  cudaEventRecord(r.e0, stream);
  g_prof.push_back(r);
}

void prof_end(cudaStream_t stream) {
  if (!g_prof_on) return;
  std::lock_guard<std::mutex> lk(g_prof_mu);
  if (!g_prof.empty()) cudaEventRecord(g_prof.back().e1, stream);
}

// ---------------------------------------------------------------------------------------------
// TMA descriptor cache
// ---------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
    else
      cudaGetLastError();
  }
  return fn;
}

struct TmKey {
  const void* ptr;
  unsigned long long inner, outer;
  long long ld;
  unsigned box_outer;
  bool operator==(const TmKey& o) const {
    return ptr == o.ptr && inner == o.inner && outer == o.outer && ld == o.ld && box_outer == o.box_outer;
  }
};
struct TmHash {
  size_t operator()(const TmKey& k) const {
    size_t h = std::hash<const void*>()(k.ptr);
    auto mix = [&](unsigned long long v) { h ^= std::hash<unsigned long long>()(v) + 0x9e3779b97f4a7c15ull + (h << 6) + (h >> 2); };
    mix(k.inner); mix(k.outer); mix((unsigned long long)k.ld); mix(k.box_outer);
    return h;
  }
};

int get_tensor_map(CUtensorMap* out, const void* ptr, unsigned long long inner, unsigned long long outer,
                   long long ld, unsigned box_outer) {
  static std::mutex mu;
  static std::unordered_map<TmKey, CUtensorMap, TmHash> cache;
  IPN_REQUIRE(ptr != nullptr, IPN_ERR_ARG, "tensor map: null pointer");
  IPN_REQUIRE(reinterpret_cast<uintptr_t>(ptr) % 16 == 0, IPN_ERR_ALIGN, "tensor map: pointer %p not 16B aligned", ptr);
  IPN_REQUIRE(ld % 8 == 0, IPN_ERR_ALIGN, "tensor map: leading dimension %lld not a multiple of 8 bf16", ld);
  IPN_REQUIRE(inner > 0 && outer > 0 && box_outer > 0 && box_outer <= 256, IPN_ERR_ARG, "tensor map: bad dims");
  TmKey key{ptr, inner, outer, ld, box_outer};
  std::lock_guard<std::mutex> lk(mu);
  auto it = cache.find(key);
  if (it != cache.end()) { *out = it->second; return IPN_OK; }
  EncodeTiledFn fn = get_encode_fn();
  IPN_REQUIRE(fn != nullptr, IPN_ERR_CUDA, "cuTensorMapEncodeTiled not available from the driver");
  cuuint64_t gdim[2] = {inner, outer};
  cuuint64_t gstr[1] = {(cuuint64_t)ld * 2};
  cuuint32_t box[2] = {64, box_outer};
  cuuint32_t estr[2] = {1, 1};
  CUtensorMap tm;
  CUresult r = fn(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), gdim, gstr, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  IPN_REQUIRE(r == CUDA_SUCCESS, IPN_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d): ptr=%p inner=%llu outer=%llu ld=%lld box=%u",
              (int)r, ptr, inner, outer, ld, box_outer);
  if (cache.size() > 8192) cache.clear();
  cache.emplace(key, tm);
  *out = tm;
  return IPN_OK;
}

int get_tensor_map_3d(CUtensorMap* out, const void* ptr, unsigned long long d0, unsigned long long d1, unsigned long long d2,
                      long long ld1, long long ld2, unsigned box1, unsigned box2) {
  IPN_REQUIRE(ptr != nullptr && reinterpret_cast<uintptr_t>(ptr) % 16 == 0, IPN_ERR_ALIGN, "tensor map 3d: bad pointer %p", ptr);
  IPN_REQUIRE(ld1 % 8 == 0 && ld2 % 8 == 0, IPN_ERR_ALIGN, "tensor map 3d: strides must be multiples of 8 bf16");
  IPN_REQUIRE(d0 > 0 && d1 > 0 && d2 > 0 && box1 > 0 && box1 <= 256 && box2 > 0 && box2 <= 256, IPN_ERR_ARG, "tensor map 3d: bad dims");
  EncodeTiledFn fn = get_encode_fn();
  IPN_REQUIRE(fn != nullptr, IPN_ERR_CUDA, "cuTensorMapEncodeTiled not available from the driver");
  cuuint64_t gdim[3] = {d0, d1, d2};
  cuuint64_t gstr[2] = {(cuuint64_t)ld1 * 2, (cuuint64_t)ld2 * 2};
  cuuint32_t box[3] = {64, box1, box2};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(ptr), gdim, gstr, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  IPN_REQUIRE(r == CUDA_SUCCESS, IPN_ERR_CUDA, "cuTensorMapEncodeTiled(3d) failed (%d)", (int)r);
  return IPN_OK;
}

}  // namespace ipn

extern "C" {
const char* ipn_last_error(void) { return ipn::g_err; }
int ipn_abi_version(void) { return IPN_ABI_VERSION; }
int ipn_struct_sizes(int* out_host, int n) {
  const int sizes[] = {(int)sizeof(IpnRowMap),      (int)sizeof(IpnGemmSeg),  (int)sizeof(IpnGemm),
                       (int)sizeof(IpnGruInproj),   (int)sizeof(IpnLstmInproj),  (int)sizeof(IpnGruDir),      (int)sizeof(IpnGruLayer), (int)sizeof(IpnGruBwdDir),
                       (int)sizeof(IpnGruLayerBwd), (int)sizeof(IpnLstmLayer), (int)sizeof(IpnLstmLayerBwd),
                       (int)sizeof(IpnCeKl),        (int)sizeof(IpnPackItem), (int)sizeof(IpnTickDecode)};
  const int m = (int)(sizeof(sizes) / sizeof(sizes[0]));
  for (int i = 0; i < n && i < m; ++i) out_host[i] = sizes[i];
  return m;
}
long long ipn_launch_count(void) { return ipn::g_launches.load(); }

void ipn_prof_enable(int on) {
  std::lock_guard<std::mutex> lk(ipn::g_prof_mu);
  ipn::g_prof_on = on != 0;
}

// Synchronises the device, aggregates the recorded launches per tag and writes lines
//   tag<TAB>launches<TAB>total_ms<TAB>flops<TAB>bytes
// into buf (NUL terminated, truncated to cap). Clears the records. Returns the number of tags.
int ipn_prof_report(char* buf_host, int cap) {
  using namespace ipn;
  cudaDeviceSynchronize();
  std::lock_guard<std::mutex> lk(g_prof_mu);
  struct Agg { long long n = 0; double ms = 0, flops = 0, bytes = 0; };
  std::map<std::string, Agg> agg;
  for (auto& r : g_prof) {
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, r.e0, r.e1) != cudaSuccess) { cudaGetLastError(); ms = 0.f; }
    Agg& a = agg[r.tag];
    a.n += 1; a.ms += ms; a.flops += r.flops; a.bytes += r.bytes;
    g_event_pool.push_back(r.e0);
    g_event_pool.push_back(r.e1);
  }
  g_prof.clear();
  std::string out;
  char line[512];
  for (auto& kv : agg) {
    snprintf(line, sizeof(line), "%s\t%lld\t%.6f\t%.6e\t%.6e\n", kv.first.c_str(), kv.second.n, kv.second.ms,
             kv.second.flops, kv.second.bytes);
    out += line;
  }
  if (buf_host && cap > 0) {
    const int n = (int)std::min<size_t>(out.size(), (size_t)cap - 1);
    memcpy(buf_host, out.data(), n);
    buf_host[n] = 0;
  }
  return (int)agg.size();
}
int ipn_device_check(int dev, int* sm_count_host) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess || dev >= n) {
    cudaGetLastError();
    ipn::set_error("no CUDA device %d (count %d): this library has no CPU fallback", dev, n);
    return IPN_ERR_ARCH;
  }
  int major = 0, sms = 0;
  cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  if (sm_count_host) *sm_count_host = sms;
  if (major != 10) {
    ipn::set_error("device %d is not sm_100", dev);
    return IPN_ERR_ARCH;
  }
  return IPN_OK;
}
}
