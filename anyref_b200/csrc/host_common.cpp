#include "host_common.h"

#include <stdarg.h>
#include <stdio.h>
#include <string.h>

#include <atomic>
#include <mutex>
#include <vector>

namespace samhost {

static thread_local char g_err[1024] = "";

int set_error(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}
const char* last_error() { return g_err; }

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// The driver entry point is resolved through the runtime so the library has no link-time libcuda dependency.
static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (fn) return fn;
  void* p = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
  if (e != cudaSuccess || q != cudaDriverEntryPointSuccess || !p) return nullptr;
  fn = reinterpret_cast<EncodeTiledFn>(p);
  return fn;
}

int encode_tmap_nd(CUtensorMap* out, int elem_bytes, int is_bf16, const void* base, int rank, const uint64_t* dims,
                   const uint64_t* strides_bytes, const uint32_t* box, int swizzle) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) return set_error(3, "cuTensorMapEncodeTiled entry point unavailable (no CUDA driver?)");
  CUtensorMapDataType dt;
  if (elem_bytes == 2)
    dt = is_bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16;
  else if (elem_bytes == 4)
    dt = CU_TENSOR_MAP_DATA_TYPE_FLOAT32;
  else
    return set_error(1, "encode_tmap: unsupported element size %d", elem_bytes);
  cuuint64_t gdim[5];
  cuuint64_t gstr[4];
  cuuint32_t bx[5];
  cuuint32_t es[5];
  for (int i = 0; i < rank; ++i) {
    gdim[i] = dims[i];
    bx[i] = box[i];
    es[i] = 1;
    if (i > 0) gstr[i - 1] = strides_bytes[i];
  }
  CUresult r = fn(out, dt, static_cast<cuuint32_t>(rank), const_cast<void*>(base), gdim, gstr, bx, es,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, static_cast<CUtensorMapSwizzle>(swizzle),
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return set_error(3, "cuTensorMapEncodeTiled failed: CUresult=%d (rank=%d dims0=%llu box0=%u swz=%d)", (int)r,
                     rank, (unsigned long long)dims[0], box[0], swizzle);
  return 0;
}

int encode_tmap_2d(CUtensorMap* out, int elem_bytes, int is_bf16, const void* base, uint64_t inner, uint64_t outer,
                   uint64_t outer_stride_bytes, uint32_t box_inner, uint32_t box_outer, int swizzle) {
  uint64_t dims[2] = {inner, outer};
  uint64_t strides[2] = {static_cast<uint64_t>(elem_bytes), outer_stride_bytes};
  uint32_t box[2] = {box_inner, box_outer};
  return encode_tmap_nd(out, elem_bytes, is_bf16, base, 2, dims, strides, box, swizzle);
}

int device_slot() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) {
    (void)cudaGetLastError();
    return 0;
  }
  return dev < 0 ? 0 : (dev > 63 ? 63 : dev);
}

int sm_count() {
  static std::atomic<int> cache[64];
  const int slot = device_slot();
  int n = cache[slot].load(std::memory_order_relaxed);
  if (n) return n;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 148;
  if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
  cache[slot].store(n, std::memory_order_relaxed);
  return n;
}

namespace {
std::mutex g_once_mu;
}
bool PerDeviceOnce::need() const {
  std::lock_guard<std::mutex> lk(g_once_mu);
  return ((mask_ >> device_slot()) & 1ull) == 0;
}
void PerDeviceOnce::done() {
  std::lock_guard<std::mutex> lk(g_once_mu);
  mask_ |= 1ull << device_slot();
}

// ------------------------------------------------------------------------------------------------- launch accounting
namespace {
std::atomic<long long> g_launches{0};
std::atomic<int> g_profile_on{0};
struct Rec {
  cudaEvent_t a, b;
  int cls;
  double flops, bytes;
  int launches;
};
std::mutex g_mu;
std::vector<Rec> g_recs;      // event pool; entries [0, g_used) are live
size_t g_used = 0;
struct Tot {
  double ms = 0, flops = 0, bytes = 0;
  long long launches = 0;
} g_tot[KC_COUNT];
}  // namespace

static thread_local int g_class_override = -1;
ClassOverride::ClassOverride(int cls) : prev_(g_class_override) { g_class_override = cls; }
ClassOverride::~ClassOverride() { g_class_override = prev_; }

LaunchScope::LaunchScope(int cls, cudaStream_t stream, double flops, double bytes, int launches)
    : slot_(-1), stream_(stream) {
  g_launches.fetch_add(launches, std::memory_order_relaxed);
  if (!g_profile_on.load(std::memory_order_relaxed)) return;
  if (g_class_override >= 0) cls = g_class_override;
  std::lock_guard<std::mutex> lk(g_mu);
  if (g_used == g_recs.size()) {
    Rec r{};
    if (cudaEventCreate(&r.a) != cudaSuccess || cudaEventCreate(&r.b) != cudaSuccess) return;
    g_recs.push_back(r);
  }
  Rec& r = g_recs[g_used];
  r.cls = cls;
  r.flops = flops;
  r.bytes = bytes;
  r.launches = launches;
  slot_ = static_cast<int>(g_used++);
  cudaEventRecord(r.a, stream);
}
LaunchScope::~LaunchScope() {
  if (slot_ < 0) return;
  std::lock_guard<std::mutex> lk(g_mu);
  cudaEventRecord(g_recs[slot_].b, stream_);
}
long long launch_count() { return g_launches.load(); }
void profile_enable(int on) { g_profile_on.store(on ? 1 : 0); }
int profile_collect() {
  std::lock_guard<std::mutex> lk(g_mu);
  for (size_t i = 0; i < g_used; ++i) {
    Rec& r = g_recs[i];
    cudaError_t e = cudaEventSynchronize(r.b);
    if (e != cudaSuccess) return set_error(2, "profile_collect: %s", cudaGetErrorString(e));
    float ms = 0.f;
    e = cudaEventElapsedTime(&ms, r.a, r.b);
    if (e != cudaSuccess) return set_error(2, "profile_collect: %s", cudaGetErrorString(e));
    Tot& t = g_tot[r.cls];
    t.ms += ms;
    t.flops += r.flops;
    t.bytes += r.bytes;
    t.launches += r.launches;
  }
  g_used = 0;
  return 0;
}
void profile_reset() {
  std::lock_guard<std::mutex> lk(g_mu);
  g_used = 0;
  for (auto& t : g_tot) t = Tot();
}
void profile_get(int cls, double* ms, long long* launches, double* flops, double* bytes) {
  std::lock_guard<std::mutex> lk(g_mu);
  if (cls < 0 || cls >= KC_COUNT) return;
  *ms = g_tot[cls].ms;
  *launches = g_tot[cls].launches;
  *flops = g_tot[cls].flops;
  *bytes = g_tot[cls].bytes;
}

}  // namespace samhost
