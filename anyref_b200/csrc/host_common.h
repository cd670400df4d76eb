// Host-side helpers shared by the C-ABI entry points: error reporting and TMA tensor-map encoding.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace samhost {

// Records a message retrievable through sam_last_error(); returns `code` (non-zero).
int set_error(int code, const char* fmt, ...);
const char* last_error();

#define SAM_CHECK_CUDA(expr)                                                                              \
  do {                                                                                                    \
    cudaError_t _e = (expr);                                                                              \
    if (_e != cudaSuccess)                                                                                \
      return samhost::set_error(2, "%s failed at %s:%d: %s", #expr, __FILE__, __LINE__,                   \
                                cudaGetErrorString(_e));                                                  \
  } while (0)

#define SAM_REQUIRE(cond, ...)                                     \
  do {                                                             \
    if (!(cond)) return samhost::set_error(1, __VA_ARGS__);        \
  } while (0)

// Row-major 2-D tensor [outer, inner] of 2- or 4-byte elements; box = [box_outer, box_inner].
// swizzle: 0 none, 1 32B, 2 64B, 3 128B (CUtensorMapSwizzle values). Returns 0 on success.
int encode_tmap_2d(CUtensorMap* out, int elem_bytes, int is_bf16, const void* base, uint64_t inner, uint64_t outer,
                   uint64_t outer_stride_bytes, uint32_t box_inner, uint32_t box_outer, int swizzle);

// Generic up-to-5-D encode (dims/strides innermost first; strides[0] is implied by the element size).
int encode_tmap_nd(CUtensorMap* out, int elem_bytes, int is_bf16, const void* base, int rank, const uint64_t* dims,
                   const uint64_t* strides_bytes, const uint32_t* box, int swizzle);

// SM count of the CURRENT device (cached per device).
int sm_count();
// Ordinal of the current device clamped to [0, 63] (per-device caches are small fixed arrays).
int device_slot();

// Per-device one-time initialisation (cudaFuncSetAttribute opt-ins, occupancy queries): function attributes belong to
// the device that is current when they are set, so a process-wide `static bool` would leave every other GPU of the
// process without the opt-in.   if (once.need()) { ...; once.done(); }
struct PerDeviceOnce {
  bool need() const;
  void done();
  unsigned long long mask_ = 0;
};

// ---------------------------------------------------------------------------------------------------------------
// Launch accounting + optional per-kernel-class timing (bench.py: `gpu_launches` and the roofline leg).
// Every kernel launch site opens a LaunchScope.  It always bumps the launch counter; when profiling is enabled it
// also brackets the launch with a CUDA event pair on the launch stream.  profile_collect() synchronises the recorded
// events and folds them into per-class totals (time, launches, algorithmic FLOPs and bytes as stated by the site).
// ---------------------------------------------------------------------------------------------------------------
enum KernelClass {
  KC_GEMM = 0, KC_ATTN_WINDOW, KC_ATTN_GLOBAL, KC_LAYERNORM, KC_LAYOUT, KC_DECODER, KC_POSTPROCESS, KC_COUNT
};
struct LaunchScope {
  LaunchScope(int cls, cudaStream_t stream, double flops = 0.0, double bytes = 0.0, int launches = 1);
  ~LaunchScope();
  int slot_;
  cudaStream_t stream_;
};
// While alive on this thread, every LaunchScope is accounted to class `cls` (the mask decoder's tensor-core GEMMs are
// launched through samk_gemm but belong to the decoder's time, not to the encoder GEMM roofline).
struct ClassOverride {
  explicit ClassOverride(int cls);
  ~ClassOverride();
  int prev_;
};
long long launch_count();
void profile_enable(int on);
// Folds all recorded event pairs into the totals; returns 0 or a CUDA error code.
int profile_collect();
void profile_reset();
// totals since the last reset for class `cls`
void profile_get(int cls, double* ms, long long* launches, double* flops, double* bytes);

}  // namespace samhost
