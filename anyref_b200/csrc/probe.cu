// Test-only kernel: one 128 x N x K tcgen05 MMA tile whose shared-memory operands are written by ordinary
// threads (generic proxy) in a chosen canonical UMMA layout.  Used by tests/test_umma_layouts.py to pin the
// descriptor conventions (swizzle XOR pattern, LBO/SBO meaning, MN-major operands) that the attention
// kernels rely on.  Not part of the product path.
#include "host_common.h"
#include "kernels.h"
#include "ptx.cuh"

namespace {

// Operand fill modes. `rows` = 128 for A, N for B.  K-major: source is [rows, K] row-major.
// MN-major (B only): source is [K, N] row-major (N contiguous), i.e. D = A * B.
//   0  K-major  SWIZZLE_128B : per 64-wide K chunk a [rows x 128B] block, 16B unit c of row r at c ^ (r & 7)
//   1  K-major  SWIZZLE_32B  : per 16-wide K step a [rows x 32B] block, unit c at c ^ ((r >> 2) & 1)
//   2  K-major  no swizzle   : per K step, 8-row groups of 256B = two 128B core matrices (k-lo, k-hi)
//   3  MN-major SWIZZLE_128B : per 64-wide N chunk a [K x 128B] block, unit c of k-row at c ^ (k & 7)
//   4  MN-major SWIZZLE_32B  : per 16-wide N chunk a [K x 32B] block
//   5  MN-major no swizzle   : per 8-wide N group a [K x 16B] block
//   6  MN-major SWIZZLE_64B  : per 32-wide N chunk a [K x 64B] block
struct ModeInfo {
  uint32_t lbo, sbo, swz, kstep_small, ksteps_per_chunk, chunk_stride;
};

__device__ __forceinline__ ModeInfo mode_info(int mode, int rows, int K) {
  ModeInfo m;
  switch (mode) {
    case 0: m = {16u, 1024u, ptx::kSwz128, 32u, 4u, (uint32_t)rows * 128u}; break;
    case 1: m = {16u, 256u, ptx::kSwz32, 0u, 1u, (uint32_t)rows * 32u}; break;
    case 2: m = {128u, 256u, ptx::kSwzNone, 0u, 1u, (uint32_t)rows * 32u}; break;
    case 3: m = {(uint32_t)K * 128u, 1024u, ptx::kSwz128, 0u, 1u, 2048u}; break;
    case 4: m = {(uint32_t)K * 32u, 256u, ptx::kSwz32, 0u, 1u, 512u}; break;
    case 5: m = {128u, (uint32_t)K * 16u, ptx::kSwzNone, 0u, 1u, 256u}; break;
    default: m = {(uint32_t)K * 64u, 512u, ptx::kSwz64, 0u, 1u, 1024u}; break;
  }
  return m;
}

// byte offset of element (r = row in M/N, k) for K-major modes
__device__ __forceinline__ uint32_t off_kmajor(int mode, int rows, int r, int k) {
  if (mode == 0) {
    const int kc = k >> 6, c = (k & 63) >> 3;
    return kc * rows * 128 + r * 128 + ((c ^ (r & 7)) << 4) + (k & 7) * 2;
  } else if (mode == 1) {
    const int ks = k >> 4, c = (k & 15) >> 3;
    return ks * rows * 32 + r * 32 + ((c ^ ((r >> 2) & 1)) << 4) + (k & 7) * 2;
  } else {
    const int ks = k >> 4, c = (k & 15) >> 3;
    return ks * rows * 32 + (r >> 3) * 256 + c * 128 + (r & 7) * 16 + (k & 7) * 2;
  }
}
// byte offset of element (k, n) for MN-major modes
__device__ __forceinline__ uint32_t off_mnmajor(int mode, int K, int k, int n) {
  if (mode == 3) {
    return (n >> 6) * K * 128 + (k >> 3) * 1024 + (k & 7) * 128 + ((((n & 63) >> 3) ^ (k & 7)) << 4) + (n & 7) * 2;
  } else if (mode == 4) {
    return (n >> 4) * K * 32 + (k >> 3) * 256 + (k & 7) * 32 + ((((n & 15) >> 3) ^ ((k & 7) >> 2)) << 4) + (n & 7) * 2;
  } else if (mode == 5) {
    return (n >> 3) * K * 16 + (k >> 3) * 128 + (k & 7) * 16 + (n & 7) * 2;
  } else {
    return (n >> 5) * K * 64 + (k >> 3) * 512 + (k & 7) * 64 + ((((n & 31) >> 3) ^ ((k & 7) >> 1)) << 4) + (n & 7) * 2;
  }
}

__global__ void __launch_bounds__(128, 1)
umma_probe_kernel(const uint16_t* __restrict__ A, const uint16_t* __restrict__ B, float* __restrict__ D,
                  const UmmaProbe p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;                      // 128 * K * 2 bytes (<= 64 KB)
  uint8_t* sB = smem + 65536;              // N * K * 2 bytes (<= 128 KB)
  __shared__ uint64_t done_bar;
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int N = p.N, K = p.K;

  if (p.a_mode != 7) {
    for (int i = tid; i < 128 * K; i += 128) {
      const int r = i / K, k = i % K;
      *reinterpret_cast<uint16_t*>(sA + off_kmajor(p.a_mode, 128, r, k)) = A[i];
    }
  }
  if (p.b_mode <= 2) {
    for (int i = tid; i < N * K; i += 128) {
      const int r = i / K, k = i % K;
      *reinterpret_cast<uint16_t*>(sB + off_kmajor(p.b_mode, N, r, k)) = B[i];
    }
  } else {
    for (int i = tid; i < N * K; i += 128) {
      const int k = i / N, n = i % N;
      *reinterpret_cast<uint16_t*>(sB + off_mnmajor(p.b_mode, K, k, n)) = B[i];
    }
  }
  if (tid == 0) {
    ptx::mbar_init(&done_bar, 1);
    ptx::fence_mbar_init();
  }
  if (warp == 0) {
    ptx::tmem_alloc(&tmem_slot, 512);
    ptx::tmem_relinquish();
  }
  if (p.a_mode == 7) {
    // a_mode 7: A lives in TENSOR MEMORY (tcgen05.mma "ts" form): lane = row, 32-bit column j = elements (2j, 2j+1)
    // of the row, element 2j in the low half.  Written with tcgen05.st by the thread that owns the lane, as a softmax
    // warp would write its probabilities.  A sits at columns [256, 256 + K/2).
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tm = tmem_slot;
    const int row = warp * 32 + lane;
    for (int c = 0; c < K / 16; ++c) {
      uint32_t v[8];
#pragma unroll
      for (int j = 0; j < 8; ++j)
        v[j] = static_cast<uint32_t>(A[row * K + c * 16 + 2 * j]) | (static_cast<uint32_t>(A[row * K + c * 16 + 2 * j + 1]) << 16);
      asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(
                       tm + (static_cast<uint32_t>(warp * 32) << 16) + 256 + c * 8),
                   "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
                   : "memory");
    }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  }
  ptx::fence_proxy_async_smem();
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = tmem_slot;

  if (tid == 0) {
    ModeInfo ma = mode_info(p.a_mode, 128, K), mb = mode_info(p.b_mode, N, K);
    if (p.a_lbo >= 0) ma.lbo = p.a_lbo;
    if (p.a_sbo >= 0) ma.sbo = p.a_sbo;
    if (p.b_lbo >= 0) mb.lbo = p.b_lbo;
    if (p.b_sbo >= 0) mb.sbo = p.b_sbo;
    const uint32_t idesc = ptx::make_idesc((uint32_t)p.fmt, 128, (uint32_t)N, 0, p.b_mode >= 3 ? 1u : 0u);
    const uint32_t a0 = ptx::smem_u32(sA), b0 = ptx::smem_u32(sB);
    for (int ks = 0; ks < K / 16 && p.a_mode == 7; ++ks) {
      const uint32_t bb = b0 + (ks / mb.ksteps_per_chunk) * mb.chunk_stride + (ks % mb.ksteps_per_chunk) * mb.kstep_small;
      ptx::mma_f16_ts(tmem, tmem + 256 + ks * 8, ptx::make_smem_desc(bb, mb.lbo, mb.sbo, mb.swz), idesc, ks != 0);
    }
    for (int ks = 0; ks < K / 16 && p.a_mode != 7; ++ks) {
      const uint32_t aa = a0 + (ks / ma.ksteps_per_chunk) * ma.chunk_stride + (ks % ma.ksteps_per_chunk) * ma.kstep_small;
      const uint32_t bb = b0 + (ks / mb.ksteps_per_chunk) * mb.chunk_stride + (ks % mb.ksteps_per_chunk) * mb.kstep_small;
      ptx::mma_f16_ss(tmem, ptx::make_smem_desc(aa, ma.lbo, ma.sbo, ma.swz),
                      ptx::make_smem_desc(bb, mb.lbo, mb.sbo, mb.swz), idesc, ks != 0);
    }
    ptx::mma_commit(&done_bar);
  }
  ptx::mbar_wait(&done_bar, 0);
  ptx::tc_fence_after();
  const int row = warp * 32 + lane;
  for (int c = 0; c < N / 16; ++c) {
    uint32_t v[16];
    ptx::tmem_ld_32x32b_x16(tmem + (static_cast<uint32_t>(warp * 32) << 16) + c * 16, v);
    ptx::tmem_ld_wait();
#pragma unroll
    for (int i = 0; i < 16; ++i) D[row * N + c * 16 + i] = __uint_as_float(v[i]);
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem, 512);
  }
}

}  // namespace

int samk_umma_probe(const void* A, const void* B, float* D, const UmmaProbe& p, cudaStream_t stream) {
  SAM_REQUIRE(p.N % 16 == 0 && p.N >= 16 && p.N <= 256, "probe: bad N %d", p.N);
  SAM_REQUIRE(p.K % 16 == 0 && p.K >= 16 && p.K <= 256, "probe: bad K %d", p.K);
  SAM_REQUIRE(((p.a_mode >= 0 && p.a_mode <= 2) || p.a_mode == 7) && p.b_mode >= 0 && p.b_mode <= 6, "probe: bad mode");
  const int smem = 65536 + 131072 + 1024;
  static samhost::PerDeviceOnce attr_once;
  if (attr_once.need()) {
    SAM_CHECK_CUDA(cudaFuncSetAttribute(umma_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    attr_once.done();
  }
  umma_probe_kernel<<<1, 128, smem, stream>>>(static_cast<const uint16_t*>(A), static_cast<const uint16_t*>(B), D, p);
  SAM_CHECK_CUDA(cudaGetLastError());
  return 0;
}
