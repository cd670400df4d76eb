// Micro-benchmark: throughput of tcgen05.ld (32x32b) by vector width.  Each warp sweeps the same 64 fp32 columns of
// its lane quadrant `iters` times as 4 x .x16, 2 x .x32 or 1 x .x64 (loads back to back, one wait::ld per sweep).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/probes/tmem_ld_probe tools/probes/tmem_ld_probe.cu
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

#define LD16(addr, v, o)                                                                                              \
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];" \
               : "=r"(v[o + 0]), "=r"(v[o + 1]), "=r"(v[o + 2]), "=r"(v[o + 3]), "=r"(v[o + 4]), "=r"(v[o + 5]),      \
                 "=r"(v[o + 6]), "=r"(v[o + 7]), "=r"(v[o + 8]), "=r"(v[o + 9]), "=r"(v[o + 10]), "=r"(v[o + 11]),    \
                 "=r"(v[o + 12]), "=r"(v[o + 13]), "=r"(v[o + 14]), "=r"(v[o + 15])                                  \
               : "r"(addr) : "memory")
#define LD32(addr, v, o)                                                                                              \
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];" \
               : "=r"(v[o + 0]), "=r"(v[o + 1]), "=r"(v[o + 2]), "=r"(v[o + 3]), "=r"(v[o + 4]), "=r"(v[o + 5]),      \
                 "=r"(v[o + 6]), "=r"(v[o + 7]), "=r"(v[o + 8]), "=r"(v[o + 9]), "=r"(v[o + 10]), "=r"(v[o + 11]),    \
                 "=r"(v[o + 12]), "=r"(v[o + 13]), "=r"(v[o + 14]), "=r"(v[o + 15]), "=r"(v[o + 16]), "=r"(v[o + 17]), \
                 "=r"(v[o + 18]), "=r"(v[o + 19]), "=r"(v[o + 20]), "=r"(v[o + 21]), "=r"(v[o + 22]), "=r"(v[o + 23]), \
                 "=r"(v[o + 24]), "=r"(v[o + 25]), "=r"(v[o + 26]), "=r"(v[o + 27]), "=r"(v[o + 28]), "=r"(v[o + 29]), \
                 "=r"(v[o + 30]), "=r"(v[o + 31])                                                                     \
               : "r"(addr) : "memory")

template <int MODE>   // 0: 4 x x16, 1: 2 x x32, 2: 8 x x8-like (x16 issued on 8-column strides: 8 loads, half overlapping)
__global__ void probe(int iters, long long* cycles, uint32_t* sink) {
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t base = slot + (static_cast<uint32_t>((warp & 3) * 32) << 16) + (warp >> 2) * 64;
  uint32_t v[64];
  uint32_t acc = 0;
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    if (MODE == 0) {
      LD16(base, v, 0); LD16(base + 16, v, 16); LD16(base + 32, v, 32); LD16(base + 48, v, 48);
    } else {
      LD32(base, v, 0); LD32(base + 32, v, 32);
    }
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 64; ++i) acc ^= v[i];
  }
  const long long t1 = clock64();
  if ((threadIdx.x & 31) == 0) cycles[blockIdx.x * (blockDim.x >> 5) + warp] = t1 - t0;
  sink[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(slot), "r"(512u) : "memory");
}

int main() {
  const int iters = 20000;
  long long* cyc;
  uint32_t* sink;
  cudaMalloc(&cyc, 148 * 16 * sizeof(long long));
  cudaMalloc(&sink, 148 * 512 * sizeof(uint32_t));
  for (int warps = 4; warps <= 16; warps *= 2) {
    for (int mode = 0; mode < 2; ++mode) {
      if (mode == 0) probe<0><<<148, warps * 32>>>(iters, cyc, sink); else probe<1><<<148, warps * 32>>>(iters, cyc, sink);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("error: %s\n", cudaGetErrorString(e)); return 1; }
      long long h[148 * 16];
      cudaMemcpy(h, cyc, sizeof(long long) * 148 * warps, cudaMemcpyDeviceToHost);
      double mx = 0;
      for (int i = 0; i < 148 * warps; ++i) mx = h[i] > mx ? h[i] : mx;
      const double per_sweep = mx / iters;
      const double bytes_per_quadrant = 32.0 * 64 * 4 * (warps / 4);
      printf("%2d warps (%d per quadrant), %s: %.1f cycles per 64-column sweep per warp -> %.1f B/clk per quadrant, %.1f B/clk per SM\n",
             warps, warps / 4, mode == 0 ? "4 x .x16" : "2 x .x32", per_sweep, bytes_per_quadrant / per_sweep,
             4 * bytes_per_quadrant / per_sweep);
    }
  }
  return 0;
}
