// MUFU.EX2 issue-rate probe: cycles per warp-level ex2.approx.ftz.f32 as a function of warps per SM sub-partition, with
// and without FFMA2 filler between the MUFUs.   nvcc -arch=sm_100a -O3 -o mufu_probe mufu_probe.cu && ./mufu_probe
#include <cstdio>
#include <cuda_runtime.h>

template <int FILL>
__global__ void probe(float* out, long long* cyc, int iters) {
  float x[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) x[i] = -0.001f * (threadIdx.x + i);
  float f0 = 1.0f, f1 = 0.5f;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x[i]));
#pragma unroll
      for (int k = 0; k < FILL; ++k) asm volatile("fma.rn.f32 %0, %0, %1, %1;" : "+f"(f0) : "f"(f1));
    }
  }
  const long long t1 = clock64();
  float s = f0;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += x[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

int main() {
  float* out;
  long long* cyc;
  cudaMalloc(&out, 148 * 1024 * 4);
  cudaMalloc(&cyc, 148 * 8);
  const int iters = 2000;
  for (int fill = 0; fill <= 2; ++fill)
    for (int warps : {1, 2, 4, 8, 12, 16}) {
      for (int rep = 0; rep < 2; ++rep) {
        if (fill == 0) probe<0><<<148, warps * 32>>>(out, cyc, iters);
        if (fill == 1) probe<3><<<148, warps * 32>>>(out, cyc, iters);
        if (fill == 2) probe<6><<<148, warps * 32>>>(out, cyc, iters);
        cudaDeviceSynchronize();
      }
      long long h[148];
      cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
      const double per_sm = double(h[0]) / (double(iters) * 16);   // cycles per round of (one MUFU per warp)
      printf("fill=%d warps/SM=%2d (per sub-partition %.1f): %.2f cycles per MUFU per warp, SM rate %.2f lanes/clk\n",
             fill * 3, warps, warps / 4.0, per_sm, warps * 32.0 / per_sm);
    }
  return 0;
}
