// MUFU.EX2 rate probe, fp32 vs packed half: cycles per warp-level ex2.approx.ftz.f32 and ex2.approx.f16x2 (two
// results per instruction) as a function of warps per SM; also the f32x2 -> f16x2 convert + ex2.f16x2 pair.
//   nvcc -arch=sm_100a -O3 -o mufu_f16x2_probe mufu_f16x2_probe.cu && ./mufu_f16x2_probe
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

template <int MODE>   // 0: ex2.f32   1: ex2.f16x2   2: cvt.f16x2.f32 + ex2.f16x2   3: ex2.bf16x2
__global__ void probe(float* out, long long* cyc, int iters) {
  float x[16];
  uint32_t h[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    x[i] = -0.001f * (threadIdx.x + i);
    h[i] = 0xb800b800u + i;   // -0.5 in both halves (approximately)
  }
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      if (MODE == 0) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x[i]));
      if (MODE == 1) asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(h[i]));
      if (MODE == 2) {
        asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(h[i]) : "f"(x[i]), "f"(x[(i + 1) & 15]));
        asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(h[i]));
      }
      if (MODE == 3) asm volatile("ex2.approx.ftz.bf16x2 %0, %0;" : "+r"(h[i]));
    }
  }
  const long long t1 = clock64();
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += x[i] + static_cast<float>(h[i] & 0xffff);
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

int main() {
  float* out;
  long long* cyc;
  cudaMalloc(&out, 148 * 1024 * 4);
  cudaMalloc(&cyc, 148 * 8);
  const int iters = 2000;
  const char* names[4] = {"ex2.f32", "ex2.f16x2", "cvt+ex2.f16x2", "ex2.bf16x2"};
  for (int mode = 0; mode < 4; ++mode)
    for (int warps : {1, 4, 8, 16}) {
      for (int rep = 0; rep < 2; ++rep) {
        if (mode == 0) probe<0><<<148, warps * 32>>>(out, cyc, iters);
        if (mode == 1) probe<1><<<148, warps * 32>>>(out, cyc, iters);
        if (mode == 2) probe<2><<<148, warps * 32>>>(out, cyc, iters);
        if (mode == 3) probe<3><<<148, warps * 32>>>(out, cyc, iters);
        cudaDeviceSynchronize();
      }
      long long h[148];
      cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
      const double per = double(h[0]) / (double(iters) * 16);
      printf("%-14s warps/SM=%2d: %.2f cycles per instruction per warp, SM rate %.2f instr-lanes/clk = %.1f results/clk\n",
             names[mode], warps, per, warps * 32.0 / per, warps * 32.0 / per * (mode == 0 ? 1 : 2));
    }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
