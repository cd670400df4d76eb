// Internal C++ interface between the kernel translation units and the C-ABI layer (capi.cpp).
// Every function enqueues work on `stream`, allocates nothing, never synchronises, and returns 0 or an error code
// whose text is available through samhost::last_error().
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

// Operand / storage formats used across the library.
//   0 = fp16, 1 = bf16 (tensor-core operand formats; identical to the tcgen05 idesc encoding), 2 = fp32
enum SamFmt : int { SAM_F16 = 0, SAM_BF16 = 1, SAM_F32 = 2 };

struct GemmEpilogue {
  void* out;         // [M, ldo] in out_fmt
  int ldo;           // elements
  int out_fmt;       // SamFmt
  const float* bias; // [N] or null
  int act;           // 0 none, 1 exact-erf GELU
  const float* res;  // fp32 residual source or null; row used = (row % res_mod); may alias `out` (fp32, in place)
  int ldr;
  int res_mod;
};

int samk_gemm(const void* A, int lda, const void* W, int ldw, int M, int N, int K, int fmt, const GemmEpilogue& ep,
              cudaStream_t stream);

// UMMA layout probe (test-only kernel, see probe.cu).
struct UmmaProbe {
  int N;          // MMA N (multiple of 16, <= 256); M is fixed at 128
  int K;          // multiple of 16, <= 256
  int fmt;        // 0 fp16, 1 bf16
  int a_mode;     // smem fill + descriptor mode for A (see probe.cu)
  int b_mode;     // same for B
  int a_lbo, a_sbo, a_kstep;  // descriptor byte offsets and per-16-K start-address advance (bytes)
  int b_lbo, b_sbo, b_kstep;
};
int samk_umma_probe(const void* A, const void* B, float* D, const UmmaProbe& p, cudaStream_t stream);
