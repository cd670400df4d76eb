// Training path of the mask decoder: forward that keeps its intermediates + backward.
//
// AnyRef fine-tunes SAM's mask decoder with the image encoder and the prompt encoder frozen (model/anyref.py:108-113);
// the loss reaches it through `mask_decoder(...)` -> `postprocess_masks` (model/anyref.py:406-450).  The gradients this
// file produces are therefore those of every mask_decoder.* parameter and of `sparse_prompt_embeddings` (which carries
// them on to text_hidden_fcs and the LLM); image embeddings, the dense prompt embedding and image_pe get none.
//
// Design.  The inference decoder (decoder.cu) fuses whole sub-blocks into single kernels and keeps nothing; a backward
// needs the intermediates.  This path is a separate, deliberately plain fp32 composition: a small set of kernels
// (one strided / batched / split-K SGEMM, LayerNorm, softmax, GELU / ReLU, column sums, layout gathers), each with its
// adjoint, composed on the host exactly in the order of the reference modules
//   MaskDecoder.predict_masks            modeling/mask_decoder.py:116-179
//   TwoWayTransformer.forward            modeling/transformer.py:62-106
//   TwoWayAttentionBlock.forward         modeling/transformer.py:151-182
//   Attention.forward                    modeling/transformer.py:220-242
// Every forward op pushes the closure of its adjoint on a tape; sam_decoder_backward runs the tape in reverse.  Every
// adjoint ACCUMULATES into the gradient buffers of its inputs (zeroed once), so fan-out needs no special handling.
// Reduction orders are fixed (split-K and per-chunk partials are folded in index order): the gradients are deterministic.
// Products with >= 8192 rows (the image side) run on the tensor cores: operands split into bf16 pairs (x = hi + lo,
// x.w ~ hi.hi + hi.lo + lo.hi as ONE tcgen05 GEMM over 3K, fp32-accurate) through the 2-CTA GEMM of the inference path
// -- forward, input gradient (in-place reduce-add epilogue) and weight gradient (block-diagonal mode: one GEMM for all
// row chunks, Tape::gemm_tc_dw).  Everything else is fp32 FMA on the SGEMM below.  The same composition without its tape
// is the inference path for prompts with more than 16 tokens (samk_decoder_forward_generic).  DESIGN.md 4b.
#include <cuda_runtime.h>
#include <stdint.h>

#include <algorithm>
#include <cstdlib>
#include <functional>
#include <vector>

#include "host_common.h"
#include "kernels.h"
#include "ptx.cuh"

namespace {

typedef long long i64;

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float load_any(const void* p, int fmt, size_t i) {
  if (fmt == 2) return static_cast<const float*>(p)[i];
  return ptx::unpack1(static_cast<const uint16_t*>(p)[i], fmt);
}

// ---------------------------------------------------------------------------------------------------------------
// SGEMM:  C[b] (+)= alpha * A[b] . B[b] (+ bias),  A(i,k) = A[i*ars + k*acs], B(k,j) = B[k*brs + j*bcs], C(i,j) = C[i*crs + j]
// (any operand may be a transposed or strided view), two-level batch (z / nb2, z % nb2) with independent strides, and
// split-K into a partial buffer [split][batch][M][N] that splitk_reduce_kernel folds in index order.
// 64 x 64 x 16 tiles, 256 threads, 4 x 4 outputs per thread.
// ---------------------------------------------------------------------------------------------------------------
struct Gemm {
  const float* A;
  const float* B;
  float* C;
  const float* bias;   // [bias_mod] or null: added as bias[j % bias_mod]
  int M, N, K;
  i64 ars, acs, brs, bcs, crs;
  i64 ccs;   // column stride of C (0 = 1)
  int nbatch, nb2;
  i64 a_b1, a_b2, b_b1, b_b2, c_b1, c_b2;
  int bias_mod;
  float alpha;
  int accumulate;
  int splits, kchunk;
  float* partial;
  int part_rows;   // rows per partial block (0: M) -- the tensor-core weight gradient pads its blocks to 256 rows
};

constexpr int GT = 64, GK = 16;       // the general tile
constexpr int HM = 16, HN = 32, HK = 64;   // the thin tile: M <= 16, N <= 32 and a long K (attention against 6 tokens)

// TM x TN x TK tiles, 256 threads as 16 x 16, (TM/16) x (TN/16) outputs per thread
template <int TM, int TN, int TK>
__global__ void __launch_bounds__(256) sgemm_kernel(const Gemm g) {
  constexpr int RM = TM / 16, RN = TN / 16;
  constexpr int SA = TM * TK / 256, SB = TK * TN / 256;   // tile-load slots per thread
  __shared__ __align__(16) float As[TK][TM + 4];
  __shared__ __align__(16) float Bs[TK][TN + 4];
  const int z = blockIdx.z;
  const int batch = z / g.splits, split = z - batch * g.splits;
  const int b1 = batch / g.nb2, b2 = batch - b1 * g.nb2;
  const float* __restrict__ A = g.A + b1 * g.a_b1 + b2 * g.a_b2;
  const float* __restrict__ B = g.B + b1 * g.b_b1 + b2 * g.b_b2;
  const int k0 = split * g.kchunk;
  const int k1 = min(g.K, k0 + g.kchunk);
  const int m0 = blockIdx.x * TM, n0 = blockIdx.y * TN;
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  // consecutive threads walk whichever index of the operand is contiguous in memory
  const bool a_kfast = g.acs == 1 && g.ars != 1;
  const bool b_kfast = g.brs == 1 && g.bcs != 1;
  float acc[RM][RN];
#pragma unroll
  for (int i = 0; i < RM; ++i)
#pragma unroll
    for (int j = 0; j < RN; ++j) acc[i][j] = 0.f;
  // per-thread slots of the tile loads, fixed over the k loop
  int ai[SA], ak[SA], bj[SB], bk[SB];
#pragma unroll
  for (int r = 0; r < SA; ++r) {
    const int idx = tid + r * 256;
    if (a_kfast) { ak[r] = idx % TK; ai[r] = idx / TK; } else { ai[r] = idx % TM; ak[r] = idx / TM; }
  }
#pragma unroll
  for (int r = 0; r < SB; ++r) {
    const int idx = tid + r * 256;
    if (b_kfast) { bk[r] = idx % TK; bj[r] = idx / TK; } else { bj[r] = idx % TN; bk[r] = idx / TN; }
  }
  float ra[SA], rb[SB];
  auto fetch = [&](int kk) {
#pragma unroll
    for (int r = 0; r < SA; ++r) {
      const int gi = m0 + ai[r], gka = kk + ak[r];
      ra[r] = (gi < g.M && gka < k1) ? A[gi * g.ars + gka * g.acs] : 0.f;
    }
#pragma unroll
    for (int r = 0; r < SB; ++r) {
      const int gj = n0 + bj[r], gkb = kk + bk[r];
      rb[r] = (gj < g.N && gkb < k1) ? B[gkb * g.brs + gj * g.bcs] : 0.f;
    }
  };
  fetch(k0);
  for (int kk = k0; kk < k1; kk += TK) {
#pragma unroll
    for (int r = 0; r < SA; ++r) As[ak[r]][ai[r]] = ra[r];
#pragma unroll
    for (int r = 0; r < SB; ++r) Bs[bk[r]][bj[r]] = rb[r];
    __syncthreads();
    if (kk + TK < k1) fetch(kk + TK);   // the next tile's global loads fly behind this tile's arithmetic
#pragma unroll
    for (int k = 0; k < TK; ++k) {
      float av[RM], bv[RN];
      if constexpr (RM == 4) {
        const float4 a = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
        av[0] = a.x; av[1] = a.y; av[2] = a.z; av[3] = a.w;
      } else {
#pragma unroll
        for (int i = 0; i < RM; ++i) av[i] = As[k][ty * RM + i];
      }
      if constexpr (RN == 4) {
        const float4 b = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
        bv[0] = b.x; bv[1] = b.y; bv[2] = b.z; bv[3] = b.w;
      } else {
#pragma unroll
        for (int j = 0; j < RN; ++j) bv[j] = Bs[k][tx * RN + j];
      }
#pragma unroll
      for (int i = 0; i < RM; ++i)
#pragma unroll
        for (int j = 0; j < RN; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }
  if (g.splits > 1) {
    float* P = g.partial + (static_cast<i64>(split) * g.nbatch + batch) * g.M * g.N;
#pragma unroll
    for (int i = 0; i < RM; ++i) {
      const int gi = m0 + ty * RM + i;
      if (gi >= g.M) continue;
#pragma unroll
      for (int j = 0; j < RN; ++j) {
        const int gj = n0 + tx * RN + j;
        if (gj < g.N) P[static_cast<i64>(gi) * g.N + gj] = acc[i][j];
      }
    }
    return;
  }
  float* __restrict__ Cb = g.C + b1 * g.c_b1 + b2 * g.c_b2;
#pragma unroll
  for (int i = 0; i < RM; ++i) {
    const int gi = m0 + ty * RM + i;
    if (gi >= g.M) continue;
#pragma unroll
    for (int j = 0; j < RN; ++j) {
      const int gj = n0 + tx * RN + j;
      if (gj >= g.N) continue;
      float v = g.alpha * acc[i][j];
      if (g.bias) v += g.bias[gj % g.bias_mod];
      float* c = Cb + gi * g.crs + gj * (g.ccs ? g.ccs : 1);
      *c = g.accumulate ? *c + v : v;
    }
  }
}

// Tall and shallow products (M >= 1024 rows, N <= 32, K <= 32: attention of 4096 pixels against 6 tokens, the mask
// product, ...): one thread per output row, B (K x N) in shared memory; no tile padding, every byte read is used.
// grid (ceil(M / 256), batch), block 256
template <int NT>
__global__ void __launch_bounds__(256) sgemm_rows_kernel(const Gemm g) {
  __shared__ float Bs[32][NT + 1];
  const int batch = blockIdx.y;
  const int b1 = batch / g.nb2, b2 = batch - b1 * g.nb2;
  const float* __restrict__ A = g.A + b1 * g.a_b1 + b2 * g.a_b2;
  const float* __restrict__ B = g.B + b1 * g.b_b1 + b2 * g.b_b2;
  for (int idx = threadIdx.x; idx < g.K * NT; idx += 256) {
    const int k = idx / NT, j = idx - k * NT;
    Bs[k][j] = j < g.N ? B[k * g.brs + j * g.bcs] : 0.f;
  }
  __syncthreads();
  const int row = blockIdx.x * 256 + threadIdx.x;
  if (row >= g.M) return;
  float acc[NT];
#pragma unroll
  for (int j = 0; j < NT; ++j) acc[j] = 0.f;
  const float* a = A + row * g.ars;
  for (int k = 0; k < g.K; ++k) {
    const float av = a[k * g.acs];
#pragma unroll
    for (int j = 0; j < NT; ++j) acc[j] = fmaf(av, Bs[k][j], acc[j]);
  }
  float* c = g.C + b1 * g.c_b1 + b2 * g.c_b2 + row * g.crs;
  const i64 ccs = g.ccs ? g.ccs : 1;
#pragma unroll
  for (int j = 0; j < NT; ++j) {
    if (j < g.N) {
      float v = g.alpha * acc[j];
      if (g.bias) v += g.bias[j % g.bias_mod];
      c[j * ccs] = g.accumulate ? c[j * ccs] + v : v;
    }
  }
}

__global__ void splitk_reduce_kernel(const Gemm g) {
  const i64 per = static_cast<i64>(g.M) * g.N;
  const i64 pitch = static_cast<i64>(g.part_rows ? g.part_rows : g.M) * g.N;   // elements per partial block
  const i64 total = per * g.nbatch;
  for (i64 t = static_cast<i64>(blockIdx.x) * blockDim.x + threadIdx.x; t < total; t += static_cast<i64>(gridDim.x) * blockDim.x) {
    const int batch = static_cast<int>(t / per);
    const i64 r = t - batch * per;
    const int i = static_cast<int>(r / g.N), j = static_cast<int>(r - static_cast<i64>(i) * g.N);
    // four independent chains (loads in flight), combined in a fixed order: deterministic
    float s4[4] = {0.f, 0.f, 0.f, 0.f};
    const float* pp = g.partial + static_cast<i64>(batch) * pitch + r;
    const i64 step = static_cast<i64>(g.nbatch) * pitch;
    int sp = 0;
    for (; sp + 4 <= g.splits; sp += 4) {
#pragma unroll
      for (int u = 0; u < 4; ++u) s4[u] += pp[(sp + u) * step];
    }
    for (; sp < g.splits; ++sp) s4[0] += pp[sp * step];
    float v = g.alpha * ((s4[0] + s4[1]) + (s4[2] + s4[3]));
    if (g.bias) v += g.bias[j % g.bias_mod];
    const int b1 = batch / g.nb2, b2 = batch - b1 * g.nb2;
    float* c = g.C + b1 * g.c_b1 + b2 * g.c_b2 + i * g.crs + j * (g.ccs ? g.ccs : 1);
    *c = g.accumulate ? *c + v : v;
  }
}

// ---------------------------------------------------------------------------------------------------------------
// Column sums:  out[j % mod] (+)= sum_r X[r, j]   (bias gradients, LayerNorm weight gradients from block partials)
// stage 1: grid (ceil(N/32), chunks), block (32, 8) -> partial[chunk][N];  stage 2: one thread per output column
// ---------------------------------------------------------------------------------------------------------------
constexpr int kColRows = 1024;   // rows per chunk

__global__ void __launch_bounds__(256) colsum_partial_kernel(const float* __restrict__ X, i64 ld, i64 R, int N, float* __restrict__ partial) {
  __shared__ float red[8][33];
  const int j = blockIdx.x * 32 + threadIdx.x;
  const i64 r0 = static_cast<i64>(blockIdx.y) * kColRows;
  const i64 r1 = min(R, r0 + kColRows);
  float s = 0.f;
  if (j < N)
    for (i64 r = r0 + threadIdx.y; r < r1; r += 8) s += X[r * ld + j];
  red[threadIdx.y][threadIdx.x] = s;
  __syncthreads();
  if (threadIdx.y == 0 && j < N) {
    float t = 0.f;
#pragma unroll
    for (int y = 0; y < 8; ++y) t += red[y][threadIdx.x];
    partial[static_cast<i64>(blockIdx.y) * N + j] = t;
  }
}
// R <= kColRows and one output per column: a single launch adds straight into out
__global__ void __launch_bounds__(256) colsum_small_kernel(const float* __restrict__ X, i64 ld, i64 R, int N, float* __restrict__ out) {
  __shared__ float red[8][33];
  const int j = blockIdx.x * 32 + threadIdx.x;
  float s = 0.f;
  if (j < N)
    for (i64 r = threadIdx.y; r < R; r += 8) s += X[r * ld + j];
  red[threadIdx.y][threadIdx.x] = s;
  __syncthreads();
  if (threadIdx.y == 0 && j < N) {
    float t = 0.f;
#pragma unroll
    for (int y = 0; y < 8; ++y) t += red[y][threadIdx.x];
    out[j] += t;
  }
}
// one warp per output column: lanes stride over the (chunk, column group) partials, fixed-order warp reduction
__global__ void __launch_bounds__(256) colsum_finish_kernel(const float* __restrict__ partial, int chunks, int N, int mod, float* __restrict__ out) {
  const int j = blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (j >= mod) return;
  const int groups = N / mod;
  float s = 0.f;
  for (int i = lane; i < chunks * groups; i += 32) {
    const int c = i / groups, gq = i - c * groups;
    s += partial[static_cast<i64>(c) * N + gq * mod + j];
  }
  s = warp_sum(s);
  if (lane == 0) out[j] += s;
}

// ---------------------------------------------------------------------------------------------------------------
// LayerNorm over the last dimension (C % 32 == 0, C <= 256), one warp per row
// ---------------------------------------------------------------------------------------------------------------
constexpr int LNV = 8;   // values per lane

__global__ void __launch_bounds__(256) ln_fwd_kernel(const float* __restrict__ X, const float* __restrict__ gw, const float* __restrict__ gb,
                                                     float* __restrict__ Y, float2* __restrict__ stats, i64 R, int C, float eps) {
  const i64 row = static_cast<i64>(blockIdx.x) * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= R) return;
  const int nv = C >> 5;
  float v[LNV];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < LNV; ++i)
    if (i < nv) {
      v[i] = X[row * C + lane + 32 * i];
      s += v[i];
    }
  const float mean = warp_sum(s) / C;
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < LNV; ++i)
    if (i < nv) {
      v[i] -= mean;
      q = fmaf(v[i], v[i], q);
    }
  const float rstd = 1.0f / sqrtf(warp_sum(q) / C + eps);
#pragma unroll
  for (int i = 0; i < LNV; ++i)
    if (i < nv) {
      const int c = lane + 32 * i;
      Y[row * C + c] = v[i] * rstd * gw[c] + gb[c];
    }
  if (lane == 0) stats[row] = make_float2(mean, rstd);
}

// dX += rstd * (dy*g - mean(dy*g) - xhat * mean(dy*g*xhat));  block partials of sum dy*xhat | sum dy -> part[block][2C]
constexpr int kLnRowsPerBlock = 64;

__global__ void __launch_bounds__(256) ln_bwd_kernel(const float* __restrict__ dY, const float* __restrict__ X, const float2* __restrict__ stats,
                                                     const float* __restrict__ gw, float* __restrict__ dX, float* __restrict__ part,
                                                     i64 R, int C) {
  __shared__ float red[8][2 * 256];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nv = C >> 5;
  float ag[LNV], ab[LNV];
#pragma unroll
  for (int i = 0; i < LNV; ++i) ag[i] = ab[i] = 0.f;
  const i64 r0 = static_cast<i64>(blockIdx.x) * kLnRowsPerBlock;
  for (int rr = warp; rr < kLnRowsPerBlock; rr += 8) {
    const i64 row = r0 + rr;
    if (row >= R) break;
    const float2 st = stats[row];
    float xh[LNV], dg[LNV];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < LNV; ++i)
      if (i < nv) {
        const int c = lane + 32 * i;
        const float dy = dY[row * C + c];
        xh[i] = (X[row * C + c] - st.x) * st.y;
        dg[i] = dy * gw[c];
        s1 += dg[i];
        s2 = fmaf(dg[i], xh[i], s2);
        ag[i] = fmaf(dy, xh[i], ag[i]);
        ab[i] += dy;
      }
    s1 = warp_sum(s1) / C;
    s2 = warp_sum(s2) / C;
    if (dX) {
#pragma unroll
      for (int i = 0; i < LNV; ++i)
        if (i < nv) dX[row * C + lane + 32 * i] += st.y * (dg[i] - s1 - xh[i] * s2);
    }
  }
#pragma unroll
  for (int i = 0; i < LNV; ++i)
    if (i < nv) {
      red[warp][lane + 32 * i] = ag[i];
      red[warp][C + lane + 32 * i] = ab[i];
    }
  __syncthreads();
  for (int c = threadIdx.x; c < 2 * C; c += 256) {
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) s += red[w][c];
    part[static_cast<i64>(blockIdx.x) * 2 * C + c] = s;
  }
}

// ---------------------------------------------------------------------------------------------------------------
// softmax over rows of length L (in place), and its adjoint dS = P * (dP - sum_j dP_j P_j) (in place on dP)
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) softmax_fwd_kernel(float* __restrict__ S, i64 R, int L) {
  const i64 row = static_cast<i64>(blockIdx.x) * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= R) return;
  float* s = S + row * L;
  float m = -INFINITY;
  for (int j = lane; j < L; j += 32) m = fmaxf(m, s[j]);
  m = warp_max(m);
  float sum = 0.f;
  for (int j = lane; j < L; j += 32) {
    const float e = expf(s[j] - m);
    s[j] = e;
    sum += e;
  }
  const float inv = 1.0f / warp_sum(sum);
  for (int j = lane; j < L; j += 32) s[j] *= inv;
}
__global__ void __launch_bounds__(256) softmax_bwd_kernel(float* __restrict__ dP, const float* __restrict__ P, i64 R, int L) {
  const i64 row = static_cast<i64>(blockIdx.x) * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= R) return;
  float* d = dP + row * L;
  const float* p = P + row * L;
  float dot = 0.f;
  for (int j = lane; j < L; j += 32) dot = fmaf(d[j], p[j], dot);
  dot = warp_sum(dot);
  for (int j = lane; j < L; j += 32) d[j] = p[j] * (d[j] - dot);
}

// ---------------------------------------------------------------------------------------------------------------
// element-wise
// ---------------------------------------------------------------------------------------------------------------
#define GRID_STRIDE(i, n) for (i64 i = static_cast<i64>(blockIdx.x) * blockDim.x + threadIdx.x; i < (n); i += static_cast<i64>(gridDim.x) * blockDim.x)

__global__ void add_kernel(const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ c, i64 n, i64 bmod) {
  GRID_STRIDE(i, n) c[i] = a[i] + b[i % bmod];
}
__global__ void acc_kernel(float* __restrict__ dst, const float* __restrict__ src, i64 n) {
  GRID_STRIDE(i, n) dst[i] += src[i];
}
__global__ void relu_fwd_kernel(const float* __restrict__ x, float* __restrict__ y, i64 n) {
  GRID_STRIDE(i, n) y[i] = fmaxf(x[i], 0.f);
}
__global__ void relu_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ y, float* __restrict__ dx, i64 n) {
  GRID_STRIDE(i, n) if (y[i] > 0.f) dx[i] += dy[i];
}
__global__ void relu_mask_kernel(float* __restrict__ dy, const float* __restrict__ y, i64 n) {
  GRID_STRIDE(i, n) if (!(y[i] > 0.f)) dy[i] = 0.f;
}
// exact-erf GELU (common.py:18 / nn.GELU default) and its derivative Phi(x) + x phi(x)
__global__ void gelu_fwd_kernel(const float* __restrict__ x, float* __restrict__ y, i64 n) {
  GRID_STRIDE(i, n) y[i] = 0.5f * x[i] * (1.0f + erff(x[i] * 0.70710678118654752f));
}
__global__ void gelu_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ x, float* __restrict__ dx, i64 n) {
  GRID_STRIDE(i, n) {
    const float v = x[i];
    const float cdf = 0.5f * (1.0f + erff(v * 0.70710678118654752f));
    const float pdf = 0.3989422804014327f * expf(-0.5f * v * v);
    dx[i] += dy[i] * (cdf + v * pdf);
  }
}

// split-bf16 operands of the tensor-core GEMM (x = hi + lo, x.w ~ hi.hi + hi.lo + lo.hi as ONE product over 3K):
// activations [rows, K] -> [rows, 3K] = [hi | hi | lo], weights [N, K] -> [N, 3K] = [hi | lo | hi]
__device__ __forceinline__ void split_bf16(float x, uint16_t& hi, uint16_t& lo) {
  const __nv_bfloat16 h = __float2bfloat16_rn(x);
  const __nv_bfloat16 l = __float2bfloat16_rn(x - __bfloat162float(h));
  hi = *reinterpret_cast<const uint16_t*>(&h);
  lo = *reinterpret_cast<const uint16_t*>(&l);
}
// four consecutive elements per thread (K % 4 == 0): one 16-byte load, three 8-byte stores
__global__ void split_act_kernel(const float* __restrict__ X, uint16_t* __restrict__ A3, i64 rows, int K) {
  const int K4 = K >> 2;
  GRID_STRIDE(i, rows * K4) {
    const i64 r = i / K4;
    const int c = static_cast<int>(i - r * K4) * 4;
    const float4 v = *reinterpret_cast<const float4*>(X + r * K + c);
    uint16_t h[4], l[4];
    split_bf16(v.x, h[0], l[0]); split_bf16(v.y, h[1], l[1]); split_bf16(v.z, h[2], l[2]); split_bf16(v.w, h[3], l[3]);
    const uint2 H = make_uint2(h[0] | (uint32_t(h[1]) << 16), h[2] | (uint32_t(h[3]) << 16));
    const uint2 Lo = make_uint2(l[0] | (uint32_t(l[1]) << 16), l[2] | (uint32_t(l[3]) << 16));
    uint16_t* o = A3 + r * (3 * K) + c;
    *reinterpret_cast<uint2*>(o) = H;
    *reinterpret_cast<uint2*>(o + K) = H;
    *reinterpret_cast<uint2*>(o + 2 * K) = Lo;
  }
}
// W [N, K] (element (n, k) at W[n*K + k]) -> out [R, 3*Cc] with (r, c) = (n, k), or (k, n) when transposed
__global__ void split_weight_kernel(const float* __restrict__ W, uint16_t* __restrict__ out, int N, int K, int transposed) {
  const int R = transposed ? K : N, Cc = transposed ? N : K;
  GRID_STRIDE(i, static_cast<i64>(R) * Cc) {
    const int r = static_cast<int>(i / Cc), c = static_cast<int>(i - static_cast<i64>(r) * Cc);
    const float x = transposed ? W[static_cast<i64>(c) * K + r] : W[static_cast<i64>(r) * K + c];
    uint16_t h, l;
    split_bf16(x, h, l);
    uint16_t* o = out + static_cast<i64>(r) * (3 * Cc) + c;
    o[0] = h; o[Cc] = l; o[2 * Cc] = h;
  }
}
// Operands of a weight gradient dW[N, K] = dY^T . X on the tensor cores.  The reduction runs over the rows, so both
// operands are needed transposed; they are written chunk-major for the block-diagonal GEMM (GemmEpilogue::diag_*):
//   out[(s * pitch + i) * 3L + c] = split(Y[s * L + c, i])   for chunk s of L rows (zero beyond the last row),
// [hi | hi | lo] for the A role (dY), [hi | lo | hi] for the W role (X).   grid (L/64, ceil(N/32), S), block (32, 8)
__global__ void __launch_bounds__(256)
split_transposed_kernel(const float* __restrict__ Y, i64 ld, i64 rows, int N, int L, int pitch, int w_role, uint16_t* __restrict__ out) {
  __shared__ float tile[64][33];   // [c][i]
  const int s = blockIdx.z;
  const int c0 = blockIdx.x * 64, i0 = blockIdx.y * 32;
  for (int r = threadIdx.y; r < 64; r += 8) {
    const i64 row = static_cast<i64>(s) * L + c0 + r;
    const int i = i0 + threadIdx.x;
    tile[r][threadIdx.x] = (row < rows && i < N) ? Y[row * ld + i] : 0.f;
  }
  __syncthreads();
  // a thread writes the pair (c, c + 1) of one output row: 128 bytes per warp and store
  for (int r = threadIdx.y; r < 32; r += 8) {
    const int i = i0 + r, c = c0 + 2 * threadIdx.x;
    if (i >= N) continue;
    uint16_t h0, l0, h1, l1;
    split_bf16(tile[2 * threadIdx.x][r], h0, l0);
    split_bf16(tile[2 * threadIdx.x + 1][r], h1, l1);
    const uint32_t H = h0 | (uint32_t(h1) << 16), Lo = l0 | (uint32_t(l1) << 16);
    uint16_t* o = out + (static_cast<i64>(s) * pitch + i) * (3 * L) + c;
    *reinterpret_cast<uint32_t*>(o) = H;
    *reinterpret_cast<uint32_t*>(o + L) = w_role ? Lo : H;
    *reinterpret_cast<uint32_t*>(o + 2 * L) = w_role ? H : Lo;
  }
}
__global__ void tile_bias_kernel(const float* __restrict__ b, float* __restrict__ out, int N, int mod) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j < N) out[j] = b[j % mod];
}

__global__ void to_f32_kernel(const void* __restrict__ src, int fmt, float* __restrict__ dst, i64 n) {
  GRID_STRIDE(i, n) dst[i] = load_any(src, fmt, static_cast<size_t>(i));
}
__global__ void from_f32_kernel(const float* __restrict__ src, void* __restrict__ dst, int fmt, i64 n) {
  GRID_STRIDE(i, n) {
    if (fmt == 2)
      static_cast<float*>(dst)[i] = src[i];
    else
      static_cast<uint16_t*>(dst)[i] = ptx::pack1(src[i], fmt);
  }
}

// ---------------------------------------------------------------------------------------------------------------
// layout
// ---------------------------------------------------------------------------------------------------------------
// dst[p, t, c] = src[img(p), c, t] (+ dense_vec[c] | + dense_full[p, c, t]);   grid (HW/32, C/32, n), block (32, 8)
__global__ void __launch_bounds__(256)
train_nchw_to_tokens_kernel(const void* __restrict__ src, int src_fmt, const int* __restrict__ img_index, const void* __restrict__ dense_vec,
                            const void* __restrict__ dense_full, int dense_fmt, float* __restrict__ dst, int C, int HW) {
  __shared__ float tile[32][33];
  const int p = blockIdx.z;
  const int img = img_index ? img_index[p] : 0;
  const int t0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  for (int i = threadIdx.y; i < 32; i += 8) {
    const int c = c0 + i, t = t0 + threadIdx.x;
    float v = load_any(src, src_fmt, (static_cast<size_t>(img) * C + c) * HW + t);
    if (dense_vec) v += load_any(dense_vec, dense_fmt, c);
    if (dense_full) v += load_any(dense_full, dense_fmt, (static_cast<size_t>(p) * C + c) * HW + t);
    tile[i][threadIdx.x] = v;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += 8) {
    const int t = t0 + i, c = c0 + threadIdx.x;
    dst[(static_cast<size_t>(p) * HW + t) * C + c] = tile[threadIdx.x][i];
  }
}

// tokens[p, t, :] = iou_token | mask_tokens[t-1] | sparse[p, t-1-nm]   (mask_decoder.py:137-141)
__global__ void tok_assemble_kernel(const float* __restrict__ iou_tok, const float* __restrict__ mask_tok, const float* __restrict__ sparse,
                                    float* __restrict__ out, int n, int nm, int k, int C) {
  const int T = 1 + nm + k;
  GRID_STRIDE(i, static_cast<i64>(n) * T * C) {
    const int c = static_cast<int>(i % C);
    const int t = static_cast<int>((i / C) % T);
    const int p = static_cast<int>(i / (static_cast<i64>(C) * T));
    out[i] = t == 0 ? iou_tok[c] : t <= nm ? mask_tok[(t - 1) * C + c] : sparse[(static_cast<i64>(p) * k + (t - 1 - nm)) * C + c];
  }
}
// adjoint: the learned tokens collect the sum over prompts (index order), the prompt embeddings their own rows
__global__ void tok_assemble_bwd_kernel(const float* __restrict__ dout, float* __restrict__ d_iou_tok, float* __restrict__ d_mask_tok,
                                        float* __restrict__ d_sparse, int n, int nm, int k, int C) {
  const int T = 1 + nm + k;
  GRID_STRIDE(i, static_cast<i64>(1 + nm) * C + static_cast<i64>(n) * k * C) {
    if (i < static_cast<i64>(1 + nm) * C) {
      const int c = static_cast<int>(i % C), t = static_cast<int>(i / C);
      float s = 0.f;
      for (int p = 0; p < n; ++p) s += dout[(static_cast<i64>(p) * T + t) * C + c];
      if (t == 0) d_iou_tok[c] += s; else d_mask_tok[(t - 1) * C + c] += s;
    } else {
      const i64 r = i - static_cast<i64>(1 + nm) * C;
      const int c = static_cast<int>(r % C);
      const int j = static_cast<int>((r / C) % k);
      const int p = static_cast<int>(r / (static_cast<i64>(C) * k));
      d_sparse[r] += dout[(static_cast<i64>(p) * T + 1 + nm + j) * C + c];
    }
  }
}

// mask rows [n, pos = ((pixel*4 + sub)*4 + s2), nm]  <->  masks [n, nm, 4g, 4g]  (two stride-2 ConvTranspose2d:
// sub = (dy,dx) of the first, s2 = (ey,ex) of the second; Y = 4y + 2dy + ey, X = 4x + 2dx + ex)
__global__ void mask_rows_kernel(float* __restrict__ rows, float* __restrict__ masks, int n, int nm, int g, int to_rows) {
  const int G4 = 4 * g;
  const i64 per = static_cast<i64>(G4) * G4;
  GRID_STRIDE(i, static_cast<i64>(n) * per * nm) {
    const int kk = static_cast<int>(i % nm);
    const i64 pos = (i / nm) % per;
    const int p = static_cast<int>(i / (per * nm));
    const int s2 = static_cast<int>(pos & 3), sub = static_cast<int>((pos >> 2) & 3);
    const int pix = static_cast<int>(pos >> 4);
    const int y = pix / g, x = pix - y * g;
    const int Y = 4 * y + 2 * (sub >> 1) + (s2 >> 1), X = 4 * x + 2 * (sub & 1) + (s2 & 1);
    const i64 m = ((static_cast<i64>(p) * nm + kk) * G4 + Y) * G4 + X;
    if (to_rows) rows[i] = masks[m]; else masks[m] = rows[i];
  }
}

// ---------------------------------------------------------------------------------------------------------------
// Adjoint of Sam.postprocess_masks (sam.py:159-172): two bilinear resamplings (align_corners=False) with a crop
// between them.  Scatter form with float atomics, as PyTorch's own upsample_bilinear2d backward.
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void tap(int o, float scale, int in_size, int& i0, int& i1, float& l1) {
  // area_pixel_compute_source_index: max(0, scale * (o + 0.5) - 0.5)
  const float src = fmaxf(scale * (o + 0.5f) - 0.5f, 0.f);
  i0 = static_cast<int>(src);
  i1 = i0 + (i0 < in_size - 1 ? 1 : 0);
  l1 = src - i0;
}
// d_in[m, iy, ix] += taps * d_out[m, oy, ox];   d_out [maps, out_h, out_w], d_in [maps, in_h, in_w], source index
// = max(0, scale * (o + 0.5) - 0.5) as area_pixel_compute_source_index does for align_corners=False
__global__ void bilinear_bwd_kernel(const float* __restrict__ d_out, float* __restrict__ d_in, int maps, int out_h, int out_w, int in_h,
                                    int in_w, float sy, float sx) {
  GRID_STRIDE(i, static_cast<i64>(maps) * out_h * out_w) {
    const int ox = static_cast<int>(i % out_w);
    const int oy = static_cast<int>((i / out_w) % out_h);
    const i64 m = i / (static_cast<i64>(out_w) * out_h);
    int y0, y1, x0, x1;
    float ly, lx;
    tap(oy, sy, in_h, y0, y1, ly);
    tap(ox, sx, in_w, x0, x1, lx);
    const float d = d_out[i];
    float* base = d_in + m * in_h * in_w;
    atomicAdd(base + static_cast<i64>(y0) * in_w + x0, (1.f - ly) * (1.f - lx) * d);
    atomicAdd(base + static_cast<i64>(y0) * in_w + x1, (1.f - ly) * lx * d);
    atomicAdd(base + static_cast<i64>(y1) * in_w + x0, ly * (1.f - lx) * d);
    atomicAdd(base + static_cast<i64>(y1) * in_w + x1, ly * lx * d);
  }
}

// ---------------------------------------------------------------------------------------------------------------
// Host: tensors, tape, ops
// ---------------------------------------------------------------------------------------------------------------
struct Ten {
  float* p = nullptr;   // values
  float* g = nullptr;   // gradient (null: none wanted)
  i64 rows = 0;
  int cols = 0;
  i64 ld = 0;
  i64 numel() const { return rows * cols; }
};

inline unsigned blocks_for(i64 n) { return static_cast<unsigned>(std::min<i64>((n + 255) / 256, 148 * 16)); }
inline size_t align256(size_t x) { return (x + 255) & ~static_cast<size_t>(255); }

struct Tape {
  SamDecoderShape s{};
  int n = 0, k = 0;
  cudaStream_t st = nullptr;
  bool dry = true;
  uint8_t* base = nullptr;
  size_t act_cap = 0, grad_cap = 0, scratch_cap = 0;   // region sizes (from the dry pass)
  size_t act_off = 0, grad_off = 0, scratch_off = 0, scratch_peak = 0;
  std::vector<std::function<int()>> back;
  std::vector<int> back_group;   // 0: always; 1 + i: hypernetwork MLP of mask token i; 100: IoU head (skipped when no
  int group = 0;                 // gradient reaches them: multimask_output selects a sub-range of the mask tokens)
  Ten masks, iou, sparse;
  float* gblob = nullptr;   // gradient of the weight blob (inside the gradient region)
  size_t blob_elems = 0;
  bool overflow = false;   // an op asked for more scratch than the sizing pass reserved
  bool use_tc = getenv("SAM_TRAIN_SIMT") == nullptr;   // image-side products (>= 8192 rows) on the split-bf16 tcgen05 GEMM; SAM_TRAIN_SIMT=1 turns it off
  bool tape_on = true;     // false: plain forward (no gradient buffers, no adjoints) -- the inference path for T > 16

  Ten alloc(i64 rows, int cols, bool grad = true) {
    Ten t;
    t.rows = rows; t.cols = cols; t.ld = cols;
    const size_t bytes = align256(static_cast<size_t>(rows) * cols * 4);
    if (!dry) t.p = reinterpret_cast<float*>(base + act_off);
    act_off += bytes;
    if (grad && tape_on) {
      if (!dry) t.g = reinterpret_cast<float*>(base + act_cap + grad_off);
      grad_off += bytes;
    }
    return t;
  }
  // op-local temporary (valid until the next scratch_reset)
  float* scratch(size_t bytes) {
    float* p = dry ? nullptr : reinterpret_cast<float*>(base + act_cap + grad_cap + scratch_off);
    scratch_off += align256(bytes);
    scratch_peak = std::max(scratch_peak, scratch_off);
    if (!dry && scratch_cap && scratch_off > scratch_cap) overflow = true;
    return p;
  }
  void scratch_reset() { scratch_off = 0; }

  // ---- launches (no-ops in the dry pass, which only sizes the regions) ----
  int gemm(Gemm g) {
    if (g.nbatch <= 0) { g.nbatch = 1; }
    if (g.nb2 <= 0) g.nb2 = 1;
    if (g.bias_mod <= 0) g.bias_mod = g.N;
    // few output tiles + a long reduction (weight gradients, attention over 4096 pixels): split K so that ~4 CTAs per SM
    // exist (a CTA's k-step is latency-bound: one tile of prefetch); the partials are folded in index order (deterministic)
    // tall and shallow (or, transposed, wide and shallow): the one-thread-per-row kernel
    if (!g.bias && g.K <= 32 && g.K >= 1 && ((g.N <= 32 && g.M >= 1024) || (g.M <= 32 && g.N >= 1024))) {
      if (g.M <= 32 && g.N >= 1024 && !(g.N <= 32 && g.M >= 1024)) {
        // C^T = B^T . A^T: swap the operands and the roles of the output's strides
        Gemm t2 = g;
        t2.A = g.B; t2.ars = g.bcs; t2.acs = g.brs; t2.a_b1 = g.b_b1; t2.a_b2 = g.b_b2;
        t2.B = g.A; t2.brs = g.acs; t2.bcs = g.ars; t2.b_b1 = g.a_b1; t2.b_b2 = g.a_b2;
        t2.M = g.N; t2.N = g.M;
        t2.crs = g.ccs ? g.ccs : 1; t2.ccs = g.crs;
        g = t2;
      }
      {
        g.splits = 1;
        g.kchunk = g.K;
        if (dry || g.M == 0 || g.N == 0) return 0;
        SAM_REQUIRE(g.nbatch <= 65535, "decoder training: gemm grid too large");
        samhost::LaunchScope scope(samhost::KC_DECODER, st, 2.0 * g.M * g.N * g.K * g.nbatch);
        const dim3 grid((g.M + 255) / 256, g.nbatch);
        if (g.N <= 8)
          sgemm_rows_kernel<8><<<grid, 256, 0, st>>>(g);
        else if (g.N <= 16)
          sgemm_rows_kernel<16><<<grid, 256, 0, st>>>(g);
        else
          sgemm_rows_kernel<32><<<grid, 256, 0, st>>>(g);
        SAM_CHECK_CUDA(cudaGetLastError());
        return 0;
      }
    }
    const bool thin = g.M <= HM && g.N <= HN && g.K >= 256;
    const int tm = thin ? HM : GT, tn = thin ? HN : GT, tk = thin ? HK : GK;
    const i64 tiles = static_cast<i64>((g.M + tm - 1) / tm) * ((g.N + tn - 1) / tn) * g.nbatch;
    g.splits = 1;
    g.kchunk = g.K;
    if (tiles < 148 && g.K >= 256) {
      int want = static_cast<int>(std::min<i64>((592 + tiles - 1) / tiles, g.K / (4 * tk) > 0 ? g.K / (4 * tk) : 1));
      if (want > 1) {
        g.kchunk = (((g.K + want - 1) / want) + tk - 1) / tk * tk;
        g.splits = (g.K + g.kchunk - 1) / g.kchunk;
      }
    }
    if (g.splits > 1) g.partial = scratch(static_cast<size_t>(g.splits) * g.nbatch * g.M * g.N * 4);
    if (dry || g.M == 0 || g.N == 0) return 0;
    SAM_REQUIRE(!overflow, "decoder training: scratch region too small (internal sizing error)");
    SAM_REQUIRE(static_cast<i64>(g.nbatch) * g.splits <= 65535 && (g.N + tn - 1) / tn <= 65535, "decoder training: gemm grid too large");
    {
      samhost::LaunchScope scope(samhost::KC_DECODER, st, 2.0 * g.M * g.N * g.K * g.nbatch);
      dim3 grid((g.M + tm - 1) / tm, (g.N + tn - 1) / tn, g.nbatch * g.splits);
      if (thin)
        sgemm_kernel<HM, HN, HK><<<grid, 256, 0, st>>>(g);
      else
        sgemm_kernel<GT, GT, GK><<<grid, 256, 0, st>>>(g);
      SAM_CHECK_CUDA(cudaGetLastError());
    }
    if (g.splits > 1) {
      samhost::LaunchScope scope(samhost::KC_DECODER, st);
      splitk_reduce_kernel<<<blocks_for(static_cast<i64>(g.M) * g.N * g.nbatch), 256, 0, st>>>(g);
      SAM_CHECK_CUDA(cudaGetLastError());
    }
    return 0;
  }
  // Image-side products on the tensor cores: C[M, N] (+)= A[M, K] . op(W) with both operands split into bf16 pairs
  // (fp32-accurate: the dropped lo.lo term is 2^-18 relative) and run by the tcgen05 GEMM of the inference path.
  //   transposed = false: C = A . W^T + bias   (W [N, K]);   transposed = true: C += A . W   (W [K, N] i.e. dX = dY . W)
  static bool tc_eligible(i64 M, int N, int K, i64 lda, i64 ldc) {
    return M >= 8192 && M <= 0x7fffffff && K % 64 == 0 && N % 8 == 0 && lda == K && ldc == N;
  }
  int gemm_tc(const float* A, i64 M, int K, const float* W, int N, bool transposed, const float* bias, int bias_mod, float* Cout,
              bool accumulate) {
    // scratch is claimed in the dry pass too (sizing)
    uint16_t* A3 = reinterpret_cast<uint16_t*>(scratch(static_cast<size_t>(M) * 3 * K * 2));
    uint16_t* W3 = reinterpret_cast<uint16_t*>(scratch(static_cast<size_t>(N) * 3 * K * 2));
    float* btile = (bias && bias_mod != N) ? scratch(static_cast<size_t>(N) * 4) : nullptr;
    if (dry) return 0;
    SAM_REQUIRE(!overflow, "decoder training: scratch region too small (internal sizing error)");
    {
      samhost::LaunchScope scope(samhost::KC_DECODER, st, 0.0, static_cast<double>(M) * K * 10.0, 2);
      split_act_kernel<<<blocks_for(M * K / 4), 256, 0, st>>>(A, A3, M, K);
      // the GEMM wants W as [N, K] rows: for dX = dY . W that is W^T
      split_weight_kernel<<<blocks_for(static_cast<i64>(N) * K), 256, 0, st>>>(W, W3, transposed ? K : N, transposed ? N : K,
                                                                             transposed ? 1 : 0);
      SAM_CHECK_CUDA(cudaGetLastError());
    }
    if (btile) {
      samhost::LaunchScope scope(samhost::KC_DECODER, st);
      tile_bias_kernel<<<(N + 255) / 256, 256, 0, st>>>(bias, btile, N, bias_mod);
      SAM_CHECK_CUDA(cudaGetLastError());
      bias = btile;
    }
    samhost::ClassOverride as_decoder(samhost::KC_DECODER);
    GemmEpilogue ep{Cout, N, SAM_F32, bias, 0, accumulate ? Cout : nullptr, accumulate ? N : 0, accumulate ? static_cast<int>(M) : 0};
    return samk_gemm(A3, 3 * K, W3, 3 * K, static_cast<int>(M), N, 3 * K, SAM_BF16, ep, st);
  }
  // Weight gradient dW[N, K] += dY[rows, N]^T . X[rows, K] on the tensor cores: the rows are cut into S chunks of L, both
  // operands are written transposed / split / chunk-major, ONE block-diagonal tcgen05 GEMM produces the S partial
  // products [S, 256, K] and splitk_reduce_kernel folds them in chunk order (deterministic).
  static void dw_plan(i64 rows, int* L, int* S) {
    i64 l = (rows + 255) / 256;             // at most 256 chunks
    l = std::max<i64>(rows >= 16384 ? 512 : 256, (l + 63) / 64 * 64);
    *L = static_cast<int>(l);
    *S = static_cast<int>((rows + l - 1) / l);
  }
  static bool dw_eligible(i64 rows, int N, int K, i64 ldy, i64 ldx) {
    return rows >= 8192 && N <= 256 && K <= 256 && N % 8 == 0 && K % 8 == 0 && ldy == N && ldx == K;
  }
  static size_t dw_scratch_bytes(i64 rows, int K) {
    int L, S;
    dw_plan(rows, &L, &S);
    return align256(static_cast<size_t>(S) * 256 * 3 * L * 2) + align256(static_cast<size_t>(S) * K * 3 * L * 2 + 256 * 3 * L * 2) +
           align256(static_cast<size_t>(S) * 256 * K * 4);
  }
  int gemm_tc_dw(const float* dY, const float* X, i64 rows, int N, int K, float* dW) {
    int L, S;
    dw_plan(rows, &L, &S);
    uint16_t* A3 = reinterpret_cast<uint16_t*>(scratch(static_cast<size_t>(S) * 256 * 3 * L * 2));
    // the second CTA of a pair reads W rows [128, 256) of a chunk: with K < 256 those are the next chunk's rows or, for
    // the last chunk, the 256 rows of slack behind the buffer (their products land in output columns >= K: clipped)
    uint16_t* W3 = reinterpret_cast<uint16_t*>(scratch(static_cast<size_t>(S) * K * 3 * L * 2 + 256 * 3 * L * 2));
    float* part = scratch(static_cast<size_t>(S) * 256 * K * 4);
    if (dry) return 0;
    SAM_REQUIRE(!overflow, "decoder training: scratch region too small (internal sizing error)");
    {
      samhost::LaunchScope scope(samhost::KC_DECODER, st, 0.0, static_cast<double>(rows) * (N + K) * 10.0, 2);
      split_transposed_kernel<<<dim3(L / 64, (N + 31) / 32, S), dim3(32, 8), 0, st>>>(dY, N, rows, N, L, 256, 0, A3);
      split_transposed_kernel<<<dim3(L / 64, (K + 31) / 32, S), dim3(32, 8), 0, st>>>(X, K, rows, K, L, K, 1, W3);
      SAM_CHECK_CUDA(cudaGetLastError());
    }
    {
      samhost::ClassOverride as_decoder(samhost::KC_DECODER);
      GemmEpilogue ep{part, K, SAM_F32, nullptr, 0, nullptr, 0, 0};
      ep.diag_mt = 1;
      ep.diag_wrows = K;
      ep.diag_wtotal = S * K;
      if (int rc = samk_gemm(A3, 3 * L, W3, 3 * L, S * 256, K, 3 * L, SAM_BF16, ep, st)) return rc;
    }
    // partial[s][i < 256][k]: rows i >= N of a chunk are the products of the unwritten operand rows -- never read
    Gemm g{};
    g.C = dW; g.crs = K; g.accumulate = 1; g.alpha = 1.f;
    g.M = N; g.N = K; g.nbatch = 1; g.nb2 = 1; g.splits = S; g.bias_mod = K;
    g.partial = part;
    g.part_rows = 256;
    samhost::LaunchScope scope(samhost::KC_DECODER, st);
    splitk_reduce_kernel<<<blocks_for(static_cast<i64>(N) * K), 256, 0, st>>>(g);
    SAM_CHECK_CUDA(cudaGetLastError());
    return 0;
  }
  // an op's adjoint will need this much scratch: registered at forward time so that the sizing pass sees it
  void need_bwd(size_t bytes) { scratch_peak = std::max(scratch_peak, align256(bytes) + 4096); }
  // out[j % mod] += sum_r X[r, j]
  int colsum(const float* X, i64 ld, i64 R, int N, int mod, float* out) {
    const int chunks = static_cast<int>((R + kColRows - 1) / kColRows);
    float* part = scratch(static_cast<size_t>(chunks) * N * 4);
    if (dry || R == 0) return 0;
    SAM_REQUIRE(!overflow, "decoder training: scratch region too small (internal sizing error)");
    if (chunks == 1 && mod == N) {
      samhost::LaunchScope scope(samhost::KC_DECODER, st);
      colsum_small_kernel<<<(N + 31) / 32, dim3(32, 8), 0, st>>>(X, ld, R, N, out);
      SAM_CHECK_CUDA(cudaGetLastError());
      return 0;
    }
    SAM_REQUIRE(chunks <= 65535, "decoder training: colsum grid too large");
    {
      samhost::LaunchScope scope(samhost::KC_DECODER, st);
      colsum_partial_kernel<<<dim3((N + 31) / 32, chunks), dim3(32, 8), 0, st>>>(X, ld, R, N, part);
      SAM_CHECK_CUDA(cudaGetLastError());
    }
    {
      samhost::LaunchScope scope(samhost::KC_DECODER, st);
      colsum_finish_kernel<<<(mod + 7) / 8, 256, 0, st>>>(part, chunks, N, mod, out);
      SAM_CHECK_CUDA(cudaGetLastError());
    }
    return 0;
  }
  template <typename F>
  int ew(i64 n, F&& launch) {
    if (dry || n == 0) return 0;
    samhost::LaunchScope scope(samhost::KC_DECODER, st);
    launch(blocks_for(n));
    SAM_CHECK_CUDA(cudaGetLastError());
    return 0;
  }
  void push(std::function<int()> f) {
    if (dry || !tape_on) return;
    back.push_back(std::move(f));
    back_group.push_back(group);
  }
};

#define TRY(expr)               \
  do {                          \
    if (int _rc = (expr)) return _rc; \
  } while (0)

// ---- ops: forward now, adjoint on the tape ----

// Y = X . W^T + b      W [N, K] (parameter), bias [bias_mod] broadcast as b[j % bias_mod];  `out` may be a strided view
int linear(Tape& t, const Ten& X, const Ten& W, const Ten& b, Ten* Y, int bias_mod = 0, const Ten* out = nullptr) {
  const int N = static_cast<int>(W.rows), K = W.cols;
  if (X.cols != K) return samhost::set_error(1, "decoder training: linear K mismatch (%d vs %d)", X.cols, K);
  *Y = out ? *out : t.alloc(X.rows, N);
  const int M = static_cast<int>(X.rows);
  const int bm = bias_mod ? bias_mod : N;
  const bool tc_fwd = t.use_tc && Tape::tc_eligible(X.rows, N, K, X.ld, Y->ld);
  if (tc_fwd) {
    TRY(t.gemm_tc(X.p, X.rows, K, W.p, N, false, b.p, bm, Y->p, false));
  } else {
    Gemm g{};
    g.A = X.p; g.ars = X.ld; g.acs = 1;
    g.B = W.p; g.brs = 1; g.bcs = K;
    g.C = Y->p; g.crs = Y->ld;
    g.bias = b.p; g.bias_mod = bm;
    g.M = M; g.N = N; g.K = K; g.alpha = 1.f;
    if (t.dry) g.bias = nullptr;
    TRY(t.gemm(g));
  }
  t.scratch_reset();
  const Ten Yc = *Y;
  if (t.use_tc && t.tape_on) {   // scratch of the adjoint's tensor-core products (the sizing pass only runs the forward)
    size_t need = 0;
    if (Tape::tc_eligible(M, K, N, Yc.ld, X.ld))
      need = align256(static_cast<size_t>(M) * 3 * N * 2) + align256(static_cast<size_t>(K) * 3 * N * 2);
    if (Tape::dw_eligible(M, N, K, Yc.ld, X.ld)) need = std::max(need, Tape::dw_scratch_bytes(M, K));
    // + the column-sum partials of the bias gradient, claimed after them in the same adjoint
    t.need_bwd(need + align256(static_cast<size_t>((static_cast<i64>(M) * (N / bm) + kColRows - 1) / kColRows) * N * 4));
  }
  Tape* tp = &t;
  t.push([tp, X, W, b, Yc, M, N, K, bm]() -> int {
    Tape& t = *tp;
    if (X.g && t.use_tc && Tape::tc_eligible(M, K, N, Yc.ld, X.ld)) {   // dX += dY . W on the tensor cores
      TRY(t.gemm_tc(Yc.g, M, N, W.p, K, true, nullptr, 0, X.g, true));
    } else if (X.g) {   // dX += dY . W
      Gemm g{};
      g.A = Yc.g; g.ars = Yc.ld; g.acs = 1;
      g.B = W.p; g.brs = K; g.bcs = 1;
      g.C = X.g; g.crs = X.ld; g.accumulate = 1;
      g.M = M; g.N = K; g.K = N; g.alpha = 1.f;
      TRY(t.gemm(g));
    }
    t.scratch_reset();
    if (W.g && t.use_tc && Tape::dw_eligible(M, N, K, Yc.ld, X.ld)) {   // dW += dY^T . X on the tensor cores
      TRY(t.gemm_tc_dw(Yc.g, X.p, M, N, K, W.g));
    } else if (W.g) {   // dW += dY^T . X   (reduction over the rows: split-K)
      Gemm g{};
      g.A = Yc.g; g.ars = 1; g.acs = Yc.ld;
      g.B = X.p; g.brs = X.ld; g.bcs = 1;
      g.C = W.g; g.crs = K; g.accumulate = 1;
      g.M = N; g.N = K; g.K = M; g.alpha = 1.f;
      TRY(t.gemm(g));
    }
    if (b.g) {
      if (bm == N) {
        TRY(t.colsum(Yc.g, Yc.ld, M, N, N, b.g));
      } else {   // bias shared by N / bm column groups: rows of the [M * N / bm, bm] view (needs a contiguous Y)
        TRY(t.colsum(Yc.g, bm, static_cast<i64>(M) * (N / bm), bm, bm, b.g));
      }
    }
    t.scratch_reset();
    return 0;
  });
  return 0;
}

// C = A + B[i % bmod]    (B may be a constant broadcast over the leading dimension)
int add(Tape& t, const Ten& A, const Ten& B, Ten* C) {
  *C = t.alloc(A.rows, A.cols);
  const i64 n = A.numel(), bmod = B.numel();
  const Ten Cc = *C;
  TRY(t.ew(n, [&](unsigned nb) { add_kernel<<<nb, 256, 0, t.st>>>(A.p, B.p, Cc.p, n, bmod); }));
  Tape* tp = &t;
  t.push([tp, A, B, Cc, n, bmod]() -> int {
    Tape& t = *tp;
    if (A.g) TRY(t.ew(n, [&](unsigned nb) { acc_kernel<<<nb, 256, 0, t.st>>>(A.g, Cc.g, n); }));
    if (B.g) {
      if (bmod != n) return samhost::set_error(1, "decoder training: gradient of a broadcast addend is not supported");
      TRY(t.ew(n, [&](unsigned nb) { acc_kernel<<<nb, 256, 0, t.st>>>(B.g, Cc.g, n); }));
    }
    return 0;
  });
  return 0;
}

int layernorm(Tape& t, const Ten& X, const Ten& gw, const Ten& gb, float eps, Ten* Y) {
  const int C = X.cols;
  const i64 R = X.rows;
  if (C % 32 != 0 || C > 32 * LNV) return samhost::set_error(1, "decoder training: LayerNorm width %d", C);
  *Y = t.alloc(R, C);
  float2* stats = reinterpret_cast<float2*>(t.alloc(R, 2, false).p);
  const Ten Yc = *Y;
  if (!t.dry) {
    samhost::LaunchScope scope(samhost::KC_DECODER, t.st);
    ln_fwd_kernel<<<static_cast<unsigned>((R + 7) / 8), 256, 0, t.st>>>(X.p, gw.p, gb.p, Yc.p, stats, R, C, eps);
    SAM_CHECK_CUDA(cudaGetLastError());
  }
  Tape* tp = &t;
  t.push([tp, X, gw, gb, Yc, stats, R, C]() -> int {
    Tape& t = *tp;
    const int nblk = static_cast<int>((R + kLnRowsPerBlock - 1) / kLnRowsPerBlock);
    float* part = t.scratch(static_cast<size_t>(nblk) * 2 * C * 4);
    {
      samhost::LaunchScope scope(samhost::KC_DECODER, t.st);
      ln_bwd_kernel<<<nblk, 256, 0, t.st>>>(Yc.g, X.p, stats, gw.p, X.g, part, R, C);
      SAM_CHECK_CUDA(cudaGetLastError());
    }
    // weight | bias gradients are adjacent in the blob (weight then bias): one column sum over the block partials
    if (gw.g) {
      if (gb.g != gw.g + C) return samhost::set_error(1, "decoder training: LayerNorm weight / bias gradients must be adjacent");
      TRY(t.colsum(part, 2 * C, nblk, 2 * C, 2 * C, gw.g));
    }
    t.scratch_reset();
    return 0;
  });
  return 0;
}

int relu(Tape& t, const Ten& X, Ten* Y) {
  *Y = t.alloc(X.rows, X.cols);
  const i64 n = X.numel();
  const Ten Yc = *Y;
  TRY(t.ew(n, [&](unsigned nb) { relu_fwd_kernel<<<nb, 256, 0, t.st>>>(X.p, Yc.p, n); }));
  Tape* tp = &t;
  t.push([tp, X, Yc, n]() -> int {
    Tape& t = *tp;
    return t.ew(n, [&](unsigned nb) { relu_bwd_kernel<<<nb, 256, 0, t.st>>>(Yc.g, Yc.p, X.g, n); });
  });
  return 0;
}
int gelu(Tape& t, const Ten& X, Ten* Y) {
  *Y = t.alloc(X.rows, X.cols);
  const i64 n = X.numel();
  const Ten Yc = *Y;
  TRY(t.ew(n, [&](unsigned nb) { gelu_fwd_kernel<<<nb, 256, 0, t.st>>>(X.p, Yc.p, n); }));
  Tape* tp = &t;
  t.push([tp, X, Yc, n]() -> int {
    Tape& t = *tp;
    return t.ew(n, [&](unsigned nb) { gelu_bwd_kernel<<<nb, 256, 0, t.st>>>(Yc.g, X.p, X.g, n); });
  });
  return 0;
}

// softmax(q k^T / sqrt(dh)) v per (prompt, head)   (transformer.py:232-239);  q [n*Nq, D], k / v [n*Nk, D] -> o [n*Nq, D]
int attention_core(Tape& t, const Ten& q, const Ten& k, const Ten& v, int n, int Nq, int Nk, int heads, Ten* o) {
  const int D = q.cols, dh = D / heads;
  const float scale = 1.0f / sqrtf(static_cast<float>(dh));
  *o = t.alloc(static_cast<i64>(n) * Nq, D);
  Ten P = t.alloc(static_cast<i64>(n) * heads * Nq, Nk, false);
  const Ten oc = *o;
  auto batch = [&](Gemm& g) { g.nbatch = n * heads; g.nb2 = heads; };
  {
    Gemm g{};   // S = scale * Q K^T
    batch(g);
    g.A = q.p; g.ars = q.ld; g.acs = 1; g.a_b1 = static_cast<i64>(Nq) * q.ld; g.a_b2 = dh;
    g.B = k.p; g.brs = 1; g.bcs = k.ld; g.b_b1 = static_cast<i64>(Nk) * k.ld; g.b_b2 = dh;
    g.C = P.p; g.crs = Nk; g.c_b1 = static_cast<i64>(heads) * Nq * Nk; g.c_b2 = static_cast<i64>(Nq) * Nk;
    g.M = Nq; g.N = Nk; g.K = dh; g.alpha = scale;
    TRY(t.gemm(g));
  }
  const i64 R = static_cast<i64>(n) * heads * Nq;
  if (!t.dry) {
    samhost::LaunchScope scope(samhost::KC_DECODER, t.st);
    softmax_fwd_kernel<<<static_cast<unsigned>((R + 7) / 8), 256, 0, t.st>>>(P.p, R, Nk);
    SAM_CHECK_CUDA(cudaGetLastError());
  }
  {
    Gemm g{};   // O = P V
    batch(g);
    g.A = P.p; g.ars = Nk; g.acs = 1; g.a_b1 = static_cast<i64>(heads) * Nq * Nk; g.a_b2 = static_cast<i64>(Nq) * Nk;
    g.B = v.p; g.brs = v.ld; g.bcs = 1; g.b_b1 = static_cast<i64>(Nk) * v.ld; g.b_b2 = dh;
    g.C = oc.p; g.crs = oc.ld; g.c_b1 = static_cast<i64>(Nq) * oc.ld; g.c_b2 = dh;
    g.M = Nq; g.N = dh; g.K = Nk; g.alpha = 1.f;
    TRY(t.gemm(g));
  }
  t.scratch_reset();
  Tape* tp = &t;
  t.push([tp, q, k, v, oc, P, n, Nq, Nk, heads, dh, scale, R]() -> int {
    Tape& t = *tp;
    auto batch = [&](Gemm& g) { g.nbatch = n * heads; g.nb2 = heads; };
    const i64 pb1 = static_cast<i64>(heads) * Nq * Nk, pb2 = static_cast<i64>(Nq) * Nk;
    float* dP = t.scratch(static_cast<size_t>(R) * Nk * 4);
    {
      Gemm g{};   // dP = dO V^T
      batch(g);
      g.A = oc.g; g.ars = oc.ld; g.acs = 1; g.a_b1 = static_cast<i64>(Nq) * oc.ld; g.a_b2 = dh;
      g.B = v.p; g.brs = 1; g.bcs = v.ld; g.b_b1 = static_cast<i64>(Nk) * v.ld; g.b_b2 = dh;
      g.C = dP; g.crs = Nk; g.c_b1 = pb1; g.c_b2 = pb2;
      g.M = Nq; g.N = Nk; g.K = dh; g.alpha = 1.f;
      TRY(t.gemm(g));
    }
    if (v.g) {
      Gemm g{};   // dV += P^T dO
      batch(g);
      g.A = P.p; g.ars = 1; g.acs = Nk; g.a_b1 = pb1; g.a_b2 = pb2;
      g.B = oc.g; g.brs = oc.ld; g.bcs = 1; g.b_b1 = static_cast<i64>(Nq) * oc.ld; g.b_b2 = dh;
      g.C = v.g; g.crs = v.ld; g.c_b1 = static_cast<i64>(Nk) * v.ld; g.c_b2 = dh; g.accumulate = 1;
      g.M = Nk; g.N = dh; g.K = Nq; g.alpha = 1.f;
      TRY(t.gemm(g));
    }
    {
      samhost::LaunchScope scope(samhost::KC_DECODER, t.st);
      softmax_bwd_kernel<<<static_cast<unsigned>((R + 7) / 8), 256, 0, t.st>>>(dP, P.p, R, Nk);
      SAM_CHECK_CUDA(cudaGetLastError());
    }
    if (q.g) {
      Gemm g{};   // dQ += scale * dS K
      batch(g);
      g.A = dP; g.ars = Nk; g.acs = 1; g.a_b1 = pb1; g.a_b2 = pb2;
      g.B = k.p; g.brs = k.ld; g.bcs = 1; g.b_b1 = static_cast<i64>(Nk) * k.ld; g.b_b2 = dh;
      g.C = q.g; g.crs = q.ld; g.c_b1 = static_cast<i64>(Nq) * q.ld; g.c_b2 = dh; g.accumulate = 1;
      g.M = Nq; g.N = dh; g.K = Nk; g.alpha = scale;
      TRY(t.gemm(g));
    }
    if (k.g) {
      Gemm g{};   // dK += scale * dS^T Q
      batch(g);
      g.A = dP; g.ars = 1; g.acs = Nk; g.a_b1 = pb1; g.a_b2 = pb2;
      g.B = q.p; g.brs = q.ld; g.bcs = 1; g.b_b1 = static_cast<i64>(Nq) * q.ld; g.b_b2 = dh;
      g.C = k.g; g.crs = k.ld; g.c_b1 = static_cast<i64>(Nk) * k.ld; g.c_b2 = dh; g.accumulate = 1;
      g.M = Nk; g.N = dh; g.K = Nq; g.alpha = scale;
      TRY(t.gemm(g));
    }
    t.scratch_reset();
    return 0;
  });
  return 0;
}

// parameters: pointers into the weight blob and, at the same offset, into its gradient
struct ParamWalk {
  const float* blob;
  float* gblob;
  size_t off = 0;
  Ten take(i64 rows, int cols) {
    Ten t;
    t.p = const_cast<float*>(blob) + off;
    t.g = gblob ? gblob + off : nullptr;
    t.rows = rows; t.cols = cols; t.ld = cols;
    off += static_cast<size_t>(rows) * cols;
    return t;
  }
};
struct AttnP { Ten qw, qb, kw, kb, vw, vb, ow, ob; };
struct Mlp3P { Ten w0, b0, w1, b1, w2, b2; };
struct LayerP {
  AttnP self_attn, t2i, i2t;
  Ten n1w, n1b, n2w, n2b, n3w, n3b, n4w, n4b, l1w, l1b, l2w, l2b;
};
struct Params {
  Ten iou_token, mask_tokens;
  LayerP layer[8];
  AttnP final_attn;
  Ten nfw, nfb, up0w, up0b, upln_w, upln_b, up1w, up1b;
  Mlp3P hyper[4], iou_head;
  size_t total;
};
// same order as csrc/decoder.cu carve_weights / _pack.pack_decoder (== state_dict order of mask_decoder.*)
void carve_params(const SamDecoderShape& s, const float* blob, float* gblob, Params* P) {
  const int C = s.C, Ci = C / 2, H = s.mlp_dim, nm = s.num_mask_tokens;
  ParamWalk w{blob, gblob};
  auto attn = [&](AttnP& a, int inner) {
    a.qw = w.take(inner, C); a.qb = w.take(1, inner);
    a.kw = w.take(inner, C); a.kb = w.take(1, inner);
    a.vw = w.take(inner, C); a.vb = w.take(1, inner);
    a.ow = w.take(C, inner); a.ob = w.take(1, C);
  };
  P->iou_token = w.take(1, C);
  P->mask_tokens = w.take(nm, C);
  for (int l = 0; l < s.depth; ++l) {
    LayerP& L = P->layer[l];
    attn(L.self_attn, C);
    L.n1w = w.take(1, C); L.n1b = w.take(1, C);
    attn(L.t2i, Ci);
    L.n2w = w.take(1, C); L.n2b = w.take(1, C);
    L.l1w = w.take(H, C); L.l1b = w.take(1, H);
    L.l2w = w.take(C, H); L.l2b = w.take(1, C);
    L.n3w = w.take(1, C); L.n3b = w.take(1, C);
    L.n4w = w.take(1, C); L.n4b = w.take(1, C);
    attn(L.i2t, Ci);
  }
  attn(P->final_attn, Ci);
  P->nfw = w.take(1, C); P->nfb = w.take(1, C);
  const int C1 = C / 4, C2 = C / 8;
  P->up0w = w.take(4 * C1, C); P->up0b = w.take(1, 4 * C1);
  P->upln_w = w.take(1, C1); P->upln_b = w.take(1, C1);
  P->up1w = w.take(4 * C2, C1); P->up1b = w.take(1, C2);
  for (int i = 0; i < nm; ++i) {
    Mlp3P& m = P->hyper[i];
    m.w0 = w.take(C, C); m.b0 = w.take(1, C);
    m.w1 = w.take(C, C); m.b1 = w.take(1, C);
    m.w2 = w.take(C2, C); m.b2 = w.take(1, C2);
  }
  const int Hi = s.iou_hidden;
  Mlp3P& h = P->iou_head;
  h.w0 = w.take(Hi, C); h.b0 = w.take(1, Hi);
  h.w1 = w.take(Hi, Hi); h.b1 = w.take(1, Hi);
  h.w2 = w.take(nm, Hi); h.b2 = w.take(1, nm);
  P->total = w.off;
}

// Attention module (transformer.py:185-242): projections + attention_core + out_proj
int attention(Tape& t, const AttnP& a, const Ten& xq, const Ten& xk, const Ten& xv, int n, int Nq, int Nk, int heads, Ten* out) {
  Ten q, k, v, o;
  TRY(linear(t, xq, a.qw, a.qb, &q));
  TRY(linear(t, xk, a.kw, a.kb, &k));
  TRY(linear(t, xv, a.vw, a.vb, &v));
  TRY(attention_core(t, q, k, v, n, Nq, Nk, heads, &o));
  return linear(t, o, a.ow, a.ob, out);
}
int mlp3(Tape& t, const Mlp3P& m, const Ten& x, Ten* y, const Ten* out = nullptr) {
  Ten h0, r0, h1, r1;
  TRY(linear(t, x, m.w0, m.b0, &h0));
  TRY(relu(t, h0, &r0));
  TRY(linear(t, r0, m.w1, m.b1, &h1));
  TRY(relu(t, h1, &r1));
  return linear(t, r1, m.w2, m.b2, y, 0, out);
}

struct TrainInputs {
  const float* blob;
  const void* emb; int emb_fmt; int n_images; const int* img_index;
  const void* sparse; int sparse_fmt;
  const void* dense_vec; const void* dense_full; int dense_fmt;
  const void* image_pe; int pe_fmt;      // [1, C, g, g], or
  const float* pe_tokens;                // the token-major fp32 copy [HW, C] (sam_decoder_prepare's pe_t)
};

// The whole forward; in the dry pass (t.dry) it only advances the allocators.
int build_forward(Tape& t, const TrainInputs& in) {
  const SamDecoderShape& s = t.s;
  const int C = s.C, nm = s.num_mask_tokens, g = s.grid, HW = g * g, n = t.n, k = t.k, T = 1 + nm + k, heads = s.heads;
  const int C1 = C / 4, C2 = C / 8;
  t.gblob = nullptr;
  if (t.tape_on) {
    t.gblob = t.dry ? nullptr : reinterpret_cast<float*>(t.base + t.act_cap + t.grad_off);
    t.grad_off += align256(t.blob_elems * 4);
  }
  Params P;
  carve_params(s, in.blob, t.gblob, &P);
  if (P.total != t.blob_elems) return samhost::set_error(1, "decoder training: weight layout mismatch");

  // inputs
  t.sparse = t.alloc(static_cast<i64>(n) * k, C);
  {
    const Ten sp = t.sparse;
    const void* src = in.sparse;
    const int fmt = in.sparse_fmt;
    TRY(t.ew(sp.numel(), [&](unsigned nb) { to_f32_kernel<<<nb, 256, 0, t.st>>>(src, fmt, sp.p, sp.numel()); }));
  }
  Ten tokens0 = t.alloc(static_cast<i64>(n) * T, C);
  {
    const Ten sp = t.sparse, iou_t = P.iou_token, mask_t = P.mask_tokens;
    TRY(t.ew(tokens0.numel(), [&](unsigned nb) {
      tok_assemble_kernel<<<nb, 256, 0, t.st>>>(iou_t.p, mask_t.p, sp.p, tokens0.p, n, nm, k, C);
    }));
    Tape* tp = &t;
    t.push([tp, tokens0, sp, iou_t, mask_t, n, nm, k, C]() -> int {
      Tape& t = *tp;
      return t.ew(static_cast<i64>(1 + nm + n * k) * C, [&](unsigned nb) {
        tok_assemble_bwd_kernel<<<nb, 256, 0, t.st>>>(tokens0.g, iou_t.g, mask_t.g, sp.g, n, nm, k, C);
      });
    });
  }
  Ten keys = t.alloc(static_cast<i64>(n) * HW, C, false);   // image embedding + dense prompt embedding: constants
  Ten pe = t.alloc(HW, C, false);
  if (!t.dry) {
    samhost::LaunchScope scope(samhost::KC_DECODER, t.st, 0.0, 0.0, 2);
    train_nchw_to_tokens_kernel<<<dim3(HW / 32, C / 32, n), dim3(32, 8), 0, t.st>>>(in.emb, in.emb_fmt, in.img_index, in.dense_vec,
                                                                                  in.dense_full, in.dense_fmt, keys.p, C, HW);
    SAM_CHECK_CUDA(cudaGetLastError());
    if (in.pe_tokens) {
      SAM_CHECK_CUDA(cudaMemcpyAsync(pe.p, in.pe_tokens, static_cast<size_t>(HW) * C * 4, cudaMemcpyDeviceToDevice, t.st));
    } else {
      train_nchw_to_tokens_kernel<<<dim3(HW / 32, C / 32, 1), dim3(32, 8), 0, t.st>>>(in.image_pe, in.pe_fmt, nullptr, nullptr, nullptr,
                                                                                    2, pe.p, C, HW);
      SAM_CHECK_CUDA(cudaGetLastError());
    }
  }

  Ten queries = tokens0;
  for (int l = 0; l < s.depth; ++l) {
    const LayerP& L = P.layer[l];
    Ten a, q, kpe, x;
    if (l == 0) {   // skip_first_layer_pe: the self-attention output replaces the queries (transformer.py:153-155)
      TRY(attention(t, L.self_attn, queries, queries, queries, n, T, T, heads, &x));
    } else {
      TRY(add(t, queries, tokens0, &q));
      TRY(attention(t, L.self_attn, q, q, queries, n, T, T, heads, &a));
      TRY(add(t, queries, a, &x));
    }
    TRY(layernorm(t, x, L.n1w, L.n1b, 1e-5f, &queries));
    TRY(add(t, queries, tokens0, &q));
    TRY(add(t, keys, pe, &kpe));
    TRY(attention(t, L.t2i, q, kpe, keys, n, T, HW, heads, &a));
    TRY(add(t, queries, a, &x));
    TRY(layernorm(t, x, L.n2w, L.n2b, 1e-5f, &queries));
    Ten h, r, m;
    TRY(linear(t, queries, L.l1w, L.l1b, &h));
    TRY(relu(t, h, &r));
    TRY(linear(t, r, L.l2w, L.l2b, &m));
    TRY(add(t, queries, m, &x));
    TRY(layernorm(t, x, L.n3w, L.n3b, 1e-5f, &queries));
    TRY(add(t, queries, tokens0, &q));
    TRY(attention(t, L.i2t, kpe, q, queries, n, HW, T, heads, &a));   // keys + key_pe is unchanged since the t2i attention
    TRY(add(t, keys, a, &x));
    TRY(layernorm(t, x, L.n4w, L.n4b, 1e-5f, &keys));
  }
  {
    Ten a, q, kpe, x;
    TRY(add(t, queries, tokens0, &q));
    TRY(add(t, keys, pe, &kpe));
    TRY(attention(t, P.final_attn, q, kpe, keys, n, T, HW, heads, &a));
    TRY(add(t, queries, a, &x));
    TRY(layernorm(t, x, P.nfw, P.nfb, 1e-5f, &queries));
  }
  const Ten hs = queries;   // [n*T, C]
  auto token_row = [&](int tok) {   // hs[:, tok, :] as a strided [n, C] view
    Ten v = hs;
    v.p = hs.p ? hs.p + static_cast<i64>(tok) * C : nullptr;
    v.g = hs.g ? hs.g + static_cast<i64>(tok) * C : nullptr;
    v.rows = n; v.ld = static_cast<i64>(T) * C;
    return v;
  };
  // output_upscaling (mask_decoder.py:53-63): both stride-2 ConvTranspose2d are per-pixel linears
  Ten U, Un, Ug, Z, Zg;
  TRY(linear(t, keys, P.up0w, P.up0b, &U));                    // [n*HW, (dy,dx,oc)]
  U.rows *= 4; U.cols = C1; U.ld = C1;                          // -> [(pixel, sub), oc]
  TRY(layernorm(t, U, P.upln_w, P.upln_b, 1e-6f, &Un));        // LayerNorm2d: over the channels of each output pixel
  TRY(gelu(t, Un, &Ug));
  TRY(linear(t, Ug, P.up1w, P.up1b, &Z, C2));                  // [(pixel, sub), (ey,ex,oc)]
  TRY(gelu(t, Z, &Zg));
  Ten up = Zg;
  up.rows *= 4; up.cols = C2; up.ld = C2;                       // [(pixel, sub, s2), oc]
  // hypernetwork MLPs (mask_decoder.py:163-169) write their rows of hyper_in [n, nm, C2]
  Ten hyper = t.alloc(n, nm * C2);
  for (int i = 0; i < nm; ++i) {
    Ten view = hyper, y;
    view.p = hyper.p ? hyper.p + i * C2 : nullptr;
    view.g = hyper.g ? hyper.g + i * C2 : nullptr;
    view.cols = C2;
    t.group = 1 + i;
    TRY(mlp3(t, P.hyper[i], token_row(1 + i), &y, &view));
    t.group = 0;
  }
  // masks = hyper_in @ upscaled (mask_decoder.py:171-174), as rows [pos, nm] per prompt, then the layout gather
  const i64 per = static_cast<i64>(HW) * 16;
  Ten mrows = t.alloc(static_cast<i64>(n) * per, nm);
  {
    Gemm gm{};
    gm.nbatch = n; gm.nb2 = 1;
    gm.A = up.p; gm.ars = C2; gm.acs = 1; gm.a_b1 = per * C2;
    gm.B = hyper.p; gm.brs = 1; gm.bcs = C2; gm.b_b1 = static_cast<i64>(nm) * C2;
    gm.C = mrows.p; gm.crs = nm; gm.c_b1 = per * nm;
    gm.M = static_cast<int>(per); gm.N = nm; gm.K = C2; gm.alpha = 1.f;
    TRY(t.gemm(gm));
    t.scratch_reset();
    Tape* tp = &t;
    t.push([tp, up, hyper, mrows, n, nm, C2, per]() -> int {
      Tape& t = *tp;
      {
        Gemm g{};   // d_up += dM . hyper
        g.nbatch = n; g.nb2 = 1;
        g.A = mrows.g; g.ars = nm; g.acs = 1; g.a_b1 = per * nm;
        g.B = hyper.p; g.brs = C2; g.bcs = 1; g.b_b1 = static_cast<i64>(nm) * C2;
        g.C = up.g; g.crs = C2; g.c_b1 = per * C2; g.accumulate = 1;
        g.M = static_cast<int>(per); g.N = C2; g.K = nm; g.alpha = 1.f;
        TRY(t.gemm(g));
      }
      {
        Gemm g{};   // d_hyper += dM^T . up  (reduction over the 65536 positions: split-K)
        g.nbatch = n; g.nb2 = 1;
        g.A = mrows.g; g.ars = 1; g.acs = nm; g.a_b1 = per * nm;
        g.B = up.p; g.brs = C2; g.bcs = 1; g.b_b1 = per * C2;
        g.C = hyper.g; g.crs = C2; g.c_b1 = static_cast<i64>(nm) * C2; g.accumulate = 1;
        g.M = nm; g.N = C2; g.K = static_cast<int>(per); g.alpha = 1.f;
        TRY(t.gemm(g));
      }
      t.scratch_reset();
      return 0;
    });
  }
  t.masks = t.alloc(static_cast<i64>(n) * nm, 16 * HW);
  {
    const Ten mk = t.masks;
    const i64 tot = mrows.numel();
    TRY(t.ew(tot, [&](unsigned nb) { mask_rows_kernel<<<nb, 256, 0, t.st>>>(mrows.p, mk.p, n, nm, g, 0); }));
    Tape* tp = &t;
    t.push([tp, mrows, mk, n, nm, g, tot]() -> int {
      Tape& t = *tp;   // mrows.g is zero here (this is its only consumer): a plain gather
      return t.ew(tot, [&](unsigned nb) { mask_rows_kernel<<<nb, 256, 0, t.st>>>(mrows.g, mk.g, n, nm, g, 1); });
    });
  }
  t.group = 100;
  TRY(mlp3(t, P.iou_head, token_row(0), &t.iou));
  t.group = 0;
  return 0;
}

size_t region_total(const Tape& t) { return t.act_cap + t.grad_cap + t.scratch_cap + 256; }

// Sizes the three regions (host only).  The backward's scratch needs are a function of the same shapes; they are
// covered by running the adjoint closures' sizing rules here: the largest are the attention dP (as big as P) and the
// split-K partials, both bounded by `bound` below.
int size_regions(Tape& t) {
  t.dry = true;
  t.act_off = t.grad_off = t.scratch_off = t.scratch_peak = 0;
  TrainInputs none{};
  TRY(build_forward(t, none));
  const SamDecoderShape& s = t.s;
  const i64 HW = static_cast<i64>(s.grid) * s.grid, T = 1 + s.num_mask_tokens + t.k;
  const i64 Tm = std::max<i64>(T, 1);
  size_t bound = 0;
  // attention dP: n * heads * Nq * Nk floats, largest with one side = HW
  bound = std::max(bound, align256(static_cast<size_t>(t.n) * s.heads * std::max(Tm * HW, Tm * Tm) * 4));
  // split-K partials: at most 592 + 148 output tiles of 64 x 64 floats (Tape::gemm)
  bound = std::max(bound, align256(static_cast<size_t>(768) * GT * GT * 4));
  const i64 rows_max = static_cast<i64>(t.n) * HW * 4;
  // column-sum partials and LayerNorm block partials: rows / 64 * 2C floats at most
  bound = std::max(bound, align256(static_cast<size_t>((rows_max + 63) / 64 + 1) * 2 * s.C * 4));
  // split-bf16 operands of the tensor-core products: rows x 3 x width x 2 bytes, widest for the gradient of the second
  // ConvTranspose2d ([n * 4 HW, 128]) and for the [n * HW, 256] tensors; plus the split weight and a tiled bias
  bound = std::max(bound, align256(static_cast<size_t>(rows_max) * 3 * (s.C / 2) * 2) + align256(static_cast<size_t>(s.C) * 3 * s.C * 2 * 2) + 4096);
  t.act_cap = align256(t.act_off);
  t.grad_cap = align256(t.grad_off);
  t.scratch_cap = 2 * std::max(bound, t.scratch_peak) + 4096;
  return 0;
}

int check_train_shape(const SamDecoderShape& s, int n, int k) {
  SAM_REQUIRE(s.C == 256 && s.C % s.heads == 0 && (s.C / 2) % s.heads == 0, "decoder training: transformer_dim must be 256 (got %d)", s.C);
  SAM_REQUIRE(s.depth >= 1 && s.depth <= 8 && s.num_mask_tokens >= 1 && s.num_mask_tokens <= 4, "decoder training: depth / mask tokens");
  SAM_REQUIRE(s.grid > 0 && (s.grid * s.grid) % 32 == 0, "decoder training: grid %d", s.grid);
  SAM_REQUIRE(n >= 1 && k >= 0, "decoder training: n = %d prompts, k = %d embeddings per prompt", n, k);
  return 0;
}

}  // namespace

size_t samk_decoder_train_workspace_bytes(const SamDecoderShape& s, int n, int k) {
  if (check_train_shape(s, n, k)) return 0;
  Tape t;
  t.s = s; t.n = n; t.k = k;
  t.blob_elems = samk_decoder_weight_elems(s);
  if (size_regions(t)) return 0;
  return region_total(t);
}

int samk_decoder_train_forward(const SamDecoderShape& s, const float* blob, const void* image_embeddings, int emb_fmt, int n_images,
                               const int* img_index, const float* sparse, int n, int k, const void* dense_vec, const void* dense_full,
                               int dense_fmt, const void* image_pe, int pe_fmt, float* masks, float* iou, void* workspace,
                               size_t workspace_bytes, void** tape_out, cudaStream_t st) {
  SAM_REQUIRE(tape_out != nullptr, "decoder training: NULL tape_out");
  *tape_out = nullptr;
  if (int rc = check_train_shape(s, n, k)) return rc;
  SAM_REQUIRE(blob && image_embeddings && image_pe && masks && iou && workspace && (k == 0 || sparse), "decoder training: NULL argument");
  SAM_REQUIRE(n_images >= 1 && (img_index || n_images == 1), "decoder training: img_index is required with several images");
  SAM_REQUIRE(!(dense_vec && dense_full), "decoder training: pass the dense embedding as a vector OR in full");
  Tape* t = new Tape();
  t->s = s; t->n = n; t->k = k; t->st = st;
  t->blob_elems = samk_decoder_weight_elems(s);
  int rc = size_regions(*t);
  if (!rc && region_total(*t) > workspace_bytes)
    rc = samhost::set_error(1, "decoder training: workspace of %zu bytes, %zu needed", workspace_bytes, region_total(*t));
  if (!rc) {
    t->dry = false;
    t->base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(workspace) + 255) & ~static_cast<uintptr_t>(255));
    t->act_off = t->grad_off = t->scratch_off = 0;
    TrainInputs in{blob, image_embeddings, emb_fmt, n_images, img_index, sparse, SAM_F32, dense_vec, dense_full, dense_fmt, image_pe,
                   pe_fmt, nullptr};
    rc = build_forward(*t, in);
  }
  if (!rc) {
    const size_t mb = static_cast<size_t>(n) * s.num_mask_tokens * 16 * s.grid * s.grid * 4;
    cudaError_t e = cudaMemcpyAsync(masks, t->masks.p, mb, cudaMemcpyDeviceToDevice, st);
    if (e == cudaSuccess) e = cudaMemcpyAsync(iou, t->iou.p, static_cast<size_t>(n) * s.num_mask_tokens * 4, cudaMemcpyDeviceToDevice, st);
    if (e != cudaSuccess) rc = samhost::set_error(2, "decoder training: output copy failed: %s", cudaGetErrorString(e));
  }
  if (rc) {
    delete t;
    return rc;
  }
  *tape_out = t;
  return 0;
}

int samk_decoder_backward(void* tape, const float* d_masks, int mask_lo, int mask_hi, const float* d_iou, float* d_weights,
                          float* d_sparse, cudaStream_t st) {
  SAM_REQUIRE(tape != nullptr, "decoder backward: NULL tape");
  Tape& t = *static_cast<Tape*>(tape);
  SAM_REQUIRE(!t.back.empty(), "decoder backward: this tape has already been consumed");
  SAM_REQUIRE(d_weights != nullptr, "decoder backward: NULL d_weights");
  SAM_REQUIRE(0 <= mask_lo && mask_lo <= mask_hi && mask_hi <= t.s.num_mask_tokens, "decoder backward: mask range [%d, %d)", mask_lo, mask_hi);
  t.st = st;
  const SamDecoderShape& s = t.s;
  SAM_CHECK_CUDA(cudaMemsetAsync(t.base + t.act_cap, 0, t.grad_cap, st));
  if (d_masks)
    SAM_CHECK_CUDA(cudaMemcpyAsync(t.masks.g, d_masks, static_cast<size_t>(t.masks.numel()) * 4, cudaMemcpyDeviceToDevice, st));
  if (d_iou)
    SAM_CHECK_CUDA(cudaMemcpyAsync(t.iou.g, d_iou, static_cast<size_t>(t.n) * s.num_mask_tokens * 4, cudaMemcpyDeviceToDevice, st));
  for (size_t i = t.back.size(); i-- > 0;) {
    const int grp = t.back_group[i];
    // mask tokens outside [mask_lo, mask_hi) and an unused IoU prediction have an all-zero cotangent: their MLPs
    // would add exact zeros everywhere
    if (grp == 100 && !d_iou) continue;
    if (grp >= 1 && grp < 100 && (!d_masks || grp - 1 < mask_lo || grp - 1 >= mask_hi)) continue;
    t.scratch_reset();
    if (int rc = t.back[i]()) return rc;
  }
  t.back.clear();
  {
    samhost::LaunchScope scope(samhost::KC_DECODER, st);
    acc_kernel<<<blocks_for(static_cast<i64>(t.blob_elems)), 256, 0, st>>>(d_weights, t.gblob, static_cast<i64>(t.blob_elems));
    SAM_CHECK_CUDA(cudaGetLastError());
  }
  if (d_sparse && t.k > 0)
    SAM_CHECK_CUDA(cudaMemcpyAsync(d_sparse, t.sparse.g, static_cast<size_t>(t.n) * t.k * s.C * 4, cudaMemcpyDeviceToDevice, st));
  return 0;
}

void samk_decoder_tape_free(void* tape) { delete static_cast<Tape*>(tape); }

// The same composition without a tape: the inference path for prompts with more tokens than the fused kernels of
// decoder.cu hold in shared memory (T > 16, e.g. SamPredictor calls with more than 10 points).  Slower, no limit on k.
size_t samk_decoder_generic_workspace_bytes(const SamDecoderShape& s, int n, int k) {
  if (check_train_shape(s, n, k)) return 0;
  Tape t;
  t.s = s; t.n = n; t.k = k; t.tape_on = false;
  t.blob_elems = samk_decoder_weight_elems(s);
  if (size_regions(t)) return 0;
  return region_total(t);
}

int samk_decoder_forward_generic(const SamDecoderShape& s, const float* blob, const float* pe_tokens, const void* image_embeddings,
                                 int emb_fmt, int n_images, const int* img_index, const void* sparse, int sparse_fmt, int n, int k,
                                 const void* dense_vec, const void* dense_full, int dense_fmt, void* masks, void* iou, int out_fmt,
                                 void* workspace, size_t workspace_bytes, cudaStream_t st) {
  if (int rc = check_train_shape(s, n, k)) return rc;
  SAM_REQUIRE(blob && pe_tokens && image_embeddings && masks && iou && workspace && (k == 0 || sparse), "mask decoder: NULL argument");
  Tape t;
  t.s = s; t.n = n; t.k = k; t.st = st; t.tape_on = false;
  t.blob_elems = samk_decoder_weight_elems(s);
  TRY(size_regions(t));
  SAM_REQUIRE(region_total(t) <= workspace_bytes, "mask decoder: workspace of %zu bytes, %zu needed", workspace_bytes, region_total(t));
  t.dry = false;
  t.base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(workspace) + 255) & ~static_cast<uintptr_t>(255));
  t.act_off = t.grad_off = t.scratch_off = 0;
  TrainInputs in{blob, image_embeddings, emb_fmt, n_images, img_index, sparse, sparse_fmt, dense_vec, dense_full, dense_fmt, nullptr, 2,
                 pe_tokens};
  TRY(build_forward(t, in));
  const i64 nmask = t.masks.numel(), niou = static_cast<i64>(n) * s.num_mask_tokens;
  samhost::LaunchScope scope(samhost::KC_DECODER, st, 0.0, 0.0, 2);
  from_f32_kernel<<<blocks_for(nmask), 256, 0, st>>>(t.masks.p, masks, out_fmt, nmask);
  from_f32_kernel<<<blocks_for(niou), 256, 0, st>>>(t.iou.p, iou, out_fmt, niou);
  SAM_CHECK_CUDA(cudaGetLastError());
  return 0;
}

// ---------------------------------------------------------------------------------------------------------------
// fp32 linear forward / backward for text_hidden_fcs (model/anyref.py:116-124) in training
// ---------------------------------------------------------------------------------------------------------------
size_t samk_linear_f32_scratch_bytes(int M, int N, int K) {
  (void)K;
  const size_t splitk = static_cast<size_t>(768) * GT * GT * 4;   // bound of Tape::gemm's split-K partials
  const size_t cols = static_cast<size_t>((M + kColRows - 1) / kColRows) * N * 4;
  return align256(splitk) + align256(cols) + 1024;
}

int samk_linear_f32_forward(const float* X, const float* W, const float* b, float* Y, int M, int N, int K, int relu_act, void* scratch,
                            size_t scratch_bytes, cudaStream_t st) {
  SAM_REQUIRE(X && W && Y && M >= 0 && N > 0 && K > 0, "linear_f32_forward: bad arguments");
  SAM_REQUIRE(scratch && scratch_bytes >= samk_linear_f32_scratch_bytes(M, N, K), "linear_f32_forward: scratch too small");
  if (M == 0) return 0;
  // a few [SEG] rows against a 4096 x 4096 weight is a memory-bound product with 64 output tiles: Tape::gemm splits K
  Tape t;
  t.dry = false; t.st = st;
  t.base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(scratch) + 255) & ~static_cast<uintptr_t>(255));
  Gemm g{};
  g.A = X; g.ars = K; g.acs = 1;
  g.B = W; g.brs = 1; g.bcs = K;
  g.C = Y; g.crs = N;
  g.bias = b; g.bias_mod = N;
  g.M = M; g.N = N; g.K = K; g.alpha = 1.f;
  {
    samhost::ClassOverride as_gemm(samhost::KC_GEMM);
    TRY(t.gemm(g));
  }
  if (relu_act) {
    samhost::LaunchScope scope(samhost::KC_GEMM, st);
    const i64 n = static_cast<i64>(M) * N;
    relu_fwd_kernel<<<blocks_for(n), 256, 0, st>>>(Y, Y, n);
    SAM_CHECK_CUDA(cudaGetLastError());
  }
  return 0;
}

// dX = dY . W and dW = dY^T . X (both overwritten; either may be NULL), db += column sums of dY.  With relu_y (the forward's output of a
// fused ReLU) dY is first masked IN PLACE with (relu_y > 0)
int samk_linear_f32_backward(float* dY, const float* relu_y, const float* X, const float* W, float* dX, float* dW, float* db, int M, int N, int K,
                             void* scratch, size_t scratch_bytes, cudaStream_t st) {
  SAM_REQUIRE(dY && X && W && M >= 0 && N > 0 && K > 0, "linear_f32_backward: bad arguments");
  SAM_REQUIRE(scratch && scratch_bytes >= samk_linear_f32_scratch_bytes(M, N, K), "linear_f32_backward: scratch too small");
  if (M == 0) return 0;
  Tape t;
  t.dry = false; t.st = st;
  t.base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(scratch) + 255) & ~static_cast<uintptr_t>(255));
  if (relu_y) {
    samhost::LaunchScope scope(samhost::KC_GEMM, st);
    const i64 n = static_cast<i64>(M) * N;
    relu_mask_kernel<<<blocks_for(n), 256, 0, st>>>(dY, relu_y, n);
    SAM_CHECK_CUDA(cudaGetLastError());
  }
  if (dX) {
    Gemm g{};
    g.A = dY; g.ars = N; g.acs = 1;
    g.B = W; g.brs = K; g.bcs = 1;
    g.C = dX; g.crs = K;
    g.M = M; g.N = K; g.K = N; g.alpha = 1.f;
    TRY(t.gemm(g));
  }
  if (dW) {
    Gemm g{};
    g.A = dY; g.ars = 1; g.acs = N;
    g.B = X; g.brs = K; g.bcs = 1;
    g.C = dW; g.crs = K; g.accumulate = 0;
    g.M = N; g.N = K; g.K = M; g.alpha = 1.f;
    TRY(t.gemm(g));
  }
  if (db) TRY(t.colsum(dY, N, M, N, N, db));
  return 0;
}

// d_low [maps, low, low] (+)= adjoint of postprocess_masks applied to d_out [maps, out_h, out_w];  tmp [maps, in_h, in_w] fp32
int samk_postprocess_backward(const float* d_out, int maps, int low, int img, int in_h, int in_w, int out_h, int out_w, float* tmp,
                              float* d_low, cudaStream_t st) {
  SAM_REQUIRE(d_out && tmp && d_low && maps >= 0, "postprocess_backward: NULL argument");
  SAM_REQUIRE(low > 0 && img > 0 && in_h > 0 && in_w > 0 && in_h <= img && in_w <= img && out_h > 0 && out_w > 0,
              "postprocess_backward: bad sizes");
  if (maps == 0) return 0;
  SAM_CHECK_CUDA(cudaMemsetAsync(tmp, 0, static_cast<size_t>(maps) * in_h * in_w * 4, st));
  SAM_CHECK_CUDA(cudaMemsetAsync(d_low, 0, static_cast<size_t>(maps) * low * low * 4, st));
  {
    // adjoint of the second resampling: [in_h, in_w] (the crop of the img x img map) -> [out_h, out_w]
    samhost::LaunchScope scope(samhost::KC_POSTPROCESS, st);
    bilinear_bwd_kernel<<<blocks_for(static_cast<i64>(maps) * out_h * out_w), 256, 0, st>>>(
        d_out, tmp, maps, out_h, out_w, in_h, in_w, static_cast<float>(in_h) / out_h, static_cast<float>(in_w) / out_w);
    SAM_CHECK_CUDA(cudaGetLastError());
  }
  {
    // adjoint of the first resampling [low, low] -> [img, img]; only the [in_h, in_w] crop of its output carries gradient
    samhost::LaunchScope scope(samhost::KC_POSTPROCESS, st);
    const float sc = static_cast<float>(low) / img;
    bilinear_bwd_kernel<<<blocks_for(static_cast<i64>(maps) * in_h * in_w), 256, 0, st>>>(tmp, d_low, maps, in_h, in_w, low, low, sc, sc);
    SAM_CHECK_CUDA(cudaGetLastError());
  }
  return 0;
}
