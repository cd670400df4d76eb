// [SEG]-prompted mask decoder of AnyRef's SAM path on sm_100a -- fp32 CUDA-core kernels (the decoder is 0.06 % of the
// path's FLOPs and latency/HBM bound; fp32 keeps the mask logits within 1e-5 of the reference, SURVEY 7 hard part 1).
//
// Replaces, for all prompts of a batch in ONE call (the reference loops per image in Python, model/anyref.py:797-819):
//   mask_decoder.py:116-179  MaskDecoder.predict_masks        (token assembly, src = image_embedding + dense, upscaling,
//                                                             hypernetwork MLPs, mask product, IoU head)
//   transformer.py:62-106    TwoWayTransformer.forward
//   transformer.py:151-182   TwoWayAttentionBlock.forward
//   transformer.py:220-242   Attention.forward                (self / token->image / image->token)
//
// Data layout: tokens ("queries") [n, T, C] and image tokens ("keys") [n, HW, C] are fp32, token-major.  Weights come
// as one fp32 blob in state_dict order (see DecoderWeights below and anyref_b200/segment_anything/_pack.py).
#include <math.h>

#include "host_common.h"
#include "kernels.h"
#include "ptx.cuh"

namespace {

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
__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752f)); }

__device__ __forceinline__ float load_any(const void* p, int fmt, size_t i) {
  if (fmt == 2) return static_cast<const float*>(p)[i];
  return ptx::unpack1(static_cast<const uint16_t*>(p)[i], fmt);
}
__device__ __forceinline__ void store_any(void* p, int fmt, size_t i, float v) {
  if (fmt == 2)
    static_cast<float*>(p)[i] = v;
  else
    static_cast<uint16_t*>(p)[i] = ptx::pack1(v, fmt);
}

// ---------------------------------------------------------------------------------------------------------------
// Prologue: NCHW -> token-major transposition (+ dense prompt embedding), token assembly
// ---------------------------------------------------------------------------------------------------------------
// dst[p, t, c] = src[img(p), c, t] (+ dense_vec[c] | + dense_full[p, c, t]);   grid (HW/32, C/32, n), block (32, 8)
__global__ void __launch_bounds__(256)
nchw_to_tokens_kernel(const void* __restrict__ src, int src_fmt, const int* __restrict__ img_index,
                      const void* __restrict__ dense_vec, const void* __restrict__ dense_full, int dense_fmt,
                      float* __restrict__ dst, int C, int HW) {
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

// tokens[p, :, :] = cat(iou_token[1,C], mask_tokens[nm,C], sparse[p, k, C])          (mask_decoder.py:126-141)
__global__ void assemble_tokens_kernel(const float* __restrict__ iou_token, const float* __restrict__ mask_tokens,
                                       const void* __restrict__ sparse, int sparse_fmt, float* __restrict__ tokens,
                                       int n, int nm, int k, int C) {
  const int T = 1 + nm + k;
  const size_t total = static_cast<size_t>(n) * T * C;
  for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int c = i % C;
    const int t = (i / C) % T;
    const int p = i / (static_cast<size_t>(C) * T);
    float v;
    if (t == 0)
      v = iou_token[c];
    else if (t <= nm)
      v = mask_tokens[(t - 1) * C + c];
    else
      v = load_any(sparse, sparse_fmt, (static_cast<size_t>(p) * k + (t - 1 - nm)) * C + c);
    tokens[i] = v;
  }
}

// ---------------------------------------------------------------------------------------------------------------
// fp32 linear:  Y[M,N] = act((X [+ X2[row % x2_mod]]) . W[N,K]^T + b) [+ R]
// 128 x 64 x 16 tiles, 256 threads, 8 x 4 outputs per thread, register-prefetched double buffering.
// ---------------------------------------------------------------------------------------------------------------
constexpr int LBM = 128, LBN = 64, LBK = 16;

struct LinArgs {
  const float* X; int ldx;
  const float* X2; int ldx2; int x2_mod;
  const float* W;
  const float* b;
  const float* R; int ldr;
  float* Y; int ldy;
  int M, N, K, act;
};

__global__ void __launch_bounds__(256)
dec_linear_kernel(const LinArgs a) {
  __shared__ __align__(16) float As[2][LBK][LBM + 4];
  __shared__ __align__(16) float Ws[2][LBK][LBN + 4];
  const int tid = threadIdx.x;
  const int ty = tid >> 4, tx = tid & 15;
  const int m0 = blockIdx.x * LBM, n0 = blockIdx.y * LBN;
  // global-load assignment
  const int ar0 = tid >> 2, akq = tid & 3;      // A rows ar0 and ar0 + 64, k-quad akq
  const int wr = tid >> 2, wkq = tid & 3;       // W row wr, k-quad wkq
  float4 ra[2], rw;
  auto gload = [&](int k0) {
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int row = m0 + ar0 + i * 64;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (row < a.M) {
        v = *reinterpret_cast<const float4*>(a.X + static_cast<size_t>(row) * a.ldx + k0 + akq * 4);
        if (a.X2) {
          const float4 u = *reinterpret_cast<const float4*>(a.X2 + static_cast<size_t>(row % a.x2_mod) * a.ldx2 + k0 + akq * 4);
          v.x += u.x; v.y += u.y; v.z += u.z; v.w += u.w;
        }
      }
      ra[i] = v;
    }
    rw = __ldg(reinterpret_cast<const float4*>(a.W + static_cast<size_t>(n0 + wr) * a.K + k0 + wkq * 4));
  };
  auto sstore = [&](int buf) {
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int r = ar0 + i * 64;
      As[buf][akq * 4 + 0][r] = ra[i].x;
      As[buf][akq * 4 + 1][r] = ra[i].y;
      As[buf][akq * 4 + 2][r] = ra[i].z;
      As[buf][akq * 4 + 3][r] = ra[i].w;
    }
    Ws[buf][wkq * 4 + 0][wr] = rw.x;
    Ws[buf][wkq * 4 + 1][wr] = rw.y;
    Ws[buf][wkq * 4 + 2][wr] = rw.z;
    Ws[buf][wkq * 4 + 3][wr] = rw.w;
  };
  float acc[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  const int nk = a.K / LBK;
  gload(0);
  sstore(0);
  __syncthreads();
  for (int kb = 0; kb < nk; ++kb) {
    const int buf = kb & 1;
    if (kb + 1 < nk) gload((kb + 1) * LBK);
#pragma unroll
    for (int k = 0; k < LBK; ++k) {
      const float4 a0 = *reinterpret_cast<const float4*>(&As[buf][k][ty * 8]);
      const float4 a1 = *reinterpret_cast<const float4*>(&As[buf][k][ty * 8 + 4]);
      const float4 w = *reinterpret_cast<const float4*>(&Ws[buf][k][tx * 4]);
      const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float wv[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], wv[j], acc[i][j]);
    }
    if (kb + 1 < nk) {
      sstore(buf ^ 1);
      __syncthreads();
    }
  }
  const int col = n0 + tx * 4;
  float4 bb = make_float4(0.f, 0.f, 0.f, 0.f);
  if (a.b) bb = __ldg(reinterpret_cast<const float4*>(a.b + col));
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int row = m0 + ty * 8 + i;
    if (row >= a.M) continue;
    float4 y = make_float4(acc[i][0] + bb.x, acc[i][1] + bb.y, acc[i][2] + bb.z, acc[i][3] + bb.w);
    if (a.act == 1) {
      y.x = fmaxf(y.x, 0.f); y.y = fmaxf(y.y, 0.f); y.z = fmaxf(y.z, 0.f); y.w = fmaxf(y.w, 0.f);
    }
    if (a.R) {
      const float4 r = *reinterpret_cast<const float4*>(a.R + static_cast<size_t>(row) * a.ldr + col);
      y.x += r.x; y.y += r.y; y.z += r.z; y.w += r.w;
    }
    *reinterpret_cast<float4*>(a.Y + static_cast<size_t>(row) * a.ldy + col) = y;
  }
}

// Same contract, 128 x 16 x 16 tiles (2 x 4 outputs per thread) for launches with only a few hundred rows.
// The accumulation order over k is the same as in dec_linear_kernel (k ascending, one fmaf per term), so both kernels
// give bit-identical results.
__global__ void __launch_bounds__(256)
dec_linear_skinny_kernel(const LinArgs a) {
  constexpr int SBN = 16;
  __shared__ __align__(16) float As[2][LBK][LBM + 4];
  __shared__ __align__(16) float Ws[2][LBK][SBN + 4];
  const int tid = threadIdx.x;
  const int ty = tid >> 2, tx = tid & 3;          // rows 2 ty, 2 ty + 1; columns 4 tx .. 4 tx + 3
  const int m0 = blockIdx.x * LBM, n0 = blockIdx.y * SBN;
  const int ar0 = tid >> 2, akq = tid & 3;        // A rows ar0 and ar0 + 64, k-quad akq
  const int wr = tid >> 2, wkq = tid & 3;         // W row wr (threads 0..63), k-quad wkq
  float4 ra[2], rw = make_float4(0.f, 0.f, 0.f, 0.f);
  auto gload = [&](int k0) {
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int row = m0 + ar0 + i * 64;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (row < a.M) {
        v = *reinterpret_cast<const float4*>(a.X + static_cast<size_t>(row) * a.ldx + k0 + akq * 4);
        if (a.X2) {
          const float4 u = *reinterpret_cast<const float4*>(a.X2 + static_cast<size_t>(row % a.x2_mod) * a.ldx2 + k0 + akq * 4);
          v.x += u.x; v.y += u.y; v.z += u.z; v.w += u.w;
        }
      }
      ra[i] = v;
    }
    if (tid < 4 * SBN) rw = __ldg(reinterpret_cast<const float4*>(a.W + static_cast<size_t>(n0 + wr) * a.K + k0 + wkq * 4));
  };
  auto sstore = [&](int buf) {
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int r = ar0 + i * 64;
      As[buf][akq * 4 + 0][r] = ra[i].x;
      As[buf][akq * 4 + 1][r] = ra[i].y;
      As[buf][akq * 4 + 2][r] = ra[i].z;
      As[buf][akq * 4 + 3][r] = ra[i].w;
    }
    if (tid < 4 * SBN) {
      Ws[buf][wkq * 4 + 0][wr] = rw.x;
      Ws[buf][wkq * 4 + 1][wr] = rw.y;
      Ws[buf][wkq * 4 + 2][wr] = rw.z;
      Ws[buf][wkq * 4 + 3][wr] = rw.w;
    }
  };
  float acc[2][4];
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  const int nk = a.K / LBK;
  gload(0);
  sstore(0);
  __syncthreads();
  for (int kb = 0; kb < nk; ++kb) {
    const int buf = kb & 1;
    if (kb + 1 < nk) gload((kb + 1) * LBK);
#pragma unroll
    for (int k = 0; k < LBK; ++k) {
      const float2 av = *reinterpret_cast<const float2*>(&As[buf][k][ty * 2]);
      const float4 w = *reinterpret_cast<const float4*>(&Ws[buf][k][tx * 4]);
      acc[0][0] = fmaf(av.x, w.x, acc[0][0]); acc[0][1] = fmaf(av.x, w.y, acc[0][1]);
      acc[0][2] = fmaf(av.x, w.z, acc[0][2]); acc[0][3] = fmaf(av.x, w.w, acc[0][3]);
      acc[1][0] = fmaf(av.y, w.x, acc[1][0]); acc[1][1] = fmaf(av.y, w.y, acc[1][1]);
      acc[1][2] = fmaf(av.y, w.z, acc[1][2]); acc[1][3] = fmaf(av.y, w.w, acc[1][3]);
    }
    if (kb + 1 < nk) {
      sstore(buf ^ 1);
      __syncthreads();
    }
  }
  const int col = n0 + tx * 4;
  float4 bb = make_float4(0.f, 0.f, 0.f, 0.f);
  if (a.b) bb = __ldg(reinterpret_cast<const float4*>(a.b + col));
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const int row = m0 + ty * 2 + i;
    if (row >= a.M) continue;
    float4 y = make_float4(acc[i][0] + bb.x, acc[i][1] + bb.y, acc[i][2] + bb.z, acc[i][3] + bb.w);
    if (a.act == 1) {
      y.x = fmaxf(y.x, 0.f); y.y = fmaxf(y.y, 0.f); y.z = fmaxf(y.z, 0.f); y.w = fmaxf(y.w, 0.f);
    }
    if (a.R) {
      const float4 r = *reinterpret_cast<const float4*>(a.R + static_cast<size_t>(row) * a.ldr + col);
      y.x += r.x; y.y += r.y; y.z += r.z; y.w += r.w;
    }
    *reinterpret_cast<float4*>(a.Y + static_cast<size_t>(row) * a.ldy + col) = y;
  }
}

// ---------------------------------------------------------------------------------------------------------------
// Tensor-core path for the image-token-side linears (M = n*4096 rows): fp32 accuracy from bf16 tcgen05 MMAs by
// operand splitting.  x = hi + lo with hi = bf16(x), lo = bf16(x - hi)  (16 mantissa bits together);
//   x.w ~= hi_x.hi_w + hi_x.lo_w + lo_x.hi_w          (the dropped lo.lo term is ~2^-18 relative)
// which is ONE GEMM over a 3x longer K:  A' = [hi_x | hi_x | lo_x],  W' = [hi_w | lo_w | hi_w], fp32 accumulation in
// TMEM (gemm2.cu).  The activation split (with the positional-encoding add fused in) is one elementwise pass; the
// weight split is done once per weight change (samk_decoder_prepare).
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
dec_split3_rows_kernel(const float* __restrict__ X, int ldx, const float* __restrict__ X2, int ldx2, int x2_mod,
                       uint16_t* __restrict__ out, int M, int K) {
  const int kq = K >> 2;
  const size_t total = static_cast<size_t>(M) * kq;
  for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int row = static_cast<int>(i / kq), k = static_cast<int>(i % kq) * 4;
    float4 v = *reinterpret_cast<const float4*>(X + static_cast<size_t>(row) * ldx + k);
    if (X2) {
      const float4 u = *reinterpret_cast<const float4*>(X2 + static_cast<size_t>(row % x2_mod) * ldx2 + k);
      v.x += u.x; v.y += u.y; v.z += u.z; v.w += u.w;
    }
    const float f[4] = {v.x, v.y, v.z, v.w};
    uint16_t hi[4], lo[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const __nv_bfloat16 h = __float2bfloat16_rn(f[j]);
      const __nv_bfloat16 l = __float2bfloat16_rn(f[j] - __bfloat162float(h));
      hi[j] = *reinterpret_cast<const uint16_t*>(&h);
      lo[j] = *reinterpret_cast<const uint16_t*>(&l);
    }
    const uint2 H = make_uint2(hi[0] | (uint32_t(hi[1]) << 16), hi[2] | (uint32_t(hi[3]) << 16));
    const uint2 L = make_uint2(lo[0] | (uint32_t(lo[1]) << 16), lo[2] | (uint32_t(lo[3]) << 16));
    uint16_t* o = out + static_cast<size_t>(row) * (3 * K) + k;
    *reinterpret_cast<uint2*>(o) = H;
    *reinterpret_cast<uint2*>(o + K) = H;
    *reinterpret_cast<uint2*>(o + 2 * K) = L;
  }
}

// W fp32 [N, K] -> [N, 3K] bf16 = [hi | lo | hi]
__global__ void dec_wsplit3_kernel(const float* __restrict__ W, uint16_t* __restrict__ out, int N, int K) {
  const int total = N * K;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int n = i / K, k = i % K;
    const float x = W[i];
    const __nv_bfloat16 h = __float2bfloat16_rn(x);
    const __nv_bfloat16 l = __float2bfloat16_rn(x - __bfloat162float(h));
    uint16_t* o = out + static_cast<size_t>(n) * (3 * K) + k;
    o[0] = *reinterpret_cast<const uint16_t*>(&h);
    o[K] = *reinterpret_cast<const uint16_t*>(&l);
    o[2 * K] = *reinterpret_cast<const uint16_t*>(&h);
  }
}

// ---------------------------------------------------------------------------------------------------------------
// Attention kernels (transformer.py:220-242).  Inputs are the already-projected q / k / v.
// ---------------------------------------------------------------------------------------------------------------
constexpr int TMAX = 16;    // max tokens per prompt (1 IoU + 4 mask + up to 11 prompt embeddings)
constexpr int TQ = 8;       // query chunk of the token->image kernel

// token -> image: q [n,T,C], K/V [n,HW,C] -> out [n,T,C]; head dim 16.  grid (heads, n, ceil(T/TQ)), block 256.
// dynamic smem: TQ*HW scores + 16*TQ*16 reduction scratch + TQ*16 q
__global__ void __launch_bounds__(256)
dec_attn_t2i_kernel(const float* __restrict__ q, const float* __restrict__ Kp, const float* __restrict__ Vp,
                    float* __restrict__ out, int T, int HW, int C, float scale) {
  constexpr int DH = 16;
  extern __shared__ float sm[];
  float* sc = sm;                        // [TQ][HW]
  float* red = sc + TQ * HW;             // [16][TQ][DH]
  float* qs = red + 16 * TQ * DH;        // [TQ][DH]
  float* stat = qs + TQ * DH;            // [8 warps][TQ] scratch, then [TQ] results at stat + 64
  const int h = blockIdx.x, p = blockIdx.y, t0 = blockIdx.z * TQ;
  const int tq = min(TQ, T - t0);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid < TQ * DH) {
    const int t = tid / DH, d = tid % DH;
    qs[tid] = (t < tq) ? q[(static_cast<size_t>(p) * T + t0 + t) * C + h * DH + d] * scale : 0.f;
  }
  __syncthreads();
  float mx[TQ];
#pragma unroll
  for (int t = 0; t < TQ; ++t) mx[t] = -INFINITY;
  for (int j = tid; j < HW; j += 256) {
    const float4* kr = reinterpret_cast<const float4*>(Kp + (static_cast<size_t>(p) * HW + j) * C + h * DH);
    const float4 k0 = kr[0], k1 = kr[1], k2 = kr[2], k3 = kr[3];
#pragma unroll
    for (int t = 0; t < TQ; ++t) {
      const float4* qq = reinterpret_cast<const float4*>(qs + t * DH);
      const float4 q0 = qq[0], q1 = qq[1], q2 = qq[2], q3 = qq[3];
      float s = q0.x * k0.x;
      s = fmaf(q0.y, k0.y, s); s = fmaf(q0.z, k0.z, s); s = fmaf(q0.w, k0.w, s);
      s = fmaf(q1.x, k1.x, s); s = fmaf(q1.y, k1.y, s); s = fmaf(q1.z, k1.z, s); s = fmaf(q1.w, k1.w, s);
      s = fmaf(q2.x, k2.x, s); s = fmaf(q2.y, k2.y, s); s = fmaf(q2.z, k2.z, s); s = fmaf(q2.w, k2.w, s);
      s = fmaf(q3.x, k3.x, s); s = fmaf(q3.y, k3.y, s); s = fmaf(q3.z, k3.z, s); s = fmaf(q3.w, k3.w, s);
      sc[t * HW + j] = s;
      mx[t] = fmaxf(mx[t], s);
    }
  }
#pragma unroll
  for (int t = 0; t < TQ; ++t) {
    const float m = warp_max(mx[t]);
    if (lane == 0) stat[warp * TQ + t] = m;
  }
  __syncthreads();
  if (tid < TQ) {
    float m = stat[tid];
    for (int w = 1; w < 8; ++w) m = fmaxf(m, stat[w * TQ + tid]);
    stat[64 + tid] = m;
  }
  __syncthreads();
  float sum[TQ];
#pragma unroll
  for (int t = 0; t < TQ; ++t) {
    sum[t] = 0.f;
    mx[t] = stat[64 + t];
  }
  for (int j = tid; j < HW; j += 256) {
#pragma unroll
    for (int t = 0; t < TQ; ++t) {
      const float e = expf(sc[t * HW + j] - mx[t]);
      sc[t * HW + j] = e;
      sum[t] += e;
    }
  }
  __syncthreads();  // all of stat[0..64) consumed, scores final
#pragma unroll
  for (int t = 0; t < TQ; ++t) {
    const float s = warp_sum(sum[t]);
    if (lane == 0) stat[warp * TQ + t] = s;
  }
  // P.V : thread (kg, d) covers keys [kg*HW/16, (kg+1)*HW/16)
  const int kg = tid >> 4, d = tid & 15;
  float acc[TQ];
#pragma unroll
  for (int t = 0; t < TQ; ++t) acc[t] = 0.f;
  const int per = HW / 16;
  for (int j = kg * per; j < (kg + 1) * per; ++j) {
    const float v = Vp[(static_cast<size_t>(p) * HW + j) * C + h * DH + d];
#pragma unroll
    for (int t = 0; t < TQ; ++t) acc[t] = fmaf(sc[t * HW + j], v, acc[t]);
  }
#pragma unroll
  for (int t = 0; t < TQ; ++t) red[(kg * TQ + t) * DH + d] = acc[t];
  __syncthreads();
  if (tid < TQ * DH) {
    const int t = tid / DH, dd = tid % DH;
    if (t < tq) {
      float l = 0.f;
      for (int w = 0; w < 8; ++w) l += stat[w * TQ + t];
      float o = 0.f;
      for (int g = 0; g < 16; ++g) o += red[(g * TQ + t) * DH + dd];
      out[(static_cast<size_t>(p) * T + t0 + t) * C + h * DH + dd] = o / l;
    }
  }
}

// image -> token: q [n,HW,C], k/v [n,T,C] -> out [n,HW,C]; head dim 16, C = 128.
// block 256 = 32 image tokens x 8 heads; grid (HW/32, n).  out may alias q.
__global__ void __launch_bounds__(256)
dec_attn_i2t_kernel(const float* q, const float* __restrict__ kp, const float* __restrict__ vp, float* out, int T,
                    int HW, int C, int heads, float scale) {
  constexpr int DH = 16;
  __shared__ __align__(16) float ks[TMAX * 128];
  __shared__ __align__(16) float vs[TMAX * 128];
  const int p = blockIdx.y;
  const int tid = threadIdx.x;
  for (int i = tid; i < T * C; i += 256) {
    ks[i] = kp[static_cast<size_t>(p) * T * C + i];
    vs[i] = vp[static_cast<size_t>(p) * T * C + i];
  }
  __syncthreads();
  const int row = blockIdx.x * 32 + tid / heads, h = tid % heads;
  if (row >= HW) return;
  const size_t off = (static_cast<size_t>(p) * HW + row) * C + h * DH;
  float qv[DH];
  {
    const float4* q4 = reinterpret_cast<const float4*>(q + off);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float4 t = q4[i];
      qv[4 * i] = t.x * scale; qv[4 * i + 1] = t.y * scale; qv[4 * i + 2] = t.z * scale; qv[4 * i + 3] = t.w * scale;
    }
  }
  float s[TMAX];
  float m = -INFINITY;
#pragma unroll
  for (int t = 0; t < TMAX; ++t) {
    s[t] = -INFINITY;
    if (t < T) {
      float a = 0.f;
#pragma unroll
      for (int d = 0; d < DH; ++d) a = fmaf(qv[d], ks[t * C + h * DH + d], a);
      s[t] = a;
      m = fmaxf(m, a);
    }
  }
  float l = 0.f;
#pragma unroll
  for (int t = 0; t < TMAX; ++t) {
    s[t] = (t < T) ? expf(s[t] - m) : 0.f;
    l += s[t];
  }
  const float inv = 1.0f / l;
  float o[DH];
#pragma unroll
  for (int d = 0; d < DH; ++d) o[d] = 0.f;
#pragma unroll
  for (int t = 0; t < TMAX; ++t) {
    if (t < T) {
#pragma unroll
      for (int d = 0; d < DH; ++d) o[d] = fmaf(s[t], vs[t * C + h * DH + d], o[d]);
    }
  }
  float4* o4 = reinterpret_cast<float4*>(out + off);
#pragma unroll
  for (int i = 0; i < 4; ++i) o4[i] = make_float4(o[4 * i] * inv, o[4 * i + 1] * inv, o[4 * i + 2] * inv, o[4 * i + 3] * inv);
}

// token self-attention: q/k/v [n,T,C] (C = heads*32) -> out [n,T,C].  One CTA per prompt.
__global__ void __launch_bounds__(256)
dec_self_attn_kernel(const float* __restrict__ q, const float* __restrict__ k, const float* __restrict__ v,
                     float* __restrict__ out, int T, int C, int heads, float scale) {
  extern __shared__ float sm[];
  float* qs = sm;               // [T][C]
  float* ks = qs + T * C;
  float* vs = ks + T * C;
  float* sc = vs + T * C;       // [heads][T][T]
  const int p = blockIdx.x, tid = threadIdx.x;
  const int dh = C / heads;
  for (int i = tid; i < T * C; i += 256) {
    qs[i] = q[static_cast<size_t>(p) * T * C + i];
    ks[i] = k[static_cast<size_t>(p) * T * C + i];
    vs[i] = v[static_cast<size_t>(p) * T * C + i];
  }
  __syncthreads();
  for (int i = tid; i < heads * T * T; i += 256) {
    const int t2 = i % T, t1 = (i / T) % T, h = i / (T * T);
    float a = 0.f;
    for (int d = 0; d < dh; ++d) a = fmaf(qs[t1 * C + h * dh + d], ks[t2 * C + h * dh + d], a);
    sc[i] = a * scale;
  }
  __syncthreads();
  for (int i = tid; i < heads * T; i += 256) {
    float* r = sc + i * T;
    float m = -INFINITY;
    for (int t = 0; t < T; ++t) m = fmaxf(m, r[t]);
    float l = 0.f;
    for (int t = 0; t < T; ++t) {
      r[t] = expf(r[t] - m);
      l += r[t];
    }
    const float inv = 1.0f / l;
    for (int t = 0; t < T; ++t) r[t] *= inv;
  }
  __syncthreads();
  for (int i = tid; i < T * C; i += 256) {
    const int c = i % C, t1 = i / C, h = c / dh;
    float a = 0.f;
    for (int t2 = 0; t2 < T; ++t2) a = fmaf(sc[(h * T + t1) * T + t2], vs[t2 * C + c], a);
    out[static_cast<size_t>(p) * T * C + i] = a;
  }
}

// ---------------------------------------------------------------------------------------------------------------
// Hypernetwork MLPs + IoU head (mask_decoder.py:159-177, :184-206): grid (nm + 1, n), block 256.
// which < nm: hyper_in[p, which, :] = MLP_which(hs[p, 1 + which]);   which == nm: iou[p, :] = head(hs[p, 0])
// ---------------------------------------------------------------------------------------------------------------
struct Mlp3 {
  const float *w0, *b0, *w1, *b1, *w2, *b2;
};
struct HyperArgs {
  Mlp3 mlp[5];
  int nm, C, hidden_iou, out_hyper, T;
  const float* hs;
  float* hyper;    // [n, nm, out_hyper] fp32
  void* iou;       // [n, nm] in iou_fmt
  int iou_fmt;
};

__device__ void mlp_layer(const float* __restrict__ W, const float* __restrict__ b, const float* x, float* y, int nout,
                          int nin, bool relu) {
  // one warp per output row, four rows in flight per warp (independent load -> fma chains: the layer is latency-bound);
  // the summation order of every output (k = lane, lane + 32, ..., then the warp tree) does not depend on the grouping
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int o = warp * 4; o < nout; o += 32) {
    float a[4] = {0.f, 0.f, 0.f, 0.f};
    for (int k = lane; k < nin; k += 32) {
      const float xv = x[k];
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (o + j < nout) a[j] = fmaf(__ldg(W + static_cast<size_t>(o + j) * nin + k), xv, a[j]);
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float t = warp_sum(a[j]);
      if (lane == 0 && o + j < nout) {
        const float v = t + b[o + j];
        y[o + j] = relu ? fmaxf(v, 0.f) : v;
      }
    }
  }
  __syncthreads();
}

__global__ void __launch_bounds__(256)
dec_hyper_kernel(const HyperArgs a) {
  __shared__ float x0[256], x1[256], x2[256];
  const int which = blockIdx.x, p = blockIdx.y;
  const int tok = (which < a.nm) ? 1 + which : 0;
  const int hidden = (which < a.nm) ? a.C : a.hidden_iou;
  const int nout = (which < a.nm) ? a.out_hyper : a.nm;
  for (int i = threadIdx.x; i < a.C; i += 256) x0[i] = a.hs[(static_cast<size_t>(p) * a.T + tok) * a.C + i];
  __syncthreads();
  const Mlp3& m = a.mlp[which];
  mlp_layer(m.w0, m.b0, x0, x1, hidden, a.C, true);
  mlp_layer(m.w1, m.b1, x1, x2, hidden, hidden, true);
  mlp_layer(m.w2, m.b2, x2, x0, nout, hidden, false);
  if (threadIdx.x < nout) {
    if (which < a.nm)
      a.hyper[(static_cast<size_t>(p) * a.nm + which) * a.out_hyper + threadIdx.x] = x0[threadIdx.x];
    else
      store_any(a.iou, a.iou_fmt, static_cast<size_t>(p) * a.nm + threadIdx.x, x0[threadIdx.x]);
  }
}

// ---------------------------------------------------------------------------------------------------------------
// Upscaling tail (mask_decoder.py:53-63, :157-158, :171-174).  The first ConvTranspose2d(k=2,s=2) is a per-pixel
// linear (dec_linear with the weight rearranged to [(dy,dx,oc), ic]); this kernel does, per (pixel, dy, dx):
// LayerNorm2d(64) -> GELU -> second ConvTranspose2d(64->32, k=2,s=2) -> GELU -> dot with the 4 hypernetwork vectors,
// writing the four 256x256 mask logits directly (the [n,32,256,256] tensor is never materialised).
// U [n*HW, 4*C1] fp32;  w1r [4 (ey,ex)][C2][C1];  masks [n, nm, 4g, 4g] in out_fmt.   C1 = 64, C2 = 32, nm <= 4.
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
dec_upscale_tail_kernel(const float* __restrict__ U, const float* __restrict__ ln_w, const float* __restrict__ ln_b,
                        const float* __restrict__ w1r, const float* __restrict__ b1, const float* __restrict__ hyper,
                        void* __restrict__ masks, int out_fmt, int g, int nm) {
  constexpr int C1 = 64, C2 = 32;
  __shared__ __align__(16) float ws[4 * C2 * C1];
  __shared__ float bs[C2], lw[C1], lb[C1], hy[4 * C2];
  const int HW = g * g;
  const size_t gt = static_cast<size_t>(blockIdx.x) * 256 + threadIdx.x;  // (p, pixel, sub)
  const int p = static_cast<int>(gt / (static_cast<size_t>(HW) * 4));
  for (int i = threadIdx.x; i < 4 * C2 * C1; i += 256) ws[i] = w1r[i];
  if (threadIdx.x < C2) bs[threadIdx.x] = b1[threadIdx.x];
  if (threadIdx.x < C1) {
    lw[threadIdx.x] = ln_w[threadIdx.x];
    lb[threadIdx.x] = ln_b[threadIdx.x];
  }
  if (threadIdx.x < nm * C2) hy[threadIdx.x] = hyper[static_cast<size_t>(p) * nm * C2 + threadIdx.x];
  __syncthreads();
  const int sub = gt & 3;
  const int pix = (gt >> 2) % HW;
  const int y = pix / g, x = pix % g;
  float a[C1];
  {
    const float4* u4 = reinterpret_cast<const float4*>(U + (gt >> 2) * (4 * C1) + sub * C1);
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < C1 / 4; ++i) {
      const float4 t = u4[i];
      a[4 * i] = t.x; a[4 * i + 1] = t.y; a[4 * i + 2] = t.z; a[4 * i + 3] = t.w;
      s += (t.x + t.y) + (t.z + t.w);
    }
    const float mean = s / C1;
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < C1; ++i) {
      a[i] -= mean;
      q = fmaf(a[i], a[i], q);
    }
    const float rstd = 1.0f / sqrtf(q / C1 + 1e-6f);
#pragma unroll
    for (int i = 0; i < C1; ++i) a[i] = gelu_erf(a[i] * rstd * lw[i] + lb[i]);
  }
  const int G4 = 4 * g;
  const int Y0 = 4 * y + 2 * (sub >> 1), X0 = 4 * x + 2 * (sub & 1);
#pragma unroll 1
  for (int s2 = 0; s2 < 4; ++s2) {
    float m[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll 2
    for (int oc = 0; oc < C2; ++oc) {
      const float4* w4 = reinterpret_cast<const float4*>(ws + (s2 * C2 + oc) * C1);
      float z0 = bs[oc], z1 = 0.f;
#pragma unroll
      for (int i = 0; i < C1 / 4; i += 2) {
        const float4 w0 = w4[i], w1 = w4[i + 1];
        z0 = fmaf(a[4 * i], w0.x, z0); z0 = fmaf(a[4 * i + 1], w0.y, z0);
        z0 = fmaf(a[4 * i + 2], w0.z, z0); z0 = fmaf(a[4 * i + 3], w0.w, z0);
        z1 = fmaf(a[4 * i + 4], w1.x, z1); z1 = fmaf(a[4 * i + 5], w1.y, z1);
        z1 = fmaf(a[4 * i + 6], w1.z, z1); z1 = fmaf(a[4 * i + 7], w1.w, z1);
      }
      const float gl = gelu_erf(z0 + z1);
#pragma unroll
      for (int k = 0; k < 4; ++k) m[k] = fmaf(hy[k * C2 + oc], gl, m[k]);
    }
    const int Y = Y0 + (s2 >> 1), X = X0 + (s2 & 1);
    for (int k = 0; k < nm; ++k)
      store_any(masks, out_fmt, ((static_cast<size_t>(p) * nm + k) * G4 + Y) * G4 + X, m[k]);
  }
}

// ---------------------------------------------------------------------------------------------------------------
// Host side
// ---------------------------------------------------------------------------------------------------------------
int launch_linear(const float* X, int ldx, const float* X2, int ldx2, int x2_mod, const float* W, const float* b,
                  const float* R, int ldr, float* Y, int ldy, int M, int N, int K, int act, cudaStream_t st) {
  SAM_REQUIRE(N % LBN == 0 && K % LBK == 0, "dec_linear: N=%d must be a multiple of %d and K=%d of %d", N, LBN, K, LBK);
  SAM_REQUIRE(ldx % 4 == 0 && ldy % 4 == 0 && (!X2 || ldx2 % 4 == 0) && (!R || ldr % 4 == 0), "dec_linear: ld %% 4");
  LinArgs a{X, ldx, X2, ldx2, x2_mod > 0 ? x2_mod : M, W, b, R, ldr, Y, ldy, M, N, K, act};
  samhost::LaunchScope scope(samhost::KC_DECODER, st, 2.0 * M * N * K);
  if (M <= 4 * LBM) {
    // token-side linears (a few hundred rows): 16-column tiles put N / 16 CTAs per row tile on the machine instead of
    // N / 64 -- these launches are latency-bound (a K = 2048 reduction walked by 4 CTAs took 133 us)
    dim3 grid((M + LBM - 1) / LBM, N / 16);
    dec_linear_skinny_kernel<<<grid, 256, 0, st>>>(a);
  } else {
    dim3 grid((M + LBM - 1) / LBM, N / LBN);
    dec_linear_kernel<<<grid, 256, 0, st>>>(a);
  }
  SAM_CHECK_CUDA(cudaGetLastError());
  return 0;
}

struct AttnW {
  const float *qw, *qb, *kw, *kb, *vw, *vb, *ow, *ob;
};
struct LayerW {
  AttnW self_attn;
  const float *n1w, *n1b;
  AttnW t2i;
  const float *n2w, *n2b;
  const float *l1w, *l1b, *l2w, *l2b;
  const float *n3w, *n3b, *n4w, *n4b;
  AttnW i2t;
};
struct DecoderWeights {
  const float *iou_token, *mask_tokens;
  LayerW layer[8];
  AttnW final_attn;
  const float *nfw, *nfb;
  const float *up0w, *up0b, *upln_w, *upln_b, *up1w, *up1b;
  Mlp3 hyper[4];
  Mlp3 iou_head;
  size_t total;
};

// Walks the blob in the documented order (== state_dict order of mask_decoder.*, conv weights rearranged).
int carve_weights(const SamDecoderShape& s, const float* blob, DecoderWeights* w) {
  const int C = s.C, Ci = C / 2, H = s.mlp_dim, nm = s.num_mask_tokens;
  size_t off = 0;
  auto take = [&](size_t n) {
    const float* p = blob ? blob + off : nullptr;
    off += n;
    return p;
  };
  auto attn = [&](AttnW& a, int internal) {
    a.qw = take((size_t)internal * C); a.qb = take(internal);
    a.kw = take((size_t)internal * C); a.kb = take(internal);
    a.vw = take((size_t)internal * C); a.vb = take(internal);
    a.ow = take((size_t)C * internal); a.ob = take(C);
  };
  w->iou_token = take(C);
  w->mask_tokens = take((size_t)nm * C);
  for (int l = 0; l < s.depth; ++l) {
    LayerW& L = w->layer[l];
    attn(L.self_attn, C);
    L.n1w = take(C); L.n1b = take(C);
    attn(L.t2i, Ci);
    L.n2w = take(C); L.n2b = take(C);
    L.l1w = take((size_t)H * C); L.l1b = take(H);
    L.l2w = take((size_t)C * H); L.l2b = take(C);
    L.n3w = take(C); L.n3b = take(C);
    L.n4w = take(C); L.n4b = take(C);
    attn(L.i2t, Ci);
  }
  attn(w->final_attn, Ci);
  w->nfw = take(C); w->nfb = take(C);
  const int C1 = C / 4, C2 = C / 8;
  w->up0w = take((size_t)4 * C1 * C); w->up0b = take(4 * C1);
  w->upln_w = take(C1); w->upln_b = take(C1);
  w->up1w = take((size_t)4 * C2 * C1); w->up1b = take(C2);
  for (int i = 0; i < nm; ++i) {
    Mlp3& m = w->hyper[i];
    m.w0 = take((size_t)C * C); m.b0 = take(C);
    m.w1 = take((size_t)C * C); m.b1 = take(C);
    m.w2 = take((size_t)C2 * C); m.b2 = take(C2);
  }
  const int Hi = s.iou_hidden;
  w->iou_head.w0 = take((size_t)Hi * C); w->iou_head.b0 = take(Hi);
  w->iou_head.w1 = take((size_t)Hi * Hi); w->iou_head.b1 = take(Hi);
  w->iou_head.w2 = take((size_t)nm * Hi); w->iou_head.b2 = take(nm);
  w->total = off;
  return 0;
}

// bf16x3-split copies of the weights that multiply image-token-sized operands (layout of the `derived` buffer)
struct DerivedW {
  const uint16_t *t2i_k[9], *t2i_v[9];   // per layer; index depth = final_attn_token_to_image
  const uint16_t *i2t_q[8], *i2t_o[8];
  const uint16_t* up0;
  size_t total;   // elements
};
void carve_derived(const SamDecoderShape& s, const uint16_t* base, DerivedW* d) {
  const size_t C = s.C, Ci = C / 2;
  size_t off = 0;
  auto take = [&](size_t n) {
    const uint16_t* p = base ? base + off : nullptr;
    off += (n + 7) & ~size_t(7);
    return p;
  };
  for (int l = 0; l <= s.depth; ++l) {
    d->t2i_k[l] = take(Ci * 3 * C);
    d->t2i_v[l] = take(Ci * 3 * C);
  }
  for (int l = 0; l < s.depth; ++l) {
    d->i2t_q[l] = take(Ci * 3 * C);
    d->i2t_o[l] = take(C * 3 * Ci);
  }
  d->up0 = take(C * 3 * C);
  d->total = off;
}

int check_shape(const SamDecoderShape& s) {
  SAM_REQUIRE(s.C == 256 && s.heads == 8, "mask decoder: transformer_dim must be 256 with 8 heads (got %d, %d)", s.C, s.heads);
  SAM_REQUIRE(s.depth >= 1 && s.depth <= 8, "mask decoder: depth %d unsupported", s.depth);
  SAM_REQUIRE(s.num_mask_tokens >= 1 && s.num_mask_tokens <= 4, "mask decoder: num_mask_tokens %d unsupported", s.num_mask_tokens);
  SAM_REQUIRE(s.mlp_dim % 64 == 0 && s.iou_hidden <= 256 && s.iou_hidden % 32 == 0, "mask decoder: mlp_dim/iou_hidden unsupported");
  SAM_REQUIRE(s.grid % 32 == 0, "mask decoder: embedding grid %d must be a multiple of 32", s.grid);
  return 0;
}

}  // namespace

size_t samk_decoder_weight_elems(const SamDecoderShape& s) {
  DecoderWeights w;
  carve_weights(s, nullptr, &w);
  return w.total;
}

size_t samk_decoder_derived_bytes(const SamDecoderShape& s) {
  DerivedW d;
  carve_derived(s, nullptr, &d);
  return d.total * sizeof(uint16_t);
}

int samk_decoder_prepare(const SamDecoderShape& s, const float* blob, void* derived, cudaStream_t st) {
  if (int rc = check_shape(s)) return rc;
  SAM_REQUIRE((reinterpret_cast<uintptr_t>(derived) & 15) == 0, "mask decoder: derived-weight buffer must be 16-byte aligned");
  DecoderWeights w;
  carve_weights(s, blob, &w);
  DerivedW d;
  carve_derived(s, static_cast<const uint16_t*>(derived), &d);
  const int C = s.C, Ci = C / 2;
  auto split = [&](const float* W, const uint16_t* out, int N, int K) {
    samhost::LaunchScope scope(samhost::KC_DECODER, st);
    dec_wsplit3_kernel<<<(N * K + 255) / 256, 256, 0, st>>>(W, const_cast<uint16_t*>(out), N, K);
  };
  for (int l = 0; l <= s.depth; ++l) {
    const AttnW& a = (l < s.depth) ? w.layer[l].t2i : w.final_attn;
    split(a.kw, d.t2i_k[l], Ci, C);
    split(a.vw, d.t2i_v[l], Ci, C);
  }
  for (int l = 0; l < s.depth; ++l) {
    split(w.layer[l].i2t.qw, d.i2t_q[l], Ci, C);
    split(w.layer[l].i2t.ow, d.i2t_o[l], C, Ci);
  }
  split(w.up0w, d.up0, C, C);
  SAM_CHECK_CUDA(cudaGetLastError());
  return 0;
}

size_t samk_decoder_workspace_bytes(const SamDecoderShape& s, int n, int k) {
  const size_t HW = static_cast<size_t>(s.grid) * s.grid, C = s.C, T = 1 + s.num_mask_tokens + k;
  size_t f = 0;
  f += HW * C;                         // pe_t
  f += 2 * n * HW * C;                 // keys, tmp
  f += 3 * n * HW * (C / 2);           // kbuf, vbuf, qbuf
  f += n * T * C * 8;                  // tokens0, queries, tq, tk, tv, ta, tb + slack
  f += n * T * s.mlp_dim;              // mlp hidden
  f += n * s.num_mask_tokens * (C / 8);  // hyper_in
  f += (3 * n * HW * C) / 2 + 16;         // a3: bf16 [n*HW, 3C] operand of the split-bf16 GEMMs
  return f * sizeof(float) + 256;
}

int samk_decoder_forward(const SamDecoderShape& s, const float* blob, const void* derived, const void* image_embeddings, int emb_fmt,
                         const int* img_index, const void* image_pe, int pe_fmt, const void* sparse, int sparse_fmt,
                         int n, int k, const void* dense_vec, const void* dense_full, int dense_fmt, void* masks,
                         void* iou, int out_fmt, void* workspace, size_t workspace_bytes, cudaStream_t st) {
  if (int rc = check_shape(s)) return rc;
  SAM_REQUIRE(n > 0 && k >= 0, "mask decoder: need at least one prompt");
  const int C = s.C, Ci = C / 2, nm = s.num_mask_tokens, T = 1 + nm + k, g = s.grid, HW = g * g;
  SAM_REQUIRE(T <= TMAX, "mask decoder: %d tokens per prompt exceed the supported maximum %d", T, TMAX);
  SAM_REQUIRE(workspace_bytes >= samk_decoder_workspace_bytes(s, n, k), "mask decoder: workspace too small");
  SAM_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 15) == 0 && (reinterpret_cast<uintptr_t>(blob) & 15) == 0,
              "mask decoder: workspace / weight blob must be 16-byte aligned");
  DecoderWeights w;
  carve_weights(s, blob, &w);
  DerivedW dw;
  carve_derived(s, static_cast<const uint16_t*>(derived), &dw);
  SAM_REQUIRE(derived != nullptr && (reinterpret_cast<uintptr_t>(derived) & 15) == 0,
              "mask decoder: derived weights missing (call sam_decoder_prepare) or misaligned");

  float* f = static_cast<float*>(workspace);
  auto take = [&](size_t nelem) {
    float* p = f;
    f += (nelem + 3) & ~size_t(3);
    return p;
  };
  float* pe_t = take((size_t)HW * C);
  float* keys = take((size_t)n * HW * C);
  float* tmp = take((size_t)n * HW * C);
  float* kbuf = take((size_t)n * HW * Ci);
  float* vbuf = take((size_t)n * HW * Ci);
  float* qbuf = take((size_t)n * HW * Ci);
  const size_t tc = (size_t)n * T * C;
  float* tok0 = take(tc);     // initial tokens == query_pe (transformer.py:95)
  float* qry = take(tc);      // running queries
  float* tq = take(tc);
  float* tk = take(tc);
  float* tv = take(tc);
  float* ta = take(tc);
  float* tb = take(tc);
  float* hid = take((size_t)n * T * s.mlp_dim);
  float* hyper = take((size_t)n * nm * (C / 8));
  uint16_t* a3 = reinterpret_cast<uint16_t*>(take(((size_t)3 * n * HW * C) / 2 + 8));
  const int MT = n * T, MK = n * HW;
  const float sc_self = 1.0f / sqrtf(static_cast<float>(C / s.heads));
  const float sc_cross = 1.0f / sqrtf(static_cast<float>(Ci / s.heads));

  // ---- prologue (mask_decoder.py:126-149, transformer.py:82-84)
  {
    dim3 grid(HW / 32, C / 32, n), blk(32, 8);
    samhost::LaunchScope scope(samhost::KC_DECODER, st, 0.0, 0.0, 3);
    nchw_to_tokens_kernel<<<grid, blk, 0, st>>>(image_embeddings, emb_fmt, img_index, dense_vec, dense_full, dense_fmt,
                                                keys, C, HW);
    dim3 grid1(HW / 32, C / 32, 1);
    nchw_to_tokens_kernel<<<grid1, blk, 0, st>>>(image_pe, pe_fmt, nullptr, nullptr, nullptr, 2, pe_t, C, HW);
    assemble_tokens_kernel<<<(MT * C + 255) / 256, 256, 0, st>>>(w.iou_token, w.mask_tokens, sparse, sparse_fmt, tok0,
                                                                n, nm, k, C);
    SAM_CHECK_CUDA(cudaGetLastError());
  }
  static samhost::PerDeviceOnce attr_once;
  const int t2i_smem = (TQ * HW + 16 * TQ * 16 + TQ * 16 + 64 + TQ) * sizeof(float);
  const int self_smem = (3 * TMAX * C + s.heads * TMAX * TMAX) * sizeof(float);
  if (attr_once.need()) {
    SAM_CHECK_CUDA(cudaFuncSetAttribute(dec_attn_t2i_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    SAM_CHECK_CUDA(cudaFuncSetAttribute(dec_self_attn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, self_smem));
    attr_once.done();
  }
  SAM_REQUIRE(t2i_smem <= 200 * 1024, "mask decoder: embedding grid %d too large for the token->image kernel", g);

#define LIN(...)                                   \
  do {                                             \
    if (int rc_ = launch_linear(__VA_ARGS__, st)) return rc_; \
  } while (0)
  // image-token-side linear on the tensor cores: Y = (X [+ X2]) . W^T + b, or Y += ... when `inplace`
  auto lin_tc = [&](const float* X, int ldx, const float* X2, int ldx2, int x2_mod, const uint16_t* W3, const float* b,
                    bool inplace, float* Y, int ldy, int M, int N, int K) -> int {
    {
      samhost::LaunchScope scope(samhost::KC_DECODER, st);
      const size_t total = static_cast<size_t>(M) * (K / 4);
      dec_split3_rows_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, st>>>(X, ldx, X2, ldx2,
                                                                                         x2_mod > 0 ? x2_mod : M, a3, M, K);
      SAM_CHECK_CUDA(cudaGetLastError());
    }
    GemmEpilogue ep{Y, ldy, SAM_F32, b, 0, inplace ? Y : nullptr, ldy, M};
    return samk_gemm(a3, 3 * K, W3, 3 * K, M, N, 3 * K, SAM_BF16, ep, st);
  };
#define LINTC(...)                              \
  do {                                          \
    if (int rc_ = lin_tc(__VA_ARGS__)) return rc_; \
  } while (0)
#define LNORM(x, res, gw, gb, out, M)                                                                       \
  do {                                                                                                      \
    if (int rc_ = samk_layernorm_rows(x, C, res, C, gw, gb, 1e-5f, out, C, SAM_F32, M, C, 1, st)) return rc_; \
  } while (0)

  // token -> image attention: queries(+pe) attend to keys(+pe); result (after out_proj) added to `qry`, then LN.
  auto token_to_image = [&](const AttnW& a, const uint16_t* k3, const uint16_t* v3, const float* gw,
                            const float* gb) -> int {
    LIN(qry, C, tok0, C, MT, a.qw, a.qb, nullptr, 0, tq, Ci, MT, Ci, C, 0);
    LINTC(keys, C, pe_t, C, HW, k3, a.kb, false, kbuf, Ci, MK, Ci, C);
    LINTC(keys, C, nullptr, 0, 0, v3, a.vb, false, vbuf, Ci, MK, Ci, C);
    dim3 grid(s.heads, n, (T + TQ - 1) / TQ);
    samhost::LaunchScope scope(samhost::KC_DECODER, st, 4.0 * n * T * HW * Ci);
    dec_attn_t2i_kernel<<<grid, 256, t2i_smem, st>>>(tq, kbuf, vbuf, ta, T, HW, Ci, sc_cross);
    SAM_CHECK_CUDA(cudaGetLastError());
    LIN(ta, Ci, nullptr, 0, 0, a.ow, a.ob, qry, C, tb, C, MT, C, Ci, 0);
    LNORM(tb, nullptr, gw, gb, qry, MT);
    return 0;
  };

  // layer loop (transformer.py:151-182)
  for (int l = 0; l < s.depth; ++l) {
    const LayerW& L = w.layer[l];
    const float* src = (l == 0) ? tok0 : qry;
    const float* pe = (l == 0) ? nullptr : tok0;   // skip_first_layer_pe
    // (1) self attention
    LIN(src, C, pe, C, MT, L.self_attn.qw, L.self_attn.qb, nullptr, 0, tq, C, MT, C, C, 0);
    LIN(src, C, pe, C, MT, L.self_attn.kw, L.self_attn.kb, nullptr, 0, tk, C, MT, C, C, 0);
    LIN(src, C, nullptr, 0, 0, L.self_attn.vw, L.self_attn.vb, nullptr, 0, tv, C, MT, C, C, 0);
    samhost::LaunchScope scope_sa(samhost::KC_DECODER, st, 4.0 * n * T * T * C);
    dec_self_attn_kernel<<<n, 256, (3 * T * C + s.heads * T * T) * sizeof(float), st>>>(tq, tk, tv, ta, T, C, s.heads,
                                                                                        sc_self);
    SAM_CHECK_CUDA(cudaGetLastError());
    LIN(ta, C, nullptr, 0, 0, L.self_attn.ow, L.self_attn.ob, (l == 0) ? nullptr : qry, C, tb, C, MT, C, C, 0);
    LNORM(tb, nullptr, L.n1w, L.n1b, qry, MT);
    // (2) token -> image cross attention
    if (int rc = token_to_image(L.t2i, dw.t2i_k[l], dw.t2i_v[l], L.n2w, L.n2b)) return rc;
    // (3) MLP (ReLU)
    LIN(qry, C, nullptr, 0, 0, L.l1w, L.l1b, nullptr, 0, hid, s.mlp_dim, MT, s.mlp_dim, C, 1);
    LIN(hid, s.mlp_dim, nullptr, 0, 0, L.l2w, L.l2b, qry, C, tb, C, MT, C, s.mlp_dim, 0);
    LNORM(tb, nullptr, L.n3w, L.n3b, qry, MT);
    // (4) image -> token cross attention
    LINTC(keys, C, pe_t, C, HW, dw.i2t_q[l], L.i2t.qb, false, qbuf, Ci, MK, Ci, C);
    LIN(qry, C, tok0, C, MT, L.i2t.kw, L.i2t.kb, nullptr, 0, tk, Ci, MT, Ci, C, 0);
    LIN(qry, C, nullptr, 0, 0, L.i2t.vw, L.i2t.vb, nullptr, 0, tv, Ci, MT, Ci, C, 0);
    {
      dim3 grid(HW / 32, n);
      samhost::LaunchScope scope(samhost::KC_DECODER, st, 4.0 * n * T * HW * Ci);
      dec_attn_i2t_kernel<<<grid, 256, 0, st>>>(qbuf, tk, tv, qbuf, T, HW, Ci, s.heads, sc_cross);
      SAM_CHECK_CUDA(cudaGetLastError());
    }
    LINTC(qbuf, Ci, nullptr, 0, 0, dw.i2t_o[l], L.i2t.ob, true, keys, C, MK, C, Ci);   // keys += out_proj(attn)
    LNORM(keys, nullptr, L.n4w, L.n4b, keys, MK);
  }
  // final token -> image attention + norm (transformer.py:99-104)
  if (int rc = token_to_image(w.final_attn, dw.t2i_k[s.depth], dw.t2i_v[s.depth], w.nfw, w.nfb)) return rc;

  // hypernetwork MLPs + IoU head
  {
    HyperArgs h;
    for (int i = 0; i < nm; ++i) h.mlp[i] = w.hyper[i];
    h.mlp[nm] = w.iou_head;
    h.nm = nm; h.C = C; h.hidden_iou = s.iou_hidden; h.out_hyper = C / 8; h.T = T;
    h.hs = qry; h.hyper = hyper; h.iou = iou; h.iou_fmt = out_fmt;
    dim3 grid(nm + 1, n);
    samhost::LaunchScope scope(samhost::KC_DECODER, st, 2.0 * n * (nm * (2.0 * C * C + C * C / 8) + 2.0 * C * C));
    dec_hyper_kernel<<<grid, 256, 0, st>>>(h);
    SAM_CHECK_CUDA(cudaGetLastError());
  }
  // upscaling + mask product
  LINTC(keys, C, nullptr, 0, 0, dw.up0, w.up0b, false, tmp, C, MK, C, C);
  samhost::LaunchScope scope_up(samhost::KC_DECODER, st, 2.0 * MK * 4 * (4.0 * (C / 8) * (C / 4) + 4.0 * nm * (C / 8)));
  dec_upscale_tail_kernel<<<static_cast<unsigned>((size_t)MK * 4 / 256), 256, 0, st>>>(
      tmp, w.upln_w, w.upln_b, w.up1w, w.up1b, hyper, masks, out_fmt, g, nm);
  SAM_CHECK_CUDA(cudaGetLastError());
#undef LIN
#undef LINTC
#undef LNORM
  return 0;
}
