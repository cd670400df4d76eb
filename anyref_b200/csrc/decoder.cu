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
//
// Launch plan (22 launches for depth 2; round 1: 45+).  Token side = one CTA per prompt holding its tokens in shared
// memory; image side = merged tcgen05 GEMMs over split-bf16 operands:
//   keys0 (+split)                       | token_first: assemble, self-attn, norm1, t2i q-proj
//   per layer l:  GEMM  keys.[Wk_t2i | Wv_t2i | Wq_i2t]^T   (one A operand for three projections: (keys + pe).W^T =
//                       keys.W^T + pe.W^T, and pe.W^T is a constant of the model kept in `derived`)
//                 attn_t2i (split over 16 key chunks) -> token_mlp (out-proj, norm2, MLP slice; 8 CTAs per prompt)
//                 -> token_tail (norm3, i2t k/v proj, NEXT layer's self-attn block or the final q-proj)
//                 attn_i2t (warp per pixel, writes the split A operand) -> GEMM out_proj -> ln_split (norm4 + split)
//   final:        GEMM  keys.[Wk_final | Wv_final | W_upscale0]^T -> attn_t2i -> hyper (out-proj, norm_final,
//                 hypernetwork MLPs, IoU head) -> up_prologue (LN2d, GELU, split) -> GEMM ConvT1 (+ GELU) -> up_hyper_dot
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

// ---------------------------------------------------------------------------------------------------------------
// Tensor-core path for the image-token-side linears (M = n*4096 rows): fp32 accuracy from bf16 tcgen05 MMAs by
// operand splitting.  x = hi + lo with hi = bf16(x), lo = bf16(x - hi)  (16 mantissa bits together);
//   x.w ~= hi_x.hi_w + hi_x.lo_w + lo_x.hi_w          (the dropped lo.lo term is ~2^-18 relative)
// which is ONE GEMM over a 3x longer K:  A' = [hi_x | hi_x | lo_x],  W' = [hi_w | lo_w | hi_w], fp32 accumulation in
// TMEM (gemm2.cu).  The activation split (with the positional-encoding add fused in) is one elementwise pass; the
// weight split is done once per weight change (samk_decoder_prepare).
// ---------------------------------------------------------------------------------------------------------------
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


// [N, K] fp32 -> [K, ldo] fp32 transposed copy at column col0 (token-side weights: a thread owns an output column and
// walks k, so consecutive threads must read consecutive addresses)
__global__ void dec_transpose_kernel(const float* __restrict__ W, float* __restrict__ out, int N, int K, int ldo, int col0) {
  __shared__ float tile[32][33];
  const int n0 = blockIdx.x * 32, k0 = blockIdx.y * 32;
  for (int i = threadIdx.y; i < 32; i += 8) {
    const int n = n0 + i, k = k0 + threadIdx.x;
    tile[i][threadIdx.x] = (n < N && k < K) ? W[static_cast<size_t>(n) * K + k] : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += 8) {
    const int k = k0 + i, n = n0 + threadIdx.x;
    if (n < N && k < K) out[static_cast<size_t>(k) * ldo + col0 + n] = tile[threadIdx.x][i];
  }
}
// out[i] = src ? src[i] : 0   (bias vectors of the merged GEMMs)
__global__ void dec_fill_kernel(float* __restrict__ out, const float* __restrict__ src, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = src ? src[i] : 0.f;
}

__device__ __forceinline__ void split_bf16(float x, uint16_t& hi, uint16_t& lo) {
  const __nv_bfloat16 h = __float2bfloat16_rn(x);
  const __nv_bfloat16 l = __float2bfloat16_rn(x - __bfloat162float(h));
  hi = *reinterpret_cast<const uint16_t*>(&h);
  lo = *reinterpret_cast<const uint16_t*>(&l);
}
// four consecutive fp32 -> 8 bytes of hi and 8 bytes of lo
__device__ __forceinline__ void split4(const float4 v, uint2& H, uint2& L) {
  uint16_t h[4], l[4];
  split_bf16(v.x, h[0], l[0]); split_bf16(v.y, h[1], l[1]); split_bf16(v.z, h[2], l[2]); split_bf16(v.w, h[3], l[3]);
  H = make_uint2(h[0] | (uint32_t(h[1]) << 16), h[2] | (uint32_t(h[3]) << 16));
  L = make_uint2(l[0] | (uint32_t(l[1]) << 16), l[2] | (uint32_t(l[3]) << 16));
}

// ---------------------------------------------------------------------------------------------------------------
// Prologue: keys0[s, t, c] = emb[img(s), c, t] + dense  (mask_decoder.py:146-147; token-major, transformer.py:82-84)
// written as fp32 AND as the split-bf16 A operand [hi | hi | lo] of the first merged GEMM.  s runs over SOURCES: the
// images when the dense embedding is the no_mask_embed broadcast (all prompts of an image share keys0 -- the repeat of
// mask_decoder.py:146 is never materialised), the prompts when each has its own dense mask embedding.
// grid (HW/32, C/32, n_src), block (32, 8)
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
dec_keys0_kernel(const void* __restrict__ src, int src_fmt, const int* __restrict__ img_index, const void* __restrict__ dense_vec,
                 const void* __restrict__ dense_full, int dense_fmt, float* __restrict__ keys, uint16_t* __restrict__ a3,
                 int C, int HW) {
  __shared__ float tile[32][33];
  const int s = blockIdx.z;
  const int img = dense_full ? (img_index ? img_index[s] : 0) : s;
  const int t0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  for (int i = threadIdx.y; i < 32; i += 8) {
    const int c = c0 + i, t = t0 + threadIdx.x;
    float v = load_any(src, src_fmt, (static_cast<size_t>(img) * C + c) * HW + t);
    if (dense_vec) v += load_any(dense_vec, dense_fmt, c);
    if (dense_full) v += load_any(dense_full, dense_fmt, (static_cast<size_t>(s) * C + c) * HW + t);
    tile[i][threadIdx.x] = v;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += 8) {
    const int t = t0 + i, c = c0 + threadIdx.x;
    const float v = tile[threadIdx.x][i];
    const size_t row = static_cast<size_t>(s) * HW + t;
    keys[row * C + c] = v;
    uint16_t hi, lo;
    split_bf16(v, hi, lo);
    uint16_t* o = a3 + row * (3 * C) + c;
    o[0] = hi;
    o[C] = hi;
    o[2 * C] = lo;
  }
}

// ---------------------------------------------------------------------------------------------------------------
// Token side.  One CTA owns ALL tokens of one prompt (T <= 16 rows of C = 256) in shared memory and runs whole
// sub-blocks of TwoWayAttentionBlock (transformer.py:151-182) without leaving the SM; the weights are read as
// [K, N]-transposed copies so that thread c walks column c with coalesced loads.  Round 1 ran every linear, norm and
// attention of the token side as its own launch (22 skinny-GEMM launches of ~25 us each per call).
// ---------------------------------------------------------------------------------------------------------------
constexpr int TMAX = 16;    // max tokens per prompt (1 IoU + 4 mask + up to 11 prompt embeddings)
constexpr int DC = 256;     // transformer_dim
constexpr int DCI = 128;    // cross-attention internal dim
constexpr int NCH = 16;     // key chunks of the split token->image attention
constexpr int PART = 16 + DCI;   // floats per (chunk, token): m[8 heads], l[8 heads], o[128]

constexpr int kTokThreads = 1024;   // token kernels: 256 columns x a 4-way split of the reduction dimension
constexpr int kKSplit = kTokThreads / 256;

// ys[t][c] = act(sum_k xs[t][k] * Wt[k][c] + b[c]) for t < TT, c < ncols <= 256 (block of kTokThreads threads, ALL of
// which must call this: it contains two block barriers and ends with one, so ys may be consumed right after).
// xs / ys in shared memory (row strides ldx / ldy floats, 16-byte aligned rows); Wt [K, ldw] fp32 in global memory;
// rs: shared scratch [kKSplit - 1][TT][256].  Thread (ks, c) walks column c over its quarter of K with 16 loads in
// flight; with 256 threads per prompt and 8 loads in flight the layer was bound by the L2 latency of the weight
// stream (ncu r02: 30 us per 256 x 256 linear).  The partials are summed in a fixed order: bit-reproducible.
template <int TT>
__device__ __forceinline__ void tok_linear(const float* __restrict__ Wt, int ldw, const float* __restrict__ b, int ncols,
                                           const float* xs, int ldx, int K, float* ys, int ldy, bool relu, float* rs) {
  const int c = threadIdx.x & 255, ks = threadIdx.x >> 8;
  const int kq = K / kKSplit;
  float acc[TT];
#pragma unroll
  for (int t = 0; t < TT; ++t) acc[t] = 0.f;
  if (c < ncols) {
    const float* w = Wt + static_cast<size_t>(ks * kq) * ldw + c;
    const float* x0 = xs + ks * kq;
#pragma unroll 4
    for (int k = 0; k < kq; k += 4) {
      const float w0 = __ldg(w + static_cast<size_t>(k) * ldw), w1 = __ldg(w + static_cast<size_t>(k + 1) * ldw);
      const float w2 = __ldg(w + static_cast<size_t>(k + 2) * ldw), w3 = __ldg(w + static_cast<size_t>(k + 3) * ldw);
#pragma unroll
      for (int t = 0; t < TT; ++t) {
        const float4 x = *reinterpret_cast<const float4*>(x0 + t * ldx + k);
        acc[t] = fmaf(x.x, w0, acc[t]);
        acc[t] = fmaf(x.y, w1, acc[t]);
        acc[t] = fmaf(x.z, w2, acc[t]);
        acc[t] = fmaf(x.w, w3, acc[t]);
      }
    }
    if (ks > 0) {
#pragma unroll
      for (int t = 0; t < TT; ++t) rs[((ks - 1) * TT + t) * 256 + c] = acc[t];
    }
  }
  __syncthreads();
  if (ks == 0 && c < ncols) {
    const float bias = b ? __ldg(b + c) : 0.f;
#pragma unroll
    for (int t = 0; t < TT; ++t) {
      float v = acc[t] + bias;
#pragma unroll
      for (int j = 0; j < kKSplit - 1; ++j) v += rs[(j * TT + t) * 256 + c];
      ys[t * ldy + c] = relu ? fmaxf(v, 0.f) : v;
    }
  }
  __syncthreads();
}

// rows t < T of xs [.., DC] (optionally + res rows) -> LayerNorm(eps 1e-5) -> out (shared) and out_g (global, optional)
__device__ __forceinline__ void tok_layernorm(const float* xs, const float* res, const float* __restrict__ gw,
                                              const float* __restrict__ gb, float* out, float* out_g, int T) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  for (int t = warp; t < T; t += nwarps) {
    float v[8];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int c = lane + 32 * i;
      v[i] = xs[t * DC + c] + (res ? res[t * DC + c] : 0.f);
      s += v[i];
    }
    const float mean = warp_sum(s) * (1.0f / DC);
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      v[i] -= mean;
      q = fmaf(v[i], v[i], q);
    }
    const float rstd = 1.0f / sqrtf(warp_sum(q) * (1.0f / DC) + 1e-5f);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int c = lane + 32 * i;
      const float y = v[i] * rstd * __ldg(gw + c) + __ldg(gb + c);
      out[t * DC + c] = y;
      if (out_g) out_g[t * DC + c] = y;
    }
  }
}

struct SelfW {   // transposed [K, N] weights + biases of one self-attention + norm1 + the t2i q projection that follows
  const float *qw, *qb, *kw, *kb, *vw, *vb, *ow, *ob;   // [DC, DC]
  const float *n1w, *n1b;
  const float *cqw, *cqb;                               // [DC, DCI] cross_attn_token_to_image.q_proj
};

// self-attention sub-block on the tokens in shared memory (transformer.py:153-161): xs <- LN1(xs + attn) (layer 0:
// LN1(attn), skip_first_layer_pe), then tq_g <- ((xs + pe) . Wq^T + bq) / sqrt(16) for the token->image attention.
// smem: xs, pe [TT][DC]; b0, b1, b2 [TT][DC] scratch; sc [8][TT][TT]
template <int TT>
__device__ __forceinline__ void tok_self_block(const SelfW& w, bool first, float* xs, const float* pe, float* b0, float* b1,
                                               float* b2, float* sc, float* rs, int T, int heads, float* q_g, float* tq_g) {
  const int tid = threadIdx.x;
  // b2 = xs + pe (layer 0: xs), the q / k input
  for (int i = tid; i < T * DC; i += kTokThreads) b2[i] = first ? xs[i] : xs[i] + pe[i];
  __syncthreads();
  tok_linear<TT>(w.qw, DC, w.qb, DC, b2, DC, DC, b0, DC, false, rs);    // q
  tok_linear<TT>(w.kw, DC, w.kb, DC, b2, DC, DC, b1, DC, false, rs);    // k
  __syncthreads();
  tok_linear<TT>(w.vw, DC, w.vb, DC, xs, DC, DC, b2, DC, false, rs);    // v (b2's old content is dead)
  __syncthreads();
  const int dh = DC / heads;
  const float scale = 1.0f / sqrtf(static_cast<float>(dh));
  for (int i = tid; i < heads * T * T; i += kTokThreads) {
    const int t2 = i % T, t1 = (i / T) % T, h = i / (T * T);
    float a = 0.f;
    for (int d = 0; d < dh; ++d) a = fmaf(b0[t1 * DC + h * dh + d], b1[t2 * DC + h * dh + d], a);
    sc[i] = a * scale;
  }
  __syncthreads();
  for (int i = tid; i < heads * T; i += kTokThreads) {
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
  for (int i = tid; i < T * DC; i += kTokThreads) {      // b1 <- attn (k is dead after the scores)
    const int c = i % DC, t1 = i / DC, h = c / dh;
    float a = 0.f;
    for (int t2 = 0; t2 < T; ++t2) a = fmaf(sc[(h * T + t1) * T + t2], b2[t2 * DC + c], a);
    b1[i] = a;
  }
  __syncthreads();
  tok_linear<TT>(w.ow, DC, w.ob, DC, b1, DC, DC, b0, DC, false, rs);   // out_proj -> b0
  __syncthreads();
  tok_layernorm(b0, first ? nullptr : xs, w.n1w, w.n1b, xs, q_g, T);
  __syncthreads();
  for (int i = tid; i < T * DC; i += kTokThreads) b2[i] = xs[i] + pe[i];
  __syncthreads();
  tok_linear<TT>(w.cqw, DCI, w.cqb, DCI, b2, DC, DC, b0, DCI, false, rs);
  __syncthreads();
  const float sc_cross = 1.0f / sqrtf(static_cast<float>(DCI / heads));
  for (int i = tid; i < T * DCI; i += kTokThreads) tq_g[i] = b0[i] * sc_cross;
}

constexpr int mlp_smem_floats(int TT) { return TT * DCI + 2 * TT * DC + (kKSplit - 1) * TT * 256; }
constexpr int tok_smem_floats(int TT) { return 5 * TT * DC + 8 * TT * TT + (kKSplit - 1) * TT * 256; }

// T1 of layer 0: token assembly (mask_decoder.py:126-141) + the first self-attention sub-block.   grid n, block 256
template <int TT>
__global__ void __launch_bounds__(kTokThreads)
dec_token_first_kernel(const float* __restrict__ iou_token, const float* __restrict__ mask_tokens, const void* __restrict__ sparse,
                       int sparse_fmt, int nm, int k, const SelfW w, int heads, float* __restrict__ tok0_g,
                       float* __restrict__ q_g, float* __restrict__ tq_g) {
  extern __shared__ __align__(16) float tsm[];
  const int T = 1 + nm + k, p = blockIdx.x, tid = threadIdx.x;
  float* xs = tsm;
  float* pe = xs + TT * DC;
  float* b0 = pe + TT * DC;
  float* b1 = b0 + TT * DC;
  float* b2 = b1 + TT * DC;
  float* sc = b2 + TT * DC;
  float* rs = sc + 8 * TT * TT;
  for (int i = tid; i < T * DC; i += kTokThreads) {
    const int c = i % DC, t = i / DC;
    float v;
    if (t == 0)
      v = iou_token[c];
    else if (t <= nm)
      v = mask_tokens[(t - 1) * DC + c];
    else
      v = load_any(sparse, sparse_fmt, (static_cast<size_t>(p) * k + (t - 1 - nm)) * DC + c);
    xs[i] = v;
    pe[i] = v;
    tok0_g[static_cast<size_t>(p) * T * DC + i] = v;    // query_pe of every later layer (transformer.py:95)
  }
  __syncthreads();
  tok_self_block<TT>(w, true, xs, pe, b0, b1, b2, sc, rs, T, heads, q_g + static_cast<size_t>(p) * T * DC,
                     tq_g + static_cast<size_t>(p) * T * DCI);
}

// ---------------------------------------------------------------------------------------------------------------
// token -> image attention, split over key chunks (transformer.py:162-166, :220-242 with the q projection done by the
// token kernel and the k / v projections by the merged GEMM).  K = kv[.., 0:128] + rk[pixel] (rk = pe.Wk^T + bk, the
// positional half of (keys + pe).Wk^T, constant per model), V = kv[.., 128:256] (bias added by the GEMM).
// A warp reads one key row (512 B, coalesced); lane l holds dims 4l..4l+3 = a quarter of head l/4, so a score is four
// FMAs and two shuffles.  Online softmax per (token, head); the chunk's (max, sum, unnormalised out) go to `part` and
// are combined by the consumer (dec_token_mlp_kernel / dec_hyper_kernel).  grid (NCH, n), block 256.
// Source rows of prompt p: one_src ? block 0 : src_index ? block src_index[p] : block p (same in the kernels below).
// ---------------------------------------------------------------------------------------------------------------
template <int TT>
__global__ void __launch_bounds__(256)
dec_attn_t2i_kernel(const float* __restrict__ tq, const float* __restrict__ kv, int ldkv, const float* __restrict__ rk,
                    const int* __restrict__ src_index, float* __restrict__ part, int T, int HW, int one_src) {
  extern __shared__ __align__(16) float red_raw[];     // [8 warps][TT][PART]
  float (*red)[TT][PART] = reinterpret_cast<float (*)[TT][PART]>(red_raw);
  const int ch = blockIdx.x, p = blockIdx.y;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int src = one_src ? 0 : (src_index ? src_index[p] : p);
  float4 q[TT];
#pragma unroll
  for (int t = 0; t < TT; ++t)
    q[t] = (t < T) ? *reinterpret_cast<const float4*>(tq + (static_cast<size_t>(p) * T + t) * DCI + 4 * lane)
                   : make_float4(0.f, 0.f, 0.f, 0.f);
  float m[TT], l[TT];
  float4 o[TT];
#pragma unroll
  for (int t = 0; t < TT; ++t) {
    m[t] = -INFINITY;
    l[t] = 0.f;
    o[t] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  const int per = HW / NCH;
  const int j0 = ch * per;
#pragma unroll 1
  for (int j = j0 + warp; j < j0 + per; j += 8) {
    const float* row = kv + (static_cast<size_t>(src) * HW + j) * ldkv;
    float4 k4 = *reinterpret_cast<const float4*>(row + 4 * lane);
    const float4 r4 = __ldg(reinterpret_cast<const float4*>(rk + static_cast<size_t>(j) * DCI) + lane);
    const float4 v4 = *reinterpret_cast<const float4*>(row + DCI + 4 * lane);
    k4.x += r4.x; k4.y += r4.y; k4.z += r4.z; k4.w += r4.w;
#pragma unroll
    for (int t = 0; t < TT; ++t) {
      if (t < T) {
        float s = q[t].x * k4.x;
        s = fmaf(q[t].y, k4.y, s); s = fmaf(q[t].z, k4.z, s); s = fmaf(q[t].w, k4.w, s);
        s += __shfl_xor_sync(0xffffffffu, s, 1);
        s += __shfl_xor_sync(0xffffffffu, s, 2);
        const float e = expf(-fabsf(s - m[t]));      // m = -inf on the first key: e = 0
        const bool up = s > m[t];
        const float alpha = up ? e : 1.0f, pr = up ? 1.0f : e;
        m[t] = up ? s : m[t];
        l[t] = fmaf(l[t], alpha, pr);
        o[t].x = fmaf(o[t].x, alpha, pr * v4.x); o[t].y = fmaf(o[t].y, alpha, pr * v4.y);
        o[t].z = fmaf(o[t].z, alpha, pr * v4.z); o[t].w = fmaf(o[t].w, alpha, pr * v4.w);
      }
    }
  }
#pragma unroll
  for (int t = 0; t < TT; ++t) {
    if (t < T) {
      if ((lane & 3) == 0) {
        red[warp][t][lane >> 2] = m[t];
        red[warp][t][8 + (lane >> 2)] = l[t];
      }
      *reinterpret_cast<float4*>(&red[warp][t][16 + 4 * lane]) = o[t];
    }
  }
  __syncthreads();
  // combine the 8 warps: thread (t, c) for c < 128 (+ the m / l entries by c < 8)
  for (int i = threadIdx.x; i < T * DCI; i += 256) {
    const int t = i / DCI, c = i % DCI, h = c >> 4;
    float M = -INFINITY;
#pragma unroll
    for (int w = 0; w < 8; ++w) M = fmaxf(M, red[w][t][h]);
    float L = 0.f, O = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) {
      const float e = expf(red[w][t][h] - M);
      L = fmaf(red[w][t][8 + h], e, L);
      O = fmaf(red[w][t][16 + c], e, O);
    }
    float* dst = part + ((static_cast<size_t>(p) * NCH + ch) * T + t) * PART;
    dst[16 + c] = O;
    if ((c & 15) == 0) {
      dst[h] = M;
      dst[8 + h] = L;
    }
  }
}

// combine the NCH chunk partials of prompt p into the attention output rows ta[t][0:128] (shared memory), t in [t0, t1)
__device__ __forceinline__ void t2i_combine(const float* __restrict__ part, int p, int T, int t0, int t1, float* ta) {
  for (int i = threadIdx.x; i < (t1 - t0) * DCI; i += blockDim.x) {
    const int t = t0 + i / DCI, c = i % DCI, h = c >> 4;
    const float* base = part + (static_cast<size_t>(p) * NCH * T + t) * PART;
    float M = -INFINITY;
    for (int ch = 0; ch < NCH; ++ch) M = fmaxf(M, base[static_cast<size_t>(ch) * T * PART + h]);
    float L = 0.f, O = 0.f;
    for (int ch = 0; ch < NCH; ++ch) {
      const float* b = base + static_cast<size_t>(ch) * T * PART;
      const float e = expf(b[h] - M);
      L = fmaf(b[8 + h], e, L);
      O = fmaf(b[16 + c], e, O);
    }
    ta[(t - t0) * DCI + c] = O / L;
  }
}

struct CrossMlpW {   // after the token->image attention: out_proj + norm2, MLP; all weights transposed [K, N]
  const float *ow, *ob;         // [DCI, DC]
  const float *n2w, *n2b;
  const float *l1w, *l1b;       // [DC, H]
  const float *l2w;             // [H, DC]
};

// T2: queries <- LN2(queries + out_proj(attn)) (computed redundantly by the 8 CTAs of a prompt), then this CTA's
// 256-wide slice j of the MLP hidden layer: part_mlp[p][j] = relu(x.W1_j^T + b1_j) . W2[:, j]^T  (transformer.py:166-172).
// CTA j == 0 also publishes the post-norm2 queries.   grid (H / 256, n), block 256
template <int TT>
__global__ void __launch_bounds__(kTokThreads)
dec_token_mlp_kernel(const float* __restrict__ part, float* __restrict__ q_g, const CrossMlpW w, int T, int H,
                     float* __restrict__ x2_g, float* __restrict__ part_mlp) {
  extern __shared__ __align__(16) float tsm[];
  float* ta = tsm;                   // [TT][DCI]
  float* xs = ta + TT * DCI;         // [TT][DC]
  float* ys = xs + TT * DC;          // [TT][DC]
  float* rs = ys + TT * DC;          // [kKSplit - 1][TT][256]
  const int j = blockIdx.x, p = blockIdx.y, tid = threadIdx.x;
  t2i_combine(part, p, T, 0, T, ta);
  for (int i = tid; i < T * DC; i += kTokThreads) xs[i] = q_g[static_cast<size_t>(p) * T * DC + i];
  __syncthreads();
  tok_linear<TT>(w.ow, DC, w.ob, DC, ta, DCI, DCI, ys, DC, false, rs);
  __syncthreads();
  tok_layernorm(ys, xs, w.n2w, w.n2b, xs, (j == 0) ? x2_g + static_cast<size_t>(p) * T * DC : nullptr, T);
  __syncthreads();
  tok_linear<TT>(w.l1w + j * 256, H, w.l1b + j * 256, 256, xs, DC, DC, ys, DC, true, rs);
  __syncthreads();
  tok_linear<TT>(w.l2w + static_cast<size_t>(j) * 256 * DC, DC, nullptr, DC, ys, DC, 256, xs, DC, false, rs);
  __syncthreads();
  float* dst = part_mlp + (static_cast<size_t>(p) * gridDim.x + j) * T * DC;
  for (int i = tid; i < T * DC; i += kTokThreads) dst[i] = xs[i];
}

struct TailW {   // end of a layer: lin2 bias + norm3, then the image->token k / v projections (concatenated columns)
  const float* l2b;
  const float *n3w, *n3b;
  const float *kvw, *kvb;       // [DC, 2 * DCI]: k_proj | v_proj of cross_attn_image_to_token
};

// T3: queries <- LN3(x2 + sum_j part_mlp[j] + b2); tk | tv <- image->token k / v projections (k from queries + pe);
// then either the NEXT layer's self-attention sub-block (has_next) or only the q projection of the final
// token->image attention (transformer.py:172-176, :99-101).   grid n, block 256
template <int TT>
__global__ void __launch_bounds__(kTokThreads)
dec_token_tail_kernel(const float* __restrict__ x2_g, const float* __restrict__ part_mlp, int nj, const TailW w,
                      const float* __restrict__ tok0_g, int T, int heads, float* __restrict__ tk_g, float* __restrict__ tv_g,
                      int has_next, const SelfW nw, float* __restrict__ q_g, float* __restrict__ tq_g) {
  extern __shared__ __align__(16) float tsm[];
  const int p = blockIdx.x, tid = threadIdx.x;
  float* xs = tsm;
  float* pe = xs + TT * DC;
  float* b0 = pe + TT * DC;
  float* b1 = b0 + TT * DC;
  float* b2 = b1 + TT * DC;
  float* sc = b2 + TT * DC;
  float* rs = sc + 8 * TT * TT;
  for (int i = tid; i < T * DC; i += kTokThreads) {
    float a = __ldg(w.l2b + (i % DC));
    for (int j = 0; j < nj; ++j) a += part_mlp[(static_cast<size_t>(p) * nj + j) * T * DC + i];
    b0[i] = a;
    xs[i] = x2_g[static_cast<size_t>(p) * T * DC + i];
    pe[i] = tok0_g[static_cast<size_t>(p) * T * DC + i];
  }
  __syncthreads();
  tok_layernorm(b0, xs, w.n3w, w.n3b, xs, q_g + static_cast<size_t>(p) * T * DC, T);
  __syncthreads();
  for (int i = tid; i < T * DC; i += kTokThreads) b2[i] = xs[i] + pe[i];
  __syncthreads();
  // k from (queries + pe), v from queries: two half-width passes over the concatenated weight
  tok_linear<TT>(w.kvw, 2 * DCI, w.kvb, DCI, b2, DC, DC, b0, DCI, false, rs);
  tok_linear<TT>(w.kvw + DCI, 2 * DCI, w.kvb + DCI, DCI, xs, DC, DC, b1, DCI, false, rs);
  __syncthreads();
  for (int i = tid; i < T * DCI; i += kTokThreads) {
    tk_g[static_cast<size_t>(p) * T * DCI + i] = b0[i];
    tv_g[static_cast<size_t>(p) * T * DCI + i] = b1[i];
  }
  __syncthreads();
  if (has_next) {
    tok_self_block<TT>(nw, false, xs, pe, b0, b1, b2, sc, rs, T, heads, q_g + static_cast<size_t>(p) * T * DC,
                       tq_g + static_cast<size_t>(p) * T * DCI);
  } else {
    tok_linear<TT>(nw.cqw, DCI, nw.cqb, DCI, b2, DC, DC, b0, DCI, false, rs);   // b2 still holds queries + pe
    __syncthreads();
    const float sc_cross = 1.0f / sqrtf(static_cast<float>(DCI / heads));
    for (int i = tid; i < T * DCI; i += kTokThreads) tq_g[static_cast<size_t>(p) * T * DCI + i] = b0[i] * sc_cross;
  }
}

// ---------------------------------------------------------------------------------------------------------------
// image -> token attention (transformer.py:174-180): one WARP per image token.  q = kvq[.., 256:384] + rq[pixel]
// (rq = pe.Wq^T + bq), softmax over the T tokens of the prompt, output written directly as the split-bf16 A operand
// [hi | hi | lo] of the out_proj GEMM.   grid (HW / 64, n), block 256 (8 pixels per warp)
// ---------------------------------------------------------------------------------------------------------------
template <int TT>
__global__ void __launch_bounds__(256)
dec_attn_i2t_kernel(const float* __restrict__ kvq, int ldq, int qcol, const float* __restrict__ rq, const int* __restrict__ src_index,
                    const float* __restrict__ tk, const float* __restrict__ tv, uint16_t* __restrict__ a3o, int T, int HW,
                    float scale, int one_src) {
  __shared__ __align__(16) float ks[TT * DCI];
  __shared__ __align__(16) float vs[TT * DCI];
  const int p = blockIdx.y, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int i = tid; i < T * DCI; i += 256) {
    ks[i] = tk[static_cast<size_t>(p) * T * DCI + i];
    vs[i] = tv[static_cast<size_t>(p) * T * DCI + i];
  }
  __syncthreads();
  const int src = one_src ? 0 : (src_index ? src_index[p] : p);
  float4 k4[TT];
#pragma unroll
  for (int t = 0; t < TT; ++t)
    k4[t] = (t < T) ? *reinterpret_cast<const float4*>(ks + t * DCI + 4 * lane) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 1
  for (int i = 0; i < 8; ++i) {
    const int pix = blockIdx.x * 64 + warp * 8 + i;
    float4 q4 = *reinterpret_cast<const float4*>(kvq + (static_cast<size_t>(src) * HW + pix) * ldq + qcol + 4 * lane);
    const float4 r4 = __ldg(reinterpret_cast<const float4*>(rq + static_cast<size_t>(pix) * DCI) + lane);
    q4.x = (q4.x + r4.x) * scale; q4.y = (q4.y + r4.y) * scale; q4.z = (q4.z + r4.z) * scale; q4.w = (q4.w + r4.w) * scale;
    float s[TT];
    float m = -INFINITY;
#pragma unroll
    for (int t = 0; t < TT; ++t) {
      float a = q4.x * k4[t].x;
      a = fmaf(q4.y, k4[t].y, a); a = fmaf(q4.z, k4[t].z, a); a = fmaf(q4.w, k4[t].w, a);
      a += __shfl_xor_sync(0xffffffffu, a, 1);
      a += __shfl_xor_sync(0xffffffffu, a, 2);
      s[t] = (t < T) ? a : -INFINITY;
      m = fmaxf(m, s[t]);
    }
    float l = 0.f;
    float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int t = 0; t < TT; ++t) {
      if (t < T) {
        const float e = expf(s[t] - m);
        l += e;
        const float4 v4 = *reinterpret_cast<const float4*>(vs + t * DCI + 4 * lane);
        o.x = fmaf(e, v4.x, o.x); o.y = fmaf(e, v4.y, o.y); o.z = fmaf(e, v4.z, o.z); o.w = fmaf(e, v4.w, o.w);
      }
    }
    const float inv = 1.0f / l;
    o.x *= inv; o.y *= inv; o.z *= inv; o.w *= inv;
    uint2 H, L;
    split4(o, H, L);
    uint16_t* dst = a3o + (static_cast<size_t>(p) * HW + pix) * (3 * DCI) + 4 * lane;
    *reinterpret_cast<uint2*>(dst) = H;
    *reinterpret_cast<uint2*>(dst + DCI) = H;
    *reinterpret_cast<uint2*>(dst + 2 * DCI) = L;
  }
}

// keys <- LN4(prev[src row] + delta) (transformer.py:180-181; delta = out_proj output incl. bias), written as fp32 and
// as the split-bf16 A operand of the next merged GEMM.  One warp per row of C = 256.   grid rows / 8, block 256
__global__ void __launch_bounds__(256)
dec_ln_split_kernel(const float* __restrict__ delta, const float* __restrict__ prev, const int* __restrict__ src_index, int HW,
                    const float* __restrict__ gw, const float* __restrict__ gb, float* __restrict__ keys,
                    uint16_t* __restrict__ a3, size_t rows, int one_src) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const size_t row = static_cast<size_t>(blockIdx.x) * 8 + warp;
  if (row >= rows) return;
  const int p = static_cast<int>(row / HW), pix = static_cast<int>(row % HW);
  const size_t prow = one_src ? static_cast<size_t>(pix) : (src_index ? static_cast<size_t>(src_index[p]) * HW + pix : row);
  float v[8];
  {
    const float4* d4 = reinterpret_cast<const float4*>(delta + row * DC);
    const float4* p4 = reinterpret_cast<const float4*>(prev + prow * DC);
    const float4 a = d4[lane], b = d4[32 + lane], c = p4[lane], d = p4[32 + lane];
    v[0] = a.x + c.x; v[1] = a.y + c.y; v[2] = a.z + c.z; v[3] = a.w + c.w;
    v[4] = b.x + d.x; v[5] = b.y + d.y; v[6] = b.z + d.z; v[7] = b.w + d.w;
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += v[i];
  const float mean = warp_sum(s) * (1.0f / DC);
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    v[i] -= mean;
    q = fmaf(v[i], v[i], q);
  }
  const float rstd = 1.0f / sqrtf(warp_sum(q) * (1.0f / DC) + 1e-5f);
  const float4 g0 = __ldg(reinterpret_cast<const float4*>(gw) + lane), g1 = __ldg(reinterpret_cast<const float4*>(gw) + 32 + lane);
  const float4 b0 = __ldg(reinterpret_cast<const float4*>(gb) + lane), b1 = __ldg(reinterpret_cast<const float4*>(gb) + 32 + lane);
  const float4 y0 = make_float4(v[0] * rstd * g0.x + b0.x, v[1] * rstd * g0.y + b0.y, v[2] * rstd * g0.z + b0.z, v[3] * rstd * g0.w + b0.w);
  const float4 y1 = make_float4(v[4] * rstd * g1.x + b1.x, v[5] * rstd * g1.y + b1.y, v[6] * rstd * g1.z + b1.z, v[7] * rstd * g1.w + b1.w);
  float4* k4 = reinterpret_cast<float4*>(keys + row * DC);
  k4[lane] = y0;
  k4[32 + lane] = y1;
  uint2 H0, L0, H1, L1;
  split4(y0, H0, L0);
  split4(y1, H1, L1);
  uint16_t* o = a3 + row * (3 * DC);
  reinterpret_cast<uint2*>(o)[lane] = H0;
  reinterpret_cast<uint2*>(o)[32 + lane] = H1;
  reinterpret_cast<uint2*>(o + DC)[lane] = H0;
  reinterpret_cast<uint2*>(o + DC)[32 + lane] = H1;
  reinterpret_cast<uint2*>(o + 2 * DC)[lane] = L0;
  reinterpret_cast<uint2*>(o + 2 * DC)[32 + lane] = L1;
}

// ---------------------------------------------------------------------------------------------------------------
// Hypernetwork MLPs + IoU head (mask_decoder.py:159-177, :184-206): grid (nm + 1, n), block 256.
// which < nm: hyper_in[p, which, :] = MLP_which(hs[p, 1 + which]);   which == nm: iou[p, :] = head(hs[p, 0])
// hs[p, tok] = norm_final_attn(queries + out_proj(final attention))[tok] (transformer.py:99-104) is computed here for
// the ONE token the CTA needs: combine of the attention partials, a 128 -> 256 projection and a LayerNorm.
// ---------------------------------------------------------------------------------------------------------------
struct Mlp3 {
  const float *w0, *b0, *w1, *b1, *w2, *b2;
};
struct HyperArgs {
  Mlp3 mlp[5];
  int nm, C, hidden_iou, out_hyper, T;
  const float* q;       // [n, T, C] queries after the last layer (before the final attention)
  const float* part;    // chunk partials of the FINAL token->image attention
  const float *ow, *ob; // final_attn_token_to_image.out_proj, weight transposed [DCI, C]
  const float *nfw, *nfb;
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
  {
    __shared__ __align__(16) float ta[DCI];
    __shared__ float redm[8], redq[8];
    t2i_combine(a.part, p, a.T, tok, tok + 1, ta);
    __syncthreads();
    const int c = threadIdx.x;      // C == 256 == blockDim.x
    float v = __ldg(a.ob + c);
    for (int k = 0; k < DCI; ++k) v = fmaf(ta[k], __ldg(a.ow + static_cast<size_t>(k) * a.C + c), v);
    v += a.q[(static_cast<size_t>(p) * a.T + tok) * a.C + c];
    const float s = warp_sum(v);
    if ((c & 31) == 0) redm[c >> 5] = s;
    __syncthreads();
    float mean = 0.f;
    for (int w = 0; w < 8; ++w) mean += redm[w];
    mean *= 1.0f / 256.0f;
    const float d = v - mean;
    const float qq = warp_sum(d * d);
    if ((c & 31) == 0) redq[c >> 5] = qq;
    __syncthreads();
    float var = 0.f;
    for (int w = 0; w < 8; ++w) var += redq[w];
    x0[c] = d * (1.0f / sqrtf(var * (1.0f / 256.0f) + 1e-5f)) * __ldg(a.nfw + c) + __ldg(a.nfb + c);
  }
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
// Upscaling tail (mask_decoder.py:53-63, :157-158, :171-174).  The first ConvTranspose2d(k=2,s=2) is a per-pixel linear
// and part of the final merged GEMM (U = its output, one row of 4 x 64 values per pixel).  The rest runs per
// (pixel, dy, dx) row of 64 channels:
//   up_prologue : LayerNorm2d(64) -> GELU -> split-bf16 A operand [hi | hi | lo] of the second ConvTranspose2d
//   tcgen05 GEMM: [rows, 3*64] x W1r3[128 = (ey,ex,oc), 3*64]^T + bias, GELU in the epilogue            (gemm2.cu)
//   up_hyper_dot: mask logits = sum_oc hyper[p, k, oc] * z[(ey,ex), oc], written straight to the 256x256 masks
// (the [n,32,256,256] tensor is never materialised).  Round 1 / early round 2 did all of it on fp32 CUDA cores in one
// kernel: 4.3 GFLOP at 17 TFLOP/s = 251 us for 16 prompts, the largest kernel of the decoder.
// U [n*HW, ldu >= 4*64] fp32;  a3 [n*HW*4, 192] bf16;  z [n*HW*4, 128] fp32;  masks [n, nm, 4g, 4g] in out_fmt.
// ---------------------------------------------------------------------------------------------------------------
constexpr int UC1 = 64, UC2 = 32;

// sixteen threads per (pixel, sub) row, one float4 of channels each: a warp reads 512 contiguous bytes per instruction
// and writes three 256-byte runs.   grid rows4 * 16 / 256, block 256
__global__ void __launch_bounds__(256)
dec_up_prologue_kernel(const float* __restrict__ U, int ldu, const float* __restrict__ ln_w, const float* __restrict__ ln_b,
                       uint16_t* __restrict__ a3, size_t rows4) {
  const size_t t = static_cast<size_t>(blockIdx.x) * 256 + threadIdx.x;
  const size_t row = t >> 4;          // (pixel, sub); rows4 * 16 is a multiple of 256: no partial blocks
  const int q = static_cast<int>(t & 15);
  float4 v = *reinterpret_cast<const float4*>(U + (row >> 2) * ldu + (row & 3) * UC1 + 4 * q);
  float s = (v.x + v.y) + (v.z + v.w);
#pragma unroll
  for (int o = 1; o < 16; o <<= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  const float mean = s / UC1;
  v.x -= mean; v.y -= mean; v.z -= mean; v.w -= mean;
  float qq = fmaf(v.x, v.x, fmaf(v.y, v.y, fmaf(v.z, v.z, v.w * v.w)));
#pragma unroll
  for (int o = 1; o < 16; o <<= 1) qq += __shfl_xor_sync(0xffffffffu, qq, o);
  const float rstd = 1.0f / sqrtf(qq / UC1 + 1e-6f);
  const float4 g4 = __ldg(reinterpret_cast<const float4*>(ln_w) + q), b4 = __ldg(reinterpret_cast<const float4*>(ln_b) + q);
  v.x = gelu_erf(v.x * rstd * g4.x + b4.x);
  v.y = gelu_erf(v.y * rstd * g4.y + b4.y);
  v.z = gelu_erf(v.z * rstd * g4.z + b4.z);
  v.w = gelu_erf(v.w * rstd * g4.w + b4.w);
  uint2 H, L;
  split4(v, H, L);
  uint16_t* o = a3 + row * (3 * UC1) + 4 * q;
  *reinterpret_cast<uint2*>(o) = H;
  *reinterpret_cast<uint2*>(o + UC1) = H;
  *reinterpret_cast<uint2*>(o + 2 * UC1) = L;
}

// one thread per (row, (ey,ex)): 32 GELU'd channels of z dotted with the nm hypernetwork vectors of the prompt.  The
// block's 64 rows of z (32 KB, contiguous) are staged through shared memory so the global reads are coalesced; each
// 32-channel run is padded to 33 floats, which makes both the staging writes and the per-thread reads conflict-free.
// grid rows4 / 64, block 256 (all rows of a block belong to one prompt: 4 * HW % 64 == 0)
__global__ void __launch_bounds__(256)
dec_up_hyper_dot_kernel(const float* __restrict__ z, const float* __restrict__ hyper, void* __restrict__ masks, int out_fmt,
                        int g, int nm) {
  __shared__ float zs[256 * 33];
  __shared__ float hy[4 * UC2];
  const int tid = threadIdx.x;
  const size_t row0 = static_cast<size_t>(blockIdx.x) * 64;
  const int HW = g * g;
  const int p = static_cast<int>(row0 / (static_cast<size_t>(HW) * 4));
  if (tid < 4 * UC2) hy[tid] = tid < nm * UC2 ? hyper[static_cast<size_t>(p) * nm * UC2 + tid] : 0.f;
  const float4* z4 = reinterpret_cast<const float4*>(z + row0 * (4 * UC2));
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int f = i * 256 + tid;              // float4 index in the block's tile: run = f / 8, offset 4 * (f % 8)
    const float4 v = z4[f];
    float* d = zs + (f >> 3) * 33 + 4 * (f & 7);
    d[0] = v.x; d[1] = v.y; d[2] = v.z; d[3] = v.w;
  }
  __syncthreads();
  const float* mine = zs + tid * 33;          // run tid = (local row, s2)
  float m[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int i = 0; i < UC2; ++i) {
    const float v = mine[i];
#pragma unroll
    for (int k = 0; k < 4; ++k) m[k] = fmaf(hy[k * UC2 + i], v, m[k]);
  }
  const size_t row = row0 + (tid >> 2);
  const int s2 = tid & 3;
  const int sub = static_cast<int>(row & 3);
  const int pix = static_cast<int>((row >> 2) % HW);
  const int y = pix / g, x = pix % g;
  const int G4 = 4 * g;
  const int Y = 4 * y + 2 * (sub >> 1) + (s2 >> 1), X = 4 * x + 2 * (sub & 1) + (s2 & 1);
  for (int k = 0; k < nm; ++k)
    store_any(masks, out_fmt, ((static_cast<size_t>(p) * nm + k) * G4 + Y) * G4 + X, m[k]);
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
  dim3 grid((M + LBM - 1) / LBM, N / LBN);
  dec_linear_kernel<<<grid, 256, 0, st>>>(a);
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


// Everything the forward needs besides the state_dict blob, rebuilt by samk_decoder_prepare whenever the weights or the
// dense positional encoding change (layout of the `derived` buffer):
//   * split-bf16 ([hi | lo | hi] along K) weights of the merged image-side GEMMs and the per-layer out_proj,
//   * their bias vectors (zero for the k / q column blocks: those biases travel inside rk / rq),
//   * rk / rq = pe . W^T + b  [HW, 128] fp32: the positional half of (keys + pe).W^T,
//   * [K, N]-transposed fp32 copies of the token-side weights (thread c of a CTA walks column c),
//   * pe_t: the token-major dense PE (scratch of the preparation).
struct DerivedW {
  const uint16_t* kvq[8];     // [3 Ci, 3 C]   rows: k_t2i | v_t2i | q_i2t of layer l
  const uint16_t* fin;        // [2 Ci + C, 3 C] rows: k_final | v_final | upscale conv 0
  const uint16_t* i2t_o[8];   // [C, 3 Ci]
  const float* b_kvq[8];      // [3 Ci] = 0 | v bias | 0
  const float* b_fin;         // [2 Ci + C] = 0 | v bias | upscale bias
  const float* rk[9];         // [HW, Ci]  (index depth = final attention)
  const float* rq[8];         // [HW, Ci]
  const float* pe_t;          // [HW, C]
  struct Tok {
    const float *sq, *sk, *sv, *so;    // self-attention [C, C]
    const float *cq, *co;              // t2i q [C, Ci], out [Ci, C]
    const float *l1, *l2;              // [C, H], [H, C]
    const float *ikv, *ikvb;           // i2t k | v [C, 2 Ci] and their biases [2 Ci]
  } tok[8];
  const float *fq, *fo;                // final attention q [C, Ci], out [Ci, C]
  const uint16_t* up1;                 // second ConvTranspose2d as a GEMM weight [(ey,ex,oc) = 128, 3 * 64]
  const float* b_up1;                  // its bias tiled over the 4 (ey,ex) positions [128]
  size_t total;                        // bytes
};
void carve_derived(const SamDecoderShape& s, const void* base, DerivedW* d) {
  const size_t C = s.C, Ci = C / 2, H = s.mlp_dim, HW = static_cast<size_t>(s.grid) * s.grid;
  size_t off = 0;
  auto take = [&](size_t bytes) {
    const uint8_t* p = base ? static_cast<const uint8_t*>(base) + off : nullptr;
    off += (bytes + 255) & ~size_t(255);
    return p;
  };
  auto t16 = [&](size_t n) { return reinterpret_cast<const uint16_t*>(take(n * 2)); };
  auto t32 = [&](size_t n) { return reinterpret_cast<const float*>(take(n * 4)); };
  for (int l = 0; l < s.depth; ++l) {
    d->kvq[l] = t16(3 * Ci * 3 * C);
    d->i2t_o[l] = t16(C * 3 * Ci);
    d->b_kvq[l] = t32(3 * Ci);
    d->rk[l] = t32(HW * Ci);
    d->rq[l] = t32(HW * Ci);
    DerivedW::Tok& t = d->tok[l];
    t.sq = t32(C * C); t.sk = t32(C * C); t.sv = t32(C * C); t.so = t32(C * C);
    t.cq = t32(C * Ci); t.co = t32(Ci * C);
    t.l1 = t32(C * H); t.l2 = t32(H * C);
    t.ikv = t32(C * 2 * Ci); t.ikvb = t32(2 * Ci);
  }
  d->fin = t16((2 * Ci + C) * 3 * C);
  d->b_fin = t32(2 * Ci + C);
  d->rk[s.depth] = t32(HW * Ci);
  d->fq = t32(C * Ci);
  d->fo = t32(Ci * C);
  d->pe_t = t32(HW * C);
  d->up1 = t16(static_cast<size_t>(4 * (C / 8)) * 3 * (C / 4));
  d->b_up1 = t32(4 * (C / 8));
  d->total = off;
}

int check_shape(const SamDecoderShape& s) {
  SAM_REQUIRE(s.C == DC && s.heads == 8, "mask decoder: transformer_dim must be 256 with 8 heads (got %d, %d)", s.C, s.heads);
  SAM_REQUIRE(s.depth >= 1 && s.depth <= 8, "mask decoder: depth %d unsupported", s.depth);
  SAM_REQUIRE(s.num_mask_tokens >= 1 && s.num_mask_tokens <= 4, "mask decoder: num_mask_tokens %d unsupported", s.num_mask_tokens);
  SAM_REQUIRE(s.mlp_dim % 256 == 0 && s.iou_hidden <= 256 && s.iou_hidden % 32 == 0, "mask decoder: mlp_dim/iou_hidden unsupported");
  SAM_REQUIRE(s.grid % 32 == 0 && (s.grid * s.grid) % (8 * NCH) == 0, "mask decoder: embedding grid %d must be a multiple of 32", s.grid);
  return 0;
}

template <typename K>
int opt_in_smem(K kernel, int bytes) {
  SAM_CHECK_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
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
  return d.total;
}

int samk_decoder_prepare(const SamDecoderShape& s, const float* blob, const void* image_pe, int pe_fmt, void* derived,
                         cudaStream_t st) {
  if (int rc = check_shape(s)) return rc;
  SAM_REQUIRE((reinterpret_cast<uintptr_t>(derived) & 255) == 0, "mask decoder: derived-weight buffer must be 256-byte aligned");
  SAM_REQUIRE(pe_fmt >= 0 && pe_fmt <= 2, "mask decoder: bad image_pe format");
  DecoderWeights w;
  carve_weights(s, blob, &w);
  DerivedW d;
  carve_derived(s, derived, &d);
  const int C = s.C, Ci = C / 2, H = s.mlp_dim, HW = s.grid * s.grid;
  auto split = [&](const float* W, const uint16_t* out, int row0, int N, int K) {
    samhost::LaunchScope scope(samhost::KC_DECODER, st);
    dec_wsplit3_kernel<<<(N * K + 255) / 256, 256, 0, st>>>(W, const_cast<uint16_t*>(out) + static_cast<size_t>(row0) * 3 * K, N, K);
  };
  auto transpose = [&](const float* W, const float* out, int N, int K, int ldo, int col0) {
    samhost::LaunchScope scope(samhost::KC_DECODER, st);
    dim3 grid((N + 31) / 32, (K + 31) / 32), blk(32, 8);
    dec_transpose_kernel<<<grid, blk, 0, st>>>(W, const_cast<float*>(out), N, K, ldo, col0);
  };
  auto fill = [&](const float* out, int off, const float* src, int n) {
    samhost::LaunchScope scope(samhost::KC_DECODER, st);
    dec_fill_kernel<<<(n + 255) / 256, 256, 0, st>>>(const_cast<float*>(out) + off, src, n);
  };
  {
    dim3 grid(HW / 32, C / 32, 1), blk(32, 8);
    samhost::LaunchScope scope(samhost::KC_DECODER, st);
    nchw_to_tokens_kernel<<<grid, blk, 0, st>>>(image_pe, pe_fmt, nullptr, nullptr, nullptr, 2, const_cast<float*>(d.pe_t), C, HW);
  }
  // r = pe_t . W^T + b  (fp32 FMA GEMM; once per weight / PE change)
  auto pe_proj = [&](const float* W, const float* b, const float* out) -> int {
    return launch_linear(d.pe_t, C, nullptr, 0, 0, W, b, nullptr, 0, const_cast<float*>(out), Ci, HW, Ci, C, 0, st);
  };
  for (int l = 0; l < s.depth; ++l) {
    const LayerW& L = w.layer[l];
    split(L.t2i.kw, d.kvq[l], 0, Ci, C);
    split(L.t2i.vw, d.kvq[l], Ci, Ci, C);
    split(L.i2t.qw, d.kvq[l], 2 * Ci, Ci, C);
    split(L.i2t.ow, d.i2t_o[l], 0, C, Ci);
    fill(d.b_kvq[l], 0, nullptr, Ci);
    fill(d.b_kvq[l], Ci, L.t2i.vb, Ci);
    fill(d.b_kvq[l], 2 * Ci, nullptr, Ci);
    if (int rc = pe_proj(L.t2i.kw, L.t2i.kb, d.rk[l])) return rc;
    if (int rc = pe_proj(L.i2t.qw, L.i2t.qb, d.rq[l])) return rc;
    const DerivedW::Tok& t = d.tok[l];
    transpose(L.self_attn.qw, t.sq, C, C, C, 0);
    transpose(L.self_attn.kw, t.sk, C, C, C, 0);
    transpose(L.self_attn.vw, t.sv, C, C, C, 0);
    transpose(L.self_attn.ow, t.so, C, C, C, 0);
    transpose(L.t2i.qw, t.cq, Ci, C, Ci, 0);
    transpose(L.t2i.ow, t.co, C, Ci, C, 0);
    transpose(L.l1w, t.l1, H, C, H, 0);
    transpose(L.l2w, t.l2, C, H, C, 0);
    transpose(L.i2t.kw, t.ikv, Ci, C, 2 * Ci, 0);
    transpose(L.i2t.vw, t.ikv, Ci, C, 2 * Ci, Ci);
    fill(t.ikvb, 0, L.i2t.kb, Ci);
    fill(t.ikvb, Ci, L.i2t.vb, Ci);
  }
  split(w.final_attn.kw, d.fin, 0, Ci, C);
  split(w.final_attn.vw, d.fin, Ci, Ci, C);
  split(w.up0w, d.fin, 2 * Ci, C, C);
  fill(d.b_fin, 0, nullptr, Ci);
  fill(d.b_fin, Ci, w.final_attn.vb, Ci);
  fill(d.b_fin, 2 * Ci, w.up0b, C);
  if (int rc = pe_proj(w.final_attn.kw, w.final_attn.kb, d.rk[s.depth])) return rc;
  transpose(w.final_attn.qw, d.fq, Ci, C, Ci, 0);
  transpose(w.final_attn.ow, d.fo, C, Ci, C, 0);
  split(w.up1w, d.up1, 0, 4 * (C / 8), C / 4);                       // [(ey,ex), oc, ic] is already [128, 64] row-major
  for (int e = 0; e < 4; ++e) fill(d.b_up1, e * (C / 8), w.up1b, C / 8);
  SAM_CHECK_CUDA(cudaGetLastError());
  return 0;
}

namespace {
struct DecWs {   // workspace carve (floats unless noted)
  float *keys0, *keys, *delta, *kvq;
  uint16_t *a3, *a3o;
  float *tok0, *q, *x2, *tq, *tk, *tv, *part, *pmlp, *hyper;
  size_t total;   // bytes
};
void carve_ws(const SamDecoderShape& s, int n, int n_src, int k, void* base, DecWs* w) {
  const size_t HW = static_cast<size_t>(s.grid) * s.grid, C = s.C, Ci = C / 2, T = 1 + s.num_mask_tokens + k;
  const size_t nmax = n > n_src ? n : n_src;
  size_t off = 0;
  auto take = [&](size_t bytes) {
    uint8_t* p = base ? static_cast<uint8_t*>(base) + off : nullptr;
    off += (bytes + 255) & ~size_t(255);
    return p;
  };
  w->keys0 = reinterpret_cast<float*>(take(static_cast<size_t>(n_src) * HW * C * 4));
  w->keys = reinterpret_cast<float*>(take(static_cast<size_t>(n) * HW * C * 4));
  w->delta = reinterpret_cast<float*>(take(static_cast<size_t>(n) * HW * C * 4));
  w->kvq = reinterpret_cast<float*>(take(nmax * HW * (2 * Ci + C) * 4));
  w->a3 = reinterpret_cast<uint16_t*>(take(nmax * HW * 3 * C * 2));
  w->a3o = reinterpret_cast<uint16_t*>(take(static_cast<size_t>(n) * HW * 3 * Ci * 2));
  const size_t tc = static_cast<size_t>(n) * T * C * 4;
  w->tok0 = reinterpret_cast<float*>(take(tc));
  w->q = reinterpret_cast<float*>(take(tc));
  w->x2 = reinterpret_cast<float*>(take(tc));
  w->tq = reinterpret_cast<float*>(take(tc / 2));
  w->tk = reinterpret_cast<float*>(take(tc / 2));
  w->tv = reinterpret_cast<float*>(take(tc / 2));
  w->part = reinterpret_cast<float*>(take(static_cast<size_t>(n) * NCH * T * PART * 4));
  w->pmlp = reinterpret_cast<float*>(take(static_cast<size_t>(n) * (s.mlp_dim / 256) * T * C * 4));
  w->hyper = reinterpret_cast<float*>(take(static_cast<size_t>(n) * s.num_mask_tokens * (C / 8) * 4));
  w->total = off;
}
}  // namespace

size_t samk_decoder_workspace_bytes(const SamDecoderShape& s, int n_images, int n, int k) {
  if (1 + s.num_mask_tokens + k > TMAX) return samk_decoder_generic_workspace_bytes(s, n, k);   // the generic path
  DecWs w;
  carve_ws(s, n, n_images > n ? n_images : n, k, nullptr, &w);   // covers both source modes (images / prompts)
  return w.total + 256;
}

int samk_decoder_forward(const SamDecoderShape& s, const float* blob, const void* derived, const void* image_embeddings, int emb_fmt,
                         int n_images, const int* img_index, const void* sparse, int sparse_fmt, int n, int k,
                         const void* dense_vec, const void* dense_full, int dense_fmt, void* masks, void* iou, int out_fmt,
                         void* workspace, size_t workspace_bytes, cudaStream_t st) {
  if (int rc = check_shape(s)) return rc;
  SAM_REQUIRE(n > 0 && k >= 0, "mask decoder: need at least one prompt");
  const int C = s.C, Ci = C / 2, nm = s.num_mask_tokens, T = 1 + nm + k, g = s.grid, HW = g * g, H = s.mlp_dim;
  if (T > TMAX) {
    // more tokens per prompt than the fused token kernels keep in shared memory (more than 11 sparse prompt embeddings,
    // e.g. many points): the plain fp32 composition of decoder_train.cu without its tape -- slower, no limit
    SAM_REQUIRE(derived != nullptr && (reinterpret_cast<uintptr_t>(derived) & 255) == 0,
                "mask decoder: derived weights missing (call sam_decoder_prepare) or misaligned");
    DerivedW dw;
    carve_derived(s, derived, &dw);
    return samk_decoder_forward_generic(s, blob, dw.pe_t, image_embeddings, emb_fmt, n_images, img_index, sparse, sparse_fmt, n, k,
                                        dense_vec, dense_full, dense_fmt, masks, iou, out_fmt, workspace, workspace_bytes, st);
  }
  SAM_REQUIRE(n_images >= 1 && (img_index || n_images == 1), "mask decoder: img_index is required with several image embeddings");
  SAM_REQUIRE(n <= 65535, "mask decoder: at most 65535 prompts per call");
  // sources of the layer-0 image tokens: the images (shared by their prompts) unless every prompt has its own dense embedding
  const int n_src = dense_full ? n : n_images;
  SAM_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 255) == 0 && (reinterpret_cast<uintptr_t>(blob) & 15) == 0,
              "mask decoder: workspace must be 256-byte and the weight blob 16-byte aligned");
  SAM_REQUIRE(derived != nullptr && (reinterpret_cast<uintptr_t>(derived) & 255) == 0,
              "mask decoder: derived weights missing (call sam_decoder_prepare) or misaligned");
  DecoderWeights w;
  carve_weights(s, blob, &w);
  DerivedW dw;
  carve_derived(s, derived, &dw);
  DecWs ws;
  carve_ws(s, n, n_src, k, workspace, &ws);
  SAM_REQUIRE(workspace_bytes >= ws.total, "mask decoder: workspace too small (%zu < %zu)", workspace_bytes, ws.total);
  const bool big = T > 8;    // kernels are instantiated for up to 8 and up to 16 tokens per prompt
  const int tok_smem = tok_smem_floats(big ? 16 : 8) * 4;
  const int t2i_smem = 8 * (big ? 16 : 8) * PART * 4;
  const int mlp_smem = mlp_smem_floats(big ? 16 : 8) * 4;
  static samhost::PerDeviceOnce attr_once;
  if (attr_once.need()) {
    if (int rc = opt_in_smem(dec_token_first_kernel<8>, tok_smem_floats(8) * 4)) return rc;
    if (int rc = opt_in_smem(dec_token_first_kernel<16>, tok_smem_floats(16) * 4)) return rc;
    if (int rc = opt_in_smem(dec_token_tail_kernel<8>, tok_smem_floats(8) * 4)) return rc;
    if (int rc = opt_in_smem(dec_token_tail_kernel<16>, tok_smem_floats(16) * 4)) return rc;
    if (int rc = opt_in_smem(dec_token_mlp_kernel<8>, mlp_smem_floats(8) * 4)) return rc;
    if (int rc = opt_in_smem(dec_token_mlp_kernel<16>, mlp_smem_floats(16) * 4)) return rc;
    if (int rc = opt_in_smem(dec_attn_t2i_kernel<8>, 8 * 8 * PART * 4)) return rc;
    if (int rc = opt_in_smem(dec_attn_t2i_kernel<16>, 8 * 16 * PART * 4)) return rc;
    attr_once.done();
  }
  // prompt -> source row block of the layer-0 image tokens: identity with per-prompt dense embeddings, img_index[p]
  // otherwise; img_index == NULL (one image, the reference's per-image call) makes every prompt read block 0
  const int* src0 = dense_full ? nullptr : img_index;
  const bool single_src = (!dense_full && !img_index);

  auto self_w = [&](int l) {
    const LayerW& L = w.layer[l];
    const DerivedW::Tok& t = dw.tok[l];
    SelfW sw{t.sq, L.self_attn.qb, t.sk, L.self_attn.kb, t.sv, L.self_attn.vb, t.so, L.self_attn.ob,
             L.n1w, L.n1b, t.cq, L.t2i.qb};
    return sw;
  };
  // image-side GEMM on the split operands: out[M, N] = A3[M, 3K] . W3[N, 3K]^T + bias   (fp32 out)
  auto gemm3 = [&](const uint16_t* A3, const uint16_t* W3, const float* bias, float* out, int M, int N, int K) -> int {
    samhost::ClassOverride as_decoder(samhost::KC_DECODER);
    GemmEpilogue ep{out, N, SAM_F32, bias, 0, nullptr, 0, 0};
    return samk_gemm(A3, 3 * K, W3, 3 * K, M, N, 3 * K, SAM_BF16, ep, st);
  };
  auto attn_t2i = [&](const float* kv, int ldkv, const float* rk, const int* idx, bool one_src) -> int {
    samhost::LaunchScope scope(samhost::KC_DECODER, st, 4.0 * n * T * HW * Ci);
    dim3 grid(NCH, n);
    if (big)
      dec_attn_t2i_kernel<16><<<grid, 256, t2i_smem, st>>>(ws.tq, kv, ldkv, rk, idx, ws.part, T, HW, one_src ? 1 : 0);
    else
      dec_attn_t2i_kernel<8><<<grid, 256, t2i_smem, st>>>(ws.tq, kv, ldkv, rk, idx, ws.part, T, HW, one_src ? 1 : 0);
    SAM_CHECK_CUDA(cudaGetLastError());
    return 0;
  };

  // ---- layer-0 image tokens (mask_decoder.py:146-149) and the first token sub-block (independent of each other)
  {
    dim3 grid(HW / 32, C / 32, n_src), blk(32, 8);
    samhost::LaunchScope scope(samhost::KC_DECODER, st, 0.0, static_cast<double>(n_src) * HW * C * (2.0 + 4.0 + 6.0));
    dec_keys0_kernel<<<grid, blk, 0, st>>>(image_embeddings, emb_fmt, img_index, dense_vec, dense_full, dense_fmt, ws.keys0,
                                           ws.a3, C, HW);
    SAM_CHECK_CUDA(cudaGetLastError());
  }
  {
    samhost::LaunchScope scope(samhost::KC_DECODER, st, 2.0 * n * T * (4.0 * C * C + C * Ci));
    const SelfW sw = self_w(0);
    if (big)
      dec_token_first_kernel<16><<<n, kTokThreads, tok_smem, st>>>(w.iou_token, w.mask_tokens, sparse, sparse_fmt, nm, k, sw, s.heads,
                                                          ws.tok0, ws.q, ws.tq);
    else
      dec_token_first_kernel<8><<<n, kTokThreads, tok_smem, st>>>(w.iou_token, w.mask_tokens, sparse, sparse_fmt, nm, k, sw, s.heads,
                                                         ws.tok0, ws.q, ws.tq);
    SAM_CHECK_CUDA(cudaGetLastError());
  }
  const float sc_cross = 1.0f / sqrtf(static_cast<float>(Ci / s.heads));
  const float* prev = ws.keys0;     // image tokens entering the layer
  int rows_src = n_src;             // row blocks of `prev` / of the merged GEMM's output
  const int* idx = src0;
  bool one = single_src;
  for (int l = 0; l < s.depth; ++l) {
    const LayerW& L = w.layer[l];
    const DerivedW::Tok& t = dw.tok[l];
    // (a) k_t2i | v_t2i | q_i2t of this layer's image tokens in one GEMM
    if (int rc = gemm3(ws.a3, dw.kvq[l], dw.b_kvq[l], ws.kvq, rows_src * HW, 3 * Ci, C)) return rc;
    // (b) token -> image attention, out-proj + norm2 + MLP, norm3 + i2t k/v + next self-attention block
    if (int rc = attn_t2i(ws.kvq, 3 * Ci, dw.rk[l], idx, one)) return rc;
    {
      samhost::LaunchScope scope(samhost::KC_DECODER, st, 2.0 * n * T * (2.0 * C * H + 8.0 * Ci * C));
      const CrossMlpW cw{t.co, L.t2i.ob, L.n2w, L.n2b, t.l1, L.l1b, t.l2};
      dim3 grid(H / 256, n);
      if (big)
        dec_token_mlp_kernel<16><<<grid, kTokThreads, mlp_smem, st>>>(ws.part, ws.q, cw, T, H, ws.x2, ws.pmlp);
      else
        dec_token_mlp_kernel<8><<<grid, kTokThreads, mlp_smem, st>>>(ws.part, ws.q, cw, T, H, ws.x2, ws.pmlp);
      SAM_CHECK_CUDA(cudaGetLastError());
    }
    {
      const bool has_next = l + 1 < s.depth;
      samhost::LaunchScope scope(samhost::KC_DECODER, st, 2.0 * n * T * (2.0 * C * Ci + (has_next ? 4.0 * C * C : 0.0) + C * Ci));
      const TailW tw{L.l2b, L.n3w, L.n3b, t.ikv, t.ikvb};
      SelfW nw;
      if (has_next) {
        nw = self_w(l + 1);
      } else {
        nw = SelfW{};
        nw.cqw = dw.fq;
        nw.cqb = w.final_attn.qb;
      }
      if (big)
        dec_token_tail_kernel<16><<<n, kTokThreads, tok_smem, st>>>(ws.x2, ws.pmlp, H / 256, tw, ws.tok0, T, s.heads, ws.tk, ws.tv,
                                                           has_next ? 1 : 0, nw, ws.q, ws.tq);
      else
        dec_token_tail_kernel<8><<<n, kTokThreads, tok_smem, st>>>(ws.x2, ws.pmlp, H / 256, tw, ws.tok0, T, s.heads, ws.tk, ws.tv,
                                                          has_next ? 1 : 0, nw, ws.q, ws.tq);
      SAM_CHECK_CUDA(cudaGetLastError());
    }
    // (c) image -> token attention, out-proj GEMM, norm4 (+ the split operand of the next merged GEMM)
    {
      samhost::LaunchScope scope(samhost::KC_DECODER, st, 4.0 * n * T * HW * Ci);
      dim3 grid(HW / 64, n);
      if (big)
        dec_attn_i2t_kernel<16><<<grid, 256, 0, st>>>(ws.kvq, 3 * Ci, 2 * Ci, dw.rq[l], idx, ws.tk, ws.tv, ws.a3o, T, HW,
                                                     sc_cross, one ? 1 : 0);
      else
        dec_attn_i2t_kernel<8><<<grid, 256, 0, st>>>(ws.kvq, 3 * Ci, 2 * Ci, dw.rq[l], idx, ws.tk, ws.tv, ws.a3o, T, HW,
                                                    sc_cross, one ? 1 : 0);
      SAM_CHECK_CUDA(cudaGetLastError());
    }
    if (int rc = gemm3(ws.a3o, dw.i2t_o[l], L.i2t.ob, ws.delta, n * HW, C, Ci)) return rc;
    {
      const size_t rows = static_cast<size_t>(n) * HW;
      samhost::LaunchScope scope(samhost::KC_DECODER, st, 0.0, static_cast<double>(rows) * C * (4.0 + 4.0 + 4.0 + 6.0));
      dec_ln_split_kernel<<<static_cast<unsigned>((rows + 7) / 8), 256, 0, st>>>(ws.delta, prev, idx, HW, L.n4w, L.n4b, ws.keys,
                                                                                ws.a3, rows, one ? 1 : 0);
      SAM_CHECK_CUDA(cudaGetLastError());
    }
    prev = ws.keys;       // from here on every prompt owns its image tokens
    rows_src = n;
    idx = nullptr;
    one = false;
  }
  // ---- final token -> image attention + upscaling input in one GEMM: k_final | v_final | ConvT0
  if (int rc = gemm3(ws.a3, dw.fin, dw.b_fin, ws.kvq, n * HW, 2 * Ci + C, C)) return rc;
  if (int rc = attn_t2i(ws.kvq, 2 * Ci + C, dw.rk[s.depth], nullptr, false)) return rc;
  {
    HyperArgs h;
    for (int i = 0; i < nm; ++i) h.mlp[i] = w.hyper[i];
    h.mlp[nm] = w.iou_head;
    h.nm = nm; h.C = C; h.hidden_iou = s.iou_hidden; h.out_hyper = C / 8; h.T = T;
    h.q = ws.q; h.part = ws.part; h.ow = dw.fo; h.ob = w.final_attn.ob; h.nfw = w.nfw; h.nfb = w.nfb;
    h.hyper = ws.hyper; h.iou = iou; h.iou_fmt = out_fmt;
    dim3 grid(nm + 1, n);
    samhost::LaunchScope scope(samhost::KC_DECODER, st, 2.0 * n * (nm * (2.0 * C * C + C * C / 8) + 2.0 * C * C));
    dec_hyper_kernel<<<grid, 256, 0, st>>>(h);
    SAM_CHECK_CUDA(cudaGetLastError());
  }
  {
    // upscaling tail: prologue -> tensor-core GEMM (+ GELU) -> hypernetwork product.  a3 is free again (the final merged
    // GEMM has consumed it); z = keys | delta, which are adjacent in the workspace and dead by now
    const size_t rows4 = static_cast<size_t>(n) * HW * 4;
    SAM_REQUIRE(C / 4 == UC1 && C / 8 == UC2 && HW % 16 == 0, "mask decoder: upscaling widths / grid");
    SAM_REQUIRE(reinterpret_cast<uint8_t*>(ws.keys) + static_cast<size_t>(n) * HW * C * 4 == reinterpret_cast<uint8_t*>(ws.delta),
                "mask decoder: workspace layout (keys | delta must be adjacent)");
    float* z = ws.keys;
    {
      samhost::LaunchScope scope(samhost::KC_DECODER, st, 0.0, static_cast<double>(rows4) * UC1 * (4.0 + 6.0));
      dec_up_prologue_kernel<<<static_cast<unsigned>(rows4 * 16 / 256), 256, 0, st>>>(ws.kvq + 2 * Ci, 2 * Ci + C, w.upln_w, w.upln_b,
                                                                                  ws.a3, rows4);
      SAM_CHECK_CUDA(cudaGetLastError());
    }
    {
      samhost::ClassOverride as_decoder(samhost::KC_DECODER);
      GemmEpilogue ep{z, 4 * UC2, SAM_F32, dw.b_up1, 1, nullptr, 0, 0};
      if (int rc = samk_gemm(ws.a3, 3 * UC1, dw.up1, 3 * UC1, static_cast<int>(rows4), 4 * UC2, 3 * UC1, SAM_BF16, ep, st)) return rc;
    }
    {
      samhost::LaunchScope scope(samhost::KC_DECODER, st, 2.0 * rows4 * 4 * UC2 * nm, static_cast<double>(rows4) * 4 * UC2 * 4.0);
      dec_up_hyper_dot_kernel<<<static_cast<unsigned>(rows4 / 64), 256, 0, st>>>(z, ws.hyper, masks, out_fmt, g, nm);
      SAM_CHECK_CUDA(cudaGetLastError());
    }
  }
  return 0;
}
