// Sam.postprocess_masks (sam.py:159-172) as ONE kernel: bilinear L x L -> S x S (align_corners=False), crop to
// input_size, bilinear -> original_size, optional thresholding (Sam.mask_threshold, sam.py:19;
// eval_referseg.py:191 sigmoid(x) > 0.5 == x > 0).  The [n,C,S,S] intermediate of the reference never exists: every
// output pixel evaluates its 4 stage-2 taps on the fly, each from 4 low-resolution taps (the 256 KB low-res mask
// stays in L1/L2).  HBM traffic = read L*L*4 + write H*W*4 bytes per mask (the algorithmic minimum).
//
// Tap arithmetic follows ATen's area_pixel_compute_source_index (scale = in/out in fp32,
// src = max(0, scale*(dst+0.5)-0.5) evaluated as ONE fused multiply-add, which is what both ATen's vectorised CPU
// kernel and its CUDA kernel compile to; i1 = i0 + (i0 < in-1)), so tap INDICES are exact with the reference and
// values agree to fp32 rounding.
#include "host_common.h"
#include "kernels.h"
#include "ptx.cuh"

namespace {

struct Tap {
  int i0, i1;
  float l0, l1;
};

__device__ __forceinline__ Tap make_tap(int dst, float scale, int in_size) {
  const float src = fmaxf(fmaf(scale, static_cast<float>(dst) + 0.5f, -0.5f), 0.0f);
  Tap t;
  t.i0 = static_cast<int>(src);
  t.i1 = t.i0 + (t.i0 < in_size - 1 ? 1 : 0);
  t.l1 = __fsub_rn(src, static_cast<float>(t.i0));
  t.l0 = __fsub_rn(1.0f, t.l1);
  return t;
}

__device__ __forceinline__ float lerp2(float a00, float a01, float a10, float a11, const Tap& ty, const Tap& tx) {
  const float top = __fadd_rn(__fmul_rn(tx.l0, a00), __fmul_rn(tx.l1, a01));
  const float bot = __fadd_rn(__fmul_rn(tx.l0, a10), __fmul_rn(tx.l1, a11));
  return __fadd_rn(__fmul_rn(ty.l0, top), __fmul_rn(ty.l1, bot));
}

__device__ __forceinline__ float ld(const void* p, int fmt, size_t i) {
  if (fmt == 2) return __ldg(static_cast<const float*>(p) + i);
  return ptx::unpack1(__ldg(static_cast<const uint16_t*>(p) + i), fmt);
}

// value of the S x S stage-1 image at (yy, xx)
__device__ __forceinline__ float stage1(const void* low, int fmt, size_t base, int L, float scale1, int yy, int xx) {
  const Tap ty = make_tap(yy, scale1, L), tx = make_tap(xx, scale1, L);
  const float a00 = ld(low, fmt, base + static_cast<size_t>(ty.i0) * L + tx.i0);
  const float a01 = ld(low, fmt, base + static_cast<size_t>(ty.i0) * L + tx.i1);
  const float a10 = ld(low, fmt, base + static_cast<size_t>(ty.i1) * L + tx.i0);
  const float a11 = ld(low, fmt, base + static_cast<size_t>(ty.i1) * L + tx.i1);
  return lerp2(a00, a01, a10, a11, ty, tx);
}

// grid (ceil(W/128), ceil(H/8), n*C), block (128, 1): each thread produces 8 rows? -> keep it simple: 1 pixel/thread
// with 4-wide vector stores when W % 4 == 0.
// With `target` / `counts` the intersectionAndUnionGPU statistics of the thresholded mask (utils/utils.py:79-91, K = 2,
// ignore_index = 255) are accumulated in the same pass: counts[m] = {inter_0, inter_1, pred_0, pred_1, target_0,
// target_1} (int32, atomically added), so the evaluation loop (eval_referseg.py:186-211) needs neither the full
// resolution logits nor the mask in HBM.
__global__ void __launch_bounds__(256)
postprocess_kernel(const void* __restrict__ low, int low_fmt, int L, int S, int h_in, int w_in, int H, int W,
                   float* __restrict__ logits, uint8_t* __restrict__ binary, uint8_t* __restrict__ packed, float threshold,
                   const uint8_t* __restrict__ target, int* __restrict__ counts) {
  const int m = blockIdx.z;
  const int Y = blockIdx.y * blockDim.y + threadIdx.y;
  const int X4 = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
  const bool active = (Y < H && X4 < W);
  if (!active && counts == nullptr && packed == nullptr) return;
  int cnt[6] = {0, 0, 0, 0, 0, 0};
  uint32_t nib = 0;      // this thread's 4 thresholded pixels, first pixel in bit 3 (numpy.packbits order)
  if (active) {
  const float scale1 = static_cast<float>(L) / static_cast<float>(S);
  const float sy = static_cast<float>(h_in) / static_cast<float>(H);
  const float sx = static_cast<float>(w_in) / static_cast<float>(W);
  const size_t base = static_cast<size_t>(m) * L * L;
  const Tap ty = make_tap(Y, sy, h_in);
  float v[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int X = X4 + i;
    v[i] = 0.f;
    if (X < W) {
      const Tap tx = make_tap(X, sx, w_in);
      const float a00 = stage1(low, low_fmt, base, L, scale1, ty.i0, tx.i0);
      const float a01 = stage1(low, low_fmt, base, L, scale1, ty.i0, tx.i1);
      const float a10 = stage1(low, low_fmt, base, L, scale1, ty.i1, tx.i0);
      const float a11 = stage1(low, low_fmt, base, L, scale1, ty.i1, tx.i1);
      v[i] = lerp2(a00, a01, a10, a11, ty, tx);
    }
  }
  const size_t o = (static_cast<size_t>(m) * H + Y) * W + X4;
  if (logits) {
    if ((W & 3) == 0) {
      *reinterpret_cast<float4*>(logits + o) = make_float4(v[0], v[1], v[2], v[3]);
    } else {
#pragma unroll
      for (int i = 0; i < 4; ++i)
        if (X4 + i < W) logits[o + i] = v[i];
    }
  }
  if (binary) {
    if ((W & 3) == 0) {
      uchar4 b;
      b.x = v[0] > threshold; b.y = v[1] > threshold; b.z = v[2] > threshold; b.w = v[3] > threshold;
      *reinterpret_cast<uchar4*>(binary + o) = b;
    } else {
#pragma unroll
      for (int i = 0; i < 4; ++i)
        if (X4 + i < W) binary[o + i] = v[i] > threshold;
    }
  }
  if (packed)
    nib = (v[0] > threshold ? 8u : 0u) | (v[1] > threshold ? 4u : 0u) | (v[2] > threshold ? 2u : 0u) | (v[3] > threshold ? 1u : 0u);
  if (counts) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      if (X4 + i < W) {
        const int t = target[o + i];
        if (t != 255) {
          const int p = v[i] > threshold ? 1 : 0;
          cnt[2 + p] += 1;
          if (t < 2) {
            cnt[4 + t] += 1;
            if (t == p) cnt[p] += 1;
          }
        }
      }
    }
  }
  }  // active
  if (packed) {
    // W % 8 == 0 (checked by the launcher): lanes 2k / 2k+1 hold the high / low nibble of one byte of the flattened
    // [num_masks, H, W] bit stream, both inside the row or both outside it
    const uint32_t lo = __shfl_down_sync(0xffffffffu, nib, 1);
    if (active && (threadIdx.x & 1) == 0)
      packed[((static_cast<size_t>(m) * H + Y) * W + X4) >> 3] = static_cast<uint8_t>((nib << 4) | lo);
  }
  if (counts) {
    __shared__ int s_cnt[6];
    const int tid = threadIdx.y * blockDim.x + threadIdx.x;
    if (tid < 6) s_cnt[tid] = 0;
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 6; ++k) {
      const int w = __reduce_add_sync(0xffffffffu, cnt[k]);
      if ((tid & 31) == 0 && w) atomicAdd(&s_cnt[k], w);
    }
    __syncthreads();
    if (tid < 6 && s_cnt[tid]) atomicAdd(&counts[m * 6 + tid], s_cnt[tid]);
  }
}

// Folds per-mask counts into the running evaluation statistics (eval_referseg.py:197-211):
//   stats[0:2] += intersection, stats[2:4] += union, stats[4:6] += intersection / (union + 1e-5) (+1 where union == 0:
//   "no-object target"), stats[6] += 1 per mask.  fp64 accumulation in mask order: deterministic.
__global__ void iou_finalize_kernel(const int* __restrict__ counts, int n, double* __restrict__ stats) {
  if (blockIdx.x != 0 || threadIdx.x != 0) return;
  double acc[7] = {0, 0, 0, 0, 0, 0, 0};
  for (int m = 0; m < n; ++m) {
    for (int k = 0; k < 2; ++k) {
      const float inter = static_cast<float>(counts[m * 6 + k]);
      const float uni = static_cast<float>(counts[m * 6 + 2 + k] + counts[m * 6 + 4 + k] - counts[m * 6 + k]);
      float a = inter / (uni + 1e-5f);   // fp32 like the torch tensors of the reference
      if (uni == 0.0f) a += 1.0f;
      acc[k] += inter;
      acc[2 + k] += uni;
      acc[4 + k] += a;
    }
    acc[6] += 1.0;
  }
  for (int k = 0; k < 7; ++k) stats[k] += acc[k];
}

}  // namespace

int samk_postprocess(const void* low, int low_fmt, int num_masks, int L, int S, int h_in, int w_in, int H, int W,
                     float* logits, uint8_t* binary, float threshold, cudaStream_t stream) {
  return samk_postprocess_iou(low, low_fmt, num_masks, L, S, h_in, w_in, H, W, logits, binary, nullptr, threshold, nullptr,
                              nullptr, stream);
}

int samk_iou_finalize(const int* counts, int n, double* stats, cudaStream_t stream) {
  SAM_REQUIRE(counts && stats && n >= 0, "iou_finalize: bad arguments");
  if (n == 0) return 0;
  samhost::LaunchScope scope(samhost::KC_POSTPROCESS, stream);
  iou_finalize_kernel<<<1, 32, 0, stream>>>(counts, n, stats);
  SAM_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int samk_postprocess_iou(const void* low, int low_fmt, int num_masks, int L, int S, int h_in, int w_in, int H, int W,
                         float* logits, uint8_t* binary, uint8_t* packed, float threshold, const uint8_t* target,
                         int* counts, cudaStream_t stream) {
  SAM_REQUIRE(num_masks > 0 && L > 0 && S > 0 && H > 0 && W > 0, "postprocess: empty problem");
  SAM_REQUIRE(h_in > 0 && w_in > 0 && h_in <= S && w_in <= S, "postprocess: input_size (%d,%d) outside the %d canvas",
              h_in, w_in, S);
  SAM_REQUIRE(logits || binary || counts || packed, "postprocess: no output requested");
  SAM_REQUIRE(!packed || W % 8 == 0, "postprocess: the bit-packed output needs W %% 8 == 0 (got %d)", W);
  SAM_REQUIRE((target == nullptr) == (counts == nullptr), "postprocess: target and counts go together");
  SAM_REQUIRE(low_fmt >= 0 && low_fmt <= 2, "postprocess: bad input format");
  SAM_REQUIRE(num_masks <= 65535, "postprocess: at most 65535 masks per call");
  dim3 blk(64, 4);
  dim3 grid(((W + 3) / 4 + blk.x - 1) / blk.x, (H + blk.y - 1) / blk.y, num_masks);
  samhost::LaunchScope scope(samhost::KC_POSTPROCESS, stream, 0.0,
                             static_cast<double>(num_masks) *
                                 (static_cast<double>(L) * L * (low_fmt == 2 ? 4.0 : 2.0) +
                                  static_cast<double>(H) * W * ((logits ? 4.0 : 0.0) + (binary ? 1.0 : 0.0) + (packed ? 0.125 : 0.0) + (target ? 1.0 : 0.0))));
  postprocess_kernel<<<grid, blk, 0, stream>>>(low, low_fmt, L, S, h_in, w_in, H, W, logits, binary, packed, threshold,
                                               target, counts);
  SAM_CHECK_CUDA(cudaGetLastError());
  return 0;
}

// PromptEncoder.get_dense_pe (prompt_encoder.py:67-76, PositionEmbeddingRandom :203-219):
//   pe[c, y, x] = sin | cos ( 2*pi * ( (2*(x+0.5)/g - 1) * G[0,c'] + (2*(y+0.5)/g - 1) * G[1,c'] ) ),  c' = c mod C/2
namespace {
__global__ void dense_pe_kernel(const float* __restrict__ gauss, void* __restrict__ out, int out_fmt, int C, int g) {
  const int half = C / 2;
  const int total = C * g * g;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int x = i % g, y = (i / g) % g, c = i / (g * g);
    const int cc = c % half;
    const float cx = __fsub_rn(__fmul_rn(2.0f, (static_cast<float>(x) + 0.5f) / static_cast<float>(g)), 1.0f);
    const float cy = __fsub_rn(__fmul_rn(2.0f, (static_cast<float>(y) + 0.5f) / static_cast<float>(g)), 1.0f);
    const float d = __fadd_rn(__fmul_rn(cx, gauss[cc]), __fmul_rn(cy, gauss[half + cc]));
    const float a = __fmul_rn(6.283185307179586f, d);
    const float v = (c < half) ? sinf(a) : cosf(a);
    if (out_fmt == 2)
      static_cast<float*>(out)[i] = v;
    else
      static_cast<uint16_t*>(out)[i] = ptx::pack1(v, out_fmt);
  }
}
}  // namespace

int samk_dense_pe(const float* gauss, void* out, int out_fmt, int C, int g, cudaStream_t stream) {
  SAM_REQUIRE(C % 2 == 0 && g > 0, "dense_pe: bad shape");
  samhost::LaunchScope scope(samhost::KC_LAYOUT, stream);
  dense_pe_kernel<<<(C * g * g + 255) / 256, 256, 0, stream>>>(gauss, out, out_fmt, C, g);
  SAM_CHECK_CUDA(cudaGetLastError());
  return 0;
}
