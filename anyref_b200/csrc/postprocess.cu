// Sam.postprocess_masks (sam.py:159-172) as ONE kernel: bilinear L x L -> S x S (align_corners=False), crop to
// input_size, bilinear -> original_size, optional thresholding (Sam.mask_threshold, sam.py:19;
// eval_referseg.py:191 sigmoid(x) > 0.5 == x > 0).  The [n,C,S,S] intermediate of the reference never exists.
// HBM traffic = read L*L*4 + write H*W*4 bytes per mask (the algorithmic minimum).
//
// "Column walker": a thread owns ONE output column X and walks a strip of output rows.  Both resizes are separable
// in exactly the order ATen evaluates them (horizontal lerp of the two source rows, then the vertical lerp), so the
// thread keeps, in registers,
//   * the horizontal lerps h(i) of the two most recent LOW-RES rows at its two stage-1 columns (4 loads per new
//     low-res row -- one every four stage-1 rows),
//   * the stage-2 horizontal lerps g(y) of the two most recent STAGE-1 rows y,
// and an output pixel costs one vertical lerp + one coalesced store.  All cache decisions depend on the row only, so
// they are warp-uniform; arbitrary scales (up- and down-sampling, crop) take the same code and simply reuse less.
// Round 1 re-derived 4 stage-1 values (16 taps, 16 loads) per output pixel: LSU-bound at 5 % of the HBM rate.
//
// Tap arithmetic follows ATen's area_pixel_compute_source_index (scale = in/out in fp32,
// src = max(0, scale*(dst+0.5)-0.5) evaluated as ONE fused multiply-add, which is what both ATen's vectorised CPU
// kernel and its CUDA kernel compile to; i1 = i0 + (i0 < in-1)), so tap INDICES are exact with the reference and
// values agree to fp32 rounding (products and sums are kept un-fused in ATen's order).
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

// l0 * a + l1 * b with separately rounded products (ATen's order; no FMA contraction)
__device__ __forceinline__ float lerp1(float l0, float a, float l1, float b) {
  return __fadd_rn(__fmul_rn(l0, a), __fmul_rn(l1, b));
}

// FMT is a template parameter of the kernel: with a run-time format every load sat in its own basic block behind a
// three-way branch, so the 4 NC loads of a new low-res row were issued one after the other (no memory-level parallelism)
template <int FMT>
__device__ __forceinline__ float ld(const void* p, size_t i) {
  if (FMT == 2) return __ldg(static_cast<const float*>(p) + i);
  return ptx::unpack1(__ldg(static_cast<const uint16_t*>(p) + i), FMT);
}

constexpr int kPostRows = 32;      // output rows per CTA strip
constexpr int kPostThreads = 128;  // a thread owns NC columns, 32 apart (lane-contiguous: coalesced stores)

// grid (ceil(W / (NC * 128)), ceil(H / 32), num_masks), block 128.  Warp w of a CTA covers columns
// [x_cta + w * 32 * NC, + 32 * NC): column j of lane l is x_warp + 32 j + l.  The row-dependent work (row taps, cache
// decisions, loop overhead) is shared by the NC columns of a thread -- with one column per thread the kernel was
// issue-bound on exactly that work (ncu r02: SM 82 %, DRAM 2 %).
// IDENT: input_size == original_size, i.e. the second resize is the identity (every stage-2 tap is (i, i, 1, 0)): the
// output IS the stage-1 image, so the stage-2 lerps, the second column taps and half of the low-res loads drop out --
// values equal to the general path's (l0 = 1, l1 = 0 exactly), which is the shape the throughput bench and square
// inputs run.
// With `target` / `counts` the intersectionAndUnionGPU statistics of the thresholded mask (utils/utils.py:79-91, K = 2,
// ignore_index = 255) are accumulated in the same pass: counts[m] = {inter_0, inter_1, pred_0, pred_1, target_0,
// target_1} (int32, atomically added), so the evaluation loop (eval_referseg.py:186-211) needs neither the full
// resolution logits nor the mask in HBM.
template <int NC, int FMT, bool IDENT>
__global__ void __launch_bounds__(kPostThreads)
postprocess_kernel(const void* __restrict__ low, int L, int S, int h_in, int w_in, int H, int W,
                   float* __restrict__ logits, uint8_t* __restrict__ binary, uint8_t* __restrict__ packed, float threshold,
                   const uint8_t* __restrict__ target, int* __restrict__ counts) {
  const int m = blockIdx.z;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int xw = (blockIdx.x * (kPostThreads / 32) + warp) * (32 * NC);     // first column of this warp
  const int Y0 = blockIdx.y * kPostRows;
  const int Y1 = min(Y0 + kPostRows, H);
  const float scale1 = static_cast<float>(L) / static_cast<float>(S);
  const float sy = static_cast<float>(h_in) / static_cast<float>(H);
  const float sx = static_cast<float>(w_in) / static_cast<float>(W);
  const size_t base = static_cast<size_t>(m) * L * L;
  // column taps: stage 2 (X -> stage-1 columns x0, x1), stage 1 (x0, x1 -> low-res columns)
  bool col_ok[NC];
  Tap tx[NC], ca[NC], cb[NC];
#pragma unroll
  for (int j = 0; j < NC; ++j) {
    const int Xr = xw + 32 * j + lane;
    col_ok[j] = Xr < W;
    tx[j] = make_tap(col_ok[j] ? Xr : W - 1, sx, w_in);    // out-of-range columns shadow the last one (no stores)
    ca[j] = make_tap(tx[j].i0, scale1, L);
    cb[j] = IDENT ? ca[j] : make_tap(tx[j].i1, scale1, L);
  }
  // low-res row cache (rows hr0, hr1): horizontal lerps at stage-1 columns x0 (a) and x1 (b) of every column
  int hr0 = -1, hr1 = -1;
  float ha0[NC], hb0[NC], ha1[NC], hb1[NC];
  auto load_h = [&](int i, float (&ha)[NC], float (&hb)[NC]) {
    const size_t r = base + static_cast<size_t>(i) * L;
#pragma unroll
    for (int j = 0; j < NC; ++j) {
      ha[j] = lerp1(ca[j].l0, ld<FMT>(low, r + ca[j].i0), ca[j].l1, ld<FMT>(low, r + ca[j].i1));
      if (!IDENT) hb[j] = lerp1(cb[j].l0, ld<FMT>(low, r + cb[j].i0), cb[j].l1, ld<FMT>(low, r + cb[j].i1));
    }
  };
  // stage-2 horizontal lerp g(y) of stage-1 row y (all branches depend on y only: warp-uniform)
  auto stage1_row = [&](int y, float (&g)[NC]) {
    const Tap t = make_tap(y, scale1, L);
    if (t.i0 == hr1) {
      hr0 = hr1;
#pragma unroll
      for (int j = 0; j < NC; ++j) {
        ha0[j] = ha1[j];
        if (!IDENT) hb0[j] = hb1[j];
      }
      hr1 = -1;
    }
    if (t.i0 != hr0) {
      load_h(t.i0, ha0, hb0);
      hr0 = t.i0;
    }
    if (t.i1 != t.i0) {
      if (t.i1 != hr1) {
        load_h(t.i1, ha1, hb1);
        hr1 = t.i1;
      }
#pragma unroll
      for (int j = 0; j < NC; ++j) {
        const float sa = lerp1(t.l0, ha0[j], t.l1, ha1[j]);     // stage-1 value at (y, x0)
        if (IDENT) {
          g[j] = sa;
        } else {
          const float sb = lerp1(t.l0, hb0[j], t.l1, hb1[j]);   // stage-1 value at (y, x1)
          g[j] = lerp1(tx[j].l0, sa, tx[j].l1, sb);
        }
      }
    } else {
#pragma unroll
      for (int j = 0; j < NC; ++j) {
        const float sa = lerp1(t.l0, ha0[j], t.l1, ha0[j]);
        if (IDENT) {
          g[j] = sa;
        } else {
          const float sb = lerp1(t.l0, hb0[j], t.l1, hb0[j]);
          g[j] = lerp1(tx[j].l0, sa, tx[j].l1, sb);
        }
      }
    }
  };
  int gr0 = -1, gr1 = -1;
  float g0[NC], g1[NC];
  int cnt[6] = {0, 0, 0, 0, 0, 0};
#pragma unroll 1
  for (int Y = Y0; Y < Y1; ++Y) {
    bool two = false;
    Tap ty;
    if (IDENT) {
      stage1_row(Y, g0);           // the stage-2 taps of row Y are (Y, Y, 1, 0): the output row is stage-1 row Y
      ty.l0 = 1.0f;
      ty.l1 = 0.0f;
    } else {
      ty = make_tap(Y, sy, h_in);
      if (ty.i0 == gr1) {
        gr0 = gr1;
#pragma unroll
        for (int j = 0; j < NC; ++j) g0[j] = g1[j];
        gr1 = -1;
      }
      if (ty.i0 != gr0) {
        stage1_row(ty.i0, g0);
        gr0 = ty.i0;
      }
      two = ty.i1 != ty.i0;
      if (two && ty.i1 != gr1) {
        stage1_row(ty.i1, g1);
        gr1 = ty.i1;
      }
    }
    const size_t orow = (static_cast<size_t>(m) * H + Y) * W;
#pragma unroll
    for (int j = 0; j < NC; ++j) {
      const float v = IDENT ? g0[j] : lerp1(ty.l0, g0[j], ty.l1, two ? g1[j] : g0[j]);
      const bool on = v > threshold;
      const int Xr = xw + 32 * j + lane;
      const size_t o = orow + Xr;
      if (col_ok[j]) {
        if (logits) logits[o] = v;
        if (binary) binary[o] = on ? 1 : 0;
        if (counts) {
          const int t = target[o];
          if (t != 255) {
            const int p = on ? 1 : 0;          // static indices only: the counters stay in registers
            cnt[2] += (p == 0); cnt[3] += (p == 1);
            cnt[4] += (t == 0); cnt[5] += (t == 1);
            cnt[0] += (t == 0 && p == 0); cnt[1] += (t == 1 && p == 1);
          }
        }
      }
      if (packed) {
        // numpy.packbits order: pixel 8k of the flattened [num_masks, H, W] stream is the MSB of byte k.  The 32 lanes
        // hold 32 consecutive columns = 4 bytes; W % 8 == 0 (checked by the launcher), so a byte is inside the row or
        // outside it.
        const uint32_t bits = __ballot_sync(0xffffffffu, on && col_ok[j]);
        const uint32_t bytes = __byte_perm(__brev(bits), 0, 0x0123);
        const int xb = xw + 32 * j + 8 * lane;     // first column of byte `lane` (lanes 0..3)
        if (lane < 4 && xb < W) packed[(orow + xb) >> 3] = static_cast<uint8_t>(bytes >> (8 * lane));
      }
    }
  }
  if (counts) {
    __shared__ int s_cnt[6];
    const int tid = threadIdx.x;
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
  const int nc = (W > 256) ? 4 : 1;    // columns per thread (narrow outputs: one, so that the columns spread over warps)
  dim3 blk(kPostThreads);
  dim3 grid((W + nc * kPostThreads - 1) / (nc * kPostThreads), (H + kPostRows - 1) / kPostRows, num_masks);
  samhost::LaunchScope scope(samhost::KC_POSTPROCESS, stream, 0.0,
                             static_cast<double>(num_masks) *
                                 (static_cast<double>(L) * L * (low_fmt == 2 ? 4.0 : 2.0) +
                                  static_cast<double>(H) * W * ((logits ? 4.0 : 0.0) + (binary ? 1.0 : 0.0) + (packed ? 0.125 : 0.0) + (target ? 1.0 : 0.0))));
  typedef void (*Fn)(const void*, int, int, int, int, int, int, float*, uint8_t*, uint8_t*, float, const uint8_t*, int*);
  static const Fn fns[2][2][3] = {
      {{postprocess_kernel<1, 0, false>, postprocess_kernel<1, 1, false>, postprocess_kernel<1, 2, false>},
       {postprocess_kernel<4, 0, false>, postprocess_kernel<4, 1, false>, postprocess_kernel<4, 2, false>}},
      {{postprocess_kernel<1, 0, true>, postprocess_kernel<1, 1, true>, postprocess_kernel<1, 2, true>},
       {postprocess_kernel<4, 0, true>, postprocess_kernel<4, 1, true>, postprocess_kernel<4, 2, true>}}};
  const bool ident = (h_in == H && w_in == W);
  fns[ident][nc == 4][low_fmt]<<<grid, blk, 0, stream>>>(low, L, S, h_in, w_in, H, W, logits, binary, packed, threshold,
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
