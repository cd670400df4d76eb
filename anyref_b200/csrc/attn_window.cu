// 14x14 windowed attention of the SAM ViT-H encoder with the decomposed relative-position bias fused into the
// softmax -- replaces image_encoder.py:235-257 (Attention.forward minus qkv/proj), :263-318 (window partition /
// unpartition, done here by index arithmetic + TMA boxes) and :354-392 (add_decomposed_rel_pos).
//
// Input  qkv  [B*64*64, 3*E] operand format (fp16/bf16), token-major, UN-partitioned; columns (which, head, d).
// Output out  [B*64*64, E]   operand format, token-major, heads merged -- the proj GEMM's A operand.
//
// One CTA (128 threads, 2 CTAs/SM) = one (image, window, head, query tile); query tile 0 = window rows 0..8
// (126 tokens), tile 1 = rows 9..13 (70 tokens).  All matrix products run on tcgen05 with TMEM accumulators:
//   Th = Q.Rh^T, Tw = Q.Rw^T  (N=32 each: the 27 rel_pos_h / rel_pos_w rows) -> per-row bias look-up tables
//   S = Q.K^T   (N=208: 196 keys padded to a multiple of 16)
//   O = P.V     (N=64 + N=16, V consumed MN-major straight from the TMA tile)
// Q/K/V tiles arrive by 4-D TMA boxes over the [B,64,64,3E] view (128B- and 32B-swizzled for the 64+16 split of
// head_dim 80).  Window padding (image_encoder.py:277-283 pads AFTER norm1, so padded tokens have q/k/v == qkv bias
// and DO take part in the softmax) is reproduced exactly: TMA zero-fills the out-of-image rows and the kernel
// overwrites them with the bias; padded queries are computed but never stored.
// Softmax: one thread per query row (TMEM lane), fp32, exp2 with the scale folded in; P is rounded to the operand
// format, the row sum is applied to O in fp32.
#include <stdlib.h>

#include "host_common.h"
#include "kernels.h"
#include "ptx.cuh"

namespace {

constexpr int HD = 80;
constexpr int WS = 14;
constexpr int NTOK = WS * WS;  // 196
constexpr int NKEY = 208;      // keys padded to the UMMA N/K granularity
constexpr int kThreads = 256;  // 8 warps: warp w and w+4 share TMEM lanes 32*(w&3).. and split the key columns

// shared-memory map (bytes from the 1024-aligned base)
constexpr int OFF_Q64 = 0;       // 128 x 128B  SWIZZLE_128B
constexpr int OFF_K64 = 16384;   // 208 x 128B
constexpr int OFF_R64 = 43008;   //  64 x 128B
constexpr int OFF_Q16 = 51200;   // 128 x 32B   SWIZZLE_32B
constexpr int OFF_K16 = 55296;   // 208 x 32B
constexpr int OFF_R16 = 61952;   //  64 x 32B
constexpr int OFF_V64 = 64512;   // 208 x 128B  (MN-major operand of P.V); loaded AFTER the rel-pos phase
constexpr int OFF_V16 = 91136;   // 208 x 32B
constexpr int OFF_XCH = 97792;   // 2 x (2 x 128) fp32: row max and row sum exchange between the two column halves
constexpr int OFF_BAR = 99840;
constexpr int kSmemBytes = OFF_BAR + 64 + 1024;
// overlays
constexpr int OFF_P = 0;         // P: 3 x (128 x 128B) SWIZZLE_128B + 128 x 32B SWIZZLE_32B at +49152 (over Q/K/R)
constexpr int OFF_P16 = 49152;
constexpr int OFF_TH = OFF_V64;           // 128 x 27 fp32: q . rel_pos_h rows   (scratch, before V is loaded)
constexpr int OFF_TW = OFF_V64 + 13824;   // 128 x 27 fp32: q . rel_pos_w rows

struct WinAttnMaps {
  CUtensorMap kv64, kv16;    // box {64|16, 14, 14, 1}
  CUtensorMap qa64, qa16;    // box {64|16, 14, 9, 1}   query tile 0
  CUtensorMap qb64, qb16;    // box {64|16, 14, 5, 1}   query tile 1
  CUtensorMap r64, r16;      // rel-pos operand table [64, 80]: box {64|16, 64}
};

__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

__device__ __forceinline__ uint32_t row_off64(int r, int c) { return r * 128 + ((c ^ (r & 7)) << 4); }
__device__ __forceinline__ uint32_t row_off16(int r, int c) { return r * 32 + ((c ^ ((r >> 2) & 1)) << 4); }

// copies the 80-element operand-format vector `src` into row r of a (64 + 16)-split K-major tile
__device__ __forceinline__ void fill_row(uint8_t* t64, uint8_t* t16, int r, const uint16_t* __restrict__ src) {
#pragma unroll
  for (int c = 0; c < 10; ++c) {
    const uint4 v = __ldg(reinterpret_cast<const uint4*>(src) + c);
    if (c < 8)
      *reinterpret_cast<uint4*>(t64 + row_off64(r, c)) = v;
    else
      *reinterpret_cast<uint4*>(t16 + row_off16(r, c - 8)) = v;
  }
}

// Loads chunk c (32 key columns; chunk 6 holds the last 16) of this thread's S row from TMEM.
template <int C>
__device__ __forceinline__ void load_s_chunk(uint32_t trow, uint32_t (&v)[32]) {
  if (C < 6) {
    ptx::tmem_ld_32x32b_x32(trow + C * 32, v);
  } else {
    uint32_t t16[16];
    ptx::tmem_ld_32x32b_x16(trow + 192, t16);
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = t16[i];
#pragma unroll
    for (int i = 16; i < 32; ++i) v[i] = 0;
  }
  ptx::tmem_ld_wait();
}

// logit (log2 domain) of key column J:  s * scale*log2e + rel_h[k_h] + rel_w[k_w],  k_h = J / 14, k_w = J % 14
#define WIN_LOGIT(J, VAL) (fmaf(__uint_as_float(VAL), scale_log2e, relh[(J) / WS]) + relw[(J) % WS])

template <int HALF, int CC>
__device__ __forceinline__ void max_chunk(uint32_t trow, const float (&relh)[WS], const float (&relw)[WS],
                                          float scale_log2e, float& mx) {
  constexpr int C = HALF * 3 + CC;
  uint32_t v[32];
  load_s_chunk<C>(trow, v);
#pragma unroll
  for (int i = 0; i < 32; ++i) {
    const int j = C * 32 + i;
    if (j < NTOK) mx = fmaxf(mx, WIN_LOGIT(j < NTOK ? j : 0, v[i]));
  }
}

template <int HALF>
__device__ __forceinline__ float row_max(uint32_t trow, const float (&relh)[WS], const float (&relw)[WS],
                                         float scale_log2e) {
  float mx = -INFINITY;
  max_chunk<HALF, 0>(trow, relh, relw, scale_log2e, mx);
  max_chunk<HALF, 1>(trow, relh, relw, scale_log2e, mx);
  max_chunk<HALF, 2>(trow, relh, relw, scale_log2e, mx);
  if (HALF == 1) max_chunk<1, 3>(trow, relh, relw, scale_log2e, mx);
  return mx;
}

// exp2 of one chunk (relh already has the row max subtracted), P written to shared memory in operand format
template <int HALF, int CC>
__device__ __forceinline__ void exp_chunk(uint32_t trow, const float (&relh)[WS], const float (&relw)[WS],
                                          float scale_log2e, uint8_t* smem, int row, int fmt, float& sum) {
  constexpr int C = HALF * 3 + CC;
  uint32_t v[32];
  load_s_chunk<C>(trow, v);
  float p[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) {
    const int j = C * 32 + i;
    if (j < NTOK) {
      p[i] = ex2(WIN_LOGIT(j < NTOK ? j : 0, v[i]));
      sum += p[i];
    } else {
      p[i] = 0.f;
    }
  }
#pragma unroll
  for (int g = 0; g < 4; ++g) {
    const int j0 = C * 32 + g * 8;
    if (j0 < NKEY) {
      uint4 u;
      u.x = ptx::pack2(p[g * 8 + 0], p[g * 8 + 1], fmt);
      u.y = ptx::pack2(p[g * 8 + 2], p[g * 8 + 3], fmt);
      u.z = ptx::pack2(p[g * 8 + 4], p[g * 8 + 5], fmt);
      u.w = ptx::pack2(p[g * 8 + 6], p[g * 8 + 7], fmt);
      if (j0 < 192)
        *reinterpret_cast<uint4*>(smem + OFF_P + (j0 >> 6) * 16384 + row_off64(row, (j0 & 63) >> 3)) = u;
      else
        *reinterpret_cast<uint4*>(smem + OFF_P16 + row_off16(row, (j0 - 192) >> 3)) = u;
    }
  }
}

template <int HALF>
__device__ __forceinline__ float row_exp(uint32_t trow, const float (&relh)[WS], const float (&relw)[WS],
                                         float scale_log2e, uint8_t* smem, int row, int fmt) {
  float sum = 0.f;
  exp_chunk<HALF, 0>(trow, relh, relw, scale_log2e, smem, row, fmt, sum);
  exp_chunk<HALF, 1>(trow, relh, relw, scale_log2e, smem, row, fmt, sum);
  exp_chunk<HALF, 2>(trow, relh, relw, scale_log2e, smem, row, fmt, sum);
  if (HALF == 1) exp_chunk<1, 3>(trow, relh, relw, scale_log2e, smem, row, fmt, sum);
  return sum;
}

__global__ void __launch_bounds__(kThreads, 2)
win_attn_kernel(const __grid_constant__ WinAttnMaps maps, const uint16_t* __restrict__ bias_op,
                uint16_t* __restrict__ out, const int E, const int heads,
                const int fmt, const float scale_log2e) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bar_load = reinterpret_cast<uint64_t*>(smem + OFF_BAR);
  uint64_t* bar_v = bar_load + 1;
  uint64_t* bar_mma = bar_load + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_load + 3);
  float* xmax = reinterpret_cast<float*>(smem + OFF_XCH);
  float* xsum = xmax + 256;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int row = (warp & 3) * 32 + lane;   // query row of the tile == TMEM lane
  const int half = warp >> 2;               // which half of the key columns this thread handles
  // work decomposition: blockIdx.x = ((b*25 + win)*heads + head)*2 + mtile
  int w = blockIdx.x;
  const int mtile = w & 1;
  w >>= 1;
  const int head = w % heads;
  w /= heads;
  const int win = w % 25;
  const int b = w / 25;
  const int wy = win / 5, wx = win % 5;
  const int iy0 = mtile ? 9 : 0;
  const int nq = mtile ? 70 : 126;
  const uint32_t sbase = ptx::smem_u32(smem);
  const int cq = head * HD, ck = E + head * HD, cv = 2 * E + head * HD;
  const int x0 = wx * WS, y0 = wy * WS;
  const bool padded_window = (wy == 4) || (wx == 4);

  if (tid == 0) {
    ptx::prefetch_tmap(&maps.kv64);
    ptx::prefetch_tmap(&maps.kv16);
    ptx::mbar_init(bar_load, 1);
    ptx::mbar_init(bar_v, 1);
    ptx::mbar_init(bar_mma, 1);
    ptx::fence_mbar_init();
    // Q and K tiles: issued first so their latency overlaps the table load and the TMEM allocation
    ptx::mbar_expect_tx(bar_load, static_cast<uint32_t>((nq + NTOK + 64) * HD * 2));
    ptx::tma_load_2d(smem + OFF_R64, &maps.r64, bar_load, 0, 0);
    ptx::tma_load_2d(smem + OFF_R16, &maps.r16, bar_load, 64, 0);
    const CUtensorMap* q64 = mtile ? &maps.qb64 : &maps.qa64;
    const CUtensorMap* q16 = mtile ? &maps.qb16 : &maps.qa16;
    ptx::tma_load_4d(smem + OFF_Q64, q64, bar_load, cq, x0, y0 + iy0, b);
    ptx::tma_load_4d(smem + OFF_Q16, q16, bar_load, cq + 64, x0, y0 + iy0, b);
    ptx::tma_load_4d(smem + OFF_K64, &maps.kv64, bar_load, ck, x0, y0, b);
    ptx::tma_load_4d(smem + OFF_K16, &maps.kv16, bar_load, ck + 64, x0, y0, b);
  }
  if (warp == 0) {
    ptx::tmem_alloc(tmem_slot, 256);
    ptx::tmem_relinquish();
  }
  // (Rows of Q beyond nq and rows of K beyond 196 stay uninitialised: they only feed accumulator rows / columns that
  // are never read.)
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  ptx::mbar_wait(bar_load, 0);

  // padded tokens (outside the 64x64 grid): q/k := qkv bias (image_encoder.py:281 pads the LN output with zeros)
  if (padded_window) {
    if (tid < NTOK) {
      const int iy = tid / WS, ix = tid % WS;
      if (y0 + iy >= 64 || x0 + ix >= 64) fill_row(smem + OFF_K64, smem + OFF_K16, tid, bias_op + E + head * HD);
    }
    if (tid < nq) {
      const int iy = iy0 + tid / WS, ix = tid % WS;
      if (y0 + iy >= 64 || x0 + ix >= 64) fill_row(smem + OFF_Q64, smem + OFF_Q16, tid, bias_op + head * HD);
    }
  }
  ptx::fence_proxy_async_smem();
  __syncthreads();

  const uint32_t id_T = ptx::make_idesc((uint32_t)fmt, 128, 32, 0, 0);
  const uint32_t id_S = ptx::make_idesc((uint32_t)fmt, 128, NKEY, 0, 0);
  const uint32_t id_O64 = ptx::make_idesc((uint32_t)fmt, 128, 64, 0, 1);
  const uint32_t id_O16 = ptx::make_idesc((uint32_t)fmt, 128, 16, 0, 1);
  constexpr uint32_t kColTw = 196, kColTh = 224;

  // ---- one MMA batch:  S = Q.K^T -> TMEM cols [0,208)   (cols 196..207 belong to pad keys and are dead)
  //                      Tw = Q.Rw^T -> cols [196,228)     (27 used; issued after S, so it may overwrite S's dead cols)
  //                      Th = Q.Rh^T -> cols [224,256)     (27 used)
  if (tid == 0) {
    ptx::tc_fence_after();
    const uint64_t dq64 = ptx::make_smem_desc(sbase + OFF_Q64, 16, 1024, ptx::kSwz128);
    const uint64_t dq16 = ptx::make_smem_desc(sbase + OFF_Q16, 16, 256, ptx::kSwz32);
    const uint64_t dk64 = ptx::make_smem_desc(sbase + OFF_K64, 16, 1024, ptx::kSwz128);
    const uint64_t dk16 = ptx::make_smem_desc(sbase + OFF_K16, 16, 256, ptx::kSwz32);
    const uint64_t dr64 = ptx::make_smem_desc(sbase + OFF_R64, 16, 1024, ptx::kSwz128);
    const uint64_t dr16 = ptx::make_smem_desc(sbase + OFF_R16, 16, 256, ptx::kSwz32);
#pragma unroll
    for (int k = 0; k < 4; ++k) ptx::mma_f16_ss(tmem, dq64 + 2 * k, dk64 + 2 * k, id_S, k != 0);
    ptx::mma_f16_ss(tmem, dq16, dk16, id_S, 1);
    // table rows 32..63 = rel_pos_w: +32 rows = +4096 B (64-wide part) / +1024 B (16-wide part)
#pragma unroll
    for (int k = 0; k < 4; ++k) ptx::mma_f16_ss(tmem + kColTw, dq64 + 2 * k, dr64 + (4096 >> 4) + 2 * k, id_T, k != 0);
    ptx::mma_f16_ss(tmem + kColTw, dq16, dr16 + (1024 >> 4), id_T, 1);
#pragma unroll
    for (int k = 0; k < 4; ++k) ptx::mma_f16_ss(tmem + kColTh, dq64 + 2 * k, dr64 + 2 * k, id_T, k != 0);
    ptx::mma_f16_ss(tmem + kColTh, dq16, dr16, id_T, 1);
    ptx::mma_commit(bar_mma);
  }
  ptx::mbar_wait(bar_mma, 0);
  ptx::tc_fence_after();

  const uint32_t trow = tmem + (static_cast<uint32_t>((warp & 3) * 32) << 16);
  const int qiy = iy0 + row / WS;  // query position inside the window (garbage rows >= nq are never stored)
  const int qix = row % WS;
  float relh[WS], relw[WS];
  {
    // half 0 spills the rel_pos_h products of its row to scratch, half 1 the rel_pos_w products; after the barrier
    // both threads of the row gather their 14 + 14 bias terms (index = q - k + 13, image_encoder.py:347-351)
    const float kLog2e = 1.4426950408889634f;
    float* th = reinterpret_cast<float*>(smem + OFF_TH) + row * 27;
    float* tw = reinterpret_cast<float*>(smem + OFF_TW) + row * 27;
    uint32_t v[32];
    ptx::tmem_ld_32x32b_x32(trow + (half ? kColTw : kColTh), v);
    ptx::tmem_ld_wait();
    float* dst = half ? tw : th;
#pragma unroll
    for (int j = 0; j < 27; ++j) dst[j] = __uint_as_float(v[j]) * kLog2e;
    __syncthreads();
    const int qh = (qiy < WS) ? qiy : (WS - 1);
#pragma unroll
    for (int kh = 0; kh < WS; ++kh) relh[kh] = th[qh - kh + (WS - 1)];
#pragma unroll
    for (int kw = 0; kw < WS; ++kw) relw[kw] = tw[qix - kw + (WS - 1)];
  }
  ptx::fence_proxy_async_smem();   // scratch (generic proxy) is about to be overwritten by the V tile (async proxy)
  __syncthreads();
  if (tid == 0) {
    ptx::mbar_expect_tx(bar_v, static_cast<uint32_t>(NTOK * HD * 2));
    ptx::tma_load_4d(smem + OFF_V64, &maps.kv64, bar_v, cv, x0, y0, b);
    ptx::tma_load_4d(smem + OFF_V16, &maps.kv16, bar_v, cv + 64, x0, y0, b);
  }

  // ---- softmax over the 196 keys (padded keys included, exactly as the reference); half 0 owns key columns
  //      [0,96), half 1 [96,208); partial max / sum are exchanged through shared memory
  float mx = half ? row_max<1>(trow, relh, relw, scale_log2e) : row_max<0>(trow, relh, relw, scale_log2e);
  xmax[half * 128 + row] = mx;
  __syncthreads();
  mx = fmaxf(xmax[row], xmax[128 + row]);
#pragma unroll
  for (int kh = 0; kh < WS; ++kh) relh[kh] -= mx;
  const float sum = half ? row_exp<1>(trow, relh, relw, scale_log2e, smem, row, fmt)
                         : row_exp<0>(trow, relh, relw, scale_log2e, smem, row, fmt);
  xsum[half * 128 + row] = sum;
  // V tile: zero the 12 pad rows (P is 0 there, but 0 x garbage could be NaN) and patch padded tokens with the bias
  ptx::mbar_wait(bar_v, 0);
  if (tid < (NKEY - NTOK) * 10) {
    const int r = NTOK + tid / 10, c = tid % 10;
    const uint4 z = make_uint4(0, 0, 0, 0);
    if (c < 8)
      *reinterpret_cast<uint4*>(smem + OFF_V64 + row_off64(r, c)) = z;
    else
      *reinterpret_cast<uint4*>(smem + OFF_V16 + row_off16(r, c - 8)) = z;
  }
  if (padded_window && tid < NTOK) {
    const int iy = tid / WS, ix = tid % WS;
    if (y0 + iy >= 64 || x0 + ix >= 64) fill_row(smem + OFF_V64, smem + OFF_V16, tid, bias_op + 2 * E + head * HD);
  }
  ptx::fence_proxy_async_smem();
  ptx::tc_fence_before();
  __syncthreads();

  // ---- O = P . V -> TMEM cols [0,80)  (V is the MN-major B operand: rows = keys, d contiguous)
  if (tid == 0) {
    ptx::tc_fence_after();
    const uint64_t dp64 = ptx::make_smem_desc(sbase + OFF_P, 16, 1024, ptx::kSwz128);
    const uint64_t dp16 = ptx::make_smem_desc(sbase + OFF_P16, 16, 256, ptx::kSwz32);
    const uint64_t dv64 = ptx::make_smem_desc(sbase + OFF_V64, NKEY * 128, 1024, ptx::kSwz128);
    const uint64_t dv16 = ptx::make_smem_desc(sbase + OFF_V16, NKEY * 32, 256, ptx::kSwz32);
#pragma unroll
    for (int ks = 0; ks < NKEY / 16; ++ks) {
      // descriptor start-address field counts 16-byte units: all offsets below are compile-time constants
      const uint64_t da = (ks < 12) ? dp64 + (((ks >> 2) * 16384 + (ks & 3) * 32) >> 4) : dp16;
      ptx::mma_f16_ss(tmem, da, dv64 + ((ks * 2048) >> 4), id_O64, ks != 0);
      ptx::mma_f16_ss(tmem + 64, da, dv16 + ((ks * 512) >> 4), id_O16, ks != 0);
    }
    ptx::mma_commit(bar_mma);
  }
  ptx::mbar_wait(bar_mma, 1);
  ptx::tc_fence_after();

  {
    const float inv = 1.0f / (xsum[row] + xsum[128 + row]);
    const int y = wy * WS + qiy, x = wx * WS + qix;
    const bool ok = (row < nq) && (y < 64) && (x < 64);
    uint16_t* dst = out + (static_cast<size_t>(b) * 4096 + (ok ? (y * 64 + x) : 0)) * E + head * HD;
    // half 0 stores head-dim columns [0,48), half 1 [48,80)
    const int c0 = half ? 3 : 0, c1 = half ? 5 : 3;
#pragma unroll
    for (int cc = 0; cc < 3; ++cc) {
      const int c = c0 + cc;
      if (c < c1) {
        uint32_t v[16];
        ptx::tmem_ld_32x32b_x16(trow + c * 16, v);
        ptx::tmem_ld_wait();
        if (ok) {
          uint4 u0, u1;
          u0.x = ptx::pack2(__uint_as_float(v[0]) * inv, __uint_as_float(v[1]) * inv, fmt);
          u0.y = ptx::pack2(__uint_as_float(v[2]) * inv, __uint_as_float(v[3]) * inv, fmt);
          u0.z = ptx::pack2(__uint_as_float(v[4]) * inv, __uint_as_float(v[5]) * inv, fmt);
          u0.w = ptx::pack2(__uint_as_float(v[6]) * inv, __uint_as_float(v[7]) * inv, fmt);
          u1.x = ptx::pack2(__uint_as_float(v[8]) * inv, __uint_as_float(v[9]) * inv, fmt);
          u1.y = ptx::pack2(__uint_as_float(v[10]) * inv, __uint_as_float(v[11]) * inv, fmt);
          u1.z = ptx::pack2(__uint_as_float(v[12]) * inv, __uint_as_float(v[13]) * inv, fmt);
          u1.w = ptx::pack2(__uint_as_float(v[14]) * inv, __uint_as_float(v[15]) * inv, fmt);
          reinterpret_cast<uint4*>(dst + c * 16)[0] = u0;
          reinterpret_cast<uint4*>(dst + c * 16)[1] = u1;
        }
      }
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem, 256);
  }
}

}  // namespace

int samk_attn_window(const void* qkv, const void* bias_op, const void* rel_tab, void* out, int B, int E, int heads,
                     int fmt, cudaStream_t stream) {
  SAM_REQUIRE(fmt == 0 || fmt == 1, "attn_window: fmt must be fp16/bf16");
  SAM_REQUIRE(B > 0, "attn_window: empty batch");
  // default: the tensor-memory-P kernel (attn_window4.cu); SAM_ATTN_WINDOW_V3=1 selects the shared-memory-P persistent
  // kernel (attn_window3.cu), SAM_ATTN_WINDOW_V2=1 this file's one-CTA-per-tile kernel (the simplest implementation)
  static const bool use_v2 = getenv("SAM_ATTN_WINDOW_V2") != nullptr;
  static const bool use_v3 = getenv("SAM_ATTN_WINDOW_V3") != nullptr;
  if (!use_v2 && !use_v3) return samk_attn_window4(qkv, bias_op, rel_tab, out, B, E, heads, fmt, stream);
  SAM_REQUIRE(E == heads * HD, "attn_window: the v2 / v3 kernels need head_dim 80 (E=%d heads=%d)", E, heads);
  if (use_v3) return samk_attn_window3(qkv, bias_op, rel_tab, out, B, E, heads, fmt, stream);
  WinAttnMaps maps;
  const int is_bf16 = (fmt == 1);
  const uint64_t ld = static_cast<uint64_t>(3) * E * 2;  // bytes per token row
  const uint64_t dims[4] = {static_cast<uint64_t>(3 * E), 64, 64, static_cast<uint64_t>(B)};
  const uint64_t strides[4] = {2, ld, 64 * ld, 4096 * ld};
  struct { CUtensorMap* m; uint32_t c, rows; int swz; } specs[6] = {
      {&maps.kv64, 64, 14, 3}, {&maps.kv16, 16, 14, 1}, {&maps.qa64, 64, 9, 3},
      {&maps.qa16, 16, 9, 1},  {&maps.qb64, 64, 5, 3},  {&maps.qb16, 16, 5, 1}};
  for (auto& s : specs) {
    const uint32_t box[4] = {s.c, 14, s.rows, 1};
    int rc = samhost::encode_tmap_nd(s.m, 2, is_bf16, qkv, 4, dims, strides, box, s.swz);
    if (rc) return rc;
  }
  {
    int rc = samhost::encode_tmap_2d(&maps.r64, 2, is_bf16, rel_tab, HD, 64, HD * 2, 64, 64, 3);
    if (rc) return rc;
    rc = samhost::encode_tmap_2d(&maps.r16, 2, is_bf16, rel_tab, HD, 64, HD * 2, 16, 64, 1);
    if (rc) return rc;
  }
  static bool attr_done = false;
  if (!attr_done) {
    SAM_CHECK_CUDA(cudaFuncSetAttribute(win_attn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
    attr_done = true;
  }
  const float scale_log2e = 1.4426950408889634f / sqrtf(static_cast<float>(HD));
  const int grid = B * 25 * heads * 2;
  // algorithmic work per (window, head): QK^T + PV over 196 x 196 x 80 and the two rel-pos products (196 x 14 x 80 x 2)
  const double wh = static_cast<double>(B) * 25 * heads;
  samhost::LaunchScope scope(samhost::KC_ATTN_WINDOW, stream, wh * (4.0 * 196 * 196 * 80 + 4.0 * 196 * 14 * 80),
                             static_cast<double>(B) * 4096 * E * 2 * 4);
  win_attn_kernel<<<grid, kThreads, kSmemBytes, stream>>>(maps, static_cast<const uint16_t*>(bias_op),
                                                                                                                      static_cast<uint16_t*>(out), E, heads, fmt, scale_log2e);
  SAM_CHECK_CUDA(cudaGetLastError());
  return 0;
}
