// 14x14 windowed attention of the SAM ViT-H encoder with the decomposed relative-position bias fused into the
// softmax -- replaces image_encoder.py:235-257 (Attention.forward minus qkv/proj), :263-318 (window partition /
// unpartition, done here by index arithmetic + TMA boxes) and :354-392 (add_decomposed_rel_pos).
//
// Input  qkv  [B*64*64, 3*E] operand format (fp16/bf16), token-major, UN-partitioned; columns (which, head, d).
// Output out  [B*64*64, E]   operand format, token-major, heads merged -- the proj GEMM's A operand.
//
// One CTA (128 threads, 2 CTAs/SM) = one (image, window, head, query tile); query tile 0 = window rows 0..8
// (126 tokens), tile 1 = rows 9..13 (70 tokens).  All matrix products run on tcgen05 with TMEM accumulators:
//   T = Q.R^T   (N=64 : 27 rel_pos_h rows | 27 rel_pos_w rows)  -> per-row bias look-up tables
//   S = Q.K^T   (N=208: 196 keys padded to a multiple of 16)
//   O = P.V     (N=64 + N=16, V consumed MN-major straight from the TMA tile)
// Q/K/V tiles arrive by 4-D TMA boxes over the [B,64,64,3E] view (128B- and 32B-swizzled for the 64+16 split of
// head_dim 80).  Window padding (image_encoder.py:277-283 pads AFTER norm1, so padded tokens have q/k/v == qkv bias
// and DO take part in the softmax) is reproduced exactly: TMA zero-fills the out-of-image rows and the kernel
// overwrites them with the bias; padded queries are computed but never stored.
// Softmax: one thread per query row (TMEM lane), fp32, exp2 with the scale folded in; P is rounded to the operand
// format, the row sum is applied to O in fp32.
#include "host_common.h"
#include "kernels.h"
#include "ptx.cuh"

namespace {

constexpr int HD = 80;
constexpr int WS = 14;
constexpr int NTOK = WS * WS;  // 196
constexpr int NKEY = 208;      // keys padded to the UMMA N/K granularity
constexpr int kThreads = 128;

// shared-memory map (bytes from the 1024-aligned base)
constexpr int OFF_Q64 = 0;       // 128 x 128B  SWIZZLE_128B
constexpr int OFF_K64 = 16384;   // 208 x 128B
constexpr int OFF_R64 = 43008;   //  64 x 128B
constexpr int OFF_Q16 = 51200;   // 128 x 32B   SWIZZLE_32B
constexpr int OFF_K16 = 55296;   // 208 x 32B
constexpr int OFF_R16 = 61952;   //  64 x 32B
constexpr int OFF_V64 = 64512;   // 208 x 128B  (MN-major operand of P.V)
constexpr int OFF_V16 = 91136;   // 208 x 32B
constexpr int OFF_TL = 97792;    // 128 x 27 fp32 bias look-up scratch
constexpr int OFF_BAR = 111616;
constexpr int kSmemBytes = OFF_BAR + 64 + 1024;
// P (probabilities, operand format) overlays Q/K/R once S and T have been consumed
constexpr int OFF_P = 0;         // 3 x (128 x 128B) SWIZZLE_128B + 128 x 32B SWIZZLE_32B at +49152
constexpr int OFF_P16 = 49152;

struct WinAttnMaps {
  CUtensorMap kv64, kv16;    // box {64|16, 14, 14, 1}
  CUtensorMap qa64, qa16;    // box {64|16, 14, 9, 1}   query tile 0
  CUtensorMap qb64, qb16;    // box {64|16, 14, 5, 1}   query tile 1
};

__device__ __forceinline__ float ex2(float x) {
  float y;
  asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

__device__ __forceinline__ uint32_t row_off64(int r, int c) { return r * 128 + ((c ^ (r & 7)) << 4); }
__device__ __forceinline__ uint32_t row_off16(int r, int c) { return r * 32 + ((c ^ ((r >> 2) & 1)) << 4); }

__global__ void __launch_bounds__(kThreads, 2)
win_attn_kernel(const __grid_constant__ WinAttnMaps maps, const uint16_t* __restrict__ bias_op,
                const uint16_t* __restrict__ rel_tab, uint16_t* __restrict__ out, const int E, const int heads,
                const int fmt, const float scale_log2e) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bar_load = reinterpret_cast<uint64_t*>(smem + OFF_BAR);
  uint64_t* bar_mma = bar_load + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_load + 2);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
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

  if (tid == 0) {
    ptx::prefetch_tmap(&maps.kv64);
    ptx::prefetch_tmap(&maps.kv16);
    ptx::mbar_init(bar_load, 1);
    ptx::mbar_init(bar_mma, 1);
    ptx::fence_mbar_init();
  }
  if (warp == 0) {
    ptx::tmem_alloc(tmem_slot, 256);
    ptx::tmem_relinquish();
  }
  // zero the rows TMA never writes (they are read by the MMAs as padding and must be finite)
  {
    const uint4 z = make_uint4(0, 0, 0, 0);
    for (int i = tid; i < (128 - nq) * 8; i += kThreads) *reinterpret_cast<uint4*>(smem + OFF_Q64 + nq * 128 + i * 16) = z;
    for (int i = tid; i < (128 - nq) * 2; i += kThreads) *reinterpret_cast<uint4*>(smem + OFF_Q16 + nq * 32 + i * 16) = z;
    for (int i = tid; i < (NKEY - NTOK) * 8; i += kThreads) {
      *reinterpret_cast<uint4*>(smem + OFF_K64 + NTOK * 128 + i * 16) = z;
      *reinterpret_cast<uint4*>(smem + OFF_V64 + NTOK * 128 + i * 16) = z;
    }
    for (int i = tid; i < (NKEY - NTOK) * 2; i += kThreads) {
      *reinterpret_cast<uint4*>(smem + OFF_K16 + NTOK * 32 + i * 16) = z;
      *reinterpret_cast<uint4*>(smem + OFF_V16 + NTOK * 32 + i * 16) = z;
    }
    // rel-pos table R [64 rows x 80] -> K-major 64+16 split
    for (int i = tid; i < 64 * 10; i += kThreads) {
      const int r = i / 10, c = i % 10;
      const uint4 v = __ldg(reinterpret_cast<const uint4*>(rel_tab + r * HD) + c);
      if (c < 8)
        *reinterpret_cast<uint4*>(smem + OFF_R64 + row_off64(r, c)) = v;
      else
        *reinterpret_cast<uint4*>(smem + OFF_R16 + row_off16(r, c - 8)) = v;
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (tid == 0) {
    const int cq = head * HD, ck = E + head * HD, cv = 2 * E + head * HD;
    const int x0 = wx * WS, y0 = wy * WS;
    ptx::mbar_expect_tx(bar_load, static_cast<uint32_t>((nq + 2 * NTOK) * HD * 2));
    const CUtensorMap* q64 = mtile ? &maps.qb64 : &maps.qa64;
    const CUtensorMap* q16 = mtile ? &maps.qb16 : &maps.qa16;
    ptx::tma_load_4d(smem + OFF_Q64, q64, bar_load, cq, x0, y0 + iy0, b);
    ptx::tma_load_4d(smem + OFF_Q16, q16, bar_load, cq + 64, x0, y0 + iy0, b);
    ptx::tma_load_4d(smem + OFF_K64, &maps.kv64, bar_load, ck, x0, y0, b);
    ptx::tma_load_4d(smem + OFF_K16, &maps.kv16, bar_load, ck + 64, x0, y0, b);
    ptx::tma_load_4d(smem + OFF_V64, &maps.kv64, bar_load, cv, x0, y0, b);
    ptx::tma_load_4d(smem + OFF_V16, &maps.kv16, bar_load, cv + 64, x0, y0, b);
  }
  ptx::mbar_wait(bar_load, 0);

  // padded tokens (outside the 64x64 grid): q/k/v := qkv bias (image_encoder.py:281 pads the LN output with zeros)
  if (wy == 4 || wx == 4) {
    const int total = (nq + 2 * NTOK) * 10;
    for (int i = tid; i < total; i += kThreads) {
      int r = i / 10;
      const int c = i % 10;
      int which, off64, off16, iy;
      if (r < nq) {
        which = 0; off64 = OFF_Q64; off16 = OFF_Q16; iy = iy0 + r / WS;
      } else if (r < nq + NTOK) {
        r -= nq; which = 1; off64 = OFF_K64; off16 = OFF_K16; iy = r / WS;
      } else {
        r -= nq + NTOK; which = 2; off64 = OFF_V64; off16 = OFF_V16; iy = r / WS;
      }
      const int ix = r % WS;
      if (wy * WS + iy >= 64 || wx * WS + ix >= 64) {
        const uint4 v = __ldg(reinterpret_cast<const uint4*>(bias_op + which * E + head * HD) + c);
        if (c < 8)
          *reinterpret_cast<uint4*>(smem + off64 + row_off64(r, c)) = v;
        else
          *reinterpret_cast<uint4*>(smem + off16 + row_off16(r, c - 8)) = v;
      }
    }
  }
  ptx::fence_proxy_async_smem();
  __syncthreads();

  const uint32_t id_T = ptx::make_idesc((uint32_t)fmt, 128, 64, 0, 0);
  const uint32_t id_S = ptx::make_idesc((uint32_t)fmt, 128, NKEY, 0, 0);
  const uint32_t id_O64 = ptx::make_idesc((uint32_t)fmt, 128, 64, 0, 1);
  const uint32_t id_O16 = ptx::make_idesc((uint32_t)fmt, 128, 16, 0, 1);

  // ---- T = Q . R^T  -> TMEM cols [0,64)
  if (tid == 0) {
    ptx::tc_fence_after();
#pragma unroll
    for (int k = 0; k < 4; ++k)
      ptx::mma_f16_ss(tmem, ptx::make_smem_desc(sbase + OFF_Q64 + k * 32, 16, 1024, ptx::kSwz128),
                      ptx::make_smem_desc(sbase + OFF_R64 + k * 32, 16, 1024, ptx::kSwz128), id_T, k != 0);
    ptx::mma_f16_ss(tmem, ptx::make_smem_desc(sbase + OFF_Q16, 16, 256, ptx::kSwz32),
                    ptx::make_smem_desc(sbase + OFF_R16, 16, 256, ptx::kSwz32), id_T, 1);
    ptx::mma_commit(bar_mma);
  }
  ptx::mbar_wait(bar_mma, 0);
  ptx::tc_fence_after();

  const uint32_t trow = tmem + (static_cast<uint32_t>(warp * 32) << 16);
  const int qiy = iy0 + tid / WS;  // query position inside the window (garbage rows >= nq are never stored)
  const int qix = tid % WS;
  float relh[WS], relw[WS];
  {
    const float kLog2e = 1.4426950408889634f;
    float* tl = reinterpret_cast<float*>(smem + OFF_TL) + tid * 27;
    uint32_t v[32];
    ptx::tmem_ld_32x32b_x32(trow, v);
    ptx::tmem_ld_wait();
#pragma unroll
    for (int j = 0; j < 27; ++j) tl[j] = __uint_as_float(v[j]);
    const int qh = (qiy < WS) ? qiy : (WS - 1);
#pragma unroll
    for (int kh = 0; kh < WS; ++kh) relh[kh] = tl[qh - kh + (WS - 1)] * kLog2e;
    uint32_t u[32];
    ptx::tmem_ld_32x32b_x32(trow + 32, u);
    ptx::tmem_ld_wait();
    // columns 27..53 hold q . rel_pos_w[0..26]
#pragma unroll
    for (int j = 27; j < 32; ++j) tl[j - 27] = __uint_as_float(v[j]);
#pragma unroll
    for (int j = 32; j < 54; ++j) tl[j - 27] = __uint_as_float(u[j - 32]);
#pragma unroll
    for (int kw = 0; kw < WS; ++kw) relw[kw] = tl[qix - kw + (WS - 1)] * kLog2e;
  }
  ptx::tc_fence_before();
  __syncthreads();

  // ---- S = Q . K^T -> TMEM cols [0,208)
  if (tid == 0) {
    ptx::tc_fence_after();
#pragma unroll
    for (int k = 0; k < 4; ++k)
      ptx::mma_f16_ss(tmem, ptx::make_smem_desc(sbase + OFF_Q64 + k * 32, 16, 1024, ptx::kSwz128),
                      ptx::make_smem_desc(sbase + OFF_K64 + k * 32, 16, 1024, ptx::kSwz128), id_S, k != 0);
    ptx::mma_f16_ss(tmem, ptx::make_smem_desc(sbase + OFF_Q16, 16, 256, ptx::kSwz32),
                    ptx::make_smem_desc(sbase + OFF_K16, 16, 256, ptx::kSwz32), id_S, 1);
    ptx::mma_commit(bar_mma);
  }
  ptx::mbar_wait(bar_mma, 1);
  ptx::tc_fence_after();

  // ---- softmax over the 196 keys (padded keys included, exactly as the reference)
  float mx = -INFINITY;
#pragma unroll
  for (int c = 0; c < 7; ++c) {
    uint32_t v[32];
    if (c < 6) {
      ptx::tmem_ld_32x32b_x32(trow + c * 32, v);
    } else {
      uint32_t t16[16];
      ptx::tmem_ld_32x32b_x16(trow + 192, t16);
#pragma unroll
      for (int i = 0; i < 16; ++i) v[i] = t16[i];
    }
    ptx::tmem_ld_wait();
#pragma unroll
    for (int i = 0; i < 32; ++i) {
      const int j = c * 32 + i;
      if (j < NTOK) mx = fmaxf(mx, fmaf(__uint_as_float(v[i]), scale_log2e, relh[j / WS]) + relw[j % WS]);
    }
  }
#pragma unroll
  for (int kh = 0; kh < WS; ++kh) relh[kh] -= mx;
  float sum = 0.f;
#pragma unroll
  for (int c = 0; c < 7; ++c) {
    uint32_t v[32];
    if (c < 6) {
      ptx::tmem_ld_32x32b_x32(trow + c * 32, v);
    } else {
      uint32_t t16[16];
      ptx::tmem_ld_32x32b_x16(trow + 192, t16);
#pragma unroll
      for (int i = 0; i < 16; ++i) v[i] = t16[i];
    }
    ptx::tmem_ld_wait();
    float p[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) {
      const int j = c * 32 + i;
      if (j < NTOK) {
        p[i] = ex2(fmaf(__uint_as_float(v[i]), scale_log2e, relh[j / WS]) + relw[j % WS]);
        sum += p[i];
      } else {
        p[i] = 0.f;
      }
    }
#pragma unroll
    for (int g = 0; g < 4; ++g) {
      const int j0 = c * 32 + g * 8;
      if (j0 < NKEY) {
        uint4 u;
        u.x = ptx::pack2(p[g * 8 + 0], p[g * 8 + 1], fmt);
        u.y = ptx::pack2(p[g * 8 + 2], p[g * 8 + 3], fmt);
        u.z = ptx::pack2(p[g * 8 + 4], p[g * 8 + 5], fmt);
        u.w = ptx::pack2(p[g * 8 + 6], p[g * 8 + 7], fmt);
        if (j0 < 192)
          *reinterpret_cast<uint4*>(smem + OFF_P + (j0 >> 6) * 16384 + row_off64(tid, (j0 & 63) >> 3)) = u;
        else
          *reinterpret_cast<uint4*>(smem + OFF_P16 + row_off16(tid, (j0 - 192) >> 3)) = u;
      }
    }
  }
  ptx::fence_proxy_async_smem();
  ptx::tc_fence_before();
  __syncthreads();

  // ---- O = P . V -> TMEM cols [0,80)  (V is the MN-major B operand: rows = keys, d contiguous)
  if (tid == 0) {
    ptx::tc_fence_after();
#pragma unroll 1
    for (int ks = 0; ks < NKEY / 16; ++ks) {
      uint64_t da;
      if (ks < 12)
        da = ptx::make_smem_desc(sbase + OFF_P + (ks >> 2) * 16384 + (ks & 3) * 32, 16, 1024, ptx::kSwz128);
      else
        da = ptx::make_smem_desc(sbase + OFF_P16, 16, 256, ptx::kSwz32);
      ptx::mma_f16_ss(tmem, da, ptx::make_smem_desc(sbase + OFF_V64 + ks * 2048, NKEY * 128, 1024, ptx::kSwz128), id_O64,
                      ks != 0);
      ptx::mma_f16_ss(tmem + 64, da, ptx::make_smem_desc(sbase + OFF_V16 + ks * 512, NKEY * 32, 256, ptx::kSwz32), id_O16,
                      ks != 0);
    }
    ptx::mma_commit(bar_mma);
  }
  ptx::mbar_wait(bar_mma, 0);
  ptx::tc_fence_after();

  {
    const float inv = 1.0f / sum;
    const int y = wy * WS + qiy, x = wx * WS + qix;
    const bool ok = (tid < nq) && (y < 64) && (x < 64);
    uint16_t* dst = out + (static_cast<size_t>(b) * 4096 + (ok ? (y * 64 + x) : 0)) * E + head * HD;
#pragma unroll
    for (int c = 0; c < 5; ++c) {
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
  SAM_REQUIRE(E == heads * HD, "attn_window: head_dim must be 80 (E=%d heads=%d)", E, heads);
  SAM_REQUIRE(B > 0, "attn_window: empty batch");
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
                                                           static_cast<const uint16_t*>(rel_tab),
                                                           static_cast<uint16_t*>(out), E, heads, fmt, scale_log2e);
  SAM_CHECK_CUDA(cudaGetLastError());
  return 0;
}
