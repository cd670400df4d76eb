// 14x14 windowed attention of the SAM ViT-H encoder, persistent warp-specialised version ("v3").
// Same contract as attn_window.cu (replaces image_encoder.py:235-257, :263-318, :354-392); see that file for the maths.
//
// One CTA per SM loops over work items (image, window, head).  Both query tiles of an item (window rows 0..8 = 126
// tokens, rows 9..13 = 70 tokens) are processed concurrently by two softmax warpgroups that share the K and V tiles:
//
//   warp 0, 11     : producers (Q0/Q1/K and V): stream item i+1 while item i is in its softmax; patch the zero-padded
//                    window tokens with the qkv bias (image_encoder.py:281) off the critical path
//   warps 1, 10    : MMA issuers (one elected thread each) for tile 0 / tile 1:  S = Q_g.K^T (N=208), Tw = Q_g.Rw^T,
//                    Th = Q_g.Rh^T -> TMEM slot g, later O = P_g.V (N = 64 + 16) into the same slot
//   warps 2..5     : softmax of tile 0 (one thread per query row / TMEM lane)
//   warps 6..9     : softmax of tile 1
//
// All hand-offs are mbarriers; per-CTA set-up (TMEM allocation, rel-pos table, barrier init) happens once.
// TMEM: 2 slots x 256 columns: S [0,208) (cols 196..207 are pad keys), Tw [196,228), Th [224,256), O [0,80).
#include <stdio.h>
#include <stdlib.h>

#include "host_common.h"
#include "kernels.h"
#include "ptx.cuh"

namespace {

constexpr int HD = 80;
constexpr int WS = 14;
constexpr int NTOK = WS * WS;  // 196
constexpr int NKEY = 208;
constexpr int kThreads3 = 384;

// shared-memory map (bytes from the 1024-aligned base)
constexpr int OFF_Q64 = 0;            // 2 x (128 x 128B) SWIZZLE_128B   (tile g at + g*16384)
constexpr int OFF_K64 = 32768;        // 208 x 128B
constexpr int OFF_R64 = 59392;        //  64 x 128B
constexpr int OFF_Q16 = 67584;        // 2 x (128 x 32B) SWIZZLE_32B     (tile g at + g*4096)
constexpr int OFF_K16 = 75776;        // 208 x 32B
constexpr int OFF_R16 = 82432;        //  64 x 32B
constexpr int OFF_V64 = 84992;        // 208 x 128B (MN-major operand of P.V)
constexpr int OFF_V16 = 111616;       // 208 x 32B
constexpr int OFF_P = 118784;         // 2 x 53248: per tile 3 x (128 x 128B) + 128 x 32B; doubles as rel-pos scratch
constexpr int kPBytes = 53248;
constexpr int OFF_BAR = OFF_P + 2 * kPBytes;   // 225280
constexpr int kSmemBytes3 = OFF_BAR + 256 + 1024;
constexpr int kScratchStride = 55;    // floats per row of the rel-pos scratch (odd: conflict-free)

struct WinAttnMaps3 {
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
__device__ __forceinline__ void fill_row_regs(uint8_t* t64, uint8_t* t16, int r, const uint4 (&b)[10]) {
#pragma unroll
  for (int c = 0; c < 10; ++c) {
    if (c < 8)
      *reinterpret_cast<uint4*>(t64 + row_off64(r, c)) = b[c];
    else
      *reinterpret_cast<uint4*>(t16 + row_off16(r, c - 8)) = b[c];
  }
}
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

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

#define WIN3_LOGIT(J, VAL) (fmaf(__uint_as_float(VAL), scale_log2e, relh[(J) / WS]) + relw[(J) % WS])

template <int C>
__device__ __forceinline__ void max_chunk(uint32_t trow, const float (&relh)[WS], const float (&relw)[WS],
                                          float scale_log2e, float& mx) {
  uint32_t v[32];
  load_s_chunk<C>(trow, v);
#pragma unroll
  for (int i = 0; i < 32; ++i) {
    const int j = C * 32 + i;
    if (j < NTOK) mx = fmaxf(mx, WIN3_LOGIT(j < NTOK ? j : 0, v[i]));
  }
}

template <int C, int FMT>
__device__ __forceinline__ void exp_chunk(uint32_t trow, const float (&relh)[WS], const float (&relw)[WS],
                                          float scale_log2e, uint32_t pbase, int row, float& sum) {
  uint32_t v[32];
  load_s_chunk<C>(trow, v);
  float p[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) {
    const int j = C * 32 + i;
    if (j < NTOK) {
      p[i] = ex2(WIN3_LOGIT(j < NTOK ? j : 0, v[i]));
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
      u.x = ptx::pack2t<FMT>(p[g * 8 + 0], p[g * 8 + 1]);
      u.y = ptx::pack2t<FMT>(p[g * 8 + 2], p[g * 8 + 3]);
      u.z = ptx::pack2t<FMT>(p[g * 8 + 4], p[g * 8 + 5]);
      u.w = ptx::pack2t<FMT>(p[g * 8 + 6], p[g * 8 + 7]);
      if (j0 < 192)
        ptx::st_shared_v4(pbase + (j0 >> 6) * 16384 + row_off64(row, (j0 & 63) >> 3), u);
      else
        ptx::st_shared_v4(pbase + 49152 + row_off16(row, (j0 - 192) >> 3), u);
    }
  }
}

// Single-pass variant working on an already loaded chunk: logits relative to the reference maximum (folded into relh),
// two at a time on the packed fp32 pipe; four partial row sums (two packed accumulators) keep the add chain short.
using ptx::add2;
using ptx::f32x2;
using ptx::fma2;
using ptx::pk2;
using ptx::upk2;
template <int C, int FMT>
__device__ __forceinline__ void exp_chunk_regs(const uint32_t (&v)[32], const float (&relh)[WS], const float (&relw)[WS],
                                               f32x2 sc2, uint32_t pbase, int row, f32x2& s0, f32x2& s1) {
  constexpr int kN = (C < 6) ? 32 : 16;
  uint32_t pk[kN / 2];
#pragma unroll
  for (int i = 0; i < kN; i += 2) {
    const int j = C * 32 + i;   // even; NTOK and WS are even, so j and j + 1 share their key row and validity
    if (j < NTOK) {
      const int jj = j < NTOK ? j : 0;
      f32x2 x = fma2(pk2(__uint_as_float(v[i]), __uint_as_float(v[i + 1])), sc2, pk2(relh[jj / WS], relh[jj / WS]));
      x = add2(x, pk2(relw[jj % WS], relw[jj % WS + 1]));
      float x0, x1;
      upk2(x, x0, x1);
      const float p0 = ex2(x0), p1 = ex2(x1);
      if ((i >> 1) & 1)
        s1 = add2(s1, pk2(p0, p1));
      else
        s0 = add2(s0, pk2(p0, p1));
      pk[i >> 1] = ptx::pack2t<FMT>(p0, p1);
    } else {
      pk[i >> 1] = 0u;
    }
  }
#pragma unroll
  for (int g = 0; g < kN / 8; ++g) {
    const int j0 = C * 32 + g * 8;
    const uint4 u = make_uint4(pk[g * 4], pk[g * 4 + 1], pk[g * 4 + 2], pk[g * 4 + 3]);
    if (j0 < 192)
      ptx::st_shared_v4(pbase + (j0 >> 6) * 16384 + row_off64(row, (j0 & 63) >> 3), u);
    else
      ptx::st_shared_v4(pbase + 49152 + row_off16(row, (j0 - 192) >> 3), u);
  }
}

struct Item {
  int b, wy, wx, head;
};
__device__ __forceinline__ Item decode_item(int it, int heads) {
  Item r;
  r.head = it % heads;
  it /= heads;
  const int win = it % 25;
  r.b = it / 25;
  r.wy = win / 5;
  r.wx = win % 5;
  return r;
}

// Debug timeline (SAM_WIN3_TRACE=1): block 0 records clock64() of the hand-offs of its first 8 items.
#define WIN3_TR(role, ev)                                                                   \
  do {                                                                                      \
    if (trace != nullptr && blockIdx.x == 0 && n < 8) trace[((role) * 8 + (ev)) * 8 + n] = clock64(); \
  } while (0)

template <int FMT>
__global__ void __launch_bounds__(kThreads3, 1)
win_attn3_kernel(const __grid_constant__ WinAttnMaps3 maps, const uint16_t* __restrict__ bias_op,
                 uint16_t* __restrict__ out, const int E, const int heads, const int num_items,
                 const float scale_log2e, long long* __restrict__ trace, const int skew) {
  constexpr int fmt = FMT;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + OFF_BAR);
  uint64_t* qk_full = bars + 0;    // TMA bytes of Q0/Q1/K (+ R on the first item)
  uint64_t* qk_ready = bars + 1;   // ... and padded tokens patched            (TMA warp -> MMA)
  uint64_t* qk_free = bars + 2;    // S/T MMAs of both tiles done              (MMA -> TMA)
  uint64_t* v_full = bars + 3;
  uint64_t* v_ready = bars + 4;
  uint64_t* v_free = bars + 5;     // PV MMAs of both tiles done               (MMA -> TMA)
  uint64_t* s_full = bars + 6;     // [2] S/T of tile g in TMEM                (MMA -> softmax g)
  uint64_t* p_ready = bars + 8;    // [2] P_g in smem, S_g consumed            (softmax g -> MMA), count 128
  uint64_t* o_full = bars + 10;    // [2] O_g in TMEM                          (MMA -> softmax g)
  uint64_t* o_done = bars + 12;    // [2] O_g read out, slot g free            (softmax g -> MMA), count 128
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 14);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t sbase = ptx::smem_u32(smem);

  if (tid == 0) {
    ptx::prefetch_tmap(&maps.kv64);
    ptx::prefetch_tmap(&maps.kv16);
    ptx::prefetch_tmap(&maps.qa64);
    ptx::prefetch_tmap(&maps.qb64);
    ptx::mbar_init(qk_full, 1);
    ptx::mbar_init(qk_ready, 1);
    ptx::mbar_init(qk_free, 2);
    ptx::mbar_init(v_full, 1);
    ptx::mbar_init(v_ready, 1);
    ptx::mbar_init(v_free, 2);
    for (int g = 0; g < 2; ++g) {
      ptx::mbar_init(&s_full[g], 1);
      ptx::mbar_init(&p_ready[g], 128);
      ptx::mbar_init(&o_full[g], 1);
      ptx::mbar_init(&o_done[g], 128);
    }
    ptx::fence_mbar_init();
  }
  if (warp == 1) {
    ptx::tmem_alloc(tmem_slot, 512);
    ptx::tmem_relinquish();
  }
  // V pad rows 196..207 are never written by TMA: zero them once (P is 0 there, but 0 x garbage could be NaN)
  for (int i = tid; i < (NKEY - NTOK) * 10; i += kThreads3) {
    const int r = NTOK + i / 10, c = i % 10;
    const uint4 z = make_uint4(0, 0, 0, 0);
    if (c < 8)
      *reinterpret_cast<uint4*>(smem + OFF_V64 + row_off64(r, c)) = z;
    else
      *reinterpret_cast<uint4*>(smem + OFF_V16 + row_off16(r, c - 8)) = z;
  }
  ptx::fence_proxy_async_smem();
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  if (tmem != 0) {   // a CTA that owns all 512 columns gets base 0; the MMA issuers rely on it (uniform addresses)
    if (tid == 0) printf("win_attn3: unexpected TMEM base %u\n", tmem);
    __trap();
  }

  if (warp == 0) {
    // ============================================================ Q / K producer: TMA + padded-token patch
    int n = 0;
    for (int it = blockIdx.x; it < num_items; it += gridDim.x, ++n) {
      const Item w = decode_item(it, heads);
      const uint32_t ph = n & 1;
      const int cq = w.head * HD, ck = E + w.head * HD;
      const int x0 = w.wx * WS, y0 = w.wy * WS;
      const bool padded = (w.wy == 4) || (w.wx == 4);
      if (lane == 0) {
        if (n > 0) ptx::mbar_wait(qk_free, ph ^ 1);
        WIN3_TR(0, 0);
        uint32_t bytes = static_cast<uint32_t>((2 * NTOK) * HD * 2);   // Q0 (126 rows) + Q1 (70 rows) + K (196 rows)
        if (n == 0) bytes += 64 * HD * 2;
        ptx::mbar_expect_tx(qk_full, bytes);
        if (n == 0) {
          ptx::tma_load_2d(smem + OFF_R64, &maps.r64, qk_full, 0, 0);
          ptx::tma_load_2d(smem + OFF_R16, &maps.r16, qk_full, 64, 0);
        }
        ptx::tma_load_4d(smem + OFF_Q64, &maps.qa64, qk_full, cq, x0, y0, w.b);
        ptx::tma_load_4d(smem + OFF_Q16, &maps.qa16, qk_full, cq + 64, x0, y0, w.b);
        ptx::tma_load_4d(smem + OFF_Q64 + 16384, &maps.qb64, qk_full, cq, x0, y0 + 9, w.b);
        ptx::tma_load_4d(smem + OFF_Q16 + 4096, &maps.qb16, qk_full, cq + 64, x0, y0 + 9, w.b);
        ptx::tma_load_4d(smem + OFF_K64, &maps.kv64, qk_full, ck, x0, y0, w.b);
        ptx::tma_load_4d(smem + OFF_K16, &maps.kv16, qk_full, ck + 64, x0, y0, w.b);
      }
      __syncwarp();
      if (padded) {
        // token r of the window (iy = r / 14, ix = r % 14) lies outside the 64x64 grid -> q / k := qkv bias
        // (image_encoder.py:281 pads x with zeros BEFORE the qkv projection).  The two bias rows are fetched into
        // registers while the TMA is in flight, so the patch itself is shared-memory stores only.
        uint4 bq[10], bk[10];
#pragma unroll
        for (int c = 0; c < 10; ++c) {
          bq[c] = __ldg(reinterpret_cast<const uint4*>(bias_op + w.head * HD) + c);
          bk[c] = __ldg(reinterpret_cast<const uint4*>(bias_op + E + w.head * HD) + c);
        }
        ptx::mbar_wait(qk_full, ph);
        for (int r = lane; r < NTOK; r += 32) {
          const int iy = r / WS, ix = r % WS;
          if (y0 + iy >= 64 || x0 + ix >= 64) {
            fill_row_regs(smem + OFF_K64, smem + OFF_K16, r, bk);
            if (r < 126)
              fill_row_regs(smem + OFF_Q64, smem + OFF_Q16, r, bq);
            else
              fill_row_regs(smem + OFF_Q64 + 16384, smem + OFF_Q16 + 4096, r - 126, bq);
          }
        }
        ptx::fence_proxy_async_smem();
      } else {
        ptx::mbar_wait(qk_full, ph);
      }
      if (lane == 0) WIN3_TR(0, 1);
      __syncwarp();
      if (lane == 0) {
        WIN3_TR(0, 2);
        ptx::mbar_arrive(qk_ready);
      }
    }
  } else if (warp == 11) {
    // ============================================================ V producer (own warp: its TMA round trip and patch
    // run concurrently with the Q / K producer's instead of behind them)
    int n = 0;
    for (int it = blockIdx.x; it < num_items; it += gridDim.x, ++n) {
      const Item w = decode_item(it, heads);
      const uint32_t ph = n & 1;
      const int cv = 2 * E + w.head * HD;
      const int x0 = w.wx * WS, y0 = w.wy * WS;
      const bool padded = (w.wy == 4) || (w.wx == 4);
      if (lane == 0) {
        if (n > 0) ptx::mbar_wait(v_free, ph ^ 1);
        WIN3_TR(1, 0);
        ptx::mbar_expect_tx(v_full, static_cast<uint32_t>(NTOK * HD * 2));
        ptx::tma_load_4d(smem + OFF_V64, &maps.kv64, v_full, cv, x0, y0, w.b);
        ptx::tma_load_4d(smem + OFF_V16, &maps.kv16, v_full, cv + 64, x0, y0, w.b);
      }
      __syncwarp();
      if (padded) {
        uint4 bv[10];
#pragma unroll
        for (int c = 0; c < 10; ++c) bv[c] = __ldg(reinterpret_cast<const uint4*>(bias_op + cv) + c);
        ptx::mbar_wait(v_full, ph);
        for (int r = lane; r < NTOK; r += 32) {
          const int iy = r / WS, ix = r % WS;
          if (y0 + iy >= 64 || x0 + ix >= 64) fill_row_regs(smem + OFF_V64, smem + OFF_V16, r, bv);
        }
        ptx::fence_proxy_async_smem();
      } else {
        ptx::mbar_wait(v_full, ph);
      }
      __syncwarp();
      if (lane == 0) {
        WIN3_TR(1, 2);
        ptx::mbar_arrive(v_ready);
      }
    }
  } else if (warp == 1 || warp == 10) {
    // ============================================================ MMA issuers: warp 1 -> tile 0, warp 10 -> tile 1
    // elect.sync + a constant TMEM base keep descriptors / addresses on the uniform datapath (2 SASS instructions per
    // UTCHMMA instead of a 12-instruction ELECT / R2UR waterfall); one issuer per tile halves the per-thread MMA count.
    if (ptx::elect_one()) {
      const int g = (warp == 1) ? 0 : 1;
      const uint32_t slot = g * 256;   // TMEM base is 0: this CTA owns all 512 columns (checked after the allocation)
      const uint32_t id_T = ptx::make_idesc((uint32_t)fmt, 128, 32, 0, 0);
      const uint32_t id_S = ptx::make_idesc((uint32_t)fmt, 128, NKEY, 0, 0);
      const uint32_t id_O64 = ptx::make_idesc((uint32_t)fmt, 128, 64, 0, 1);
      const uint32_t id_O16 = ptx::make_idesc((uint32_t)fmt, 128, 16, 0, 1);
      const uint64_t dk64 = ptx::make_smem_desc(sbase + OFF_K64, 16, 1024, ptx::kSwz128);
      const uint64_t dk16 = ptx::make_smem_desc(sbase + OFF_K16, 16, 256, ptx::kSwz32);
      const uint64_t dr64 = ptx::make_smem_desc(sbase + OFF_R64, 16, 1024, ptx::kSwz128);
      const uint64_t dr16 = ptx::make_smem_desc(sbase + OFF_R16, 16, 256, ptx::kSwz32);
      const uint64_t dv64 = ptx::make_smem_desc(sbase + OFF_V64, NKEY * 128, 1024, ptx::kSwz128);
      const uint64_t dv16 = ptx::make_smem_desc(sbase + OFF_V16, NKEY * 32, 256, ptx::kSwz32);
      const uint64_t dq64 = ptx::make_smem_desc(sbase + OFF_Q64 + g * 16384, 16, 1024, ptx::kSwz128);
      const uint64_t dq16 = ptx::make_smem_desc(sbase + OFF_Q16 + g * 4096, 16, 256, ptx::kSwz32);
      const uint64_t dp64 = ptx::make_smem_desc(sbase + OFF_P + g * kPBytes, 16, 1024, ptx::kSwz128);
      const uint64_t dp16 = ptx::make_smem_desc(sbase + OFF_P + g * kPBytes + 49152, 16, 256, ptx::kSwz32);
      int n = 0;
      for (int it = blockIdx.x; it < num_items; it += gridDim.x, ++n) {
        const uint32_t ph = n & 1;
        ptx::mbar_wait(qk_ready, ph);
        // start-up skew: tile 1 enters its first item half a period late (once tile 0's softmax is done), so from then
        // on one tile's MMAs / epilogue run under the other tile's softmax instead of both idling the tensor pipe
        if (n == 0 && g == 1 && skew) ptx::mbar_test_spin(&p_ready[0], 0);
        WIN3_TR(2 + g, 0);
        if (n > 0) ptx::mbar_wait(&o_done[g], ph ^ 1);   // slot g drained by the previous item's epilogue
        WIN3_TR(2 + g, 1);
        ptx::tc_fence_after();
#pragma unroll
        for (int k = 0; k < 4; ++k) ptx::mma_f16_ss(slot, dq64 + 2 * k, dk64 + 2 * k, id_S, k != 0);
        ptx::mma_f16_ss(slot, dq16, dk16, id_S, 1);
        // table rows 32..63 = rel_pos_w (+4096 B / +1024 B), rows 0..31 = rel_pos_h.  Tw is issued after S on
        // purpose: it overwrites the dead pad-key columns 196..207 of S.
#pragma unroll
        for (int k = 0; k < 4; ++k) ptx::mma_f16_ss(slot + 196, dq64 + 2 * k, dr64 + (4096 >> 4) + 2 * k, id_T, k != 0);
        ptx::mma_f16_ss(slot + 196, dq16, dr16 + (1024 >> 4), id_T, 1);
#pragma unroll
        for (int k = 0; k < 4; ++k) ptx::mma_f16_ss(slot + 224, dq64 + 2 * k, dr64 + 2 * k, id_T, k != 0);
        ptx::mma_f16_ss(slot + 224, dq16, dr16, id_T, 1);
        ptx::mma_commit(&s_full[g]);
        ptx::mma_commit(qk_free);       // count 2: both issuers
        WIN3_TR(2 + g, 2);
        ptx::mbar_wait(v_ready, ph);
        WIN3_TR(2 + g, 3);
        ptx::mbar_wait(&p_ready[g], ph);
        WIN3_TR(2 + g, 4);
        ptx::tc_fence_after();
#pragma unroll
        for (int ks = 0; ks < NKEY / 16; ++ks) {
          const uint64_t da = (ks < 12) ? dp64 + (((ks >> 2) * 16384 + (ks & 3) * 32) >> 4) : dp16;
          ptx::mma_f16_ss(slot, da, dv64 + ((ks * 2048) >> 4), id_O64, ks != 0);
          ptx::mma_f16_ss(slot + 64, da, dv16 + ((ks * 512) >> 4), id_O16, ks != 0);
        }
        ptx::mma_commit(&o_full[g]);
        ptx::mma_commit(v_free);        // count 2
        WIN3_TR(2 + g, 5);
      }
    }
  } else {
    // ============================================================ softmax warpgroups (g = query tile)
    const int g = (warp - 2) >> 2;
    const int row = ((warp & 3) << 5) + lane;          // TMEM lane == query row of the tile (warp & 3 = lane quadrant)
    const uint32_t trow = tmem + g * 256 + (static_cast<uint32_t>((warp & 3) * 32) << 16);
    const uint32_t pbase = sbase + OFF_P + g * kPBytes;
    float* sc = reinterpret_cast<float*>(smem + OFF_P + g * kPBytes) + row * kScratchStride;
    const int nq = g ? 70 : 126;
    const int qiy = (g ? 9 : 0) + row / WS;
    const int qix = row % WS;
    const int qh = (qiy < WS) ? qiy : (WS - 1);
    const float kLog2e = 1.4426950408889634f;
    int n = 0;
    for (int it = blockIdx.x; it < num_items; it += gridDim.x, ++n) {
      const Item w = decode_item(it, heads);
      const uint32_t ph = n & 1;
      ptx::mbar_wait(&s_full[g], ph);
      if (row == 0) WIN3_TR(4 + g, 0);
      ptx::tc_fence_after();
      float relh[WS], relw[WS];
      {
        // spill this row's 27 + 27 rel-pos products to scratch, gather the 14 + 14 terms it needs
        // (index = q - k + 13, image_encoder.py:347-351)
        uint32_t v[32];
        ptx::tmem_ld_32x32b_x32(trow + 224, v);
        ptx::tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 27; ++j) sc[j] = __uint_as_float(v[j]) * kLog2e;
        ptx::tmem_ld_32x32b_x32(trow + 196, v);
        ptx::tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 27; ++j) sc[27 + j] = __uint_as_float(v[j]) * kLog2e;
#pragma unroll
        for (int kh = 0; kh < WS; ++kh) relh[kh] = sc[qh - kh + (WS - 1)];
#pragma unroll
        for (int kw = 0; kw < WS; ++kw) relw[kw] = sc[27 + qix - kw + (WS - 1)];
      }
      named_bar_sync(1 + g, 128);   // scratch (aliases P) fully consumed by the whole warpgroup before P is written
      if (row == 0) WIN3_TR(4 + g, 1);
      // Single pass: the reference maximum is the maximum of the first 32 keys; every probability is taken against it
      // (the reference cancels in O / sum).  TMEM loads run one chunk ahead of the arithmetic.  Only if a row sum
      // leaves [0, 2^10] -- a later key beating the reference by a wide margin -- is the row redone the two-pass way.
      float sum;
      {
        uint32_t va[32], vb[32];
        ptx::tmem_ld_32x32b_x32(trow, va);
        ptx::tmem_ld_wait_dep(va);
        ptx::tmem_ld_32x32b_x32(trow + 32, vb);
        float m0 = -INFINITY, m1 = -INFINITY;
#pragma unroll
        for (int i = 0; i < 32; i += 2) {
          m0 = fmaxf(m0, WIN3_LOGIT(i, va[i]));
          m1 = fmaxf(m1, WIN3_LOGIT(i + 1, va[i + 1]));
        }
        const float mref = fmaxf(m0, m1);
#pragma unroll
        for (int kh = 0; kh < WS; ++kh) relh[kh] -= mref;
        const f32x2 sc2 = pk2(scale_log2e, scale_log2e);
        f32x2 s0 = 0ull, s1 = 0ull;
        exp_chunk_regs<0, FMT>(va, relh, relw, sc2, pbase, row, s0, s1);
        ptx::tmem_ld_wait_dep(vb);
        ptx::tmem_ld_32x32b_x32(trow + 64, va);
        exp_chunk_regs<1, FMT>(vb, relh, relw, sc2, pbase, row, s0, s1);
        ptx::tmem_ld_wait_dep(va);
        ptx::tmem_ld_32x32b_x32(trow + 96, vb);
        exp_chunk_regs<2, FMT>(va, relh, relw, sc2, pbase, row, s0, s1);
        ptx::tmem_ld_wait_dep(vb);
        ptx::tmem_ld_32x32b_x32(trow + 128, va);
        exp_chunk_regs<3, FMT>(vb, relh, relw, sc2, pbase, row, s0, s1);
        ptx::tmem_ld_wait_dep(va);
        ptx::tmem_ld_32x32b_x32(trow + 160, vb);
        exp_chunk_regs<4, FMT>(va, relh, relw, sc2, pbase, row, s0, s1);
        ptx::tmem_ld_wait_dep(vb);
        ptx::tmem_ld_32x32b_x16_lo(trow + 192, va);
        exp_chunk_regs<5, FMT>(vb, relh, relw, sc2, pbase, row, s0, s1);
        ptx::tmem_ld_wait_dep(va);
        exp_chunk_regs<6, FMT>(va, relh, relw, sc2, pbase, row, s0, s1);
        float a0, a1;
        upk2(add2(s0, s1), a0, a1);
        sum = a0 + a1;
      }
      // rows >= nq are not query rows (their Q is whatever the shared memory held): they must not trigger the redo
      if (__any_sync(0xffffffffu, (row < nq) && !(sum <= 1024.0f))) {
        float mx = -INFINITY;   // relative to the first reference
        max_chunk<0>(trow, relh, relw, scale_log2e, mx);
        max_chunk<1>(trow, relh, relw, scale_log2e, mx);
        max_chunk<2>(trow, relh, relw, scale_log2e, mx);
        max_chunk<3>(trow, relh, relw, scale_log2e, mx);
        max_chunk<4>(trow, relh, relw, scale_log2e, mx);
        max_chunk<5>(trow, relh, relw, scale_log2e, mx);
        max_chunk<6>(trow, relh, relw, scale_log2e, mx);
#pragma unroll
        for (int kh = 0; kh < WS; ++kh) relh[kh] -= mx;
        sum = 0.f;
        exp_chunk<0, FMT>(trow, relh, relw, scale_log2e, pbase, row, sum);
        exp_chunk<1, FMT>(trow, relh, relw, scale_log2e, pbase, row, sum);
        exp_chunk<2, FMT>(trow, relh, relw, scale_log2e, pbase, row, sum);
        exp_chunk<3, FMT>(trow, relh, relw, scale_log2e, pbase, row, sum);
        exp_chunk<4, FMT>(trow, relh, relw, scale_log2e, pbase, row, sum);
        exp_chunk<5, FMT>(trow, relh, relw, scale_log2e, pbase, row, sum);
        exp_chunk<6, FMT>(trow, relh, relw, scale_log2e, pbase, row, sum);
      }
      ptx::fence_proxy_async_smem();
      ptx::tc_fence_before();
      if (row == 0) WIN3_TR(4 + g, 2);
      ptx::mbar_arrive(&p_ready[g]);

      ptx::mbar_wait(&o_full[g], ph);
      if (row == 0) WIN3_TR(4 + g, 3);
      ptx::tc_fence_after();
      {
        // pull the whole O row (80 fp32) into registers with the loads back to back, hand the TMEM slot back to the
        // MMA issuer at once (its next S overlaps the scaling and the global stores below)
        uint32_t o0[32], o1[32], o2[32];
        ptx::tmem_ld_32x32b_x32(trow, o0);
        ptx::tmem_ld_32x32b_x32(trow + 32, o1);
        ptx::tmem_ld_32x32b_x16_lo(trow + 64, o2);
        ptx::tmem_ld_wait_dep(o0);
        ptx::tmem_ld_wait_dep(o1);
        ptx::tmem_ld_wait_dep(o2);
        ptx::tc_fence_before();
        if (row == 0) WIN3_TR(4 + g, 4);
        ptx::mbar_arrive(&o_done[g]);
        const float inv = 1.0f / sum;
        const int y = w.wy * WS + qiy, x = w.wx * WS + qix;
        const bool ok = (row < nq) && (y < 64) && (x < 64);
        if (ok) {
          uint4* dst = reinterpret_cast<uint4*>(out + (static_cast<size_t>(w.b) * 4096 + (y * 64 + x)) * E + w.head * HD);
#pragma unroll
          for (int c = 0; c < 10; ++c) {
            const uint32_t* v = (c < 4) ? &o0[c * 8] : (c < 8) ? &o1[(c - 4) * 8] : &o2[(c - 8) * 8];
            uint4 u;
            u.x = ptx::pack2t<FMT>(__uint_as_float(v[0]) * inv, __uint_as_float(v[1]) * inv);
            u.y = ptx::pack2t<FMT>(__uint_as_float(v[2]) * inv, __uint_as_float(v[3]) * inv);
            u.z = ptx::pack2t<FMT>(__uint_as_float(v[4]) * inv, __uint_as_float(v[5]) * inv);
            u.w = ptx::pack2t<FMT>(__uint_as_float(v[6]) * inv, __uint_as_float(v[7]) * inv);
            dst[c] = u;
          }
        }
      }
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem, 512);
  }
}

}  // namespace

int samk_attn_window3(const void* qkv, const void* bias_op, const void* rel_tab, void* out, int B, int E, int heads,
                      int fmt, cudaStream_t stream) {
  SAM_REQUIRE(fmt == 0 || fmt == 1, "attn_window: fmt must be fp16/bf16");
  SAM_REQUIRE(E == heads * HD, "attn_window: head_dim must be 80 (E=%d heads=%d)", E, heads);
  SAM_REQUIRE(B > 0, "attn_window: empty batch");
  WinAttnMaps3 maps;
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
  int rc = samhost::encode_tmap_2d(&maps.r64, 2, is_bf16, rel_tab, HD, 64, HD * 2, 64, 64, 3);
  if (rc) return rc;
  rc = samhost::encode_tmap_2d(&maps.r16, 2, is_bf16, rel_tab, HD, 64, HD * 2, 16, 64, 1);
  if (rc) return rc;
  static bool attr_done = false;
  if (!attr_done) {
    SAM_CHECK_CUDA(cudaFuncSetAttribute(win_attn3_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes3));
    SAM_CHECK_CUDA(cudaFuncSetAttribute(win_attn3_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes3));
    attr_done = true;
  }
  const float scale_log2e = 1.4426950408889634f / sqrtf(static_cast<float>(HD));
  const int num_items = B * 25 * heads;
  int grid = samhost::sm_count();
  if (grid > num_items) grid = num_items;
  const double wh = static_cast<double>(num_items);
  samhost::LaunchScope scope(samhost::KC_ATTN_WINDOW, stream, wh * (4.0 * 196 * 196 * 80 + 4.0 * 196 * 14 * 80),
                             static_cast<double>(B) * 4096 * E * 2 * 4);
  static const int skew = getenv("SAM_WIN3_NOSKEW") ? 0 : 1;
  long long* trace = nullptr;
  static const bool want_trace = getenv("SAM_WIN3_TRACE") != nullptr;   // debug only: synchronises and prints
  if (want_trace) {
    SAM_CHECK_CUDA(cudaMalloc(&trace, 6 * 8 * 8 * sizeof(long long)));
    SAM_CHECK_CUDA(cudaMemset(trace, 0, 6 * 8 * 8 * sizeof(long long)));
  }
  if (fmt == 0)
    win_attn3_kernel<0><<<grid, kThreads3, kSmemBytes3, stream>>>(maps, static_cast<const uint16_t*>(bias_op),
                                                                   static_cast<uint16_t*>(out), E, heads, num_items,
                                                                   scale_log2e, trace, skew);
  else
    win_attn3_kernel<1><<<grid, kThreads3, kSmemBytes3, stream>>>(maps, static_cast<const uint16_t*>(bias_op),
                                                                   static_cast<uint16_t*>(out), E, heads, num_items,
                                                                   scale_log2e, trace, skew);
  SAM_CHECK_CUDA(cudaGetLastError());
  if (want_trace) {
    long long h[6 * 8 * 8];
    SAM_CHECK_CUDA(cudaStreamSynchronize(stream));
    SAM_CHECK_CUDA(cudaMemcpy(h, trace, sizeof(h), cudaMemcpyDeviceToHost));
    cudaFree(trace);
    long long t0 = 0;
    for (int i = 0; i < 6 * 8 * 8; ++i)
      if (h[i] && (t0 == 0 || h[i] < t0)) t0 = h[i];
    static const char* roles[6] = {"prodQK", "prodV", "mma0", "mma1", "soft0", "soft1"};
    for (int r = 0; r < 6; ++r)
      for (int e = 0; e < 8; ++e) {
        bool any = false;
        for (int n = 0; n < 8; ++n) any |= h[(r * 8 + e) * 8 + n] != 0;
        if (!any) continue;
        fprintf(stderr, "[win3 trace] %-7s ev%d:", roles[r], e);
        for (int n = 0; n < 8; ++n) fprintf(stderr, " %7lld", h[(r * 8 + e) * 8 + n] ? h[(r * 8 + e) * 8 + n] - t0 : -1);
        fprintf(stderr, "\n");
      }
  }
  return 0;
}
