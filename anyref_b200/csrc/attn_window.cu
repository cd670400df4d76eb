// 14x14 windowed attention of the SAM ViT-H encoder: probabilities in TENSOR MEMORY, double-buffered operands.
// Replaces image_encoder.py:235-257 (Attention.forward), :263-318 (window partition / un-partition) and :354-392
// (add_decomposed_rel_pos) for the 14x14 windowed blocks.
//
// An earlier version kept the probabilities in shared memory (104 KB), so Q / K / V were single-buffered and the two
// query tiles of an item waited for each other at every hand-off.  Here
//   * P never touches shared memory: each softmax thread packs its row's probabilities and writes them with tcgen05.st
//     IN PLACE over the S columns it has already consumed; O = P.V is a tcgen05.mma with the A operand in tensor memory
//     ("ts" form, layout pinned by tools/gpu_probe_ts.py) and V as five 16-wide MN-major SWIZZLE_32B chunks, so one
//     N = 80 MMA per 16 keys (13 per tile instead of 26 and no A-operand shared-memory reads);
//   * the freed shared memory holds a 2-stage ring of Q / K / V: the producers run a whole item ahead and the two
//     tiles free-run -- one tile's MMAs and epilogue hide under the other tile's softmax;
//   * the rel-pos gather uses the tile's own (dead) Q buffer as thread-private, bank-conflict-free scratch: no barrier.
// Softmax is single pass against a reference maximum (max of the first 32 keys); if the running row sum leaves
// [0, 2^10] the probabilities written so far are rescaled in tensor memory and the reference moved -- exact softmax.
//
//   warp 0 / 11  : producers (Q0,Q1,K / V): 4-D TMA boxes straight from the un-partitioned qkv, padded-token patch
//   warp 1 / 10  : MMA issuers of tile 0 / 1 (one elected thread each)
//   warps 2..5   : softmax + epilogue of tile 0 (query rows 0..125);  warps 6..9: tile 1 (rows 126..195)
// TMEM slot g (256 columns): S [0,208) | Tw [196,228) Th [224,256) | P (16-bit pairs) [0,104) | O [112,192).
#include <stdio.h>
#include <stdlib.h>

#include "host_common.h"
#include "kernels.h"
#include "ptx.cuh"

namespace {

// head_dim HD is a template parameter: 80 (ViT-H: a 64-wide SWIZZLE_128B tile + a 16-wide SWIZZLE_32B tile per operand)
// or 64 (ViT-L / ViT-B: the 64-wide tile alone)
constexpr int WS = 14;
constexpr int NTOK = WS * WS;  // 196
constexpr int NKEY = 208;
constexpr int kThreads4 = 384;

// shared-memory map (bytes from the 1024-aligned base); stage s at + s * kStageBytes
constexpr int OFF_Q64 = 0;            // 2 x (128 x 128B) SWIZZLE_128B   (tile g at + g*16384); doubles as gather scratch
constexpr int OFF_K64 = 32768;        // 208 x 128B
constexpr int OFF_Q16 = 59392;        // 2 x (128 x 32B) SWIZZLE_32B     (tile g at + g*4096)
constexpr int OFF_K16 = 67584;        // 208 x 32B
constexpr int OFF_V = 74240;          // 5 chunks x (208 x 32B) SWIZZLE_32B, MN-major operand of P.V
constexpr int kVChunk = NKEY * 32;    // 6656
constexpr int kStageBytes = OFF_V + 5 * kVChunk;   // 107520 (multiple of 1024)
constexpr int OFF_R64 = 2 * kStageBytes;           // 64 x 128B rel-pos operand table
constexpr int OFF_R16 = OFF_R64 + 8192;            // 64 x 32B
constexpr int OFF_BAR = OFF_R16 + 2048;
constexpr int kSmemBytes4 = OFF_BAR + 256 + 1024;

constexpr uint32_t TM_O = 112;
constexpr float kSumLimit = 1024.0f;

struct WinAttnMaps4 {
  CUtensorMap kv64, kv16;    // box {64|16, 14, 14, 1}
  CUtensorMap qa64, qa16;    // box {64|16, 14, 9, 1}   query tile 0
  CUtensorMap qb64, qb16;    // box {64|16, 14, 5, 1}   query tile 1
  CUtensorMap r64, r16;      // rel-pos operand table [64, 80]: box {64|16, 64}
  CUtensorMap oa64, oa16;    // WIN4_TMA_OUT: the query-tile boxes over `out` [B, 64, 64, E] (stores clip the padded rows)
  CUtensorMap ob64, ob16;
};

// Output through shared memory + TMA stores: every softmax thread parks its normalised O row in the tile's (dead) Q
// buffer of the stage, in the layout the Q load used, and one thread per tile issues 4-D bulk tensor stores with the
// load's box -- rows of padded window tokens fall outside the tensor and are clipped.  Replaces ten STG.128 per thread
// whose warp-level instructions touched 32 different half sectors each (ncu: the item loop stalled ~1200 cycles per
// item until those stores had drained).  The stage's Q / K buffers are handed back to the producer when the store has
// READ the staging rows (checked at the top of the next item), not right after the rel-pos gather.
#ifndef WIN4_TMA_OUT
#define WIN4_TMA_OUT 1
#endif
__device__ __forceinline__ void named_bar_sync4(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
__device__ __forceinline__ void tma_store_4d(const CUtensorMap* m, const void* src, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];" ::"l"(m),
               "r"(ptx::smem_u32(src)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
               : "memory");
}
__device__ __forceinline__ void bulk_commit4() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read_all4() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_all4() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

using ptx::add2;
using ptx::f32x2;
using ptx::fma2;
using ptx::pk2;
using ptx::upk2;

__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ uint32_t row_off64(int r, int c) { return r * 128 + ((c ^ (r & 7)) << 4); }
__device__ __forceinline__ uint32_t row_off16(int r, int c) { return r * 32 + ((c ^ ((r >> 2) & 1)) << 4); }

__device__ __forceinline__ void tmem_st_x16(uint32_t taddr, const uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16};" ::"r"(taddr),
      "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]),
      "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
      : "memory");
}
__device__ __forceinline__ void tmem_st_x8(uint32_t taddr, const uint32_t (&v)[16]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(v[0]),
               "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
               : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_ld_x16(uint32_t taddr, uint32_t (&v)[16]) {
  ptx::tmem_ld_32x32b_x16(taddr, v);
}

#define WIN4_LOGIT(J, VAL) (fmaf(__uint_as_float(VAL), scale_log2e, relh[(J) / WS]) + relw[(J) % WS])

// One 32-key chunk (16 keys for C == 6) already in registers: probabilities against the reference folded into relh,
// two at a time on the packed fp32 pipe; packed 16-bit pairs go straight back to tensor memory (columns C*16 ..).
template <int C, int FMT>
__device__ __forceinline__ void exp_chunk_tm(const uint32_t (&v)[32], const float (&relh)[WS], const float (&relw)[WS],
                                             f32x2 sc2, uint32_t trow, f32x2& s0, f32x2& s1) {
  constexpr int kN = (C < 6) ? 32 : 16;
  uint32_t pk[16];
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
      pk[i >> 1] = 0u;   // pad keys 196..207: P = 0
    }
  }
  if (C < 6)
    tmem_st_x16(trow + C * 16, pk);
  else
    tmem_st_x8(trow + C * 16, pk);
}

template <int C>
__device__ __forceinline__ float chunk_max(const uint32_t (&v)[32], const float (&relh)[WS], const float (&relw)[WS],
                                           float scale_log2e) {
  constexpr int kN = (C < 6) ? 32 : 16;
  float m0 = -INFINITY, m1 = -INFINITY;
#pragma unroll
  for (int i = 0; i < kN; i += 2) {
    const int j = C * 32 + i;
    if (j < NTOK) {
      const int jj = j < NTOK ? j : 0;
      m0 = fmaxf(m0, WIN4_LOGIT(jj, v[i]));
      m1 = fmaxf(m1, WIN4_LOGIT(jj + 1, v[i + 1]));
    }
  }
  return fmaxf(m0, m1);
}

// Rare path: the running row sum left [0, kSumLimit] at chunk C.  Move the reference to this chunk's maximum, rescale
// the probabilities of chunks 0 .. C-1 in tensor memory (a power of two: exact), and redo chunk C.  Warp-uniform (the
// tcgen05 instructions are .aligned); lanes that did not overflow use delta = 0.
template <int C, int FMT>
__device__ __forceinline__ void rescale_and_redo(const uint32_t (&v)[32], float (&relh)[WS], const float (&relw)[WS],
                                                 float scale_log2e, uint32_t trow, f32x2& s0, f32x2& s1, f32x2 prev0,
                                                 f32x2 prev1, bool mine) {
  tmem_st_wait();   // the probabilities written so far must have landed before they are read back
  const float cmax = chunk_max<C>(v, relh, relw, scale_log2e);
  const float delta = mine ? fmaxf(cmax, 0.0f) : 0.0f;
  const float alpha = ex2(-delta);
#pragma unroll 1
  for (int c = 0; c < C; ++c) {
    uint32_t p[16];
    tmem_ld_x16(trow + c * 16, p);
    ptx::tmem_ld_wait();
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      const float2 f = ptx::unpack2(p[i], FMT);
      p[i] = ptx::pack2t<FMT>(f.x * alpha, f.y * alpha);
    }
    tmem_st_x16(trow + c * 16, p);
  }
#pragma unroll
  for (int kh = 0; kh < WS; ++kh) relh[kh] -= delta;
  const f32x2 a2 = pk2(alpha, alpha);
  s0 = ptx::mul2(prev0, a2);
  s1 = ptx::mul2(prev1, a2);
  exp_chunk_tm<C, FMT>(v, relh, relw, pk2(scale_log2e, scale_log2e), trow, s0, s1);
}

template <int C, int FMT>
__device__ __forceinline__ void softmax_chunk(const uint32_t (&v)[32], float (&relh)[WS], const float (&relw)[WS],
                                              float scale_log2e, uint32_t trow, f32x2& s0, f32x2& s1, bool valid) {
  const f32x2 prev0 = s0, prev1 = s1;
  exp_chunk_tm<C, FMT>(v, relh, relw, pk2(scale_log2e, scale_log2e), trow, s0, s1);
  if (C > 0) {
    float a0, a1;
    upk2(add2(s0, s1), a0, a1);
    const bool over = valid && !(a0 + a1 <= kSumLimit);   // rows >= nq hold no query: never trigger
    if (__any_sync(0xffffffffu, over)) rescale_and_redo<C, FMT>(v, relh, relw, scale_log2e, trow, s0, s1, prev0, prev1, over);
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

template <int FMT, int HD>
__global__ void __launch_bounds__(kThreads4, 1)
win_attn4_kernel(const __grid_constant__ WinAttnMaps4 maps, const uint16_t* __restrict__ bias_op,
                 uint16_t* __restrict__ out, const int E, const int heads, const int num_items,
                 const float scale_log2e) {
  constexpr int fmt = FMT;
  constexpr bool kTail = (HD > 64);   // operands have a 16-wide tail beyond the 64-wide tile
  constexpr int kU4 = HD / 8;         // 16-byte units per operand row
  constexpr int kVCh = HD / 16;       // 16-wide V chunks
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + OFF_BAR);
  uint64_t* qk_full = bars + 0;    // [2 stages] TMA bytes of Q0/Q1/K (+ R with the first item)
  uint64_t* qk_ready = bars + 2;   // [2] ... and padded tokens patched            (producer -> MMA issuers)
  uint64_t* qk_free = bars + 4;    // [2] S/T MMAs of both tiles done + gather scratch released (count 2 + 256)
  uint64_t* v_full = bars + 6;     // [2]
  uint64_t* v_ready = bars + 8;    // [2]
  uint64_t* v_free = bars + 10;    // [2] PV MMAs of both tiles done (count 2)
  uint64_t* s_full = bars + 12;    // [2 tiles] S/T in TMEM                       (MMA -> softmax)
  uint64_t* p_ready = bars + 14;   // [2 tiles] P in TMEM, count 128              (softmax -> MMA)
  uint64_t* o_full = bars + 16;    // [2 tiles] O in TMEM                         (MMA -> softmax)
  uint64_t* o_done = bars + 18;    // [2 tiles] O read out, slot free, count 128  (softmax -> MMA)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 20);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t sbase = ptx::smem_u32(smem);

  if (tid == 0) {
    ptx::prefetch_tmap(&maps.kv64);
    ptx::prefetch_tmap(&maps.kv16);
    ptx::prefetch_tmap(&maps.qa64);
    ptx::prefetch_tmap(&maps.qb64);
    for (int s = 0; s < 2; ++s) {
      ptx::mbar_init(&qk_full[s], 1);
      ptx::mbar_init(&qk_ready[s], 1);
      ptx::mbar_init(&qk_free[s], WIN4_TMA_OUT ? 2 + 2 : 2 + 256);
      ptx::mbar_init(&v_full[s], 1);
      ptx::mbar_init(&v_ready[s], 1);
      ptx::mbar_init(&v_free[s], 2);
      ptx::mbar_init(&s_full[s], 1);
      ptx::mbar_init(&p_ready[s], 128);
      ptx::mbar_init(&o_full[s], 1);
      ptx::mbar_init(&o_done[s], 128);
    }
    ptx::fence_mbar_init();
  }
  if (warp == 1) {
    ptx::tmem_alloc(tmem_slot, 512);
    ptx::tmem_relinquish();
  }
  // V pad rows 196..207 are never written by TMA: zero them once in both stages (P is 0 there, but 0 x garbage could
  // be NaN)
  for (int i = tid; i < 2 * 5 * (NKEY - NTOK) * 2; i += kThreads4) {
    const int u = i & 1, r = NTOK + (i >> 1) % (NKEY - NTOK), c = ((i >> 1) / (NKEY - NTOK)) % 5, s = (i >> 1) / ((NKEY - NTOK) * 5);
    *reinterpret_cast<uint4*>(smem + s * kStageBytes + OFF_V + c * kVChunk + row_off16(r, u)) = make_uint4(0, 0, 0, 0);
  }
  ptx::fence_proxy_async_smem();
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  if (tmem != 0) {   // a CTA that owns all 512 columns gets base 0; the MMA issuers rely on it (uniform addresses)
    if (tid == 0) printf("win_attn4: unexpected TMEM base %u\n", tmem);
    __trap();
  }

  if (warp == 0) {
    // ============================================================ Q / K producer: TMA + padded-token patch
    int n = 0;
    for (int it = blockIdx.x; it < num_items; it += gridDim.x, ++n) {
      const Item w = decode_item(it, heads);
      const int s = n & 1;
      const uint32_t sph = (n >> 1) & 1;
      uint8_t* st = smem + s * kStageBytes;
      const int cq = w.head * HD, ck = E + w.head * HD;
      const int x0 = w.wx * WS, y0 = w.wy * WS;
      const bool padded = (w.wy == 4) || (w.wx == 4);
      if (lane == 0) {
        if (n >= 2) ptx::mbar_wait(&qk_free[s], sph ^ 1);
        uint32_t bytes = static_cast<uint32_t>((2 * NTOK) * HD * 2);   // Q0 (126 rows) + Q1 (70 rows) + K (196 rows)
        if (n == 0) bytes += 64 * HD * 2;
        ptx::mbar_expect_tx(&qk_full[s], bytes);
        if (n == 0) {
          ptx::tma_load_2d(smem + OFF_R64, &maps.r64, &qk_full[s], 0, 0);
          if (kTail) ptx::tma_load_2d(smem + OFF_R16, &maps.r16, &qk_full[s], 64, 0);
        }
        ptx::tma_load_4d(st + OFF_Q64, &maps.qa64, &qk_full[s], cq, x0, y0, w.b);
        ptx::tma_load_4d(st + OFF_Q64 + 16384, &maps.qb64, &qk_full[s], cq, x0, y0 + 9, w.b);
        ptx::tma_load_4d(st + OFF_K64, &maps.kv64, &qk_full[s], ck, x0, y0, w.b);
        if (kTail) {
          ptx::tma_load_4d(st + OFF_Q16, &maps.qa16, &qk_full[s], cq + 64, x0, y0, w.b);
          ptx::tma_load_4d(st + OFF_Q16 + 4096, &maps.qb16, &qk_full[s], cq + 64, x0, y0 + 9, w.b);
          ptx::tma_load_4d(st + OFF_K16, &maps.kv16, &qk_full[s], ck + 64, x0, y0, w.b);
        }
      }
      __syncwarp();
      if (padded) {
        // token r of the window (iy = r / 14, ix = r % 14) lies outside the 64x64 grid -> q / k := qkv bias
        // (image_encoder.py:281 pads x with zeros BEFORE the qkv projection).  The two bias rows are fetched into
        // registers while the TMA is in flight, so the patch itself is shared-memory stores only.
        uint4 bq[kU4], bk[kU4];
#pragma unroll
        for (int c = 0; c < kU4; ++c) {
          bq[c] = __ldg(reinterpret_cast<const uint4*>(bias_op + w.head * HD) + c);
          bk[c] = __ldg(reinterpret_cast<const uint4*>(bias_op + E + w.head * HD) + c);
        }
        ptx::mbar_wait(&qk_full[s], sph);
        for (int r = lane; r < NTOK; r += 32) {
          const int iy = r / WS, ix = r % WS;
          if (y0 + iy >= 64 || x0 + ix >= 64) {
            uint8_t* q64 = st + OFF_Q64 + (r < 126 ? 0 : 16384);
            uint8_t* q16 = st + OFF_Q16 + (r < 126 ? 0 : 4096);
            const int rq = r < 126 ? r : r - 126;
#pragma unroll
            for (int c = 0; c < kU4; ++c) {
              if (c < 8) {
                *reinterpret_cast<uint4*>(st + OFF_K64 + row_off64(r, c)) = bk[c];
                *reinterpret_cast<uint4*>(q64 + row_off64(rq, c)) = bq[c];
              } else {
                *reinterpret_cast<uint4*>(st + OFF_K16 + row_off16(r, c - 8)) = bk[c];
                *reinterpret_cast<uint4*>(q16 + row_off16(rq, c - 8)) = bq[c];
              }
            }
          }
        }
        ptx::fence_proxy_async_smem();
      } else {
        ptx::mbar_wait(&qk_full[s], sph);
      }
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&qk_ready[s]);
    }
  } else if (warp == 11) {
    // ============================================================ V producer
    int n = 0;
    for (int it = blockIdx.x; it < num_items; it += gridDim.x, ++n) {
      const Item w = decode_item(it, heads);
      const int s = n & 1;
      const uint32_t sph = (n >> 1) & 1;
      uint8_t* sv = smem + s * kStageBytes + OFF_V;
      const int cv = 2 * E + w.head * HD;
      const int x0 = w.wx * WS, y0 = w.wy * WS;
      const bool padded = (w.wy == 4) || (w.wx == 4);
      if (lane == 0) {
        if (n >= 2) ptx::mbar_wait(&v_free[s], sph ^ 1);
        ptx::mbar_expect_tx(&v_full[s], static_cast<uint32_t>(NTOK * HD * 2));
#pragma unroll
        for (int c = 0; c < kVCh; ++c) ptx::tma_load_4d(sv + c * kVChunk, &maps.kv16, &v_full[s], cv + 16 * c, x0, y0, w.b);
      }
      __syncwarp();
      if (padded) {
        uint4 bv[kU4];
#pragma unroll
        for (int c = 0; c < kU4; ++c) bv[c] = __ldg(reinterpret_cast<const uint4*>(bias_op + cv) + c);
        ptx::mbar_wait(&v_full[s], sph);
        for (int r = lane; r < NTOK; r += 32) {
          const int iy = r / WS, ix = r % WS;
          if (y0 + iy >= 64 || x0 + ix >= 64) {
#pragma unroll
            for (int c = 0; c < kU4; ++c)
              *reinterpret_cast<uint4*>(sv + (c >> 1) * kVChunk + row_off16(r, c & 1)) = bv[c];
          }
        }
        ptx::fence_proxy_async_smem();
      } else {
        ptx::mbar_wait(&v_full[s], sph);
      }
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&v_ready[s]);
    }
  } else if (warp == 1 || warp == 10) {
    // ============================================================ MMA issuers: warp 1 -> tile 0, warp 10 -> tile 1
    if (ptx::elect_one()) {
      const int g = (warp == 1) ? 0 : 1;
      const uint32_t slot = g * 256;   // TMEM base is 0 (checked above)
      const uint32_t id_T = ptx::make_idesc((uint32_t)fmt, 128, 32, 0, 0);
      const uint32_t id_S = ptx::make_idesc((uint32_t)fmt, 128, NKEY, 0, 0);
      const uint32_t id_O = ptx::make_idesc((uint32_t)fmt, 128, HD, 0, 1);
      const uint64_t dr64 = ptx::make_smem_desc(sbase + OFF_R64, 16, 1024, ptx::kSwz128);
      const uint64_t dr16 = ptx::make_smem_desc(sbase + OFF_R16, 16, 256, ptx::kSwz32);
      int n = 0;
      for (int it = blockIdx.x; it < num_items; it += gridDim.x, ++n) {
        const int s = n & 1;
        const uint32_t sph = (n >> 1) & 1, ph = n & 1;
        const uint32_t sb = sbase + s * kStageBytes;
        const uint64_t dk64 = ptx::make_smem_desc(sb + OFF_K64, 16, 1024, ptx::kSwz128);
        const uint64_t dk16 = ptx::make_smem_desc(sb + OFF_K16, 16, 256, ptx::kSwz32);
        const uint64_t dq64 = ptx::make_smem_desc(sb + OFF_Q64 + g * 16384, 16, 1024, ptx::kSwz128);
        const uint64_t dq16 = ptx::make_smem_desc(sb + OFF_Q16 + g * 4096, 16, 256, ptx::kSwz32);
        // V: MN-major, five 16-wide SWIZZLE_32B chunks (LBO = chunk stride), 8-key groups of 256 B (SBO)
        const uint64_t dv = ptx::make_smem_desc(sb + OFF_V, kVChunk, 256, ptx::kSwz32);
        ptx::mbar_wait(&qk_ready[s], sph);
        if (n > 0) ptx::mbar_wait(&o_done[g], ph ^ 1);   // slot g drained by the previous item's epilogue
        ptx::tc_fence_after();
#pragma unroll
        for (int k = 0; k < 4; ++k) ptx::mma_f16_ss(slot, dq64 + 2 * k, dk64 + 2 * k, id_S, k != 0);
        if (kTail) ptx::mma_f16_ss(slot, dq16, dk16, id_S, 1);
        // table rows 32..63 = rel_pos_w (+4096 B / +1024 B), rows 0..31 = rel_pos_h.  Tw is issued after S on
        // purpose: it overwrites the dead pad-key columns 196..207 of S.
#pragma unroll
        for (int k = 0; k < 4; ++k) ptx::mma_f16_ss(slot + 196, dq64 + 2 * k, dr64 + (4096 >> 4) + 2 * k, id_T, k != 0);
        if (kTail) ptx::mma_f16_ss(slot + 196, dq16, dr16 + (1024 >> 4), id_T, 1);
#pragma unroll
        for (int k = 0; k < 4; ++k) ptx::mma_f16_ss(slot + 224, dq64 + 2 * k, dr64 + 2 * k, id_T, k != 0);
        if (kTail) ptx::mma_f16_ss(slot + 224, dq16, dr16, id_T, 1);
        ptx::mma_commit(&s_full[g]);
        ptx::mma_commit(&qk_free[s]);
        ptx::mbar_wait(&v_ready[s], sph);
        ptx::mbar_wait(&p_ready[g], ph);
        ptx::tc_fence_after();
#pragma unroll
        for (int ks = 0; ks < NKEY / 16; ++ks)
          ptx::mma_f16_ts(slot + TM_O, slot + ks * 8, dv + ((ks * 512) >> 4), id_O, ks != 0);
        ptx::mma_commit(&o_full[g]);
        ptx::mma_commit(&v_free[s]);
      }
    }
  } else {
    // ============================================================ softmax warpgroups (g = query tile)
    const int g = (warp - 2) >> 2;
    const int row = ((warp & 3) << 5) + lane;          // TMEM lane == query row of the tile (warp & 3 = lane quadrant)
    const uint32_t trow = tmem + g * 256 + (static_cast<uint32_t>((warp & 3) * 32) << 16);
    const int nq = g ? 70 : 126;
    const int qiy = (g ? 9 : 0) + row / WS;
    const int qix = row % WS;
    const int qh = (qiy < WS) ? qiy : (WS - 1);
    const float kLog2e = 1.4426950408889634f;
    int n = 0;
    for (int it = blockIdx.x; it < num_items; it += gridDim.x, ++n) {
      const Item w = decode_item(it, heads);
      const int s = n & 1;
      const uint32_t ph = n & 1;
      if (WIN4_TMA_OUT && n > 0 && (warp & 3) == 2 && lane == 0) {
        // the previous item's output store has read its staging rows: that stage's Q / K may be overwritten
        bulk_wait_read_all4();
        ptx::mbar_arrive(&qk_free[s ^ 1]);
      }
      ptx::mbar_wait(&s_full[g], ph);
      ptx::tc_fence_after();
      float relh[WS], relw[WS];
      {
        // rel-pos products of this row: Th[27] (cols 224..250), Tw[27] (cols 196..222).  The 14 + 14 terms the row needs
        // sit at a row-dependent offset (index = q - k + 13, image_encoder.py:347-351): bounce them through thread-
        // private scratch in this tile's Q buffer (dead once S / T are in TMEM), word (j, row) at j*128 + row: no
        // bank conflicts, no other thread involved, no barrier.
        uint32_t th[32], tw[32];
        ptx::tmem_ld_32x32b_x32(trow + 224, th);
        ptx::tmem_ld_32x32b_x32(trow + 196, tw);
        ptx::tmem_ld_wait_dep(th);
        ptx::tmem_ld_wait_dep(tw);
        const uint32_t sc = sbase + s * kStageBytes + OFF_Q64 + g * 16384 + row * 4;
#pragma unroll
        for (int j = 0; j < 27; ++j)
          asm volatile("st.shared.f32 [%0], %1;" ::"r"(sc + j * 512), "f"(__uint_as_float(th[j]) * kLog2e) : "memory");
        const uint32_t ah = sc + (qh + (WS - 1)) * 512;
#pragma unroll
        for (int kh = 0; kh < WS; ++kh)
          asm volatile("ld.shared.f32 %0, [%1];" : "=f"(relh[kh]) : "r"(ah - kh * 512) : "memory");
#pragma unroll
        for (int j = 0; j < 27; ++j)
          asm volatile("st.shared.f32 [%0], %1;" ::"r"(sc + j * 512), "f"(__uint_as_float(tw[j]) * kLog2e) : "memory");
        const uint32_t aw = sc + (qix + (WS - 1)) * 512;
#pragma unroll
        for (int kw = 0; kw < WS; ++kw)
          asm volatile("ld.shared.f32 %0, [%1];" : "=f"(relw[kw]) : "r"(aw - kw * 512) : "memory");
      }
      // the scratch stores above went through the generic proxy; the producer's next TMA into this stage writes the
      // same bytes through the async proxy.  Without this fence a late scratch store can land on top of the freshly
      // loaded Q rows (seen as a few wrong rows of one item in ~4 % of stress runs).
      if (!WIN4_TMA_OUT) {
        ptx::fence_proxy_async_smem();
        ptx::mbar_arrive(&qk_free[s]);   // Q / K of this stage may be overwritten (count 2 MMA commits + 256 threads)
      }

      f32x2 s0 = 0ull, s1 = 0ull;
      {
        uint32_t va[32], vb[32];
        ptx::tmem_ld_32x32b_x32(trow, va);
        ptx::tmem_ld_wait_dep(va);
        ptx::tmem_ld_32x32b_x32(trow + 32, vb);
        const float mref = chunk_max<0>(va, relh, relw, scale_log2e);
#pragma unroll
        for (int kh = 0; kh < WS; ++kh) relh[kh] -= mref;
        softmax_chunk<0, FMT>(va, relh, relw, scale_log2e, trow, s0, s1, row < nq);
        ptx::tmem_ld_wait_dep(vb);
        ptx::tmem_ld_32x32b_x32(trow + 64, va);
        softmax_chunk<1, FMT>(vb, relh, relw, scale_log2e, trow, s0, s1, row < nq);
        ptx::tmem_ld_wait_dep(va);
        ptx::tmem_ld_32x32b_x32(trow + 96, vb);
        softmax_chunk<2, FMT>(va, relh, relw, scale_log2e, trow, s0, s1, row < nq);
        ptx::tmem_ld_wait_dep(vb);
        ptx::tmem_ld_32x32b_x32(trow + 128, va);
        softmax_chunk<3, FMT>(vb, relh, relw, scale_log2e, trow, s0, s1, row < nq);
        ptx::tmem_ld_wait_dep(va);
        ptx::tmem_ld_32x32b_x32(trow + 160, vb);
        softmax_chunk<4, FMT>(va, relh, relw, scale_log2e, trow, s0, s1, row < nq);
        ptx::tmem_ld_wait_dep(vb);
        ptx::tmem_ld_32x32b_x16_lo(trow + 192, va);
        softmax_chunk<5, FMT>(vb, relh, relw, scale_log2e, trow, s0, s1, row < nq);
        ptx::tmem_ld_wait_dep(va);
        softmax_chunk<6, FMT>(va, relh, relw, scale_log2e, trow, s0, s1, row < nq);
      }
      float a0, a1;
      upk2(add2(s0, s1), a0, a1);
      const float sum = a0 + a1;
      tmem_st_wait();
      ptx::tc_fence_before();
      ptx::mbar_arrive(&p_ready[g]);

      ptx::mbar_wait(&o_full[g], ph);
      ptx::tc_fence_after();
      {
        // pull the whole O row (80 fp32) into registers with the loads back to back, hand the TMEM slot back to the
        // MMA issuer at once (its next S overlaps the scaling and the global stores below)
        uint32_t o0[32], o1[32], o2[32];
        ptx::tmem_ld_32x32b_x32(trow + TM_O, o0);
        ptx::tmem_ld_32x32b_x32(trow + TM_O + 32, o1);
        if (kTail) ptx::tmem_ld_32x32b_x16_lo(trow + TM_O + 64, o2);
        ptx::tmem_ld_wait_dep(o0);
        ptx::tmem_ld_wait_dep(o1);
        if (kTail) ptx::tmem_ld_wait_dep(o2);
        ptx::tc_fence_before();
        ptx::mbar_arrive(&o_done[g]);
        const float inv = 1.0f / sum;
        const int y = w.wy * WS + qiy, x = w.wx * WS + qix;
        const bool ok = (row < nq) && (y < 64) && (x < 64);
        if (WIN4_TMA_OUT) {
          uint8_t* st = smem + s * kStageBytes;
          if (row < nq) {
#pragma unroll
            for (int c = 0; c < kU4; ++c) {
              const uint32_t* v = (c < 4) ? &o0[c * 8] : (c < 8) ? &o1[(c - 4) * 8] : &o2[(c - 8) * 8];
              uint4 u;
              u.x = ptx::pack2t<FMT>(__uint_as_float(v[0]) * inv, __uint_as_float(v[1]) * inv);
              u.y = ptx::pack2t<FMT>(__uint_as_float(v[2]) * inv, __uint_as_float(v[3]) * inv);
              u.z = ptx::pack2t<FMT>(__uint_as_float(v[4]) * inv, __uint_as_float(v[5]) * inv);
              u.w = ptx::pack2t<FMT>(__uint_as_float(v[6]) * inv, __uint_as_float(v[7]) * inv);
              if (c < 8)
                *reinterpret_cast<uint4*>(st + OFF_Q64 + g * 16384 + row_off64(row, c)) = u;
              else
                *reinterpret_cast<uint4*>(st + OFF_Q16 + g * 4096 + row_off16(row, c - 8)) = u;
            }
          }
          ptx::fence_proxy_async_smem();   // staging rows (and the gather scratch before them) -> async proxy
          named_bar_sync4(1 + g, 128);
          if ((warp & 3) == 2 && lane == 0) {
            const int co = w.head * HD, x0 = w.wx * WS, y0 = w.wy * WS + (g ? 9 : 0);
            tma_store_4d(g ? &maps.ob64 : &maps.oa64, st + OFF_Q64 + g * 16384, co, x0, y0, w.b);
            if (kTail) tma_store_4d(g ? &maps.ob16 : &maps.oa16, st + OFF_Q16 + g * 4096, co + 64, x0, y0, w.b);
            bulk_commit4();
          }
        } else if (ok) {
          uint4* dst = reinterpret_cast<uint4*>(out + (static_cast<size_t>(w.b) * 4096 + (y * 64 + x)) * E + w.head * HD);
#pragma unroll
          for (int c = 0; c < kU4; ++c) {
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

  if (WIN4_TMA_OUT && warp >= 2 && warp <= 9 && (warp & 3) == 2 && lane == 0) bulk_wait_all4();   // output stores landed
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem, 512);
  }
}

}  // namespace

int samk_attn_window(const void* qkv, const void* bias_op, const void* rel_tab, void* out, int B, int E, int heads,
                      int fmt, cudaStream_t stream) {
  SAM_REQUIRE(fmt == 0 || fmt == 1, "attn_window: fmt must be fp16/bf16");
  SAM_REQUIRE(heads > 0 && E % heads == 0 && (E / heads == 80 || E / heads == 64),
              "attn_window: head_dim must be 80 (ViT-H) or 64 (ViT-L / ViT-B), got E=%d heads=%d", E, heads);
  const int HD = E / heads;
  SAM_REQUIRE(B > 0, "attn_window: empty batch");
  WinAttnMaps4 maps;
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
    const uint64_t ldo = static_cast<uint64_t>(E) * 2;   // bytes per output token row
    const uint64_t odims[4] = {static_cast<uint64_t>(E), 64, 64, static_cast<uint64_t>(B)};
    const uint64_t ostrides[4] = {2, ldo, 64 * ldo, 4096 * ldo};
    struct { CUtensorMap* m; uint32_t c, rows; int swz; } ospecs[4] = {
        {&maps.oa64, 64, 9, 3}, {&maps.oa16, 16, 9, 1}, {&maps.ob64, 64, 5, 3}, {&maps.ob16, 16, 5, 1}};
    for (auto& s : ospecs) {
      const uint32_t box[4] = {s.c, 14, s.rows, 1};
      int rc = samhost::encode_tmap_nd(s.m, 2, is_bf16, out, 4, odims, ostrides, box, s.swz);
      if (rc) return rc;
    }
  }
  int rc = samhost::encode_tmap_2d(&maps.r64, 2, is_bf16, rel_tab, HD, 64, HD * 2, 64, 64, 3);
  if (rc) return rc;
  rc = samhost::encode_tmap_2d(&maps.r16, 2, is_bf16, rel_tab, HD, 64, HD * 2, 16, 64, 1);
  if (rc) return rc;
  static samhost::PerDeviceOnce attr_once;
  if (attr_once.need()) {
    SAM_CHECK_CUDA(cudaFuncSetAttribute(win_attn4_kernel<0, 80>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes4));
    SAM_CHECK_CUDA(cudaFuncSetAttribute(win_attn4_kernel<1, 80>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes4));
    SAM_CHECK_CUDA(cudaFuncSetAttribute(win_attn4_kernel<0, 64>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes4));
    SAM_CHECK_CUDA(cudaFuncSetAttribute(win_attn4_kernel<1, 64>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes4));
    attr_once.done();
  }
  const float scale_log2e = 1.4426950408889634f / sqrtf(static_cast<float>(HD));
  const int num_items = B * 25 * heads;
  int grid = samhost::sm_count();
  if (grid > num_items) grid = num_items;
  const double wh = static_cast<double>(num_items);
  samhost::LaunchScope scope(samhost::KC_ATTN_WINDOW, stream, wh * (4.0 * 196 * 196 * HD + 4.0 * 196 * 14 * HD),
                             static_cast<double>(B) * 4096 * E * 2 * 4);
  typedef void (*KernelFn)(WinAttnMaps4, const uint16_t*, uint16_t*, int, int, int, float);
  const KernelFn kernel = (HD == 80) ? (fmt == 0 ? win_attn4_kernel<0, 80> : win_attn4_kernel<1, 80>)
                                     : (fmt == 0 ? win_attn4_kernel<0, 64> : win_attn4_kernel<1, 64>);
  kernel<<<grid, kThreads4, kSmemBytes4, stream>>>(maps, static_cast<const uint16_t*>(bias_op), static_cast<uint16_t*>(out),
                                                   E, heads, num_items, scale_log2e);
  SAM_CHECK_CUDA(cudaGetLastError());
  return 0;
}
