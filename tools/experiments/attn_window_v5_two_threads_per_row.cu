// 14x14 windowed attention of the SAM ViT-H encoder (replaces image_encoder.py:235-257, :263-318, :354-392):
// probabilities in TENSOR MEMORY, double-buffered operands, TWO softmax threads per query row.
//
//   * window partition / padding / un-partition are never materialised: 4-D TMA boxes over the [B, 64, 64, 3E] view
//     of the un-partitioned qkv gather a window's q / k / v, zero-fill + a bias patch reproduce the padding, the
//     output goes back through 4-D TMA stores that clip the padded rows;
//   * S = Q.K^T (N = 208) and the two rel-pos products T = Q.R^T land in tensor memory; each softmax thread packs its
//     probabilities and writes them with tcgen05.st IN PLACE over S columns it has already consumed; O = P.V is the "ts"
//     form of tcgen05.mma (A operand in tensor memory), V as five 16-wide MN-major SWIZZLE_32B chunks;
//   * the shared memory holds a 2-stage ring of Q / K / V: the producers run a whole item ahead, the two query tiles
//     free-run;
//   * round 2: a query row is shared by TWO threads (keys 0..95 | keys 96..195), i.e. 16 softmax warps per CTA instead
//     of 8.  With one thread per row every SM sub-partition ran two softmax warps whose load -> exp -> pack -> store
//     chains could not fill it (ncu r01e: issue slots 34 %, half of the issued instructions barrier polls, item =
//     5670 cycles against ~2800 of MUFU / issue work).  The two threads agree on the reference maximum through shared
//     memory, run the single-pass softmax on their own half (each with its own lazy rescale), reconcile the references
//     if one of them had to move, add their row sums, and split the O read-out (40 columns each).
// Softmax is single pass against a reference maximum (max of the first 16 keys of both halves); if a running half-row
// sum leaves [0, 2^10] the probabilities written so far are rescaled in tensor memory and the reference moved -- exact.
//
//   warp 0 / 3   : producers (Q0,Q1,K / V): 4-D TMA boxes straight from the un-partitioned qkv, padded-token patch
//   warp 1 / 2   : MMA issuers of tile 0 / 1 (one elected thread each)
//   warps 4..11  : softmax + epilogue of tile 0 (query rows 0..125);  warps 12..19: tile 1 (rows 126..195);
//                  within a tile: warps 0..3 = key half 0 of lane quadrants 0..3, warps 4..7 = key half 1
// TMEM slot g (256 columns): S [0,208) | Tw [196,228) Th [224,256) | P half 0 [0,48), half 1 [96,152) | O [152,232).
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
constexpr int kThreads4 = 640;
constexpr int kHalfKeys = 96;       // key half 0 = keys [0, 96), half 1 = keys [96, 196) (+ 12 pad keys)

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
constexpr int OFF_XCH = OFF_BAR + 256;              // float2 [2 tiles][128 rows][2 halves] exchange of the row pairs
constexpr int kSmemBytes4 = OFF_XCH + 2 * 128 * 2 * 8 + 1024;

constexpr uint32_t TM_O = 152;
constexpr uint32_t TM_P1 = 96;       // P columns of key half 1
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
// the 16 softmax warps wait for their tile's MMAs with the parked mbarrier wait (suspend-time hint): spinning, they took
// the issue slots of the MMA issuers and producers that share their SM sub-partitions (ncu r02: 67 % of the executed
// instructions were barrier polls)
#ifndef WIN5_PARK
#define WIN5_PARK 1
#endif
#if WIN5_PARK
#define WIN5_WAIT(bar, ph) ptx::mbar_wait_parked(bar, ph)
#else
#define WIN5_WAIT(bar, ph) ptx::mbar_wait(bar, ph)
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

// -DWIN5_TRACE=1: block 0 records clock64 stamps of items 3 and 4 for the Q/K producer, the tile-0 MMA issuer and the two
// tile-0 softmax threads of row 0, and prints them when the kernel ends (timeline debugging; timing builds only)
#ifndef WIN5_TRACE
#define WIN5_TRACE 0
#endif
#if WIN5_TRACE
__device__ long long g_win5_trace[2 * 4 * 12];
#define WT(role, ev)                                                                                   \
  do {                                                                                                 \
    if (blockIdx.x == 0 && n >= 3 && n < 5) g_win5_trace[((n - 3) * 4 + (role)) * 12 + (ev)] = clock64(); \
  } while (0)
#else
#define WT(role, ev) do { } while (0)
#endif

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
__device__ __forceinline__ void tmem_st_x8(uint32_t taddr, const uint32_t (&v)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(v[0]),
               "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
               : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_ld_x16(uint32_t taddr, uint32_t (&v)[16]) {
  ptx::tmem_ld_32x32b_x16(taddr, v);
}

#define WIN4_LOGIT(J, VAL) (fmaf(__uint_as_float(VAL), scale_log2e, relh[(J) / WS - KH0]) + relw[(J) % WS])

// One 16-key chunk (keys J0 .. J0+15) already in registers: probabilities against the reference folded into relh, two
// at a time on the packed fp32 pipe; the packed 16-bit pairs go straight back to tensor memory (8 columns at pcol).
// KH0 = first key row of this thread's half (relh holds key rows KH0 ..).
template <int J0, int KH0, int FMT>
__device__ __forceinline__ void exp_chunk_tm(const uint32_t (&v)[16], const float (&relh)[8], const float (&relw)[WS],
                                             f32x2 sc2, uint32_t pcol, f32x2& s0, f32x2& s1) {
  uint32_t pk[8];
#pragma unroll
  for (int i = 0; i < 16; i += 2) {
    const int j = J0 + i;   // even; NTOK and WS are even, so j and j + 1 share their key row and validity
    if (j < NTOK) {
      f32x2 x = fma2(pk2(__uint_as_float(v[i]), __uint_as_float(v[i + 1])), sc2,
                     pk2(relh[j / WS - KH0], relh[j / WS - KH0]));
      x = add2(x, pk2(relw[j % WS], relw[j % WS + 1]));
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
  tmem_st_x8(pcol, pk);
}

template <int J0, int KH0>
__device__ __forceinline__ float chunk_max(const uint32_t (&v)[16], const float (&relh)[8], const float (&relw)[WS],
                                           float scale_log2e) {
  float m0 = -INFINITY, m1 = -INFINITY;
#pragma unroll
  for (int i = 0; i < 16; i += 2) {
    const int j = J0 + i;
    if (j < NTOK) {
      m0 = fmaxf(m0, WIN4_LOGIT(j, v[i]));
      m1 = fmaxf(m1, WIN4_LOGIT(j + 1, v[i + 1]));
    }
  }
  return fmaxf(m0, m1);
}

// multiply the NP 8-column probability chunks at pbase by alpha (warp-collective tcgen05.ld / st)
template <int FMT>
__device__ __forceinline__ void rescale_p(uint32_t pbase, int np, float alpha) {
#pragma unroll 1
  for (int c = 0; c < np; ++c) {
    uint32_t p[8];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(p[0]), "=r"(p[1]), "=r"(p[2]), "=r"(p[3]), "=r"(p[4]), "=r"(p[5]), "=r"(p[6]), "=r"(p[7])
                 : "r"(pbase + c * 8)
                 : "memory");
    ptx::tmem_ld_wait();
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float2 f = ptx::unpack2(p[i], FMT);
      p[i] = ptx::pack2t<FMT>(f.x * alpha, f.y * alpha);
    }
    tmem_st_x8(pbase + c * 8, p);
  }
}

// Rare path: the running half-row sum left [0, kSumLimit] at chunk C (C previous chunks already written).  Move this
// thread's reference to the chunk's maximum, rescale its probabilities of chunks 0 .. C-1 in tensor memory and redo
// chunk C.  Warp-uniform (the tcgen05 instructions are .aligned); lanes that did not overflow use delta = 0.
template <int C, int J0, int KH0, int FMT>
__device__ __forceinline__ void rescale_and_redo(const uint32_t (&v)[16], float (&relh)[8], const float (&relw)[WS],
                                                 float scale_log2e, uint32_t pbase, f32x2& s0, f32x2& s1, f32x2 prev0,
                                                 f32x2 prev1, bool mine, float& moved) {
  tmem_st_wait();   // the probabilities written so far must have landed before they are read back
  const float cmax = chunk_max<J0, KH0>(v, relh, relw, scale_log2e);
  const float delta = mine ? fmaxf(cmax, 0.0f) : 0.0f;
  const float alpha = ex2(-delta);
  rescale_p<FMT>(pbase, C, alpha);
#pragma unroll
  for (int kh = 0; kh < 8; ++kh) relh[kh] -= delta;
  moved += delta;
  const f32x2 a2 = pk2(alpha, alpha);
  s0 = ptx::mul2(prev0, a2);
  s1 = ptx::mul2(prev1, a2);
  exp_chunk_tm<J0, KH0, FMT>(v, relh, relw, pk2(scale_log2e, scale_log2e), pbase + C * 8, s0, s1);
}

template <int C, int J0, int KH0, int FMT>
__device__ __forceinline__ void softmax_chunk(const uint32_t (&v)[16], float (&relh)[8], const float (&relw)[WS],
                                              float scale_log2e, uint32_t pbase, f32x2& s0, f32x2& s1, bool valid,
                                              float& moved) {
  const f32x2 prev0 = s0, prev1 = s1;
  exp_chunk_tm<J0, KH0, FMT>(v, relh, relw, pk2(scale_log2e, scale_log2e), pbase + C * 8, s0, s1);
  if (C > 0) {
    float a0, a1;
    upk2(add2(s0, s1), a0, a1);
    const bool over = valid && !(a0 + a1 <= kSumLimit);   // rows >= nq hold no query: never trigger
    if (__any_sync(0xffffffffu, over))
      rescale_and_redo<C, J0, KH0, FMT>(v, relh, relw, scale_log2e, pbase, s0, s1, prev0, prev1, over, moved);
  }
}

// The single-pass softmax of one key half: NCHK chunks of 16 keys starting at key JB (S columns from scol, P chunks from
// pbase), the tensor-memory load of chunk C + 1 in flight behind the arithmetic of chunk C.  `cur` holds chunk C.
template <int C, int NCHK, int JB, int KH0, int FMT>
struct HalfLoop {
  // one chunk buffer: with four softmax warps per SM sub-partition the tensor-memory load latency is hidden by the
  // other warps, and a second buffer pushed the loop over its 104-register budget (10 % of the loop were local loads)
  static __device__ __forceinline__ void run(uint32_t scol, uint32_t pbase, uint32_t (&cur)[16], float (&relh)[8],
                                             const float (&relw)[WS], float scale_log2e, bool valid, f32x2& s0, f32x2& s1,
                                             float& moved) {
    if constexpr (C < NCHK) {
      softmax_chunk<C, JB + 16 * C, KH0, FMT>(cur, relh, relw, scale_log2e, pbase, s0, s1, valid, moved);
      if constexpr (C + 1 < NCHK) {
        tmem_ld_x16(scol + (C + 1) * 16, cur);
        ptx::tmem_ld_wait_dep16(cur);
        HalfLoop<C + 1, NCHK, JB, KH0, FMT>::run(scol, pbase, cur, relh, relw, scale_log2e, valid, s0, s1, moved);
      }
    }
  }
};

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
  uint64_t* p_ready = bars + 14;   // [2 tiles] P in TMEM, count 256              (softmax -> MMA)
  uint64_t* o_full = bars + 16;    // [2 tiles] O in TMEM                         (MMA -> softmax)
  uint64_t* o_done = bars + 18;    // [2 tiles] O read out, slot free, count 256  (softmax -> MMA)
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
      ptx::mbar_init(&qk_free[s], 2 + 2);
      ptx::mbar_init(&v_full[s], 1);
      ptx::mbar_init(&v_ready[s], 1);
      ptx::mbar_init(&v_free[s], 2);
      ptx::mbar_init(&s_full[s], 1);
      ptx::mbar_init(&p_ready[s], 256);
      ptx::mbar_init(&o_full[s], 1);
      ptx::mbar_init(&o_done[s], 256);
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

  // register budget: the four service warps (one warpgroup) keep 64 registers, the 16 softmax warps take 104: the pool of
  // a CTA is what it was launched with, 128 x 64 + 512 x 104 = 640 x 96
  // (640 threads x 96 at launch; setmaxnreg is warpgroup-collective, so it sits ahead of the role branches)
  if (warp < 4) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 64;\n");
  } else {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 104;\n");
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
        WT(0, 0);
        if (n >= 2) ptx::mbar_wait(&qk_free[s], sph ^ 1);
        WT(0, 1);
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
        // (unit-major loop: one 16-byte unit of the two bias rows in registers at a time -- this warp runs on a
        // 64-register budget)
        ptx::mbar_wait(&qk_full[s], sph);
#pragma unroll 1
        for (int c = 0; c < kU4; ++c) {
          const uint4 bq = __ldg(reinterpret_cast<const uint4*>(bias_op + w.head * HD) + c);
          const uint4 bk = __ldg(reinterpret_cast<const uint4*>(bias_op + E + w.head * HD) + c);
          for (int r = lane; r < NTOK; r += 32) {
            const int iy = r / WS, ix = r % WS;
            if (y0 + iy >= 64 || x0 + ix >= 64) {
              uint8_t* q64 = st + OFF_Q64 + (r < 126 ? 0 : 16384);
              uint8_t* q16 = st + OFF_Q16 + (r < 126 ? 0 : 4096);
              const int rq = r < 126 ? r : r - 126;
              if (c < 8) {
                *reinterpret_cast<uint4*>(st + OFF_K64 + row_off64(r, c)) = bk;
                *reinterpret_cast<uint4*>(q64 + row_off64(rq, c)) = bq;
              } else {
                *reinterpret_cast<uint4*>(st + OFF_K16 + row_off16(r, c - 8)) = bk;
                *reinterpret_cast<uint4*>(q16 + row_off16(rq, c - 8)) = bq;
              }
            }
          }
        }
        ptx::fence_proxy_async_smem();
      } else {
        ptx::mbar_wait(&qk_full[s], sph);
      }
      __syncwarp();
      if (lane == 0) { WT(0, 2); ptx::mbar_arrive(&qk_ready[s]); }
    }
  } else if (warp == 3) {
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
        ptx::mbar_wait(&v_full[s], sph);
#pragma unroll 1
        for (int c = 0; c < kU4; ++c) {
          const uint4 bv = __ldg(reinterpret_cast<const uint4*>(bias_op + cv) + c);
          for (int r = lane; r < NTOK; r += 32) {
            const int iy = r / WS, ix = r % WS;
            if (y0 + iy >= 64 || x0 + ix >= 64)
              *reinterpret_cast<uint4*>(sv + (c >> 1) * kVChunk + row_off16(r, c & 1)) = bv;
          }
        }
        ptx::fence_proxy_async_smem();
      } else {
        ptx::mbar_wait(&v_full[s], sph);
      }
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&v_ready[s]);
    }
  } else if (warp == 1 || warp == 2) {
    // ============================================================ MMA issuers: warp 1 -> tile 0, warp 2 -> tile 1
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
        if (g == 0) {
          WT(1, 0);
#if WIN5_TRACE
          if (blockIdx.x == 0 && n >= 3 && n < 5) {
            unsigned long long gt;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
            g_win5_trace[((n - 3) * 4 + 1) * 12 + 9] = static_cast<long long>(gt);
          }
#endif
        }
        ptx::mbar_wait(&qk_ready[s], sph);
        if (g == 0) WT(1, 1);
        if (n > 0) ptx::mbar_wait(&o_done[g], ph ^ 1);   // slot g drained by the previous item's epilogue
        if (g == 0) WT(1, 2);
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
        if (g == 0) WT(1, 3);
        ptx::mbar_wait(&v_ready[s], sph);
        if (g == 0) WT(1, 4);
        ptx::mbar_wait(&p_ready[g], ph);
        if (g == 0) WT(1, 5);
        ptx::tc_fence_after();
        // P chunk ks (keys 16 ks ..) sits at column 8 ks for key half 0 (ks < 6) and at TM_P1 + 8 (ks - 6) for half 1
#pragma unroll
        for (int ks = 0; ks < NKEY / 16; ++ks)
          ptx::mma_f16_ts(slot + TM_O, slot + (ks < 6 ? ks * 8 : TM_P1 + (ks - 6) * 8), dv + ((ks * 512) >> 4), id_O,
                          ks != 0);
        ptx::mma_commit(&o_full[g]);
        ptx::mma_commit(&v_free[s]);
        if (g == 0) WT(1, 6);
      }
    }
  } else {
    // ============================================================ softmax warps (g = query tile, half = key half)
    const int sw = warp - 4;
    const int g = sw >> 3;
    const int half = (sw >> 2) & 1;
    const int quad = warp & 3;                         // TMEM lane quadrant == warp % 4
    const int row = (quad << 5) + lane;                // TMEM lane == query row of the tile
    const uint32_t trow = tmem + g * 256 + (static_cast<uint32_t>(quad * 32) << 16);
    const int pair_bar = 1 + g * 4 + quad;             // named barrier of the two warps that share these 32 rows
    const uint32_t xch = sbase + OFF_XCH + ((g * 128 + row) * 2) * 8;                // float2 [half] of this row
    const int nq = g ? 70 : 126;
    const int qiy = (g ? 9 : 0) + row / WS;
    const int qix = row % WS;
    const int qh = (qiy < WS) ? qiy : (WS - 1);
    const float kLog2e = 1.4426950408889634f;
    const bool store_thread = (half == 0) && (quad == 2) && (lane == 0);
    constexpr int kOH = HD / 2;                        // O columns read out per thread
    int n = 0;
    for (int it = blockIdx.x; it < num_items; it += gridDim.x, ++n) {
      const Item w = decode_item(it, heads);
      const int s = n & 1;
      const uint32_t ph = n & 1;
      if (n > 0 && store_thread) {
        // the previous item's output store has read its staging rows: that stage's Q / K may be overwritten
        bulk_wait_read_all4();
        ptx::mbar_arrive(&qk_free[s ^ 1]);
      }
      const bool tr = (g == 0) && (quad == 0) && (lane == 0);
      if (tr) WT(2 + half, 0);
      WIN5_WAIT(&s_full[g], ph);
      if (tr) WT(2 + half, 1);
      ptx::tc_fence_after();
      float relh[8], relw[WS];
      uint32_t first[16];
      const uint32_t scol = trow + (half ? kHalfKeys : 0);
      tmem_ld_x16(scol, first);            // first key chunk of this half (in flight behind the gather)
      {
        // rel-pos products of this row: Th[27] (cols 224..250), Tw[27] (cols 196..222).  The terms a row needs sit at a
        // row-dependent offset (index = q - k + 13, image_encoder.py:347-351): bounce them through scratch in this tile's
        // Q buffer (dead once S / T are in TMEM), word (j, row) at j*512 + row*4: no bank conflicts.  The two threads of
        // a row share the scratch: half 0 dumps Th, both read their key rows, half 1 dumps Tw, both read it.
        uint32_t t[32];
        ptx::tmem_ld_32x32b_x32(trow + (half ? 196 : 224), t);
        ptx::tmem_ld_wait();               // also completes `first`
        const uint32_t sc = sbase + s * kStageBytes + OFF_Q64 + g * 16384 + row * 4;
        if (half == 0) {
#pragma unroll
          for (int j = 0; j < 27; ++j)
            asm volatile("st.shared.f32 [%0], %1;" ::"r"(sc + j * 512), "f"(__uint_as_float(t[j]) * kLog2e) : "memory");
        }
        named_bar_sync4(pair_bar, 64);
        {
          // this half's key rows: kh = KH0 .. KH0 + 7 with KH0 = 0 (keys 0..95 -> rows 0..6) or 6 (keys 96..195 -> 6..13)
          const int kh0 = half ? 6 : 0;
          const uint32_t ah = sc + (qh + (WS - 1) - kh0) * 512;
#pragma unroll
          for (int kh = 0; kh < 8; ++kh)
            asm volatile("ld.shared.f32 %0, [%1];" : "=f"(relh[kh]) : "r"(ah - kh * 512) : "memory");
        }
        named_bar_sync4(pair_bar, 64);
        if (half == 1) {
#pragma unroll
          for (int j = 0; j < 27; ++j)
            asm volatile("st.shared.f32 [%0], %1;" ::"r"(sc + j * 512), "f"(__uint_as_float(t[j]) * kLog2e) : "memory");
        }
        named_bar_sync4(pair_bar, 64);
        const uint32_t aw = sc + (qix + (WS - 1)) * 512;
#pragma unroll
        for (int kw = 0; kw < WS; ++kw)
          asm volatile("ld.shared.f32 %0, [%1];" : "=f"(relw[kw]) : "r"(aw - kw * 512) : "memory");
      }
      if (tr) WT(2 + half, 2);
      // reference maximum of the row: max over the first 16 keys of both halves
      float mref = half ? chunk_max<kHalfKeys, 6>(first, relh, relw, scale_log2e)
                        : chunk_max<0, 0>(first, relh, relw, scale_log2e);
      asm volatile("st.shared.f32 [%0], %1;" ::"r"(xch + half * 8), "f"(mref) : "memory");
      named_bar_sync4(pair_bar, 64);
      {
        float om;
        asm volatile("ld.shared.f32 %0, [%1];" : "=f"(om) : "r"(xch + (half ^ 1) * 8) : "memory");
        mref = fmaxf(mref, om);
      }
#pragma unroll
      for (int kh = 0; kh < 8; ++kh) relh[kh] -= mref;
      named_bar_sync4(pair_bar, 64);       // both have read the maxima before the slots are reused for the sums

      if (tr) WT(2 + half, 3);
      f32x2 s0 = 0ull, s1 = 0ull;
      float moved = 0.f;                   // how far this thread's reference has moved up (lazy rescale)
      const bool valid = row < nq;
      const uint32_t pbase = trow + (half ? TM_P1 : 0);
      if (half)
        HalfLoop<0, 7, kHalfKeys, 6, FMT>::run(scol, pbase, first, relh, relw, scale_log2e, valid, s0, s1, moved);
      else
        HalfLoop<0, 6, 0, 0, FMT>::run(scol, pbase, first, relh, relw, scale_log2e, valid, s0, s1, moved);
      if (tr) WT(2 + half, 4);
      float a0, a1;
      upk2(add2(s0, s1), a0, a1);
      float sum = a0 + a1;
      // reconcile the two halves: common reference = the higher of the two; the other half rescales (rare)
      asm volatile("st.shared.v2.f32 [%0], {%1, %2};" ::"r"(xch + half * 8), "f"(sum), "f"(moved) : "memory");
      named_bar_sync4(pair_bar, 64);
      float2 other;
      asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(other.x), "=f"(other.y) : "r"(xch + (half ^ 1) * 8) : "memory");
      {
        const float top = fmaxf(moved, other.y);
        const float mine_a = ex2(moved - top), other_a = ex2(other.y - top);     // 1.0 unless a reference moved
        if (__any_sync(0xffffffffu, moved != top)) {
          tmem_st_wait();
          rescale_p<FMT>(pbase, half ? 7 : 6, mine_a);
        }
        sum = sum * mine_a + other.x * other_a;
      }
      tmem_st_wait();
      ptx::tc_fence_before();
      ptx::mbar_arrive(&p_ready[g]);
      if (tr) WT(2 + half, 5);

      WIN5_WAIT(&o_full[g], ph);
      if (tr) WT(2 + half, 6);
      ptx::tc_fence_after();
      {
        // this thread's half of the O row (HD / 2 fp32) into registers, then hand the TMEM slot back to the MMA issuer
        // at once (its next S overlaps the scaling and the stores below)
        uint32_t o0[32], o1[16];
        const uint32_t ocol = trow + TM_O + half * kOH;
        ptx::tmem_ld_32x32b_x32(ocol, o0);
        if (kTail) {
          asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                       : "=r"(o1[0]), "=r"(o1[1]), "=r"(o1[2]), "=r"(o1[3]), "=r"(o1[4]), "=r"(o1[5]), "=r"(o1[6]), "=r"(o1[7])
                       : "r"(ocol + 32)
                       : "memory");
        }
        ptx::tmem_ld_wait_dep(o0);
        if (kTail) ptx::tmem_ld_wait_dep16(o1);
        ptx::tc_fence_before();
        ptx::mbar_arrive(&o_done[g]);
        if (tr) WT(2 + half, 7);
        const float inv = 1.0f / sum;
        uint8_t* st = smem + s * kStageBytes;
        if (row < nq) {
          constexpr int kUH = kU4 / 2;     // 16-byte output units per thread (5 for head_dim 80, 4 for 64)
#pragma unroll
          for (int i = 0; i < kUH; ++i) {
            const uint32_t* v = (i < 4) ? &o0[i * 8] : &o1[(i - 4) * 8];
            uint4 u;
            u.x = ptx::pack2t<FMT>(__uint_as_float(v[0]) * inv, __uint_as_float(v[1]) * inv);
            u.y = ptx::pack2t<FMT>(__uint_as_float(v[2]) * inv, __uint_as_float(v[3]) * inv);
            u.z = ptx::pack2t<FMT>(__uint_as_float(v[4]) * inv, __uint_as_float(v[5]) * inv);
            u.w = ptx::pack2t<FMT>(__uint_as_float(v[6]) * inv, __uint_as_float(v[7]) * inv);
            const int c = half * kUH + i;  // unit of the output row
            if (c < 8)
              *reinterpret_cast<uint4*>(st + OFF_Q64 + g * 16384 + row_off64(row, c)) = u;
            else
              *reinterpret_cast<uint4*>(st + OFF_Q16 + g * 4096 + row_off16(row, c - 8)) = u;
          }
        }
        ptx::fence_proxy_async_smem();   // staging rows (and the gather scratch before them) -> async proxy
        named_bar_sync4(9 + g, 256);
        if (tr) WT(2 + half, 8);
        if (store_thread) {
          const int co = w.head * HD, x0 = w.wx * WS, y0 = w.wy * WS + (g ? 9 : 0);
          tma_store_4d(g ? &maps.ob64 : &maps.oa64, st + OFF_Q64 + g * 16384, co, x0, y0, w.b);
          if (kTail) tma_store_4d(g ? &maps.ob16 : &maps.oa16, st + OFF_Q16 + g * 4096, co + 64, x0, y0, w.b);
          bulk_commit4();
        }
      }
    }
    if (store_thread) bulk_wait_all4();   // output stores landed
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem, 512);
  }
#if WIN5_TRACE
  if (blockIdx.x == 0 && tid == 0) {
    const long long t0 = g_win5_trace[0];
    for (int i = 0; i < 2; ++i)
      for (int r = 0; r < 4; ++r) {
        const long long* e = &g_win5_trace[(i * 4 + r) * 12];
        printf("item %d role %d:", 3 + i, r);
        for (int k = 0; k < 9; ++k) printf(" %lld", e[k] ? e[k] - t0 : -1ll);
        if (r == 1) printf("  globaltimer_ns %lld", e[9] - g_win5_trace[1 * 12 + 9]);
        printf("\n");
      }
  }
#endif
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
