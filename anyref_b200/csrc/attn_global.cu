// Global (64x64 = 4096 token) attention of the 4 non-windowed SAM ViT-H blocks (replaces image_encoder.py:235-257 +
// :354-392 for blocks 7/15/23/31), decoupled pipeline:
//   * key blocks are ONE image row (64 keys), S is double-buffered in TMEM (2 x 64 columns per tile) and P is
//     double-buffered in shared memory, so Q.K^T of block j+1 is issued BEFORE the softmax of block j finishes and the
//     softmax warps find their next S already waiting;
//   * logits / probabilities are computed two at a time with the packed fp32 pipe (fma.rn.f32x2 / add.rn.f32x2);
//   * no running maximum: probabilities are taken against a reference maximum (true maximum of the first key row);
//     only if a block's probability SUM overflows 2^10 (rare) is the reference moved, O rescaled in TMEM and the block
//     redone.  The result is exactly softmax -- the reference cancels in O / l.
//   warp 0      : TMA -- Q tiles once, then K / V rows through 4-stage rings
//   warp 1      : one thread issues all tcgen05 MMAs:  S_g = Q_g.K^T (128x64x80),  O_g += P_g.V (128x80x64)
//   warps 2..5  : softmax of tile 0, one thread per query row (TMEM lane);  warps 6..9: tile 1
// Relative position (image_encoder.py:354-392): with Rrev[j] = rel_pos[126 - j], q.Rrev[63 - q_pos + k_pos] is the
// bias term.  Two prologue MMAs per tile compute T_w = Q.Rw_rev^T (all 127 offsets) and T_h = Q.Rh_rev[start..+80)^T;
// each thread keeps its 64 rel_w terms in registers (fp32) and parks its 64 rel_h terms in shared memory (fp16 pairs).
// TMEM (512 columns): tile g: S buffers at g*256 + [0,64) and [64,128), O at g*256 + [128,208).
#include "host_common.h"
#include "kernels.h"
#include "ptx.cuh"

namespace {

// head_dim HD is a template parameter: 80 (ViT-H: 64-wide SWIZZLE_128B tile + 16-wide SWIZZLE_32B tail per operand) or
// 64 (ViT-L / ViT-B: the 64-wide tile alone)
constexpr int G = 64;            // token grid
constexpr int BKV = 64;          // keys per block = one image row
constexpr int kThreadsG = 352;
constexpr int kNBlk = G * G / BKV;   // 64
constexpr int kStagesKV = 4;

// shared-memory map (bytes from the 1024-aligned base)
constexpr int OFF_Q64 = 0;          // 2 tiles x (128 x 128B) SWIZZLE_128B
constexpr int OFF_K64 = 32768;      // 4 stages x 8192
constexpr int OFF_V64 = 65536;      // 4 stages x 8192
constexpr int OFF_P = 98304;        // 2 tiles x 2 buffers x 16384 (128 rows x 128B, SWIZZLE_128B)
constexpr int OFF_Q16 = 163840;     // 2 x (128 x 32B) SWIZZLE_32B
constexpr int OFF_K16 = 172032;     // 4 stages x 2048
constexpr int OFF_V16 = 180224;     // 4 stages x 2048
constexpr int OFF_RELH = 188416;    // [32 pairs][256 rows] half2 : rel_h terms (x log2e) of every query row
constexpr int OFF_BAR = 221184;
constexpr int kSmemBytesG = OFF_BAR + 512 + 1024;
// prologue overlays (all consumed before the first K / V block lands)
constexpr int OFF_RW64 = OFF_V64;   // Rw_rev rows 0..127 (K-major B operand), 16 KB
constexpr int OFF_RW16 = OFF_V16;
constexpr int OFF_RH = OFF_K64;     // Rh_rev sub-table, 80 rows, un-swizzled core-matrix layout (5 x 80 x 32B)
constexpr int OFF_STAGE0 = OFF_K64; // tile 0: 128 x 127 fp32 staging of T_w (65024 B <= K64 + V64)
constexpr int OFF_STAGE1 = OFF_P;   // tile 1: same, over the P buffers
constexpr int kStageStride = 127;

constexpr uint32_t TM_O = 128;      // column offset of O inside a tile's 256-column slot
constexpr uint32_t TM_P = 208;      // GLOB_PTMEM: ring of six 8-column P quarters
constexpr float kSumLimit = 1024.0f;
// rel_w term of the logits preloaded into the S accumulator (tcgen05.st of rel_w / scale, the S MMAs accumulate on top)
// instead of one FADD2 per logit pair in the softmax warps (32 of the ~300 warp instructions per 64-key block)
#ifndef GLOB_PRELOAD
#define GLOB_PRELOAD 1
#endif
// Probabilities in tensor memory: P(j) goes to a ring of six 8-column quarter slots at columns [208, 256) of the tile's
// slot (tcgen05.st, quarter q of block j -> ring slot (4 j + q) mod 6) and P.V is the `ts` form of tcgen05.mma (A from
// TMEM) -- no STS, no swizzle arithmetic, no generic->async proxy fence in the softmax warps.  Quarters 0, 1 of block j
// reuse the slots of quarters 2, 3 of block j - 2; quarters 2, 3 reuse quarters 0, 1 of block j - 1, whose P.V is
// waited for halfway through the block (issued at the end of block j - 1, long finished by then).
#ifndef GLOB_PTMEM
#define GLOB_PTMEM 0
#endif
#ifndef SAM_GLOB3_POLY_EVERY
#define SAM_GLOB3_POLY_EVERY 0
#endif
constexpr int kPolyEvery = SAM_GLOB3_POLY_EVERY;   // 0: all exponentials on the MUFU.  Measured: every 4th pair on the
                                                   // FMA pipe 2.04 ms, every 2nd 2.13 ms, none 2.03 ms -- the softmax warps are
                                                   // issue / latency bound, not MUFU bound, so the offload does not pay   // a block's probability sum above this moves the reference maximum

struct GlobAttnMaps3 {
  CUtensorMap q64, q16;    // 2-D over qkv [B*4096, 3E]: box {64,128} SWIZZLE_128B and {16,128} SWIZZLE_32B
  CUtensorMap kv64, kv16;  // box {64,64} and {16,64}
};

__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
using ptx::add2;
using ptx::f32x2;
using ptx::fma2;
using ptx::pk2;
using ptx::upk2;

__device__ __forceinline__ uint32_t row_off64(int r, int c) { return r * 128 + ((c ^ (r & 7)) << 4); }
__device__ __forceinline__ uint32_t row_off16(int r, int c) { return r * 32 + ((c ^ ((r >> 2) & 1)) << 4); }

__device__ __forceinline__ void tmem_st_32x32b_x16(uint32_t taddr, const uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16};" ::"r"(taddr),
      "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]),
      "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
      : "memory");
}
__device__ __forceinline__ void tmem_st_32x32b_x8(uint32_t taddr, const uint32_t (&v)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(v[0]),
               "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
               : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// 64 fp32 words of this thread's TMEM lane starting at column address taddr
__device__ __forceinline__ void store_relw(uint32_t taddr, const float (&relw)[64]) {
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    uint32_t v[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __float_as_uint(relw[c * 16 + i]);
    tmem_st_32x32b_x16(taddr + c * 16, v);
  }
}

// 16 keys (kw = Q*16 .. Q*16+15) of a key row: logits in the log2 domain relative to the reference maximum (folded
// into rh), exp2, row-sum (two partial sums), P -> shared memory in operand format (two 16-byte units of the row).
template <int Q, int FMT, int FAKE = 0>
__device__ __forceinline__ void softmax_q16(const uint32_t (&v)[16], const float (&relw)[64], f32x2 rh2, f32x2 sc2,
                                            uint32_t prow /* row base ^ (swizzle << 4) */, int qs, f32x2& bsum2) {
  uint32_t pk[8];
#pragma unroll
  for (int i = 0; i < 16; i += 2) {
    f32x2 x = fma2(pk2(__uint_as_float(v[i]), __uint_as_float(v[i + 1])), sc2, rh2);
    if (!GLOB_PRELOAD) x = add2(x, pk2(relw[Q * 16 + i], relw[Q * 16 + i + 1]));
    float x0, x1, p0, p1;
    upk2(x, x0, x1);
    if (kPolyEvery > 0 && ((Q * 8 + (i >> 1)) % kPolyEvery) == kPolyEvery - 1) {
      // every kPolyEvery-th pair takes its exponentials on the FMA pipe instead of the MUFU:
      // 2^x = 2^n * 2^f, n = round(x) via the 1.5 * 2^23 trick, 2^f by a degree-4 polynomial on [-0.5, 0.5]
      // (max relative error 3.1e-6, far below the 16-bit rounding of P); the exponent is added as an integer.
      x0 = fmaxf(x0, -125.0f);
      x1 = fmaxf(x1, -125.0f);
      const f32x2 xc = pk2(x0, x1);
      const f32x2 t = add2(xc, pk2(12582912.0f, 12582912.0f));
      const f32x2 nf = add2(t, pk2(-12582912.0f, -12582912.0f));
      const f32x2 f = fma2(nf, pk2(-1.0f, -1.0f), xc);
      f32x2 q = fma2(pk2(0.00960039533674717f, 0.00960039533674717f), f, pk2(0.05591689422726631f, 0.05591689422726631f));
      q = fma2(q, f, pk2(0.24023719131946564f, 0.24023719131946564f));
      q = fma2(q, f, pk2(0.6931219696998596f, 0.6931219696998596f));
      q = fma2(q, f, pk2(1.0f, 1.0f));
      float q0, q1, t0, t1;
      upk2(q, q0, q1);
      upk2(t, t0, t1);
      p0 = __uint_as_float(__float_as_uint(q0) + (__float_as_uint(t0) << 23));
      p1 = __uint_as_float(__float_as_uint(q1) + (__float_as_uint(t1) << 23));
    } else if (FAKE) {
      p0 = x0 * 0.001f;   // diagnostic build only (GLOB_DIAG): no MUFU work in this tile
      p1 = x1 * 0.001f;
    } else {
      p0 = ex2(x0);
      p1 = ex2(x1);
    }
    bsum2 = add2(bsum2, pk2(p0, p1));
    pk[i >> 1] = ptx::pack2t<FMT>(p0, p1);
  }
  if (GLOB_PTMEM) {
    // prow = TMEM address of this lane's P ring; quarter Q of the block goes to ring slot (qs + Q) mod 6, qs = 4 j mod 6
    int slot = qs + Q;
    slot -= (slot >= 6) ? 6 : 0;
    tmem_st_32x32b_x8(prow + static_cast<uint32_t>(slot * 8), pk);
  } else {
    ptx::st_shared_v4(prow ^ ((Q * 2) << 4), make_uint4(pk[0], pk[1], pk[2], pk[3]));
    ptx::st_shared_v4(prow ^ ((Q * 2 + 1) << 4), make_uint4(pk[4], pk[5], pk[6], pk[7]));
  }
}

template <int Q>
__device__ __forceinline__ void max_q16(const uint32_t (&v)[16], const float (&relw)[64], float rh, float scale_log2e,
                                        float& m0, float& m1) {
#pragma unroll
  for (int i = 0; i < 16; i += 2) {
    m0 = fmaxf(m0, fmaf(__uint_as_float(v[i]), scale_log2e, rh) + (GLOB_PRELOAD ? 0.f : relw[Q * 16 + i]));
    m1 = fmaxf(m1, fmaf(__uint_as_float(v[i + 1]), scale_log2e, rh) + (GLOB_PRELOAD ? 0.f : relw[Q * 16 + i + 1]));
  }
}

// One pass over the 64 S columns of a key row, 16 at a time with the next tcgen05.ld in flight behind the arithmetic of
// the current quarter (only 32 S registers live).  MODE 0: probabilities (P -> smem, row sum);  MODE 1: maximum only.
template <int MODE, int FMT, int FAKE = 0>
__device__ __forceinline__ void block_pass(uint32_t ts, const float (&relw)[64], float rh, float scale_log2e,
                                           uint32_t prow, int qs, uint64_t* pv_prev, uint32_t pv_prev_parity,
                                           f32x2& bsum2, float& bmax) {
  const f32x2 sc2 = pk2(scale_log2e, scale_log2e), rh2 = pk2(rh, rh);
  float m0 = -INFINITY, m1 = -INFINITY;
  uint32_t a[16], b[16];
  ptx::tmem_ld_32x32b_x16(ts, a);
  ptx::tmem_ld_wait_dep16(a);
  ptx::tmem_ld_32x32b_x16(ts + 16, b);
  if (MODE == 0) softmax_q16<0, FMT, FAKE>(a, relw, rh2, sc2, prow, qs, bsum2); else max_q16<0>(a, relw, rh, scale_log2e, m0, m1);
  ptx::tmem_ld_wait_dep16(b);
  ptx::tmem_ld_32x32b_x16(ts + 32, a);
  if (MODE == 0) softmax_q16<1, FMT, FAKE>(b, relw, rh2, sc2, prow, qs, bsum2); else max_q16<1>(b, relw, rh, scale_log2e, m0, m1);
  ptx::tmem_ld_wait_dep16(a);
  ptx::tmem_ld_32x32b_x16(ts + 48, b);
  if (GLOB_PTMEM && MODE == 0 && pv_prev != nullptr) {
    ptx::mbar_wait(pv_prev, pv_prev_parity);   // P.V of the previous block has consumed the ring slots quarters 2, 3 reuse
    ptx::tc_fence_after();
  }
  if (MODE == 0) softmax_q16<2, FMT, FAKE>(a, relw, rh2, sc2, prow, qs, bsum2); else max_q16<2>(a, relw, rh, scale_log2e, m0, m1);
  ptx::tmem_ld_wait_dep16(b);
  if (MODE == 0) softmax_q16<3, FMT, FAKE>(b, relw, rh2, sc2, prow, qs, bsum2); else max_q16<3>(b, relw, rh, scale_log2e, m0, m1);
  if (MODE == 1) bmax = fmaxf(m0, m1);
}

template <int FMT, int HD>
__global__ void __launch_bounds__(kThreadsG, 1)
glob_attn3_kernel(const __grid_constant__ GlobAttnMaps3 maps, const uint16_t* __restrict__ rh_rev,
                  const uint16_t* __restrict__ rw_rev, uint16_t* __restrict__ out, const int E, const int heads,
                  const float scale_log2e) {
  constexpr int fmt = FMT;
  constexpr bool kTail = (HD > 64);
  constexpr int kU4 = HD / 8;     // 16-byte units per operand row
  constexpr int kKS = HD / 16;    // 16-wide K steps of a Q.K^T product / 16-column chunks of O
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + OFF_BAR);
  uint64_t* q_full = bars + 0;
  uint64_t* pro_done = bars + 2;   // all 256 softmax threads finished the prologue (count 256)
  uint64_t* k_full = bars + 3;     // [4]
  uint64_t* k_free = bars + 7;     // [4]
  uint64_t* v_full = bars + 11;    // [4]
  uint64_t* v_free = bars + 15;    // [4]
  uint64_t* s_full = bars + 19;    // [tile*2 + buf]  S in TMEM                    (MMA -> softmax)
  uint64_t* s_free = bars + 23;    // [tile*2 + buf]  S read out, count 128        (softmax -> MMA)
  uint64_t* p_ready = bars + 27;   // [tile*2 + buf]  P in smem, count 128         (softmax -> MMA)
  uint64_t* pv_done = bars + 31;   // [tile*2 + buf]  P.V finished: P buffer free  (MMA -> softmax)
  uint64_t* t_full = bars + 35;    // [2] prologue MMAs of tile g done
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 37);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  int w = blockIdx.x;
  const int qt = w % (G * G / 256);          // 256-query group: image rows 4*qt .. 4*qt+3
  w /= (G * G / 256);
  const int head = w % heads;
  const int b = w / heads;
  const int qh0 = qt * 4;
  const int th_start = ((60 - qh0) >> 3) << 3;   // first Rh_rev row held in T_h (multiple of 8, >= 0)
  const uint32_t sbase = ptx::smem_u32(smem);
  const int row0 = b * (G * G) + qt * 256;
  const int cq = head * HD, ck = E + head * HD, cv = 2 * E + head * HD;

  if (tid == 0) {
    ptx::prefetch_tmap(&maps.q64);
    ptx::prefetch_tmap(&maps.q16);
    ptx::prefetch_tmap(&maps.kv64);
    ptx::prefetch_tmap(&maps.kv16);
    ptx::mbar_init(q_full, 1);
    ptx::mbar_init(&t_full[0], 1);
    ptx::mbar_init(&t_full[1], 1);
    ptx::mbar_init(pro_done, 256);
    for (int i = 0; i < kStagesKV; ++i) {
      ptx::mbar_init(&k_full[i], 1);
      ptx::mbar_init(&k_free[i], 2);
      ptx::mbar_init(&v_full[i], 1);
      ptx::mbar_init(&v_free[i], 2);
      ptx::mbar_init(&s_full[i], 1);
      ptx::mbar_init(&s_free[i], 128);
      ptx::mbar_init(&p_ready[i], 128);
      ptx::mbar_init(&pv_done[i], 1);
    }
    ptx::fence_mbar_init();
    // Q tiles: issued right away (their buffers alias nothing)
    ptx::mbar_expect_tx(q_full, 2 * 128 * HD * 2);
    ptx::tma_load_2d(smem + OFF_Q64, &maps.q64, q_full, cq, row0);
    ptx::tma_load_2d(smem + OFF_Q64 + 16384, &maps.q64, q_full, cq, row0 + 128);
    if (kTail) {
      ptx::tma_load_2d(smem + OFF_Q16, &maps.q16, q_full, cq + 64, row0);
      ptx::tma_load_2d(smem + OFF_Q16 + 4096, &maps.q16, q_full, cq + 64, row0 + 128);
    }
  }
  if (warp == 1) {
    ptx::tmem_alloc(tmem_slot, 512);
    ptx::tmem_relinquish();
  }
  // rel-pos operand tables -> smem (generic proxy)
  for (int i = tid; i < 128 * kU4; i += kThreadsG) {
    const int r = i / kU4, c = i % kU4;
    const uint4 v = __ldg(reinterpret_cast<const uint4*>(rw_rev + r * HD) + c);
    if (c < 8)
      *reinterpret_cast<uint4*>(smem + OFF_RW64 + row_off64(r, c)) = v;
    else
      *reinterpret_cast<uint4*>(smem + OFF_RW16 + row_off16(r, c - 8)) = v;
  }
  for (int i = tid; i < 80 * kU4; i += kThreadsG) {
    const int r = i / kU4, c = i % kU4;  // local row r <-> Rh_rev row th_start + r (rows >= 128 do not exist: zero)
    uint4 v = make_uint4(0, 0, 0, 0);
    if (th_start + r < 128) v = __ldg(reinterpret_cast<const uint4*>(rh_rev + (th_start + r) * HD) + c);
    // K-major, no swizzle: per 16-wide K step a block of 80 rows x 32B; 8-row groups of 256B = [k-lo 128B][k-hi 128B]
    *reinterpret_cast<uint4*>(smem + OFF_RH + (c >> 1) * (80 * 32) + (r >> 3) * 256 + (c & 1) * 128 + (r & 7) * 16) = v;
  }
  ptx::fence_proxy_async_smem();
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  if (tmem != 0) {   // a CTA that owns all 512 columns gets base 0; the MMA issuers rely on it (uniform addresses)
    if (tid == 0) printf("glob_attn3: unexpected TMEM base %u\n", tmem);
    __trap();
  }

  if (warp == 0) {
    // ============================================================ TMA producer
    if (lane == 0) {
      ptx::mbar_wait(pro_done, 0);   // staging areas (alias K / V / P) are free again
      for (int j = 0; j < kNBlk; ++j) {
        const int s = j & (kStagesKV - 1);
        const uint32_t ph = (j >> 2) & 1;
        const int r = b * (G * G) + j * BKV;
        if (j >= kStagesKV) ptx::mbar_wait(&k_free[s], ph ^ 1);
        ptx::mbar_expect_tx(&k_full[s], BKV * HD * 2);
        ptx::tma_load_2d(smem + OFF_K64 + s * 8192, &maps.kv64, &k_full[s], ck, r);
        if (kTail) ptx::tma_load_2d(smem + OFF_K16 + s * 2048, &maps.kv16, &k_full[s], ck + 64, r);
        if (j >= kStagesKV) ptx::mbar_wait(&v_free[s], ph ^ 1);
        ptx::mbar_expect_tx(&v_full[s], BKV * HD * 2);
        ptx::tma_load_2d(smem + OFF_V64 + s * 8192, &maps.kv64, &v_full[s], cv, r);
        if (kTail) ptx::tma_load_2d(smem + OFF_V16 + s * 2048, &maps.kv16, &v_full[s], cv + 64, r);
      }
    }
  } else if (warp == 1 || warp == 10) {
    // ============================================================ MMA issuers: warp 1 -> tile 0, warp 10 -> tile 1
    // (ncu: with one issuing thread for both tiles that thread was busy ~90 % of the kernel -- ~12 SASS instructions
    // per UTCHMMA -- and the softmax warps starved; two issuers halve the per-thread MMA count.)
    if (ptx::elect_one()) {
      const int g = (warp == 1) ? 0 : 1;
      const uint32_t slot = g * 256;   // TMEM base is 0: this CTA owns all 512 columns (checked after the allocation)
      const uint32_t id_T = ptx::make_idesc((uint32_t)fmt, 128, 128, 0, 0);
      const uint32_t id_S = ptx::make_idesc((uint32_t)fmt, 128, 64, 0, 0);
      const uint32_t id_TH = ptx::make_idesc((uint32_t)fmt, 128, 80, 0, 0);
      const uint32_t id_O64 = ptx::make_idesc((uint32_t)fmt, 128, 64, 0, 1);
      const uint32_t id_O16 = ptx::make_idesc((uint32_t)fmt, 128, 16, 0, 1);
      const uint64_t dq64 = ptx::make_smem_desc(sbase + OFF_Q64 + g * 16384, 16, 1024, ptx::kSwz128);
      const uint64_t dq16 = ptx::make_smem_desc(sbase + OFF_Q16 + g * 4096, 16, 256, ptx::kSwz32);
      const uint64_t drw64 = ptx::make_smem_desc(sbase + OFF_RW64, 16, 1024, ptx::kSwz128);
      const uint64_t drw16 = ptx::make_smem_desc(sbase + OFF_RW16, 16, 256, ptx::kSwz32);
      const uint64_t drh = ptx::make_smem_desc(sbase + OFF_RH, 128, 256, ptx::kSwzNone);
      // stage-0 descriptors; stage s adds a constant to the (16-byte granular) start-address field
      const uint64_t dk64_0 = ptx::make_smem_desc(sbase + OFF_K64, 16, 1024, ptx::kSwz128);
      const uint64_t dk16_0 = ptx::make_smem_desc(sbase + OFF_K16, 16, 256, ptx::kSwz32);
      const uint64_t dv64_0 = ptx::make_smem_desc(sbase + OFF_V64, BKV * 128, 1024, ptx::kSwz128);
      const uint64_t dv16_0 = ptx::make_smem_desc(sbase + OFF_V16, BKV * 32, 256, ptx::kSwz32);
      const uint64_t dp_0 = ptx::make_smem_desc(sbase + OFF_P + g * 32768, 16, 1024, ptx::kSwz128);
      ptx::mbar_wait(q_full, 0);
      ptx::tc_fence_after();
      // prologue: T_w = Q . Rw_rev^T -> S columns [0,128) ;  T_h = Q . Rh_rev[th_start..+80)^T -> O columns
#pragma unroll
      for (int k = 0; k < kKS; ++k) {
        const uint64_t da = (k < 4) ? dq64 + 2 * k : dq16;
        const uint64_t dw = (k < 4) ? drw64 + 2 * k : drw16;
        ptx::mma_f16_ss(slot, da, dw, id_T, k != 0);
        ptx::mma_f16_ss(slot + TM_O, da, drh + ((k * 80 * 32) >> 4), id_TH, k != 0);
      }
      ptx::mma_commit(&t_full[g]);
      ptx::mbar_wait(pro_done, 0);

      // S_g(j) = Q_g . K(j)^T into buffer j & 1
      auto issue_s = [&](int j) {
        const int s = j & (kStagesKV - 1);
        const uint64_t dk64 = dk64_0 + static_cast<uint64_t>(s * (8192 >> 4));
        const uint64_t dk16 = dk16_0 + static_cast<uint64_t>(s * (2048 >> 4));
        const uint32_t d = slot + (j & 1) * 64;
#pragma unroll
        // GLOB_PRELOAD: the S buffer already holds rel_w / scale (written by the softmax warps), accumulate on top
        for (int k = 0; k < 4; ++k) ptx::mma_f16_ss(d, dq64 + 2 * k, dk64 + 2 * k, id_S, GLOB_PRELOAD ? 1 : (k != 0));
        if (kTail) ptx::mma_f16_ss(d, dq16, dk16, id_S, 1);
        ptx::mma_commit(&s_full[g * 2 + (j & 1)]);
        ptx::mma_commit(&k_free[s]);   // K(j) consumed by this tile (count 2: both issuers)
      };
      ptx::mbar_wait(&k_full[0], 0);
      ptx::tc_fence_after();
      issue_s(0);
      int qs = 0;   // GLOB_PTMEM: 4 j mod 6
#pragma unroll 1
      for (int j = 0; j < kNBlk; ++j, qs = (qs >= 2) ? qs - 2 : qs + 4) {
        const int s = j & (kStagesKV - 1);
        const int bf = j & 1;
        if (j + 1 < kNBlk) {
          // next S first: it only needs K(j+1) and the S buffer released by the softmax of block j-1
          const int jn = j + 1;
          ptx::mbar_wait(&k_full[jn & (kStagesKV - 1)], (jn >> 2) & 1);
          if (jn >= 2) ptx::mbar_wait(&s_free[g * 2 + (jn & 1)], ((jn >> 1) & 1) ^ 1);
          ptx::tc_fence_after();
          issue_s(jn);
        }
        const uint64_t dv64 = dv64_0 + static_cast<uint64_t>(s * (8192 >> 4));
        const uint64_t dv16 = dv16_0 + static_cast<uint64_t>(s * (2048 >> 4));
        const uint64_t dp = dp_0 + static_cast<uint64_t>(bf * (16384 >> 4));
        ptx::mbar_wait(&v_full[s], (j >> 2) & 1);
        ptx::mbar_wait(&p_ready[g * 2 + bf], (j >> 1) & 1);   // P_g(j) in smem
        ptx::tc_fence_after();
#pragma unroll
        for (int ks = 0; ks < BKV / 16; ++ks) {
          if (GLOB_PTMEM) {
            int ps = qs + ks;   // ring slot of quarter ks of block j
            ps -= (ps >= 6) ? 6 : 0;
            const uint32_t pa = slot + TM_P + static_cast<uint32_t>(ps * 8);
            ptx::mma_f16_ts(slot + TM_O, pa, dv64 + ((ks * 2048) >> 4), id_O64, (j | ks) != 0);
            if (kTail) ptx::mma_f16_ts(slot + TM_O + 64, pa, dv16 + ((ks * 512) >> 4), id_O16, (j | ks) != 0);
          } else {
            ptx::mma_f16_ss(slot + TM_O, dp + 2 * ks, dv64 + ((ks * 2048) >> 4), id_O64, (j | ks) != 0);
            if (kTail) ptx::mma_f16_ss(slot + TM_O + 64, dp + 2 * ks, dv16 + ((ks * 512) >> 4), id_O16, (j | ks) != 0);
          }
        }
        ptx::mma_commit(&pv_done[g * 2 + bf]);
        ptx::mma_commit(&v_free[s]);   // V(j) consumed by this tile (count 2)
      }
    }
  } else {
    // ============================================================ softmax warpgroups (g = query tile)
    const int g = (warp - 2) >> 2;
    const int row = ((warp & 3) << 5) + lane;            // TMEM lane == query row inside the tile
    const uint32_t slot = tmem + g * 256;
    const uint32_t trow = slot + (static_cast<uint32_t>((warp & 3) * 32) << 16);
    const int sw = row & 7;
    const uint32_t relh_addr = sbase + OFF_RELH + static_cast<uint32_t>(g * 128 + row) * 4u;   // [pair * 256] words
    const int qh = qh0 + g * 2 + (row >> 6);
    const int qw = row & 63;
    const float kLog2e = 1.4426950408889634f;
    float relw[64];   // rel_w[kw] * log2e  (GLOB_PRELOAD: rel_w[kw] / scale, the value preloaded into S)
    ptx::mbar_wait(&t_full[g], 0);
    ptx::tc_fence_after();
    {
      float* st = reinterpret_cast<float*>(smem + (g ? OFF_STAGE1 : OFF_STAGE0)) + row * kStageStride;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        uint32_t v[32];
        ptx::tmem_ld_32x32b_x32(trow + c * 32, v);
        ptx::tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i)
          if (c * 32 + i < 127) st[c * 32 + i] = __uint_as_float(v[i]);
      }
      const float* src = st + (63 - qw);
#pragma unroll
      for (int i = 0; i < 64; ++i) relw[i] = GLOB_PRELOAD ? src[i] * (kLog2e / scale_log2e) : src[i] * kLog2e;
      // rel_h: 64 consecutive T_h columns starting at a warp-uniform offset
      const uint32_t th_col = TM_O + static_cast<uint32_t>(63 - qh - th_start);
      uint32_t* relh_w = reinterpret_cast<uint32_t*>(smem + OFF_RELH) + g * 128 + row;
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        uint32_t v[32];
        ptx::tmem_ld_32x32b_x32(trow + th_col + c * 32, v);
        ptx::tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          __half2 h = __floats2half2_rn(__uint_as_float(v[2 * i]) * kLog2e, __uint_as_float(v[2 * i + 1]) * kLog2e);
          relh_w[(c * 16 + i) * 256] = *reinterpret_cast<uint32_t*>(&h);
        }
      }
    }
    if (GLOB_PRELOAD) {
      // both S buffers <- rel_w / scale (T_w has been read out of these columns above)
      store_relw(trow, relw);
      store_relw(trow + 64, relw);
      tmem_st_wait();
    }
    ptx::tc_fence_before();
    ptx::fence_proxy_async_smem();   // generic-proxy staging stores vs. the TMA writes of K / V into the same bytes
    ptx::mbar_arrive(pro_done);

    float m_ref = 0.f;   // reference maximum (log2 domain) all stored probabilities are relative to
    float l = 0.f;       // running row sum relative to m_ref
    int qs = 0;          // GLOB_PTMEM: ring slot of quarter 0 of block j = 4 j mod 6
#pragma unroll 1
    for (int j = 0; j < kNBlk; ++j, qs = (qs >= 2) ? qs - 2 : qs + 4) {
      const int bf = j & 1;
      const uint32_t ph = (j >> 1) & 1;
      const uint32_t ts = trow + bf * 64;
      const uint32_t prow = GLOB_PTMEM ? trow + TM_P
                                       : ((sbase + OFF_P + (g * 2 + bf) * 16384 + row * 128) ^ (sw << 4));   // 128-byte aligned row
      uint64_t* const pv_prev = (GLOB_PTMEM && j >= 1) ? &pv_done[g * 2 + (bf ^ 1)] : nullptr;
      const uint32_t pv_prev_parity = ((j - 1) >> 1) & 1;
      // explicit ld.shared: through the generic pointer this was S2UR SR_SWINHI + 64-bit address arithmetic + LD.E at
      // the head of every block (ncu: 7 % of the loop's stall samples)
      uint32_t rh_bits;
      asm volatile("ld.shared.b32 %0, [%1];" : "=r"(rh_bits) : "r"(relh_addr + static_cast<uint32_t>(j >> 1) * 1024u));
      const float2 rhp = __half22float2(*reinterpret_cast<const __half2*>(&rh_bits));
      float rh = (bf ? rhp.y : rhp.x);
      ptx::mbar_wait(&s_full[g * 2 + bf], ph);
      ptx::tc_fence_after();
      f32x2 bsum2 = 0ull;
      float bm = 0.f;
      if (j == 0) {
        block_pass<1, FMT>(ts, relw, rh, scale_log2e, prow, qs, pv_prev, pv_prev_parity, bsum2, bm);
        m_ref = bm;
      }
      if (j >= 2) {
        ptx::mbar_wait(&pv_done[g * 2 + bf], ph ^ 1);   // P.V of block j-2 finished: this P buffer is reusable
      }
      rh -= m_ref;
#ifdef GLOB_DIAG
      if (g == 1 || GLOB_DIAG == 2) block_pass<0, FMT, 1>(ts, relw, rh, scale_log2e, prow, qs, pv_prev, pv_prev_parity, bsum2, bm); else
#endif
      block_pass<0, FMT>(ts, relw, rh, scale_log2e, prow, qs, pv_prev, pv_prev_parity, bsum2, bm);
      float s0, s1;
      upk2(bsum2, s0, s1);
      float bsum = s0 + s1;
      if (__any_sync(0xffffffffu, !(bsum <= kSumLimit))) {
        // rare: some row of this warp has logits far above its reference.  Move the reference to the block maximum,
        // rescale the accumulated O row and row sum, and redo the block (S is still in its TMEM buffer).  (j == 0 never
        // gets here: its reference is its own maximum, so bsum <= 64.)
        block_pass<1, FMT>(ts, relw, rh, scale_log2e, prow, qs, pv_prev, pv_prev_parity, bsum2, bm);
        const float delta = fmaxf(bm, 0.f);
        const float alpha = ex2(-delta);
        m_ref += delta;
        l *= alpha;
        rh -= delta;
        ptx::mbar_wait(&pv_done[g * 2 + (bf ^ 1)], ((j - 1) >> 1) & 1);   // every earlier P.V has landed in O
        ptx::tc_fence_after();
#pragma unroll
        for (int c = 0; c < kKS; ++c) {
          uint32_t v[16];
          ptx::tmem_ld_32x32b_x16(trow + TM_O + c * 16, v);
          ptx::tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 16; ++i) v[i] = __float_as_uint(__uint_as_float(v[i]) * alpha);
          tmem_st_32x32b_x16(trow + TM_O + c * 16, v);
        }
        tmem_st_wait();
        bsum2 = 0ull;
        block_pass<0, FMT>(ts, relw, rh, scale_log2e, prow, qs, pv_prev, pv_prev_parity, bsum2, bm);
        upk2(bsum2, s0, s1);
        bsum = s0 + s1;
      }
      l += bsum;
      if (GLOB_PRELOAD) store_relw(ts, relw);   // this S buffer is read out: re-arm it for Q.K^T of block j + 2
      if (GLOB_PRELOAD || GLOB_PTMEM) tmem_st_wait();   // ... and the P quarters of this block
      ptx::tc_fence_before();
      ptx::mbar_arrive(&s_free[g * 2 + bf]);
      if (!GLOB_PTMEM) ptx::fence_proxy_async_smem();
      ptx::mbar_arrive(&p_ready[g * 2 + bf]);
    }
    // epilogue: O / l -> out
    ptx::mbar_wait(&pv_done[g * 2 + ((kNBlk - 1) & 1)], ((kNBlk - 1) >> 1) & 1);
    ptx::tc_fence_after();
    const float inv = 1.0f / l;
    uint16_t* dst = out + static_cast<size_t>(row0 + g * 128 + row) * E + head * HD;
#pragma unroll
    for (int c = 0; c < kKS; ++c) {
      uint32_t v[16];
      ptx::tmem_ld_32x32b_x16(trow + TM_O + c * 16, v);
      ptx::tmem_ld_wait();
      uint4 u0, u1;
      u0.x = ptx::pack2t<FMT>(__uint_as_float(v[0]) * inv, __uint_as_float(v[1]) * inv);
      u0.y = ptx::pack2t<FMT>(__uint_as_float(v[2]) * inv, __uint_as_float(v[3]) * inv);
      u0.z = ptx::pack2t<FMT>(__uint_as_float(v[4]) * inv, __uint_as_float(v[5]) * inv);
      u0.w = ptx::pack2t<FMT>(__uint_as_float(v[6]) * inv, __uint_as_float(v[7]) * inv);
      u1.x = ptx::pack2t<FMT>(__uint_as_float(v[8]) * inv, __uint_as_float(v[9]) * inv);
      u1.y = ptx::pack2t<FMT>(__uint_as_float(v[10]) * inv, __uint_as_float(v[11]) * inv);
      u1.z = ptx::pack2t<FMT>(__uint_as_float(v[12]) * inv, __uint_as_float(v[13]) * inv);
      u1.w = ptx::pack2t<FMT>(__uint_as_float(v[14]) * inv, __uint_as_float(v[15]) * inv);
      reinterpret_cast<uint4*>(dst + c * 16)[0] = u0;
      reinterpret_cast<uint4*>(dst + c * 16)[1] = u1;
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

int samk_attn_global(const void* qkv, const void* rh_rev, const void* rw_rev, void* out, int B, int E, int heads,
                      int fmt, cudaStream_t stream) {
  SAM_REQUIRE(fmt == 0 || fmt == 1, "attn_global: fmt must be fp16/bf16");
  SAM_REQUIRE(heads > 0 && E % heads == 0 && (E / heads == 80 || E / heads == 64),
              "attn_global: head_dim must be 80 (ViT-H) or 64 (ViT-L / ViT-B), got E=%d heads=%d", E, heads);
  const int HD = E / heads;
  SAM_REQUIRE(B > 0, "attn_global: empty batch");
  GlobAttnMaps3 maps;
  const int is_bf16 = (fmt == 1);
  const uint64_t rows = static_cast<uint64_t>(B) * G * G;
  int rc = samhost::encode_tmap_2d(&maps.q64, 2, is_bf16, qkv, 3ull * E, rows, 3ull * E * 2, 64, 128, 3);
  if (rc) return rc;
  rc = samhost::encode_tmap_2d(&maps.q16, 2, is_bf16, qkv, 3ull * E, rows, 3ull * E * 2, 16, 128, 1);
  if (rc) return rc;
  rc = samhost::encode_tmap_2d(&maps.kv64, 2, is_bf16, qkv, 3ull * E, rows, 3ull * E * 2, 64, BKV, 3);
  if (rc) return rc;
  rc = samhost::encode_tmap_2d(&maps.kv16, 2, is_bf16, qkv, 3ull * E, rows, 3ull * E * 2, 16, BKV, 1);
  if (rc) return rc;
  static samhost::PerDeviceOnce attr_once;
  if (attr_once.need()) {
    SAM_CHECK_CUDA(cudaFuncSetAttribute(glob_attn3_kernel<0, 80>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytesG));
    SAM_CHECK_CUDA(cudaFuncSetAttribute(glob_attn3_kernel<1, 80>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytesG));
    SAM_CHECK_CUDA(cudaFuncSetAttribute(glob_attn3_kernel<0, 64>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytesG));
    SAM_CHECK_CUDA(cudaFuncSetAttribute(glob_attn3_kernel<1, 64>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytesG));
    attr_once.done();
  }
  const float scale_log2e = 1.4426950408889634f / sqrtf(static_cast<float>(HD));
  const int grid = B * heads * (G * G / 256);
  const double bh = static_cast<double>(B) * heads;
  samhost::LaunchScope scope(samhost::KC_ATTN_GLOBAL, stream, bh * (4.0 * 4096 * 4096 * HD + 4.0 * 4096 * 64 * HD),
                             static_cast<double>(B) * 4096 * E * 2 * 4);
  typedef void (*KernelFn)(GlobAttnMaps3, const uint16_t*, const uint16_t*, uint16_t*, int, int, float);
  const KernelFn kernel = (HD == 80) ? (fmt == 0 ? glob_attn3_kernel<0, 80> : glob_attn3_kernel<1, 80>)
                                     : (fmt == 0 ? glob_attn3_kernel<0, 64> : glob_attn3_kernel<1, 64>);
  kernel<<<grid, kThreadsG, kSmemBytesG, stream>>>(maps, static_cast<const uint16_t*>(rh_rev),
                                                   static_cast<const uint16_t*>(rw_rev), static_cast<uint16_t*>(out), E,
                                                   heads, scale_log2e);
  SAM_CHECK_CUDA(cudaGetLastError());
  return 0;
}
