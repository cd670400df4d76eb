// Global (64x64 = 4096 token) attention of the 4 non-windowed SAM ViT-H blocks, warp-specialised version ("v2").
// Same contract as attn_global.cu (replaces image_encoder.py:235-257 + :354-392 for blocks 7/15/23/31).
//
// One CTA (1 per SM) = one (image, head) x 256 queries = two 128-query tiles (4 image rows) that ping-pong on the
// tensor core while sharing every K / V block:
//   warp 0      : TMA -- Q tiles once, then K / V key blocks (128 keys = 2 image rows) through 2-stage rings
//   warp 1      : one thread issues all tcgen05 MMAs:  S_g = Q_g.K^T (128x128x80),  O_g += P_g.V (128x80x128)
//   warps 2..5  : softmax of tile 0, one thread per query row (TMEM lane);  warps 6..9: tile 1
// O accumulates in TMEM across the 32 key blocks (never read back per block).  The softmax is single-pass with a
// lazily updated reference maximum: probabilities are computed against the running reference and the (rare) case of
// a block exceeding it by more than 2^8 rescales O in TMEM and redoes the block.  While one tile is in its softmax
// the tensor core works on the other tile, so MMA and exp2 overlap.
// Relative position (image_encoder.py:354-392): with Rrev[j] = rel_pos[126 - j], q.Rrev[63 - q_pos + k_pos] is the
// bias term.  Two prologue MMAs per tile compute T_w = Q.Rw_rev^T (all 127 offsets) and T_h = Q.Rh_rev[start..+80)^T;
// each thread keeps its 64 rel_w terms in registers (fp16 pairs) and parks its 64 rel_h terms in shared memory.
// TMEM (512 columns): tile g: S at g*256 + [0,128), O at g*256 + [128,208).
#include "host_common.h"
#include "kernels.h"
#include "ptx.cuh"

namespace {

constexpr int HD = 80;
constexpr int G = 64;           // token grid
constexpr int BKV = 128;        // keys per block
constexpr int kThreadsG = 320;
constexpr int kNBlk = G * G / BKV;   // 32

// shared-memory map (bytes from the 1024-aligned base)
constexpr int OFF_Q64 = 0;          // 2 tiles x (128 x 128B) SWIZZLE_128B
constexpr int OFF_K64 = 32768;      // 2 stages x 16384
constexpr int OFF_V64 = 65536;      // 2 stages x 16384
constexpr int OFF_P = 98304;        // 2 tiles x 32768 (two 64-key chunks of 128 x 128B)
constexpr int OFF_Q16 = 163840;     // 2 x (128 x 32B) SWIZZLE_32B
constexpr int OFF_K16 = 172032;     // 2 stages x 4096
constexpr int OFF_V16 = 180224;     // 2 stages x 4096
constexpr int OFF_RELH = 188416;    // [32 pairs][256 rows] half2 : rel_h terms (x log2e) of every query row
constexpr int OFF_BAR = 221184;
constexpr int kSmemBytesG = OFF_BAR + 256 + 1024;
// prologue overlays (all consumed before the first K / V block lands)
constexpr int OFF_RW64 = OFF_V64;   // Rw_rev rows 0..127 (K-major B operand), 16 KB
constexpr int OFF_RW16 = OFF_V16;
constexpr int OFF_RH = OFF_K64;     // Rh_rev sub-table, 80 rows, un-swizzled core-matrix layout (5 x 80 x 32B)
constexpr int OFF_STAGE0 = OFF_K64; // tile 0: 128 x 127 fp32 staging of T_w (65024 B <= K64 + V64)
constexpr int OFF_STAGE1 = OFF_P;   // tile 1: same, over both P buffers
constexpr int kStageStride = 127;

constexpr uint32_t TM_O = 128;      // column offset of O inside a tile's 256-column slot
constexpr float kRescaleThreshold = 8.0f;   // log2 units

struct GlobAttnMapsG {
  CUtensorMap t64, t16;  // 2-D over qkv [B*4096, 3E]: box {64,128} SWIZZLE_128B and {16,128} SWIZZLE_32B
};

__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
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
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// One 32-key chunk of a block: logits (log2 domain, relative to the reference max folded into rh), running block
// max, exp2, row-sum, P -> shared memory (operand format).  C = chunk index 0..3 (keys C*32 .. C*32+31).
template <int C, int FMT>
__device__ __forceinline__ void softmax_chunk(uint32_t trow, const uint32_t (&relw)[32], float rh, float scale_log2e,
                                              uint32_t pbase, int row, float& bmax, float& bsum) {
  uint32_t v[32];
  ptx::tmem_ld_32x32b_x32(trow + C * 32, v);
  ptx::tmem_ld_wait();
  float p[32];
#pragma unroll
  for (int i = 0; i < 32; i += 2) {
    const int kw = (C * 32 + i) & 63;
    const float2 rw = __half22float2(*reinterpret_cast<const __half2*>(&relw[kw >> 1]));
    const float x0 = fmaf(__uint_as_float(v[i]), scale_log2e, rh) + rw.x;
    const float x1 = fmaf(__uint_as_float(v[i + 1]), scale_log2e, rh) + rw.y;
    bmax = fmaxf(bmax, fmaxf(x0, x1));
    p[i] = ex2(x0);
    p[i + 1] = ex2(x1);
    bsum += p[i] + p[i + 1];
  }
#pragma unroll
  for (int g = 0; g < 4; ++g) {
    const int j0 = C * 32 + g * 8;
    uint4 u;
    u.x = ptx::pack2t<FMT>(p[g * 8 + 0], p[g * 8 + 1]);
    u.y = ptx::pack2t<FMT>(p[g * 8 + 2], p[g * 8 + 3]);
    u.z = ptx::pack2t<FMT>(p[g * 8 + 4], p[g * 8 + 5]);
    u.w = ptx::pack2t<FMT>(p[g * 8 + 6], p[g * 8 + 7]);
    ptx::st_shared_v4(pbase + (j0 >> 6) * 16384 + row_off64(row, (j0 & 63) >> 3), u);
  }
}

// max-only pass over one chunk (used for the very first block, whose maximum seeds the reference)
template <int C>
__device__ __forceinline__ void max_chunk(uint32_t trow, const uint32_t (&relw)[32], float rh, float scale_log2e,
                                          float& bmax) {
  uint32_t v[32];
  ptx::tmem_ld_32x32b_x32(trow + C * 32, v);
  ptx::tmem_ld_wait();
#pragma unroll
  for (int i = 0; i < 32; i += 2) {
    const int kw = (C * 32 + i) & 63;
    const float2 rw = __half22float2(*reinterpret_cast<const __half2*>(&relw[kw >> 1]));
    bmax = fmaxf(bmax, fmaxf(fmaf(__uint_as_float(v[i]), scale_log2e, rh) + rw.x,
                             fmaf(__uint_as_float(v[i + 1]), scale_log2e, rh) + rw.y));
  }
}

template <int FMT>
__global__ void __launch_bounds__(kThreadsG, 1)
glob_attn2_kernel(const __grid_constant__ GlobAttnMapsG maps, const uint16_t* __restrict__ rh_rev,
                  const uint16_t* __restrict__ rw_rev, uint16_t* __restrict__ out, const int E, const int heads,
                  const float scale_log2e) {
  constexpr int fmt = FMT;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + OFF_BAR);
  uint64_t* q_full = bars + 0;
  uint64_t* t_full = bars + 1;     // prologue MMAs done
  uint64_t* pro_done = bars + 2;   // all 256 softmax threads finished the prologue (count 256)
  uint64_t* k_full = bars + 3;     // [2]
  uint64_t* k_free = bars + 5;     // [2]
  uint64_t* v_full = bars + 7;     // [2]
  uint64_t* v_free = bars + 9;     // [2]
  uint64_t* s_full = bars + 11;    // [2 tiles]
  uint64_t* p_ready = bars + 13;   // [2 tiles] count 128
  uint64_t* pv_done = bars + 15;   // [2 tiles]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 17);

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
    ptx::prefetch_tmap(&maps.t64);
    ptx::prefetch_tmap(&maps.t16);
    ptx::mbar_init(q_full, 1);
    ptx::mbar_init(t_full, 1);
    ptx::mbar_init(pro_done, 256);
    for (int i = 0; i < 2; ++i) {
      ptx::mbar_init(&k_full[i], 1);
      ptx::mbar_init(&k_free[i], 1);
      ptx::mbar_init(&v_full[i], 1);
      ptx::mbar_init(&v_free[i], 1);
      ptx::mbar_init(&s_full[i], 1);
      ptx::mbar_init(&p_ready[i], 128);
      ptx::mbar_init(&pv_done[i], 1);
    }
    ptx::fence_mbar_init();
    // Q tiles: issued right away (their buffers alias nothing)
    ptx::mbar_expect_tx(q_full, 2 * 128 * HD * 2);
    ptx::tma_load_2d(smem + OFF_Q64, &maps.t64, q_full, cq, row0);
    ptx::tma_load_2d(smem + OFF_Q16, &maps.t16, q_full, cq + 64, row0);
    ptx::tma_load_2d(smem + OFF_Q64 + 16384, &maps.t64, q_full, cq, row0 + 128);
    ptx::tma_load_2d(smem + OFF_Q16 + 4096, &maps.t16, q_full, cq + 64, row0 + 128);
  }
  if (warp == 1) {
    ptx::tmem_alloc(tmem_slot, 512);
    ptx::tmem_relinquish();
  }
  // rel-pos operand tables -> smem (generic proxy)
  for (int i = tid; i < 128 * 10; i += kThreadsG) {
    const int r = i / 10, c = i % 10;
    const uint4 v = __ldg(reinterpret_cast<const uint4*>(rw_rev + r * HD) + c);
    if (c < 8)
      *reinterpret_cast<uint4*>(smem + OFF_RW64 + row_off64(r, c)) = v;
    else
      *reinterpret_cast<uint4*>(smem + OFF_RW16 + row_off16(r, c - 8)) = v;
  }
  for (int i = tid; i < 80 * 10; i += kThreadsG) {
    const int r = i / 10, c = i % 10;  // local row r <-> Rh_rev row th_start + r (rows >= 128 do not exist: zero)
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

  if (warp == 0) {
    // ============================================================ TMA producer
    if (lane == 0) {
      ptx::mbar_wait(pro_done, 0);   // staging areas (alias K / V / P) are free again
      for (int j = 0; j < kNBlk; ++j) {
        const int s = j & 1;
        const uint32_t ph = (j >> 1) & 1;
        const int r = b * (G * G) + j * BKV;
        if (j >= 2) ptx::mbar_wait(&k_free[s], ph ^ 1);
        ptx::mbar_expect_tx(&k_full[s], BKV * HD * 2);
        ptx::tma_load_2d(smem + OFF_K64 + s * 16384, &maps.t64, &k_full[s], ck, r);
        ptx::tma_load_2d(smem + OFF_K16 + s * 4096, &maps.t16, &k_full[s], ck + 64, r);
        if (j >= 2) ptx::mbar_wait(&v_free[s], ph ^ 1);
        ptx::mbar_expect_tx(&v_full[s], BKV * HD * 2);
        ptx::tma_load_2d(smem + OFF_V64 + s * 16384, &maps.t64, &v_full[s], cv, r);
        ptx::tma_load_2d(smem + OFF_V16 + s * 4096, &maps.t16, &v_full[s], cv + 64, r);
      }
    }
  } else if (warp == 1) {
    // ============================================================ MMA issuer
    if (ptx::elect_one()) {
      const uint32_t tmem = 0;   // experiment: uniform-datapath issue (TMEM base is 0 for a 512-column allocation)
      const uint32_t id_S = ptx::make_idesc((uint32_t)fmt, 128, 128, 0, 0);
      const uint32_t id_TH = ptx::make_idesc((uint32_t)fmt, 128, 80, 0, 0);
      const uint32_t id_O64 = ptx::make_idesc((uint32_t)fmt, 128, 64, 0, 1);
      const uint32_t id_O16 = ptx::make_idesc((uint32_t)fmt, 128, 16, 0, 1);
      uint64_t dq64[2], dq16[2], dp[2];
#pragma unroll
      for (int g = 0; g < 2; ++g) {
        dq64[g] = ptx::make_smem_desc(sbase + OFF_Q64 + g * 16384, 16, 1024, ptx::kSwz128);
        dq16[g] = ptx::make_smem_desc(sbase + OFF_Q16 + g * 4096, 16, 256, ptx::kSwz32);
        dp[g] = ptx::make_smem_desc(sbase + OFF_P + g * 32768, 16, 1024, ptx::kSwz128);
      }
      const uint64_t drw64 = ptx::make_smem_desc(sbase + OFF_RW64, 16, 1024, ptx::kSwz128);
      const uint64_t drw16 = ptx::make_smem_desc(sbase + OFF_RW16, 16, 256, ptx::kSwz32);
      const uint64_t drh = ptx::make_smem_desc(sbase + OFF_RH, 128, 256, ptx::kSwzNone);
      ptx::mbar_wait(q_full, 0);
      ptx::tc_fence_after();
      // prologue: T_w = Q . Rw_rev^T -> S columns ;  T_h = Q . Rh_rev[th_start..+80)^T -> O columns
#pragma unroll
      for (int g = 0; g < 2; ++g) {
        const uint32_t slot = tmem + g * 256;
#pragma unroll
        for (int k = 0; k < 5; ++k) {
          const uint64_t da = (k < 4) ? dq64[g] + 2 * k : dq16[g];
          const uint64_t dw = (k < 4) ? drw64 + 2 * k : drw16;
          ptx::mma_f16_ss(slot, da, dw, id_S, k != 0);
          ptx::mma_f16_ss(slot + TM_O, da, drh + ((k * 80 * 32) >> 4), id_TH, k != 0);
        }
      }
      ptx::mma_commit(t_full);
      ptx::mbar_wait(pro_done, 0);

      auto issue_s = [&](int g, int j) {
        const int s = j & 1;
        const uint64_t dk64 = ptx::make_smem_desc(sbase + OFF_K64 + s * 16384, 16, 1024, ptx::kSwz128);
        const uint64_t dk16 = ptx::make_smem_desc(sbase + OFF_K16 + s * 4096, 16, 256, ptx::kSwz32);
        const uint32_t slot = tmem + g * 256;
#pragma unroll
        for (int k = 0; k < 4; ++k) ptx::mma_f16_ss(slot, dq64[g] + 2 * k, dk64 + 2 * k, id_S, k != 0);
        ptx::mma_f16_ss(slot, dq16[g], dk16, id_S, 1);
        ptx::mma_commit(&s_full[g]);
      };
      // S of block 0 for both tiles
      ptx::mbar_wait(&k_full[0], 0);
      ptx::tc_fence_after();
      issue_s(0, 0);
      issue_s(1, 0);
      ptx::mma_commit(&k_free[0]);   // K(0) consumed by both tiles
#pragma unroll 1
      for (int j = 0; j < kNBlk; ++j) {
        const int s = j & 1;
        const uint32_t ph = (j >> 1) & 1;    // phase of the stage-indexed barriers
        const uint32_t pj = j & 1;           // phase of the per-block barriers
        const uint64_t dv64 = ptx::make_smem_desc(sbase + OFF_V64 + s * 16384, BKV * 128, 1024, ptx::kSwz128);
        const uint64_t dv16 = ptx::make_smem_desc(sbase + OFF_V16 + s * 4096, BKV * 32, 256, ptx::kSwz32);
#pragma unroll
        for (int g = 0; g < 2; ++g) {
          const uint32_t slot = tmem + g * 256;
          ptx::mbar_wait(&p_ready[g], pj);               // P_g(j) in smem, S_g(j) consumed
          if (g == 0) ptx::mbar_wait(&v_full[s], ph);
          ptx::tc_fence_after();
#pragma unroll
          for (int ks = 0; ks < BKV / 16; ++ks) {
            const uint64_t da = dp[g] + (((ks >> 2) * 16384 + (ks & 3) * 32) >> 4);
            ptx::mma_f16_ss(slot + TM_O, da, dv64 + ((ks * 2048) >> 4), id_O64, (j | ks) != 0);
            ptx::mma_f16_ss(slot + TM_O + 64, da, dv16 + ((ks * 512) >> 4), id_O16, (j | ks) != 0);
          }
          ptx::mma_commit(&pv_done[g]);
          if (g == 1) ptx::mma_commit(&v_free[s]);       // V(j) consumed by both tiles
          if (j + 1 < kNBlk) {
            if (g == 0) {
              ptx::mbar_wait(&k_full[s ^ 1], ((j + 1) >> 1) & 1);
              ptx::tc_fence_after();
            }
            issue_s(g, j + 1);
            if (g == 1) ptx::mma_commit(&k_free[s ^ 1]);  // K(j+1) consumed by both tiles
          }
        }
      }
    }
  } else {
    // ============================================================ softmax warpgroups (g = query tile)
    const int g = (warp - 2) >> 2;
    const int row = ((warp & 3) << 5) + lane;            // TMEM lane == query row inside the tile
    const uint32_t slot = tmem + g * 256;
    const uint32_t trow = slot + (static_cast<uint32_t>((warp & 3) * 32) << 16);
    const uint32_t pbase = sbase + OFF_P + g * 32768;
    uint32_t* relh_s = reinterpret_cast<uint32_t*>(smem + OFF_RELH) + g * 128 + row;   // [pair * 256]
    const int qh = qh0 + g * 2 + (row >> 6);
    const int qw = row & 63;
    const float kLog2e = 1.4426950408889634f;
    uint32_t relw[32];   // rel_w[kw] * log2e, kw = 0..63, fp16 pairs
    ptx::mbar_wait(t_full, 0);
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
      for (int i = 0; i < 32; ++i) {
        __half2 h = __floats2half2_rn(src[2 * i] * kLog2e, src[2 * i + 1] * kLog2e);
        relw[i] = *reinterpret_cast<uint32_t*>(&h);
      }
      // rel_h: 64 consecutive T_h columns starting at a warp-uniform offset
      const uint32_t th_col = TM_O + static_cast<uint32_t>(63 - qh - th_start);
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        uint32_t v[32];
        ptx::tmem_ld_32x32b_x32(trow + th_col + c * 32, v);
        ptx::tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          __half2 h = __floats2half2_rn(__uint_as_float(v[2 * i]) * kLog2e, __uint_as_float(v[2 * i + 1]) * kLog2e);
          relh_s[(c * 16 + i) * 256] = *reinterpret_cast<uint32_t*>(&h);
        }
      }
    }
    ptx::tc_fence_before();
    ptx::mbar_arrive(pro_done);

    float m_ref = 0.f;   // reference maximum (log2 domain) all stored probabilities are relative to
    float l = 0.f;       // running row sum relative to m_ref
#pragma unroll 1
    for (int j = 0; j < kNBlk; ++j) {
      const uint32_t pj = j & 1;
      const float2 rhp = __half22float2(*reinterpret_cast<const __half2*>(&relh_s[j * 256]));
      ptx::mbar_wait(&s_full[g], pj);
      ptx::tc_fence_after();
      if (j == 0) {
        float bm = -INFINITY;
        max_chunk<0>(trow, relw, rhp.x, scale_log2e, bm);
        max_chunk<1>(trow, relw, rhp.x, scale_log2e, bm);
        max_chunk<2>(trow, relw, rhp.y, scale_log2e, bm);
        max_chunk<3>(trow, relw, rhp.y, scale_log2e, bm);
        m_ref = bm;
      } else {
        ptx::mbar_wait(&pv_done[g], pj ^ 1);   // P.V of block j-1 finished: P buffer reusable, O stable
        ptx::tc_fence_after();
      }
      float bmax = -INFINITY, bsum = 0.f;
      float rh0 = rhp.x - m_ref, rh1 = rhp.y - m_ref;
      softmax_chunk<0, FMT>(trow, relw, rh0, scale_log2e, pbase, row, bmax, bsum);
      softmax_chunk<1, FMT>(trow, relw, rh0, scale_log2e, pbase, row, bmax, bsum);
      softmax_chunk<2, FMT>(trow, relw, rh1, scale_log2e, pbase, row, bmax, bsum);
      softmax_chunk<3, FMT>(trow, relw, rh1, scale_log2e, pbase, row, bmax, bsum);
      if (__any_sync(0xffffffffu, bmax > kRescaleThreshold)) {
        // rare: this block exceeds the reference by more than 2^8 for some row of the warp.  Move the reference,
        // rescale the accumulated O row and row sum, and redo the block against the new reference.
        const float delta = fmaxf(bmax, 0.f);
        const float alpha = ex2(-delta);
        m_ref += delta;
        l *= alpha;
#pragma unroll
        for (int c = 0; c < 5; ++c) {
          uint32_t v[16];
          ptx::tmem_ld_32x32b_x16(trow + TM_O + c * 16, v);
          ptx::tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 16; ++i) v[i] = __float_as_uint(__uint_as_float(v[i]) * alpha);
          tmem_st_32x32b_x16(trow + TM_O + c * 16, v);
        }
        tmem_st_wait();
        bmax = -INFINITY;
        bsum = 0.f;
        rh0 = rhp.x - m_ref;
        rh1 = rhp.y - m_ref;
        softmax_chunk<0, FMT>(trow, relw, rh0, scale_log2e, pbase, row, bmax, bsum);
        softmax_chunk<1, FMT>(trow, relw, rh0, scale_log2e, pbase, row, bmax, bsum);
        softmax_chunk<2, FMT>(trow, relw, rh1, scale_log2e, pbase, row, bmax, bsum);
        softmax_chunk<3, FMT>(trow, relw, rh1, scale_log2e, pbase, row, bmax, bsum);
      }
      l += bsum;
      ptx::fence_proxy_async_smem();
      ptx::tc_fence_before();
      ptx::mbar_arrive(&p_ready[g]);
    }
    // epilogue: O / l -> out
    ptx::mbar_wait(&pv_done[g], (kNBlk - 1) & 1);
    ptx::tc_fence_after();
    const float inv = 1.0f / l;
    uint16_t* dst = out + static_cast<size_t>(row0 + g * 128 + row) * E + head * HD;
#pragma unroll
    for (int c = 0; c < 5; ++c) {
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

int samk_attn_global2(const void* qkv, const void* rh_rev, const void* rw_rev, void* out, int B, int E, int heads,
                      int fmt, cudaStream_t stream) {
  SAM_REQUIRE(fmt == 0 || fmt == 1, "attn_global: fmt must be fp16/bf16");
  SAM_REQUIRE(E == heads * HD, "attn_global: head_dim must be 80 (E=%d heads=%d)", E, heads);
  SAM_REQUIRE(B > 0, "attn_global: empty batch");
  GlobAttnMapsG maps;
  const int is_bf16 = (fmt == 1);
  const uint64_t rows = static_cast<uint64_t>(B) * G * G;
  int rc = samhost::encode_tmap_2d(&maps.t64, 2, is_bf16, qkv, 3ull * E, rows, 3ull * E * 2, 64, 128, 3);
  if (rc) return rc;
  rc = samhost::encode_tmap_2d(&maps.t16, 2, is_bf16, qkv, 3ull * E, rows, 3ull * E * 2, 16, 128, 1);
  if (rc) return rc;
  static bool attr_done = false;
  if (!attr_done) {
    SAM_CHECK_CUDA(cudaFuncSetAttribute(glob_attn2_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytesG));
    SAM_CHECK_CUDA(cudaFuncSetAttribute(glob_attn2_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytesG));
    attr_done = true;
  }
  const float scale_log2e = 1.4426950408889634f / sqrtf(static_cast<float>(HD));
  const int grid = B * heads * (G * G / 256);
  const double bh = static_cast<double>(B) * heads;
  samhost::LaunchScope scope(samhost::KC_ATTN_GLOBAL, stream, bh * (4.0 * 4096 * 4096 * 80 + 4.0 * 4096 * 64 * 80),
                             static_cast<double>(B) * 4096 * E * 2 * 4);
  if (fmt == 0)
    glob_attn2_kernel<0><<<grid, kThreadsG, kSmemBytesG, stream>>>(maps, static_cast<const uint16_t*>(rh_rev),
                                                                    static_cast<const uint16_t*>(rw_rev),
                                                                    static_cast<uint16_t*>(out), E, heads, scale_log2e);
  else
    glob_attn2_kernel<1><<<grid, kThreadsG, kSmemBytesG, stream>>>(maps, static_cast<const uint16_t*>(rh_rev),
                                                                    static_cast<const uint16_t*>(rw_rev),
                                                                    static_cast<uint16_t*>(out), E, heads, scale_log2e);
  SAM_CHECK_CUDA(cudaGetLastError());
  return 0;
}
