// 2-CTA (cta_group::2) tcgen05 GEMM for the big encoder linears:   C[M,N] = epilogue( A[M,K] * W[N,K]^T )
//
// A CTA pair (one cluster = the two SMs of a TPC) owns a 256 x 256 output tile.  CTA r loads A rows [r*128, +128) and
// W rows [r*128, +128) of the tile; the leader CTA issues tcgen05.mma.cta_group::2 (M = 256, N = 256, K = 16), which
// reads both halves of W from the two shared memories -- each W byte is fetched from L2 and written to smem once per
// pair -- and leaves each CTA's 128 x 256 fp32 accumulator in its own TMEM.
//
// Warp roles (320 threads):  warp 0 TMA producer | warp 1 MMA issuer (leader CTA) + TMEM alloc | warps 2..9 epilogue.
// Epilogue: tcgen05.ld -> +bias -> GELU -> round -> 128B-swizzled smem staging -> TMA bulk tensor store (coalesced,
// OOB-clipped, no LSU store traffic); the in-place fp32 residual add  x += A.W^T + b  (image_encoder.py:190-191) is a
// TMA reduce-add (cp.reduce.async.bulk.tensor ... .add), so the residual stream is never loaded by the SM.
// Two TMEM accumulator stages (2 x 256 columns) let the epilogue of tile i overlap the main loop of tile i+1.
//
// LayerNorm folding (LNF).  The encoder's norm1 / norm2 (image_encoder.py:178, :191) never run as kernels of their own:
//   * producer  (OUT_FMT = fp32, LNF): the residual GEMMs (proj, lin2) load the x tile with TMA, add, store the new
//     fp32 x with TMA, store a 16-bit copy xb = round(x) for the next GEMM and write per-row partial sums
//     (mean, sum of squared deviations of each 128-column slice) -- x is read once and never again by a LayerNorm pass;
//   * consumer  (OUT_FMT = 16-bit, LNF): qkv / lin1 multiply xb by W' = gamma o W and finish the normalisation in
//     the epilogue:  LN(x).W^T + b = rstd * (xb.W'^T - mean * colsum(W')) + (beta.W^T + b).
//
// Block-diagonal mode (GemmEpilogue::diag_*, fp32 store): A [S * 256, K] and W [S * wrows, K] hold S independent chunks
// and the tile of A's row block s multiplies the W rows of chunk s -- one launch computes the S partial products of a
// weight gradient dY^T . X whose reduction dimension was cut into S chunks (decoder_train.cu, Tape::gemm_tc_dw).  The only
// difference to the plain GEMM is one term in the producer's W row coordinate.
#include <stdio.h>
#include <stdlib.h>

#include "host_common.h"
#include "kernels.h"
#include "ptx.cuh"

namespace {

constexpr int BM2 = 128;          // rows per CTA (256 per pair)
constexpr int BN2 = 256;          // tile width
constexpr int BK2 = 64;
constexpr int kStages2 = 5;
constexpr int kStageA2 = BM2 * BK2 * 2;         // 16 KB
constexpr int kStageB2 = (BN2 / 2) * BK2 * 2;   // 16 KB (this CTA's half of W)
constexpr int kStage2 = kStageA2 + kStageB2;
constexpr int kEpiWarps = 8;
constexpr int kStagingPerWarp = 2 * 4096;       // double-buffered 32 x 128 B boxes
constexpr int kThreads2 = 64 + kEpiWarps * 32;
constexpr int kSmem2 = kStages2 * kStage2 + kEpiWarps * kStagingPerWarp + 1024 /*align*/ + 256 /*barriers*/ +
                       1024 /*half_stats*/;
static_assert(kSmem2 <= 232448, "more than the 227 KB of shared memory a CTA may opt in to");
static_assert((2 * kStages2 + 4 + 2 * kEpiWarps) * 8 + 4 <= 256, "barrier block overflows its 256 bytes");
constexpr uint32_t kTmemCols2 = 512;

// which roles use the parked mbarrier wait (ptx::mbar_wait_parked): bit 0 TMA producer, bit 1 MMA issuer, bit 2 epilogue
// warps waiting for an accumulator, bit 3 epilogue warps waiting for their x-tile loads
#ifndef GEMM2_PARK_MASK
#define GEMM2_PARK_MASK 15
#endif
template <int BIT>
__device__ __forceinline__ void wait_role(uint64_t* bar, uint32_t parity) {
  if ((GEMM2_PARK_MASK >> BIT) & 1) ptx::mbar_wait_parked(bar, parity); else ptx::mbar_wait(bar, parity);
}

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_remote(uint64_t* bar, uint32_t cta) {
  asm volatile(
      "{\n"
      ".reg .b32 ra;\n"
      "mapa.shared::cluster.u32 ra, %0, %1;\n"
      "mbarrier.arrive.shared::cluster.b64 _, [ra];\n"
      "}\n" ::"r"(ptx::smem_u32(bar)),
      "r"(cta)
      : "memory");
}
// TMA load into THIS CTA's smem; completion bytes go to the LEADER CTA's mbarrier (peer bit cleared).
__device__ __forceinline__ void tma_load_2d_cg2(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(ptx::smem_u32(dst)), "l"(m), "r"(ptx::smem_u32(bar) & 0xFEFFFFFFu), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void mma_f16_ss_cg2(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                               uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive on the same-offset mbarrier of both CTAs once all previously issued MMAs have completed
__device__ __forceinline__ void mma_commit_cg2(uint64_t* bar) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
          ptx::smem_u32(bar)),
      "h"(static_cast<uint16_t>(3))
      : "memory");
}
__device__ __forceinline__ void tmem_alloc_cg2(uint32_t* dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(ptx::smem_u32(dst_smem)),
               "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_cg2(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ uint64_t l2_policy_evict_last() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ void tma_load_2d_cg2_hint(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, uint64_t pol) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%3, %4}], [%2], %5;"
      ::"r"(ptx::smem_u32(dst)), "l"(m), "r"(ptx::smem_u32(bar) & 0xFEFFFFFFu), "r"(c0), "r"(c1), "l"(pol)
      : "memory");
}
__device__ __forceinline__ void tma_store_2d_hint(const CUtensorMap* m, const void* src, int c0, int c1, uint64_t pol) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group.L2::cache_hint [%0, {%2, %3}], [%1], %4;" ::"l"(m),
               "r"(ptx::smem_u32(src)), "r"(c0), "r"(c1), "l"(pol)
               : "memory");
}
__device__ __forceinline__ void st_global_v4_hint(void* p, uint4 v, uint64_t pol) {
  asm volatile("st.global.L2::cache_hint.v4.b32 [%0], {%1, %2, %3, %4}, %5;" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w), "l"(pol) : "memory");
}
__device__ __forceinline__ float4 ld_global_v4f_hint(const void* p, uint64_t pol) {
  float4 v;
  asm volatile("ld.global.L2::cache_hint.v4.f32 {%0, %1, %2, %3}, [%4], %5;" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p), "l"(pol));
  return v;
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* m, const void* src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(m),
               "r"(ptx::smem_u32(src)), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void tma_reduce_add_2d(const CUtensorMap* m, const void* src, int c0, int c1) {
  asm volatile("cp.reduce.async.bulk.tensor.2d.global.shared::cta.add.bulk_group [%0, {%2, %3}], [%1];" ::"l"(m),
               "r"(ptx::smem_u32(src)), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ float4 ld_shared_v4f(uint32_t saddr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(saddr));
  return v;
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}

// xb rows of a 32 x 32 chunk (64 bytes per row, one row per lane) as FULL 32-byte sectors: lanes 2k / 2k+1 swap halves so
// that instruction i of the pair covers one whole sector of one row (lane 2k: bytes [32 j, +16), lane 2k+1: the next 16)
// instead of a half sector of each of its own row -- half as many L1/L2 write requests for the same bytes.
#ifndef GEMM2_XB_PAIR
#define GEMM2_XB_PAIR 1
#endif
template <typename StoreFn>
__device__ __forceinline__ void store_xb_paired(uint16_t* xb_base /* row of lane 0 of the warp */, int ldxb, int lane,
                                                const uint4 (&u)[4], StoreFn&& store) {
  const bool odd = lane & 1;
  // the even lane sends its units 1, 3 and receives the odd lane's units 0, 2; the odd lane the other way round
  uint4 r0, r1;
  {
    const uint4 s0 = odd ? u[0] : u[1], s1 = odd ? u[2] : u[3];
    r0.x = __shfl_xor_sync(0xffffffffu, s0.x, 1); r0.y = __shfl_xor_sync(0xffffffffu, s0.y, 1);
    r0.z = __shfl_xor_sync(0xffffffffu, s0.z, 1); r0.w = __shfl_xor_sync(0xffffffffu, s0.w, 1);
    r1.x = __shfl_xor_sync(0xffffffffu, s1.x, 1); r1.y = __shfl_xor_sync(0xffffffffu, s1.y, 1);
    r1.z = __shfl_xor_sync(0xffffffffu, s1.z, 1); r1.w = __shfl_xor_sync(0xffffffffu, s1.w, 1);
  }
  // ra = the even lane's row, rb = the odd lane's row; this lane writes 16-byte unit (odd ? 1 : 0) of both sectors of both
  uint16_t* ra = xb_base + static_cast<size_t>(lane & ~1) * ldxb;
  uint16_t* rb = ra + ldxb;
  const int h = odd ? 1 : 0;
  store(reinterpret_cast<uint4*>(ra) + 0 + h, odd ? r0 : u[0]);   // ra sector 0: even u0 | even u1 (received by odd)
  store(reinterpret_cast<uint4*>(ra) + 2 + h, odd ? r1 : u[2]);   // ra sector 1: even u2 | even u3
  store(reinterpret_cast<uint4*>(rb) + 0 + h, odd ? u[1] : r0);   // rb sector 0: odd u0 (received by even) | odd u1
  store(reinterpret_cast<uint4*>(rb) + 2 + h, odd ? u[3] : r1);   // rb sector 1: odd u2 | odd u3
}

// exact-erf GELU (common.py:18) with erfc from Abramowitz-Stegun 7.1.26 (|error| < 1.5e-7, far below the 16-bit
// rounding of the stored activation).  Uses  x*Phi(x) = max(x,0) - 0.5*|x|*erfc(|x|/sqrt2)  so no sign handling is
// needed; raw MUFU ex2 / rcp (ftz) -- their arguments are always in range -- 15 instructions per element.
__device__ __forceinline__ float gelu_fast(float x) {
  const float z = fabsf(x) * 0.70710678118654752f;
  float t, e;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(fmaf(0.3275911f, z, 1.0f)));
  float p = fmaf(1.061405429f, t, -1.453152027f);
  p = fmaf(p, t, 1.421413741f);
  p = fmaf(p, t, -0.284496736f);
  p = fmaf(p, t, 0.254829592f);
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(z * z * -1.4426950408889634f));
  const float erfc_z = p * t * e;
  return fmaf(-0.70710678118654752f * z, erfc_z, fmaxf(x, 0.0f));
}

// Two elements at a time on the packed fp32 pipe (FFMA2 / FMUL2): 10 issue slots per element instead of 16 -- the
// epilogue's ALU energy is what separates the GELU GEMM from the plain one under the power cap.
__device__ __forceinline__ void gelu_fast2(float& x0, float& x1) {
  using ptx::f32x2;
  const float a0 = fabsf(x0), a1 = fabsf(x1);
  const f32x2 ax = ptx::pk2(a0, a1);
  const f32x2 d = ptx::fma2(ax, ptx::pk2(0.3275911f * 0.70710678118654752f, 0.3275911f * 0.70710678118654752f),
                            ptx::pk2(1.0f, 1.0f));
  float d0, d1, t0, t1;
  ptx::upk2(d, d0, d1);
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t0) : "f"(d0));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t1) : "f"(d1));
  const f32x2 t = ptx::pk2(t0, t1);
  f32x2 p = ptx::fma2(ptx::pk2(1.061405429f, 1.061405429f), t, ptx::pk2(-1.453152027f, -1.453152027f));
  p = ptx::fma2(p, t, ptx::pk2(1.421413741f, 1.421413741f));
  p = ptx::fma2(p, t, ptx::pk2(-0.284496736f, -0.284496736f));
  p = ptx::fma2(p, t, ptx::pk2(0.254829592f, 0.254829592f));
  // exp(-z^2) with z = |x| / sqrt2:  ex2(x^2 * (-0.5 * log2 e))
  const f32x2 w = ptx::mul2(ptx::mul2(ax, ax), ptx::pk2(-0.72134752044448170f, -0.72134752044448170f));
  float w0, w1, e0, e1;
  ptx::upk2(w, w0, w1);
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e0) : "f"(w0));
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e1) : "f"(w1));
  const f32x2 erfc_z = ptx::mul2(ptx::mul2(p, t), ptx::pk2(e0, e1));
  const f32x2 r = ptx::fma2(ptx::mul2(ax, ptx::pk2(-0.5f, -0.5f)), erfc_z, ptx::pk2(fmaxf(x0, 0.0f), fmaxf(x1, 0.0f)));
  ptx::upk2(r, x0, x1);
}

// LnSlice: statistics of one 128-column slice of a residual row as (mean, M2 = sum of squared deviations from that
// mean).  The producers accumulate sums of (x - shift) with shift = the first element of the columns they sum -- a value
// within a few standard deviations of the mean, so neither sum cancels -- and the consumer combines the slices with
// Chan's formula; the textbook E[x^2] - mean^2 loses all precision for rows whose |mean| is far above their standard deviation.
// The residual producers build a slice from its two 64-column halves (each with its own shift) and combine them with
// Chan's formula for equal halves -- in exactly this form whether one warp walks both halves (whole tile) or the two
// warps of a lane quadrant hold one half each (half item, Sched), so the statistics, and with them every downstream
// bit, do not depend on the tile schedule (test_full_size_batch_is_independent_of_batch_position).
__device__ __forceinline__ float2 ln_half(float shift, float s1, float s2) {
  const float m = s1 * (1.0f / 64.0f);
  return make_float2(shift + m, fmaxf(fmaf(-s1, m, s2), 0.f));
}
__device__ __forceinline__ float2 ln_combine(float2 a, float2 b) {
  const float d = a.x - b.x;
  return make_float2(0.5f * (a.x + b.x), fmaf(32.0f * d, d, a.y + b.y));
}

struct Gemm2Params {
  const float* bias;
  int act;         // 0 none, 1 GELU
  int out_mode;    // 0: 16-bit store, 1: fp32 store, 2: fp32 reduce-add (in-place residual)
  int out_fmt;     // SamFmt of the output
  int M, N, K;
  uint32_t idesc;
  // LayerNorm folding, consumer side: per-row slice statistics (mean, M2) [M, ln_parts], colsum(W') [N]
  const float2* ln_stats;
  int ln_parts;
  const float* ln_colsum;
  float ln_inv_c, ln_eps;
  // LayerNorm folding, producer side: 16-bit copy of the new residual rows and their partial sums [M, N / 128]
  void* xb;
  int ldxb, xb_fmt;
  float2* stats_out;
  const float* xres;   // == the fp32 output (in-place residual), row stride ldx
  int ldx;
  int x_tma;           // producer: fetch the x chunks with TMA (short main loops) instead of vector loads
  int diag_mt, diag_wrows;   // block-diagonal mode (GemmEpilogue::diag_*): W rows of a tile start at chunk * diag_wrows
  int nsplit;          // split the tiles of a final, at most half-full round into two 256 x 128 halves (see Sched)
  uint32_t idesc_half; // instruction descriptor of those half tiles (M = 256, N = 128)
  int l2hint;          // L2 eviction-priority hints: W evict_last; the producer's x / xb streams evict_first, so that
                       // they do not push the A row block out of L2 before all CTA pairs of a tile row have read it
};

// Work items of one CTA pair.  Tiles [0, lim) go round-robin over the pairs as whole 256 x 256 tiles.  When the last
// round would leave at least half of the pairs idle (0 < rem <= pairs / 2), its rem tiles are cut into 2 * rem half
// items of 256 x 128 (column slice h of the tile) and pair j takes half (tile lim + j / 2, h = j & 1) as its last item:
// the round costs about half a tile period instead of a whole one.  A half item loads 64 W rows per CTA (tmBh) and
// runs the N = 128 form of the same MMA; its 128 accumulator columns are one 128-column slice of the output, so every
// column is still accumulated in the same order (bit-identical results) and a LayerNorm slice still has one owner.
struct Sched {
  int full_items;   // whole tiles of this pair: tile = cluster_id + i * num_clusters
  int n_items;      // full_items + (0 | 1 half item)
  int half_tile, half_h;
  __host__ __device__ __forceinline__ bool is_half(int it) const { return it >= full_items; }
};
__host__ __device__ __forceinline__ Sched make_sched(int num_tiles, int num_clusters, int cluster_id, int nsplit) {
  int lim = num_tiles, n_half = 0;
  if (nsplit) {
    const int full = (num_tiles / num_clusters) * num_clusters, rem = num_tiles - full;
    if (rem > 0 && 2 * rem <= num_clusters) {
      lim = full;
      n_half = 2 * rem;
    }
  }
  Sched sc;
  sc.full_items = lim > cluster_id ? (lim - cluster_id + num_clusters - 1) / num_clusters : 0;
  sc.n_items = sc.full_items + (cluster_id < n_half ? 1 : 0);
  sc.half_tile = lim + (cluster_id >> 1);
  sc.half_h = cluster_id & 1;
  return sc;
}

// OUT_FMT: SamFmt of the output (0 fp16, 1 bf16, 2 fp32);  ACT: 0 none, 1 GELU  (compile-time so the epilogue carries
// exactly one conversion / activation path)
template <int OUT_FMT, int ACT, int LNF>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads2, 1)
gemm2_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
             const __grid_constant__ CUtensorMap tmC, const __grid_constant__ CUtensorMap tmBh, const Gemm2Params p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* staging = smem + kStages2 * kStage2;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(staging + kEpiWarps * kStagingPerWarp);
  uint64_t* empty_bar = full_bar + kStages2;
  uint64_t* acc_full = empty_bar + kStages2;
  uint64_t* acc_empty = acc_full + 2;
  uint64_t* ld_bar = acc_empty + 2;                       // LNF producer: x-tile loads, 2 per epilogue warp
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(ld_bar + 2 * kEpiWarps);
  float2* half_stats = reinterpret_cast<float2*>(staging + kEpiWarps * kStagingPerWarp + 256);   // [128] LNF producer, half items

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = (rank == 0);
  const int m_tiles = (p.M + 2 * BM2 - 1) / (2 * BM2);
  const int n_tiles = (p.N + BN2 - 1) / BN2;
  const int num_tiles = m_tiles * n_tiles;
  const int num_kb = (p.K + BK2 - 1) / BK2;
  const int cluster_id = blockIdx.x >> 1;
  const int num_clusters = gridDim.x >> 1;
  const Sched sc = make_sched(num_tiles, num_clusters, cluster_id, p.nsplit);

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tmA);
    ptx::prefetch_tmap(&tmB);
    ptx::prefetch_tmap(&tmC);
    if (p.nsplit) ptx::prefetch_tmap(&tmBh);
    for (int s = 0; s < kStages2; ++s) {
      ptx::mbar_init(&full_bar[s], 1);    // leader's arrive.expect_tx (bytes of BOTH CTAs' loads)
      ptx::mbar_init(&empty_bar[s], 1);   // multicast tcgen05.commit
    }
    for (int s = 0; s < 2; ++s) {
      ptx::mbar_init(&acc_full[s], 1);                  // multicast tcgen05.commit
      ptx::mbar_init(&acc_empty[s], 2 * kEpiWarps);     // epilogue warps of BOTH CTAs (used in the leader only)
    }
    for (int s = 0; s < 2 * kEpiWarps; ++s) ptx::mbar_init(&ld_bar[s], 1);
    ptx::fence_mbar_init();
  }
  cluster_sync();   // barriers of both CTAs initialised before anyone touches a remote one
  if (warp == 1) tmem_alloc_cg2(tmem_slot, kTmemCols2);
  ptx::tc_fence_before();
  cluster_sync();   // both TMEM allocations done before the leader's first MMA writes the peer's TMEM
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer (one lane per CTA)
    if (lane == 0) {
      int s = 0;
      uint32_t ph = 0;
      const uint64_t pol_w = l2_policy_evict_last();
      for (int it = 0; it < sc.n_items; ++it) {
        const bool hf = sc.is_half(it);
        const int tile = hf ? sc.half_tile : cluster_id + it * num_clusters;
        const int m_blk = tile / n_tiles, n_blk = tile % n_tiles;
        const int row_a = m_blk * (2 * BM2) + static_cast<int>(rank) * BM2;
        // whole tile: this CTA's 128 W rows; half item: its 64 rows of column slice h (MMA N = 128 takes 64 from each CTA)
        const int row_b = hf ? n_blk * BN2 + sc.half_h * (BN2 / 2) + static_cast<int>(rank) * (BN2 / 4)
                             : n_blk * BN2 + static_cast<int>(rank) * (BN2 / 2) + (p.diag_mt ? (m_blk / p.diag_mt) * p.diag_wrows : 0);
        const CUtensorMap* mapB = hf ? &tmBh : &tmB;
        const uint32_t tx = hf ? 2 * (kStageA2 + kStageB2 / 2) : 2 * kStage2;
        for (int kb = 0; kb < num_kb; ++kb) {
          wait_role<0>(&empty_bar[s], ph ^ 1);
          uint8_t* sa = smem + s * kStage2;
          uint8_t* sb = sa + kStageA2;
          // the peer's bytes may land before this expect_tx (tx-count is signed); they cannot land in an earlier
          // phase because the peer waited for the commit that followed the MMAs of that phase
          if (leader) ptx::mbar_expect_tx(&full_bar[s], tx);
          tma_load_2d_cg2(sa, &tmA, &full_bar[s], kb * BK2, row_a);
          if (p.l2hint) tma_load_2d_cg2_hint(sb, mapB, &full_bar[s], kb * BK2, row_b, pol_w);
          else tma_load_2d_cg2(sb, mapB, &full_bar[s], kb * BK2, row_b);
          if (++s == kStages2) { s = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer (leader CTA only)
    if (leader && lane == 0) {
      int s = 0;
      uint32_t ph = 0;
      int as = 0;
      uint32_t aph = 0;
      for (int it = 0; it < sc.n_items; ++it) {
        const uint32_t idesc = sc.is_half(it) ? p.idesc_half : p.idesc;
        wait_role<1>(&acc_empty[as], aph ^ 1);
        ptx::tc_fence_after();
        const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(as * BN2);
        for (int kb = 0; kb < num_kb; ++kb) {
          wait_role<1>(&full_bar[s], ph);
          ptx::tc_fence_after();
          const uint32_t sa = ptx::smem_u32(smem + s * kStage2);
          const uint32_t sb = sa + kStageA2;
          const uint64_t da = ptx::make_smem_desc(sa, 16, 1024, ptx::kSwz128);
          const uint64_t db = ptx::make_smem_desc(sb, 16, 1024, ptx::kSwz128);
#pragma unroll
          for (int k = 0; k < BK2 / 16; ++k) mma_f16_ss_cg2(d_tmem, da + 2 * k, db + 2 * k, idesc, (kb | k) != 0);
          mma_commit_cg2(&empty_bar[s]);
          if (++s == kStages2) { s = 0; ph ^= 1; }
        }
        mma_commit_cg2(&acc_full[as]);
        if (++as == 2) { as = 0; aph ^= 1; }
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue (8 warps per CTA)
    const int ew = warp - 2;
    const int quad = warp & 3;        // TMEM lane quadrant this warp may access
    const int half = ew >> 2;         // column half [half*128, +128) of the tile
    uint8_t* stg = staging + ew * kStagingPerWarp;
    int buf = 0;
    int as = 0;
    uint32_t aph = 0;
    if constexpr (OUT_FMT == 2 && LNF == 1) {
      // ---------------- LNF producer:  x <- x + A.W^T + b ;  xb <- round(x) ;  partial row sums of the new x.
      // The x chunk of the NEXT step is loaded with plain vector loads one chunk ahead (it does not depend on the
      // accumulator; x loads through the TMA queue would sit in front of the main loop's operand loads and stall
      // them on their HBM misses); the new x leaves through the staging buffers as TMA stores.
      // chunk q of this warp: 4 chunks of 32 columns per whole tile; a half item (always the last one) has 2 -- this
      // warp's 64 columns of the item's 128-column slice
      const int total_q = 4 * sc.full_items + (sc.n_items > sc.full_items ? 2 : 0);
      auto chunk_xy = [&](int q, int& col0, int& row0) {
        const bool hf = sc.is_half(q >> 2);
        const int tile = hf ? sc.half_tile : cluster_id + (q >> 2) * num_clusters;
        const int m_blk = tile / n_tiles, n_blk = tile % n_tiles;
        row0 = m_blk * (2 * BM2) + static_cast<int>(rank) * BM2 + quad * 32;
        col0 = n_blk * BN2 + (hf ? sc.half_h * 128 + half * 64 : half * 128) + (q & 3) * 32;
      };
      // slice statistics of the finished chunks.  Whole tile: this warp owns the 128-column slice.  Half item: the two
      // warps of a lane quadrant hold 64 columns each of the same rows; they meet at a named barrier and the
      // half == 0 warp writes the combination (Chan's formula for two equal halves).
      // called after every odd chunk with the statistics of the 64 columns just finished (chunks c - 1, c)
      float2 first_half = make_float2(0.f, 0.f);
      auto write_stats = [&](int q, int row, float2 h64) {
        if (!sc.is_half(q >> 2)) {
          if ((q & 3) == 1) {
            first_half = h64;
          } else {
            const int tile = cluster_id + (q >> 2) * num_clusters;
            p.stats_out[static_cast<size_t>(row) * (p.N >> 7) + (tile % n_tiles) * 2 + half] = ln_combine(first_half, h64);
          }
        } else {
          if (half == 1) half_stats[quad * 32 + lane] = h64;
          asm volatile("bar.sync %0, 64;" ::"r"(1 + quad) : "memory");
          if (half == 0)
            p.stats_out[static_cast<size_t>(row) * (p.N >> 7) + (sc.half_tile % n_tiles) * 2 + sc.half_h] =
                ln_combine(h64, half_stats[quad * 32 + lane]);
        }
      };
      if (p.x_tma) {
        // Short main loops (proj, K = E): the x chunks stream through the two staging buffers with TMA loads issued
        // two chunks ahead (TMA load -> add in place -> TMA store).  With a long main loop (lin2) this variant
        // loses: the x loads miss to HBM and hold up the operand loads queued behind them in the TMA unit.
        uint64_t* ldb = ld_bar + 2 * ew;
        if (lane == 0) {
          for (int q = 0; q < 2 && q < total_q; ++q) {
            int c0, r0;
            chunk_xy(q, c0, r0);
            ptx::mbar_expect_tx(&ldb[q], 4096);
            ptx::tma_load_2d(stg + q * 4096, &tmC, &ldb[q], c0, r0);
          }
        }
        uint32_t lph = 0;
        float s1 = 0.f, s2 = 0.f, shift = 0.f;
#pragma unroll 1
        for (int q = 0; q < total_q; ++q) {
          const int c = q & 3, b = q & 1;
          const bool hf = sc.is_half(q >> 2);
          const int last_c = hf ? 1 : 3;
          int col0, row0;
          chunk_xy(q, col0, row0);
          if (c == 0) {
            wait_role<2>(&acc_full[as], aph);
            ptx::tc_fence_after();
          }
          if ((c & 1) == 0) {
            s1 = 0.f;
            s2 = 0.f;
          }
          const uint32_t t_row = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) +
                                 static_cast<uint32_t>(as * BN2 + (hf ? half * 64 : half * 128));
          uint32_t v[32];
          ptx::tmem_ld_32x32b_x32(t_row + c * 32, v);
          ptx::tmem_ld_wait();
          if (c == last_c) {
            ptx::tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_remote(&acc_empty[as], 0);
          }
          wait_role<3>(&ldb[b], (lph >> b) & 1u);
          lph ^= 1u << b;
          const bool valid = row0 < p.M && col0 < p.N;     // warp-uniform (M % 32 == 0, N % 32 == 0)
          const uint32_t sb = ptx::smem_u32(stg) + b * 4096 + lane * 128;
          const int row = row0 + lane;
          uint16_t* xb_row = static_cast<uint16_t*>(p.xb) + static_cast<size_t>(row) * p.ldxb + col0;
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const uint32_t sa = sb + ((i ^ (lane & 7)) << 4);
            float4 x4 = ld_shared_v4f(sa);
            float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f);
            if (p.bias && valid) b4 = __ldg(reinterpret_cast<const float4*>(p.bias + col0) + i);
            x4.x += __uint_as_float(v[4 * i + 0]) + b4.x;
            x4.y += __uint_as_float(v[4 * i + 1]) + b4.y;
            x4.z += __uint_as_float(v[4 * i + 2]) + b4.z;
            x4.w += __uint_as_float(v[4 * i + 3]) + b4.w;
            if ((c & 1) == 0 && i == 0) shift = x4.x;      // per-half shift: sums of (x - shift) do not cancel (see LnSlice)
            {
              const float dx = x4.x - shift, dy = x4.y - shift, dz = x4.z - shift, dw = x4.w - shift;
              s1 += (dx + dy) + (dz + dw);
              s2 = fmaf(dx, dx, s2); s2 = fmaf(dy, dy, s2); s2 = fmaf(dz, dz, s2); s2 = fmaf(dw, dw, s2);
            }
            ptx::st_shared_v4f(sa, x4);
            v[4 * i + 0] = __float_as_uint(x4.x); v[4 * i + 1] = __float_as_uint(x4.y);
            v[4 * i + 2] = __float_as_uint(x4.z); v[4 * i + 3] = __float_as_uint(x4.w);
          }
          ptx::fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0 && valid) {
            tma_store_2d(&tmC, stg + b * 4096, col0, row0);
            bulk_commit();
          }
          if (valid) {
            uint4 ub[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              ub[i].x = ptx::pack2(__uint_as_float(v[8 * i + 0]), __uint_as_float(v[8 * i + 1]), p.xb_fmt);
              ub[i].y = ptx::pack2(__uint_as_float(v[8 * i + 2]), __uint_as_float(v[8 * i + 3]), p.xb_fmt);
              ub[i].z = ptx::pack2(__uint_as_float(v[8 * i + 4]), __uint_as_float(v[8 * i + 5]), p.xb_fmt);
              ub[i].w = ptx::pack2(__uint_as_float(v[8 * i + 6]), __uint_as_float(v[8 * i + 7]), p.xb_fmt);
            }
            if (GEMM2_XB_PAIR) {
              store_xb_paired(xb_row - static_cast<size_t>(lane) * p.ldxb, p.ldxb, lane, ub,
                              [](uint4* dst, const uint4& val) { *dst = val; });
            } else {
#pragma unroll
              for (int i = 0; i < 4; ++i) reinterpret_cast<uint4*>(xb_row)[i] = ub[i];
            }
            if (c & 1) write_stats(q, row, ln_half(shift, s1, s2));
          }
          if (c == last_c && ++as == 2) { as = 0; aph ^= 1; }
          if (lane == 0 && q + 2 < total_q) {
            bulk_wait_read<0>();              // the store of this chunk has read buffer b
            int c2, r2;
            chunk_xy(q + 2, c2, r2);
            ptx::mbar_expect_tx(&ldb[b], 4096);
            ptx::tma_load_2d(stg + b * 4096, &tmC, &ldb[b], c2, r2);
          }
          __syncwarp();
        }
      } else {
      const uint64_t pol_stream = l2_policy_evict_first();
      // coalesced: instruction i of lane l fetches 16 B piece (l & 7) of row 4 i + (l >> 3)  (4 full lines each)
      auto load_x = [&](int q, float4 (&dst)[8]) {
        int c0, r0;
        chunk_xy(q, c0, r0);
        if (r0 < p.M && c0 < p.N) {
          const float* src = p.xres + static_cast<size_t>(r0 + (lane >> 3)) * p.ldx + c0 + (lane & 7) * 4;
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float4* ptr = reinterpret_cast<const float4*>(src + static_cast<size_t>(4 * i) * p.ldx);
            dst[i] = p.l2hint ? ld_global_v4f_hint(ptr, pol_stream) : *ptr;
          }
        }
      };
      // two chunks of x in flight per warp (register sets xa / xb, alternating): the prefetch distance covers the
      // HBM latency under load even when the main loop is short (proj)
      float4 xa[8], xb2[8];
      load_x(0, xa);
      load_x(1, xb2);
      float s1 = 0.f, s2 = 0.f, shift = 0.f;
      auto process = [&](const int q, float4 (&xn)[8]) {
        const int c = q & 3;
        const bool hf = sc.is_half(q >> 2);
        const int last_c = hf ? 1 : 3;
        int col0, row0;
        chunk_xy(q, col0, row0);
        const bool valid = row0 < p.M && col0 < p.N;     // warp-uniform (M % 32 == 0, N % 32 == 0)
        const uint32_t sbuf = ptx::smem_u32(stg) + buf * 4096;
        if (valid) {
          // park the prefetched x chunk in the staging buffer (coalesced layout in, own row out later) so that
          // the registers are free for the next prefetch, which then has a whole chunk period to arrive
          if (lane == 0) bulk_wait_read<1>();
          __syncwarp();
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int r = 4 * i + (lane >> 3);
            ptx::st_shared_v4f(sbuf + r * 128 + (((lane & 7) ^ (r & 7)) << 4), xn[i]);
          }
        }
        if (q + 2 < total_q) load_x(q + 2, xn);
        if (c == 0) {
          wait_role<2>(&acc_full[as], aph);
          ptx::tc_fence_after();
        }
        if ((c & 1) == 0) {
          s1 = 0.f;
          s2 = 0.f;
        }
        const uint32_t t_row = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) +
                               static_cast<uint32_t>(as * BN2 + (hf ? half * 64 : half * 128));
        uint32_t v[32];
        ptx::tmem_ld_32x32b_x32(t_row + c * 32, v);
        ptx::tmem_ld_wait();
        if (c == last_c) {
          ptx::tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive_remote(&acc_empty[as], 0);
        }
        if (valid) {
          __syncwarp();
          const uint32_t sb = sbuf + lane * 128;
          float4 xr[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) xr[i] = ld_shared_v4f(sb + ((i ^ (lane & 7)) << 4));
          const int row = row0 + lane;
          uint16_t* xb_row = static_cast<uint16_t*>(p.xb) + static_cast<size_t>(row) * p.ldxb + col0;
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            float4 x4 = xr[i];
            float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f);
            if (p.bias) b4 = __ldg(reinterpret_cast<const float4*>(p.bias + col0) + i);
            x4.x += __uint_as_float(v[4 * i + 0]) + b4.x;
            x4.y += __uint_as_float(v[4 * i + 1]) + b4.y;
            x4.z += __uint_as_float(v[4 * i + 2]) + b4.z;
            x4.w += __uint_as_float(v[4 * i + 3]) + b4.w;
            if ((c & 1) == 0 && i == 0) shift = x4.x;
            {
              const float dx = x4.x - shift, dy = x4.y - shift, dz = x4.z - shift, dw = x4.w - shift;
              s1 += (dx + dy) + (dz + dw);
              s2 = fmaf(dx, dx, s2); s2 = fmaf(dy, dy, s2); s2 = fmaf(dz, dz, s2); s2 = fmaf(dw, dw, s2);
            }
            ptx::st_shared_v4f(sb + ((i ^ (lane & 7)) << 4), x4);
            xr[i] = x4;
          }
          ptx::fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) {
            if (p.l2hint) tma_store_2d_hint(&tmC, stg + buf * 4096, col0, row0, pol_stream);
            else tma_store_2d(&tmC, stg + buf * 4096, col0, row0);
            bulk_commit();
          }
          buf ^= 1;
          uint4 ub[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            ub[i].x = ptx::pack2(xr[2 * i].x, xr[2 * i].y, p.xb_fmt);
            ub[i].y = ptx::pack2(xr[2 * i].z, xr[2 * i].w, p.xb_fmt);
            ub[i].z = ptx::pack2(xr[2 * i + 1].x, xr[2 * i + 1].y, p.xb_fmt);
            ub[i].w = ptx::pack2(xr[2 * i + 1].z, xr[2 * i + 1].w, p.xb_fmt);
          }
          const bool hint = p.l2hint != 0;
          auto store = [&](uint4* dst, const uint4& val) {
            if (hint) st_global_v4_hint(dst, val, pol_stream); else *dst = val;
          };
          if (GEMM2_XB_PAIR) {
            store_xb_paired(xb_row - static_cast<size_t>(lane) * p.ldxb, p.ldxb, lane, ub, store);
          } else {
#pragma unroll
            for (int i = 0; i < 4; ++i) store(reinterpret_cast<uint4*>(xb_row) + i, ub[i]);
          }
          if (c & 1) write_stats(q, row, ln_half(shift, s1, s2));
        }
        if (c == last_c && ++as == 2) { as = 0; aph ^= 1; }
      };
#pragma unroll 1
      for (int q = 0; q < total_q; q += 2) {      // total_q is even
        process(q, xa);
        process(q + 1, xb2);
      }
      }
    } else
    for (int it = 0; it < sc.n_items; ++it) {
      const bool hf = sc.is_half(it);
      const int tile = hf ? sc.half_tile : cluster_id + it * num_clusters;
      const int m_blk = tile / n_tiles, n_blk = tile % n_tiles;
      const int row0 = m_blk * (2 * BM2) + static_cast<int>(rank) * BM2 + quad * 32;
      // first output / accumulator column of this warp: its 128-column half of a whole tile (4 chunks of 32), or its
      // 64 columns of the half item's 128-column slice (2 chunks)
      const int col_base = n_blk * BN2 + (hf ? sc.half_h * 128 + half * 64 : half * 128);
      const int ncp = hf ? 1 : 2;
      float ln_mu = 0.f, ln_r = 0.f;
      if constexpr (LNF == 1) {
        // LNF consumer: finish the row statistics of this thread's row while the main loop of the tile runs
        const int row = row0 + lane;
        if (row < p.M) {
          // combine the equal-sized slices (mean_i, M2_i): mean = avg(mean_i), M2 = sum M2_i + n_i * sum (mean_i - mean)^2
          const float2* st = p.ln_stats + static_cast<size_t>(row) * p.ln_parts;
          float sm = 0.f;
          for (int i = 0; i < p.ln_parts; ++i) sm += __ldg(st + i).x;
          ln_mu = sm / static_cast<float>(p.ln_parts);
          float m2 = 0.f, dev = 0.f;
          for (int i = 0; i < p.ln_parts; ++i) {
            const float2 t = __ldg(st + i);
            const float d = t.x - ln_mu;
            m2 += t.y;
            dev = fmaf(d, d, dev);
          }
          const float n_i = 1.0f / (p.ln_inv_c * static_cast<float>(p.ln_parts));     // columns per slice
          ln_r = rsqrtf(fmaf(n_i, dev, m2) * p.ln_inv_c + p.ln_eps);
        }
      }
      wait_role<2>(&acc_full[as], aph);
      ptx::tc_fence_after();
      const uint32_t t_row = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) +
                             static_cast<uint32_t>(as * BN2 + (hf ? half * 64 : half * 128));
      // one 32-column chunk of the accumulator, already in registers
      auto chunk = [&](const int c, const uint32_t (&v)[32]) {
        const int col0 = col_base + c * 32;
        // tile overhang (all conditions are warp-uniform); TMA clips partially out-of-range boxes itself
        const bool valid = row0 < p.M && col0 < p.N;
        const bool pair_valid = row0 < p.M && (col0 - (c & 1) * 32) < p.N;   // 16-bit mode: chunks c-1 | c share a box
        if (OUT_FMT != 2 ? !pair_valid : !valid) return;
        float f[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) f[i] = __uint_as_float(v[i]);
        if (LNF == 1) {
          // rstd * (acc - mean * colsum(W')) + (beta.W^T + b)
          if (valid) {
            const float4* b4 = reinterpret_cast<const float4*>(p.bias + col0);
            const float4* c4 = reinterpret_cast<const float4*>(p.ln_colsum + col0);
            const ptx::f32x2 nmu = ptx::pk2(-ln_mu, -ln_mu), rr = ptx::pk2(ln_r, ln_r);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              if (col0 + 4 * i < p.N) {
                const float4 b = __ldg(b4 + i);
                const float4 cs = __ldg(c4 + i);
                ptx::upk2(ptx::fma2(rr, ptx::fma2(nmu, ptx::pk2(cs.x, cs.y), ptx::pk2(f[4 * i + 0], f[4 * i + 1])), ptx::pk2(b.x, b.y)),
                          f[4 * i + 0], f[4 * i + 1]);
                ptx::upk2(ptx::fma2(rr, ptx::fma2(nmu, ptx::pk2(cs.z, cs.w), ptx::pk2(f[4 * i + 2], f[4 * i + 3])), ptx::pk2(b.z, b.w)),
                          f[4 * i + 2], f[4 * i + 3]);
              }
            }
          }
        } else if (p.bias && valid) {
          const float4* b4 = reinterpret_cast<const float4*>(p.bias + col0);
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            if (col0 + 4 * i < p.N) {
              const float4 b = __ldg(b4 + i);
              ptx::upk2(ptx::add2(ptx::pk2(f[4 * i + 0], f[4 * i + 1]), ptx::pk2(b.x, b.y)), f[4 * i + 0], f[4 * i + 1]);
              ptx::upk2(ptx::add2(ptx::pk2(f[4 * i + 2], f[4 * i + 3]), ptx::pk2(b.z, b.w)), f[4 * i + 2], f[4 * i + 3]);
            }
          }
        }
        if (ACT == 1) {
#pragma unroll
          for (int i = 0; i < 32; i += 2) gelu_fast2(f[i], f[i + 1]);
        }
        if (OUT_FMT != 2) {
          // 16-bit output: two 32-column chunks share one 32 x 64 (128 B rows) staging box
          if ((c & 1) == 0) {
            if (lane == 0) bulk_wait_read<1>();
            __syncwarp();
          }
          const uint32_t sb = ptx::smem_u32(stg) + buf * 4096 + lane * 128;
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            uint4 u;
            u.x = ptx::pack2t<OUT_FMT>(f[8 * i + 0], f[8 * i + 1]);
            u.y = ptx::pack2t<OUT_FMT>(f[8 * i + 2], f[8 * i + 3]);
            u.z = ptx::pack2t<OUT_FMT>(f[8 * i + 4], f[8 * i + 5]);
            u.w = ptx::pack2t<OUT_FMT>(f[8 * i + 6], f[8 * i + 7]);
            const int chunk = (c & 1) * 4 + i;
            ptx::st_shared_v4(sb + ((chunk ^ (lane & 7)) << 4), u);
          }
          if ((c & 1) == 1) {
            ptx::fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) {
              tma_store_2d(&tmC, stg + buf * 4096, col0 - 32, row0);
              bulk_commit();
            }
            buf ^= 1;
          }
        } else {
          if (lane == 0) bulk_wait_read<1>();
          __syncwarp();
          const uint32_t sb = ptx::smem_u32(stg) + buf * 4096 + lane * 128;
#pragma unroll
          for (int i = 0; i < 8; ++i)
            ptx::st_shared_v4f(sb + ((i ^ (lane & 7)) << 4), make_float4(f[4 * i], f[4 * i + 1], f[4 * i + 2], f[4 * i + 3]));
          ptx::fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) {
            if (p.out_mode == 2)
              tma_reduce_add_2d(&tmC, stg + buf * 4096, col0, row0);
            else
              tma_store_2d(&tmC, stg + buf * 4096, col0, row0);
            bulk_commit();
          }
          buf ^= 1;
        }
      };
      // the tensor-memory load of chunk c + 1 is in flight behind the arithmetic of chunk c (the GELU epilogue keeps
      // the epilogue warps busy ~75 % of a tile period: an exposed tcgen05.ld latency per chunk stalls the MMA issuer
      // on acc_empty)
      uint32_t va[32], vb[32];
      ptx::tmem_ld_32x32b_x32(t_row, va);
      ptx::tmem_ld_wait_dep(va);
#pragma unroll 1
      for (int cp = 0; cp < ncp; ++cp) {
        const bool more = cp + 1 < ncp;
        ptx::tmem_ld_32x32b_x32(t_row + (2 * cp + 1) * 32, vb);
        chunk(2 * cp, va);
        ptx::tmem_ld_wait_dep(vb);
        if (more) {
          ptx::tmem_ld_32x32b_x32(t_row + 64, va);
        } else {
          // accumulator stage fully read by this warp -> hand it back to the MMA issuer early
          ptx::tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive_remote(&acc_empty[as], 0);
        }
        chunk(2 * cp + 1, vb);
        if (more) ptx::tmem_ld_wait_dep(va);
      }
      if (++as == 2) { as = 0; aph ^= 1; }
    }
    if (lane == 0) bulk_wait_read<0>();   // staging smem must outlive the bulk reads
  }

  ptx::tc_fence_before();
  cluster_sync();   // nobody exits (or frees TMEM) while the peer may still touch this CTA's smem / TMEM / barriers
  if (warp == 1) {
    ptx::tc_fence_after();
    tmem_dealloc_cg2(tmem_base, kTmemCols2);
  }
}

}  // namespace

// Work items of one CTA pair as the kernel enumerates them (host copy of the same function, for the CPU tests):
// out[4 * i + {0, 1, 2, 3}] = tile, is_half, column slice h, 0.  Returns the number of items, or -1 if they exceed `cap`.
int samk_gemm2_schedule(int num_tiles, int num_clusters, int cluster_id, int nsplit, int* out, int cap) {
  const Sched sc = make_sched(num_tiles, num_clusters, cluster_id, nsplit);
  if (sc.n_items > cap) return -1;
  for (int it = 0; it < sc.n_items; ++it) {
    const bool hf = sc.is_half(it);
    out[4 * it + 0] = hf ? sc.half_tile : cluster_id + it * num_clusters;
    out[4 * it + 1] = hf ? 1 : 0;
    out[4 * it + 2] = hf ? sc.half_h : 0;
    out[4 * it + 3] = 0;
  }
  return sc.n_items;
}

// -1: policy (launches of fewer than 8 whole rounds or at most 16384 rows), 0: whole tiles only, 1: whenever the last round allows
static int g_tile_split_mode = -1;
void samk_gemm2_set_tile_split(int mode) { g_tile_split_mode = mode < 0 ? -1 : (mode ? 1 : 0); }

// Returns -1 if this kernel does not cover the request (caller falls back to the 1-CTA kernel), else 0 / error code.
int samk_gemm2(const void* A, int lda, const void* W, int ldw, int M, int N, int K, int fmt, const GemmEpilogue& ep,
               cudaStream_t stream) {
  int out_mode;
  const bool ln_consumer = ep.ln_stats != nullptr, ln_producer = ep.xb != nullptr;
  if (ep.out_fmt != SAM_F32) {
    if (ep.res || ln_producer) return -1;
    out_mode = 0;
  } else if (ln_consumer) {
    return -1;
  } else if (!ep.res) {
    if (ln_producer) return -1;
    out_mode = 1;
  } else if (ep.res == static_cast<const float*>(ep.out) && ep.res_mod == M && ep.ldr == ep.ldo) {
    out_mode = ln_producer ? 3 : 2;
  } else {
    return -1;
  }
  if (ln_consumer) {
    SAM_REQUIRE(ep.ln_colsum && ep.bias && ep.ln_parts > 0 && ep.ln_c > 0, "gemm (LN-fold consumer): colsum / bias / parts missing");
  }
  if (ln_producer) {
    SAM_REQUIRE(ep.stats_out && N % 128 == 0 && M % 32 == 0 && ep.ldxb % 8 == 0 &&
                    (reinterpret_cast<uintptr_t>(ep.xb) & 15) == 0 && ep.act == 0,
                "gemm (LN-fold producer): needs stats_out, N %% 128 == 0, M %% 32 == 0, 16-byte aligned xb rows, no activation");
  }
  if (N % 8 != 0 || M < 1) return -1;
  if (ep.diag_mt) {
    SAM_REQUIRE(out_mode == 1 && !ln_consumer && ep.act == 0 && !ep.bias && ep.diag_wrows > 0 && ep.diag_wtotal >= ep.diag_wrows &&
                    M % (2 * BM2 * ep.diag_mt) == 0 && N <= BN2,
                "gemm (block-diagonal mode): fp32 store without epilogue extras, whole row blocks per chunk, N <= 256");
  }
  CUtensorMap tmA, tmB, tmC, tmBh;
  const int is_bf16 = (fmt == 1);
  int rc = samhost::encode_tmap_2d(&tmA, 2, is_bf16, A, (uint64_t)K, (uint64_t)M, (uint64_t)lda * 2, BK2, BM2, 3);
  if (rc) return rc;
  rc = samhost::encode_tmap_2d(&tmB, 2, is_bf16, W, (uint64_t)K, (uint64_t)(ep.diag_mt ? ep.diag_wtotal : N), (uint64_t)ldw * 2, BK2,
                               BN2 / 2, 3);
  if (rc) return rc;
  // W rows of a half item (Sched): 64 per CTA
  rc = samhost::encode_tmap_2d(&tmBh, 2, is_bf16, W, (uint64_t)K, (uint64_t)(ep.diag_mt ? ep.diag_wtotal : N), (uint64_t)ldw * 2, BK2,
                               BN2 / 4, 3);
  if (rc) return rc;
  if (out_mode == 0)
    rc = samhost::encode_tmap_2d(&tmC, 2, ep.out_fmt == SAM_BF16, ep.out, (uint64_t)N, (uint64_t)M, (uint64_t)ep.ldo * 2, 64, 32, 3);
  else
    rc = samhost::encode_tmap_2d(&tmC, 4, 0, ep.out, (uint64_t)N, (uint64_t)M, (uint64_t)ep.ldo * 4, 32, 32, 3);
  if (rc) return rc;
  typedef void (*KernelFn)(CUtensorMap, CUtensorMap, CUtensorMap, CUtensorMap, Gemm2Params);
  // [out_fmt][act][LNF]; the fp32 kernels have no activation, their LNF variant is the residual producer
  static const KernelFn kernels[3][2][2] = {{{gemm2_kernel<0, 0, 0>, gemm2_kernel<0, 0, 1>}, {gemm2_kernel<0, 1, 0>, gemm2_kernel<0, 1, 1>}},
                                            {{gemm2_kernel<1, 0, 0>, gemm2_kernel<1, 0, 1>}, {gemm2_kernel<1, 1, 0>, gemm2_kernel<1, 1, 1>}},
                                            {{gemm2_kernel<2, 0, 0>, gemm2_kernel<2, 0, 1>}, {gemm2_kernel<2, 1, 0>, gemm2_kernel<2, 0, 1>}}};
  static samhost::PerDeviceOnce attr_once;
  if (attr_once.need()) {
    for (int a = 0; a < 3; ++a)
      for (int b = 0; b < 2; ++b)
        for (int c = 0; c < 2; ++c)
          SAM_CHECK_CUDA(cudaFuncSetAttribute(kernels[a][b][c], cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem2));
    attr_once.done();
  }
  if (ep.act != 0 && ep.act != 1) return -1;
  const KernelFn kernel = kernels[ep.out_fmt][ep.act][(ln_consumer || ln_producer) ? 1 : 0];
  Gemm2Params p;
  p.bias = ep.bias;
  p.act = ep.act;
  p.out_mode = out_mode;
  p.out_fmt = ep.out_fmt;
  p.M = M;
  p.N = N;
  p.K = K;
  p.idesc = ptx::make_idesc((uint32_t)fmt, 2 * BM2, BN2, 0, 0);
  p.idesc_half = ptx::make_idesc((uint32_t)fmt, 2 * BM2, BN2 / 2, 0, 0);
  p.ln_stats = static_cast<const float2*>(ep.ln_stats);
  p.ln_parts = ep.ln_parts;
  p.ln_colsum = ep.ln_colsum;
  p.ln_inv_c = ep.ln_c > 0 ? 1.0f / static_cast<float>(ep.ln_c) : 0.f;
  p.ln_eps = ep.ln_eps;
  p.xb = ep.xb;
  p.ldxb = ep.ldxb;
  p.xb_fmt = fmt;
  p.stats_out = static_cast<float2*>(ep.stats_out);
  p.diag_mt = ep.diag_mt;
  p.diag_wrows = ep.diag_wrows;
  p.xres = static_cast<const float*>(ep.out);
  p.ldx = ep.ldo;
  {
    static const char* xl = getenv("SAM_GEMM_XLOAD");   // experiment switch: "tma" | "lsu"
    p.x_tma = xl ? (xl[0] == 't') : (K <= 2048);
    static const char* lh = getenv("SAM_GEMM_L2HINT");   // "0" switches the L2 eviction hints off (A/B measurements)
    p.l2hint = lh ? (lh[0] == '1') : 1;
  }
  const int m_tiles = (M + 2 * BM2 - 1) / (2 * BM2), n_tiles = (N + BN2 - 1) / BN2;
  int clusters = m_tiles * n_tiles;
  // persistent kernel: never launch more CTA pairs than can be co-resident (GPCs with an odd number of free SMs make
  // this smaller than sm_count / 2), otherwise the late pairs serialise behind the early ones
  static int max_clusters_dev[64] = {0};   // per device
  int& max_clusters = max_clusters_dev[samhost::device_slot()];
  if (max_clusters == 0) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(samhost::sm_count());
    cfg.blockDim = dim3(kThreads2);
    cfg.dynamicSmemBytes = kSmem2;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = 2;
    at[0].val.clusterDim.y = 1;
    at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    int n = 0;
    if (cudaOccupancyMaxActiveClusters(&n, kernel, &cfg) != cudaSuccess || n <= 0) {
      (void)cudaGetLastError();
      n = samhost::sm_count() / 2;
    }
    max_clusters = n;
    if (getenv("SAM_GEMM_DEBUG")) fprintf(stderr, "[anyref_sam] gemm2: max co-resident CTA pairs = %d\n", n);
  }
  {
    // Half items in the last round (Sched) pay off for small batches: measured on the whole path (forced on / off) at
    // 1 / 2 / 4 images 7.96 -> 7.67, 13.43 -> 13.09 and 24.31 -> 24.07 ms; at 16 images the step is energy-bound under the
    // power cap, the idle pairs of the last round give their power to the busy ones, and halves -- more operand bytes
    // per FLOP -- measure equal or slower (92.9 vs 92.6 ms).  Hence: launches of fewer than 8 whole rounds, or of at
    // most 16384 rows (4 images).
    static const char* ns = getenv("SAM_GEMM_NSPLIT");   // "0": never, "1": whenever the last round allows (A/B measurements)
    const int mode = g_tile_split_mode >= 0 ? g_tile_split_mode : (ns ? (ns[0] == '1') : -1);
    const bool few_rounds = clusters < 8 * max_clusters || M <= 16384;
    p.nsplit = (mode >= 0 ? mode != 0 : few_rounds) && !ep.diag_mt;
  }
  if (clusters > max_clusters) clusters = max_clusters;
  else if (p.nsplit && 2 * clusters <= max_clusters) clusters *= 2;   // fewer tiles than half the pairs: all of them as half items
  const double out_b = (ep.out_fmt == 2) ? 4.0 : 2.0;
  samhost::LaunchScope scope(samhost::KC_GEMM, stream, 2.0 * M * N * K,
                             2.0 * (static_cast<double>(M) * K + static_cast<double>(N) * K) + out_b * M * N +
                                 (ep.res ? 4.0 * M * N : 0.0) + (ln_producer ? 2.0 * M * N : 0.0));
  kernel<<<2 * clusters, kThreads2, kSmem2, stream>>>(tmA, tmB, tmC, tmBh, p);
  SAM_CHECK_CUDA(cudaGetLastError());
  return 0;
}
