// Global (64x64 = 4096 token) attention of the 4 non-windowed SAM ViT-H blocks with the decomposed relative-position
// bias fused into an online softmax -- replaces image_encoder.py:235-257 + :354-392 for blocks 7/15/23/31.  The
// [16, 4096, 4096] attention matrix (537 MB per image in the reference) is never materialised.
//
// Input  qkv [B*4096, 3E] operand format, columns (which, head, d);   output out [B*4096, E] operand format.
//
// One CTA = one (image, head, 128-query tile = two image rows); 160 threads:
//   warps 0..3 : softmax -- one thread per query row (TMEM lane); fp32 running max / sum, O kept in registers
//   warp  4    : lane 0 drives TMA (K/V key blocks of 128 = two image rows) and issues the tcgen05 MMAs
// Per key block: S = Q.K^T (128x128x80) into TMEM -> softmax with bias -> P (operand format) to smem ->
// O_blk = P.V into TMEM -> rescale-and-accumulate in registers.
// Relative position: with Rrev[j] = rel_pos[126 - j],  q.Rrev[63 - q_pos + k_pos] is the bias term; two prologue
// MMAs compute T_w = Q.Rw_rev^T (all 127 offsets; each thread keeps its 64 rel_w values, fp16-packed, in registers)
// and T_h = Q.Rh_rev[start..start+80)^T, which stays in TMEM: the two rel_h values a key block needs sit in
// adjacent columns at a warp-uniform offset.
#include <stdlib.h>

#include "host_common.h"
#include "kernels.h"
#include "ptx.cuh"

namespace {

constexpr int HD = 80;
constexpr int G = 64;          // token grid
constexpr int BQ = 128;        // queries per CTA
constexpr int BKV = 128;       // keys per block
constexpr int kThreads = 160;

constexpr int OFF_Q64 = 0;        // 128 x 128B SWIZZLE_128B
constexpr int OFF_K64 = 16384;
constexpr int OFF_V64 = 32768;
constexpr int OFF_P = 49152;      // 2 x (128 x 128B)
constexpr int OFF_Q16 = 81920;    // 128 x 32B SWIZZLE_32B
constexpr int OFF_K16 = 86016;
constexpr int OFF_V16 = 90112;
constexpr int OFF_BAR = 94208;
constexpr int kSmemBytes = OFF_BAR + 128 + 1024;
// prologue overlays
constexpr int OFF_RW64 = OFF_V64;  // Rw_rev rows 0..127 (K-major B operand)
constexpr int OFF_RW16 = OFF_V16;
constexpr int OFF_RH = OFF_K64;    // Rh_rev sub-table, 80 rows, un-swizzled core-matrix layout (5 x 80 x 32B)
constexpr int OFF_STAGE = OFF_K64; // 128 x 127 fp32 staging of T_w (65024 B, ends below OFF_Q16)
constexpr int kStageStride = 127;

constexpr uint32_t TM_S = 0;      // S / O_blk / T_w   (128 columns)
constexpr uint32_t TM_TH = 128;   // T_h               (80 columns)

struct GlobAttnMaps {
  CUtensorMap t64, t16;  // 2-D over qkv [B*4096, 3E]: box {64,128} SWIZZLE_128B and {16,128} SWIZZLE_32B
};

__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ uint32_t row_off64(int r, int c) { return r * 128 + ((c ^ (r & 7)) << 4); }
__device__ __forceinline__ uint32_t row_off16(int r, int c) { return r * 32 + ((c ^ ((r >> 2) & 1)) << 4); }
__device__ __forceinline__ void tmem_ld_x2(uint32_t taddr, uint32_t& a, uint32_t& b) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x2.b32 {%0, %1}, [%2];" : "=r"(a), "=r"(b) : "r"(taddr) : "memory");
}

__global__ void __launch_bounds__(kThreads, 2)
glob_attn_kernel(const __grid_constant__ GlobAttnMaps maps, const uint16_t* __restrict__ rh_rev,
                 const uint16_t* __restrict__ rw_rev, uint16_t* __restrict__ out, const int E, const int heads,
                 const int fmt, const float scale_log2e) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + OFF_BAR);
  uint64_t* bar_q = bars + 0;       // Q tile landed
  uint64_t* bar_t = bars + 1;       // prologue MMAs done
  uint64_t* bar_pro = bars + 2;     // softmax threads finished the prologue (count 128)
  uint64_t* k_full = bars + 3;
  uint64_t* v_full = bars + 4;
  uint64_t* s_full = bars + 5;
  uint64_t* p_ready = bars + 6;     // count 128
  uint64_t* o_full = bars + 7;
  uint64_t* o_read = bars + 8;      // count 128
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 9);

  const int tid = threadIdx.x, warp = tid >> 5;
  int w = blockIdx.x;
  const int qt = w % (G * G / BQ);
  w /= (G * G / BQ);
  const int head = w % heads;
  const int b = w / heads;
  const int qh0 = qt * 2;                       // first image row of this query tile
  const int th_start = ((62 - qh0) >> 3) << 3;  // first Rh_rev row held in T_h (multiple of 8, >= 0)
  const uint32_t sbase = ptx::smem_u32(smem);
  const int row0 = b * (G * G) + qt * BQ;
  const int cq = head * HD, ck = E + head * HD, cv = 2 * E + head * HD;
  constexpr int nblk = G * G / BKV;

  if (tid == 0) {
    ptx::prefetch_tmap(&maps.t64);
    ptx::prefetch_tmap(&maps.t16);
    ptx::mbar_init(bar_q, 1);
    ptx::mbar_init(bar_t, 1);
    ptx::mbar_init(bar_pro, 128);
    ptx::mbar_init(k_full, 1);
    ptx::mbar_init(v_full, 1);
    ptx::mbar_init(s_full, 1);
    ptx::mbar_init(p_ready, 128);
    ptx::mbar_init(o_full, 1);
    ptx::mbar_init(o_read, 128);
    ptx::fence_mbar_init();
  }
  if (warp == 4) {
    ptx::tmem_alloc(tmem_slot, 256);
    ptx::tmem_relinquish();
  }
  // rel-pos operand tables -> smem (generic proxy)
  for (int i = tid; i < 128 * 10; i += kThreads) {
    const int r = i / 10, c = i % 10;
    const uint4 v = __ldg(reinterpret_cast<const uint4*>(rw_rev + r * HD) + c);
    if (c < 8)
      *reinterpret_cast<uint4*>(smem + OFF_RW64 + row_off64(r, c)) = v;
    else
      *reinterpret_cast<uint4*>(smem + OFF_RW16 + row_off16(r, c - 8)) = v;
  }
  for (int i = tid; i < 80 * 10; i += kThreads) {
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

  if (warp == 4) {
    // ============================================================ control thread: TMA + MMA issue
    if ((tid & 31) == 0) {
      const uint32_t id_S = ptx::make_idesc((uint32_t)fmt, 128, 128, 0, 0);
      const uint32_t id_TH = ptx::make_idesc((uint32_t)fmt, 128, 80, 0, 0);
      const uint32_t id_O64 = ptx::make_idesc((uint32_t)fmt, 128, 64, 0, 1);
      const uint32_t id_O16 = ptx::make_idesc((uint32_t)fmt, 128, 16, 0, 1);
      ptx::mbar_expect_tx(bar_q, BQ * HD * 2);
      ptx::tma_load_2d(smem + OFF_Q64, &maps.t64, bar_q, cq, row0);
      ptx::tma_load_2d(smem + OFF_Q16, &maps.t16, bar_q, cq + 64, row0);
      ptx::mbar_wait(bar_q, 0);
      ptx::tc_fence_after();
      // T_w = Q . Rw_rev^T -> TM_S ;  T_h = Q . Rh_rev[th_start..+80)^T -> TM_TH
#pragma unroll
      for (int k = 0; k < 5; ++k) {
        const uint64_t da = (k < 4) ? ptx::make_smem_desc(sbase + OFF_Q64 + k * 32, 16, 1024, ptx::kSwz128)
                                    : ptx::make_smem_desc(sbase + OFF_Q16, 16, 256, ptx::kSwz32);
        const uint64_t dw = (k < 4) ? ptx::make_smem_desc(sbase + OFF_RW64 + k * 32, 16, 1024, ptx::kSwz128)
                                    : ptx::make_smem_desc(sbase + OFF_RW16, 16, 256, ptx::kSwz32);
        ptx::mma_f16_ss(tmem + TM_S, da, dw, id_S, k != 0);
        ptx::mma_f16_ss(tmem + TM_TH, da, ptx::make_smem_desc(sbase + OFF_RH + k * (80 * 32), 128, 256, ptx::kSwzNone),
                        id_TH, k != 0);
      }
      ptx::mma_commit(bar_t);
      ptx::mbar_wait(bar_pro, 0);  // staging area (aliases K/V/P) is free again
      ptx::mbar_expect_tx(k_full, BKV * HD * 2);
      ptx::tma_load_2d(smem + OFF_K64, &maps.t64, k_full, ck, b * (G * G));
      ptx::tma_load_2d(smem + OFF_K16, &maps.t16, k_full, ck + 64, b * (G * G));
      ptx::mbar_expect_tx(v_full, BKV * HD * 2);
      ptx::tma_load_2d(smem + OFF_V64, &maps.t64, v_full, cv, b * (G * G));
      ptx::tma_load_2d(smem + OFF_V16, &maps.t16, v_full, cv + 64, b * (G * G));
#pragma unroll 1
      for (int i = 0; i < nblk; ++i) {
        const uint32_t ph = i & 1;
        ptx::mbar_wait(k_full, ph);
        if (i > 0) ptx::mbar_wait(o_read, ph ^ 1);
        ptx::tc_fence_after();
#pragma unroll
        for (int k = 0; k < 5; ++k) {
          const uint64_t da = (k < 4) ? ptx::make_smem_desc(sbase + OFF_Q64 + k * 32, 16, 1024, ptx::kSwz128)
                                      : ptx::make_smem_desc(sbase + OFF_Q16, 16, 256, ptx::kSwz32);
          const uint64_t db = (k < 4) ? ptx::make_smem_desc(sbase + OFF_K64 + k * 32, 16, 1024, ptx::kSwz128)
                                      : ptx::make_smem_desc(sbase + OFF_K16, 16, 256, ptx::kSwz32);
          ptx::mma_f16_ss(tmem + TM_S, da, db, id_S, k != 0);
        }
        ptx::mma_commit(s_full);
        ptx::mbar_wait(s_full, ph);  // K tile consumed
        if (i + 1 < nblk) {
          const int r = b * (G * G) + (i + 1) * BKV;
          ptx::mbar_expect_tx(k_full, BKV * HD * 2);
          ptx::tma_load_2d(smem + OFF_K64, &maps.t64, k_full, ck, r);
          ptx::tma_load_2d(smem + OFF_K16, &maps.t16, k_full, ck + 64, r);
        }
        ptx::mbar_wait(p_ready, ph);
        ptx::mbar_wait(v_full, ph);
        ptx::tc_fence_after();
#pragma unroll
        for (int ks = 0; ks < BKV / 16; ++ks) {
          const uint64_t da = ptx::make_smem_desc(sbase + OFF_P + (ks >> 2) * 16384 + (ks & 3) * 32, 16, 1024, ptx::kSwz128);
          ptx::mma_f16_ss(tmem + TM_S, da, ptx::make_smem_desc(sbase + OFF_V64 + ks * 2048, BKV * 128, 1024, ptx::kSwz128),
                          id_O64, ks != 0);
          ptx::mma_f16_ss(tmem + TM_S + 64, da, ptx::make_smem_desc(sbase + OFF_V16 + ks * 512, BKV * 32, 256, ptx::kSwz32),
                          id_O16, ks != 0);
        }
        ptx::mma_commit(o_full);
        ptx::mbar_wait(o_full, ph);  // V tile and P consumed
        if (i + 1 < nblk) {
          const int r = b * (G * G) + (i + 1) * BKV;
          ptx::mbar_expect_tx(v_full, BKV * HD * 2);
          ptx::tma_load_2d(smem + OFF_V64, &maps.t64, v_full, cv, r);
          ptx::tma_load_2d(smem + OFF_V16, &maps.t16, v_full, cv + 64, r);
        }
      }
    }
  } else {
    // ============================================================ softmax threads (row = tid)
    const uint32_t trow = tmem + (static_cast<uint32_t>(warp * 32) << 16);
    const int qh = qh0 + (tid >> 6);
    const int qw = tid & 63;
    const float kLog2e = 1.4426950408889634f;
    uint32_t relw[32];  // rel_w[kw] * log2e, kw = 0..63, packed as fp16 pairs
    ptx::mbar_wait(bar_t, 0);
    ptx::tc_fence_after();
    {
      float* st = reinterpret_cast<float*>(smem + OFF_STAGE) + tid * kStageStride;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        uint32_t v[32];
        ptx::tmem_ld_32x32b_x32(trow + TM_S + c * 32, v);
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
    }
    ptx::tc_fence_before();
    ptx::mbar_arrive(bar_pro);

    float o[HD];
#pragma unroll
    for (int d = 0; d < HD; ++d) o[d] = 0.f;
    float m = -INFINITY, l = 0.f;
    const uint32_t th_col = TM_TH + static_cast<uint32_t>(63 - qh - th_start);

#pragma unroll 1
    for (int i = 0; i < nblk; ++i) {
      const uint32_t ph = i & 1;
      ptx::mbar_wait(s_full, ph);
      ptx::tc_fence_after();
      uint32_t h0, h1;
      tmem_ld_x2(trow + th_col + 2 * i, h0, h1);
      ptx::tmem_ld_wait();
      float rh[2] = {__uint_as_float(h0) * kLog2e, __uint_as_float(h1) * kLog2e};
      float bmax = -INFINITY;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        uint32_t v[32];
        ptx::tmem_ld_32x32b_x32(trow + TM_S + c * 32, v);
        ptx::tmem_ld_wait();
#pragma unroll
        for (int i2 = 0; i2 < 32; i2 += 2) {
          const int j = c * 32 + i2;
          const float2 rw = __half22float2(*reinterpret_cast<const __half2*>(&relw[(j & 63) >> 1]));
          bmax = fmaxf(bmax, fmaf(__uint_as_float(v[i2]), scale_log2e, rh[j >> 6]) + rw.x);
          bmax = fmaxf(bmax, fmaf(__uint_as_float(v[i2 + 1]), scale_log2e, rh[j >> 6]) + rw.y);
        }
      }
      const float m_new = fmaxf(m, bmax);
      const float alpha = ex2(m - m_new);
      rh[0] -= m_new;
      rh[1] -= m_new;
      float bsum = 0.f;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        uint32_t v[32];
        ptx::tmem_ld_32x32b_x32(trow + TM_S + c * 32, v);
        ptx::tmem_ld_wait();
        float p[32];
#pragma unroll
        for (int i2 = 0; i2 < 32; i2 += 2) {
          const int j = c * 32 + i2;
          const float2 rw = __half22float2(*reinterpret_cast<const __half2*>(&relw[(j & 63) >> 1]));
          p[i2] = ex2(fmaf(__uint_as_float(v[i2]), scale_log2e, rh[j >> 6]) + rw.x);
          p[i2 + 1] = ex2(fmaf(__uint_as_float(v[i2 + 1]), scale_log2e, rh[j >> 6]) + rw.y);
          bsum += p[i2] + p[i2 + 1];
        }
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          const int j0 = c * 32 + g * 8;
          uint4 u;
          u.x = ptx::pack2(p[g * 8 + 0], p[g * 8 + 1], fmt);
          u.y = ptx::pack2(p[g * 8 + 2], p[g * 8 + 3], fmt);
          u.z = ptx::pack2(p[g * 8 + 4], p[g * 8 + 5], fmt);
          u.w = ptx::pack2(p[g * 8 + 6], p[g * 8 + 7], fmt);
          *reinterpret_cast<uint4*>(smem + OFF_P + (j0 >> 6) * 16384 + row_off64(tid, (j0 & 63) >> 3)) = u;
        }
      }
      l = l * alpha + bsum;
      m = m_new;
      ptx::fence_proxy_async_smem();
      ptx::tc_fence_before();
      ptx::mbar_arrive(p_ready);
      ptx::mbar_wait(o_full, ph);
      ptx::tc_fence_after();
#pragma unroll
      for (int c = 0; c < 5; ++c) {
        uint32_t v[16];
        ptx::tmem_ld_32x32b_x16(trow + TM_S + c * 16, v);
        ptx::tmem_ld_wait();
#pragma unroll
        for (int d = 0; d < 16; ++d) o[c * 16 + d] = fmaf(o[c * 16 + d], alpha, __uint_as_float(v[d]));
      }
      ptx::tc_fence_before();
      ptx::mbar_arrive(o_read);
    }
    const float inv = 1.0f / l;
    uint16_t* dst = out + static_cast<size_t>(row0 + tid) * E + head * HD;
#pragma unroll
    for (int c = 0; c < 10; ++c) {
      uint4 u;
      u.x = ptx::pack2(o[c * 8 + 0] * inv, o[c * 8 + 1] * inv, fmt);
      u.y = ptx::pack2(o[c * 8 + 2] * inv, o[c * 8 + 3] * inv, fmt);
      u.z = ptx::pack2(o[c * 8 + 4] * inv, o[c * 8 + 5] * inv, fmt);
      u.w = ptx::pack2(o[c * 8 + 6] * inv, o[c * 8 + 7] * inv, fmt);
      reinterpret_cast<uint4*>(dst)[c] = u;
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 4) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem, 256);
  }
}

}  // namespace

int samk_attn_global(const void* qkv, const void* rh_rev, const void* rw_rev, void* out, int B, int E, int heads,
                     int fmt, cudaStream_t stream) {
  SAM_REQUIRE(fmt == 0 || fmt == 1, "attn_global: fmt must be fp16/bf16");
  SAM_REQUIRE(B > 0, "attn_global: empty batch");
  // default: the decoupled-pipeline kernel (attn_global3.cu); SAM_ATTN_GLOBAL_V2=1 selects the two-tile ping-pong kernel
  // (attn_global2.cu), SAM_ATTN_GLOBAL_V1=1 this file's simpler one-tile-per-CTA kernel
  static const bool use_v1 = getenv("SAM_ATTN_GLOBAL_V1") != nullptr;
  static const bool use_v2 = getenv("SAM_ATTN_GLOBAL_V2") != nullptr;
  if (!use_v1 && !use_v2) return samk_attn_global3(qkv, rh_rev, rw_rev, out, B, E, heads, fmt, stream);
  SAM_REQUIRE(E == heads * HD, "attn_global: the v1 / v2 kernels need head_dim 80 (E=%d heads=%d)", E, heads);
  if (use_v2) return samk_attn_global2(qkv, rh_rev, rw_rev, out, B, E, heads, fmt, stream);
  GlobAttnMaps maps;
  const int is_bf16 = (fmt == 1);
  const uint64_t rows = static_cast<uint64_t>(B) * G * G;
  int rc = samhost::encode_tmap_2d(&maps.t64, 2, is_bf16, qkv, 3ull * E, rows, 3ull * E * 2, 64, 128, 3);
  if (rc) return rc;
  rc = samhost::encode_tmap_2d(&maps.t16, 2, is_bf16, qkv, 3ull * E, rows, 3ull * E * 2, 16, 128, 1);
  if (rc) return rc;
  static bool attr_done = false;
  if (!attr_done) {
    SAM_CHECK_CUDA(cudaFuncSetAttribute(glob_attn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
    attr_done = true;
  }
  const float scale_log2e = 1.4426950408889634f / sqrtf(static_cast<float>(HD));
  const int grid = B * heads * (G * G / BQ);
  const double bh = static_cast<double>(B) * heads;
  samhost::LaunchScope scope(samhost::KC_ATTN_GLOBAL, stream, bh * (4.0 * 4096 * 4096 * 80 + 4.0 * 4096 * 64 * 80),
                             static_cast<double>(B) * 4096 * E * 2 * 4);
  glob_attn_kernel<<<grid, kThreads, kSmemBytes, stream>>>(maps, static_cast<const uint16_t*>(rh_rev),
                                                            static_cast<const uint16_t*>(rw_rev),
                                                            static_cast<uint16_t*>(out), E, heads, fmt, scale_log2e);
  SAM_CHECK_CUDA(cudaGetLastError());
  return 0;
}
