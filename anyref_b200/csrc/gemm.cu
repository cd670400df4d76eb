// Dense GEMM for the SAM ViT-H encoder on sm_100a:   C[M,N] = epilogue( A[M,K] * W[N,K]^T )
//
// Replaces the cuBLAS calls behind nn.Linear / 1x1 conv on the reference path
// (image_encoder.py:238 qkv, :258 proj, common.py:26 lin1/lin2, image_encoder.py:93 neck, :418 patch-embed).
//
// Structure (one CTA per SM, persistent over output tiles):
//   warp 0      : TMA producer  -- A/W tiles (128B-swizzled, K-major) into a kStages-deep smem ring
//   warp 1      : MMA issuer    -- tcgen05.mma 128 x BN x 16 (kind::f16, fp32 accumulate) into TMEM;
//                                  two accumulator stages so the epilogue of tile i overlaps tile i+1
//   warps 2..5  : epilogue      -- tcgen05.ld accumulator rows, + bias / GELU(erf) / + residual, store
//
// Operands are bf16 or fp16 (runtime `fmt`, same tensor-core rate); accumulation and epilogue math are fp32.
#include <stdio.h>
#include <stdlib.h>

#include "host_common.h"
#include "kernels.h"
#include "ptx.cuh"

namespace {

constexpr int BM = 128;
constexpr int BK = 64;  // 64 x 2B = one 128-byte swizzle row
constexpr int kGemmThreads = 192;

template <int BN>
struct GemmCfg {
  static constexpr int kStageA = BM * BK * 2;
  static constexpr int kStageB = BN * BK * 2;
  static constexpr int kStage = kStageA + kStageB;
  static constexpr int kStages = (BN == 256) ? 4 : 6;
  static constexpr int kSmem = kStages * kStage + 1024 /*align slack*/ + 256 /*barriers*/;
  static constexpr int kTmemCols = 2 * BN;
};

__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752f)); }

template <int BN>
__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_tn_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
               const GemmEpilogue ep, const int M, const int N, const int K, const uint32_t idesc) {
  using Cfg = GemmCfg<BN>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + Cfg::kStages * Cfg::kStage);
  uint64_t* empty_bar = full_bar + Cfg::kStages;
  uint64_t* acc_full = empty_bar + Cfg::kStages;
  uint64_t* acc_empty = acc_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int m_tiles = (M + BM - 1) / BM;
  const int n_tiles = (N + BN - 1) / BN;
  const int num_tiles = m_tiles * n_tiles;
  const int num_kb = (K + BK - 1) / BK;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tmA);
    ptx::prefetch_tmap(&tmB);
    for (int s = 0; s < Cfg::kStages; ++s) {
      ptx::mbar_init(&full_bar[s], 1);
      ptx::mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      ptx::mbar_init(&acc_full[s], 1);
      ptx::mbar_init(&acc_empty[s], 4);
    }
    ptx::fence_mbar_init();
  }
  if (warp == 1) {
    ptx::tmem_alloc(tmem_slot, Cfg::kTmemCols);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      int s = 0;
      uint32_t ph = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int m_blk = tile / n_tiles, n_blk = tile % n_tiles;
        for (int kb = 0; kb < num_kb; ++kb) {
          ptx::mbar_wait(&empty_bar[s], ph ^ 1);
          uint8_t* sa = smem + s * Cfg::kStage;
          uint8_t* sb = sa + Cfg::kStageA;
          ptx::mbar_expect_tx(&full_bar[s], Cfg::kStage);
          ptx::tma_load_2d(sa, &tmA, &full_bar[s], kb * BK, m_blk * BM);
          ptx::tma_load_2d(sb, &tmB, &full_bar[s], kb * BK, n_blk * BN);
          if (++s == Cfg::kStages) { s = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      int s = 0;
      uint32_t ph = 0;
      int as = 0;
      uint32_t aph = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        ptx::mbar_wait(&acc_empty[as], aph ^ 1);
        ptx::tc_fence_after();
        const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(as * BN);
        for (int kb = 0; kb < num_kb; ++kb) {
          ptx::mbar_wait(&full_bar[s], ph);
          ptx::tc_fence_after();
          const uint32_t sa = ptx::smem_u32(smem + s * Cfg::kStage);
          const uint32_t sb = sa + Cfg::kStageA;
          const uint64_t da = ptx::make_smem_desc(sa, 16, 1024, ptx::kSwz128);
          const uint64_t db = ptx::make_smem_desc(sb, 16, 1024, ptx::kSwz128);
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
            // advancing 16 elements (32 B) along K inside the 128B swizzle row: +2 in the (addr>>4) field
            ptx::mma_f16_ss(d_tmem, da + 2 * k, db + 2 * k, idesc, (kb | k) != 0);
          }
          ptx::mma_commit(&empty_bar[s]);
          if (++s == Cfg::kStages) { s = 0; ph ^= 1; }
        }
        ptx::mma_commit(&acc_full[as]);
        if (++as == 2) { as = 0; aph ^= 1; }
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue (4 warps)
    const int quad = warp & 3;  // TMEM lane quadrant this warp may access
    int as = 0;
    uint32_t aph = 0;
    const int fmt = ep.out_fmt;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int m_blk = tile / n_tiles, n_blk = tile % n_tiles;
      ptx::mbar_wait(&acc_full[as], aph);
      ptx::tc_fence_after();
      const int row = m_blk * BM + quad * 32 + lane;
      const bool row_ok = row < M;
      const uint32_t t_row = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + static_cast<uint32_t>(as * BN);
      const float* res_row = nullptr;
      if (ep.res && row_ok) res_row = ep.res + static_cast<size_t>(row % ep.res_mod) * ep.ldr;
#pragma unroll 1
      for (int c = 0; c < BN / 32; ++c) {
        uint32_t v[32];
        ptx::tmem_ld_32x32b_x32(t_row + c * 32, v);
        ptx::tmem_ld_wait();
        const int col0 = n_blk * BN + c * 32;
        if (row_ok && col0 < N) {
          float f[32];
#pragma unroll
          for (int i = 0; i < 32; ++i) f[i] = __uint_as_float(v[i]);
          if (ep.bias) {
            const float4* b4 = reinterpret_cast<const float4*>(ep.bias + col0);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const float4 b = __ldg(b4 + i);
              f[4 * i + 0] += b.x; f[4 * i + 1] += b.y; f[4 * i + 2] += b.z; f[4 * i + 3] += b.w;
            }
          }
          if (ep.act == 1) {
#pragma unroll
            for (int i = 0; i < 32; ++i) f[i] = gelu_erf(f[i]);
          } else if (ep.act == 2) {   // ReLU (text_hidden_fcs, model/anyref.py:118-123)
#pragma unroll
            for (int i = 0; i < 32; ++i) f[i] = fmaxf(f[i], 0.0f);
          }
          if (res_row) {
            const float4* r4 = reinterpret_cast<const float4*>(res_row + col0);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const float4 r = r4[i];
              f[4 * i + 0] += r.x; f[4 * i + 1] += r.y; f[4 * i + 2] += r.z; f[4 * i + 3] += r.w;
            }
          }
          if (fmt == 2) {
            float4* o4 = reinterpret_cast<float4*>(static_cast<float*>(ep.out) + static_cast<size_t>(row) * ep.ldo + col0);
#pragma unroll
            for (int i = 0; i < 8; ++i) o4[i] = make_float4(f[4 * i], f[4 * i + 1], f[4 * i + 2], f[4 * i + 3]);
          } else {
            uint4* o4 = reinterpret_cast<uint4*>(static_cast<uint16_t*>(ep.out) + static_cast<size_t>(row) * ep.ldo + col0);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              uint4 u;
              u.x = ptx::pack2(f[8 * i + 0], f[8 * i + 1], fmt);
              u.y = ptx::pack2(f[8 * i + 2], f[8 * i + 3], fmt);
              u.z = ptx::pack2(f[8 * i + 4], f[8 * i + 5], fmt);
              u.w = ptx::pack2(f[8 * i + 6], f[8 * i + 7], fmt);
              o4[i] = u;
            }
          }
        }
      }
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&acc_empty[as]);
      if (++as == 2) { as = 0; aph ^= 1; }
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, Cfg::kTmemCols);
  }
}

template <int BN>
int launch_gemm(const void* A, int lda, const void* W, int ldw, int M, int N, int K, int fmt, const GemmEpilogue& ep,
                int max_ctas, cudaStream_t stream) {
  using Cfg = GemmCfg<BN>;
  CUtensorMap tmA, tmB;
  const int is_bf16 = (fmt == 1);
  int rc = samhost::encode_tmap_2d(&tmA, 2, is_bf16, A, (uint64_t)K, (uint64_t)M, (uint64_t)lda * 2, BK, BM, 3);
  if (rc) return rc;
  rc = samhost::encode_tmap_2d(&tmB, 2, is_bf16, W, (uint64_t)K, (uint64_t)N, (uint64_t)ldw * 2, BK, BN, 3);
  if (rc) return rc;
  static samhost::PerDeviceOnce attr_once;
  if (attr_once.need()) {
    SAM_CHECK_CUDA(cudaFuncSetAttribute(gemm_tn_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmem));
    attr_once.done();
  }
  const int m_tiles = (M + BM - 1) / BM, n_tiles = (N + BN - 1) / BN;
  int grid = m_tiles * n_tiles;
  int cap = samhost::sm_count();
  if (max_ctas > 0 && max_ctas < cap) cap = max_ctas;
  if (grid > cap) grid = cap;
  const uint32_t idesc = ptx::make_idesc((uint32_t)fmt, BM, BN, 0, 0);
  const double out_b = (ep.out_fmt == 2) ? 4.0 : 2.0;
  samhost::LaunchScope scope(samhost::KC_GEMM, stream, 2.0 * M * N * K,
                             2.0 * (static_cast<double>(M) * K + static_cast<double>(N) * K) +
                                 out_b * M * N + (ep.res ? 4.0 * M * N : 0.0));
  gemm_tn_kernel<BN><<<grid, kGemmThreads, Cfg::kSmem, stream>>>(tmA, tmB, ep, M, N, K, idesc);
  SAM_CHECK_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace

int samk_gemm(const void* A, int lda, const void* W, int ldw, int M, int N, int K, int fmt, const GemmEpilogue& ep,
              cudaStream_t stream) {
  SAM_REQUIRE(fmt == 0 || fmt == 1, "gemm: operand fmt must be 0 (fp16) or 1 (bf16), got %d", fmt);
  SAM_REQUIRE(ep.out_fmt >= 0 && ep.out_fmt <= 2, "gemm: bad out_fmt %d", ep.out_fmt);
  SAM_REQUIRE(M > 0 && N > 0 && K > 0, "gemm: empty problem %dx%dx%d", M, N, K);
  SAM_REQUIRE(N % 32 == 0, "gemm: N=%d must be a multiple of 32", N);
  SAM_REQUIRE(K % 8 == 0 && lda % 8 == 0 && ldw % 8 == 0, "gemm: K/lda/ldw must be multiples of 8 (16-byte rows)");
  SAM_REQUIRE((reinterpret_cast<uintptr_t>(A) & 15) == 0 && (reinterpret_cast<uintptr_t>(W) & 15) == 0,
              "gemm: operands must be 16-byte aligned");
  SAM_REQUIRE(ep.ldo % 8 == 0 && (reinterpret_cast<uintptr_t>(ep.out) & 15) == 0, "gemm: output must be 16B aligned");
  if (ep.res) SAM_REQUIRE(ep.res_mod > 0 && ep.ldr % 4 == 0, "gemm: residual needs res_mod>0, ldr%%4==0");
  // big problems go to the 2-CTA kernel (gemm2.cu); it declines (-1) epilogues it does not implement
  static const bool force_v1 = getenv("SAM_GEMM_V1") != nullptr;
  if (!force_v1) {
    const int rc2 = samk_gemm2(A, lda, W, ldw, M, N, K, fmt, ep, stream);
    if (rc2 >= 0) return rc2;
  }
  SAM_REQUIRE(!ep.ln_stats && !ep.xb && !ep.diag_mt,
              "gemm: the LayerNorm-folding / block-diagonal modes exist only in the 2-CTA kernel (shape %dx%dx%d declined)", M, N, K);
  if (N % 256 == 0 || N > 256) return launch_gemm<256>(A, lda, W, ldw, M, N, K, fmt, ep, 0, stream);
  return launch_gemm<128>(A, lda, W, ldw, M, N, K, fmt, ep, 0, stream);
}
