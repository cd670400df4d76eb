// Memory-bound glue kernels of the SAM ViT-H encoder (vectorised, one pass over HBM each):
//   samk_layernorm_rows  : nn.LayerNorm over the channel dim of token-major fp32 rows (optionally of x + res, the
//                          decoder's post-residual norms, transformer.py:157-181) -> operand format or fp32
//                          (image_encoder.py:179 norm1, :191 norm2; also LayerNorm2d on NHWC rows, common.py:38-43)
//   samk_patch_im2col    : NCHW image -> [B*g*g, 3*p*p] patch matrix (A operand of the patch-embed GEMM,
//                          image_encoder.py:418-426)
//   samk_im2col3x3       : NHWC [B,g,g,C] -> [B*g*g, 9C] with zero padding (neck 3x3 conv, image_encoder.py:100-106)
//   samk_ln_nhwc_to_nchw : LayerNorm2d (common.py:31-43) on NHWC fp32 rows fused with the NHWC->NCHW transposition
//                          that produces the encoder output [B,C,g,g] (image_encoder.py:107, :124)
#include <stdlib.h>

#include "host_common.h"
#include "kernels.h"
#include "ptx.cuh"

namespace {

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

constexpr int kLnMaxVec = 10;  // float4 per lane: supports C <= 1280

// One warp per row. normalize == 0 degenerates to a dtype cast (used in front of the neck GEMM).
__global__ void __launch_bounds__(256)
layernorm_rows_kernel(const float* x, int ldx, const float* res, int ldr, const float* __restrict__ gamma,
                      const float* __restrict__ beta, float eps, void* out, int ldo, int out_fmt, int M, int C,
                      int normalize) {
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= M) return;
  const int nvec = C >> 2;
  const float4* xr = reinterpret_cast<const float4*>(x + static_cast<size_t>(row) * ldx);
  const float4* rr = res ? reinterpret_cast<const float4*>(res + static_cast<size_t>(row) * ldr) : nullptr;
  float4 v[kLnMaxVec];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < kLnMaxVec; ++i) {
    const int idx = lane + i * 32;
    if (idx < nvec) {
      v[i] = xr[idx];
      if (rr) {
        const float4 r = rr[idx];
        v[i].x += r.x; v[i].y += r.y; v[i].z += r.z; v[i].w += r.w;
      }
      s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
    }
  }
  float mean = 0.f, rstd = 1.f;
  if (normalize) {
    mean = warp_sum(s) / static_cast<float>(C);
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < kLnMaxVec; ++i) {
      const int idx = lane + i * 32;
      if (idx < nvec) {
        const float a = v[i].x - mean, b = v[i].y - mean, c = v[i].z - mean, d = v[i].w - mean;
        q += (a * a + b * b) + (c * c + d * d);
      }
    }
    rstd = rsqrtf(warp_sum(q) / static_cast<float>(C) + eps);
  }
#pragma unroll
  for (int i = 0; i < kLnMaxVec; ++i) {
    const int idx = lane + i * 32;
    if (idx < nvec) {
      float4 y = v[i];
      if (normalize) {
        const float4 g = __ldg(reinterpret_cast<const float4*>(gamma) + idx);
        const float4 b = __ldg(reinterpret_cast<const float4*>(beta) + idx);
        y.x = (y.x - mean) * rstd * g.x + b.x;
        y.y = (y.y - mean) * rstd * g.y + b.y;
        y.z = (y.z - mean) * rstd * g.z + b.z;
        y.w = (y.w - mean) * rstd * g.w + b.w;
      }
      if (out_fmt == 2) {
        reinterpret_cast<float4*>(static_cast<float*>(out) + static_cast<size_t>(row) * ldo)[idx] = y;
      } else {
        uint2 u;
        u.x = ptx::pack2(y.x, y.y, out_fmt);
        u.y = ptx::pack2(y.z, y.w, out_fmt);
        reinterpret_cast<uint2*>(static_cast<uint16_t*>(out) + static_cast<size_t>(row) * ldo)[idx] = u;
      }
    }
  }
}


// ---------------------------------------------------------------------------------------------------------------
// Streaming LayerNorm for the encoder's big activations (M = B*4096 rows of 1280 fp32 -> 16-bit operand rows).
// The register-resident kernel above keeps only ~120 KB of loads in flight per SM (one row per warp, bounded by
// the register file), which is half of what HBM3e needs.  Here one persistent CTA per SM keeps a 32-row ring in
// shared memory filled by 1-D bulk async copies (TMA engine, mbarrier completion): 8 consumer warps normalise rows
// straight out of shared memory while the producer thread keeps ~160 KB of reads in flight.
// ---------------------------------------------------------------------------------------------------------------
constexpr int kLnStages = 5;      // ring slots (100 KB: two CTAs per SM)
constexpr int kLnGroup = 4;       // consecutive rows per slot (one bulk copy)

constexpr int kLnConsumerWarps = 16;

__global__ void __launch_bounds__(32 * (kLnConsumerWarps + 1), 2)
layernorm_stream_kernel(const float* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ beta,
                        float eps, void* __restrict__ out, int ldo, int out_fmt, int M, int C) {
  extern __shared__ __align__(128) uint8_t ln_smem[];
  const int row_bytes = C * 4;
  const int slot_bytes = kLnGroup * row_bytes;
  uint64_t* full = reinterpret_cast<uint64_t*>(ln_smem + kLnStages * slot_bytes);
  uint64_t* empty = full + kLnStages;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < kLnStages; ++s) {
      ptx::mbar_init(&full[s], 1);
      ptx::mbar_init(&empty[s], kLnGroup);
    }
    ptx::fence_mbar_init();
  }
  __syncthreads();
  // row groups of this CTA: blockIdx.x, blockIdx.x + gridDim.x, ...  (x rows are contiguous: ldx == C)
  const int num_groups = M / kLnGroup;
  const int my_groups = (num_groups - static_cast<int>(blockIdx.x) + static_cast<int>(gridDim.x) - 1) / static_cast<int>(gridDim.x);
  if (warp == kLnConsumerWarps) {
    if (lane == 0) {
      for (int i = 0; i < my_groups; ++i) {
        const int s = i % kLnStages;
        const uint32_t ph = (i / kLnStages) & 1;
        if (i >= kLnStages) ptx::mbar_wait(&empty[s], ph ^ 1);
        const size_t grp = static_cast<size_t>(blockIdx.x) + static_cast<size_t>(i) * gridDim.x;
        ptx::mbar_expect_tx(&full[s], slot_bytes);
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                         ptx::smem_u32(ln_smem + s * slot_bytes)),
                     "l"(x + grp * kLnGroup * C), "r"(slot_bytes), "r"(ptx::smem_u32(&full[s]))
                     : "memory");
      }
    }
    return;
  }
  // consumer warp w: row (w & 3) of every (kLnConsumerWarps / 4)-th group, starting with group (w >> 2)
  const int nvec = C >> 2;
  const int sub = warp & (kLnGroup - 1);
  for (int i = warp >> 2; i < my_groups; i += kLnConsumerWarps / kLnGroup) {
    const int s = i % kLnStages;
    const uint32_t ph = (i / kLnStages) & 1;
    ptx::mbar_wait(&full[s], ph);
    const float4* xr = reinterpret_cast<const float4*>(ln_smem + s * slot_bytes + sub * row_bytes);
    float4 v[kLnMaxVec];
    float sum = 0.f;
#pragma unroll
    for (int j = 0; j < kLnMaxVec; ++j) {
      const int idx = lane + j * 32;
      if (idx < nvec) {
        v[j] = xr[idx];
        sum += (v[j].x + v[j].y) + (v[j].z + v[j].w);
      }
    }
    __syncwarp();
    if (lane == 0) ptx::mbar_arrive(&empty[s]);   // row copied to registers: hand the slot back
    const float mean = warp_sum(sum) / static_cast<float>(C);
    float q = 0.f;
#pragma unroll
    for (int j = 0; j < kLnMaxVec; ++j) {
      const int idx = lane + j * 32;
      if (idx < nvec) {
        const float a = v[j].x - mean, b = v[j].y - mean, c = v[j].z - mean, d = v[j].w - mean;
        q += (a * a + b * b) + (c * c + d * d);
      }
    }
    const float rstd = rsqrtf(warp_sum(q) / static_cast<float>(C) + eps);
    const size_t row = (static_cast<size_t>(blockIdx.x) + static_cast<size_t>(i) * gridDim.x) * kLnGroup + sub;
#pragma unroll
    for (int j = 0; j < kLnMaxVec; ++j) {
      const int idx = lane + j * 32;
      if (idx < nvec) {
        const float4 g = __ldg(reinterpret_cast<const float4*>(gamma) + idx);
        const float4 b = __ldg(reinterpret_cast<const float4*>(beta) + idx);
        float4 y;
        y.x = (v[j].x - mean) * rstd * g.x + b.x;
        y.y = (v[j].y - mean) * rstd * g.y + b.y;
        y.z = (v[j].z - mean) * rstd * g.z + b.z;
        y.w = (v[j].w - mean) * rstd * g.w + b.w;
        if (out_fmt == 2) {
          reinterpret_cast<float4*>(static_cast<float*>(out) + row * ldo)[idx] = y;
        } else {
          uint2 u;
          u.x = ptx::pack2(y.x, y.y, out_fmt);
          u.y = ptx::pack2(y.z, y.w, out_fmt);
          reinterpret_cast<uint2*>(static_cast<uint16_t*>(out) + row * ldo)[idx] = u;
        }
      }
    }
  }
}

__device__ __forceinline__ void load8(const void* img, int fmt, size_t idx, float (&f)[8]) {
  if (fmt == 2) {
    const float4 a = reinterpret_cast<const float4*>(static_cast<const float*>(img) + idx)[0];
    const float4 b = reinterpret_cast<const float4*>(static_cast<const float*>(img) + idx)[1];
    f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w; f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
  } else {
    const uint4 u = *reinterpret_cast<const uint4*>(static_cast<const uint16_t*>(img) + idx);
    const float2 a = ptx::unpack2(u.x, fmt), b = ptx::unpack2(u.y, fmt), c = ptx::unpack2(u.z, fmt),
                 d = ptx::unpack2(u.w, fmt);
    f[0] = a.x; f[1] = a.y; f[2] = b.x; f[3] = b.y; f[4] = c.x; f[5] = c.y; f[6] = d.x; f[7] = d.y;
  }
}

// Each thread moves 8 consecutive pixels of one patch row: out[(b,py,px), c*p*p + ky*p + kx0..kx0+7].
// grid (ceil(S / 8 / 128), S, B * 3), block 128: the image row and the (batch, channel) plane come from the block indices, so the
// only index arithmetic per thread is one 32-bit division by the patch width (the flat 64-bit index decomposition of
// round 1 made this copy ALU-bound: ncu SM 75 %, DRAM 30 %).
__global__ void __launch_bounds__(256)
patch_im2col_kernel(const void* __restrict__ img, int in_fmt, uint16_t* __restrict__ out, int out_fmt, int B, int S,
                    int p) {
  const int g = S / p;
  const int x8 = blockIdx.x * blockDim.x + threadIdx.x;      // 8-pixel group of the image row
  if (x8 * 8 >= S) return;
  const int y = blockIdx.y;
  const int c = blockIdx.z % 3, b = blockIdx.z / 3;
  const int px = (x8 * 8) / p, kx0 = (x8 * 8) % p;
  const int py = y / p, ky = y % p;
  float f[8];
  load8(img, in_fmt, ((static_cast<size_t>(b) * 3 + c) * S + y) * S + x8 * 8, f);
  uint4 u;
  u.x = ptx::pack2(f[0], f[1], out_fmt);
  u.y = ptx::pack2(f[2], f[3], out_fmt);
  u.z = ptx::pack2(f[4], f[5], out_fmt);
  u.w = ptx::pack2(f[6], f[7], out_fmt);
  const size_t orow = (static_cast<size_t>(b) * g + py) * g + px;
  *reinterpret_cast<uint4*>(out + orow * (3 * p * p) + c * p * p + ky * p + kx0) = u;
}

// out[(b,y,x), (ky*3+kx)*C + c] = in[b, y+ky-1, x+kx-1, c] (zero outside).  One warp per output token: the lanes copy
// the 9 neighbour rows (C * 2 bytes each, lane-contiguous 16-byte pieces) -- no per-thread index decomposition.
// grid (g * g / 8, B), block 256.
__global__ void __launch_bounds__(256)
im2col3x3_kernel(const uint16_t* __restrict__ in, uint16_t* __restrict__ out, int B, int g, int C) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tok = blockIdx.x * 8 + warp;
  if (tok >= g * g) return;
  const int b = blockIdx.y;
  const int y = tok / g, x = tok % g;
  const int cv = C / 8;
  const uint4* src = reinterpret_cast<const uint4*>(in) + static_cast<size_t>(b) * g * g * cv;
  uint4* dst = reinterpret_cast<uint4*>(out) + (static_cast<size_t>(b) * g * g + tok) * 9 * cv;
#pragma unroll
  for (int tap = 0; tap < 9; ++tap) {
    const int yy = y + tap / 3 - 1, xx = x + tap % 3 - 1;
    const bool in_range = yy >= 0 && yy < g && xx >= 0 && xx < g;
    for (int c8 = lane; c8 < cv; c8 += 32) {
      uint4 v = make_uint4(0, 0, 0, 0);
      if (in_range) v = __ldg(src + static_cast<size_t>(yy * g + xx) * cv + c8);
      dst[tap * cv + c8] = v;
    }
  }
}

// Block = 32 consecutive tokens x C channels (C <= 256): per-token LayerNorm over channels, then a transposed,
// coalesced store into NCHW.
__global__ void __launch_bounds__(256)
ln_nhwc_to_nchw_kernel(const float* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ beta,
                       float eps, void* __restrict__ out, int out_fmt, int tokens_per_img, int C) {
  __shared__ float tile[256][33];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const size_t tok0 = static_cast<size_t>(blockIdx.x) * 32;
  for (int p = warp; p < 32; p += 8) {
    const float* xr = x + (tok0 + p) * C;
    float v[8];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int c = lane + i * 32;
      v[i] = (c < C) ? xr[c] : 0.f;
      s += v[i];
    }
    const float mean = warp_sum(s) / static_cast<float>(C);
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int c = lane + i * 32;
      if (c < C) q += (v[i] - mean) * (v[i] - mean);
    }
    const float rstd = rsqrtf(warp_sum(q) / static_cast<float>(C) + eps);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int c = lane + i * 32;
      if (c < C) tile[c][p] = (v[i] - mean) * rstd * __ldg(gamma + c) + __ldg(beta + c);
    }
  }
  __syncthreads();
  const size_t img = tok0 / tokens_per_img;
  const size_t pix0 = tok0 % tokens_per_img;
  for (int c = warp; c < C; c += 8) {
    const float y = tile[c][lane];
    const size_t o = (img * C + c) * tokens_per_img + pix0 + lane;
    if (out_fmt == 2)
      static_cast<float*>(out)[o] = y;
    else
      static_cast<uint16_t*>(out)[o] = ptx::pack1(y, out_fmt);
  }
}

// First producer of the LayerNorm-folding chain (gemm2.cu): xb = round(x) in the operand format plus the per-row
// partial statistics (mean, sum of squared deviations) of every 128-column slice -- what the residual GEMM epilogues emit for all later
// blocks.  One warp per row; lane l holds columns 128 j + 4 l .. + 3 of slice j.
// With `pos` the broadcast rows pos[row % pos_mod] (pos_embed, image_encoder.py:112-113) are added first and the sum is
// written back to x, so the patch-embed GEMM can use the plain fp32 store epilogue of the 2-CTA kernel.
__global__ void __launch_bounds__(256)
cast_stats_kernel(float* __restrict__ x, int ldx, void* __restrict__ xb, int ldxb, int fmt,
                  float2* __restrict__ stats, int M, int C, const float* __restrict__ pos, int pos_mod) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int parts = C >> 7;
  for (int row = blockIdx.x * 8 + warp; row < M; row += gridDim.x * 8) {
    float4* xr = reinterpret_cast<float4*>(x + static_cast<size_t>(row) * ldx);
    const float4* pr = pos ? reinterpret_cast<const float4*>(pos + static_cast<size_t>(row % pos_mod) * C) : nullptr;
    uint2* orow = reinterpret_cast<uint2*>(static_cast<uint16_t*>(xb) + static_cast<size_t>(row) * ldxb);
    for (int j = 0; j < parts; ++j) {
      float4 v = xr[j * 32 + lane];
      if (pr) {
        const float4 q = __ldg(pr + j * 32 + lane);
        v.x += q.x; v.y += q.y; v.z += q.z; v.w += q.w;
        xr[j * 32 + lane] = v;
      }
      uint2 u;
      u.x = ptx::pack2(v.x, v.y, fmt);
      u.y = ptx::pack2(v.z, v.w, fmt);
      orow[j * 32 + lane] = u;
      // (mean, M2 = sum of squared deviations from that mean) of the slice: the consumer combines the slices with
      // Chan's formula, so no E[x^2] - mean^2 cancellation occurs even for rows whose |mean| is far above their std
      const float mean = warp_sum((v.x + v.y) + (v.z + v.w)) * (1.0f / 128.0f);
      const float dx = v.x - mean, dy = v.y - mean, dz = v.z - mean, dw = v.w - mean;
      const float m2 = warp_sum(fmaf(dx, dx, fmaf(dy, dy, fmaf(dz, dz, dw * dw))));
      if (lane == 0) stats[static_cast<size_t>(row) * parts + j] = make_float2(mean, m2);
    }
  }
}

}  // namespace

int samk_cast_stats(float* x, int ldx, void* xb, int ldxb, int fmt, void* stats, int M, int C, const float* pos,
                    int pos_mod, cudaStream_t stream) {
  SAM_REQUIRE(C % 128 == 0 && ldx % 4 == 0 && ldxb % 4 == 0 && M > 0, "cast_stats: C=%d must be a multiple of 128", C);
  SAM_REQUIRE(fmt == 0 || fmt == 1, "cast_stats: output must be fp16/bf16");
  samhost::LaunchScope scope(samhost::KC_LAYERNORM, stream, 0.0, static_cast<double>(M) * C * 6.0);
  int grid = (M + 7) / 8;
  const int cap = samhost::sm_count() * 16;
  if (grid > cap) grid = cap;
  SAM_REQUIRE(!pos || pos_mod > 0, "cast_stats: pos_mod must be positive");
  cast_stats_kernel<<<grid, 256, 0, stream>>>(x, ldx, xb, ldxb, fmt, static_cast<float2*>(stats), M, C, pos, pos_mod);
  SAM_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int samk_layernorm_rows(const float* x, int ldx, const float* res, int ldr, const float* gamma, const float* beta,
                        float eps, void* out, int ldo, int out_fmt, int M, int C, int normalize, cudaStream_t stream) {
  SAM_REQUIRE(C % 4 == 0 && C <= kLnMaxVec * 128, "layernorm: C=%d must be a multiple of 4 and <= %d", C, kLnMaxVec * 128);
  SAM_REQUIRE(ldx % 4 == 0 && ldo % 4 == 0 && (!res || ldr % 4 == 0), "layernorm: leading dimensions must be multiples of 4");
  SAM_REQUIRE(M > 0, "layernorm: empty input");
  SAM_REQUIRE(!normalize || (gamma && beta), "layernorm: affine parameters missing");
  samhost::LaunchScope scope(samhost::KC_LAYERNORM, stream, 0.0,
                             static_cast<double>(M) * C * ((res ? 8.0 : 4.0) + (out_fmt == 2 ? 4.0 : 2.0)));
  // big plain LayerNorms (the encoder's norm1 / norm2) stream through the shared-memory ring
  static const bool no_stream = getenv("SAM_LN_NO_STREAM") != nullptr;
  if (!no_stream && normalize && !res && M >= 8192 && M % kLnGroup == 0 && (C * 4) % 16 == 0 && ldx == C &&
      (reinterpret_cast<uintptr_t>(x) & 15) == 0 && x != out) {
    const int smem = kLnStages * kLnGroup * C * 4 + 2 * kLnStages * 8;
    static int smem_set[64] = {0};   // per device (function attributes are per device)
    const int slot = samhost::device_slot();
    if (smem > smem_set[slot]) {
      SAM_CHECK_CUDA(cudaFuncSetAttribute(layernorm_stream_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
      smem_set[slot] = smem;
    }
    int grid = 2 * samhost::sm_count();
    layernorm_stream_kernel<<<grid, 32 * (kLnConsumerWarps + 1), smem, stream>>>(x, gamma, beta, eps, out, ldo, out_fmt, M, C);
    SAM_CHECK_CUDA(cudaGetLastError());
    return 0;
  }
  layernorm_rows_kernel<<<(M + 7) / 8, 256, 0, stream>>>(x, ldx, res, ldr, gamma, beta, eps, out, ldo, out_fmt, M, C,
                                                         normalize);
  SAM_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int samk_patch_im2col(const void* img, int in_fmt, void* out, int out_fmt, int B, int S, int p, cudaStream_t stream) {
  SAM_REQUIRE(p % 8 == 0 && S % p == 0, "patch_im2col: patch %d / image %d unsupported", p, S);
  SAM_REQUIRE(out_fmt == 0 || out_fmt == 1, "patch_im2col: output must be fp16/bf16");
  const size_t total = static_cast<size_t>(B) * 3 * S * (S / p) * (p / 8);
  samhost::LaunchScope scope(samhost::KC_LAYOUT, stream, 0.0,
                             static_cast<double>(B) * 3 * S * S * ((in_fmt == 2 ? 4.0 : 2.0) + 2.0));
  (void)total;
  SAM_REQUIRE(S <= 65535 && B * 3 <= 65535, "patch_im2col: image side / batch too large for the launch grid");
  dim3 grid((S / 8 + 127) / 128, S, B * 3);
  patch_im2col_kernel<<<grid, 128, 0, stream>>>(img, in_fmt, static_cast<uint16_t*>(out), out_fmt, B, S, p);
  SAM_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int samk_im2col3x3(const void* in, void* out, int B, int g, int C, cudaStream_t stream) {
  SAM_REQUIRE(C % 8 == 0, "im2col3x3: C must be a multiple of 8");
  const size_t total = static_cast<size_t>(B) * g * g * 9 * (C / 8);
  samhost::LaunchScope scope(samhost::KC_LAYOUT, stream, 0.0, static_cast<double>(B) * g * g * C * 2.0 * 10);
  (void)total;
  SAM_REQUIRE(B <= 65535, "im2col3x3: batch too large for the launch grid");
  dim3 grid((g * g + 7) / 8, B);
  im2col3x3_kernel<<<grid, 256, 0, stream>>>(static_cast<const uint16_t*>(in), static_cast<uint16_t*>(out), B, g, C);
  SAM_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int samk_ln_nhwc_to_nchw(const float* x, const float* gamma, const float* beta, float eps, void* out, int out_fmt,
                         int B, int tokens_per_img, int C, cudaStream_t stream) {
  SAM_REQUIRE(C <= 256 && tokens_per_img % 32 == 0, "ln_nhwc_to_nchw: C<=256 and tokens%%32==0 required");
  const size_t blocks = static_cast<size_t>(B) * tokens_per_img / 32;
  samhost::LaunchScope scope(samhost::KC_LAYERNORM, stream, 0.0,
                             static_cast<double>(B) * tokens_per_img * C * (4.0 + (out_fmt == 2 ? 4.0 : 2.0)));
  ln_nhwc_to_nchw_kernel<<<static_cast<unsigned>(blocks), 256, 0, stream>>>(x, gamma, beta, eps, out, out_fmt,
                                                                           tokens_per_img, C);
  SAM_CHECK_CUDA(cudaGetLastError());
  return 0;
}
