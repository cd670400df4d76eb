// ImageEncoderViT.forward (image_encoder.py:110-125) as one enqueue of hand-written sm_100a kernels.
//
// Per block (Block.forward, image_encoder.py:177-193) with the fp32 residual stream x [B*g*g, E]:
//   a   = LN1(x)                                  layernorm_rows -> operand format
//   qkv = a . Wqkv^T + b                          tcgen05 GEMM   (un-partitioned: padded window rows are never computed)
//   o   = attention(qkv)                          windowed (partition / rel-pos / softmax / un-partition fused) or global
//   x  += o . Wproj^T + b                         tcgen05 GEMM, fp32 residual epilogue (in place)
//   a   = LN2(x)
//   h   = GELU(a . W1^T + b1)                     tcgen05 GEMM, erf-GELU epilogue -> operand format
//   x  += h . W2^T + b2                           tcgen05 GEMM, fp32 residual epilogue (in place)
// Patch embedding = im2col + GEMM with (+bias, +pos_embed) epilogue; neck = 1x1 GEMM -> LN2d -> 3x3 (im2col GEMM) ->
// LN2d fused with the NHWC->NCHW store.
//
// With SamEncoderShape.ln_fold the two LayerNorm launches of a block disappear (5 launches per block): the residual
// GEMMs also emit xb = round(x) and per-row partial sums, and qkv / lin1 (packed as gamma o W) finish the
// normalisation in their epilogues (gemm2.cu, "LayerNorm folding"):
//   qkv = rstd * (xb . Wqkv'^T - mean * colsum) + (beta1 . Wqkv^T + b)
//   x  += o . Wproj^T + b ;  xb, stats <- x
//   h   = GELU(rstd * (xb . W1'^T - mean * colsum) + (beta2 . W1^T + b1))
//   x  += h . W2^T + b2   ;  xb, stats <- x
#include <stdio.h>

#include "host_common.h"
#include "kernels.h"

namespace {

struct BlockW16 {
  const uint16_t *qkv_w, *proj_w, *lin1_w, *lin2_w, *qkv_b_op, *rel;
};
struct BlockW32 {
  const float *n1w, *n1b, *qkv_b, *proj_b, *n2w, *n2b, *lin1_b, *lin2_b;
};

inline size_t align8(size_t n) { return (n + 7) & ~size_t(7); }

}  // namespace

// Blob layouts (all segments padded to a multiple of 8 elements so every pointer stays 16-byte aligned):
//   w16: patch_w [E, 3*p*p] | per block { qkv_w [3E,E], proj_w [E,E], lin1_w [mlp,E], lin2_w [E,mlp], qkv_bias_op [3E],
//        rel [2*128*hd] (windowed: window table [64,hd] then zeros; global: rh_rev [128,hd], rw_rev [128,hd]) }
//        | neck0_w [C,E] | neck2_w [C, 9*C] with columns ordered (ky, kx, c)
//   w32: pos_embed [g*g, E] | patch_b [E] | per block { n1w, n1b [E], qkv_b [3E], proj_b [E], n2w, n2b [E], lin1_b [mlp],
//        lin2_b [E] } | neck_ln1 w, b [C] | neck_ln2 w, b [C]
//   ln_fold: qkv_w / lin1_w hold gamma o W, qkv_b / lin1_b hold beta . W^T + b, n1w .. n2b are unused, and every block
//        gains { qkv_colsum [3E], lin1_colsum [mlp] } (row sums of the ROUNDED gamma o W) at its end of w32
size_t samk_encoder_w16_elems(const SamEncoderShape& s) {
  const size_t E = s.embed_dim, M = s.mlp_dim, C = s.out_chans, hd = E / s.heads;
  size_t n = align8(E * 3 * s.patch * s.patch);
  n += s.depth * (align8(3 * E * E) + align8(E * E) + align8(M * E) + align8(E * M) + align8(3 * E) + align8(2 * 128 * hd));
  n += align8(C * E) + align8(C * 9 * C);
  return n;
}
size_t samk_encoder_w32_elems(const SamEncoderShape& s) {
  const size_t E = s.embed_dim, M = s.mlp_dim, C = s.out_chans, g = s.img / s.patch;
  size_t n = align8(g * g * E) + align8(E);
  n += s.depth * (6 * align8(E) + align8(3 * E) + align8(M));
  if (s.ln_fold) n += s.depth * (align8(3 * E) + align8(M));
  n += 4 * align8(C);
  return n;
}
size_t samk_encoder_workspace_bytes(const SamEncoderShape& s, int B) {
  const size_t E = s.embed_dim, C = s.out_chans, g = s.img / s.patch, M = static_cast<size_t>(B) * g * g;
  size_t big = 3 * E;
  if (static_cast<size_t>(s.mlp_dim) > big) big = s.mlp_dim;
  if (static_cast<size_t>(3 * s.patch * s.patch) > big) big = 3 * s.patch * s.patch;
  if (9 * C > big) big = 9 * C;
  size_t bytes = 0;
  bytes += M * E * 4;        // x (fp32 residual stream)
  bytes += M * E * 2;        // a16
  bytes += M * big * 2;      // big16: qkv | mlp hidden | patch matrix | 3x3 im2col
  bytes += M * C * 4;        // n32
  bytes += M * C * 2;        // n16
  if (s.ln_fold) {
    bytes += M * E * 2;            // xb: operand-format copy of the residual stream
    bytes += M * (E / 128) * 8;    // per-row (mean, M2) of each 128-column slice
  }
  return bytes + 1024;
}

int samk_encoder_forward(const SamEncoderShape& s, const void* w16v, const float* w32, const void* images, int in_fmt,
                         int B, void* out, int out_fmt, void* workspace, size_t workspace_bytes, cudaStream_t st) {
  const int E = s.embed_dim, C = s.out_chans, g = s.img / s.patch, mlp = s.mlp_dim, fmt = s.fmt;
  SAM_REQUIRE(B > 0, "image encoder: empty batch");
  SAM_REQUIRE(fmt == 0 || fmt == 1, "image encoder: operand format must be fp16 (0) or bf16 (1)");
  SAM_REQUIRE(s.img == 1024 && s.patch == 16 && s.window == 14, "image encoder: only 1024/16 images with 14x14 windows");
  SAM_REQUIRE(E % s.heads == 0 && (E / s.heads == 80 || E / s.heads == 64),
              "image encoder: head_dim must be 80 (ViT-H) or 64 (ViT-L / ViT-B), got %d/%d", E, s.heads);
  SAM_REQUIRE(s.depth <= 64, "image encoder: depth %d > 64", s.depth);
  SAM_REQUIRE(!s.ln_fold || E % 256 == 0, "image encoder: ln_fold needs embed_dim %% 256 == 0, got %d", E);
  SAM_REQUIRE(workspace_bytes >= samk_encoder_workspace_bytes(s, B), "image encoder: workspace too small");
  SAM_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 1023) == 0, "image encoder: workspace must be 1024-byte aligned");
  const size_t M = static_cast<size_t>(B) * g * g;
  SAM_REQUIRE(M * 5120 < (1ull << 31), "image encoder: batch %d too large for one call (chunk it)", B);
  const int hd = E / s.heads;
  const uint16_t* w16 = static_cast<const uint16_t*>(w16v);

  // ---- carve weights
  size_t o16 = 0, o32 = 0;
  auto t16 = [&](size_t n) { const uint16_t* p = w16 + o16; o16 += align8(n); return p; };
  auto t32 = [&](size_t n) { const float* p = w32 + o32; o32 += align8(n); return p; };
  const uint16_t* patch_w = t16(static_cast<size_t>(E) * 3 * s.patch * s.patch);
  const float* pos = t32(static_cast<size_t>(g) * g * E);
  const float* patch_b = t32(E);

  // ---- carve workspace
  size_t big = 3 * E;
  if (static_cast<size_t>(mlp) > big) big = mlp;
  if (static_cast<size_t>(3 * s.patch * s.patch) > big) big = 3 * s.patch * s.patch;
  if (static_cast<size_t>(9 * C) > big) big = 9 * C;
  uint8_t* wp = static_cast<uint8_t*>(workspace);
  float* x = reinterpret_cast<float*>(wp); wp += M * E * 4;
  uint16_t* a16 = reinterpret_cast<uint16_t*>(wp); wp += M * E * 2;
  uint16_t* big16 = reinterpret_cast<uint16_t*>(wp); wp += M * big * 2;
  float* n32 = reinterpret_cast<float*>(wp); wp += M * C * 4;
  uint16_t* n16 = reinterpret_cast<uint16_t*>(wp); wp += M * C * 2;
  uint16_t* xb = nullptr;
  void* stats = nullptr;
  const int parts = E / 128;
  if (s.ln_fold) {
    xb = reinterpret_cast<uint16_t*>(wp); wp += M * E * 2;
    stats = wp;
  }

  auto gemm = [&](const void* A, int lda, const void* W, int K, int N, void* o, int ldo, int ofmt, const float* bias,
                  int act, const float* res, int ldr, int res_mod) {
    GemmEpilogue ep{o, ldo, ofmt, bias, act, res, ldr, res_mod};
    return samk_gemm(A, lda, W, K, static_cast<int>(M), N, K, fmt, ep, st);
  };
  // LayerNorm-folding GEMMs: consumer (A = xb, W = gamma o W) and producer (in-place residual + xb + stats)
  auto gemm_ln = [&](const void* W, int N, void* o, const float* bias_fold, const float* colsum, int act) {
    GemmEpilogue ep{o, N, fmt, bias_fold, act, nullptr, 0, 0};
    ep.ln_stats = stats;
    ep.ln_parts = parts;
    ep.ln_colsum = colsum;
    ep.ln_c = E;
    ep.ln_eps = 1e-6f;
    return samk_gemm(xb, E, W, E, static_cast<int>(M), N, E, fmt, ep, st);
  };
  auto gemm_res = [&](const void* A, int K, const void* W, const float* bias) {
    GemmEpilogue ep{x, E, SAM_F32, bias, 0, x, E, static_cast<int>(M)};
    if (s.ln_fold) {
      ep.xb = xb;
      ep.ldxb = E;
      ep.stats_out = stats;
    }
    return samk_gemm(A, K, W, K, static_cast<int>(M), E, K, fmt, ep, st);
  };
#define RUN(expr)            \
  do {                       \
    if (int rc_ = (expr)) return rc_; \
  } while (0)

  // ---- patch embedding + pos_embed (image_encoder.py:112-113, :418-426)
  const int pk = 3 * s.patch * s.patch;
  RUN(samk_patch_im2col(images, in_fmt, big16, fmt, B, s.img, s.patch, st));
  if (s.ln_fold) {
    // plain fp32 store epilogue (2-CTA kernel); pos_embed is added by the pass that makes xb / stats anyway
    RUN(gemm(big16, pk, patch_w, pk, E, x, E, SAM_F32, patch_b, 0, nullptr, 0, 0));
    RUN(samk_cast_stats(x, E, xb, E, fmt, stats, static_cast<int>(M), E, pos, g * g, st));
  } else {
    RUN(gemm(big16, pk, patch_w, pk, E, x, E, SAM_F32, patch_b, 0, pos, E, g * g));
  }

  // ---- transformer blocks
  for (int i = 0; i < s.depth; ++i) {
    BlockW16 b16;
    BlockW32 b32;
    b16.qkv_w = t16(3ull * E * E);
    b16.proj_w = t16(static_cast<size_t>(E) * E);
    b16.lin1_w = t16(static_cast<size_t>(mlp) * E);
    b16.lin2_w = t16(static_cast<size_t>(E) * mlp);
    b16.qkv_b_op = t16(3 * E);
    b16.rel = t16(2 * 128 * hd);
    b32.n1w = t32(E); b32.n1b = t32(E);
    b32.qkv_b = t32(3 * E);
    b32.proj_b = t32(E);
    b32.n2w = t32(E); b32.n2b = t32(E);
    b32.lin1_b = t32(mlp);
    b32.lin2_b = t32(E);
    const bool is_global = (s.global_mask >> i) & 1ull;

    const float* qkv_cs = s.ln_fold ? t32(3 * E) : nullptr;
    const float* lin1_cs = s.ln_fold ? t32(mlp) : nullptr;

    if (s.ln_fold)
      RUN(gemm_ln(b16.qkv_w, 3 * E, big16, b32.qkv_b, qkv_cs, 0));
    else {
      RUN(samk_layernorm_rows(x, E, nullptr, 0, b32.n1w, b32.n1b, 1e-6f, a16, E, fmt, static_cast<int>(M), E, 1, st));
      RUN(gemm(a16, E, b16.qkv_w, E, 3 * E, big16, 3 * E, fmt, b32.qkv_b, 0, nullptr, 0, 0));
    }
    if (is_global)
      RUN(samk_attn_global(big16, b16.rel, b16.rel + 128 * hd, a16, B, E, s.heads, fmt, st));
    else
      RUN(samk_attn_window(big16, b16.qkv_b_op, b16.rel, a16, B, E, s.heads, fmt, st));
    RUN(gemm_res(a16, E, b16.proj_w, b32.proj_b));
    if (s.ln_fold)
      RUN(gemm_ln(b16.lin1_w, mlp, big16, b32.lin1_b, lin1_cs, 1));
    else {
      RUN(samk_layernorm_rows(x, E, nullptr, 0, b32.n2w, b32.n2b, 1e-6f, a16, E, fmt, static_cast<int>(M), E, 1, st));
      RUN(gemm(a16, E, b16.lin1_w, E, mlp, big16, mlp, fmt, b32.lin1_b, 1, nullptr, 0, 0));
    }
    RUN(gemm_res(big16, mlp, b16.lin2_w, b32.lin2_b));
    if (s.tap_block == i && s.tap_out) {
      // test hook: copy the residual stream after block i (fp32 [M, E])
      if (cudaMemcpyAsync(s.tap_out, x, M * E * 4, cudaMemcpyDeviceToDevice, st) != cudaSuccess)
        return samhost::set_error(2, "image encoder: tap copy failed");
    }
  }

  // ---- neck (image_encoder.py:92-108)
  const uint16_t* neck0 = t16(static_cast<size_t>(C) * E);
  const uint16_t* neck2 = t16(static_cast<size_t>(C) * 9 * C);
  const float* ln1w = t32(C); const float* ln1b = t32(C);
  const float* ln2w = t32(C); const float* ln2b = t32(C);
  const uint16_t* neck_in = xb;   // ln_fold: the last lin2 epilogue already wrote the operand-format copy of x
  if (!s.ln_fold) {
    RUN(samk_layernorm_rows(x, E, nullptr, 0, nullptr, nullptr, 0.f, a16, E, fmt, static_cast<int>(M), E, 0, st));
    neck_in = a16;
  }
  RUN(gemm(neck_in, E, neck0, E, C, n32, C, SAM_F32, nullptr, 0, nullptr, 0, 0));
  RUN(samk_layernorm_rows(n32, C, nullptr, 0, ln1w, ln1b, 1e-6f, n16, C, fmt, static_cast<int>(M), C, 1, st));
  RUN(samk_im2col3x3(n16, big16, B, g, C, st));
  RUN(gemm(big16, 9 * C, neck2, 9 * C, C, n32, C, SAM_F32, nullptr, 0, nullptr, 0, 0));
  RUN(samk_ln_nhwc_to_nchw(n32, ln2w, ln2b, 1e-6f, out, out_fmt, B, g * g, C, st));
#undef RUN
  return 0;
}
