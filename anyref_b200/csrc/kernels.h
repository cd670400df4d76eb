// Internal C++ interface between the kernel translation units and the C-ABI layer (capi.cpp).
// Every function enqueues work on `stream`, allocates nothing, never synchronises, and returns 0 or an error code
// whose text is available through samhost::last_error().
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/anyref_sam.h"  // SamEncoderShape / SamDecoderShape

// Operand / storage formats used across the library.
//   0 = fp16, 1 = bf16 (tensor-core operand formats; identical to the tcgen05 idesc encoding), 2 = fp32
enum SamFmt : int { SAM_F16 = 0, SAM_BF16 = 1, SAM_F32 = 2 };

struct GemmEpilogue {
  void* out;         // [M, ldo] in out_fmt
  int ldo;           // elements
  int out_fmt;       // SamFmt
  const float* bias; // [N] or null
  int act;           // 0 none, 1 exact-erf GELU, 2 ReLU
  const float* res;  // fp32 residual source or null; row used = (row % res_mod); may alias `out` (fp32, in place)
  int ldr;
  int res_mod;
  // LayerNorm folding (gemm2.cu), all optional (zero = off):
  //   consumer: A = 16-bit copy of the un-normalised rows, W = gamma o W; the epilogue computes
  //             rstd * (acc - mean * ln_colsum[n]) + bias[n]  with mean / rstd from the partial sums ln_stats [M, ln_parts]
  //             ((mean, sum of squared deviations) per slice of the ln_c-wide row, combined with Chan's formula); bias must hold beta.W^T + b
  const void* ln_stats = nullptr;
  int ln_parts = 0;
  const float* ln_colsum = nullptr;
  int ln_c = 0;
  float ln_eps = 0.f;
  //   producer (in-place fp32 residual only): also writes xb [M, ldxb] = round(out) in the operand format and the
  //             partial sums stats_out [M, N / 128] of the new rows
  void* xb = nullptr;
  int ldxb = 0;
  void* stats_out = nullptr;
  // Block-diagonal mode (gemm2.cu only; fp32 store): A is [S * diag_mt * 256, K] and W [S * diag_wrows, K]; the output
  // tile of A's row block m_blk multiplies the W rows of chunk s = m_blk / diag_mt, i.e. S independent products
  // C_s = A_s . W_s^T stacked along the rows of C -- the split-K partials of a weight gradient dY^T . X
  int diag_mt = 0;
  int diag_wrows = 0;
  int diag_wtotal = 0;   // S * diag_wrows
};

int samk_gemm(const void* A, int lda, const void* W, int ldw, int M, int N, int K, int fmt, const GemmEpilogue& ep,
              cudaStream_t stream);

// 2-CTA variant (gemm2.cu): returns -1 when it does not implement the requested epilogue (caller falls back).
int samk_gemm2(const void* A, int lda, const void* W, int ldw, int M, int N, int K, int fmt, const GemmEpilogue& ep,
               cudaStream_t stream);

// Half-tile items in the last round of samk_gemm2 (gemm2.cu, Sched): -1 policy, 0 off, 1 whenever possible.
void samk_gemm2_set_tile_split(int mode);
int samk_gemm2_schedule(int num_tiles, int num_clusters, int cluster_id, int nsplit, int* out, int cap);

// UMMA layout probe (test-only kernel, see probe.cu).
struct UmmaProbe {
  int N;          // MMA N (multiple of 16, <= 256); M is fixed at 128
  int K;          // multiple of 16, <= 256
  int fmt;        // 0 fp16, 1 bf16
  int a_mode;     // smem fill + descriptor mode for A (see probe.cu)
  int b_mode;     // same for B
  int a_lbo, a_sbo, a_kstep;  // descriptor byte offsets and per-16-K start-address advance (bytes)
  int b_lbo, b_sbo, b_kstep;
};
int samk_umma_probe(const void* A, const void* B, float* D, const UmmaProbe& p, cudaStream_t stream);

// 14x14 windowed attention with fused decomposed rel-pos bias (attn_window.cu).
//   qkv [B*4096, 3E] op-format; bias_op [3E] op-format (qkv bias, used for padded window tokens);
//   rel_tab [64, 80] op-format: rows 0..26 rel_pos_h, rows 32..58 rel_pos_w, rest zero; out [B*4096, E] op-format.
int samk_attn_window(const void* qkv, const void* bias_op, const void* rel_tab, void* out, int B, int E, int heads,
                     int fmt, cudaStream_t stream);
int samk_attn_window3(const void* qkv, const void* bias_op, const void* rel_tab, void* out, int B, int E, int heads,
                      int fmt, cudaStream_t stream);
int samk_attn_window4(const void* qkv, const void* bias_op, const void* rel_tab, void* out, int B, int E, int heads,
                      int fmt, cudaStream_t stream);
// Global 64x64 attention with fused rel-pos bias (attn_global.cu).
//   rh_rev / rw_rev [128, 80] op-format: row j = rel_pos_{h,w}[126 - j] for j < 127, row 127 zero.
int samk_attn_global(const void* qkv, const void* rh_rev, const void* rw_rev, void* out, int B, int E, int heads,
                     int fmt, cudaStream_t stream);

int samk_attn_global2(const void* qkv, const void* rh_rev, const void* rw_rev, void* out, int B, int E, int heads,
                      int fmt, cudaStream_t stream);
int samk_attn_global3(const void* qkv, const void* rh_rev, const void* rw_rev, void* out, int B, int E, int heads,
                      int fmt, cudaStream_t stream);

// Memory-bound glue (glue.cu).  All fp32 activations are token-major rows.
//   layernorm_rows : out[row] = LN(x[row] (+ res[row])) * gamma + beta  (normalize == 0: plain dtype cast)
int samk_layernorm_rows(const float* x, int ldx, const float* res, int ldr, const float* gamma, const float* beta,
                        float eps, void* out, int ldo, int out_fmt, int M, int C, int normalize, cudaStream_t stream);
//   cast_stats : xb = round(x) (operand format) + per-row (mean, M2) of every 128-column slice [M, C/128]
//                (pos != null: x[row] += pos[row % pos_mod] first, written back to x)
int samk_cast_stats(float* x, int ldx, void* xb, int ldxb, int fmt, void* stats, int M, int C, const float* pos,
                    int pos_mod, cudaStream_t stream);
int samk_patch_im2col(const void* img, int in_fmt, void* out, int out_fmt, int B, int S, int p, cudaStream_t stream);
int samk_im2col3x3(const void* in, void* out, int B, int g, int C, cudaStream_t stream);
int samk_ln_nhwc_to_nchw(const float* x, const float* gamma, const float* beta, float eps, void* out, int out_fmt,
                         int B, int tokens_per_img, int C, cudaStream_t stream);

// Whole-module drivers (encoder.cpp, decoder.cu, postprocess.cu); see include/anyref_sam.h for the semantics.
size_t samk_encoder_w16_elems(const SamEncoderShape& s);
size_t samk_encoder_w32_elems(const SamEncoderShape& s);
size_t samk_encoder_workspace_bytes(const SamEncoderShape& s, int B);
int samk_encoder_forward(const SamEncoderShape& s, const void* w16, const float* w32, const void* images, int in_fmt,
                         int B, void* out, int out_fmt, void* workspace, size_t workspace_bytes, cudaStream_t st);
size_t samk_decoder_weight_elems(const SamDecoderShape& s);
size_t samk_decoder_workspace_bytes(const SamDecoderShape& s, int n_images, int n, int k);
size_t samk_decoder_derived_bytes(const SamDecoderShape& s);
int samk_decoder_prepare(const SamDecoderShape& s, const float* blob, const void* image_pe, int pe_fmt, void* derived,
                         cudaStream_t st);
int samk_decoder_forward(const SamDecoderShape& s, const float* blob, const void* derived, const void* image_embeddings, int emb_fmt,
                         int n_images, const int* img_index, const void* sparse, int sparse_fmt, int n, int k,
                         const void* dense_vec, const void* dense_full, int dense_fmt, void* masks, void* iou, int out_fmt,
                         void* workspace, size_t workspace_bytes, cudaStream_t st);
// Training path of the mask decoder (decoder_train.cu): forward that keeps its intermediates on a tape + backward
size_t samk_decoder_train_workspace_bytes(const SamDecoderShape& s, int n, int k);
int samk_decoder_train_forward(const SamDecoderShape& s, const float* blob, const void* image_embeddings, int emb_fmt, int n_images,
                               const int* img_index, const float* sparse, int n, int k, const void* dense_vec, const void* dense_full,
                               int dense_fmt, const void* image_pe, int pe_fmt, float* masks, float* iou, void* workspace,
                               size_t workspace_bytes, void** tape_out, cudaStream_t st);
int samk_decoder_backward(void* tape, const float* d_masks, int mask_lo, int mask_hi, const float* d_iou, float* d_weights,
                          float* d_sparse, cudaStream_t st);
void samk_decoder_tape_free(void* tape);
// the same composition without a tape: inference for more tokens per prompt than decoder.cu's fused kernels hold
size_t samk_decoder_generic_workspace_bytes(const SamDecoderShape& s, int n, int k);
int samk_decoder_forward_generic(const SamDecoderShape& s, const float* blob, const float* pe_tokens, const void* image_embeddings,
                                 int emb_fmt, int n_images, const int* img_index, const void* sparse, int sparse_fmt, int n, int k,
                                 const void* dense_vec, const void* dense_full, int dense_fmt, void* masks, void* iou, int out_fmt,
                                 void* workspace, size_t workspace_bytes, cudaStream_t st);
size_t samk_linear_f32_scratch_bytes(int M, int N, int K);
int samk_linear_f32_forward(const float* X, const float* W, const float* b, float* Y, int M, int N, int K, int relu_act, void* scratch,
                            size_t scratch_bytes, cudaStream_t st);
int samk_linear_f32_backward(float* dY, const float* relu_y, const float* X, const float* W, float* dX, float* dW, float* db, int M, int N, int K,
                             void* scratch, size_t scratch_bytes, cudaStream_t st);
int samk_postprocess_backward(const float* d_out, int maps, int low, int img, int in_h, int in_w, int out_h, int out_w, float* tmp,
                              float* d_low, cudaStream_t st);

int samk_postprocess(const void* low, int low_fmt, int num_masks, int L, int S, int h_in, int w_in, int H, int W,
                     float* logits, uint8_t* binary, float threshold, cudaStream_t stream);
int samk_postprocess_iou(const void* low, int low_fmt, int num_masks, int L, int S, int h_in, int w_in, int H, int W,
                         float* logits, uint8_t* binary, uint8_t* packed, float threshold, const uint8_t* target,
                         int* counts, cudaStream_t stream);
int samk_iou_finalize(const int* counts, int n, double* stats, cudaStream_t stream);
int samk_dense_pe(const float* gauss, void* out, int out_fmt, int C, int g, cudaStream_t stream);

// PromptEncoder point / box / mask prompts and Sam.preprocess (prompt.cu).
int samk_prompt_sparse(const float* coords, const float* labels, const float* gauss, const float* table, float* out,
                       int n, int n_in, int pad, int mode, int C, int img_h, int img_w, int ld_tokens, int tok0,
                       cudaStream_t stream);
size_t samk_prompt_mask_blob_elems(int mask_in_chans, int C);
int samk_prompt_mask_embed(const void* masks, int in_fmt, const float* blob, int mask_in_chans, void* out, int out_fmt,
                           int n, int g, int C, cudaStream_t stream);
int samk_preprocess(const void* img, int in_fmt, void* out, int out_fmt, int B, int h, int w, int S, const float* mean,
                    const float* std, cudaStream_t stream);
// ResizeLongestSide.apply_image: PIL-exact separable bilinear resize of an HWC uint8 image (prompt.cu).
int samk_resize_u8(const uint8_t* in, int H, int W, int C, uint8_t* tmp, uint8_t* out, int new_h, int new_w,
                   const int* xbounds, const int* xcoeff, int xk, const int* ybounds, const int* ycoeff, int yk,
                   cudaStream_t stream);
