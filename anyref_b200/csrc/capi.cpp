// extern "C" boundary of libanyref_sam.so (declared in include/anyref_sam.h).
#include "../../include/anyref_sam.h"

#include "host_common.h"
#include "kernels.h"

static inline cudaStream_t S(void* s) { return static_cast<cudaStream_t>(s); }

extern "C" {

const char* sam_last_error(void) { return samhost::last_error(); }
int sam_abi_version(void) { return 4; }

int sam_gemm(const void* A, int lda, const void* W, int ldw, int M, int N, int K, int fmt, void* out, int ldo,
             int out_fmt, const float* bias, int act, const float* res, int ldr, int res_mod, void* stream) {
  GemmEpilogue ep;
  ep.out = out;
  ep.ldo = ldo;
  ep.out_fmt = out_fmt;
  ep.bias = bias;
  ep.act = act;
  ep.res = res;
  ep.ldr = ldr;
  ep.res_mod = res_mod;
  return samk_gemm(A, lda, W, ldw, M, N, K, fmt, ep, S(stream));
}

int sam_gemm_residual_ln(const void* A, int lda, const void* W, int ldw, int M, int N, int K, int fmt, float* x, int ldx,
                         const float* bias, void* xb, int ldxb, void* stats, void* stream) {
  if (!A || !W || !x || !xb || !stats) return samhost::set_error(1, "sam_gemm_residual_ln: NULL argument");
  GemmEpilogue ep{x, ldx, SAM_F32, bias, 0, x, ldx, M};
  ep.xb = xb;
  ep.ldxb = ldxb;
  ep.stats_out = stats;
  return samk_gemm(A, lda, W, ldw, M, N, K, fmt, ep, S(stream));
}
int sam_cast_stats(const float* x, int ldx, void* xb, int ldxb, int fmt, void* stats, int M, int C, void* stream) {
  if (!x || !xb || !stats) return samhost::set_error(1, "sam_cast_stats: NULL argument");
  return samk_cast_stats(const_cast<float*>(x), ldx, xb, ldxb, fmt, stats, M, C, nullptr, 0, S(stream));
}
int sam_gemm_ln(const void* xb, int lda, const void* Wg, int ldw, int M, int N, int K, int fmt, void* out, int ldo,
                int out_fmt, const float* bias_fold, const float* colsum, const void* stats, int parts, float eps,
                int act, void* stream) {
  if (!xb || !Wg || !out || !bias_fold || !colsum || !stats) return samhost::set_error(1, "sam_gemm_ln: NULL argument");
  if (out_fmt != 0 && out_fmt != 1) return samhost::set_error(1, "sam_gemm_ln: output must be fp16/bf16");
  GemmEpilogue ep{out, ldo, out_fmt, bias_fold, act, nullptr, 0, 0};
  ep.ln_stats = stats;
  ep.ln_parts = parts;
  ep.ln_colsum = colsum;
  ep.ln_c = K;
  ep.ln_eps = eps;
  return samk_gemm(xb, lda, Wg, ldw, M, N, K, fmt, ep, S(stream));
}

int sam_umma_probe(const void* A, const void* B, float* D, int N, int K, int fmt, int a_mode, int b_mode, int a_lbo,
                   int a_sbo, int b_lbo, int b_sbo, void* stream) {
  UmmaProbe p;
  p.N = N;
  p.K = K;
  p.fmt = fmt;
  p.a_mode = a_mode;
  p.b_mode = b_mode;
  p.a_lbo = a_lbo;
  p.a_sbo = a_sbo;
  p.b_lbo = b_lbo;
  p.b_sbo = b_sbo;
  p.a_kstep = p.b_kstep = 0;
  return samk_umma_probe(A, B, D, p, S(stream));
}

int sam_layernorm(const float* x, int ldx, const float* res, int ldr, const float* gamma, const float* beta, float eps,
                  void* out, int ldo, int out_fmt, int M, int C, int normalize, void* stream) {
  return samk_layernorm_rows(x, ldx, res, ldr, gamma, beta, eps, out, ldo, out_fmt, M, C, normalize, S(stream));
}
int sam_patch_im2col(const void* img, int in_fmt, void* out, int out_fmt, int B, int Sz, int p, void* stream) {
  return samk_patch_im2col(img, in_fmt, out, out_fmt, B, Sz, p, S(stream));
}
int sam_im2col3x3(const void* in, void* out, int B, int g, int C, void* stream) {
  return samk_im2col3x3(in, out, B, g, C, S(stream));
}
int sam_ln_nhwc_to_nchw(const float* x, const float* gamma, const float* beta, float eps, void* out, int out_fmt,
                        int B, int tokens_per_img, int C, void* stream) {
  return samk_ln_nhwc_to_nchw(x, gamma, beta, eps, out, out_fmt, B, tokens_per_img, C, S(stream));
}
int sam_attn_window(const void* qkv, const void* bias_op, const void* rel_tab, void* out, int B, int E, int heads,
                    int fmt, void* stream) {
  return samk_attn_window(qkv, bias_op, rel_tab, out, B, E, heads, fmt, S(stream));
}
int sam_attn_global(const void* qkv, const void* rh_rev, const void* rw_rev, void* out, int B, int E, int heads,
                    int fmt, void* stream) {
  return samk_attn_global(qkv, rh_rev, rw_rev, out, B, E, heads, fmt, S(stream));
}

size_t sam_encoder_w16_elems(const SamEncoderShape* s) { return samk_encoder_w16_elems(*s); }
size_t sam_encoder_w32_elems(const SamEncoderShape* s) { return samk_encoder_w32_elems(*s); }
size_t sam_encoder_workspace_bytes(const SamEncoderShape* s, int B) { return samk_encoder_workspace_bytes(*s, B); }
int sam_encoder_forward(const SamEncoderShape* s, const void* w16, const float* w32, const void* images, int in_fmt,
                        int B, void* out, int out_fmt, void* workspace, size_t workspace_bytes, void* stream) {
  if (!s || !w16 || !w32 || !images || !out || !workspace) return samhost::set_error(1, "sam_encoder_forward: NULL argument");
  return samk_encoder_forward(*s, w16, w32, images, in_fmt, B, out, out_fmt, workspace, workspace_bytes, S(stream));
}
size_t sam_decoder_weight_elems(const SamDecoderShape* s) { return samk_decoder_weight_elems(*s); }
size_t sam_decoder_workspace_bytes(const SamDecoderShape* s, int n_images, int n, int k) {
  return samk_decoder_workspace_bytes(*s, n_images, n, k);
}
size_t sam_decoder_derived_bytes(const SamDecoderShape* s) { return samk_decoder_derived_bytes(*s); }
int sam_decoder_prepare(const SamDecoderShape* s, const float* weights, const void* image_pe, int pe_fmt, void* derived,
                        void* stream) {
  if (!s || !weights || !image_pe || !derived) return samhost::set_error(1, "sam_decoder_prepare: NULL argument");
  return samk_decoder_prepare(*s, weights, image_pe, pe_fmt, derived, S(stream));
}
int sam_decoder_forward(const SamDecoderShape* s, const float* weights, const void* derived, const void* image_embeddings,
                        int emb_fmt, int n_images, const int* img_index, const void* sparse, int sparse_fmt, int n, int k,
                        const void* dense_vec, const void* dense_full, int dense_fmt, void* masks, void* iou, int out_fmt,
                        void* workspace, size_t workspace_bytes, void* stream) {
  if (!s || !weights || !derived || !image_embeddings || !masks || !iou || !workspace)
    return samhost::set_error(1, "sam_decoder_forward: NULL argument");
  if (k > 0 && !sparse) return samhost::set_error(1, "sam_decoder_forward: sparse embeddings missing");
  return samk_decoder_forward(*s, weights, derived, image_embeddings, emb_fmt, n_images, img_index, sparse, sparse_fmt, n, k,
                              dense_vec, dense_full, dense_fmt, masks, iou, out_fmt, workspace, workspace_bytes, S(stream));
}
size_t sam_decoder_train_workspace_bytes(const SamDecoderShape* s, int n, int k) {
  return s ? samk_decoder_train_workspace_bytes(*s, n, k) : 0;
}
int sam_decoder_train_forward(const SamDecoderShape* s, const float* weights, const void* image_embeddings, int emb_fmt, int n_images,
                              const int* img_index, const float* sparse, int n, int k, const void* dense_vec, const void* dense_full,
                              int dense_fmt, const void* image_pe, int pe_fmt, float* masks, float* iou, void* workspace,
                              size_t workspace_bytes, void** tape, void* stream) {
  if (!s || !tape) return samhost::set_error(1, "sam_decoder_train_forward: NULL argument");
  return samk_decoder_train_forward(*s, weights, image_embeddings, emb_fmt, n_images, img_index, sparse, n, k, dense_vec, dense_full,
                                    dense_fmt, image_pe, pe_fmt, masks, iou, workspace, workspace_bytes, tape, S(stream));
}
int sam_decoder_backward(void* tape, const float* d_masks, int mask_lo, int mask_hi, const float* d_iou, float* d_weights,
                         float* d_sparse, void* stream) {
  return samk_decoder_backward(tape, d_masks, mask_lo, mask_hi, d_iou, d_weights, d_sparse, S(stream));
}
void sam_decoder_tape_free(void* tape) { samk_decoder_tape_free(tape); }
size_t sam_linear_f32_scratch_bytes(int M, int N, int K) { return samk_linear_f32_scratch_bytes(M, N, K); }
int sam_linear_f32_forward(const float* X, const float* W, const float* b, float* Y, int M, int N, int K, int relu, void* scratch,
                           size_t scratch_bytes, void* stream) {
  return samk_linear_f32_forward(X, W, b, Y, M, N, K, relu, scratch, scratch_bytes, S(stream));
}
int sam_linear_f32_backward(float* dY, const float* relu_y, const float* X, const float* W, float* dX, float* dW, float* db, int M,
                            int N, int K, void* scratch, size_t scratch_bytes, void* stream) {
  return samk_linear_f32_backward(dY, relu_y, X, W, dX, dW, db, M, N, K, scratch, scratch_bytes, S(stream));
}
int sam_postprocess_masks_backward(const float* d_logits, int num_masks, int L, int Sz, int h_in, int w_in, int H, int W, float* tmp,
                                   float* d_low, void* stream) {
  return samk_postprocess_backward(d_logits, num_masks, L, Sz, h_in, w_in, H, W, tmp, d_low, S(stream));
}
int sam_postprocess_masks(const void* low, int low_fmt, int num_masks, int L, int Sz, int h_in, int w_in, int H, int W,
                          float* logits, unsigned char* binary, float threshold, void* stream) {
  return samk_postprocess(low, low_fmt, num_masks, L, Sz, h_in, w_in, H, W, logits, binary, threshold, S(stream));
}

int sam_postprocess_masks_iou(const void* low, int low_fmt, int num_masks, int L, int Sz, int h_in, int w_in, int H, int W,
                              float* logits, unsigned char* binary, float threshold, const unsigned char* target,
                              int* counts, void* stream) {
  return samk_postprocess_iou(low, low_fmt, num_masks, L, Sz, h_in, w_in, H, W, logits, binary, nullptr, threshold, target,
                              counts, S(stream));
}
int sam_postprocess_masks_packed(const void* low, int low_fmt, int num_masks, int L, int Sz, int h_in, int w_in, int H, int W,
                                 unsigned char* packed, float threshold, const unsigned char* target, int* counts,
                                 void* stream) {
  if (!low || !packed) return samhost::set_error(1, "sam_postprocess_masks_packed: NULL argument");
  return samk_postprocess_iou(low, low_fmt, num_masks, L, Sz, h_in, w_in, H, W, nullptr, nullptr, packed, threshold, target,
                              counts, S(stream));
}
int sam_iou_finalize(const int* counts, int n, double* stats, void* stream) {
  return samk_iou_finalize(counts, n, stats, S(stream));
}

int sam_dense_pe(const float* gauss, void* out, int out_fmt, int C, int g, void* stream) {
  return samk_dense_pe(gauss, out, out_fmt, C, g, S(stream));
}

int sam_prompt_sparse(const float* coords, const float* labels, const float* gauss, const float* table, float* out, int n,
                      int n_in, int pad, int mode, int C, int img_h, int img_w, int ld_tokens, int tok0, void* stream) {
  if (!gauss || !table || !out || (n_in > 0 && !coords))
    return samhost::set_error(1, "sam_prompt_sparse: NULL argument");
  return samk_prompt_sparse(coords, labels, gauss, table, out, n, n_in, pad, mode, C, img_h, img_w, ld_tokens, tok0,
                            S(stream));
}
size_t sam_prompt_mask_blob_elems(int mask_in_chans, int C) { return samk_prompt_mask_blob_elems(mask_in_chans, C); }
int sam_prompt_mask_embed(const void* masks, int in_fmt, const float* blob, int mask_in_chans, void* out, int out_fmt,
                          int n, int g, int C, void* stream) {
  if (!masks || !blob || !out) return samhost::set_error(1, "sam_prompt_mask_embed: NULL argument");
  return samk_prompt_mask_embed(masks, in_fmt, blob, mask_in_chans, out, out_fmt, n, g, C, S(stream));
}
int sam_resize_u8(const unsigned char* in, int H, int W, int C, unsigned char* tmp, unsigned char* out, int new_h, int new_w,
                  const int* xbounds, const int* xcoeff, int xk, const int* ybounds, const int* ycoeff, int yk,
                  void* stream) {
  if (!in || !out) return samhost::set_error(1, "sam_resize_u8: NULL image");
  return samk_resize_u8(in, H, W, C, tmp, out, new_h, new_w, xbounds, xcoeff, xk, ybounds, ycoeff, yk, S(stream));
}
int sam_preprocess(const void* img, int in_fmt, void* out, int out_fmt, int B, int h, int w, int Sz, const float* mean,
                   const float* std, void* stream) {
  if (!img || !out || !mean || !std) return samhost::set_error(1, "sam_preprocess: NULL argument");
  return samk_preprocess(img, in_fmt, out, out_fmt, B, h, w, Sz, mean, std, S(stream));
}

void sam_gemm_set_tile_split(int mode) { samk_gemm2_set_tile_split(mode); }
int sam_gemm_schedule(int num_tiles, int num_pairs, int pair, int split, int* out, int cap) {
  return samk_gemm2_schedule(num_tiles, num_pairs, pair, split, out, cap);
}
long long sam_launch_count(void) { return samhost::launch_count(); }
void sam_profile_enable(int on) { samhost::profile_enable(on); }
void sam_profile_reset(void) { samhost::profile_reset(); }
int sam_profile_collect(void) { return samhost::profile_collect(); }
void sam_profile_get(int cls, double* ms, long long* launches, double* flops, double* bytes) {
  samhost::profile_get(cls, ms, launches, flops, bytes);
}

}  // extern "C"
