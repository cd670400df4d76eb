/*
 * anyref_sam.h -- C ABI of libanyref_sam.so, the B200 (sm_100a) implementation of AnyRef's SAM ViT-H grounding path.
 *
 * The reference (jwh97nn/AnyRef) has no FFI: its boundary for this path is the Python nn.Module tree
 * `visual_model` (model/anyref.py:106) whose forwards run torch library kernels.  Each entry point below replaces
 * the library calls behind one reference function; the reference file:line is cited on every declaration.
 * The Python modules in anyref_b200/segment_anything/ keep the reference class names / signatures / state_dict
 * and call these functions through ctypes (see INTEGRATION.md for the binding).
 *
 * Conventions
 *   - all pointers are DEVICE pointers unless a parameter name ends in `_host`;
 *   - `stream` is a cudaStream_t passed as void*; work is enqueued, nothing is allocated, nothing synchronises;
 *   - return value 0 = ok, non-zero = error; the text is returned by sam_last_error() (thread-local);
 *   - `fmt`: 0 = fp16, 1 = bf16 (tensor-core operand formats), 2 = fp32.
 */
#ifndef ANYREF_SAM_H_
#define ANYREF_SAM_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif


/* Shape of the ViT image encoder (arguments of _build_sam, build_sam.py:56-88). */
typedef struct SamEncoderShape {
  int embed_dim;          /* 1280 for ViT-H */
  int depth;              /* 32 */
  int heads;              /* 16 (head_dim must be 80) */
  int mlp_dim;            /* embed_dim * mlp_ratio = 5120 */
  int img;                /* 1024 */
  int patch;              /* 16 */
  int window;             /* 14 */
  int out_chans;          /* 256 */
  int fmt;                /* tensor-core operand format: 0 fp16, 1 bf16 */
  unsigned long long global_mask; /* bit i set <=> block i uses global attention (global_attn_indexes) */
  int tap_block;          /* test hook: copy the fp32 residual stream after this block to tap_out (-1 = off) */
  float* tap_out;         /* device [B*64*64, embed_dim] fp32 or NULL */
  int ln_fold;            /* 1: norm1 / norm2 folded into the neighbouring GEMMs (needs embed_dim % 256 == 0 and the
                             folded weight-blob layout, csrc/encoder.cpp); 0: LayerNorm kernels */
} SamEncoderShape;

/* Shape of the mask decoder (mask_decoder.py:17-73, transformer.py:16-60). */
typedef struct SamDecoderShape {
  int C;                  /* transformer_dim = 256 */
  int heads;              /* 8 */
  int depth;              /* 2 */
  int mlp_dim;            /* 2048 */
  int num_mask_tokens;    /* num_multimask_outputs + 1 = 4 */
  int iou_hidden;         /* 256 */
  int grid;               /* image embedding size = 64 */
} SamDecoderShape;

/* Text of the last error raised on the calling thread ("" if none). */
const char* sam_last_error(void);
/* Library ABI version (bumped on incompatible change). */
int sam_abi_version(void);

/*
 * C[M,N] = epilogue(A[M,K] . W[N,K]^T)  -- tcgen05/TMEM GEMM, TMA-fed.
 * Replaces F.linear / 1x1 conv library GEMMs: image_encoder.py:238 (qkv), :258 (proj), common.py:26 (lin1, lin2),
 * image_encoder.py:93 (neck 1x1), :418 (patch-embed as a patch GEMM), :100 (neck 3x3 after im2col).
 *   A, W       : operand format `fmt` (0 fp16 / 1 bf16), row-major with leading dimensions lda / ldw (elements)
 *   out        : out_fmt 0/1/2, leading dimension ldo
 *   bias       : fp32 [N] or NULL
 *   act        : 0 none, 1 exact-erf GELU (common.py:18), 2 ReLU (text_hidden_fcs, model/anyref.py:118-123)
 *   res        : fp32 residual [res_mod, ldr] or NULL; out[row] += res[row % res_mod]; may alias out (in place)
 */
int sam_gemm(const void* A, int lda, const void* W, int ldw, int M, int N, int K, int fmt, void* out, int ldo,
             int out_fmt, const float* bias, int act, const float* res, int ldr, int res_mod, void* stream);

/*
 * Test-only: one 128 x N x K tcgen05 tile with thread-written shared-memory operands in a chosen canonical layout
 * (pins the UMMA descriptor conventions the attention kernels depend on).  A [128,K]; B [N,K] (b_mode 0..2) or
 * [K,N] (b_mode 3..6); D fp32 [128,N].  lbo/sbo < 0 selects the mode's default.
 */
int sam_umma_probe(const void* A, const void* B, float* D, int N, int K, int fmt, int a_mode, int b_mode, int a_lbo,
                   int a_sbo, int b_lbo, int b_sbo, void* stream);

/*
 * Row LayerNorm: out[row] = (v - mean(v)) / sqrt(var(v) + eps) * gamma + beta with v = x[row] (+ res[row]).
 * Replaces nn.LayerNorm(eps=1e-6) of Block.norm1 / norm2 (image_encoder.py:179, :191; eps from build_sam.py:73),
 * LayerNorm2d on channels-last rows (common.py:38-43) and the decoder's post-residual norms (transformer.py:157-181,
 * eps 1e-5).  x, res fp32 [M, ld*]; out fmt 0/1/2.  normalize == 0 -> plain cast of x (+res) to out_fmt.
 */
int sam_layernorm(const float* x, int ldx, const float* res, int ldr, const float* gamma, const float* beta, float eps,
                  void* out, int ldo, int out_fmt, int M, int C, int normalize, void* stream);

/*
 * LayerNorm folded into the GEMMs on either side of it (Block.forward, image_encoder.py:177-193: norm1 -> attn.qkv,
 * norm2 -> mlp.lin1, eps 1e-6), so that x is never re-read by a normalisation pass:
 *
 *   sam_gemm_residual_ln  x[M,N] += A[M,K].W[N,K]^T + bias  in place (fp32; image_encoder.py:190, :192), and in the
 *                         same epilogue  xb[M,ldxb] = round(x) in format `fmt`  and  stats[M, N/128] (float2) = the
 *                         (mean, sum of squared deviations from it) of every 128-column slice of the new row.  N % 128 == 0, M % 32 == 0.
 *   sam_cast_stats        the same xb / stats from an existing fp32 x [M, C] (first block).  C % 128 == 0.
 *   sam_gemm_ln           out[M,N] = act( rstd_r * (xb[M,K].Wg[N,K]^T - mean_r * colsum[n]) + bias_fold[n] )  with
 *                         Wg = gamma o W rounded to `fmt`, colsum[n] = sum_k Wg[n,k], bias_fold = beta.W^T + b,
 *                         mean_r / rstd_r from stats[M, parts] over the K-wide row  ==  act(LN(x).W^T + b).
 *                         out_fmt 0/1; act 0 none, 1 exact-erf GELU.
 */
int sam_gemm_residual_ln(const void* A, int lda, const void* W, int ldw, int M, int N, int K, int fmt, float* x, int ldx,
                         const float* bias, void* xb, int ldxb, void* stats, void* stream);
int sam_cast_stats(const float* x, int ldx, void* xb, int ldxb, int fmt, void* stats, int M, int C, void* stream);
int sam_gemm_ln(const void* xb, int lda, const void* Wg, int ldw, int M, int N, int K, int fmt, void* out, int ldo,
                int out_fmt, const float* bias_fold, const float* colsum, const void* stats, int parts, float eps,
                int act, void* stream);

/*
 * PatchEmbed im2col (image_encoder.py:418-426): img [B,3,S,S] (fmt in_fmt) -> out [B*(S/p)^2, 3*p*p] (fmt out_fmt),
 * column order (c, ky, kx) == the flattened Conv2d weight, so patch-embed becomes sam_gemm with W = proj.weight.
 */
int sam_patch_im2col(const void* img, int in_fmt, void* out, int out_fmt, int B, int S, int p, void* stream);

/*
 * 3x3 / pad 1 im2col on channels-last activations for the neck's second conv (image_encoder.py:100-106):
 * in [B,g,g,C] (16-bit) -> out [B*g*g, 9*C], column order (ky, kx, c).
 */
int sam_im2col3x3(const void* in, void* out, int B, int g, int C, void* stream);

/*
 * LayerNorm2d (common.py:31-43, eps 1e-6) on channels-last fp32 rows fused with the NHWC -> NCHW transposition that
 * yields the encoder output [B,C,g,g] (image_encoder.py:107, :124).  x [B*tokens, C] fp32; out fmt 0/1/2.
 */
int sam_ln_nhwc_to_nchw(const float* x, const float* gamma, const float* beta, float eps, void* out, int out_fmt,
                        int B, int tokens_per_img, int C, void* stream);

/*
 * Windowed (14x14) encoder attention with window partition / un-partition and the decomposed rel-pos bias fused in.
 * Replaces image_encoder.py:235-257 (Attention.forward without qkv/proj), :263-318 (window_partition/unpartition)
 * and :354-392 (add_decomposed_rel_pos) for the 28 windowed blocks.
 *   qkv     [B*64*64, 3E] fmt 0/1, token-major, NOT partitioned, columns ordered (q|k|v, head, d)
 *   bias_op [3E]          qkv bias in operand format -- q/k/v of the zero-padded window tokens (image_encoder.py:281)
 *   rel_tab [64, 80]      operand format: rows 0..26 rel_pos_h, rows 32..58 rel_pos_w, rest zero
 *   out     [B*64*64, E]  operand format, heads merged (the A operand of the proj GEMM)
 * head_dim must be 80 (ViT-H), grid 64x64.
 */
int sam_attn_window(const void* qkv, const void* bias_op, const void* rel_tab, void* out, int B, int E, int heads,
                    int fmt, void* stream);

/*
 * Global (64x64 tokens) encoder attention, flash-style, decomposed rel-pos bias fused into the online softmax.
 * Replaces image_encoder.py:235-257 + :354-392 for blocks 7/15/23/31.
 *   rh_rev / rw_rev [128, 80] operand format: row j = rel_pos_{h,w}[126 - j] for j < 127, row 127 zero.
 */
int sam_attn_global(const void* qkv, const void* rh_rev, const void* rw_rev, void* out, int B, int E, int heads,
                    int fmt, void* stream);

/*
 * ImageEncoderViT.forward (image_encoder.py:110-125): images [B,3,1024,1024] (in_fmt 0/1/2) -> out [B,256,64,64]
 * (out_fmt 0/1/2).  One call enqueues the whole encoder (patch-embed GEMM, 32 x {LN, QKV GEMM, fused attention,
 * proj GEMM + residual, LN, MLP GEMMs with GELU / residual epilogues}, neck).
 *   w16 / w32 : weight blobs in the layout documented in csrc/encoder.cpp (operand-format matrices / fp32 vectors);
 *               sizes from sam_encoder_w16_elems / sam_encoder_w32_elems; packed by segment_anything/_pack.py
 *   workspace : device scratch of at least sam_encoder_workspace_bytes(shape, B) bytes, 1024-byte aligned
 */
size_t sam_encoder_w16_elems(const SamEncoderShape* shape);
size_t sam_encoder_w32_elems(const SamEncoderShape* shape);
size_t sam_encoder_workspace_bytes(const SamEncoderShape* shape, int B);
int sam_encoder_forward(const SamEncoderShape* shape, const void* w16, const float* w32, const void* images,
                        int in_fmt, int B, void* out, int out_fmt, void* workspace, size_t workspace_bytes,
                        void* stream);

/*
 * MaskDecoder.forward / predict_masks (mask_decoder.py:75-179) incl. TwoWayTransformer (transformer.py:62-106) for n
 * prompts in one call.  Prompt p uses image embedding img_index[p] (img_index == NULL: n_images must be 1 and all
 * prompts use it -- the reference's per-image call, model/anyref.py:807).
 *   weights          fp32 blob, state_dict order of mask_decoder.* with the two ConvTranspose2d weights rearranged
 *                    (see csrc/decoder.cu carve_weights); size from sam_decoder_weight_elems
 *   derived          device buffer of sam_decoder_derived_bytes() bytes (256-byte aligned) filled by
 *                    sam_decoder_prepare whenever `weights` or the dense positional encoding change: split-bf16 and
 *                    transposed weight copies and the positional halves pe.W^T + b of the image-side projections
 *                    ((keys + pe).W^T = keys.W^T + pe.W^T; prompt_encoder.get_dense_pe() is a constant of the model)
 *   image_pe         [1, C, g, g] pe_fmt, PromptEncoder.get_dense_pe()  (argument of sam_decoder_prepare)
 *   image_embeddings [n_images, C, g, g] emb_fmt
 *   sparse           [n, k, C] sparse_fmt (the [SEG] embeddings; tokens = [iou, mask x4, sparse...]); up to 5 + k = 16
 *                    tokens run on the fused token kernels, more on a plain (slower) fp32 composition of the same call
 *   dense_vec        [C] (no_mask_embed broadcast, prompt_encoder.py:181-184) or NULL;
 *   dense_full       [n, C, g, g] or NULL (mask prompts); both in dense_fmt
 *   masks            [n, num_mask_tokens, 4g, 4g] out_fmt -- ALL mask tokens (caller slices [0:1] or [1:], :106-111)
 *   iou              [n, num_mask_tokens] out_fmt
 *   workspace        256-byte aligned scratch of sam_decoder_workspace_bytes(shape, n_images, n, k) bytes
 */
size_t sam_decoder_weight_elems(const SamDecoderShape* shape);
size_t sam_decoder_workspace_bytes(const SamDecoderShape* shape, int n_images, int n, int k);
size_t sam_decoder_derived_bytes(const SamDecoderShape* shape);
int sam_decoder_prepare(const SamDecoderShape* shape, const float* weights, const void* image_pe, int pe_fmt,
                        void* derived, void* stream);
int sam_decoder_forward(const SamDecoderShape* shape, const float* weights, const void* derived,
                        const void* image_embeddings, int emb_fmt, int n_images, const int* img_index,
                        const void* sparse, int sparse_fmt, int n, int k, const void* dense_vec, const void* dense_full,
                        int dense_fmt, void* masks, void* iou, int out_fmt, void* workspace, size_t workspace_bytes,
                        void* stream);

/*
 * Training path of the mask decoder.  AnyRef fine-tunes mask_decoder.* with the encoders frozen
 * (model/anyref.py:108-113); the loss reaches it through mask_decoder -> postprocess_masks (model/anyref.py:406-450).
 * sam_decoder_train_forward is the same function as sam_decoder_forward (all arithmetic fp32; masks / iou fp32 out;
 * sparse fp32; image_pe is passed directly, no `derived` buffer) but keeps every intermediate in `workspace` and
 * returns a tape.  sam_decoder_backward consumes the tape once:
 *   d_masks   [n, num_mask_tokens, 4g, 4g] fp32 or NULL;   d_iou [n, num_mask_tokens] fp32 or NULL
 *   mask_lo, mask_hi   the caller's promise that d_masks is zero outside mask tokens [mask_lo, mask_hi)
 *             (multimask_output=False uses token 0, True tokens 1.., mask_decoder.py:106-111): the hypernetwork MLPs of
 *             the other tokens are skipped; pass 0, num_mask_tokens when unknown
 *   d_weights fp32, layout of `weights`: the gradient is ADDED (zero it for a fresh gradient)
 *   d_sparse  [n, k, C] fp32, overwritten (may be NULL)
 * Image embeddings, dense prompt embeddings and image_pe receive no gradient (frozen in the reference).  The workspace
 * must stay untouched between the two calls; sam_decoder_tape_free releases the host-side tape (also after backward).
 */
size_t sam_decoder_train_workspace_bytes(const SamDecoderShape* shape, int n, int k);
int sam_decoder_train_forward(const SamDecoderShape* shape, const float* weights, const void* image_embeddings,
                              int emb_fmt, int n_images, const int* img_index, const float* sparse, int n, int k,
                              const void* dense_vec, const void* dense_full, int dense_fmt, const void* image_pe,
                              int pe_fmt, float* masks, float* iou, void* workspace, size_t workspace_bytes,
                              void** tape, void* stream);
int sam_decoder_backward(void* tape, const float* d_masks, int mask_lo, int mask_hi, const float* d_iou,
                         float* d_weights, float* d_sparse, void* stream);
void sam_decoder_tape_free(void* tape);

/*
 * fp32 nn.Linear forward / backward for text_hidden_fcs in training (model/anyref.py:116-124, :395-401):
 * Y [M, N] = X [M, K] . W[N, K]^T + b (optionally ReLU);  dX = dY . W and dW = dY^T . X (both overwritten, either may be
 * NULL), db += column sums of dY (zero it first).  With relu_y (the output of a forward that fused the ReLU) dY is first masked IN PLACE with
 * (relu_y > 0).  scratch (both calls): sam_linear_f32_scratch_bytes(M, N, K) bytes.
 */
size_t sam_linear_f32_scratch_bytes(int M, int N, int K);
int sam_linear_f32_forward(const float* X, const float* W, const float* b, float* Y, int M, int N, int K, int relu,
                           void* scratch, size_t scratch_bytes, void* stream);
int sam_linear_f32_backward(float* dY, const float* relu_y, const float* X, const float* W, float* dX, float* dW,
                            float* db, int M, int N, int K, void* scratch, size_t scratch_bytes, void* stream);

/*
 * Adjoint of sam_postprocess_masks (the loss of model/anyref.py:432-450 is taken on the post-processed logits):
 * d_low [num_masks, L, L] fp32 (overwritten) from d_logits [num_masks, H, W] fp32; tmp: num_masks * h_in * w_in floats.
 */
int sam_postprocess_masks_backward(const float* d_logits, int num_masks, int L, int S, int h_in, int w_in, int H, int W,
                                   float* tmp, float* d_low, void* stream);

/*
 * Sam.postprocess_masks (sam.py:159-172) fused: bilinear L x L -> S x S, crop [:h_in, :w_in], bilinear -> H x W, both
 * align_corners=False, no antialias.  low [num_masks, L, L] (low_fmt); logits fp32 [num_masks, H, W] and/or binary
 * uint8 [num_masks, H, W] = logits > threshold (Sam.mask_threshold, sam.py:19).  Either output may be NULL.
 */
int sam_postprocess_masks(const void* low, int low_fmt, int num_masks, int L, int S, int h_in, int w_in, int H, int W,
                          float* logits, unsigned char* binary, float threshold, void* stream);

/*
 * postprocess_masks fused with the evaluation statistics (eval_referseg.py:186-211 over intersectionAndUnionGPU,
 * utils/utils.py:79-91, K = 2, ignore_index = 255): target uint8 [num_masks, H, W] (0 / 1 / 255); counts int32
 * [num_masks, 6] = {inter_0, inter_1, pred_0, pred_1, target_0, target_1} is ADDED to (zero it first).  logits and
 * binary may both be NULL: then nothing of full resolution is written at all.
 * sam_iou_finalize folds counts into stats fp64 [7] = {inter_bg, inter_fg, union_bg, union_fg, acc_iou_bg, acc_iou_fg,
 * count} (+=), the per-rank vector that one ncclAllReduce(SUM) combines.
 */
int sam_postprocess_masks_iou(const void* low, int low_fmt, int num_masks, int L, int S, int h_in, int w_in, int H, int W,
                              float* logits, unsigned char* binary, float threshold, const unsigned char* target,
                              int* counts, void* stream);

/*
 * postprocess_masks + thresholding with the BIT-PACKED mask as the only full-resolution output: packed uint8
 * [num_masks * H * W / 8], bit 7 of byte 0 = pixel 0 of the flattened [num_masks, H, W] array (numpy.packbits order;
 * needs W % 8 == 0).  This is what the optional ncclAllGather of the masks moves (131,072 B per 1024^2 mask
 * instead of 4 MB of logits); HBM bytes per mask = L*L*4 read + H*W/8 written.  target / counts as in
 * sam_postprocess_masks_iou, or both NULL.
 */
int sam_postprocess_masks_packed(const void* low, int low_fmt, int num_masks, int L, int S, int h_in, int w_in, int H, int W,
                                 unsigned char* packed, float threshold, const unsigned char* target, int* counts,
                                 void* stream);
int sam_iou_finalize(const int* counts, int n, double* stats, void* stream);

/*
 * PromptEncoder.get_dense_pe (prompt_encoder.py:67-76; PositionEmbeddingRandom :203-219): gauss fp32 [2, C/2]
 * (positional_encoding_gaussian_matrix) -> out [1, C, g, g] in out_fmt.
 */
int sam_dense_pe(const float* gauss, void* out, int out_fmt, int C, int g, void* stream);

/*
 * PromptEncoder point / box prompts (prompt_encoder.py:78-109; PositionEmbeddingRandom.forward_with_coords :231-238).
 *   mode 0: coords fp32 [n, n_in, 2] (x, y in input-image pixels), labels fp32 [n, n_in] (-1 not-a-point, 0 negative,
 *           1 positive, anything else: bare encoding); pad = 1 appends the (0, 0) / -1 padding point (:86-90).
 *   mode 1: coords fp32 [n, 2, 2] = box corners (x1, y1), (x2, y2); corner j gets point_embeddings[2 + j].
 * gauss fp32 [2, C/2]; table fp32 [5, C] = point_embeddings[0..3].weight, not_a_point_embed.weight.
 * Writes n_in + pad tokens per prompt into out fp32 [n, ld_tokens, C] starting at token tok0 (so points, boxes and
 * text embeddings can be laid side by side as the reference's torch.cat does, :165-177).
 */
int sam_prompt_sparse(const float* coords, const float* labels, const float* gauss, const float* table, float* out, int n,
                      int n_in, int pad, int mode, int C, int img_h, int img_w, int ld_tokens, int tok0, void* stream);

/*
 * PromptEncoder mask prompts: mask_downscaling (prompt_encoder.py:56-64, :111-114) fused per output pixel.
 * blob fp32 in state_dict order: mask_downscaling.0.{weight,bias}, .1.{weight,bias}, .3.{weight,bias}, .4.{weight,bias},
 * .6.{weight,bias} (sam_prompt_mask_blob_elems floats).  masks [n, 1, 4g, 4g] (in_fmt) -> out [n, C, g, g] (out_fmt).
 */
size_t sam_prompt_mask_blob_elems(int mask_in_chans, int C);
int sam_prompt_mask_embed(const void* masks, int in_fmt, const float* blob, int mask_in_chans, void* out, int out_fmt,
                          int n, int g, int C, void* stream);

/*
 * Sam.preprocess (sam.py:174-184; AnyRef's sam_preprocess, utils/refer_seg.py:560-593): (img - mean) / std, zero-pad
 * to S x S, cast.  img [B, 3, h, w]: in_fmt 0 fp16, 1 bf16, 2 fp32, 3 uint8;  out [B, 3, S, S] in out_fmt (0 / 1 / 2).
 * mean / std: HOST pointers to 3 floats (Sam.pixel_mean / pixel_std, sam.py:46-49).
 */
int sam_preprocess(const void* img, int in_fmt, void* out, int out_fmt, int B, int h, int w, int S, const float* mean,
                   const float* std, void* stream);

/*
 * ResizeLongestSide.apply_image (utils/transforms.py:27-34; torchvision resize of a PIL image = PIL BILINEAR with
 * antialiasing) for an HWC uint8 image on the device: separable 22-bit fixed-point resample, horizontal pass first with
 * a uint8 intermediate, bit-exact with Pillow's Resample.c given Pillow's coefficient tables
 *   xbounds [new_w, 2] int32 (first source column, tap count), xcoeff [new_w, xk] int32 (weights * 2^22, rounded);
 *   ybounds / ycoeff likewise for rows.  A table may be NULL when that axis keeps its size.
 *   tmp: [H, new_w, C] uint8 scratch (needed when both axes change); out: [new_h, new_w, C] uint8.
 */
int sam_resize_u8(const unsigned char* in, int H, int W, int C, unsigned char* tmp, unsigned char* out, int new_h,
                  int new_w, const int* xbounds, const int* xcoeff, int xk, const int* ybounds, const int* ycoeff,
                  int yk, void* stream);

/*
 * Scheduling switch of the 2-CTA GEMM behind every sam_gemm* entry point and the encoder (no reference counterpart:
 * the reference's nn.Linear calls, modeling/image_encoder.py:238, :257, modeling/common.py:26, leave tiling to cuBLAS).
 * When the last round of 256 x 256 output tiles would leave at least half of the CTA pairs idle, its tiles are cut into
 * 256 x 128 halves, one per pair.  Results are bit-identical either way except the LayerNorm slice statistics of the
 * residual producers, which combine two 64-column halves (same value to fp32 rounding).
 * mode: -1 = policy (launches of fewer than 8 whole rounds or of at most 16384 rows, i.e. batches of up to 4 images),
 *       0 = never, 1 = whenever possible.
 */
void sam_gemm_set_tile_split(int mode);
/*
 * Test-only, host code (no GPU needed): the work items CTA pair `pair` of `num_pairs` walks for a launch of `num_tiles`
 * output tiles, from the same function the kernel calls.  out[4 i + 0] = tile, [4 i + 1] = 1 for a 256 x 128 half item,
 * [4 i + 2] = its column slice (0 | 1).  Returns the number of items, -1 if more than `cap`.
 */
int sam_gemm_schedule(int num_tiles, int num_pairs, int pair, int split, int* out, int cap);

/*
 * Launch accounting and per-kernel-class timing (used by bench.py for `gpu_launches` and the roofline leg).
 * sam_launch_count: kernels launched by this library since load.  With profiling enabled every launch is bracketed by
 * a CUDA event pair on its stream; sam_profile_collect synchronises those events and adds them to per-class totals.
 * Classes: 0 GEMM, 1 windowed attention, 2 global attention, 3 LayerNorm, 4 layout (im2col / dense PE), 5 mask
 * decoder, 6 postprocess.  flops / bytes are the ALGORITHMIC work stated by each launch site.
 */
long long sam_launch_count(void);
void sam_profile_enable(int on);
void sam_profile_reset(void);
int sam_profile_collect(void);
void sam_profile_get(int cls, double* ms, long long* launches, double* flops, double* bytes);

#ifdef __cplusplus
}
#endif
#endif /* ANYREF_SAM_H_ */
