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

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

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
 *   act        : 0 none, 1 exact-erf GELU (common.py:18)
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

#ifdef __cplusplus
}
#endif
#endif /* ANYREF_SAM_H_ */
