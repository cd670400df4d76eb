// extern "C" boundary of libanyref_sam.so (declared in include/anyref_sam.h).
#include "../../include/anyref_sam.h"

#include "host_common.h"
#include "kernels.h"

static inline cudaStream_t S(void* s) { return static_cast<cudaStream_t>(s); }

extern "C" {

const char* sam_last_error(void) { return samhost::last_error(); }
int sam_abi_version(void) { return 1; }

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

}  // extern "C"
