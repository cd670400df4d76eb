"""Thin Python wrappers over the C ABI, one per kernel entry point.

These take torch CUDA tensors (PyTorch owns all memory), pass raw device pointers + the current stream to
libanyref_sam.so, and raise RuntimeError on any non-zero status.  No torch math happens here.
"""
from __future__ import annotations

import torch

from . import _lib
from ._lib import check, fmt_of, ptr, stream_ptr


def _req_cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise RuntimeError("anyref_b200 kernels need CUDA tensors (no CPU fallback)")


@_lib.device_scoped
def gemm(a: torch.Tensor, w: torch.Tensor, *, bias=None, act: str = "none", residual=None, res_mod: int | None = None,
         out: torch.Tensor | None = None, out_dtype=None) -> torch.Tensor:
    """out[M,N] = epilogue(a[M,K] @ w[N,K]^T); a, w fp16/bf16; bias/residual fp32; see sam_gemm."""
    _req_cuda(a, w, bias, residual, out)
    assert a.dim() == 2 and w.dim() == 2 and a.shape[1] == w.shape[1] and a.dtype == w.dtype
    assert a.stride(1) == 1 and w.stride(1) == 1
    M, K = a.shape
    N = w.shape[0]
    if out is None:
        out = torch.empty((M, N), device=a.device, dtype=out_dtype or a.dtype)
    assert out.shape == (M, N) and out.stride(1) == 1
    ldr, rmod = 0, 0
    if residual is not None:
        assert residual.dtype == torch.float32 and residual.stride(1) == 1
        ldr = residual.stride(0)
        rmod = res_mod if res_mod is not None else residual.shape[0]
    if bias is not None:
        assert bias.dtype == torch.float32 and bias.numel() == N and bias.is_contiguous()
    lib = _lib.load()
    rc = lib.sam_gemm(ptr(a), a.stride(0), ptr(w), w.stride(0), M, N, K, fmt_of(a.dtype), ptr(out), out.stride(0),
                      fmt_of(out.dtype), ptr(bias), {"none": 0, "gelu": 1, "relu": 2}[act], ptr(residual), ldr, rmod,
                      stream_ptr(a.device))
    check(rc, "sam_gemm")
    return out


@_lib.device_scoped
def cast_stats(x: torch.Tensor, out_dtype):
    """-> (xb = x.to(out_dtype), stats [M, C/128, 2] fp32 per-slice (mean, sum of squared deviations) of every 128 columns); see sam_cast_stats."""
    _req_cuda(x)
    assert x.dim() == 2 and x.dtype == torch.float32 and x.stride(1) == 1
    M, Cc = x.shape
    xb = torch.empty((M, Cc), device=x.device, dtype=out_dtype)
    stats = torch.empty((M, Cc // 128, 2), device=x.device, dtype=torch.float32)
    rc = _lib.load().sam_cast_stats(ptr(x), x.stride(0), ptr(xb), xb.stride(0), fmt_of(out_dtype), ptr(stats), M, Cc,
                                    stream_ptr(x.device))
    check(rc, "sam_cast_stats")
    return xb, stats


@_lib.device_scoped
def gemm_residual_ln(a: torch.Tensor, w: torch.Tensor, x: torch.Tensor, bias=None, *, xb=None, stats=None):
    """x += a @ w^T + bias in place (fp32); -> (xb = round(x) in a.dtype, stats [M, N/128, 2]); see sam_gemm_residual_ln."""
    _req_cuda(a, w, x, bias)
    M, K = a.shape
    N = w.shape[0]
    assert x.shape == (M, N) and x.dtype == torch.float32 and x.stride(1) == 1 and a.dtype == w.dtype
    if xb is None:
        xb = torch.empty((M, N), device=a.device, dtype=a.dtype)
    if stats is None:
        stats = torch.empty((M, N // 128, 2), device=a.device, dtype=torch.float32)
    rc = _lib.load().sam_gemm_residual_ln(ptr(a), a.stride(0), ptr(w), w.stride(0), M, N, K, fmt_of(a.dtype), ptr(x),
                                          x.stride(0), ptr(bias), ptr(xb), xb.stride(0), ptr(stats), stream_ptr(a.device))
    check(rc, "sam_gemm_residual_ln")
    return xb, stats


@_lib.device_scoped
def gemm_ln(xb: torch.Tensor, wg: torch.Tensor, bias_fold: torch.Tensor, colsum: torch.Tensor, stats: torch.Tensor,
            eps: float, act: str = "none", out: torch.Tensor | None = None) -> torch.Tensor:
    """act(LN(x) @ W^T + b) from the folded operands (see sam_gemm_ln / segment_anything/_pack.py::_fold_layernorm)."""
    _req_cuda(xb, wg, bias_fold, colsum, stats)
    M, K = xb.shape
    N = wg.shape[0]
    assert stats.dtype == torch.float32 and stats.is_contiguous() and stats.shape[0] == M and stats.shape[2] == 2
    if out is None:
        out = torch.empty((M, N), device=xb.device, dtype=xb.dtype)
    rc = _lib.load().sam_gemm_ln(ptr(xb), xb.stride(0), ptr(wg), wg.stride(0), M, N, K, fmt_of(xb.dtype), ptr(out),
                                 out.stride(0), fmt_of(out.dtype), ptr(bias_fold), ptr(colsum), ptr(stats),
                                 stats.shape[1], float(eps), {"none": 0, "gelu": 1}[act], stream_ptr(xb.device))
    check(rc, "sam_gemm_ln")
    return out


@_lib.device_scoped
def umma_probe(a: torch.Tensor, b: torch.Tensor, N: int, K: int, a_mode: int, b_mode: int, a_lbo=-1, a_sbo=-1,
               b_lbo=-1, b_sbo=-1) -> torch.Tensor:
    _req_cuda(a, b)
    d = torch.zeros((128, N), device=a.device, dtype=torch.float32)
    lib = _lib.load()
    rc = lib.sam_umma_probe(ptr(a), ptr(b), ptr(d), N, K, fmt_of(a.dtype), a_mode, b_mode, a_lbo, a_sbo, b_lbo, b_sbo,
                            stream_ptr(a.device))
    check(rc, "sam_umma_probe")
    return d


@_lib.device_scoped
def layernorm(x: torch.Tensor, gamma, beta, eps: float, out_dtype, *, residual=None, normalize: bool = True,
              out: torch.Tensor | None = None) -> torch.Tensor:
    """Row LayerNorm of fp32 [M, C] (optionally of x + residual) -> out_dtype; see sam_layernorm."""
    _req_cuda(x, gamma, beta, residual, out)
    assert x.dim() == 2 and x.dtype == torch.float32 and x.stride(1) == 1
    M, Cc = x.shape
    if out is None:
        out = torch.empty((M, Cc), device=x.device, dtype=out_dtype)
    rc = _lib.load().sam_layernorm(ptr(x), x.stride(0), ptr(residual), residual.stride(0) if residual is not None else 0,
                                   ptr(gamma), ptr(beta), float(eps), ptr(out), out.stride(0), fmt_of(out.dtype), M, Cc,
                                   1 if normalize else 0, stream_ptr(x.device))
    check(rc, "sam_layernorm")
    return out


@_lib.device_scoped
def patch_im2col(img: torch.Tensor, patch: int, out_dtype) -> torch.Tensor:
    _req_cuda(img)
    B, ch, S, S2 = img.shape
    assert ch == 3 and S == S2 and img.is_contiguous()
    g = S // patch
    out = torch.empty((B * g * g, 3 * patch * patch), device=img.device, dtype=out_dtype)
    rc = _lib.load().sam_patch_im2col(ptr(img), fmt_of(img.dtype), ptr(out), fmt_of(out_dtype), B, S, patch,
                                      stream_ptr(img.device))
    check(rc, "sam_patch_im2col")
    return out


@_lib.device_scoped
def im2col3x3(x: torch.Tensor, B: int, g: int) -> torch.Tensor:
    _req_cuda(x)
    Cc = x.shape[-1]
    assert x.is_contiguous() and x.numel() == B * g * g * Cc and x.element_size() == 2
    out = torch.empty((B * g * g, 9 * Cc), device=x.device, dtype=x.dtype)
    rc = _lib.load().sam_im2col3x3(ptr(x), ptr(out), B, g, Cc, stream_ptr(x.device))
    check(rc, "sam_im2col3x3")
    return out


@_lib.device_scoped
def ln_nhwc_to_nchw(x: torch.Tensor, gamma, beta, eps: float, B: int, g: int, out_dtype) -> torch.Tensor:
    _req_cuda(x, gamma, beta)
    Cc = x.shape[-1]
    assert x.dtype == torch.float32 and x.is_contiguous()
    out = torch.empty((B, Cc, g, g), device=x.device, dtype=out_dtype)
    rc = _lib.load().sam_ln_nhwc_to_nchw(ptr(x), ptr(gamma), ptr(beta), float(eps), ptr(out), fmt_of(out_dtype), B, g * g,
                                         Cc, stream_ptr(x.device))
    check(rc, "sam_ln_nhwc_to_nchw")
    return out


def window_rel_table(rel_pos_h: torch.Tensor, rel_pos_w: torch.Tensor, dtype) -> torch.Tensor:
    """[64, hd] operand table of sam_attn_window: rows 0..26 rel_pos_h, rows 32..58 rel_pos_w, rest zero."""
    n, hd = rel_pos_h.shape
    t = torch.zeros((64, hd), device=rel_pos_h.device, dtype=dtype)
    t[:n] = rel_pos_h.to(dtype)
    t[32:32 + n] = rel_pos_w.to(dtype)
    return t


def global_rel_table(rel_pos: torch.Tensor, dtype) -> torch.Tensor:
    """[128, hd] operand table of sam_attn_global: row j = rel_pos[126 - j] (j < 127), row 127 zero."""
    n, hd = rel_pos.shape
    t = torch.zeros((n + 1, hd), device=rel_pos.device, dtype=dtype)
    t[:n] = rel_pos.flip(0).to(dtype)
    return t


@_lib.device_scoped
def attn_window(qkv: torch.Tensor, bias_op: torch.Tensor, rel_tab: torch.Tensor, B: int, heads: int) -> torch.Tensor:
    _req_cuda(qkv, bias_op, rel_tab)
    E = qkv.shape[1] // 3
    assert qkv.is_contiguous() and qkv.shape[0] == B * 4096 and bias_op.dtype == qkv.dtype == rel_tab.dtype
    out = torch.empty((B * 4096, E), device=qkv.device, dtype=qkv.dtype)
    rc = _lib.load().sam_attn_window(ptr(qkv), ptr(bias_op), ptr(rel_tab), ptr(out), B, E, heads, fmt_of(qkv.dtype),
                                     stream_ptr(qkv.device))
    check(rc, "sam_attn_window")
    return out


@_lib.device_scoped
def attn_global(qkv: torch.Tensor, rh_rev: torch.Tensor, rw_rev: torch.Tensor, B: int, heads: int) -> torch.Tensor:
    _req_cuda(qkv, rh_rev, rw_rev)
    E = qkv.shape[1] // 3
    assert qkv.is_contiguous() and qkv.shape[0] == B * 4096
    out = torch.empty((B * 4096, E), device=qkv.device, dtype=qkv.dtype)
    rc = _lib.load().sam_attn_global(ptr(qkv), ptr(rh_rev), ptr(rw_rev), ptr(out), B, E, heads, fmt_of(qkv.dtype),
                                     stream_ptr(qkv.device))
    check(rc, "sam_attn_global")
    return out
