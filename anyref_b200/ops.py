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
                      fmt_of(out.dtype), ptr(bias), {"none": 0, "gelu": 1}[act], ptr(residual), ldr, rmod,
                      stream_ptr(a.device))
    check(rc, "sam_gemm")
    return out


def umma_probe(a: torch.Tensor, b: torch.Tensor, N: int, K: int, a_mode: int, b_mode: int, a_lbo=-1, a_sbo=-1,
               b_lbo=-1, b_sbo=-1) -> torch.Tensor:
    _req_cuda(a, b)
    d = torch.zeros((128, N), device=a.device, dtype=torch.float32)
    lib = _lib.load()
    rc = lib.sam_umma_probe(ptr(a), ptr(b), ptr(d), N, K, fmt_of(a.dtype), a_mode, b_mode, a_lbo, a_sbo, b_lbo, b_sbo,
                            stream_ptr(a.device))
    check(rc, "sam_umma_probe")
    return d
