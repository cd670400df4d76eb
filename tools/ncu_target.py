"""Tiny single-kernel drivers for `ncu --set full` captures:
    [NCU_M=rows] [NCU_SPLIT=-1|0|1] python tools/ncu_target.py <lin1|proj|qkv|lin2|proj_ln|lin2_ln|qkv_fold|lin1_fold|attn_window|attn_global|layernorm> [iters] [fp16|bf16]
NCU_M: token rows of the GEMM targets (default 65536 = 16 images; 4096 = one image); NCU_SPLIT: sam_gemm_set_tile_split mode."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from anyref_b200 import ops

dev = "cuda"
M = int(os.environ.get("NCU_M", "65536"))
if "NCU_SPLIT" in os.environ:
    from anyref_b200 import _lib
    _lib.gemm_set_tile_split(int(os.environ["NCU_SPLIT"]))
dt = torch.float16 if (len(sys.argv) > 3 and sys.argv[3] == "fp16") else torch.bfloat16
which = sys.argv[1]
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 3
torch.manual_seed(0)
if which in ("lin1", "proj", "qkv", "lin2", "proj_ln", "lin2_ln", "qkv_fold", "lin1_fold"):
    N, K = {"lin1": (5120, 1280), "proj": (1280, 1280), "qkv": (3840, 1280), "lin2": (1280, 5120),
            "proj_ln": (1280, 1280), "lin2_ln": (1280, 5120), "qkv_fold": (3840, 1280), "lin1_fold": (5120, 1280)}[which]
    a = (torch.randn(M, K, device=dev) * 0.5).to(dt)
    w = (torch.randn(N, K, device=dev) * 0.05).to(dt)
    bias = torch.randn(N, device=dev)
    if which == "lin1":
        out = torch.empty(M, N, device=dev, dtype=dt)
        fn = lambda: ops.gemm(a, w, bias=bias, act="gelu", out=out)
    elif which == "qkv":
        out = torch.empty(M, N, device=dev, dtype=dt)
        fn = lambda: ops.gemm(a, w, bias=bias, out=out)
    elif which.endswith("_fold"):
        out = torch.empty(M, N, device=dev, dtype=dt)
        stats = torch.zeros(M, K // 128, 2, device=dev)
        stats[..., 1] = 128.0
        colsum = torch.randn(N, device=dev)
        fn = lambda: ops.gemm_ln(a, w, bias, colsum, stats, 1e-6, act="gelu" if which == "lin1_fold" else "none", out=out)
    elif which.endswith("_ln"):
        x = torch.randn(M, N, device=dev)
        xb = torch.empty(M, N, device=dev, dtype=dt)
        stats = torch.empty(M, N // 128, 2, device=dev)
        fn = lambda: ops.gemm_residual_ln(a, w, x, bias, xb=xb, stats=stats)
    else:
        x = torch.randn(M, N, device=dev)
        fn = lambda: ops.gemm(a, w, bias=bias, residual=x, out=x)
elif which in ("attn_window", "attn_global"):
    B, heads, E = 16, 16, 1280
    qkv = torch.randn(B * 4096, 3 * E, device=dev).to(dt)
    bias = torch.randn(3 * E, device=dev).to(dt)
    tab = ops.window_rel_table(torch.randn(27, 80, device=dev) * 0.1, torch.randn(27, 80, device=dev) * 0.1, dt)
    gh = ops.global_rel_table(torch.randn(127, 80, device=dev) * 0.1, dt)
    gw = ops.global_rel_table(torch.randn(127, 80, device=dev) * 0.1, dt)
    fn = (lambda: ops.attn_window(qkv, bias, tab, B, heads)) if which == "attn_window" else (lambda: ops.attn_global(qkv, gh, gw, B, heads))
elif which == "layernorm":
    x = torch.randn(M, 1280, device=dev)
    g = torch.ones(1280, device=dev)
    b = torch.zeros(1280, device=dev)
    out = torch.empty(M, 1280, device=dev, dtype=dt)
    fn = lambda: ops.layernorm(x, g, b, 1e-6, dt, out=out)
else:
    raise SystemExit(f"unknown target {which}")
for _ in range(iters):
    fn()
torch.cuda.synchronize()
print("ok", which)
