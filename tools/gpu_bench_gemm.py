"""Sustained timing of the encoder's GEMM shapes WITH their real epilogues (B=16), beside torch.matmul (cuBLAS)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from anyref_b200 import ops

dev = "cuda"
M = 65536
dt = torch.bfloat16


def timeit(fn, iters):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def main():
    iters = int(sys.argv[1]) if len(sys.argv) > 1 else 200
    x32 = torch.randn(M, 1280, device=dev)
    only = sys.argv[2].split(",") if len(sys.argv) > 2 else None
    stats = torch.zeros(M, 10, 2, device=dev)
    stats[..., 1] = 128.0
    colsum = torch.randn(5120, device=dev)
    xb = torch.empty(M, 1280, device=dev, dtype=dt)
    for name, N, K, kind in (("qkv", 3840, 1280, "bias16"), ("qkv_fold", 3840, 1280, "fold16"), ("proj", 1280, 1280, "res32"),
                             ("proj_ln", 1280, 1280, "resln"), ("lin1", 5120, 1280, "gelu16"), ("lin1_fold", 5120, 1280, "foldgelu"),
                             ("lin2", 1280, 5120, "res32"), ("lin2_ln", 1280, 5120, "resln"), ("lin1_noact", 5120, 1280, "bias16"), ("lin1", 5120, 1280, "gelu16"),
                             ("plain_qkv", 3840, 1280, "plain16"),
                             ("plain_proj", 1280, 1280, "plain16"), ("proj_f32out", 1280, 1280, "f32")):
        if only and name not in only:
            continue
        a = (torch.randn(M, K, device=dev) * 0.5).to(dt)
        w = (torch.randn(N, K, device=dev) * 0.05).to(dt)
        bias = torch.randn(N, device=dev)
        out16 = torch.empty(M, N, device=dev, dtype=dt)
        if kind == "bias16":
            fn = lambda: ops.gemm(a, w, bias=bias, out=out16)
        elif kind == "gelu16":
            fn = lambda: ops.gemm(a, w, bias=bias, act="gelu", out=out16)
        elif kind == "fold16":
            fn = lambda: ops.gemm_ln(a, w, bias, colsum[:N], stats, 1e-6, out=out16)
        elif kind == "foldgelu":
            fn = lambda: ops.gemm_ln(a, w, bias, colsum[:N], stats, 1e-6, act="gelu", out=out16)
        elif kind == "resln":
            fn = lambda: ops.gemm_residual_ln(a, w, x32, bias, xb=xb, stats=stats)
        elif kind == "res32":
            fn = lambda: ops.gemm(a, w, bias=bias, residual=x32, out=x32)
        elif kind == "f32":
            o32 = torch.empty(M, N, device=dev)
            fn = lambda: ops.gemm(a, w, bias=bias, out=o32)
        else:
            fn = lambda: ops.gemm(a, w, out=out16)
        ms = timeit(fn, iters)
        ms_t = timeit(lambda: torch.matmul(a, w.t(), out=out16), iters)
        fl = 2.0 * M * N * K
        print(f"{name:12s} {kind:8s} N={N} K={K}: ours {ms:.3f} ms {fl / ms / 1e9:7.1f} TF/s | cuBLAS plain {ms_t:.3f} ms "
              f"{fl / ms_t / 1e9:7.1f} TF/s", flush=True)


if __name__ == "__main__":
    main()
