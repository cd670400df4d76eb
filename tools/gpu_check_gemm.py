"""Diagnostic run on a B200: UMMA layout probes + GEMM correctness/timing.  Prints, never asserts."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from anyref_b200 import ops


def probe_suite():
    torch.manual_seed(0)
    dev = "cuda"
    print("== UMMA probes (max abs err vs fp32 reference; ~1e-2 or less = layout correct)")
    for dt in (torch.bfloat16, torch.float16):
        for (a_mode, b_mode, N, K) in [
            (0, 0, 64, 64), (0, 0, 256, 128), (0, 0, 208, 128), (0, 0, 80, 64),
            (1, 1, 64, 32), (1, 1, 208, 16), (2, 2, 64, 32), (0, 1, 64, 64), (0, 2, 128, 64),
            (0, 3, 64, 32), (0, 3, 64, 128), (0, 3, 128, 64),
            (0, 4, 16, 32), (0, 4, 80, 128), (0, 4, 64, 64),
            (0, 5, 16, 32), (0, 5, 80, 64),
            (0, 6, 32, 32), (0, 6, 64, 64), (0, 6, 96, 128),
        ]:
            a = torch.randn(128, K, device=dev).to(dt)
            if b_mode <= 2:
                b = torch.randn(N, K, device=dev).to(dt)
                ref = a.float() @ b.float().t()
            else:
                b = torch.randn(K, N, device=dev).to(dt)
                ref = a.float() @ b.float()
            try:
                d = ops.umma_probe(a, b, N, K, a_mode, b_mode)
                torch.cuda.synchronize()
                err = (d - ref).abs().max().item()
                print(f"probe dt={str(dt)[6:]:9s} a_mode={a_mode} b_mode={b_mode} N={N:3d} K={K:3d} max_err={err:.4e} "
                      f"ref_max={ref.abs().max().item():.2f} {'OK' if err < 0.05 else 'MISMATCH'}")
            except Exception as e:  # noqa
                print(f"probe a_mode={a_mode} b_mode={b_mode} N={N} K={K} EXC {e}")
                return
    # alternative LBO/SBO hypotheses for modes that mismatched can be added here


def gemm_suite():
    dev = "cuda"
    torch.manual_seed(1)
    print("== GEMM correctness")
    cases = [
        (128, 256, 64), (256, 512, 128), (300, 256, 192), (4096, 1280, 1280), (4900, 3840, 1280), (1000, 128, 320),
        (4096, 256, 2304), (513, 5120, 1280), (4096, 1280, 5120), (4096, 1280, 768),
    ]
    for dt in (torch.bfloat16, torch.float16):
        for (M, N, K) in cases:
            a = (torch.randn(M, K, device=dev) * 0.5).to(dt)
            w = (torch.randn(N, K, device=dev) * 0.05).to(dt)
            bias = torch.randn(N, device=dev)
            ref = a.float() @ w.float().t()
            out = ops.gemm(a, w, out_dtype=torch.float32)
            torch.cuda.synchronize()
            e0 = (out - ref).abs().max().item()
            out1 = ops.gemm(a, w, bias=bias, act="gelu", out_dtype=dt)
            ref1 = torch.nn.functional.gelu(ref + bias)
            e1 = (out1.float() - ref1).abs().max().item()
            res = torch.randn(M, N, device=dev)
            x = res.clone()
            ops.gemm(a, w, bias=bias, residual=x, out=x)
            e2 = (x - (ref + bias + res)).abs().max().item()
            pos = torch.randn(128, N, device=dev)
            out3 = ops.gemm(a, w, bias=bias, residual=pos, res_mod=128, out_dtype=torch.float32)
            idx = torch.arange(M, device=dev) % 128
            e3 = (out3 - (ref + bias + pos[idx])).abs().max().item()
            torch.cuda.synchronize()
            print(f"gemm dt={str(dt)[6:]:9s} M={M:5d} N={N:5d} K={K:5d} plain={e0:.3e} bias+gelu={e1:.3e} "
                  f"inplace_res={e2:.3e} modres={e3:.3e} ref_max={ref.abs().max().item():.2f}")


def gemm_bench():
    dev = "cuda"
    print("== GEMM timing (CUDA events, 20 iters after 5 warm-up; torch.matmul beside it)")
    for (M, N, K, name) in [(65536, 3840, 1280, "qkv"), (65536, 1280, 1280, "proj"), (65536, 5120, 1280, "lin1"),
                            (65536, 1280, 5120, "lin2"), (65536, 256, 1280, "neck1"), (8192, 8192, 8192, "sq8k")]:
        a = (torch.randn(M, K, device=dev) * 0.5).to(torch.bfloat16)
        w = (torch.randn(N, K, device=dev) * 0.05).to(torch.bfloat16)
        out = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
        for fn, label in ((lambda: ops.gemm(a, w, out=out), "ours"), (lambda: torch.matmul(a, w.t(), out=out), "torch")):
            for _ in range(5):
                fn()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(20):
                fn()
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 20
            print(f"{name:6s} {label:5s} M={M} N={N} K={K}: {ms:.3f} ms  {2.0 * M * N * K / ms / 1e9:.1f} TFLOP/s")


if __name__ == "__main__":
    print(torch.cuda.get_device_name(0))
    t = time.time()
    which = sys.argv[1:] or ["probe", "gemm", "bench"]
    if "probe" in which:
        probe_suite()
    if "gemm" in which:
        gemm_suite()
    if "bench" in which:
        gemm_bench()
    print(f"done in {time.time() - t:.1f}s")
