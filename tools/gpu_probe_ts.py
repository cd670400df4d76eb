"""UMMA probe of the tensor-memory A operand (tcgen05.mma 'ts' form): python tools/gpu_probe_ts.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from anyref_b200 import ops

torch.manual_seed(0)
for dt in (torch.bfloat16, torch.float16):
    for (b_mode, N, K) in [(0, 64, 64), (0, 128, 128), (3, 64, 64), (3, 64, 208), (4, 16, 208), (4, 80, 64), (4, 80, 208), (3, 64, 128)]:
        a = torch.randn(128, K, device="cuda").to(dt)
        if b_mode <= 2:
            b = torch.randn(N, K, device="cuda").to(dt)
            ref = a.float() @ b.float().t()
        else:
            b = torch.randn(K, N, device="cuda").to(dt)
            ref = a.float() @ b.float()
        d = ops.umma_probe(a, b, N, K, 7, b_mode)
        torch.cuda.synchronize()
        err = (d - ref).abs().max().item()
        print(f"ts-probe dt={str(dt)[6:]:9s} b_mode={b_mode} N={N:3d} K={K:3d} max_err={err:.4e} {'OK' if err < 0.05 else 'MISMATCH'}")
