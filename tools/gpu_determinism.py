"""Run every encoder-side kernel repeatedly on identical inputs and report bitwise run-to-run differences.
usage: python tools/gpu_determinism.py [reps]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from anyref_b200 import ops

dev = "cuda"
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 10
torch.manual_seed(0)


def check(name, fn):
    ref = fn().clone()
    bad, worst = 0, 0.0
    for _ in range(reps):
        out = fn()
        if not torch.equal(out, ref):
            bad += 1
            worst = max(worst, (out.float() - ref.float()).abs().max().item())
    torch.cuda.synchronize()
    print(f"{name:58s} {'DIFFERS in %d/%d runs, max|d|=%.3e' % (bad, reps, worst) if bad else 'deterministic'}", flush=True)


for dt in (torch.float16, torch.bfloat16):
    tag = str(dt)[6:]
    for (B, heads) in ((1, 2), (2, 16)):
        E = heads * 80
        M = B * 4096
        qkv = torch.randn(M, 3 * E, device=dev).to(dt)
        bias = torch.randn(3 * E, device=dev).to(dt)
        tab = ops.window_rel_table(torch.randn(27, 80, device=dev) * 0.1, torch.randn(27, 80, device=dev) * 0.1, dt)
        gh = ops.global_rel_table(torch.randn(127, 80, device=dev) * 0.1, dt)
        gw = ops.global_rel_table(torch.randn(127, 80, device=dev) * 0.1, dt)
        check(f"attn_window {tag} B={B} heads={heads}", lambda: ops.attn_window(qkv, bias, tab, B, heads))
        check(f"attn_global {tag} B={B} heads={heads}", lambda: ops.attn_global(qkv, gh, gw, B, heads))
        x = torch.randn(M, E, device=dev)
        g = torch.randn(E, device=dev)
        b = torch.randn(E, device=dev)
        check(f"layernorm {tag} M={M} C={E}", lambda: ops.layernorm(x, g, b, 1e-6, dt))
        a = (torch.randn(M, E, device=dev) * 0.5).to(dt)
        for (N, K, kind) in ((3 * E, E, "bias"), (4 * E, E, "gelu"), (E, E, "res"), (256, E, "plain")):
            w = (torch.randn(N, K, device=dev) * 0.05).to(dt)
            bb = torch.randn(N, device=dev)
            if kind == "bias":
                check(f"gemm {tag} {M}x{N}x{K} +bias", lambda: ops.gemm(a, w, bias=bb))
            elif kind == "gelu":
                check(f"gemm {tag} {M}x{N}x{K} +bias gelu", lambda: ops.gemm(a, w, bias=bb, act="gelu"))
            elif kind == "plain":
                check(f"gemm {tag} {M}x{N}x{K} fp32 out", lambda: ops.gemm(a, w, out_dtype=torch.float32))
            else:
                x0 = torch.randn(M, N, device=dev)

                def f():
                    xx = x0.clone()
                    ops.gemm(a, w, bias=bb, residual=xx, out=xx)
                    return xx

                check(f"gemm {tag} {M}x{N}x{K} +bias +residual (in place)", f)

# whole tiny encoder
from anyref_b200.segment_anything import build_sam_from_config
from anyref_b200.synthetic import CONFIGS, synthetic_images, synthetic_state_dict

cfg = CONFIGS["vit_tiny80"]
sam = build_sam_from_config(cfg)
sam.load_state_dict(synthetic_state_dict(cfg), strict=True)
sam = sam.cuda()
x = synthetic_images(2, seed=0).cuda()
for dt in (torch.float16, torch.bfloat16):
    sam.image_encoder.set_operand_dtype(dt)
    for nb in (1, 2):
        check(f"encoder tiny80 {str(dt)[6:]} B={nb}", lambda: sam.image_encoder(x[:nb]))
        for blk in (0, 1):
            def f():
                tap = torch.empty(nb * 4096, cfg.embed_dim, device="cuda")
                sam.image_encoder(x[:nb], _tap=(blk, tap))
                return tap
            check(f"  tap after block {blk} B={nb}", f)
print("done")
