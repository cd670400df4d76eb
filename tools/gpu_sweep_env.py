"""A/B timing of one attention kernel under different values of an experiment environment switch (read per launch):
   python tools/gpu_sweep_env.py attn_global SAM_GLOB_STAGGER 0 400 800 1200
Prints the median CUDA-event time of 20 launches per value (B = 16, ViT-H shapes, bf16) and checks that every value
produces bit-identical output."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from anyref_b200 import ops

which, var, values = sys.argv[1], sys.argv[2], sys.argv[3:]
dev, dt = "cuda", torch.bfloat16
B, heads, E = 16, 16, 1280
torch.manual_seed(0)
qkv = torch.randn(B * 4096, 3 * E, device=dev).to(dt)
bias = torch.randn(3 * E, device=dev).to(dt)
tab = ops.window_rel_table(torch.randn(27, 80, device=dev) * 0.1, torch.randn(27, 80, device=dev) * 0.1, dt)
gh = ops.global_rel_table(torch.randn(127, 80, device=dev) * 0.1, dt)
gw = ops.global_rel_table(torch.randn(127, 80, device=dev) * 0.1, dt)
fn = (lambda: ops.attn_window(qkv, bias, tab, B, heads)) if which == "attn_window" else (lambda: ops.attn_global(qkv, gh, gw, B, heads))
ref = None
for rep in range(2):
    for v in values:
        os.environ[var] = v
        for _ in range(3):
            out = fn()
        ts = []
        for _ in range(20):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            out = fn()
            b.record()
            torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        ts.sort()
        if ref is None:
            ref = out.clone()
        same = torch.equal(out, ref)
        print(f"{which} {var}={v}: median {ts[len(ts) // 2] * 1000:.1f} us  min {ts[0] * 1000:.1f} us  identical={same}", flush=True)
