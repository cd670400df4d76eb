"""Stand-alone timing of the two attention kernels at the bench shapes (B = 16, ViT-H), CUDA events, inputs > L2.
    [ANYREF_SAM_LIB=anyref_b200/libanyref_sam_<variant>.so] python tools/gpu_time_attn.py [window|global|both] [fp16|bf16]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from anyref_b200 import ops

which = sys.argv[1] if len(sys.argv) > 1 else "both"
dt = torch.float16 if (len(sys.argv) > 2 and sys.argv[2] == "fp16") else torch.bfloat16
dev = "cuda"
B, heads, E = 16, 16, 1280
torch.manual_seed(0)
qkv = torch.randn(B * 4096, 3 * E, device=dev).to(dt)
bias = torch.randn(3 * E, device=dev).to(dt)
tab = ops.window_rel_table(torch.randn(27, 80, device=dev) * 0.1, torch.randn(27, 80, device=dev) * 0.1, dt)
gh = ops.global_rel_table(torch.randn(127, 80, device=dev) * 0.1, dt)
gw = ops.global_rel_table(torch.randn(127, 80, device=dev) * 0.1, dt)


def timeit(fn, iters):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2], ts[0]


lib = os.environ.get("ANYREF_SAM_LIB", "default")
if which in ("window", "both"):
    med, mn = timeit(lambda: ops.attn_window(qkv, bias, tab, B, heads), 30)
    print(f"attn_window {dt} lib={lib}: median {med * 1e3:.1f} us, min {mn * 1e3:.1f} us")
if which in ("global", "both"):
    med, mn = timeit(lambda: ops.attn_global(qkv, gh, gw, B, heads), 10)
    print(f"attn_global {dt} lib={lib}: median {med * 1e3:.1f} us, min {mn * 1e3:.1f} us")
