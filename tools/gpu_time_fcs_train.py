"""text_hidden_fcs (Linear(4096, 4096) - ReLU - Linear(4096, 256), model/anyref.py:116-124) in training: forward + backward
of this path's fp32 linears against stock PyTorch (fp32, TF32 off) for a handful of [SEG] rows."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from anyref_b200.seg_head import build_text_hidden_fcs

torch.backends.cuda.matmul.allow_tf32 = False
H = 4096
M = int(sys.argv[1]) if len(sys.argv) > 1 else 3
fcs = build_text_hidden_fcs(H, 256).cuda()
ref = torch.nn.Sequential(torch.nn.Linear(H, H), torch.nn.ReLU(), torch.nn.Linear(H, 256)).cuda()
ref[0].load_state_dict(fcs[0][0].state_dict())
ref[2].load_state_dict(fcs[0][2].state_dict())
x0 = torch.randn(M, H, device="cuda")
cot = torch.randn(M, 256, device="cuda")


def run(mod):
    x = x0.clone().requires_grad_(True)
    (mod(x) * cot).sum().backward()
    return x.grad


def timed(fn, iters=30):
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


ga, gb = run(fcs[0]), run(ref)
print(f"M={M}: dX rel diff {float((ga - gb).norm() / gb.norm()):.2e}; this path {timed(lambda: run(fcs[0])):.3f} ms, "
      f"stock PyTorch {timed(lambda: run(ref)):.3f} ms per forward + backward")
