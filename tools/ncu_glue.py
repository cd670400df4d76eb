"""Memory-bound kernels of the path at the bench shapes (B = 16), one after the other -- target for ncu captures
(`dram__bytes_*`, duration -> GB/s against the measured copy bandwidth) and for CUDA-event timing.

    python tools/ncu_glue.py [iters]        # prints per-kernel median time, algorithmic bytes and GB/s (events, L2 flushed)
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from anyref_b200 import ops
from anyref_b200.segment_anything import build_sam_from_config
from anyref_b200.synthetic import CONFIGS, synthetic_state_dict

dev = "cuda"
iters = int(sys.argv[1]) if len(sys.argv) > 1 else 5
B, g, E = 16, 64, 1280
M = B * g * g
dt = torch.bfloat16
torch.manual_seed(0)
cfg = CONFIGS["vit_tiny80"]
sam = build_sam_from_config(cfg)
sam.load_state_dict(synthetic_state_dict(cfg))
sam = sam.cuda()

low = torch.randn(B, 1, 256, 256, device=dev)
gt = (torch.rand(B, 1, 1024, 1024, device=dev) > 0.5).to(torch.uint8)
x32 = torch.randn(M, E, device=dev)
n16 = torch.randn(M, 256, device=dev).to(dt)
n32 = torch.randn(M, 256, device=dev)
g256, b256 = torch.ones(256, device=dev), torch.zeros(256, device=dev)
img_u8 = torch.randint(0, 256, (B, 3, 1024, 683), device=dev, dtype=torch.uint8)
img = torch.randn(B, 3, 1024, 1024, device=dev).to(dt)
flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)

cases = {
    # name: (fn, algorithmic bytes)
    "postprocess_logits": (lambda: sam.postprocess_masks(low, (1024, 1024), (1024, 1024)),
                           B * (256 * 256 * 4 + 1024 * 1024 * 4)),
    "postprocess_c3": (lambda: sam.postprocess_masks(low, (1024, 683), (640, 427)), B * (256 * 256 * 4 + 640 * 427 * 4)),
    "postprocess_packed_iou": (lambda: sam.postprocess_and_score(low, (1024, 1024), (1024, 1024), gt, return_packed=True),
                               B * (256 * 256 * 4 + 1024 * 1024 * (1 + 0.125))),
    "cast_stats": (lambda: ops.cast_stats(x32, dt), M * E * (4 + 2) + M * 10 * 8),
    "im2col3x3": (lambda: ops.im2col3x3(n16, B, g), M * 256 * 2 * (1 + 9)),
    "ln_nhwc_to_nchw": (lambda: ops.ln_nhwc_to_nchw(n32, g256, b256, 1e-6, B, g, dt), M * 256 * (4 + 2)),
    "patch_im2col": (lambda: ops.patch_im2col(img, 16, dt), B * 3 * 1024 * 1024 * 2 * 2),
    "preprocess_u8": (lambda: sam.preprocess(img_u8, out_dtype=dt), B * 3 * (1024 * 683 + 1024 * 1024 * 2)),
}
for name, (fn, nbytes) in cases.items():
    for _ in range(2):
        fn()
    ts = []
    for _ in range(iters):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    med = ts[len(ts) // 2]
    print(f"{name:24s} {med * 1e3:9.1f} us  {nbytes / 1e6:9.1f} MB  {nbytes / (med * 1e-3) / 1e9:8.1f} GB/s")
print("ok")
