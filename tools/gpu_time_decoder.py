"""Decoder + postprocess timing without the per-launch profiling events (CUDA events, L2 flushed before every call).
    python tools/gpu_time_decoder.py [n_prompts=16] [n_images=n_prompts] [c3]
`c3`: BASELINE configs[2] -- multimask_output, content 1024x683 -> 640x427 masks."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from anyref_b200.segment_anything import build_sam_from_config
from anyref_b200.synthetic import CONFIGS, synthetic_state_dict

cfg = CONFIGS["vit_tiny80"]
sam = build_sam_from_config(cfg)
sam.load_state_dict(synthetic_state_dict(cfg))
sam = sam.cuda()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 16
n_img = int(sys.argv[2]) if len(sys.argv) > 2 else n
c3 = len(sys.argv) > 3 and sys.argv[3] == "c3"
emb = torch.randn(n_img, 256, 64, 64, device="cuda")
text = torch.randn(n, 1, 256, device="cuda")
idx = (torch.arange(n, device="cuda") * n_img // n).to(torch.int32)
pe = sam.prompt_encoder.get_dense_pe()
big = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")


def step():
    sparse, dense = sam.prompt_encoder(points=None, boxes=None, masks=None, text_embeds=text)
    low, iou = sam.mask_decoder.forward_batched(emb, pe, sparse, dense, idx, c3)
    if c3:
        return sam.postprocess_masks(low, (1024, 683), (640, 427))
    return sam.postprocess_masks(low, (1024, 1024), (1024, 1024))


for it in range(5):
    step()
torch.cuda.synchronize()
ts = []
for it in range(20):
    big.zero_()      # flush L2 and keep the queue busy so launch latency is hidden as in a full step
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    step()
    e1.record()
    torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1))
ts.sort()
print(f"decoder + postprocess, n={n} prompts on {n_img} images{' (C3: multimask, 1024x683 -> 640x427)' if c3 else ''}: median {ts[len(ts) // 2]:.3f} ms, min {ts[0]:.3f} ms")
