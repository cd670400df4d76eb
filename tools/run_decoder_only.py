"""Decoder + postprocess only (16 prompts on 16 images) -- target for ncu launch lists."""
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
emb = torch.randn(n, 256, 64, 64, device="cuda")
text = torch.randn(n, 1, 256, device="cuda")
idx = torch.arange(n, dtype=torch.int32, device="cuda")
pe = sam.prompt_encoder.get_dense_pe()
for it in range(3):
    sparse, dense = sam.prompt_encoder(points=None, boxes=None, masks=None, text_embeds=text)
    low, iou = sam.mask_decoder.forward_batched(emb, pe, sparse, dense, idx, False)
    out = sam.postprocess_masks(low, (1024, 1024), (1024, 1024))
torch.cuda.synchronize()
print("ok", out.shape)
