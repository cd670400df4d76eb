"""Parity of the whole path (ViT-H, fp16 / bf16 operands) against the fp32 CPU oracle on several input seeds:
relative Frobenius error of embeddings / low-res logits and the mask IoU per prompt (north_star bar: IoU >= 0.999).
    python tools/gpu_parity_seeds.py [n_seeds]"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from anyref_b200.grounding import GroundingPath
from anyref_b200.segment_anything import build_sam_vit_h
from anyref_b200.synthetic import CONFIGS, synthetic_images, synthetic_seg_embeddings, synthetic_state_dict
from oracle import sam_oracle as O

n_seeds = int(sys.argv[1]) if len(sys.argv) > 1 else 3
cfg = CONFIGS["vit_h"]
sd = synthetic_state_dict(cfg, seed=1234)
sam = build_sam_vit_h(None)
sam.load_state_dict(sd, strict=True)
sam = sam.cuda()
path = GroundingPath(sam)
torch.set_num_threads(os.cpu_count() or 1)
rows = []
for seed in range(100, 100 + n_seeds):
    x = synthetic_images(1, seed=seed)
    seg = synthetic_seg_embeddings(1, 3, seed=seed)
    sizes_in, sizes_out = [(1024, 683)], [(640, 427)]
    want, emb_ref, low_ref, _ = O.grounding_path(sd, cfg, x, [seg[0]], sizes_in, sizes_out, multimask_output=False,
                                                 return_intermediates=True)
    for dt in (torch.float16, torch.bfloat16):
        sam.image_encoder.set_operand_dtype(dt)
        emb = sam.image_encoder(x.cuda()).float().cpu()
        got = path(x.cuda(), [seg[0].cuda()], sizes_in, sizes_out)[0].cpu()
        w = want[0]
        ious = [(((got[i] > 0) & (w[i] > 0)).sum().item() / max(((got[i] > 0) | (w[i] > 0)).sum().item(), 1)) for i in range(w.shape[0])]
        rows.append({"seed": seed, "dtype": str(dt).split(".")[-1], "emb_rel_fro": ((emb - emb_ref).norm() / emb_ref.norm()).item(),
                     "logits_rel_fro": ((got - w).norm() / w.norm()).item(), "mask_iou_min": min(ious), "mask_iou": ious})
        print(json.dumps(rows[-1]), flush=True)
for dt in ("float16", "bfloat16"):
    sel = [r for r in rows if r["dtype"] == dt]
    print(f"{dt}: min IoU over {len(sel)} inputs x 3 prompts = {min(r['mask_iou_min'] for r in sel):.5f}, "
          f"max embeddings rel-Fro = {max(r['emb_rel_fro'] for r in sel):.2e}")
