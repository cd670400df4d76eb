"""Diagnostic: fp32 gradients of this path and of autograd-over-the-oracle against an fp64 run of the oracle."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from anyref_b200.segment_anything import build_sam_from_config
from anyref_b200.synthetic import CONFIGS, synthetic_state_dict
from oracle import sam_oracle as O

torch.backends.cuda.matmul.allow_tf32 = False
torch.backends.cudnn.allow_tf32 = False
cfg = CONFIGS["vit_tiny80"]
sd = synthetic_state_dict(cfg)
sam = build_sam_from_config(cfg)
sam.load_state_dict(sd)
sam = sam.cuda()
for p in sam.parameters():
    p.requires_grad_(False)
for p in sam.mask_decoder.parameters():
    p.requires_grad_(True)
sam.mask_decoder.train()
n, k = int(sys.argv[1]), int(sys.argv[2])
g = torch.Generator().manual_seed(5)
emb = (torch.randn(2, 256, 64, 64, generator=g) * 0.5).cuda()[1:2]
g = torch.Generator().manual_seed(100 + 10 * n + k)
sparse0 = torch.randn(n, k, 256, generator=g).cuda()


def run(dtype, mine=False):
    osd = {kk: v.detach().clone().cuda().to(dtype).requires_grad_(kk.startswith("mask_decoder.")) for kk, v in sd.items()}
    with torch.no_grad():
        pe = O.dense_pe({kk: v.cuda() for kk, v in sd.items()}, cfg).to(dtype)
    dense = osd["prompt_encoder.no_mask_embed.weight"].detach().reshape(1, -1, 1, 1).expand(n, -1, 64, 64)
    sp = sparse0.clone().to(dtype).requires_grad_(True)
    total = 0.0
    gg = torch.Generator().manual_seed(7)
    for multi in (False, True):
        if mine:
            m, i = sam.mask_decoder(image_embeddings=emb, image_pe=pe, sparse_prompt_embeddings=sp, dense_prompt_embeddings=dense,
                                    multimask_output=multi)
        else:
            m, i = O.mask_decoder(osd, cfg, emb.to(dtype), pe, sp, dense, multi)
        rm = torch.randn(m.shape, generator=gg).cuda().to(dtype)
        ri = torch.randn(i.shape, generator=gg).cuda().to(dtype)
        total = total + (m * rm).sum() / 256.0 + (i * ri).sum()
    total.backward()
    grads = {"sparse": sp.grad.double()}
    if mine:
        for name, p in sam.mask_decoder.named_parameters():
            grads[name] = p.grad.double()
            p.grad = None
    else:
        for name, _ in sam.mask_decoder.named_parameters():
            grads[name] = osd["mask_decoder." + name].grad.double()
    return grads


ref64 = run(torch.float64)
ref32 = run(torch.float32)
got32 = run(torch.float32, mine=True)
gmax = max(float(v.norm()) for v in ref64.values())
rows = []
for name in ref64:
    den = float(ref64[name].norm()) + 1e-6 * gmax
    rows.append((float((got32[name] - ref64[name]).norm()) / den, float((ref32[name] - ref64[name]).norm()) / den, name))
rows.sort(reverse=True)
print("error vs the fp64 oracle:   this path     torch fp32 autograd")
for a, b, name in rows[:12]:
    print(f"  {name:55s} {a:.3e}   {b:.3e}")
sp = (got32["sparse"] - ref64["sparse"]).reshape(n, -1).norm(dim=1) / ref64["sparse"].reshape(n, -1).norm(dim=1)
sq = (ref32["sparse"] - ref64["sparse"]).reshape(n, -1).norm(dim=1) / ref64["sparse"].reshape(n, -1).norm(dim=1)
print("sparse grad per prompt: this path", sp.tolist(), " torch fp32", sq.tolist())
