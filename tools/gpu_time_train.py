"""Training step of the mask branch (model/anyref.py:395-450 without the LLM): decoder forward-with-tape, postprocess, BCE +
dice, backward -- this path against PyTorch autograd over the oracle's modules on the same GPU (fp32, TF32 off)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F

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
osd = {k: v.detach().clone().cuda().float().requires_grad_(k.startswith("mask_decoder.")) for k, v in sd.items()}
n = int(sys.argv[1]) if len(sys.argv) > 1 else 3
emb = torch.randn(1, 256, 64, 64, device="cuda") * 0.5
with torch.no_grad():
    pe = O.dense_pe(osd, cfg)
sparse0 = torch.randn(n, 1, 256, device="cuda")
dense = osd["prompt_encoder.no_mask_embed.weight"].detach().reshape(1, -1, 1, 1).expand(n, -1, 64, 64)
gt = (torch.rand(n, 480, 640, device="cuda") > 0.5).float()


def loss_fn(pm):
    ce = F.binary_cross_entropy_with_logits(pm, gt, reduction="none").flatten(1, 2).mean(1).sum()
    s = pm.sigmoid().flatten(1, 2)
    t = gt.flatten(1, 2)
    dice = (1 - (2 * (s * t).sum(-1) + 1) / (s.sum(-1) + t.sum(-1) + 1)).sum()
    return 2.0 * ce + 0.5 * dice


def mine():
    sp = sparse0.clone().requires_grad_(True)
    low, _ = sam.mask_decoder(image_embeddings=emb, image_pe=pe, sparse_prompt_embeddings=sp, dense_prompt_embeddings=dense,
                              multimask_output=False)
    pm = sam.postprocess_masks(low, input_size=(768, 1024), original_size=(480, 640)).squeeze(1)
    loss_fn(pm).backward()


def theirs():
    sp = sparse0.clone().requires_grad_(True)
    low, _ = O.mask_decoder(osd, cfg, emb, pe, sp, dense, False)
    pm = O.postprocess_masks(low, (768, 1024), (480, 640)).squeeze(1)
    loss_fn(pm).backward()


def timeit(fn, name):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(10):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    print(f"{name}: n={n} prompts, forward + loss + backward median {ts[len(ts) // 2]:.2f} ms, min {ts[0]:.2f} ms", flush=True)


if os.environ.get("TRAIN_PROFILE"):      # under ncu: one warm-up + one step of this path only
    mine()
    torch.cuda.synchronize()
    torch.cuda.nvtx.range_push("step")
    mine()
    torch.cuda.synchronize()
    torch.cuda.nvtx.range_pop()
    sys.exit(0)
timeit(mine, "anyref_b200 training path")
timeit(theirs, "stock PyTorch autograd over the reference modules (fp32)")
