"""Diagnostic run on a B200: the module tree (encoder / prompt encoder / decoder / postprocess) vs the CPU oracle
(vit_tiny80, full tensors) and vs the committed reference goldens (vit_h).  Prints, never asserts."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from anyref_b200.segment_anything import build_sam_from_config
from anyref_b200.synthetic import CONFIGS, synthetic_images, synthetic_seg_embeddings, synthetic_state_dict
from oracle import sam_oracle as O
from oracle.make_goldens import SIZES, sub

GOLD = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def stats(name, got, ref):
    got, ref = got.float().cpu(), ref.float().cpu()
    d = got - ref
    print(f"{name}: max_abs={d.abs().max().item():.3e} rel_fro={(d.norm() / ref.norm()).item():.3e} "
          f"ref_max={ref.abs().max().item():.3f} finite={bool(torch.isfinite(got).all())}", flush=True)


def iou(a, b):
    a, b = a.cpu() > 0, b.cpu() > 0
    return ((a & b).sum().item() + 1e-9) / ((a | b).sum().item() + 1e-9)


@torch.no_grad()
def tiny():
    cfg = CONFIGS["vit_tiny80"]
    sd = synthetic_state_dict(cfg)
    x = synthetic_images(2, seed=0)
    seg = synthetic_seg_embeddings(2, 3, seed=0)
    taps = {}
    t0 = time.time()
    emb_ref = O.image_encoder(sd, x, cfg, taps)
    print(f"oracle encoder (tiny80, B=2) {time.time() - t0:.1f}s")
    sam = build_sam_from_config(cfg)
    sam.load_state_dict(sd, strict=True)
    sam = sam.cuda()
    xc = x.cuda()
    for dt in (torch.float16, torch.bfloat16):
        sam.image_encoder.set_operand_dtype(dt)
        for blk in (0, 1):
            tap = torch.empty(2 * 4096, cfg.embed_dim, device="cuda")
            sam.image_encoder(xc, _tap=(blk, tap))
            stats(f"tiny80 op={str(dt)[6:]} block{blk}", tap.view(2, 64, 64, -1), taps[f"block{blk}"])
        emb = sam.image_encoder(xc)
        stats(f"tiny80 op={str(dt)[6:]} embeddings", emb, emb_ref)
    # decoder + postprocess on the ORACLE embeddings (isolates the decoder)
    pe_ref = O.dense_pe(sd, cfg)
    pe = sam.prompt_encoder.get_dense_pe()
    stats("dense_pe", pe, pe_ref)
    for b in range(2):
        sparse_ref, dense_ref = O.prompt_encoder(sd, cfg, text_embeds=seg[b])
        sparse, dense = sam.prompt_encoder(points=None, boxes=None, masks=None, text_embeds=seg[b].cuda())
        for mm in (False, True):
            low_ref, iou_ref = O.mask_decoder(sd, cfg, emb_ref[b:b + 1], pe_ref, sparse_ref, dense_ref, mm)
            low, iou_p = sam.mask_decoder(image_embeddings=emb_ref[b:b + 1].cuda(), image_pe=pe, sparse_prompt_embeddings=sparse,
                                          dense_prompt_embeddings=dense, multimask_output=mm)
            stats(f"decoder img{b} multimask={mm} low_res", low, low_ref)
            stats(f"decoder img{b} multimask={mm} iou", iou_p, iou_ref)
            for inp, orig in SIZES + [((1024, 1024), (333, 517))]:
                post_ref = O.postprocess_masks(low_ref, inp, orig, 1024)
                post, binm = sam.postprocess_masks(low_ref.cuda(), inp, orig, return_binary=True)
                d = (post.cpu() - post_ref).abs().max().item()
                flips = ((post.cpu() > 0) != (post_ref > 0)).sum().item()
                bin_ok = bool(((post > 0).to(torch.uint8) == binm).all())
                print(f"  postprocess {inp}->{orig}: max_abs={d:.2e} sign_flips={flips}/{post_ref.numel()} binary_consistent={bin_ok}")
    # batched decoder entry: all prompts of both images in one call
    sp = torch.cat([seg[0], seg[1]]).cuda()
    sparse, dense = sam.prompt_encoder(points=None, boxes=None, masks=None, text_embeds=sp)
    idx = torch.tensor([0, 0, 0, 1, 1, 1], dtype=torch.int32, device="cuda")
    low, _ = sam.mask_decoder.forward_batched(emb_ref.cuda(), pe, sparse, dense, idx, True)
    refs = []
    for b in range(2):
        s_r, d_r = O.prompt_encoder(sd, cfg, text_embeds=seg[b])
        refs.append(O.mask_decoder(sd, cfg, emb_ref[b:b + 1], pe_ref, s_r, d_r, True)[0])
    stats("decoder batched (2 images x 3 prompts)", low, torch.cat(refs))


@torch.no_grad()
def vit_h():
    g = torch.load(os.path.join(GOLD, "vit_h_seed1234_in0.pt"), weights_only=False)
    cfg = CONFIGS["vit_h"]
    sd = synthetic_state_dict(cfg, seed=g["meta"]["seed_ckpt"])
    x = synthetic_images(1, seed=g["meta"]["seed_in"])
    seg = synthetic_seg_embeddings(1, g["meta"]["n_seg"], seed=g["meta"]["seed_in"])[0]
    sam = build_sam_from_config(cfg)
    sam.load_state_dict(sd, strict=True)
    sam = sam.cuda()
    xc = x.cuda()
    for dt in (torch.float16, torch.bfloat16):
        sam.image_encoder.set_operand_dtype(dt)
        for name, blk in (("block0", 0), ("block_first_global", 7), ("block_last", 31)):
            tap = torch.empty(4096, cfg.embed_dim, device="cuda")
            sam.image_encoder(xc, _tap=(blk, tap))
            stats(f"vit_h op={str(dt)[6:]} {name}", sub(tap.view(1, 64, 64, -1), (1, 8, 8, 16)), g[f"tap_{name}_sub"])
        torch.cuda.synchronize()
        t0 = time.time()
        emb = sam.image_encoder(xc)
        torch.cuda.synchronize()
        print(f"vit_h encoder B=1 wall {1e3 * (time.time() - t0):.1f} ms")
        stats(f"vit_h op={str(dt)[6:]} embeddings(sub)", sub(emb, (1, 4, 4, 4)), g["emb_sub"])
        pe = sam.prompt_encoder.get_dense_pe()
        sparse, dense = sam.prompt_encoder(points=None, boxes=None, masks=None, text_embeds=seg.cuda())
        for mm in (False, True):
            tag = "multi" if mm else "single"
            low, iou_p = sam.mask_decoder(image_embeddings=emb, image_pe=pe, sparse_prompt_embeddings=sparse,
                                          dense_prompt_embeddings=dense, multimask_output=mm)
            stats(f"vit_h op={str(dt)[6:]} low_{tag}(sub)", sub(low, (1, 1, 4, 4)), g[f"low_{tag}_sub"])
            stats(f"vit_h op={str(dt)[6:]} iou_{tag}", iou_p, g[f"iou_{tag}"])
            if not mm:
                for inp, orig in SIZES:
                    key = f"post_{tag}_{inp[0]}x{inp[1]}_{orig[0]}x{orig[1]}"
                    post = sam.postprocess_masks(low, inp, orig)
                    bits = torch.from_numpy(np.packbits((post > 0).cpu().numpy().reshape(-1)))
                    want = np.unpackbits(g[key + "_bits"].numpy())[:post.numel()].reshape(post.shape).astype(bool)
                    got = (post > 0).cpu().numpy()
                    for i in range(post.shape[0]):
                        inter = (want[i] & got[i]).sum()
                        union = (want[i] | got[i]).sum()
                        print(f"  vit_h op={str(dt)[6:]} {key} mask{i} IoU={inter / max(union, 1):.5f} fg_ref={want[i].mean():.3f}")


if __name__ == "__main__":
    print(torch.cuda.get_device_name(0))
    which = sys.argv[1:] or ["tiny", "vit_h"]
    if "tiny" in which:
        tiny()
    if "vit_h" in which:
        vit_h()
