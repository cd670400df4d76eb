"""C4 (BASELINE.json configs[3]): AnyRef-7B end-to-end referring-segmentation forward on synthetic image + text.

The LLM side stays STOCK PyTorch (transformers LlamaModel with AnyRef-7B's shape: hidden 4096, 32 layers, 32 heads,
random weights -- there are no checkpoints offline; CLIP-L's 256 projected image tokens are random stand-ins, exactly the
tensor `model/anyref.py:718` splices into the sequence).  It hands over what `generate` consumes: last-layer hidden
states [B, L + 255, 4096] and the positions of the [SEG] tokens (model/anyref.py:756-758).  The SAM path -- text_hidden_fcs,
image encoder, prompt encoder, mask decoder, postprocess -- is this repo (anyref_b200.seg_head.SegHead).

    python tools/demo_c4.py [--layers 32] [--batch 2] [--text-len 64]

Prints one JSON line with the wall-time split (CUDA events).  A functional drop-in demonstration, not a roofline number.
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--layers", type=int, default=32)
    ap.add_argument("--batch", type=int, default=2)
    ap.add_argument("--text-len", type=int, default=64)
    ap.add_argument("--steps", type=int, default=3)
    args = ap.parse_args()
    from transformers import LlamaConfig, LlamaModel

    from anyref_b200.seg_head import SegHead, build_text_hidden_fcs
    from anyref_b200.segment_anything import build_sam_vit_h
    from anyref_b200.synthetic import synthetic_images, synthetic_state_dict

    dev = torch.device("cuda")
    dt = torch.bfloat16
    torch.manual_seed(0)
    cfg = LlamaConfig(hidden_size=4096, intermediate_size=11008, num_hidden_layers=args.layers, num_attention_heads=32,
                      vocab_size=32004, max_position_embeddings=2048)
    old = torch.get_default_dtype()
    torch.set_default_dtype(dt)
    with torch.device(dev):
        llm = LlamaModel(cfg)
    torch.set_default_dtype(old)
    llm.eval()
    n_params = sum(p.numel() for p in llm.parameters())

    sam = build_sam_vit_h(None)
    sam.load_state_dict(synthetic_state_dict("vit_h", seed=1234), strict=True)
    sam = sam.to(dev)
    sam.image_encoder.set_operand_dtype(dt)
    fcs = build_text_hidden_fcs(4096, 256).to(dev)
    head = SegHead(sam, fcs)

    B, L = args.batch, args.text_len
    SEG_ID = 32003
    ids = torch.randint(3, 32000, (B, L), device=dev)
    ids[:, L - 8] = SEG_ID                      # every sample asks for one mask ...
    ids[0, L - 3] = SEG_ID                      # ... the first one for two
    image_tokens = torch.randn(B, 256, 4096, device=dev, dtype=dt) * 0.02   # CLIP-L features after mm_projector
    sam_images = synthetic_images(B, seed=0).to(dev, dt)
    sizes = [(1024, 1024)] * B

    def llm_forward():
        emb = llm.embed_tokens(ids)
        # <image> placeholder at position 1 expands to 256 tokens: sequence length L + 255 (model/anyref.py:718)
        x = torch.cat([emb[:, :1], image_tokens, emb[:, 2:]], dim=1)
        out = llm(inputs_embeds=x, output_hidden_states=True, use_cache=False)
        return out.hidden_states[-1]

    def sam_forward(hidden):
        idx = torch.where(ids[:, 1:] == SEG_ID)
        return head(hidden, idx, sam_images, sizes, sizes)

    with torch.no_grad():
        hidden = llm_forward()
        masks = sam_forward(hidden)
        torch.cuda.synchronize()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        t_llm = t_sam = 0.0
        for _ in range(args.steps):
            ev[0].record()
            hidden = llm_forward()
            ev[1].record()
            masks = sam_forward(hidden)
            ev[2].record()
            torch.cuda.synchronize()
            t_llm += ev[0].elapsed_time(ev[1])
            t_sam += ev[1].elapsed_time(ev[2])
    assert hidden.shape == (B, L + 255, 4096)
    assert [tuple(m.shape) for m in masks] == [(2, 1024, 1024)] + [(1, 1024, 1024)] * (B - 1)
    assert all(bool(torch.isfinite(m).all()) for m in masks)
    print(json.dumps({"config": "C4: LLaMA-7B-shaped stock PyTorch LLM (random weights) + CLIP-token stand-ins -> SegHead "
                                "(text_hidden_fcs + SAM ViT-H path on the B200 kernels)",
                      "llm_layers": args.layers, "llm_params": n_params, "batch": B, "seq_len": L + 255,
                      "masks": sum(m.shape[0] for m in masks),
                      "llm_forward_ms": t_llm / args.steps, "sam_path_ms": t_sam / args.steps,
                      "sam_share": t_sam / (t_llm + t_sam)}))


if __name__ == "__main__":
    main()
