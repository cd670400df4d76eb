"""Generates tests/golden/*.pt by running the UNMODIFIED reference modules imported from /root/reference/model
(segment_anything) on the deterministic synthetic checkpoint / inputs of anyref_b200.synthetic.

    PYTHONDONTWRITEBYTECODE=1 python -m oracle.make_goldens          (only works where /root/reference exists)

The reference ships no golden vectors for this path, so these files are the pin: the oracle
(oracle/sam_oracle.py) must reproduce them on any machine (tests/test_goldens.py), and the CUDA path is compared with
the oracle.  To keep fixtures small only strided sub-samples, bit-packed binary masks and fp64 checksums are stored.
"""
from __future__ import annotations

import os
import sys
import time

import numpy as np
import torch

sys.dont_write_bytecode = True
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference/model")

from anyref_b200.synthetic import CONFIGS, synthetic_images, synthetic_seg_embeddings, synthetic_state_dict  # noqa
from tests.refutil import build_reference_sam  # noqa

OUT = os.path.join(ROOT, "tests", "golden")

# the (input_size, original_size) pairs exercised (square, portrait crop + downsample, landscape)
SIZES = [((1024, 1024), (1024, 1024)), ((1024, 683), (640, 427)), ((768, 1024), (480, 640))]


def checksum(t: torch.Tensor) -> dict:
    d = t.double()
    return {"sum": d.sum().item(), "abs_sum": d.abs().sum().item(), "sq_sum": (d * d).sum().item(),
            "shape": tuple(t.shape)}


def sub(t: torch.Tensor, steps) -> torch.Tensor:
    idx = tuple(slice(None, None, s) for s in steps)
    return t[idx].clone()


def pack_mask(logits: torch.Tensor) -> torch.Tensor:
    return torch.from_numpy(np.packbits((logits > 0).numpy().reshape(-1)))


@torch.no_grad()
def run(name: str, seed_ckpt: int = 1234, seed_in: int = 0, n_seg: int = 2) -> dict:
    import segment_anything as sa  # the reference package

    cfg = CONFIGS[name]
    sd = synthetic_state_dict(cfg, seed=seed_ckpt)
    ref = build_reference_sam(sa, cfg)
    ref.load_state_dict(sd, strict=True)
    x = synthetic_images(1, seed=seed_in)
    seg = synthetic_seg_embeddings(1, n_seg, seed=seed_in)[0]

    taps = {}
    hooks = []
    watch = {"patch": ref.image_encoder.patch_embed, "block0": ref.image_encoder.blocks[0],
             "block_first_global": ref.image_encoder.blocks[cfg.global_attn_indexes[0]],
             "block_last": ref.image_encoder.blocks[cfg.depth - 1]}
    for k, m in watch.items():
        hooks.append(m.register_forward_hook(lambda mod, i, o, k=k: taps.__setitem__(k, o.detach())))
    t0 = time.time()
    emb = ref.image_encoder(x)
    enc_s = time.time() - t0
    for h in hooks:
        h.remove()

    g = {"meta": {"config": name, "seed_ckpt": seed_ckpt, "seed_in": seed_in, "n_seg": n_seg,
                  "torch": torch.__version__, "encoder_seconds": enc_s,
                  "generator": "oracle/make_goldens.py (reference modules from /root/reference/model/segment_anything)"}}
    g["emb_sub"] = sub(emb, (1, 4, 4, 4))
    g["emb_sum"] = checksum(emb)
    for k, v in taps.items():
        g[f"tap_{k}_sub"] = sub(v, (1, 8, 8, 16))
        g[f"tap_{k}_sum"] = checksum(v)

    pe = ref.prompt_encoder.get_dense_pe()
    g["dense_pe_sub"] = sub(pe, (1, 8, 8, 8))
    g["dense_pe_sum"] = checksum(pe)
    sparse, dense = ref.prompt_encoder(points=None, boxes=None, masks=None, text_embeds=seg)
    for mm in (False, True):
        low, iou = ref.mask_decoder(image_embeddings=emb, image_pe=pe, sparse_prompt_embeddings=sparse,
                                    dense_prompt_embeddings=dense, multimask_output=mm)
        tag = "multi" if mm else "single"
        g[f"low_{tag}_sub"] = sub(low, (1, 1, 4, 4))
        g[f"low_{tag}_sum"] = checksum(low)
        g[f"iou_{tag}"] = iou.clone()
        for inp, orig in SIZES:
            post = ref.postprocess_masks(low, input_size=inp, original_size=orig)
            key = f"post_{tag}_{inp[0]}x{inp[1]}_{orig[0]}x{orig[1]}"
            g[key + "_sub"] = sub(post, (1, 1, 16, 16))
            g[key + "_sum"] = checksum(post)
            g[key + "_fg"] = int((post > 0).sum().item())
            if not mm:
                g[key + "_bits"] = pack_mask(post)
    return g


def main():
    os.makedirs(OUT, exist_ok=True)
    for name in (sys.argv[1:] or ["vit_tiny80", "vit_h"]):
        t0 = time.time()
        g = run(name)
        path = os.path.join(OUT, f"{name}_seed1234_in0.pt")
        torch.save(g, path)
        print(f"{name}: wrote {path} ({os.path.getsize(path) / 1024:.0f} KiB) in {time.time() - t0:.1f}s; "
              f"emb std {g['emb_sum']['sq_sum'] ** 0.5:.3f}")


if __name__ == "__main__":
    main()
