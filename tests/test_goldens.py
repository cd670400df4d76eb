"""The oracle must reproduce the committed golden vectors (outputs of the imported reference, see
oracle/make_goldens.py) on any machine -- this is what pins it where /root/reference does not exist."""
import os

import numpy as np
import pytest
import torch

from anyref_b200.synthetic import CONFIGS, synthetic_images, synthetic_seg_embeddings, synthetic_state_dict
from oracle import sam_oracle as O
from oracle.make_goldens import SIZES, sub

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def close(got: torch.Tensor, want: torch.Tensor, tol: float):
    # same torch build => bit-identical; the tolerance only absorbs a different CPU BLAS kernel selection
    assert got.shape == want.shape
    assert (got - want).abs().max().item() <= tol, (got - want).abs().max().item()


def check_sum(t: torch.Tensor, want: dict, rel: float = 1e-5):
    assert tuple(t.shape) == tuple(want["shape"])
    d = t.double()
    assert d.sum().item() == pytest.approx(want["sum"], rel=rel, abs=rel * want["abs_sum"])
    assert d.abs().sum().item() == pytest.approx(want["abs_sum"], rel=rel)
    assert (d * d).sum().item() == pytest.approx(want["sq_sum"], rel=rel)


@torch.no_grad()
@pytest.mark.parametrize("name", ["vit_tiny80", "vit_h"])
def test_oracle_reproduces_goldens(name):
    g = torch.load(os.path.join(GOLD, f"{name}_seed1234_in0.pt"), weights_only=False)
    meta = g["meta"]
    cfg = CONFIGS[name]
    sd = synthetic_state_dict(cfg, seed=meta["seed_ckpt"])
    x = synthetic_images(1, seed=meta["seed_in"])
    seg = synthetic_seg_embeddings(1, meta["n_seg"], seed=meta["seed_in"])[0]
    taps = {}
    emb = O.image_encoder(sd, x, cfg, taps)
    close(sub(emb, (1, 4, 4, 4)), g["emb_sub"], 2e-4)
    check_sum(emb, g["emb_sum"])
    # NB the reference hook on patch_embed fires before +pos_embed
    pos = sd["image_encoder.pos_embed"]
    close(sub(taps["patch_embed"] - pos, (1, 8, 8, 16)), g["tap_patch_sub"], 1e-4)
    close(sub(taps["block0"], (1, 8, 8, 16)), g["tap_block0_sub"], 2e-4)
    close(sub(taps[f"block{cfg.global_attn_indexes[0]}"], (1, 8, 8, 16)), g["tap_block_first_global_sub"], 5e-4)
    close(sub(taps[f"block{cfg.depth - 1}"], (1, 8, 8, 16)), g["tap_block_last_sub"], 2e-3)
    pe = O.dense_pe(sd, cfg)
    close(sub(pe, (1, 8, 8, 8)), g["dense_pe_sub"], 1e-6)
    sparse, dense = O.prompt_encoder(sd, cfg, text_embeds=seg)
    for mm in (False, True):
        tag = "multi" if mm else "single"
        low, iou = O.mask_decoder(sd, cfg, emb, pe, sparse, dense, mm)
        close(sub(low, (1, 1, 4, 4)), g[f"low_{tag}_sub"], 2e-5)
        check_sum(low, g[f"low_{tag}_sum"], rel=1e-4)
        close(iou, g[f"iou_{tag}"], 2e-5)
        for inp, orig in SIZES:
            key = f"post_{tag}_{inp[0]}x{inp[1]}_{orig[0]}x{orig[1]}"
            post = O.postprocess_masks(low, inp, orig, cfg.img_size)
            close(sub(post, (1, 1, 16, 16)), g[key + "_sub"], 2e-5)
            fg = int((post > 0).sum().item())
            assert abs(fg - g[key + "_fg"]) <= max(2, int(2e-5 * post.numel()))
            if not mm:
                bits = torch.from_numpy(np.packbits((post > 0).numpy().reshape(-1)))
                diff = np.unpackbits((bits ^ g[key + "_bits"]).numpy()).sum()
                assert diff <= max(2, int(2e-5 * post.numel()))
