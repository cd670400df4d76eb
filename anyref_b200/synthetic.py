"""Deterministic synthetic SAM checkpoints and inputs (no real weights or datasets exist offline).

`sam_tensor_specs` enumerates the state_dict layout of the reference `Sam` module
(/root/reference/model/segment_anything/build_sam.py:56-102 builds it; names/shapes verified against the
reference in tests/test_state_dict_layout.py).  `synthetic_state_dict` fills it from ONE seeded CPU generator in
that fixed order, so the same tensors are produced in this container and on the GPU box:

  * Linear / Conv / ConvTranspose weights and biases: U(-1/sqrt(fan_in), 1/sqrt(fan_in))  (torch's default init)
  * LayerNorm / LayerNorm2d: weight U(0.9, 1.1), bias U(-0.1, 0.1)   (so the affine part is exercised)
  * nn.Embedding tables: N(0, 1);  PE gaussian matrix: N(0, 1)
  * pos_embed and every rel_pos_h / rel_pos_w: trunc-normal(std 0.02) -- the reference zero-initialises them
    (image_encoder.py:70-74, :232-233), which would leave the position paths untested.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Iterator

import torch


@dataclass(frozen=True)
class SamConfig:
    """Hyper-parameters of `_build_sam` (build_sam.py:56-102); defaults = ViT-H."""
    embed_dim: int = 1280
    depth: int = 32
    num_heads: int = 16
    global_attn_indexes: tuple = (7, 15, 23, 31)
    img_size: int = 1024
    patch_size: int = 16
    window_size: int = 14
    mlp_ratio: float = 4.0
    out_chans: int = 256          # == prompt_embed_dim == transformer_dim
    mask_in_chans: int = 16
    dec_depth: int = 2
    dec_heads: int = 8
    dec_mlp_dim: int = 2048
    num_multimask_outputs: int = 3
    iou_head_depth: int = 3
    iou_head_hidden_dim: int = 256

    @property
    def grid(self) -> int:
        return self.img_size // self.patch_size

    @property
    def head_dim(self) -> int:
        return self.embed_dim // self.num_heads


CONFIGS = {
    "vit_h": SamConfig(),
    "vit_l": SamConfig(embed_dim=1024, depth=24, num_heads=16, global_attn_indexes=(5, 11, 17, 23)),
    "vit_b": SamConfig(embed_dim=768, depth=12, num_heads=12, global_attn_indexes=(2, 5, 8, 11)),
    # test-size model with ViT-H's head_dim (80), one windowed + one global block
    "vit_tiny80": SamConfig(embed_dim=160, depth=2, num_heads=2, global_attn_indexes=(1,)),
    # same with ViT-L / ViT-B's head_dim (64)
    "vit_tiny64": SamConfig(embed_dim=128, depth=2, num_heads=2, global_attn_indexes=(1,)),
    # smallest width the folded-LayerNorm GEMMs accept (embed_dim % 256 == 0), head_dim 64
    "vit_tiny256": SamConfig(embed_dim=256, depth=2, num_heads=4, global_attn_indexes=(1,)),
}


def sam_tensor_specs(cfg: SamConfig) -> Iterator[tuple]:
    """Yields (name, shape, kind, fan_in) in state_dict order."""
    E, g, hd = cfg.embed_dim, cfg.grid, cfg.head_dim
    mlp = int(E * cfg.mlp_ratio)
    yield "image_encoder.pos_embed", (1, g, g, E), "pos", 0
    k = 3 * cfg.patch_size * cfg.patch_size
    yield "image_encoder.patch_embed.proj.weight", (E, 3, cfg.patch_size, cfg.patch_size), "w", k
    yield "image_encoder.patch_embed.proj.bias", (E,), "b", k
    for i in range(cfg.depth):
        p = f"image_encoder.blocks.{i}."
        s = g if i in cfg.global_attn_indexes else cfg.window_size
        yield p + "norm1.weight", (E,), "ln_w", 0
        yield p + "norm1.bias", (E,), "ln_b", 0
        yield p + "attn.rel_pos_h", (2 * s - 1, hd), "pos", 0
        yield p + "attn.rel_pos_w", (2 * s - 1, hd), "pos", 0
        yield p + "attn.qkv.weight", (3 * E, E), "w", E
        yield p + "attn.qkv.bias", (3 * E,), "b", E
        yield p + "attn.proj.weight", (E, E), "w", E
        yield p + "attn.proj.bias", (E,), "b", E
        yield p + "norm2.weight", (E,), "ln_w", 0
        yield p + "norm2.bias", (E,), "ln_b", 0
        yield p + "mlp.lin1.weight", (mlp, E), "w", E
        yield p + "mlp.lin1.bias", (mlp,), "b", E
        yield p + "mlp.lin2.weight", (E, mlp), "w", mlp
        yield p + "mlp.lin2.bias", (E,), "b", mlp
    C = cfg.out_chans
    yield "image_encoder.neck.0.weight", (C, E, 1, 1), "w", E
    yield "image_encoder.neck.1.weight", (C,), "ln_w", 0
    yield "image_encoder.neck.1.bias", (C,), "ln_b", 0
    yield "image_encoder.neck.2.weight", (C, C, 3, 3), "w", 9 * C
    yield "image_encoder.neck.3.weight", (C,), "ln_w", 0
    yield "image_encoder.neck.3.bias", (C,), "ln_b", 0
    # prompt encoder (prompt_encoder.py:39-64)
    yield "prompt_encoder.pe_layer.positional_encoding_gaussian_matrix", (2, C // 2), "gauss", 0
    for i in range(4):
        yield f"prompt_encoder.point_embeddings.{i}.weight", (1, C), "embed", 0
    yield "prompt_encoder.not_a_point_embed.weight", (1, C), "embed", 0
    m4, m = cfg.mask_in_chans // 4, cfg.mask_in_chans
    yield "prompt_encoder.mask_downscaling.0.weight", (m4, 1, 2, 2), "w", 4
    yield "prompt_encoder.mask_downscaling.0.bias", (m4,), "b", 4
    yield "prompt_encoder.mask_downscaling.1.weight", (m4,), "ln_w", 0
    yield "prompt_encoder.mask_downscaling.1.bias", (m4,), "ln_b", 0
    yield "prompt_encoder.mask_downscaling.3.weight", (m, m4, 2, 2), "w", 4 * m4
    yield "prompt_encoder.mask_downscaling.3.bias", (m,), "b", 4 * m4
    yield "prompt_encoder.mask_downscaling.4.weight", (m,), "ln_w", 0
    yield "prompt_encoder.mask_downscaling.4.bias", (m,), "ln_b", 0
    yield "prompt_encoder.mask_downscaling.6.weight", (C, m, 1, 1), "w", m
    yield "prompt_encoder.mask_downscaling.6.bias", (C,), "b", m
    yield "prompt_encoder.no_mask_embed.weight", (1, C), "embed", 0

    # mask decoder (mask_decoder.py:44-73, transformer.py:35-60, :129-147, :197-210)
    def attn(prefix, internal):
        for nm, (o, i) in (("q_proj", (internal, C)), ("k_proj", (internal, C)), ("v_proj", (internal, C)),
                           ("out_proj", (C, internal))):
            yield prefix + nm + ".weight", (o, i), "w", i
            yield prefix + nm + ".bias", (o,), "b", i

    for l in range(cfg.dec_depth):
        p = f"mask_decoder.transformer.layers.{l}."
        yield from attn(p + "self_attn.", C)
        yield p + "norm1.weight", (C,), "ln_w", 0
        yield p + "norm1.bias", (C,), "ln_b", 0
        yield from attn(p + "cross_attn_token_to_image.", C // 2)
        yield p + "norm2.weight", (C,), "ln_w", 0
        yield p + "norm2.bias", (C,), "ln_b", 0
        yield p + "mlp.lin1.weight", (cfg.dec_mlp_dim, C), "w", C
        yield p + "mlp.lin1.bias", (cfg.dec_mlp_dim,), "b", C
        yield p + "mlp.lin2.weight", (C, cfg.dec_mlp_dim), "w", cfg.dec_mlp_dim
        yield p + "mlp.lin2.bias", (C,), "b", cfg.dec_mlp_dim
        yield p + "norm3.weight", (C,), "ln_w", 0
        yield p + "norm3.bias", (C,), "ln_b", 0
        yield p + "norm4.weight", (C,), "ln_w", 0
        yield p + "norm4.bias", (C,), "ln_b", 0
        yield from attn(p + "cross_attn_image_to_token.", C // 2)
    yield from attn("mask_decoder.transformer.final_attn_token_to_image.", C // 2)
    yield "mask_decoder.transformer.norm_final_attn.weight", (C,), "ln_w", 0
    yield "mask_decoder.transformer.norm_final_attn.bias", (C,), "ln_b", 0
    yield "mask_decoder.iou_token.weight", (1, C), "embed", 0
    nm = cfg.num_multimask_outputs + 1
    yield "mask_decoder.mask_tokens.weight", (nm, C), "embed", 0
    # ConvTranspose2d weight is [in, out, kh, kw]; torch computes fan_in from dim 1 -> out*kh*kw
    yield "mask_decoder.output_upscaling.0.weight", (C, C // 4, 2, 2), "w", (C // 4) * 4
    yield "mask_decoder.output_upscaling.0.bias", (C // 4,), "b", (C // 4) * 4
    yield "mask_decoder.output_upscaling.1.weight", (C // 4,), "ln_w", 0
    yield "mask_decoder.output_upscaling.1.bias", (C // 4,), "ln_b", 0
    yield "mask_decoder.output_upscaling.3.weight", (C // 4, C // 8, 2, 2), "w", (C // 8) * 4
    yield "mask_decoder.output_upscaling.3.bias", (C // 8,), "b", (C // 8) * 4
    for i in range(nm):
        p = f"mask_decoder.output_hypernetworks_mlps.{i}.layers."
        for j, (o, ii) in enumerate(((C, C), (C, C), (C // 8, C))):
            yield p + f"{j}.weight", (o, ii), "w", ii
            yield p + f"{j}.bias", (o,), "b", ii
    h = cfg.iou_head_hidden_dim
    dims = [C] + [h] * (cfg.iou_head_depth - 1) + [nm]
    for j in range(cfg.iou_head_depth):
        yield f"mask_decoder.iou_prediction_head.layers.{j}.weight", (dims[j + 1], dims[j]), "w", dims[j]
        yield f"mask_decoder.iou_prediction_head.layers.{j}.bias", (dims[j + 1],), "b", dims[j]


def synthetic_state_dict(cfg: SamConfig | str = "vit_h", seed: int = 1234) -> dict:
    """fp32 CPU state_dict with the reference layout, deterministic in (cfg, seed)."""
    if isinstance(cfg, str):
        cfg = CONFIGS[cfg]
    g = torch.Generator(device="cpu")
    g.manual_seed(seed)
    sd = {}
    for name, shape, kind, fan_in in sam_tensor_specs(cfg):
        if kind in ("w", "b"):
            bound = 1.0 / math.sqrt(fan_in)
            t = (torch.rand(shape, generator=g, dtype=torch.float32) * 2.0 - 1.0) * bound
        elif kind == "ln_w":
            t = 0.9 + 0.2 * torch.rand(shape, generator=g, dtype=torch.float32)
        elif kind == "ln_b":
            t = -0.1 + 0.2 * torch.rand(shape, generator=g, dtype=torch.float32)
        elif kind in ("embed", "gauss"):
            t = torch.randn(shape, generator=g, dtype=torch.float32)
        elif kind == "pos":
            t = (torch.randn(shape, generator=g, dtype=torch.float32) * 0.02).clamp_(-0.04, 0.04)
        else:  # pragma: no cover
            raise ValueError(kind)
        sd[name] = t
    return sd


def synthetic_images(batch: int, seed: int = 0, img_size: int = 1024) -> torch.Tensor:
    """`sam_images` as the data pipeline would hand them over: normalised + padded, [B,3,S,S] fp32."""
    g = torch.Generator(device="cpu")
    g.manual_seed(seed)
    return torch.randn((batch, 3, img_size, img_size), generator=g, dtype=torch.float32)


def synthetic_seg_embeddings(batch: int, n_seg: int, seed: int = 0, dim: int = 256) -> torch.Tensor:
    """[SEG] embeddings (output of text_hidden_fcs, model/anyref.py:770): [B, n_seg, 1, dim] fp32."""
    g = torch.Generator(device="cpu")
    g.manual_seed(seed + 7919)
    return torch.randn((batch, n_seg, 1, dim), generator=g, dtype=torch.float32)
