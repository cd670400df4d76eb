"""Helpers that build the REFERENCE modules (imported from /root/reference) for a given SamConfig."""
from functools import partial

import torch


def build_reference_sam(sa, cfg):
    """Mirror of the reference's _build_sam argument wiring (build_sam.py:56-102) for an arbitrary SamConfig."""
    from segment_anything.modeling import ImageEncoderViT, MaskDecoder, PromptEncoder, Sam, TwoWayTransformer

    g = cfg.grid
    sam = Sam(
        image_encoder=ImageEncoderViT(depth=cfg.depth, embed_dim=cfg.embed_dim, img_size=cfg.img_size,
                                      mlp_ratio=cfg.mlp_ratio, norm_layer=partial(torch.nn.LayerNorm, eps=1e-6),
                                      num_heads=cfg.num_heads, patch_size=cfg.patch_size, qkv_bias=True,
                                      use_rel_pos=True, global_attn_indexes=list(cfg.global_attn_indexes),
                                      window_size=cfg.window_size, out_chans=cfg.out_chans),
        prompt_encoder=PromptEncoder(embed_dim=cfg.out_chans, image_embedding_size=(g, g),
                                     input_image_size=(cfg.img_size, cfg.img_size), mask_in_chans=cfg.mask_in_chans),
        mask_decoder=MaskDecoder(num_multimask_outputs=cfg.num_multimask_outputs,
                                 transformer=TwoWayTransformer(depth=cfg.dec_depth, embedding_dim=cfg.out_chans,
                                                               mlp_dim=cfg.dec_mlp_dim, num_heads=cfg.dec_heads),
                                 transformer_dim=cfg.out_chans, iou_head_depth=cfg.iou_head_depth,
                                 iou_head_hidden_dim=cfg.iou_head_hidden_dim),
    )
    sam.eval()
    return sam
