"""PromptEncoder with the reference's parameter tree and AnyRef's `text_embeds` argument
(reference: modeling/prompt_encoder.py:16-229).

On AnyRef's grounding path only `text_embeds` is used (model/anyref.py:413-416, :802-805): the sparse embedding is
the [SEG] projection itself and the dense embedding is a stride-0 broadcast of `no_mask_embed` -- no arithmetic at
all, so forward() is pure tensor plumbing exactly as in the reference (:164-186).  `get_dense_pe` (:67-76) is a CUDA
kernel (`sam_dense_pe`), cached until the gaussian matrix changes.  Point / box / mask prompts are the next scope
row and raise NotImplementedError.
"""
from __future__ import annotations

from typing import Optional, Tuple, Type

import numpy as np
import torch
import torch.nn as nn

from .. import _runtime
from ... import _lib
from .common import LayerNorm2d


class PositionEmbeddingRandom(nn.Module):
    def __init__(self, num_pos_feats: int = 64, scale: Optional[float] = None) -> None:
        super().__init__()
        if scale is None or scale <= 0.0:
            scale = 1.0
        self.register_buffer("positional_encoding_gaussian_matrix", scale * torch.randn((2, num_pos_feats)))
        self._cache = None

    @torch.no_grad()
    def forward(self, size: Tuple[int, int]) -> torch.Tensor:
        """[C, h, w] grid encoding (prompt_encoder.py:216-229)."""
        h, w = size
        if h != w:
            raise NotImplementedError("square embedding grids only")
        gm = self.positional_encoding_gaussian_matrix
        _runtime.require_cuda(gm, "PositionEmbeddingRandom")
        key = (gm.data_ptr(), gm._version, gm.dtype, h)
        if self._cache is None or self._cache[0] != key:
            Cc = 2 * gm.shape[1]
            out = torch.empty((Cc, h, w), device=gm.device, dtype=gm.dtype)
            g32 = gm.float().contiguous()
            rc = _lib.load().sam_dense_pe(g32.data_ptr(), out.data_ptr(), _lib.fmt_of(out.dtype), Cc, h,
                                          _lib.stream_ptr(gm.device))
            _lib.check(rc, "sam_dense_pe")
            self._cache = (key, out)
        return self._cache[1]


class PromptEncoder(nn.Module):
    def __init__(self, embed_dim: int, image_embedding_size: Tuple[int, int], input_image_size: Tuple[int, int],
                 mask_in_chans: int, activation: Type[nn.Module] = nn.GELU) -> None:
        super().__init__()
        self.embed_dim = embed_dim
        self.input_image_size = input_image_size
        self.image_embedding_size = image_embedding_size
        self.pe_layer = PositionEmbeddingRandom(embed_dim // 2)
        self.num_point_embeddings: int = 4  # pos/neg point + 2 box corners
        self.point_embeddings = nn.ModuleList([nn.Embedding(1, embed_dim) for _ in range(self.num_point_embeddings)])
        self.not_a_point_embed = nn.Embedding(1, embed_dim)
        self.mask_input_size = (4 * image_embedding_size[0], 4 * image_embedding_size[1])
        self.mask_downscaling = nn.Sequential(
            nn.Conv2d(1, mask_in_chans // 4, kernel_size=2, stride=2),
            LayerNorm2d(mask_in_chans // 4),
            activation(),
            nn.Conv2d(mask_in_chans // 4, mask_in_chans, kernel_size=2, stride=2),
            LayerNorm2d(mask_in_chans),
            activation(),
            nn.Conv2d(mask_in_chans, embed_dim, kernel_size=1),
        )
        self.no_mask_embed = nn.Embedding(1, embed_dim)

    def get_dense_pe(self) -> torch.Tensor:
        """[1, C, g, g] positional encoding of the image embedding grid (prompt_encoder.py:67-76)."""
        return self.pe_layer(self.image_embedding_size).unsqueeze(0)

    def _get_device(self) -> torch.device:
        return self.point_embeddings[0].weight.device

    def forward(self, points: Optional[Tuple[torch.Tensor, torch.Tensor]], boxes: Optional[torch.Tensor],
                masks: Optional[torch.Tensor], text_embeds: Optional[torch.Tensor]
                ) -> Tuple[torch.Tensor, torch.Tensor]:
        if points is not None or boxes is not None or masks is not None:
            raise NotImplementedError("point / box / mask prompts are outside AnyRef's grounding path "
                                      "(model/anyref.py:802 passes text_embeds only); see SURVEY 8(f)-2")
        bs = text_embeds.shape[0] if text_embeds is not None else 1
        dev = self._get_device()
        if text_embeds is not None:
            # cat(empty fp32 [bs,0,C], text_embeds) promotes to fp32 (prompt_encoder.py:165-177)
            sparse = text_embeds.to(device=dev, dtype=torch.promote_types(torch.float32, text_embeds.dtype))
        else:
            sparse = torch.empty((bs, 0, self.embed_dim), device=dev)
        dense = self.no_mask_embed.weight.reshape(1, -1, 1, 1).expand(bs, -1, self.image_embedding_size[0],
                                                                      self.image_embedding_size[1])
        return sparse, dense
