"""PromptEncoder with the reference's parameter tree and AnyRef's `text_embeds` argument
(reference: modeling/prompt_encoder.py:16-229).

On AnyRef's grounding path only `text_embeds` is used (model/anyref.py:413-416, :802-805): the sparse embedding is
the [SEG] projection itself and the dense embedding is a stride-0 broadcast of `no_mask_embed` -- no arithmetic at
all, so that part of forward() is pure tensor plumbing exactly as in the reference (:164-186).  `get_dense_pe`
(:67-76) is a CUDA kernel (`sam_dense_pe`), cached until the gaussian matrix changes.  Point / box prompts (:78-109)
and mask prompts (:111-114) -- used by SamPredictor / convert_avs_masks.py -- are the fused kernels
`sam_prompt_sparse` / `sam_prompt_mask_embed`; points, boxes and text embeddings are written side by side into one
[n, N, C] buffer in the reference's concatenation order.
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

    @_lib.device_scoped
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

    def _get_batch_size(self, points, boxes, masks, text_embeds) -> int:
        """prompt_encoder.py:116-135."""
        if points is not None:
            return points[0].shape[0]
        if boxes is not None:
            return boxes.shape[0]
        if masks is not None:
            return masks.shape[0]
        if text_embeds is not None:
            return text_embeds.shape[0]
        return 1

    def _sparse_tables(self, dev):
        """[5, C] fp32 = point_embeddings[0..3].weight, not_a_point_embed.weight (re-packed when a weight changes)."""
        ws = [e.weight for e in self.point_embeddings] + [self.not_a_point_embed.weight]
        key = tuple((w.data_ptr(), w._version, w.dtype) for w in ws)
        if getattr(self, "_tab_cache", None) is None or self._tab_cache[0] != key:
            tab = torch.cat([w.detach().reshape(1, -1).float() for w in ws], dim=0).contiguous()
            self._tab_cache = (key, tab)
        return self._tab_cache[1]

    def _mask_blob(self):
        """mask_downscaling parameters as one fp32 blob in state_dict order (layout: include/anyref_sam.h)."""
        ps = [p for i in (0, 1, 3, 4, 6) for p in (self.mask_downscaling[i].weight, self.mask_downscaling[i].bias)]
        key = tuple((p.data_ptr(), p._version, p.dtype) for p in ps)
        if getattr(self, "_mask_cache", None) is None or self._mask_cache[0] != key:
            blob = torch.cat([p.detach().reshape(-1).float() for p in ps]).contiguous()
            self._mask_cache = (key, blob)
        return self._mask_cache[1]

    @_lib.device_scoped
    @torch.no_grad()
    def _embed_masks(self, masks: torch.Tensor) -> torch.Tensor:
        """mask_downscaling (prompt_encoder.py:111-114) as one kernel: [n,1,4g,4g] -> [n,C,g,g] in the weights' dtype."""
        _runtime.require_cuda(masks, "PromptEncoder (mask prompt)")
        g = self.image_embedding_size[0]
        if masks.dim() != 4 or masks.shape[1] != 1 or tuple(masks.shape[2:]) != self.mask_input_size:
            raise ValueError(f"mask prompt must be [n,1,{self.mask_input_size[0]},{self.mask_input_size[1]}], got "
                             f"{tuple(masks.shape)}")
        if masks.dtype not in (torch.float16, torch.bfloat16, torch.float32):
            masks = masks.float()
        m = masks.contiguous()
        wdt = self.mask_downscaling[6].weight.dtype
        out = torch.empty((m.shape[0], self.embed_dim, g, g), device=m.device, dtype=wdt)
        blob = self._mask_blob()
        L = _lib.load()
        cin = self.mask_downscaling[3].weight.shape[0]
        if blob.numel() != L.sam_prompt_mask_blob_elems(cin, self.embed_dim):
            raise RuntimeError("mask_downscaling parameter blob has an unexpected size")
        rc = L.sam_prompt_mask_embed(m.data_ptr(), _lib.fmt_of(m.dtype), blob.data_ptr(), cin, out.data_ptr(),
                                     _lib.fmt_of(wdt), m.shape[0], g, self.embed_dim, _lib.stream_ptr(m.device))
        _lib.check(rc, "sam_prompt_mask_embed")
        return out

    @_lib.device_scoped
    def forward(self, points: Optional[Tuple[torch.Tensor, torch.Tensor]], boxes: Optional[torch.Tensor],
                masks: Optional[torch.Tensor], text_embeds: Optional[torch.Tensor]
                ) -> Tuple[torch.Tensor, torch.Tensor]:
        """prompt_encoder.py:140-186.  The [SEG] branch (text_embeds only) is pure tensor plumbing and stays inside
        autograd, as in the reference: gradients of the mask loss reach text_hidden_fcs through `sparse`
        (model/anyref.py:770-806).  The point / box / mask branches run CUDA kernels without a backward; with
        gradients enabled on their inputs they raise instead of silently detaching."""
        if points is not None or boxes is not None or masks is not None:
            if torch.is_grad_enabled() and text_embeds is not None and text_embeds.requires_grad:
                raise NotImplementedError("PromptEncoder: point / box / mask prompts combined with a text embedding "
                                          "that requires grad have no backward here; call under torch.no_grad()")
            with torch.no_grad():
                return self._forward(points, boxes, masks, text_embeds)
        return self._forward(points, boxes, masks, text_embeds)

    def _forward(self, points, boxes, masks, text_embeds):
        bs = self._get_batch_size(points, boxes, masks, text_embeds)
        dev = self._get_device()
        g = self.image_embedding_size
        if points is None and boxes is None:
            if text_embeds is not None:
                # cat(empty fp32 [bs,0,C], text_embeds) promotes to fp32 (prompt_encoder.py:165-177)
                sparse = text_embeds.to(device=dev, dtype=torch.promote_types(torch.float32, text_embeds.dtype))
            else:
                sparse = torch.empty((bs, 0, self.embed_dim), device=dev)
        else:
            _runtime.require_cuda(self.point_embeddings[0].weight, "PromptEncoder (point / box prompt)")
            n_pts = 0 if points is None else points[0].shape[1] + (1 if boxes is None else 0)
            n_box = 0 if boxes is None else 2
            n_txt = 0 if text_embeds is None else text_embeds.shape[1]
            ntot = n_pts + n_box + n_txt
            sparse = torch.empty((bs, ntot, self.embed_dim), device=dev, dtype=torch.float32)
            gauss = self.pe_layer.positional_encoding_gaussian_matrix.float().contiguous()
            tab = self._sparse_tables(dev)
            L = _lib.load()
            st = _lib.stream_ptr(dev)
            H, W = int(self.input_image_size[0]), int(self.input_image_size[1])
            if points is not None:
                coords = points[0].to(device=dev, dtype=torch.float32).contiguous()
                labels = points[1].to(device=dev, dtype=torch.float32).contiguous()
                rc = L.sam_prompt_sparse(coords.data_ptr(), labels.data_ptr(), gauss.data_ptr(), tab.data_ptr(),
                                         sparse.data_ptr(), bs, coords.shape[1], 1 if boxes is None else 0, 0,
                                         self.embed_dim, H, W, ntot, 0, st)
                _lib.check(rc, "sam_prompt_sparse")
            if boxes is not None:
                corners = boxes.to(device=dev, dtype=torch.float32).reshape(-1, 2, 2).contiguous()
                rc = L.sam_prompt_sparse(corners.data_ptr(), None, gauss.data_ptr(), tab.data_ptr(), sparse.data_ptr(),
                                         bs, 2, 0, 1, self.embed_dim, H, W, ntot, n_pts, st)
                _lib.check(rc, "sam_prompt_sparse")
            if text_embeds is not None:
                sparse[:, n_pts + n_box:, :] = text_embeds.to(device=dev, dtype=torch.float32)
        if masks is not None:
            dense = self._embed_masks(masks)
        else:
            dense = self.no_mask_embed.weight.reshape(1, -1, 1, 1).expand(bs, -1, g[0], g[1])
        return sparse, dense
