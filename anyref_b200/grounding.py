"""The grounding hot path as AnyRef drives it (model/anyref.py:793-819 `generate`, :877-905 `evaluate`):

    image_embeddings = visual_model.image_encoder(sam_images)
    for every image b:  prompt_encoder(text_embeds=[SEG] embeddings) -> mask_decoder(...) -> postprocess_masks(...)

`GroundingPath` runs the same module calls, but decodes the prompts of ALL images of the batch with one
`MaskDecoder.forward_batched` call and one post-processing launch per distinct (input_size, original_size) pair
instead of a Python loop per image.  Results are identical to the per-image loop
(tests/test_gpu_path.py::test_batched_equals_per_image_loop).
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import torch

from .segment_anything.modeling import Sam


class GroundingPath:
    def __init__(self, sam: Sam):
        self.sam = sam

    @torch.no_grad()
    def __call__(self, sam_images: torch.Tensor, seg_embeds: Sequence[torch.Tensor],
                 input_sizes: Sequence[Tuple[int, int]], original_sizes: Sequence[Tuple[int, int]],
                 multimask_output: bool = False, return_binary: bool = False):
        """sam_images [B,3,1024,1024]; seg_embeds[b] = [n_b, 1, 256] ([SEG] projections of image b, n_b may be 0);
        sizes per image.  Returns a list (len B) of fp32 logits [n_b, C, H_b, W_b] (C = 1, or 3 with
        multimask_output); with return_binary=True a list of (logits, uint8 masks) pairs."""
        sam = self.sam
        B = sam_images.shape[0]
        if not (len(seg_embeds) == len(input_sizes) == len(original_sizes) == B):
            raise ValueError("one entry per image is required for seg_embeds / input_sizes / original_sizes")
        emb = sam.image_encoder(sam_images)
        counts = [int(s.shape[0]) for s in seg_embeds]
        total = sum(counts)
        outs: List = [None] * B
        if total == 0:
            for b in range(B):
                H, W = original_sizes[b]
                z = torch.empty((0, 3 if multimask_output else 1, H, W), device=emb.device, dtype=torch.float32)
                outs[b] = (z, z.to(torch.uint8)) if return_binary else z
            return outs
        text = torch.cat([s.to(emb.device) for s in seg_embeds if s.shape[0] > 0], dim=0)
        sparse, dense = sam.prompt_encoder(points=None, boxes=None, masks=None, text_embeds=text)
        sparse = sparse.to(text.dtype)  # model/anyref.py:806
        index = torch.repeat_interleave(torch.arange(B, dtype=torch.int32),
                                        torch.tensor(counts, dtype=torch.int64)).to(emb.device, non_blocking=True)
        low, _ = sam.mask_decoder.forward_batched(emb, sam.prompt_encoder.get_dense_pe(), sparse, dense, index,
                                                  multimask_output)
        # post-process: one launch per run of images that share (input_size, original_size)
        starts = [0]
        for c in counts:
            starts.append(starts[-1] + c)
        b = 0
        while b < B:
            e = b + 1
            key = (tuple(input_sizes[b]), tuple(original_sizes[b]))
            while e < B and (tuple(input_sizes[e]), tuple(original_sizes[e])) == key:
                e += 1
            seg = low[starts[b]:starts[e]]
            res = sam.postprocess_masks(seg, input_sizes[b], original_sizes[b], return_binary=return_binary)
            for i in range(b, e):
                lo, hi = starts[i] - starts[b], starts[i + 1] - starts[b]
                outs[i] = (res[0][lo:hi], res[1][lo:hi]) if return_binary else res[lo:hi]
            b = e
        return outs

    @torch.no_grad()
    def per_image_loop(self, sam_images, seg_embeds, input_sizes, original_sizes, multimask_output: bool = False):
        """The reference's call sequence verbatim (model/anyref.py:793-819), one decoder call per image."""
        sam = self.sam
        emb = sam.image_encoder(sam_images)
        outs = []
        for b in range(emb.shape[0]):
            sparse, dense = sam.prompt_encoder(points=None, boxes=None, masks=None, text_embeds=seg_embeds[b])
            sparse = sparse.to(seg_embeds[b].dtype)
            low, _ = sam.mask_decoder(image_embeddings=emb[b:b + 1], image_pe=sam.prompt_encoder.get_dense_pe(),
                                      sparse_prompt_embeddings=sparse, dense_prompt_embeddings=dense,
                                      multimask_output=multimask_output)
            outs.append(sam.postprocess_masks(low, input_size=input_sizes[b], original_size=original_sizes[b]))
        return outs
