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
        self._index_cache: dict = {}      # (device, prompt counts) -> int32 image index of every prompt, on the device

    def _image_index(self, counts, device) -> torch.Tensor:
        key = (str(device), tuple(counts))
        idx = self._index_cache.get(key)
        if idx is None:
            if len(self._index_cache) > 64:
                self._index_cache.clear()
            idx = torch.repeat_interleave(torch.arange(len(counts), dtype=torch.int32),
                                          torch.tensor(counts, dtype=torch.int64)).to(device)
            self._index_cache[key] = idx
        return idx

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
        index = self._image_index(counts, emb.device)
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

    def host_pipeline(self, depth: int = 2) -> "HostPipeline":
        """A double-buffered host-to-host runner over this path (see HostPipeline)."""
        return HostPipeline(self, depth)

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


class HostPipeline:
    """Streams batches that live in (pinned) HOST memory through a GroundingPath and returns the mask logits to host
    memory, the way an evaluation loop consumes the path (eval_referseg.py:130-211 moves every batch to the GPU and
    every prediction back):  the upload of batch i+1 (copy-in stream) and the download of batch i-1 (copy-out
    stream) overlap the kernels of batch i (the caller's current stream).  Results are bit-identical to calling the
    path on device tensors; nothing is skipped -- every batch pays its own H2D and D2H.

        pipe = path.host_pipeline()
        for images, segs, host_out in batches:             # pinned tensors
            pipe.submit(images, segs, input_sizes, original_sizes, host_out)
        pipe.drain()                                       # all host_out buffers are complete after this
    """

    def __init__(self, path: GroundingPath, depth: int = 2):
        if depth < 1:
            raise ValueError("depth must be >= 1")
        self.path = path
        self.depth = depth
        self._slots: list = []
        self._n = 0
        self._up = None
        self._down = None

    def _slot(self, dev):
        if self._up is None:
            self._up, self._down = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
            self._slots = [{"img": None, "seg": None, "done": None, "down": None} for _ in range(self.depth)]
        s = self._slots[self._n % self.depth]
        self._n += 1
        return s

    @torch.no_grad()
    def submit(self, host_images: torch.Tensor, host_seg: torch.Tensor, input_sizes, original_sizes,
               host_out: torch.Tensor, multimask_output: bool = False, device=None) -> None:
        """host_images [B,3,1024,1024], host_seg [B,n,1,256] (the same number of [SEG] prompts per image), host_out
        [B*n, C, H, W] fp32 -- all pinned host tensors (pageable ones work but serialise the copies)."""
        if host_images.is_cuda or host_seg.is_cuda or host_out.is_cuda:
            raise ValueError("HostPipeline takes host tensors; call the GroundingPath directly for device tensors")
        dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        main = torch.cuda.current_stream(dev)
        slot = self._slot(dev)
        fresh = False
        if slot["img"] is None or slot["img"].shape != host_images.shape or slot["img"].dtype != host_images.dtype:
            slot["img"] = torch.empty(host_images.shape, dtype=host_images.dtype, device=dev)
            fresh = True
        if slot["seg"] is None or slot["seg"].shape != host_seg.shape or slot["seg"].dtype != host_seg.dtype:
            slot["seg"] = torch.empty(host_seg.shape, dtype=host_seg.dtype, device=dev)
            fresh = True
        if slot["down"] is not None:
            slot["down"].synchronize()          # bounds how far the host runs ahead of the device
        if fresh or slot["done"] is None:
            self._up.wait_stream(main)          # new buffers: ordered after whatever used that memory before
        else:
            self._up.wait_event(slot["done"])   # only the kernels that last read THIS slot, not the batch in flight
        with torch.cuda.stream(self._up):
            slot["img"].copy_(host_images, non_blocking=True)
            slot["seg"].copy_(host_seg, non_blocking=True)
        main.wait_stream(self._up)
        B = host_images.shape[0]
        outs = self.path(slot["img"], [slot["seg"][b] for b in range(B)], input_sizes, original_sizes,
                         multimask_output=multimask_output)
        slot["done"] = torch.cuda.Event()
        slot["done"].record(main)
        self._down.wait_event(slot["done"])
        with torch.cuda.stream(self._down):
            o = 0
            for t in outs:
                if t.shape[0]:
                    host_out[o:o + t.shape[0]].copy_(t, non_blocking=True)
                    t.record_stream(self._down)
                    o += t.shape[0]
            slot["down"] = torch.cuda.Event()
            slot["down"].record(self._down)

    def drain(self) -> None:
        """Blocks until every submitted batch has landed in its host_out buffer; the caller's current stream also
        waits for the copy-out stream, so an event recorded after drain() brackets the whole pipeline."""
        if self._down is not None:
            torch.cuda.current_stream().wait_stream(self._down)
            for s in self._slots:
                if s["down"] is not None:
                    s["down"].synchronize()
