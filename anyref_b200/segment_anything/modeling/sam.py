"""Sam wrapper (reference: modeling/sam.py:18-184): holds the three modules, `mask_threshold`, the pixel
normalisation buffers, and `postprocess_masks` -- here one fused CUDA kernel (`sam_postprocess_masks`) instead of two
F.interpolate calls with a [n,C,1024,1024] intermediate (sam.py:159-172).

`Sam.forward` (:54-135) is dead code on AnyRef's path and broken in the reference (it calls the prompt encoder without
the `text_embeds` argument AnyRef added, prompt_encoder.py:139-145 -> TypeError).  It is provided here
in working form -- same records in, same records out, plus an optional 'text_embeds' key -- for users of upstream SAM.
"""
from __future__ import annotations

from typing import Any, Dict, List, Optional, Tuple

import torch
import torch.nn as nn

from .. import _runtime
from ... import _lib
from .image_encoder import ImageEncoderViT
from .mask_decoder import MaskDecoder
from .prompt_encoder import PromptEncoder


class Sam(nn.Module):
    mask_threshold: float = 0.0
    image_format: str = "RGB"

    def __init__(self, image_encoder: ImageEncoderViT, prompt_encoder: PromptEncoder, mask_decoder: MaskDecoder,
                 pixel_mean: List[float] = [123.675, 116.28, 103.53],
                 pixel_std: List[float] = [58.395, 57.12, 57.375]) -> None:
        super().__init__()
        self.image_encoder = image_encoder
        self.prompt_encoder = prompt_encoder
        self.mask_decoder = mask_decoder
        self.register_buffer("pixel_mean", torch.Tensor(pixel_mean).view(-1, 1, 1), False)
        self.register_buffer("pixel_std", torch.Tensor(pixel_std).view(-1, 1, 1), False)

    @property
    def device(self) -> Any:
        return self.pixel_mean.device

    @torch.no_grad()
    def forward(self, batched_input: List[Dict[str, Any]], multimask_output: bool) -> List[Dict[str, torch.Tensor]]:
        """End-to-end prediction for a list of image records (sam.py:54-135).  Each record: 'image' [3,h,w] already
        resized to the model frame (longest side == img_size), 'original_size' (H, W), and any of 'point_coords' [n,N,2]
        + 'point_labels' [n,N], 'boxes' [n,4], 'mask_inputs' [n,1,256,256], 'text_embeds' [n,k,256] ([SEG] embeddings;
        not a key of upstream SAM).  Returns per image 'masks' bool [n,C,H,W], 'iou_predictions' [n,C],
        'low_res_logits' [n,C,256,256].  The images go through ONE batched encoder call as in the reference
        (:98-101); thresholding happens inside the postprocess kernel (no fp32 [n,C,H,W] compare pass)."""
        if len(batched_input) == 0:
            return []
        images = torch.stack([self.preprocess(rec["image"]) for rec in batched_input], dim=0)
        image_embeddings = self.image_encoder(images)
        image_pe = self.prompt_encoder.get_dense_pe()
        outputs = []
        for i, rec in enumerate(batched_input):
            points = (rec["point_coords"], rec["point_labels"]) if "point_coords" in rec else None
            sparse, dense = self.prompt_encoder(points=points, boxes=rec.get("boxes", None),
                                                masks=rec.get("mask_inputs", None),
                                                text_embeds=rec.get("text_embeds", None))
            low_res, iou = self.mask_decoder(image_embeddings=image_embeddings[i:i + 1], image_pe=image_pe,
                                             sparse_prompt_embeddings=sparse.to(image_embeddings.dtype),
                                             dense_prompt_embeddings=dense, multimask_output=multimask_output)
            _, binary = self.postprocess_masks(low_res, input_size=tuple(rec["image"].shape[-2:]),
                                               original_size=tuple(rec["original_size"]), return_binary=True)
            outputs.append({"masks": binary.view(torch.bool), "iou_predictions": iou, "low_res_logits": low_res})
        return outputs

    def postprocess_masks(self, masks: torch.Tensor, input_size: Tuple[int, ...], original_size: Tuple[int, ...],
                          return_binary: bool = False):
        """[n,C,L,L] low-res logits -> fp32 logits [n,C,H,W] (sam.py:137-172).  With return_binary=True also returns
        the uint8 mask `logits > mask_threshold` produced in the same pass.  When `masks` carries a graph (the mask
        loss of model/anyref.py:424-450 is taken on the post-processed logits) the result does too."""
        _runtime.require_cuda(masks, "Sam.postprocess_masks")
        if masks.dim() != 4 or masks.shape[2] != masks.shape[3]:
            raise ValueError(f"expected masks [n,C,L,L], got {tuple(masks.shape)}")
        if torch.is_grad_enabled() and masks.requires_grad:
            from .._train import PostprocessFn
            out = PostprocessFn.apply(masks, masks.shape[2], self.image_encoder.img_size, int(input_size[0]),
                                      int(input_size[1]), int(original_size[0]), int(original_size[1]))
            return (out, (out.detach() > self.mask_threshold).to(torch.uint8)) if return_binary else out
        return self._postprocess_masks(masks, input_size, original_size, return_binary)

    @_lib.device_scoped
    @torch.no_grad()
    def _postprocess_masks(self, masks, input_size, original_size, return_binary):
        m = masks.contiguous()
        n, ch, L, _ = m.shape
        h_in, w_in = int(input_size[0]), int(input_size[1])
        H, W = int(original_size[0]), int(original_size[1])
        S = self.image_encoder.img_size
        out = torch.empty((n, ch, H, W), device=m.device, dtype=torch.float32)
        binary = torch.empty((n, ch, H, W), device=m.device, dtype=torch.uint8) if return_binary else None
        if n * ch > 0:
            rc = _lib.load().sam_postprocess_masks(m.data_ptr(), _lib.fmt_of(m.dtype), n * ch, L, S, h_in, w_in, H, W,
                                                   out.data_ptr(), binary.data_ptr() if return_binary else None,
                                                   float(self.mask_threshold), _lib.stream_ptr(m.device))
            _lib.check(rc, "sam_postprocess_masks")
        return (out, binary) if return_binary else out

    @_lib.device_scoped
    @torch.no_grad()
    def postprocess_and_score(self, masks: torch.Tensor, input_size: Tuple[int, ...], original_size: Tuple[int, ...],
                              gt_masks: torch.Tensor, stats: Optional[torch.Tensor] = None,
                              return_binary: bool = False, return_packed: bool = False):
        """postprocess_masks + thresholding + the evaluation statistics of eval_referseg.py:186-211 in ONE pass:
        `gt_masks` uint8 [n,C,H,W] (0 / 1, 255 = ignore).  Adds this call's masks to `stats` (fp64 [7] =
        intersection bg/fg, union bg/fg, accumulated IoU bg/fg, count; created when None) and returns it -- that
        vector is what `anyref_b200.dp.all_reduce_stats` sums across ranks.  Nothing of full resolution is written
        unless return_binary=True (then the uint8 masks are returned as well) or return_packed=True (then the
        bit-packed masks, numpy.packbits order over the flattened [n,C,H,W] array, W % 8 == 0 -- the payload of
        `anyref_b200.dp.all_gather_packed`: H*W/8 bytes per mask)."""
        _runtime.require_cuda(masks, "Sam.postprocess_and_score")
        m = masks.contiguous()
        n, ch, L, _ = m.shape
        H, W = int(original_size[0]), int(original_size[1])
        if tuple(gt_masks.shape) != (n, ch, H, W) or gt_masks.dtype != torch.uint8:
            raise ValueError(f"gt_masks must be uint8 [{n},{ch},{H},{W}], got {gt_masks.dtype} {tuple(gt_masks.shape)}")
        gt = gt_masks.to(m.device).contiguous()
        if stats is None:
            stats = torch.zeros(7, dtype=torch.float64, device=m.device)
        if return_binary and return_packed:
            raise ValueError("return_binary and return_packed are mutually exclusive")
        if return_packed and W % 8 != 0:
            raise ValueError(f"return_packed needs W % 8 == 0, got W={W} (use return_binary + dp.pack_bits)")
        binary = torch.empty((n, ch, H, W), device=m.device, dtype=torch.uint8) if return_binary else None
        if return_packed:
            binary = torch.empty((n * ch * H * W // 8,), device=m.device, dtype=torch.uint8)
        if n * ch > 0 and return_packed:
            counts = torch.zeros((n * ch, 6), dtype=torch.int32, device=m.device)
            L_ = _lib.load()
            st = _lib.stream_ptr(m.device)
            rc = L_.sam_postprocess_masks_packed(m.data_ptr(), _lib.fmt_of(m.dtype), n * ch, L, self.image_encoder.img_size,
                                                 int(input_size[0]), int(input_size[1]), H, W, binary.data_ptr(),
                                                 float(self.mask_threshold), gt.data_ptr(), counts.data_ptr(), st)
            _lib.check(rc, "sam_postprocess_masks_packed")
            rc = L_.sam_iou_finalize(counts.data_ptr(), n * ch, stats.data_ptr(), st)
            _lib.check(rc, "sam_iou_finalize")
        elif n * ch > 0:
            counts = torch.zeros((n * ch, 6), dtype=torch.int32, device=m.device)
            L_ = _lib.load()
            st = _lib.stream_ptr(m.device)
            rc = L_.sam_postprocess_masks_iou(m.data_ptr(), _lib.fmt_of(m.dtype), n * ch, L, self.image_encoder.img_size,
                                              int(input_size[0]), int(input_size[1]), H, W, None,
                                              binary.data_ptr() if return_binary else None, float(self.mask_threshold),
                                              gt.data_ptr(), counts.data_ptr(), st)
            _lib.check(rc, "sam_postprocess_masks_iou")
            rc = L_.sam_iou_finalize(counts.data_ptr(), n * ch, stats.data_ptr(), st)
            _lib.check(rc, "sam_iou_finalize")
        return (stats, binary) if (return_binary or return_packed) else stats

    @_lib.device_scoped
    @torch.no_grad()
    def preprocess(self, x: torch.Tensor, out_dtype: Optional[torch.dtype] = None) -> torch.Tensor:
        """Normalise pixel values and pad to a square (sam.py:174-184; AnyRef's sam_preprocess,
        utils/refer_seg.py:560-593) as ONE kernel that also does the cast to the encoder's input dtype:
        [3,h,w] or [B,3,h,w], uint8 / fp16 / bf16 / fp32 in 0..255 -> [..., 3, S, S] in `out_dtype` (default: fp32 for
        float inputs as in the reference, which keeps the input dtype)."""
        _runtime.require_cuda(x, "Sam.preprocess")
        squeeze = x.dim() == 3
        xb = (x.unsqueeze(0) if squeeze else x).contiguous()
        if xb.dim() != 4 or xb.shape[1] != 3:
            raise ValueError(f"expected [3,h,w] or [B,3,h,w], got {tuple(x.shape)}")
        B, _, h, w = xb.shape
        S = self.image_encoder.img_size
        if xb.dtype == torch.uint8:
            in_fmt = 3
        else:
            in_fmt = _lib.fmt_of(xb.dtype)
        if out_dtype is None:
            out_dtype = torch.float32 if xb.dtype == torch.uint8 else xb.dtype
        out = torch.empty((B, 3, S, S), device=xb.device, dtype=out_dtype)
        import ctypes

        mean = (ctypes.c_float * 3)(*[float(v) for v in self.pixel_mean.flatten().tolist()])
        std = (ctypes.c_float * 3)(*[float(v) for v in self.pixel_std.flatten().tolist()])
        rc = _lib.load().sam_preprocess(xb.data_ptr(), in_fmt, out.data_ptr(), _lib.fmt_of(out_dtype), B, h, w, S, mean,
                                        std, _lib.stream_ptr(xb.device))
        _lib.check(rc, "sam_preprocess")
        return out[0] if squeeze else out
