"""The caller side of the grounding path: from the LLM's hidden states to masks.

Reference: `AnyRefModel.initialize_anyref_modules` builds `text_hidden_fcs = ModuleList([Sequential(Linear(h, h),
ReLU, Linear(h, 256), Dropout(0))])` (model/anyref.py:116-127) and `generate` / `evaluate` use it as

    hidden = hidden_states[seg_token_index[0], seg_token_index[1] + 255, :]      # :758  last layer, at the [SEG] tokens
    pred_embeddings = text_hidden_fcs[0](hidden)                                 # :770  [#seg, 256]
    image_embeddings = visual_model.image_encoder(sam_images)                    # :793
    for b: prompt_encoder(text_embeds=pred_embeddings[seg_batch == b][:, None]) -> mask_decoder -> postprocess_masks

`SegProjection` keeps the reference's module layout (children 0..3, state_dict keys `0.weight`, `0.bias`, `2.weight`,
`2.bias`) so `text_hidden_fcs.load_state_dict` is interchangeable, but its inference forward is two tcgen05 GEMMs with
the bias / ReLU fused in the epilogue.  `SegHead.__call__` is the tail of `generate` with the per-image loop replaced by
the batched `GroundingPath`.  The LLM itself (LLaVA / LLaMA, CLIP, ImageBind) stays stock PyTorch: it only has to hand
over its last-layer hidden states.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import ops
from .grounding import GroundingPath
from .segment_anything.modeling import Sam

IMAGE_TOKEN_EXPANSION = 255   # one <image> placeholder becomes 256 CLIP tokens (model/anyref.py:718, :758)


class SegProjection(nn.Sequential):
    """Linear(in, in) -> ReLU -> Linear(in, out) -> Dropout(0)  (model/anyref.py:118-123)."""

    def __init__(self, in_dim: int, out_dim: int) -> None:
        super().__init__(nn.Linear(in_dim, in_dim), nn.ReLU(inplace=True), nn.Linear(in_dim, out_dim), nn.Dropout(0.0))
        self._packed = None

    def invalidate_packed(self) -> None:
        """Drop the cached operand-format weights (see ImageEncoderViT.invalidate_packed)."""
        self._packed = None

    def _weights(self, dt: torch.dtype):
        ps = (self[0].weight, self[0].bias, self[2].weight, self[2].bias)
        key = tuple((p.data_ptr(), p._version, p.dtype) for p in ps) + (dt,)
        if self._packed is None or self._packed[0] != key:
            self._packed = (key, (ps[0].detach().to(dt).contiguous(), ps[1].detach().float().contiguous(),
                                  ps[2].detach().to(dt).contiguous(), ps[3].detach().float().contiguous()))
        return self._packed[1]

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        if not x.is_cuda:
            raise RuntimeError("SegProjection: anyref_b200 runs only on CUDA (sm_100a) tensors -- there is no CPU fallback")
        lead = x.shape[:-1]
        if torch.is_grad_enabled() and (x.requires_grad or any(p.requires_grad for p in self.parameters())):
            # training (model/anyref.py:116-124, :395-401): fp32 linears with a backward (csrc/decoder_train.cu)
            from .segment_anything._train import LinearF32Fn
            a = x.reshape(-1, x.shape[-1])
            if a.shape[0] == 0:
                return x.new_zeros((*lead, self[2].out_features))
            h = LinearF32Fn.apply(a, self[0].weight, self[0].bias, True)
            y = LinearF32Fn.apply(h, self[2].weight, self[2].bias, False)
            return y.to(x.dtype).reshape(*lead, self[2].out_features)
        dt = x.dtype if x.dtype in (torch.float16, torch.bfloat16) else torch.bfloat16
        w1, b1, w2, b2 = self._weights(dt)
        a = x.reshape(-1, x.shape[-1]).to(dt).contiguous()
        if a.shape[0] == 0:
            return x.new_zeros((*lead, w2.shape[0]))
        h = ops.gemm(a, w1, bias=b1, act="relu", out_dtype=dt)
        y = ops.gemm(h, w2, bias=b2, out_dtype=torch.float32)
        return y.to(x.dtype).reshape(*lead, w2.shape[0])


def dice_loss(inputs: torch.Tensor, targets: torch.Tensor, num_masks: float) -> torch.Tensor:
    """model/anyref.py:19-46 (the live branch: +1 smoothing, no scale): inputs are logits [n, H, W]."""
    inputs = inputs.sigmoid().flatten(1, 2)
    targets = targets.flatten(1, 2)
    numerator = 2 * (inputs * targets).sum(-1)
    denominator = inputs.sum(-1) + targets.sum(-1)
    return (1 - (numerator + 1) / (denominator + 1)).sum() / num_masks


def sigmoid_ce_loss(inputs: torch.Tensor, targets: torch.Tensor, num_masks: float) -> torch.Tensor:
    """model/anyref.py:50-67."""
    loss = F.binary_cross_entropy_with_logits(inputs, targets, reduction="none")
    return loss.flatten(1, 2).mean(1).sum() / (num_masks + 1e-8)


def build_text_hidden_fcs(in_dim: int = 4096, out_dim: int = 256) -> nn.ModuleList:
    """`self.text_hidden_fcs` of model/anyref.py:124 -- a one-element ModuleList, indexed `[0]` by the callers."""
    return nn.ModuleList([SegProjection(in_dim, out_dim)])


class SegHead:
    """Tail of `AnyRefForCausalLM.generate` (model/anyref.py:756-819): [SEG] hidden states -> masks."""

    def __init__(self, sam: Sam, text_hidden_fcs: nn.ModuleList, rephrase_weight: float = 0.0) -> None:
        self.sam = sam
        self.text_hidden_fcs = text_hidden_fcs
        self.rephrase_weight = rephrase_weight
        self.path = GroundingPath(sam)

    @torch.no_grad()
    def __call__(self, hidden_states: torch.Tensor, seg_token_index: Tuple[torch.Tensor, torch.Tensor],
                 sam_images: torch.Tensor, sam_resized_sizes: Sequence[Tuple[int, int]],
                 original_sizes: Sequence[Tuple[int, int]], rephrase_hidden_states=None,
                 multimask_output: bool = False) -> List[torch.Tensor]:
        """hidden_states [B, L + 255, H]: last LLM layer; seg_token_index = torch.where(output_ids[:, 1:] == seg_id)
        (batch index, position).  Returns, per image, fp32 mask logits [#seg_b, H_b, W_b] (`pred_mask.squeeze(1)`,
        :817); with no [SEG] token at all: a single zero mask per image as the reference does (:762-764)."""
        bs = sam_images.shape[0]
        bi, pos = seg_token_index
        if bi.numel() == 0:
            return [torch.zeros((1, *original_sizes[0]), device=sam_images.device, dtype=torch.float32)] * bs
        hidden = hidden_states[bi, pos + IMAGE_TOKEN_EXPANSION, :]
        if self.rephrase_weight > 0 and rephrase_hidden_states is not None:
            for i in range(min(bs, hidden.shape[0])):
                hidden[i] += rephrase_hidden_states[i] * self.rephrase_weight
        pred = self.text_hidden_fcs[0](hidden)                         # [#seg, 256]
        bi_host = bi.tolist()
        seg_list = [pred[[j for j, b_ in enumerate(bi_host) if b_ == b]].unsqueeze(1) for b in range(bs)]
        outs = self.path(sam_images, seg_list, sam_resized_sizes, original_sizes, multimask_output=multimask_output)
        return [o.squeeze(1) if not multimask_output else o for o in outs]

    def mask_loss(self, last_hidden_state: torch.Tensor, seg_token_index: Tuple[torch.Tensor, torch.Tensor],
                  sam_images: torch.Tensor, sam_resized_sizes: Sequence[Tuple[int, int]],
                  original_sizes: Sequence[Tuple[int, int]], gt_masks: Sequence[torch.Tensor],
                  bce_loss_weight: float = 2.0, dice_loss_weight: float = 0.5,
                  loc_embeddings: Optional[torch.Tensor] = None) -> Dict[str, torch.Tensor]:
        """The mask branch of `AnyRefForCausalLM.model_forward` (model/anyref.py:366-451) in training: frozen image
        encoder under no_grad (:366-367), `text_hidden_fcs` on the hidden states at the [SEG] tokens (:392-401, index
        already offset by the caller as `seg_token_idx_offset_one`), optional location embeddings (:403-404), prompt
        encoder, mask decoder, postprocess_masks (:406-430) and the BCE / dice losses with the reference's weights and
        normalisation (:432-450).  The per-image decoder calls of the reference are ONE batched call here (identical
        results, tests/test_gpu_train.py).  Everything that is trainable in AnyRef's recipe -- `text_hidden_fcs`, the
        mask decoder when `train_mask_decoder` -- and `last_hidden_state` itself receive gradients from the returned
        `mask_loss`; the caller adds the LM loss (:453-459)."""
        sam = self.sam
        bs = sam_images.shape[0]
        bi, pos = seg_token_index
        with torch.no_grad():
            image_embeddings = sam.image_encoder(sam_images)
        hidden = last_hidden_state[bi, pos, :]
        pred = self.text_hidden_fcs[0](hidden)                               # [#seg, 256]
        if loc_embeddings is not None:
            pred = pred + loc_embeddings
        order = torch.argsort(bi, stable=True)                               # prompts grouped by image, original order kept
        counts = torch.bincount(bi, minlength=bs).tolist()
        sparse, dense = sam.prompt_encoder(points=None, boxes=None, masks=None, text_embeds=pred[order].unsqueeze(1))
        index = bi[order].to(torch.int32)
        low, _ = sam.mask_decoder.forward_batched(image_embeddings, sam.prompt_encoder.get_dense_pe(),
                                                  sparse.to(pred.dtype), dense, index, False)
        ce = dice = 0.0
        num = 0
        pred_masks = []
        start = 0
        for b in range(bs):
            n_b = counts[b]
            if n_b == 0:                                                     # (the reference would divide 0 by 0 here)
                pred_masks.append(low.new_zeros((0, *original_sizes[b])))
                continue
            pm = sam.postprocess_masks(low[start:start + n_b], input_size=sam_resized_sizes[b],
                                       original_size=original_sizes[b]).squeeze(1)
            start += n_b
            pred_masks.append(pm)
            gt = gt_masks[b].to(pm)
            if pm.shape[-2:] != gt.shape[-2:]:                               # AVS targets (model/anyref.py:437-441)
                pm = F.interpolate(pm.unsqueeze(0), size=gt.shape[-2:], mode="bilinear", align_corners=False).squeeze(0)
            ce = ce + sigmoid_ce_loss(pm, gt, num_masks=gt.shape[0]) * gt.shape[0]
            dice = dice + dice_loss(pm, gt, num_masks=gt.shape[0]) * gt.shape[0]
            num += gt.shape[0]
        ce = bce_loss_weight * ce / (num + 1e-8)
        dice = dice_loss_weight * dice / (num + 1e-8)
        return {"ce_loss": ce, "dice_loss": dice, "mask_loss": ce + dice, "pred_masks": pred_masks}

