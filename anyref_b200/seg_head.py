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

from typing import List, Sequence, Tuple

import torch
import torch.nn as nn

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
