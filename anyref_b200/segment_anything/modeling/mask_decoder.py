"""MaskDecoder with the reference's constructor, parameter tree and call signature
(reference: modeling/mask_decoder.py:16-206), executed by `sam_decoder_forward` (csrc/decoder.cu).

forward(image_embeddings [1,C,g,g], image_pe [1,C,g,g], sparse_prompt_embeddings [n,k,C],
        dense_prompt_embeddings [n,C,g,g], multimask_output) -> (masks [n,1|3,4g,4g], iou [n,1|3])
`forward_batched` is the additional fast path of SURVEY 8(f)-1: prompts of MANY images in one call
(image_embeddings [B,C,g,g] + an image index per prompt) instead of the per-image Python loop of model/anyref.py:797.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Tuple, Type

import torch
import torch.nn as nn

from .. import _pack, _runtime
from ... import _lib
from .common import LayerNorm2d


class MLP(nn.Module):
    """Parameters of mask_decoder.py:184-206 (Linear-ReLU stack)."""

    def __init__(self, input_dim: int, hidden_dim: int, output_dim: int, num_layers: int,
                 sigmoid_output: bool = False) -> None:
        super().__init__()
        if sigmoid_output:
            raise NotImplementedError("sigmoid_output is never used by SAM's decoder")
        self.num_layers = num_layers
        h = [hidden_dim] * (num_layers - 1)
        self.layers = nn.ModuleList(nn.Linear(n, k) for n, k in zip([input_dim] + h, h + [output_dim]))
        self.sigmoid_output = sigmoid_output

    def forward(self, x):  # pragma: no cover
        raise RuntimeError("MLP runs inside MaskDecoder's fused CUDA forward; call the mask decoder")


class MaskDecoder(nn.Module):
    def __init__(self, *, transformer_dim: int, transformer: nn.Module, num_multimask_outputs: int = 3,
                 activation: Type[nn.Module] = nn.GELU, iou_head_depth: int = 3,
                 iou_head_hidden_dim: int = 256) -> None:
        super().__init__()
        if activation is not nn.GELU:
            raise NotImplementedError("the B200 decoder implements SAM's configuration: GELU upscaling")
        self.transformer_dim = transformer_dim
        self.transformer = transformer
        self.num_multimask_outputs = num_multimask_outputs
        self.iou_token = nn.Embedding(1, transformer_dim)
        self.num_mask_tokens = num_multimask_outputs + 1
        self.mask_tokens = nn.Embedding(self.num_mask_tokens, transformer_dim)
        self.output_upscaling = nn.Sequential(
            nn.ConvTranspose2d(transformer_dim, transformer_dim // 4, kernel_size=2, stride=2),
            LayerNorm2d(transformer_dim // 4),
            activation(),
            nn.ConvTranspose2d(transformer_dim // 4, transformer_dim // 8, kernel_size=2, stride=2),
            activation(),
        )
        self.output_hypernetworks_mlps = nn.ModuleList(
            [MLP(transformer_dim, transformer_dim, transformer_dim // 8, 3) for _ in range(self.num_mask_tokens)])
        self.iou_prediction_head = MLP(transformer_dim, iou_head_hidden_dim, self.num_mask_tokens, iou_head_depth)
        self._packed = None

    def invalidate_packed(self) -> None:
        """Drop the packed / derived weight blobs (see ImageEncoderViT.invalidate_packed)."""
        self._packed = None

    @_lib.device_scoped
    def _weights(self, grid: int, pe: torch.Tensor):
        """(shape, fp32 weight blob, derived buffer).  The derived buffer (split / transposed weight copies and the
        positional halves pe.W^T + b of the image-side projections) is rebuilt when a parameter or the dense positional
        encoding changes; PromptEncoder.get_dense_pe() returns a cached tensor, so this is once per model in practice."""
        sig = (_runtime.params_signature(self), grid, pe.data_ptr(), pe._version, pe.dtype)
        if self._packed is None or self._packed[0] != sig:
            shape, blob = _pack.pack_decoder(self, grid)
            lib = _lib.load()
            nbytes = lib.sam_decoder_derived_bytes(C.byref(shape))
            derived = torch.empty(nbytes + 256, dtype=torch.uint8, device=blob.device)
            dptr = (derived.data_ptr() + 255) & ~255
            rc = lib.sam_decoder_prepare(C.byref(shape), blob.data_ptr(), pe.data_ptr(), _lib.fmt_of(pe.dtype), dptr,
                                         _lib.stream_ptr(blob.device))
            _lib.check(rc, "sam_decoder_prepare")
            self._packed = (sig, shape, blob, (derived, dptr), pe)     # pe kept alive: its address is part of the key
        return self._packed[1:4]

    def forward(self, image_embeddings: torch.Tensor, image_pe: torch.Tensor,
                sparse_prompt_embeddings: torch.Tensor, dense_prompt_embeddings: torch.Tensor,
                multimask_output: bool) -> Tuple[torch.Tensor, torch.Tensor]:
        if image_embeddings.shape[0] != 1:
            raise ValueError("MaskDecoder.forward takes the embedding of ONE image (mask_decoder.py:146 repeats it per "
                             "prompt); use forward_batched for prompts of several images")
        sl = slice(1, None) if multimask_output else slice(0, 1)
        masks, iou = self.predict_masks(image_embeddings, image_pe, sparse_prompt_embeddings, dense_prompt_embeddings,
                                        _used=sl)
        return masks[:, sl, :, :], iou[:, sl]

    def forward_batched(self, image_embeddings: torch.Tensor, image_pe: torch.Tensor,
                        sparse_prompt_embeddings: torch.Tensor, dense_prompt_embeddings: torch.Tensor,
                        image_index: torch.Tensor, multimask_output: bool) -> Tuple[torch.Tensor, torch.Tensor]:
        """Prompt p is decoded against image_embeddings[image_index[p]] (int32 device tensor [n])."""
        sl = slice(1, None) if multimask_output else slice(0, 1)
        masks, iou = self.predict_masks(image_embeddings, image_pe, sparse_prompt_embeddings, dense_prompt_embeddings,
                                        image_index, _used=sl)
        return masks[:, sl, :, :], iou[:, sl]

    def predict_masks(self, image_embeddings: torch.Tensor, image_pe: torch.Tensor,
                      sparse_prompt_embeddings: torch.Tensor, dense_prompt_embeddings: torch.Tensor,
                      image_index: Optional[torch.Tensor] = None, _used: Optional[slice] = None
                      ) -> Tuple[torch.Tensor, torch.Tensor]:
        """All `num_mask_tokens` masks [n,4,4g,4g] and IoU predictions [n,4] (mask_decoder.py:116-179).  `_used`: the
        slice of mask tokens the caller keeps (forward's multimask selection) -- the training backward skips the
        hypernetwork MLPs of the others, whose cotangent is exactly zero."""
        if torch.is_grad_enabled() and (sparse_prompt_embeddings.requires_grad or
                                        (self.training and any(p.requires_grad for p in self.parameters()))):
            # model/anyref.py:108-113 fine-tunes the decoder (train(), requires_grad=True) with the encoders frozen:
            # the fp32 training path keeps its intermediates and has a backward (csrc/decoder_train.cu).  In eval()
            # with constant prompts the fused inference path runs and returns tensors without a graph
            lo, hi, _ = (_used or slice(None)).indices(self.num_mask_tokens)
            return self._predict_masks_train(image_embeddings, image_pe, sparse_prompt_embeddings,
                                             dense_prompt_embeddings, image_index, (lo, hi))
        with torch.no_grad():
            return self._predict_masks(image_embeddings, image_pe, sparse_prompt_embeddings, dense_prompt_embeddings,
                                       image_index)

    def _check_inputs(self, image_embeddings, image_pe, sparse_prompt_embeddings, dense_prompt_embeddings, image_index):
        """Shared argument checks -> (emb, pe, sparse, dense_vec, dense_full, image_index), all contiguous."""
        _runtime.require_cuda(image_embeddings, "MaskDecoder")
        emb = image_embeddings.contiguous()
        pe = image_pe.contiguous()
        _, Cc, g, g2 = emb.shape
        if g != g2 or Cc != self.transformer_dim or pe.shape[1:] != emb.shape[1:]:
            raise ValueError(f"bad embedding shapes {tuple(emb.shape)} / {tuple(pe.shape)}")
        sparse = sparse_prompt_embeddings.contiguous()
        n = sparse.shape[0]
        dense = dense_prompt_embeddings
        if dense.shape[0] != n:
            raise ValueError("dense_prompt_embeddings batch must equal the number of prompts")
        dense_vec = dense_full = None
        if dense.stride(0) == 0 and dense.stride(2) == 0 and dense.stride(3) == 0 and dense.stride(1) == 1:
            dense_vec = dense      # no_mask_embed broadcast (prompt_encoder.py:181-184): pass the [C] vector
        else:
            dense_full = dense.contiguous()
        if image_index is not None:
            if image_index.dtype != torch.int32 or not image_index.is_cuda or image_index.numel() != n:
                raise ValueError("image_index must be an int32 CUDA tensor with one entry per prompt")
            image_index = image_index.contiguous()
        elif emb.shape[0] != 1:
            raise ValueError("image_index is required when several image embeddings are given")
        return emb, pe, sparse, dense_vec, dense_full, image_index

    def _predict_masks_train(self, image_embeddings, image_pe, sparse_prompt_embeddings, dense_prompt_embeddings,
                             image_index, mask_range):
        """fp32 forward with a graph: gradients for this module's parameters and for sparse_prompt_embeddings.  Image
        embeddings, dense prompt embeddings and image_pe are constants here, as in the reference's fine-tuning where
        the image encoder and the prompt encoder are frozen (model/anyref.py:107-113); asking for their gradient
        raises instead of silently returning none.  Outputs are fp32 whatever the embedding dtype."""
        for name, t in (("image_embeddings", image_embeddings), ("dense_prompt_embeddings", dense_prompt_embeddings),
                        ("image_pe", image_pe)):
            if t.requires_grad:
                raise NotImplementedError(f"MaskDecoder training path: no gradient is produced for {name} (frozen in "
                                          "model/anyref.py:107-113); detach it")
        from .._train import DecoderTrainFn
        emb, pe, sparse, dense_vec, dense_full, image_index = self._check_inputs(
            image_embeddings, image_pe, sparse_prompt_embeddings, dense_prompt_embeddings, image_index)
        if sparse.shape[0] == 0:      # an image without a [SEG] token: empty outputs that still hang off the graph
            g = emb.shape[-1]
            zero = sparse.sum() * 0.0
            return (torch.zeros((0, self.num_mask_tokens, 4 * g, 4 * g), device=emb.device) + zero,
                    torch.zeros((0, self.num_mask_tokens), device=emb.device) + zero)
        params = tuple(self.parameters())
        return DecoderTrainFn.apply(self, mask_range, emb.detach(), pe.detach(), sparse,
                                    dense_vec.detach() if dense_vec is not None else None,
                                    dense_full.detach() if dense_full is not None else None, image_index, *params)

    @_lib.device_scoped
    def _predict_masks(self, image_embeddings, image_pe, sparse_prompt_embeddings, dense_prompt_embeddings, image_index):
        lib = _lib.load()
        emb, pe, sparse, dense_vec, dense_full, image_index = self._check_inputs(
            image_embeddings, image_pe, sparse_prompt_embeddings, dense_prompt_embeddings, image_index)
        g = emb.shape[-1]
        n, k = sparse.shape[0], sparse.shape[1]
        dense = dense_vec if dense_vec is not None else dense_full
        shape, blob, (_derived_keep, derived_ptr) = self._weights(g, pe)
        out_dtype = emb.dtype
        masks = torch.empty((n, self.num_mask_tokens, 4 * g, 4 * g), device=emb.device, dtype=out_dtype)
        iou = torch.empty((n, self.num_mask_tokens), device=emb.device, dtype=out_dtype)
        n_images = emb.shape[0]
        nbytes = lib.sam_decoder_workspace_bytes(C.byref(shape), n_images, n, k)
        keep, wsp = _runtime.workspace(emb.device, nbytes, "decoder")
        rc = lib.sam_decoder_forward(
            C.byref(shape), blob.data_ptr(), derived_ptr, emb.data_ptr(), _lib.fmt_of(emb.dtype), n_images,
            image_index.data_ptr() if image_index is not None else None,
            sparse.data_ptr() if k > 0 else None, _lib.fmt_of(sparse.dtype), n, k,
            dense_vec.data_ptr() if dense_vec is not None else None,
            dense_full.data_ptr() if dense_full is not None else None, _lib.fmt_of(dense.dtype),
            masks.data_ptr(), iou.data_ptr(), _lib.fmt_of(out_dtype), wsp, nbytes, _lib.stream_ptr(emb.device))
        _lib.check(rc, "sam_decoder_forward")
        return masks, iou
