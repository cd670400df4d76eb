"""ImageEncoderViT with the reference's constructor, parameter tree and call signature
(reference: modeling/image_encoder.py:17-125), executed by libanyref_sam.so.

forward(x [B,3,1024,1024]) -> [B,256,64,64] in x.dtype, one `sam_encoder_forward` call per chunk of images.
Sub-modules (Block, Attention, PatchEmbed) exist to hold parameters under the reference's names
(blocks.N.attn.qkv.weight, ...); they have no stand-alone forward.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional, Tuple, Type

import torch
import torch.nn as nn

from .. import _pack, _runtime
from ... import _lib
from .common import LayerNorm2d, MLPBlock


def _no_forward(self, *a, **k):  # pragma: no cover
    raise RuntimeError(f"{type(self).__name__} runs inside ImageEncoderViT's fused CUDA forward; call the encoder")


class Attention(nn.Module):
    """Parameters of image_encoder.py:196-233."""

    def __init__(self, dim: int, num_heads: int = 8, qkv_bias: bool = True, use_rel_pos: bool = False,
                 rel_pos_zero_init: bool = True, input_size: Optional[Tuple[int, int]] = None) -> None:
        super().__init__()
        self.num_heads = num_heads
        self.scale = (dim // num_heads) ** -0.5
        self.qkv = nn.Linear(dim, dim * 3, bias=qkv_bias)
        self.proj = nn.Linear(dim, dim)
        self.use_rel_pos = use_rel_pos
        if use_rel_pos:
            assert input_size is not None, "Input size must be provided if using relative positional encoding."
            hd = dim // num_heads
            self.rel_pos_h = nn.Parameter(torch.zeros(2 * input_size[0] - 1, hd))
            self.rel_pos_w = nn.Parameter(torch.zeros(2 * input_size[1] - 1, hd))

    forward = _no_forward


class Block(nn.Module):
    """Parameters of image_encoder.py:128-175."""

    def __init__(self, dim: int, num_heads: int, mlp_ratio: float = 4.0, qkv_bias: bool = True,
                 norm_layer: Type[nn.Module] = nn.LayerNorm, act_layer: Type[nn.Module] = nn.GELU,
                 use_rel_pos: bool = False, rel_pos_zero_init: bool = True, window_size: int = 0,
                 input_size: Optional[Tuple[int, int]] = None) -> None:
        super().__init__()
        self.norm1 = norm_layer(dim)
        self.attn = Attention(dim, num_heads=num_heads, qkv_bias=qkv_bias, use_rel_pos=use_rel_pos,
                              rel_pos_zero_init=rel_pos_zero_init,
                              input_size=input_size if window_size == 0 else (window_size, window_size))
        self.norm2 = norm_layer(dim)
        self.mlp = MLPBlock(embedding_dim=dim, mlp_dim=int(dim * mlp_ratio), act=act_layer)
        self.window_size = window_size

    forward = _no_forward


class PatchEmbed(nn.Module):
    """Parameters of image_encoder.py:395-420."""

    def __init__(self, kernel_size=(16, 16), stride=(16, 16), padding=(0, 0), in_chans: int = 3,
                 embed_dim: int = 768) -> None:
        super().__init__()
        self.proj = nn.Conv2d(in_chans, embed_dim, kernel_size=kernel_size, stride=stride, padding=padding)

    forward = _no_forward


class ImageEncoderViT(nn.Module):
    #: images per sam_encoder_forward call (bounds the scratch workspace: ~85 MB per image for ViT-H)
    max_chunk = 32

    def __init__(self, img_size: int = 1024, patch_size: int = 16, in_chans: int = 3, embed_dim: int = 768,
                 depth: int = 12, num_heads: int = 12, mlp_ratio: float = 4.0, out_chans: int = 256,
                 qkv_bias: bool = True, norm_layer: Type[nn.Module] = nn.LayerNorm,
                 act_layer: Type[nn.Module] = nn.GELU, use_abs_pos: bool = True, use_rel_pos: bool = False,
                 rel_pos_zero_init: bool = True, window_size: int = 0,
                 global_attn_indexes: Tuple[int, ...] = ()) -> None:
        super().__init__()
        if in_chans != 3 or not qkv_bias or not use_abs_pos or not use_rel_pos:
            raise NotImplementedError("the B200 encoder implements SAM's configuration: RGB input, qkv bias, absolute "
                                      "and decomposed relative position embeddings (build_sam.py:65-82)")
        self.img_size = img_size
        self.patch_size = patch_size
        self.embed_dim = embed_dim
        self.num_heads = num_heads
        self.out_chans = out_chans
        self.window_size = window_size
        g = img_size // patch_size
        self.patch_embed = PatchEmbed(kernel_size=(patch_size, patch_size), stride=(patch_size, patch_size),
                                      in_chans=in_chans, embed_dim=embed_dim)
        self.pos_embed = nn.Parameter(torch.zeros(1, g, g, embed_dim))
        self.blocks = nn.ModuleList()
        for i in range(depth):
            self.blocks.append(Block(dim=embed_dim, num_heads=num_heads, mlp_ratio=mlp_ratio, qkv_bias=qkv_bias,
                                     norm_layer=norm_layer, act_layer=act_layer, use_rel_pos=use_rel_pos,
                                     rel_pos_zero_init=rel_pos_zero_init,
                                     window_size=window_size if i not in global_attn_indexes else 0,
                                     input_size=(g, g)))
        self.neck = nn.Sequential(
            nn.Conv2d(embed_dim, out_chans, kernel_size=1, bias=False),
            LayerNorm2d(out_chans),
            nn.Conv2d(out_chans, out_chans, kernel_size=3, padding=1, bias=False),
            LayerNorm2d(out_chans),
        )
        self._packed = None          # (signature, op_dtype, shape, w16, w32)
        self._operand_dtype = None   # explicit override (set_operand_dtype)
        self._ln_fold = None         # explicit override (set_ln_fold)

    # ------------------------------------------------------------------------------------------------ configuration
    def set_operand_dtype(self, dtype) -> None:
        """Tensor-core operand format (torch.bfloat16 or torch.float16); accumulation, residual stream, LayerNorm,
        softmax and GELU are always fp32.  Default: the parameters' dtype if it is 16-bit, else the input's dtype if
        it is 16-bit, else $ANYREF_SAM_OPERAND (bf16|fp16, default bf16)."""
        if dtype not in (None, torch.float16, torch.bfloat16):
            raise ValueError("operand dtype must be torch.float16 or torch.bfloat16")
        self._operand_dtype = dtype

    def set_ln_fold(self, on) -> None:
        """norm1 / norm2 folded into the GEMMs around them (5 kernels per block instead of 7; csrc/gemm2.cu).
        Default (None): on when embed_dim is a multiple of 256 (ViT-H / L / B) unless $ANYREF_SAM_LN_FOLD=0."""
        if on not in (None, True, False):
            raise ValueError("ln_fold must be None, True or False")
        if on and self.embed_dim % 256 != 0:
            raise ValueError("ln_fold needs embed_dim % 256 == 0")
        self._ln_fold = on

    def _resolve_ln_fold(self) -> bool:
        if self._ln_fold is not None:
            return bool(self._ln_fold)
        return self.embed_dim % 256 == 0 and os.environ.get("ANYREF_SAM_LN_FOLD", "1") != "0"

    def _resolve_operand_dtype(self, x: torch.Tensor):
        if self._operand_dtype is not None:
            return self._operand_dtype
        pd = self.pos_embed.dtype
        if pd in (torch.float16, torch.bfloat16):
            return pd
        if x.dtype in (torch.float16, torch.bfloat16):
            return x.dtype
        env = os.environ.get("ANYREF_SAM_OPERAND", "bf16").lower()
        return torch.float16 if env in ("fp16", "float16", "half") else torch.bfloat16

    def invalidate_packed(self) -> None:
        """Drop the packed weight blobs; the next forward re-packs from the current parameters.  Needed only after
        writes the parameter fingerprint cannot see (`p.data.copy_()`, `p.data += ...`; _runtime.params_signature)."""
        self._packed = None

    def _weights(self, op_dtype):
        fold = self._resolve_ln_fold()
        sig = (_runtime.params_signature(self), op_dtype, fold)
        if self._packed is None or self._packed[0] != sig:
            shape, w16, w32 = _pack.pack_encoder(self, op_dtype, fold)
            self._packed = (sig, shape, w16, w32)
        return self._packed[1:]

    # ------------------------------------------------------------------------------------------------------ forward
    @_lib.device_scoped
    @torch.no_grad()
    def forward(self, x: torch.Tensor, _tap: Optional[Tuple[int, torch.Tensor]] = None) -> torch.Tensor:
        _runtime.require_cuda(x, "ImageEncoderViT")
        if x.dim() != 4 or x.shape[1] != 3 or x.shape[2] != self.img_size or x.shape[3] != self.img_size:
            raise ValueError(f"expected images [B,3,{self.img_size},{self.img_size}], got {tuple(x.shape)}")
        lib = _lib.load()
        op_dtype = self._resolve_operand_dtype(x)
        shape, w16, w32 = self._weights(op_dtype)
        x = x.contiguous()
        B = x.shape[0]
        g = self.img_size // self.patch_size
        out = torch.empty((B, self.out_chans, g, g), device=x.device, dtype=x.dtype)
        stream = _lib.stream_ptr(x.device)
        for b0 in range(0, B, self.max_chunk):
            nb = min(self.max_chunk, B - b0)
            nbytes = lib.sam_encoder_workspace_bytes(C.byref(shape), nb)
            keep, wsp = _runtime.workspace(x.device, nbytes)
            sh = shape
            if _tap is not None:
                sh = _lib.SamEncoderShape.from_buffer_copy(shape)
                sh.tap_block, sh.tap_out = _tap[0], _tap[1][b0 * g * g:].data_ptr()
            rc = lib.sam_encoder_forward(C.byref(sh), w16.data_ptr(), w32.data_ptr(), x[b0:b0 + nb].data_ptr(),
                                         _lib.fmt_of(x.dtype), nb, out[b0:b0 + nb].data_ptr(), _lib.fmt_of(out.dtype),
                                         wsp, nbytes, stream)
            _lib.check(rc, "sam_encoder_forward")
        return out
