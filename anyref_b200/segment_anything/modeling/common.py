"""Parameter holders shared by the SAM modules (reference: modeling/common.py).

These classes keep the reference's parameter names and shapes so state_dicts are interchangeable; the arithmetic is
done by the CUDA kernels that the owning module's forward enqueues (LayerNorm2d: common.py:31-43 -> sam_layernorm /
sam_ln_nhwc_to_nchw; MLPBlock: common.py:13-26 -> two sam_gemm calls with GELU / residual epilogues).
"""
from __future__ import annotations

import torch
import torch.nn as nn


class MLPBlock(nn.Module):
    def __init__(self, embedding_dim: int, mlp_dim: int, act=nn.GELU) -> None:
        super().__init__()
        self.lin1 = nn.Linear(embedding_dim, mlp_dim)
        self.lin2 = nn.Linear(mlp_dim, embedding_dim)
        self.act = act()

    def forward(self, x):  # pragma: no cover
        raise RuntimeError("MLPBlock is executed inside its parent's fused CUDA forward; call the parent module")


class LayerNorm2d(nn.Module):
    def __init__(self, num_channels: int, eps: float = 1e-6) -> None:
        super().__init__()
        self.weight = nn.Parameter(torch.ones(num_channels))
        self.bias = nn.Parameter(torch.zeros(num_channels))
        self.eps = eps

    def forward(self, x):  # pragma: no cover
        raise RuntimeError("LayerNorm2d is executed inside its parent's fused CUDA forward; call the parent module")
