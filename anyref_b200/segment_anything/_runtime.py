"""Shared host-side plumbing of the SAM modules: per-device scratch workspace and parameter-change tracking.

PyTorch owns all device memory: the C ABI never allocates, so the modules hand it a cached scratch
buffer that only ever grows.
"""
from __future__ import annotations

import torch

_WORKSPACES: dict = {}


def workspace(device: torch.device, nbytes: int, tag: str = "main"):
    """Returns (tensor_keepalive, 1024-byte-aligned device pointer) of at least `nbytes` bytes."""
    key = (device.type, device.index if device.index is not None else torch.cuda.current_device(), tag)
    buf = _WORKSPACES.get(key)
    need = nbytes + 1024
    if buf is None or buf.numel() < need:
        buf = None
        _WORKSPACES.pop(key, None)
        buf = torch.empty(need, dtype=torch.uint8, device=device)
        _WORKSPACES[key] = buf
    p = buf.data_ptr()
    return buf, (p + 1023) & ~1023


def release_workspaces() -> None:
    _WORKSPACES.clear()


def params_signature(module: torch.nn.Module):
    """Cheap fingerprint that changes whenever a parameter/buffer is replaced, moved, cast or written in place through
    autograd-tracked ops (load_state_dict, .to(), .half(), optimizer steps, `p.add_()` / `p.copy_()`).
    NOT detected: writes through `.data` (`p.data.copy_(w)`, `p.data += d` -- old-style checkpoint loaders and some
    peft merge code), which keep data_ptr and do not bump `_version`; after such a write call
    `module.invalidate_packed()` (ImageEncoderViT, MaskDecoder, SegProjection) so the packed weight blobs are rebuilt."""
    sig = []
    for t in list(module.parameters()) + list(module.buffers()):
        sig.append((t.data_ptr(), t._version, t.dtype))
    return hash(tuple(sig))


def require_cuda(t: torch.Tensor, what: str) -> None:
    if not t.is_cuda:
        raise RuntimeError(f"{what}: anyref_b200 runs only on CUDA (sm_100a) tensors -- there is no CPU fallback")
