"""Data-parallel evaluation of the grounding path: one process per GPU, replicated weights, images
sharded contiguously by rank, NO collective inside the forward.  After the forward each rank contributes

  * IoU statistics -- fp32 [intersection_bg, intersection_fg, union_bg, union_fg, acc_iou_bg, acc_iou_fg, count],
    summed with one all_reduce (reference analogue: AverageMeter.all_reduce, utils/utils.py:36-57, over
    intersectionAndUnionGPU outputs, utils/utils.py:79-91);
  * optionally the bit-packed binary masks, all_gather'ed to every rank.

Works with backend "nccl" (GPU) and "gloo" (CPU tests).
"""
from __future__ import annotations

import os
from typing import List, Optional, Tuple

import torch
import torch.distributed as dist


def init_from_env(backend: str | None = None) -> Tuple[int, int, int]:
    """Initialises torch.distributed from torchrun's environment.  Returns (rank, world_size, local_rank)."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local)
        kw = {"device_id": torch.device("cuda", local)} if backend == "nccl" else {}
        dist.init_process_group(backend=backend, rank=rank, world_size=world, **kw)
    return rank, world, local


def shard_range(num_items: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous shard [lo, hi) of rank `rank`: rank r gets items [r*N/P, (r+1)*N/P)."""
    return (rank * num_items) // world, ((rank + 1) * num_items) // world


def intersection_and_union(pred: torch.Tensor, target: torch.Tensor, k: int = 2):
    """intersectionAndUnionGPU (utils/utils.py:79-91) for label maps with values in [0, k)."""
    pred = pred.reshape(-1).to(torch.int64)
    target = target.reshape(-1).to(torch.int64)
    inter = pred[pred == target]
    area_i = torch.bincount(inter, minlength=k).float()
    area_p = torch.bincount(pred, minlength=k).float()
    area_t = torch.bincount(target, minlength=k).float()
    return area_i, area_p + area_t - area_i, area_t


def iou_stats(pred_masks: List[torch.Tensor], gt_masks: List[torch.Tensor]) -> torch.Tensor:
    """Per-rank 7-vector (see module docstring) accumulated over this rank's masks, as eval_referseg.py:186-211."""
    dev = pred_masks[0].device if pred_masks else torch.device("cpu")
    out = torch.zeros(7, dtype=torch.float32, device=dev)
    for p, g in zip(pred_masks, gt_masks):
        i, u, _ = intersection_and_union(p, g, 2)
        out[0:2] += i
        out[2:4] += u
        acc = i / (u + 1e-5)
        acc[u == 0] += 1.0  # no-object target (eval_referseg.py:199)
        out[4:6] += acc
        out[6] += 1
    return out


def all_reduce_stats(stats: torch.Tensor) -> torch.Tensor:
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(stats, op=dist.ReduceOp.SUM)
    return stats


def summarize(stats: torch.Tensor) -> dict:
    s = stats.detach().float().cpu()
    ciou = (s[1] / (s[3] + 1e-10)).item()
    giou = (s[5] / max(s[6].item(), 1.0)).item() if s[6] > 0 else 0.0
    return {"cIoU": ciou, "gIoU": giou, "count": int(s[6].item())}


def pack_bits(binary: torch.Tensor) -> torch.Tensor:
    """uint8 {0,1} mask [..., H, W] -> bit-packed uint8 [ceil(numel/8)] (MSB first, like numpy.packbits)."""
    flat = binary.reshape(-1).to(torch.uint8)
    pad = (-flat.numel()) % 8
    if pad:
        flat = torch.cat([flat, flat.new_zeros(pad)])
    w = torch.tensor([128, 64, 32, 16, 8, 4, 2, 1], dtype=torch.uint8, device=flat.device)
    return (flat.view(-1, 8) * w).sum(dim=1, dtype=torch.int32).to(torch.uint8)


def unpack_bits(packed: torch.Tensor, numel: int) -> torch.Tensor:
    w = torch.tensor([128, 64, 32, 16, 8, 4, 2, 1], dtype=torch.uint8, device=packed.device)
    return ((packed.view(-1, 1) & w) != 0).to(torch.uint8).reshape(-1)[:numel]


def _checksum(t: torch.Tensor) -> torch.Tensor:
    """Order-sensitive 64-bit checksum of a uint8 tensor, computed where the tensor lives: the bytes are read as int64
    words (zero-padded to a multiple of 8) and summed with weights 1 + (index mod 65521), wrap-around arithmetic."""
    n = t.numel()
    if n == 0:
        return torch.zeros(1, dtype=torch.int64, device=t.device)
    flat = t.reshape(-1)
    if n % 8 or flat.data_ptr() % 8 or not flat.is_contiguous():
        padded = flat.new_zeros((n + 7) // 8 * 8)
        padded[:n] = flat
        flat = padded
    words = flat.view(torch.int64)
    w = (torch.arange(words.numel(), device=t.device, dtype=torch.int64) % 65521) + 1
    return (words * w).sum().reshape(1)


def all_gather_packed(packed: torch.Tensor, verify: bool = True, mismatches: Optional[List[int]] = None) -> List[torch.Tensor]:
    """all_gather of per-rank bit-packed masks of possibly different lengths (padded to the max length).  With
    `verify`, every rank also contributes a checksum of what it sent and every rank checks the pieces it received
    against them: a transfer that does not reproduce the sender's bytes raises instead of returning silently different
    masks (the gathered masks are an evaluation artefact that is compared bit for bit between 1-rank and N-rank runs).
    When a list is passed as `mismatches`, the offending source ranks are appended to it instead of raising (a sweep
    that must reach its next collective on every rank)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return [packed]
    world = dist.get_world_size()
    meta = torch.cat([torch.tensor([packed.numel()], dtype=torch.int64, device=packed.device),
                      _checksum(packed) if verify else torch.zeros(1, dtype=torch.int64, device=packed.device)])
    metas = [torch.zeros_like(meta) for _ in range(world)]
    dist.all_gather(metas, meta)
    sizes = [int(m[0].item()) for m in metas]
    m = max(sizes)
    buf = packed.new_zeros(m)
    buf[:packed.numel()] = packed
    outs = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(outs, buf)
    pieces = [o[:n] for o, n in zip(outs, sizes)]
    if verify:
        for r, (piece, mt) in enumerate(zip(pieces, metas)):
            if int(_checksum(piece).item()) != int(mt[1].item()):
                if mismatches is None:
                    raise RuntimeError(f"all_gather_packed: the bytes received from rank {r} do not match what it sent")
                mismatches.append(r)
    return pieces
