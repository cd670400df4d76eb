"""Data-parallel evaluation sweep of the grounding path (BASELINE.json configs[4]):

    N synthetic images x n [SEG] prompts, sharded contiguously over the ranks (one process per GPU, replicated weights),
    every rank runs encoder -> prompt encoder -> batched mask decoder -> postprocess + threshold + IoU counts on its
    shard with NO collective inside the forward; afterwards one all_reduce of the 7 IoU statistics (the analogue of
    AverageMeter.all_reduce over intersectionAndUnionGPU outputs, utils/utils.py:36-57, :79-91; the loop it replaces is
    eval_referseg.py:130-211) and, optionally, one all_gather of the bit-packed binary masks.

    python -m anyref_b200.eval_sweep --images 1024 --n-seg 2 --batch 16            # 1 GPU
    torchrun --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 -m anyref_b200.eval_sweep --images 1024 --n-seg 2

Inputs are a pure function of the GLOBAL image index (device generator seeded per image), so an N-rank run reproduces
the 1-rank run: the gathered masks (sha256 digest) and the integer intersection / union counts bit for bit, the
accumulated per-mask IoU up to the order of the fp64 additions.  Rank 0 prints one JSON line.
"""
from __future__ import annotations

import argparse
import hashlib
import json
import sys
from typing import List, Optional, Tuple

import torch

from . import dp


def _image_inputs(idx: int, n_seg: int, device, dtype, H: int, W: int, seed: int):
    """(image [3,1024,1024], [SEG] embeddings [n_seg,1,256], target masks uint8 [n_seg,1,H,W]) of global image `idx`."""
    g = torch.Generator(device=device)
    g.manual_seed(seed * 1_000_003 + idx)
    img = torch.randn((3, 1024, 1024), generator=g, device=device, dtype=torch.float32).to(dtype)
    seg = torch.randn((n_seg, 1, 256), generator=g, device=device, dtype=torch.float32).to(dtype)
    # blocky random targets (8x8 cells) with a few ignore pixels (255), as the datasets have
    cells = torch.rand((n_seg, 1, (H + 7) // 8, (W + 7) // 8), generator=g, device=device) > 0.5
    gt = cells.repeat_interleave(8, 2).repeat_interleave(8, 3)[:, :, :H, :W].to(torch.uint8)
    gt[:, :, : max(1, H // 64), :] = 255
    return img, seg, gt.contiguous()


@torch.no_grad()
def run_shard(sam, lo: int, hi: int, n_seg: int, batch: int, device, op_dtype=torch.bfloat16,
              input_size: Tuple[int, int] = (1024, 1024), original_size: Tuple[int, int] = (1024, 1024),
              multimask_output: bool = False, keep_masks: bool = True, seed: int = 0):
    """Images [lo, hi) through the path.  Returns (stats fp64 [7] on `device`, bit-packed masks uint8 or None)."""
    H, W = original_size
    stats = torch.zeros(7, dtype=torch.float64, device=device)
    packed: List[torch.Tensor] = []
    pe = sam.prompt_encoder.get_dense_pe()
    for b0 in range(lo, hi, batch):
        ids = range(b0, min(b0 + batch, hi))
        items = [_image_inputs(i, n_seg, device, op_dtype, H, W, seed) for i in ids]
        images = torch.stack([it[0] for it in items])
        text = torch.cat([it[1] for it in items])
        gt = torch.cat([it[2] for it in items])
        emb = sam.image_encoder(images)
        sparse, dense = sam.prompt_encoder(points=None, boxes=None, masks=None, text_embeds=text)
        sparse = sparse.to(text.dtype)                                           # model/anyref.py:806
        index = torch.arange(len(items), dtype=torch.int32, device=device).repeat_interleave(n_seg)
        low, _ = sam.mask_decoder.forward_batched(emb, pe, sparse, dense, index, multimask_output)
        if multimask_output:
            gt = gt.expand(-1, low.shape[1], -1, -1).contiguous()
        if keep_masks and W % 8 == 0:
            # thresholded masks leave the kernel bit-packed (H*W/8 bytes per mask), nothing else of full resolution
            stats, bits = sam.postprocess_and_score(low, input_size, original_size, gt, stats, return_packed=True)
            packed.append(bits)
        elif keep_masks:
            stats, binary = sam.postprocess_and_score(low, input_size, original_size, gt, stats, return_binary=True)
            packed.append(dp.pack_bits(binary))
        else:
            stats = sam.postprocess_and_score(low, input_size, original_size, gt, stats)
    return stats, (torch.cat(packed) if packed else torch.zeros(0, dtype=torch.uint8, device=device)) if keep_masks else None


def sweep(sam, num_images: int, n_seg: int, batch: int, rank: int, world: int, device, gather_masks: bool = True,
          **kw) -> dict:
    """The whole C5 sweep on this rank + the two collectives.  Device time = max over ranks."""
    import torch.distributed as dist

    lo, hi = dp.shard_range(num_images, rank, world)
    if world > 1:
        # the first collective of each kind sets up NCCL's channels (tens of ms, once per process): keep that one-time
        # cost out of the timed sweep, like the kernels' warm-up batch (round 1 charged it to the 8-GPU sweep: 11 %)
        dp.all_reduce_stats(torch.zeros(7, dtype=torch.float64, device=device))
        if gather_masks:
            dp.all_gather_packed(torch.zeros(1 << 20, dtype=torch.uint8, device=device))
        dist.barrier()
    torch.cuda.synchronize(device)
    e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
    e0.record()
    stats, packed = run_shard(sam, lo, hi, n_seg, batch, device, keep_masks=gather_masks, **kw)
    e1.record()
    stats = dp.all_reduce_stats(stats)
    bad_ranks: List[int] = []
    gathered = dp.all_gather_packed(packed, mismatches=bad_ranks) if gather_masks else None
    e2.record()
    torch.cuda.synchronize(device)
    t = torch.tensor([e0.elapsed_time(e1), e0.elapsed_time(e2)], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_forward, ms_total = t.tolist()
    out = dp.summarize(stats)
    out.update({"stats": [float(v) for v in stats.tolist()], "images": num_images, "masks": num_images * n_seg,
                "ms_forward": ms_forward, "ms_total": ms_total, "images_per_s": num_images / (ms_total * 1e-3),
                "masks_per_s": num_images * n_seg / (ms_total * 1e-3), "n_gpus": world})
    if gather_masks and rank == 0:
        h = hashlib.sha256()
        nbytes = 0
        per_rank = []
        for p in gathered:
            b = p.cpu().numpy().tobytes()
            h.update(b)
            per_rank.append(hashlib.sha256(b).hexdigest()[:16])
            nbytes += len(b)
        out["mask_bytes"] = nbytes
        out["mask_sha256"] = h.hexdigest()
        out["mask_sha256_per_rank"] = per_rank      # localises a mismatch between an N-rank and a 1-rank run
        out["gather_verified"] = not bad_ranks     # every received piece reproduces its sender's checksum
        if bad_ranks:
            out["gather_mismatch_ranks"] = bad_ranks
    return out


def main(argv: Optional[List[str]] = None) -> int:
    ap = argparse.ArgumentParser()
    ap.add_argument("--images", type=int, default=1024)
    ap.add_argument("--n-seg", type=int, default=2)
    ap.add_argument("--batch", type=int, default=16)
    ap.add_argument("--dtype", default="bf16", choices=["bf16", "fp16"])
    ap.add_argument("--model", default="vit_h")
    ap.add_argument("--no-gather-masks", action="store_true")
    args = ap.parse_args(argv)
    if not torch.cuda.is_available():
        raise SystemExit("eval_sweep needs CUDA devices (there is no CPU fallback for the product path)")
    from .segment_anything import build_sam_from_config
    from .synthetic import CONFIGS, synthetic_state_dict

    rank, world, local = dp.init_from_env("nccl")
    if world == 1:
        torch.cuda.set_device(0)
    device = torch.device("cuda", torch.cuda.current_device())
    op_dtype = torch.bfloat16 if args.dtype == "bf16" else torch.float16
    sam = build_sam_from_config(CONFIGS[args.model])
    sam.load_state_dict(synthetic_state_dict(args.model, seed=1234), strict=True)
    sam = sam.to(device)
    sam.image_encoder.set_operand_dtype(op_dtype)
    run_shard(sam, 0, min(args.batch, args.images), args.n_seg, args.batch, device, op_dtype=op_dtype)   # warm-up
    res = sweep(sam, args.images, args.n_seg, args.batch, rank, world, device, gather_masks=not args.no_gather_masks,
                op_dtype=op_dtype)
    if rank == 0:
        res["config"] = {"workload": f"C5: {args.images} synthetic 1024x1024 images x {args.n_seg} [SEG], {args.model} "
                                     f"{args.dtype}, batch {args.batch} per GPU, contiguous shards over {world} GPU(s); "
                                     "all_reduce of 7 IoU statistics" + ("" if args.no_gather_masks else " + all_gather of "
                                                                         "bit-packed masks")}
        print(json.dumps(res), flush=True)
    if world > 1:
        import torch.distributed as dist

        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
