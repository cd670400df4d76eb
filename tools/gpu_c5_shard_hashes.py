"""Bit-reproducibility of the data-parallel sweep: sha256 of the bit-packed masks of every 128-image chunk of the C5 sweep.
Single process: all 8 chunks on this GPU (twice: run-to-run determinism).  Under torchrun with 8 ranks: rank r computes chunk r on
its own GPU.  The chunk digests must agree between the two modes."""
import hashlib
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from anyref_b200 import eval_sweep
from anyref_b200.segment_anything import build_sam_from_config
from anyref_b200.synthetic import CONFIGS, synthetic_state_dict

rank = int(os.environ.get("RANK", "0"))
world = int(os.environ.get("WORLD_SIZE", "1"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
cfg = CONFIGS["vit_h"]
sam = build_sam_from_config(cfg)
sam.load_state_dict(synthetic_state_dict(cfg, seed=1234))
sam = sam.to(dev).eval()
dt = torch.float16
sam.image_encoder.set_operand_dtype(dt)
chunks = [rank] if world > 1 else list(range(8))
with torch.no_grad():
    for rep in range(1 if world > 1 else 2):
        for c in chunks:
            stats, packed = eval_sweep.run_shard(sam, 128 * c, 128 * (c + 1), 2, 16, dev, op_dtype=dt)
            h = hashlib.sha256(packed.cpu().numpy().tobytes()).hexdigest()[:16]
            print(f"world={world} rank={rank} gpu={torch.cuda.get_device_name(dev)} rep={rep} chunk={c} sha={h} "
                  f"inter_fg={int(stats[1].item())} union_fg={int(stats[3].item())}", flush=True)

if world > 1:
    # the gather itself: rank 0 hashes every gathered piece, to be compared with the owners' digests above
    import torch.distributed as dist

    from anyref_b200 import dp

    dist.init_process_group("nccl", device_id=dev)
    pieces = dp.all_gather_packed(packed)
    if rank == 0:
        total = hashlib.sha256()
        for r, p in enumerate(pieces):
            b = p.cpu().numpy().tobytes()
            total.update(b)
            print(f"world={world} gathered piece {r}: {len(b)} bytes sha={hashlib.sha256(b).hexdigest()[:16]}", flush=True)
        print(f"world={world} gathered total sha={total.hexdigest()}", flush=True)
    dist.barrier()
    dist.destroy_process_group()
else:
    total = hashlib.sha256()
    for c in range(8):
        stats, packed = eval_sweep.run_shard(sam, 128 * c, 128 * (c + 1), 2, 16, dev, op_dtype=dt)
        total.update(packed.cpu().numpy().tobytes())
    print(f"world=1 total sha={total.hexdigest()}", flush=True)
