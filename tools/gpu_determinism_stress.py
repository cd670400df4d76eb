"""Run-to-run determinism of the whole path: 40 passes over two 64-image shards of the C5 sweep (ViT-H, fp16 operands) on one GPU,
sha256 of the bit-packed masks of every pass -- one digest per shard or the path has a race."""
import hashlib, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from anyref_b200 import eval_sweep
from anyref_b200.segment_anything import build_sam_from_config
from anyref_b200.synthetic import CONFIGS, synthetic_state_dict
dev = torch.device("cuda", 0)
cfg = CONFIGS["vit_h"]
sam = build_sam_from_config(cfg); sam.load_state_dict(synthetic_state_dict(cfg, seed=1234)); sam = sam.to(dev).eval()
sam.image_encoder.set_operand_dtype(torch.float16)
seen = {}
with torch.no_grad():
    for rep in range(40):
        c = rep % 2
        stats, packed = eval_sweep.run_shard(sam, 128 * c, 128 * c + 64, 2, 16, dev, op_dtype=torch.float16)
        h = hashlib.sha256(packed.cpu().numpy().tobytes()).hexdigest()[:16]
        seen.setdefault(c, set()).add(h)
print({c: sorted(v) for c, v in seen.items()})
print("deterministic" if all(len(v) == 1 for v in seen.values()) else "NONDETERMINISTIC")
