"""B = 1 (the reference's evaluation shape) step time under different GEMM policies (experiment driver).
    [SAM_GEMM_V1=1 [SAM_GEMM_V1_BN=128]] python tools/gpu_small_batch.py [fold|nofold] [B]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from anyref_b200.grounding import GroundingPath
from anyref_b200.segment_anything import build_sam_vit_h
from anyref_b200.synthetic import synthetic_images, synthetic_seg_embeddings, synthetic_state_dict

fold = (sys.argv[1] if len(sys.argv) > 1 else "fold") == "fold"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 1
sam = build_sam_vit_h(None)
sam.load_state_dict(synthetic_state_dict("vit_h", seed=1234), strict=True)
sam = sam.cuda()
sam.image_encoder.set_operand_dtype(torch.float16)
sam.image_encoder.set_ln_fold(fold)
path = GroundingPath(sam)
x = synthetic_images(B, seed=0).to(torch.float16).cuda()
seg = synthetic_seg_embeddings(B, 1, seed=0).to(torch.float16).cuda()
sizes = [(1024, 1024)] * B
fn = lambda: path(x, [seg[b] for b in range(B)], sizes, sizes)
for _ in range(5):
    fn()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(30):
    fn()
e1.record()
torch.cuda.synchronize()
print(f"B={B} ln_fold={fold} V1={os.environ.get('SAM_GEMM_V1')} BN={os.environ.get('SAM_GEMM_V1_BN')}: {e0.elapsed_time(e1) / 30:.3f} ms per step")
