"""Does replaying one whole step (encoder + decoder + postprocess, B = 16) as a CUDA graph beat stream launches?"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from anyref_b200.grounding import GroundingPath
from anyref_b200.segment_anything import build_sam_vit_h
from anyref_b200.synthetic import synthetic_images, synthetic_seg_embeddings, synthetic_state_dict

dev = torch.device("cuda", 0)
sam = build_sam_vit_h(None)
sam.load_state_dict(synthetic_state_dict("vit_h", seed=1234), strict=True)
sam = sam.to(dev)
sam.image_encoder.set_operand_dtype(torch.float16)
path = GroundingPath(sam)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
x = synthetic_images(B, seed=0).to(torch.float16).to(dev)
seg = synthetic_seg_embeddings(B, 1, seed=0).to(torch.float16).to(dev)
segs = [seg[b] for b in range(B)]
sizes = [(1024, 1024)] * B


def step():
    return path(x, segs, sizes, sizes)


def timeit(fn, n=30):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


print(f"stream launches: {timeit(step):.3f} ms/step")
g = torch.cuda.CUDAGraph()
s = torch.cuda.Stream()
s.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(s):
    step()
torch.cuda.current_stream().wait_stream(s)
with torch.cuda.graph(g):
    out = step()
print(f"graph replay   : {timeit(g.replay):.3f} ms/step")
print(f"stream launches: {timeit(step):.3f} ms/step")
print(f"graph replay   : {timeit(g.replay):.3f} ms/step")
ref = step()
g.replay()
torch.cuda.synchronize()
print("identical:", all(torch.equal(a, b) for a, b in zip(out, ref)))
