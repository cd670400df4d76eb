"""Debug timeline of the global-attention kernel (trace build): ANYREF_SAM_LIB=...gtrace.so python tools/gpu_trace_global.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from anyref_b200 import ops

B, heads, E = 16, 16, 1280
dt = torch.bfloat16
qkv = torch.randn(B * 4096, 3 * E, device="cuda").to(dt)
gh = ops.global_rel_table(torch.randn(127, 80, device="cuda") * 0.1, dt)
gw = ops.global_rel_table(torch.randn(127, 80, device="cuda") * 0.1, dt)
for _ in range(2):
    ops.attn_global(qkv, gh, gw, B, heads)
torch.cuda.synchronize()
