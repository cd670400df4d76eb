"""Debug timeline of the windowed-attention kernel: SAM_WIN3_TRACE=1 python tools/gpu_trace_win3.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from anyref_b200 import ops

B, heads, E = 16, 16, 1280
dt = torch.bfloat16
qkv = torch.randn(B * 4096, 3 * E, device="cuda").to(dt)
bias = torch.randn(3 * E, device="cuda").to(dt)
tab = ops.window_rel_table(torch.randn(27, 80, device="cuda") * 0.1, torch.randn(27, 80, device="cuda") * 0.1, dt)
for _ in range(2):
    ops.attn_window(qkv, bias, tab, B, heads)
torch.cuda.synchronize()
