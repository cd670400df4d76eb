"""Repeats the windowed-attention parity case that failed once (B=2, 16 heads, bf16) and prints the error of every run."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from anyref_b200 import ops
from tests.test_gpu_kernels import _window_reference, rel_fro

DEV = "cuda"
dt = torch.bfloat16
B, heads = 2, 16
torch.manual_seed(3)
E = heads * 80
qkv = torch.randn(B * 4096, 3 * E, device=DEV).to(dt)
bias = (torch.randn(3 * E, device=DEV) * 0.5).to(dt)
rel_h = (torch.randn(27, 80, device=DEV) * 0.2).to(dt)
rel_w = (torch.randn(27, 80, device=DEV) * 0.2).to(dt)
tab = ops.window_rel_table(rel_h, rel_w, dt)
want = _window_reference(qkv, bias, rel_h, rel_w, B, heads)
first = None
bad = 0
n = int(sys.argv[1]) if len(sys.argv) > 1 else 200
junk = torch.empty(256 << 20, dtype=torch.uint8, device=DEV)
for it in range(n):
    if it % 3 == 0:
        junk.random_(0, 255)          # perturb L2 / timing
    got = ops.attn_window(qkv, bias, tab, B, heads)
    if first is None:
        first = got.clone()
    e = rel_fro(got, want)
    same = torch.equal(got, first)
    if e >= 5e-3 or not same or not bool(torch.isfinite(got.float()).all()):
        bad += 1
        d = (got.float() - first.float()).abs()
        rows = torch.nonzero(d.amax(1) > 0).flatten()
        print(f"run {it}: rel_fro {e:.3e} identical_to_first {same} differing rows {rows.numel()} "
              f"first rows {rows[:8].tolist()} max diff {d.max().item():.3e} cols {torch.nonzero(d.amax(0) > 0).flatten()[:6].tolist()}")
print(f"windowed: {n} runs, {bad} bad, first rel_fro {rel_fro(first, want):.3e}")

# global attention: bit-reproducibility under the same perturbation (B = 2 images, 16 heads)
gh = ops.global_rel_table((torch.randn(127, 80, device=DEV) * 0.2).to(dt), dt)
gw = ops.global_rel_table((torch.randn(127, 80, device=DEV) * 0.2).to(dt), dt)
first = None
bad = 0
for it in range(n // 2):
    if it % 3 == 0:
        junk.random_(0, 255)
    got = ops.attn_global(qkv, gh, gw, B, heads)
    if first is None:
        first = got.clone()
    if not torch.equal(got, first) or not bool(torch.isfinite(got.float()).all()):
        bad += 1
        d = (got.float() - first.float()).abs()
        rows = torch.nonzero(d.amax(1) > 0).flatten()
        print(f"global run {it}: differing rows {rows.numel()} first {rows[:8].tolist()} max diff {d.max().item():.3e}")
print(f"global: {n // 2} runs, {bad} bad")
