"""ncu target: the three products of one image-side linear of the training path on the SGEMM (fp32 CUDA cores):
forward X.W^T, input gradient dY.W and weight gradient dY^T.X (split-K) at [32768, 256] x [256, 256] (8 prompts)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from anyref_b200 import _lib

lib = _lib.load()
M, N, K = 32768, 256, 256
x = torch.randn(M, K, device="cuda")
w = torch.randn(N, K, device="cuda") * 0.05
b = torch.randn(N, device="cuda")
y = torch.empty(M, N, device="cuda")
dy = torch.randn(M, N, device="cuda")
dx = torch.empty(M, K, device="cuda")
dw = torch.zeros(N, K, device="cuda")
db = torch.zeros(N, device="cuda")
nb = lib.sam_linear_f32_scratch_bytes(M, N, K)
scratch = torch.empty(nb + 256, dtype=torch.uint8, device="cuda")
for _ in range(2):
    assert lib.sam_linear_f32_forward(x.data_ptr(), w.data_ptr(), b.data_ptr(), y.data_ptr(), M, N, K, 0, scratch.data_ptr(), scratch.numel(), None) == 0
    assert lib.sam_linear_f32_backward(dy.data_ptr(), None, x.data_ptr(), w.data_ptr(), dx.data_ptr(), dw.data_ptr(), db.data_ptr(),
                                       M, N, K, scratch.data_ptr(), scratch.numel(), None) == 0
torch.cuda.synchronize()
ref = torch.nn.functional.linear(x, w, b)
print("forward max-abs vs torch:", float((y - ref).abs().max()), " dW rel:", float((dw - dy.t() @ x).norm() / (dy.t() @ x).norm()))


def timed(fn, iters=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3


if not os.environ.get("NCU"):
    gf = 2.0 * M * N * K / 1e9
    t = timed(lambda: lib.sam_linear_f32_forward(x.data_ptr(), w.data_ptr(), b.data_ptr(), y.data_ptr(), M, N, K, 0, scratch.data_ptr(), scratch.numel(), None))
    print(f"forward  {t:7.1f} us  {gf / t * 1e3:6.1f} TFLOP/s")
    t = timed(lambda: lib.sam_linear_f32_backward(dy.data_ptr(), None, x.data_ptr(), w.data_ptr(), dx.data_ptr(), None, None, M, N, K,
                                                  scratch.data_ptr(), scratch.numel(), None))
    print(f"dX       {t:7.1f} us  {gf / t * 1e3:6.1f} TFLOP/s")
    t = timed(lambda: lib.sam_linear_f32_backward(dy.data_ptr(), None, x.data_ptr(), w.data_ptr(), None, dw.data_ptr(), None, M, N, K,
                                                  scratch.data_ptr(), scratch.numel(), None))
    print(f"dW       {t:7.1f} us  {gf / t * 1e3:6.1f} TFLOP/s  (split-K + reduce)")

