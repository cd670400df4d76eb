"""Diagnostic run on a B200: fused attention + glue kernels vs the oracle's formulas evaluated in fp32 on the GPU.
Prints, never asserts."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F

from anyref_b200 import ops
from oracle import sam_oracle as O


def ref_attention(qkv, rel_h, rel_w, B, heads, window):
    """oracle.encoder_attention without the qkv/proj linears, on already-projected (rounded) qkv [B*4096, 3E]."""
    E = qkv.shape[1] // 3
    x = qkv.float().view(B, 64, 64, 3 * E)
    if window:
        x, padded = O._partition(x, window)
    b, h, w, _ = x.shape
    q, k, v = x.reshape(b, h * w, 3, heads, -1).permute(2, 0, 3, 1, 4).reshape(3, b * heads, h * w, -1).unbind(0)
    scale = (E // heads) ** -0.5
    attn = (q * scale) @ k.transpose(-2, -1)
    rh = O._rel_table(h, h, rel_h.float())
    rw = O._rel_table(w, w, rel_w.float())
    rq = q.reshape(b * heads, h, w, -1)
    a = torch.einsum("bhwc,hkc->bhwk", rq, rh)
    c = torch.einsum("bhwc,wkc->bhwk", rq, rw)
    attn = (attn.view(b * heads, h, w, h, w) + a[:, :, :, :, None] + c[:, :, :, None, :]).view(b * heads, h * w, h * w)
    attn = attn.softmax(-1)
    out = (attn @ v).view(b, heads, h, w, -1).permute(0, 2, 3, 1, 4).reshape(b, h, w, -1)
    if window:
        out = O._unpartition(out, window, padded, (64, 64))
    return out.reshape(B * 4096, E)


def stats(name, got, ref):
    d = (got.float() - ref.float())
    print(f"{name}: max_abs={d.abs().max().item():.3e} rel_fro={(d.norm() / ref.float().norm()).item():.3e} "
          f"ref_max={ref.abs().max().item():.3f} finite={bool(torch.isfinite(got.float()).all())}")


def attn_suite():
    dev = "cuda"
    torch.manual_seed(0)
    for dt in (torch.float16, torch.bfloat16):
        for (B, heads) in ((1, 2), (2, 16)):
            E = heads * 80
            qkv = (torch.randn(B * 4096, 3 * E, device=dev) * 1.0).to(dt)
            bias = (torch.randn(3 * E, device=dev) * 0.5).to(dt)
            # windowed: padded tokens carry the qkv bias -> emulate by the reference on a padded qkv built explicitly
            rel_h = (torch.randn(27, 80, device=dev) * 0.2).to(dt)
            rel_w = (torch.randn(27, 80, device=dev) * 0.2).to(dt)
            tab = ops.window_rel_table(rel_h, rel_w, dt)
            out = ops.attn_window(qkv, bias, tab, B, heads)
            torch.cuda.synchronize()
            # reference: partition with padding value == bias (pad zeros then overwrite)
            x = qkv.float().view(B, 64, 64, 3 * E)
            xp = bias.float().view(1, 1, 1, -1).expand(B, 70, 70, 3 * E).clone()
            xp[:, :64, :64] = x
            win = xp.view(B, 5, 14, 5, 14, 3 * E).permute(0, 1, 3, 2, 4, 5).reshape(-1, 14, 14, 3 * E)
            b = win.shape[0]
            q, k, v = win.reshape(b, 196, 3, heads, 80).permute(2, 0, 3, 1, 4).reshape(3, b * heads, 196, 80).unbind(0)
            attn = (q * 80 ** -0.5) @ k.transpose(-2, -1)
            rh = O._rel_table(14, 14, rel_h.float())
            rw = O._rel_table(14, 14, rel_w.float())
            rq = q.reshape(b * heads, 14, 14, 80)
            a = torch.einsum("bhwc,hkc->bhwk", rq, rh)
            c = torch.einsum("bhwc,wkc->bhwk", rq, rw)
            attn = (attn.view(b * heads, 14, 14, 14, 14) + a[..., None] + c[:, :, :, None, :]).view(b * heads, 196, 196)
            o = (attn.softmax(-1) @ v).view(b, heads, 14, 14, 80).permute(0, 2, 3, 1, 4).reshape(b, 14, 14, E)
            ref = O._unpartition(o, 14, (70, 70), (64, 64)).reshape(B * 4096, E)
            stats(f"attn_window dt={str(dt)[6:]} B={B} heads={heads}", out, ref)
            # global
            gh = (torch.randn(127, 80, device=dev) * 0.2).to(dt)
            gw = (torch.randn(127, 80, device=dev) * 0.2).to(dt)
            out = ops.attn_global(qkv, ops.global_rel_table(gh, dt), ops.global_rel_table(gw, dt), B, heads)
            torch.cuda.synchronize()
            if B * heads <= 4:
                ref = ref_attention(qkv, gh, gw, B, heads, 0)
            else:  # chunk over images to bound the [heads,4096,4096] fp32 matrix
                ref = torch.cat([ref_attention(qkv[i * 4096:(i + 1) * 4096], gh, gw, 1, heads, 0) for i in range(B)])
            stats(f"attn_global dt={str(dt)[6:]} B={B} heads={heads}", out, ref)


def glue_suite():
    dev = "cuda"
    torch.manual_seed(1)
    x = torch.randn(4096 * 2, 1280, device=dev) * 2 + 0.3
    g = torch.rand(1280, device=dev) + 0.5
    b = torch.randn(1280, device=dev) * 0.1
    for dt in (torch.float32, torch.bfloat16, torch.float16):
        out = ops.layernorm(x, g, b, 1e-6, dt)
        stats(f"layernorm 1280 -> {str(dt)[6:]}", out, F.layer_norm(x, (1280,), g, b, 1e-6))
    r = torch.randn_like(x)
    stats("layernorm(x+res)", ops.layernorm(x, g, b, 1e-5, torch.float32, residual=r), F.layer_norm(x + r, (1280,), g, b, 1e-5))
    x2 = torch.randn(300, 256, device=dev)
    stats("layernorm 256", ops.layernorm(x2, g[:256].contiguous(), b[:256].contiguous(), 1e-5, torch.float32),
          F.layer_norm(x2, (256,), g[:256], b[:256], 1e-5))
    stats("cast", ops.layernorm(x, None, None, 0.0, torch.bfloat16, normalize=False), x.to(torch.bfloat16))
    img = torch.randn(2, 3, 1024, 1024, device=dev)
    for dt_in in (torch.float32, torch.bfloat16):
        pm = ops.patch_im2col(img.to(dt_in), 16, torch.bfloat16)
        ref = F.unfold(img.to(dt_in).float(), 16, stride=16).transpose(1, 2).reshape(-1, 768).to(torch.bfloat16)
        stats(f"patch_im2col in={str(dt_in)[6:]}", pm, ref)
    y = torch.randn(2, 64, 64, 256, device=dev).to(torch.bfloat16)
    cols = ops.im2col3x3(y, 2, 64)
    yp = F.pad(y.float(), (0, 0, 1, 1, 1, 1))
    ref = torch.stack([yp[:, ky:ky + 64, kx:kx + 64] for ky in range(3) for kx in range(3)], dim=3).reshape(-1, 9 * 256)
    stats("im2col3x3", cols, ref)
    z = torch.randn(2 * 4096, 256, device=dev) * 3
    g2, b2 = g[:256].contiguous(), b[:256].contiguous()
    out = ops.ln_nhwc_to_nchw(z, g2, b2, 1e-6, 2, 64, torch.float32)
    ref = F.layer_norm(z, (256,), g2, b2, 1e-6).view(2, 64, 64, 256).permute(0, 3, 1, 2)
    stats("ln_nhwc_to_nchw", out, ref)


def attn_bench():
    dev = "cuda"
    B, heads, E = 16, 16, 1280
    dt = torch.bfloat16
    qkv = torch.randn(B * 4096, 3 * E, device=dev).to(dt)
    bias = torch.randn(3 * E, device=dev).to(dt)
    tab = ops.window_rel_table(torch.randn(27, 80, device=dev) * 0.1, torch.randn(27, 80, device=dev) * 0.1, dt)
    gh = ops.global_rel_table(torch.randn(127, 80, device=dev) * 0.1, dt)
    gw = ops.global_rel_table(torch.randn(127, 80, device=dev) * 0.1, dt)
    for fn, name, flops in ((lambda: ops.attn_window(qkv, bias, tab, B, heads), "attn_window", B * 5.27e9),
                            (lambda: ops.attn_global(qkv, gh, gw, B, heads), "attn_global", B * 87.2e9)):
        for _ in range(20):
            fn()
        torch.cuda.synchronize()
        times = []
        for rep in range(5):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(40):
                fn()
            e1.record()
            torch.cuda.synchronize()
            times.append(e0.elapsed_time(e1) / 40)
        ms = sorted(times)[len(times) // 2]
        print(f"{name} B=16: median {ms:.3f} ms (min {min(times):.3f} max {max(times):.3f})  {flops / ms / 1e9:.1f} TFLOP/s")


if __name__ == "__main__":
    print(torch.cuda.get_device_name(0))
    t = time.time()
    which = sys.argv[1:] or ["glue", "attn", "bench"]
    if "glue" in which:
        glue_suite()
    if "attn" in which:
        attn_suite()
    if "bench" in which:
        attn_bench()
    print(f"done in {time.time() - t:.1f}s")
