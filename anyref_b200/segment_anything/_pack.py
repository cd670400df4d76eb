"""Packs module parameters into the weight blobs the C ABI consumes (layouts: csrc/encoder.cpp, csrc/decoder.cu).

Runs once per parameter change (see _runtime.params_signature); pure data movement (casts, permutes, copies).
"""
from __future__ import annotations

import ctypes as C

import torch

from .. import _lib


def _align8(n: int) -> int:
    return (n + 7) & ~7


class _Blob:
    def __init__(self, n: int, dtype, device):
        self.t = torch.zeros(n, dtype=dtype, device=device)
        self.off = 0

    def put(self, src: torch.Tensor, n_expected: int | None = None):
        flat = src.detach().reshape(-1)
        n = flat.numel()
        if n_expected is not None and n != n_expected:
            raise RuntimeError(f"pack: segment has {n} elements, layout expects {n_expected}")
        self.t[self.off:self.off + n].copy_(flat)
        self.off += _align8(n)

    def put_exact(self, src: torch.Tensor):  # unpadded (decoder blob)
        flat = src.detach().reshape(-1)
        self.t[self.off:self.off + flat.numel()].copy_(flat)
        self.off += flat.numel()


def encoder_shape(enc, op_dtype, ln_fold: bool = False) -> _lib.SamEncoderShape:
    gmask = 0
    for i, blk in enumerate(enc.blocks):
        if blk.window_size == 0:
            gmask |= 1 << i
    return _lib.SamEncoderShape(embed_dim=enc.embed_dim, depth=len(enc.blocks), heads=enc.num_heads,
                                mlp_dim=enc.blocks[0].mlp.lin1.out_features, img=enc.img_size, patch=enc.patch_size,
                                window=enc.window_size, out_chans=enc.out_chans, fmt=_lib.fmt_of(op_dtype),
                                global_mask=gmask, tap_block=-1, tap_out=None, ln_fold=1 if ln_fold else 0)


def _fold_layernorm(norm, lin, op_dtype):
    """LayerNorm folded into the Linear that follows it (csrc/gemm2.cu, "LayerNorm folding"):
        LN(x) . W^T + b  =  rstd * (x . Wg^T - mean * colsum) + bias_fold
    -> (Wg = gamma o W in the operand format, colsum = row sums of the ROUNDED Wg, bias_fold = beta . W^T + b)."""
    w = lin.weight.detach().double()
    wg = (w * norm.weight.detach().double()[None, :]).float().to(op_dtype)
    colsum = wg.double().sum(dim=1).float()
    bias_fold = (w @ norm.bias.detach().double() + lin.bias.detach().double()).float()
    return wg, colsum, bias_fold


def pack_encoder(enc, op_dtype, ln_fold: bool = False):
    """-> (shape, w16 blob, w32 blob).  ln_fold selects the folded-LayerNorm blob layout (csrc/encoder.cpp)."""
    lib = _lib.load()
    shape = encoder_shape(enc, op_dtype, ln_fold)
    dev = enc.pos_embed.device
    n16 = lib.sam_encoder_w16_elems(C.byref(shape))
    n32 = lib.sam_encoder_w32_elems(C.byref(shape))
    b16 = _Blob(n16, op_dtype, dev)
    b32 = _Blob(n32, torch.float32, dev)
    E, hd = enc.embed_dim, enc.embed_dim // enc.num_heads
    g = enc.img_size // enc.patch_size
    b16.put(enc.patch_embed.proj.weight.reshape(E, -1).to(op_dtype))
    b32.put(enc.pos_embed.float(), g * g * E)
    b32.put(enc.patch_embed.proj.bias.float())
    for blk in enc.blocks:
        a = blk.attn
        if ln_fold:
            qkv_w, qkv_cs, qkv_fb = _fold_layernorm(blk.norm1, a.qkv, op_dtype)
            lin1_w, lin1_cs, lin1_fb = _fold_layernorm(blk.norm2, blk.mlp.lin1, op_dtype)
        else:
            qkv_w, qkv_fb = a.qkv.weight.to(op_dtype), a.qkv.bias.float()
            lin1_w, lin1_fb = blk.mlp.lin1.weight.to(op_dtype), blk.mlp.lin1.bias.float()
        b16.put(qkv_w)
        b16.put(a.proj.weight.to(op_dtype))
        b16.put(lin1_w)
        b16.put(blk.mlp.lin2.weight.to(op_dtype))
        b16.put(a.qkv.bias.to(op_dtype))
        s = g if blk.window_size == 0 else blk.window_size
        if a.rel_pos_h.shape != (2 * s - 1, hd) or a.rel_pos_w.shape != (2 * s - 1, hd):
            raise RuntimeError("rel_pos tables must have 2*size-1 rows (the interpolating branch of get_rel_pos, "
                               "image_encoder.py:336-343, is never taken by SAM checkpoints and is not implemented)")
        rel = torch.zeros(2 * 128 * hd, dtype=op_dtype, device=dev)
        if blk.window_size == 0:
            if s != 64:
                raise RuntimeError("global attention kernel expects a 64x64 token grid")
            rel[:127 * hd] = a.rel_pos_h.detach().flip(0).to(op_dtype).reshape(-1)
            rel[128 * hd:128 * hd + 127 * hd] = a.rel_pos_w.detach().flip(0).to(op_dtype).reshape(-1)
        else:
            if s != 14:
                raise RuntimeError("windowed attention kernel expects 14x14 windows")
            rel[:27 * hd] = a.rel_pos_h.detach().to(op_dtype).reshape(-1)
            rel[32 * hd:59 * hd] = a.rel_pos_w.detach().to(op_dtype).reshape(-1)
        b16.put(rel)
        b32.put(blk.norm1.weight.float()); b32.put(blk.norm1.bias.float())
        b32.put(qkv_fb)
        b32.put(a.proj.bias.float())
        b32.put(blk.norm2.weight.float()); b32.put(blk.norm2.bias.float())
        b32.put(lin1_fb)
        b32.put(blk.mlp.lin2.bias.float())
        if ln_fold:
            b32.put(qkv_cs)
            b32.put(lin1_cs)
    Cc = enc.out_chans
    b16.put(enc.neck[0].weight.reshape(Cc, E).to(op_dtype))
    b16.put(enc.neck[2].weight.permute(0, 2, 3, 1).reshape(Cc, 9 * Cc).to(op_dtype))
    b32.put(enc.neck[1].weight.float()); b32.put(enc.neck[1].bias.float())
    b32.put(enc.neck[3].weight.float()); b32.put(enc.neck[3].bias.float())
    if b16.off != n16 or b32.off != n32:
        raise RuntimeError(f"encoder blob layout mismatch: {b16.off}/{n16} {b32.off}/{n32}")
    return shape, b16.t, b32.t


def decoder_shape(dec, grid: int) -> _lib.SamDecoderShape:
    tr = dec.transformer
    return _lib.SamDecoderShape(C=dec.transformer_dim, heads=tr.num_heads, depth=tr.depth, mlp_dim=tr.mlp_dim,
                                num_mask_tokens=dec.num_mask_tokens,
                                iou_hidden=dec.iou_prediction_head.layers[0].out_features, grid=grid)


def _unwrap(m):
    """peft's ModulesToSaveWrapper keeps the live copy under modules_to_save[active_adapter] (the reference
    tolerates the wrapper with try/except, mask_decoder.py:126-136, :162-169)."""
    mts = getattr(m, "modules_to_save", None)
    if mts is not None:
        name = getattr(m, "active_adapter", None)
        if isinstance(name, (list, tuple)):
            name = name[0]
        if name in mts:
            return mts[name]
    return m


def _decoder_sources(dec):
    """The decoder's tensors in blob order (csrc/decoder.cu carve_weights == state_dict order of mask_decoder.*), the
    two ConvTranspose2d weights re-arranged into per-pixel linear weights."""
    out = []

    def attn(a):
        for lin in (a.q_proj, a.k_proj, a.v_proj, a.out_proj):
            out.append(lin.weight); out.append(lin.bias)

    def norm(ln):
        out.append(ln.weight); out.append(ln.bias)

    out.append(dec.iou_token.weight)
    out.append(_unwrap(dec.mask_tokens).weight)
    tr = dec.transformer
    for layer in tr.layers:
        attn(layer.self_attn); norm(layer.norm1)
        attn(layer.cross_attn_token_to_image); norm(layer.norm2)
        out.extend((layer.mlp.lin1.weight, layer.mlp.lin1.bias, layer.mlp.lin2.weight, layer.mlp.lin2.bias))
        norm(layer.norm3); norm(layer.norm4)
        attn(layer.cross_attn_image_to_token)
    attn(tr.final_attn_token_to_image); norm(tr.norm_final_attn)
    up = _unwrap(dec.output_upscaling)
    # ConvTranspose2d(k=2,s=2) weight [in, out, dy, dx] -> per-pixel linear weight [(dy,dx,out), in]
    w0 = up[0].weight.detach()
    out.append(w0.permute(2, 3, 1, 0).reshape(-1, w0.shape[0]))
    out.append(up[0].bias.detach().repeat(4))
    norm(up[1])
    out.append(up[3].weight.detach().permute(2, 3, 1, 0))     # [(ey,ex), out, in]
    out.append(up[3].bias)
    hyper = _unwrap(dec.output_hypernetworks_mlps)
    for i in range(dec.num_mask_tokens):
        m = hyper[i]
        if len(m.layers) != 3:
            raise RuntimeError("hypernetwork MLPs must have 3 layers")
        for lin in m.layers:
            out.append(lin.weight); out.append(lin.bias)
    head = dec.iou_prediction_head
    if len(head.layers) != 3:
        raise RuntimeError("iou_head_depth must be 3")
    for lin in head.layers:
        out.append(lin.weight); out.append(lin.bias)
    return [t.detach() for t in out]


def pack_decoder(dec, grid: int):
    """-> (shape, fp32 blob) in the order of csrc/decoder.cu carve_weights.  One concatenation (the training path packs
    on every forward: a parameter update must never meet a stale blob)."""
    lib = _lib.load()
    shape = decoder_shape(dec, grid)
    n = lib.sam_decoder_weight_elems(C.byref(shape))
    srcs = _decoder_sources(dec)
    dt = srcs[0].dtype
    if all(t.dtype == dt for t in srcs):
        blob = torch.cat([t.reshape(-1) for t in srcs])
        blob = blob.float() if dt != torch.float32 else blob
    else:
        blob = torch.cat([t.reshape(-1).float() for t in srcs])
    if blob.numel() != n:
        raise RuntimeError(f"decoder blob layout mismatch: packed {blob.numel()}, library expects {n}")
    return shape, blob


def unpack_decoder_grads(dec, gblob: torch.Tensor) -> dict:
    """Inverse of pack_decoder for the gradient blob of sam_decoder_backward: {id(parameter): gradient in the parameter's
    own layout}.  Same walk as pack_decoder; the ConvTranspose2d weights go back to [in, out, dy, dx] and the four
    per-sub-pixel copies of the first ConvTranspose2d bias are summed."""
    out = {}
    off = 0

    def take(p, transform=None, numel=None):
        nonlocal off
        n = p.numel() if numel is None else numel
        g = gblob[off:off + n]
        off += n
        out[id(p)] = transform(g) if transform is not None else g.reshape(p.shape)

    def attn(a):
        for lin in (a.q_proj, a.k_proj, a.v_proj, a.out_proj):
            take(lin.weight); take(lin.bias)

    def norm(ln):
        take(ln.weight); take(ln.bias)

    take(dec.iou_token.weight)
    take(_unwrap(dec.mask_tokens).weight)
    tr = dec.transformer
    for layer in tr.layers:
        attn(layer.self_attn); norm(layer.norm1)
        attn(layer.cross_attn_token_to_image); norm(layer.norm2)
        take(layer.mlp.lin1.weight); take(layer.mlp.lin1.bias)
        take(layer.mlp.lin2.weight); take(layer.mlp.lin2.bias)
        norm(layer.norm3); norm(layer.norm4)
        attn(layer.cross_attn_image_to_token)
    attn(tr.final_attn_token_to_image); norm(tr.norm_final_attn)
    up = _unwrap(dec.output_upscaling)
    cin, cout = up[0].weight.shape[0], up[0].weight.shape[1]
    take(up[0].weight, lambda g: g.reshape(2, 2, cout, cin).permute(3, 2, 0, 1).contiguous())
    take(up[0].bias, lambda g: g.reshape(4, cout).sum(0), numel=4 * cout)
    norm(up[1])
    cin1, cout1 = up[3].weight.shape[0], up[3].weight.shape[1]
    take(up[3].weight, lambda g: g.reshape(2, 2, cout1, cin1).permute(3, 2, 0, 1).contiguous())
    take(up[3].bias)
    hyper = _unwrap(dec.output_hypernetworks_mlps)
    for i in range(dec.num_mask_tokens):
        for lin in hyper[i].layers:
            take(lin.weight); take(lin.bias)
    for lin in dec.iou_prediction_head.layers:
        take(lin.weight); take(lin.bias)
    if off != gblob.numel():
        raise RuntimeError(f"decoder gradient blob layout mismatch: walked {off}, blob has {gblob.numel()}")
    return out

