"""autograd bridges of the training path: mask decoder, postprocess_masks and fp32 linear.

AnyRef fine-tunes `visual_model.mask_decoder` and `text_hidden_fcs` with everything else frozen
(model/anyref.py:108-127) and takes the mask loss on `postprocess_masks(mask_decoder(...))` (model/anyref.py:406-450).
Each Function below calls the C-ABI forward that keeps what its backward needs, and the matching C-ABI backward
(csrc/decoder_train.cu).  Nothing here computes with PyTorch ops except the re-arrangement of the gradient blob into
the parameters' own layouts.
"""
from __future__ import annotations

import ctypes as C

import torch

from .. import _lib
from . import _pack


class _Tape:
    """Owns the host-side tape of one decoder forward; freed after backward or when the graph is dropped."""

    def __init__(self, ptr: int, workspace: torch.Tensor):
        self.ptr = ptr
        self.workspace = workspace      # activations + gradient region: must outlive the tape

    def free(self) -> None:
        if self.ptr:
            _lib.load().sam_decoder_tape_free(self.ptr)
            self.ptr = None
        self.workspace = None

    def __del__(self):  # pragma: no cover - interpreter shutdown order
        try:
            self.free()
        except Exception:
            pass


class DecoderTrainFn(torch.autograd.Function):
    """(masks [n,nm,4g,4g], iou [n,nm]) = MaskDecoder.predict_masks(...) in fp32 with a gradient for the decoder's
    parameters and for sparse_prompt_embeddings."""

    @staticmethod
    def forward(ctx, dec, mask_range, emb, pe, sparse, dense_vec, dense_full, image_index, *params):
        lib = _lib.load()
        ctx.set_materialize_grads(False)      # an unused output (the IoU prediction in AnyRef) arrives as None
        ctx.mask_range = mask_range
        dev = emb.device
        with torch.cuda.device(dev):
            g = emb.shape[-1]
            shape, blob = _pack.pack_decoder(dec, g)
            n, k = sparse.shape[0], sparse.shape[1]
            sp = sparse.detach().float().contiguous()
            nm = dec.num_mask_tokens
            masks = torch.empty((n, nm, 4 * g, 4 * g), device=dev, dtype=torch.float32)
            iou = torch.empty((n, nm), device=dev, dtype=torch.float32)
            nbytes = lib.sam_decoder_train_workspace_bytes(C.byref(shape), n, k)
            if nbytes == 0:
                raise RuntimeError(f"sam_decoder_train_workspace_bytes: {lib.sam_last_error().decode()}")
            ws = torch.empty(nbytes + 256, dtype=torch.uint8, device=dev)
            dense = dense_vec if dense_vec is not None else dense_full
            tape = C.c_void_p()
            rc = lib.sam_decoder_train_forward(
                C.byref(shape), blob.data_ptr(), emb.data_ptr(), _lib.fmt_of(emb.dtype), emb.shape[0],
                image_index.data_ptr() if image_index is not None else None, sp.data_ptr() if k > 0 else None, n, k,
                dense_vec.data_ptr() if dense_vec is not None else None,
                dense_full.data_ptr() if dense_full is not None else None, _lib.fmt_of(dense.dtype),
                pe.data_ptr(), _lib.fmt_of(pe.dtype), masks.data_ptr(), iou.data_ptr(), ws.data_ptr(), ws.numel(),
                C.byref(tape), _lib.stream_ptr(dev))
            _lib.check(rc, "sam_decoder_train_forward")
        ctx.tape = _Tape(tape.value, ws)
        ctx.dec = dec
        ctx.blob_elems = blob.numel()
        ctx.sparse_meta = (sparse.shape, sparse.dtype)
        ctx.params = params
        ctx.keep = (emb, pe, dense, image_index, blob)     # read again by nothing, but the tape holds raw pointers
        return masks, iou

    @staticmethod
    def backward(ctx, d_masks, d_iou):
        tape = ctx.tape
        if tape is None or not tape.ptr:
            raise RuntimeError("MaskDecoder backward: the tape of this forward has already been consumed "
                               "(retain_graph / double backward are not supported)")
        lib = _lib.load()
        dev = tape.workspace.device
        shape, dtype = ctx.sparse_meta
        with torch.cuda.device(dev):
            dm = d_masks.float().contiguous() if d_masks is not None else None
            di = d_iou.float().contiguous() if d_iou is not None else None
            gblob = torch.zeros(ctx.blob_elems, dtype=torch.float32, device=dev)
            d_sparse = torch.empty(shape, dtype=torch.float32, device=dev)
            rc = lib.sam_decoder_backward(tape.ptr, dm.data_ptr() if dm is not None else None, ctx.mask_range[0],
                                          ctx.mask_range[1], di.data_ptr() if di is not None else None, gblob.data_ptr(),
                                          d_sparse.data_ptr() if d_sparse.numel() else None, _lib.stream_ptr(dev))
            _lib.check(rc, "sam_decoder_backward")
            # the workspace may be recycled by the allocator as soon as this stream has passed the kernels above
            tape.free()
            ctx.tape = None
            grads = _pack.unpack_decoder_grads(ctx.dec, gblob)
        out = []
        for p in ctx.params:
            gp = grads.get(id(p)) if p.requires_grad else None
            out.append(gp.to(p.dtype) if gp is not None else None)
        ds = d_sparse.to(dtype) if ctx.needs_input_grad[4] else None
        return (None, None, None, None, ds, None, None, None, *out)


class PostprocessFn(torch.autograd.Function):
    """Sam.postprocess_masks with a gradient for the low-resolution logits."""

    @staticmethod
    def forward(ctx, masks, L, S, h_in, w_in, H, W):
        lib = _lib.load()
        m = masks.contiguous()
        n, ch = m.shape[0], m.shape[1]
        out = torch.empty((n, ch, H, W), device=m.device, dtype=torch.float32)
        with torch.cuda.device(m.device):
            if n * ch > 0:
                rc = lib.sam_postprocess_masks(m.data_ptr(), _lib.fmt_of(m.dtype), n * ch, L, S, h_in, w_in, H, W,
                                               out.data_ptr(), None, 0.0, _lib.stream_ptr(m.device))
                _lib.check(rc, "sam_postprocess_masks")
        ctx.meta = (n, ch, L, S, h_in, w_in, H, W, masks.dtype)
        return out

    @staticmethod
    def backward(ctx, d_out):
        n, ch, L, S, h_in, w_in, H, W, dtype = ctx.meta
        lib = _lib.load()
        d = d_out.float().contiguous()
        d_low = torch.zeros((n, ch, L, L), device=d.device, dtype=torch.float32)
        with torch.cuda.device(d.device):
            if n * ch > 0:
                tmp = torch.empty((n * ch, h_in, w_in), device=d.device, dtype=torch.float32)
                rc = lib.sam_postprocess_masks_backward(d.data_ptr(), n * ch, L, S, h_in, w_in, H, W, tmp.data_ptr(),
                                                        d_low.data_ptr(), _lib.stream_ptr(d.device))
                _lib.check(rc, "sam_postprocess_masks_backward")
        return d_low.to(dtype), None, None, None, None, None, None


class LinearF32Fn(torch.autograd.Function):
    """y = x W^T + b (optionally ReLU) in fp32 with gradients for x, W and b (text_hidden_fcs in training)."""

    @staticmethod
    def forward(ctx, x, weight, bias, relu):
        lib = _lib.load()
        xf = x.detach().float().contiguous()
        wf = weight.detach().float().contiguous()
        bf = bias.detach().float().contiguous() if bias is not None else None
        M, K = xf.shape
        N = wf.shape[0]
        y = torch.empty((M, N), device=xf.device, dtype=torch.float32)
        with torch.cuda.device(xf.device):
            nbytes = lib.sam_linear_f32_scratch_bytes(M, N, K)
            scratch = torch.empty(nbytes + 256, dtype=torch.uint8, device=xf.device)
            rc = lib.sam_linear_f32_forward(xf.data_ptr(), wf.data_ptr(), bf.data_ptr() if bf is not None else None,
                                            y.data_ptr(), M, N, K, 1 if relu else 0, scratch.data_ptr(), scratch.numel(),
                                            _lib.stream_ptr(xf.device))
            _lib.check(rc, "sam_linear_f32_forward")
        ctx.save_for_backward(xf, wf, y if relu else None)
        ctx.meta = (x.dtype, weight.dtype, bias.dtype if bias is not None else None, relu)
        return y

    @staticmethod
    def backward(ctx, dy):
        xf, wf, y = ctx.saved_tensors
        xdt, wdt, bdt, relu = ctx.meta
        lib = _lib.load()
        # the kernel masks dY in place with the fused ReLU's (y > 0): work on a private copy of autograd's tensor
        d = dy.float().clone(memory_format=torch.contiguous_format) if relu else dy.float().contiguous()
        M, K = xf.shape
        N = wf.shape[0]
        dev = xf.device
        need_x, need_w, need_b = ctx.needs_input_grad[0], ctx.needs_input_grad[1], ctx.needs_input_grad[2]
        dx = torch.empty((M, K), device=dev, dtype=torch.float32) if need_x else None
        dw = torch.empty((N, K), device=dev, dtype=torch.float32) if need_w else None
        db = torch.zeros((N,), device=dev, dtype=torch.float32) if need_b and bdt is not None else None
        with torch.cuda.device(dev):
            nbytes = lib.sam_linear_f32_scratch_bytes(M, N, K)
            scratch = torch.empty(nbytes + 256, dtype=torch.uint8, device=dev)
            rc = lib.sam_linear_f32_backward(d.data_ptr(), y.data_ptr() if relu else None, xf.data_ptr(), wf.data_ptr(),
                                             dx.data_ptr() if dx is not None else None,
                                             dw.data_ptr() if dw is not None else None,
                                             db.data_ptr() if db is not None else None, M, N, K, scratch.data_ptr(),
                                             scratch.numel(), _lib.stream_ptr(dev))
            _lib.check(rc, "sam_linear_f32_backward")
        return (dx.to(xdt) if dx is not None else None, dw.to(wdt) if dw is not None else None,
                db.to(bdt) if db is not None else None, None)
