"""ctypes binding of libanyref_sam.so (C ABI declared in include/anyref_sam.h).

There is deliberately no fallback: if the library is missing or a call fails, a RuntimeError is raised.
"""
from __future__ import annotations

import ctypes as C
import functools
import os
from pathlib import Path

_PKG = Path(__file__).resolve().parent
# $ANYREF_SAM_LIB selects another build of the same ABI (A/B timing of kernel variants on one box)
LIB_PATH = Path(os.environ["ANYREF_SAM_LIB"]) if os.environ.get("ANYREF_SAM_LIB") else _PKG / "libanyref_sam.so"

_lib = None

c_void_p, c_int, c_float, c_char_p, c_size_t = C.c_void_p, C.c_int, C.c_float, C.c_char_p, C.c_size_t


class SamEncoderShape(C.Structure):
    """Mirror of `struct SamEncoderShape` (include/anyref_sam.h)."""
    _fields_ = [("embed_dim", c_int), ("depth", c_int), ("heads", c_int), ("mlp_dim", c_int), ("img", c_int),
                ("patch", c_int), ("window", c_int), ("out_chans", c_int), ("fmt", c_int),
                ("global_mask", C.c_ulonglong), ("tap_block", c_int), ("tap_out", c_void_p), ("ln_fold", c_int)]


class SamDecoderShape(C.Structure):
    """Mirror of `struct SamDecoderShape` (include/anyref_sam.h)."""
    _fields_ = [("C", c_int), ("heads", c_int), ("depth", c_int), ("mlp_dim", c_int), ("num_mask_tokens", c_int),
                ("iou_hidden", c_int), ("grid", c_int)]


_ENC_P, _DEC_P = C.POINTER(SamEncoderShape), C.POINTER(SamDecoderShape)

# name -> argtypes (restype is int unless listed in _RESTYPES)
_PROTOS = {
    "sam_last_error": [],
    "sam_abi_version": [],
    "sam_gemm": [c_void_p, c_int, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p, c_int, c_int, c_void_p,
                 c_int, c_void_p, c_int, c_int, c_void_p],
    "sam_gemm_residual_ln": [c_void_p, c_int, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p, c_int, c_void_p,
                             c_void_p, c_int, c_void_p, c_void_p],
    "sam_cast_stats": [c_void_p, c_int, c_void_p, c_int, c_int, c_void_p, c_int, c_int, c_void_p],
    "sam_gemm_ln": [c_void_p, c_int, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p, c_int, c_int, c_void_p,
                    c_void_p, c_void_p, c_int, c_float, c_int, c_void_p],
    "sam_umma_probe": [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int,
                       c_void_p],
    "sam_layernorm": [c_void_p, c_int, c_void_p, c_int, c_void_p, c_void_p, c_float, c_void_p, c_int, c_int, c_int, c_int,
                      c_int, c_void_p],
    "sam_patch_im2col": [c_void_p, c_int, c_void_p, c_int, c_int, c_int, c_int, c_void_p],
    "sam_im2col3x3": [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p],
    "sam_ln_nhwc_to_nchw": [c_void_p, c_void_p, c_void_p, c_float, c_void_p, c_int, c_int, c_int, c_int, c_void_p],
    "sam_attn_window": [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p],
    "sam_attn_global": [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p],
    "sam_encoder_w16_elems": [_ENC_P],
    "sam_encoder_w32_elems": [_ENC_P],
    "sam_encoder_workspace_bytes": [_ENC_P, c_int],
    "sam_encoder_forward": [_ENC_P, c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p, c_int, c_void_p, c_size_t,
                            c_void_p],
    "sam_decoder_weight_elems": [_DEC_P],
    "sam_decoder_workspace_bytes": [_DEC_P, c_int, c_int, c_int],
    "sam_decoder_derived_bytes": [_DEC_P],
    "sam_decoder_prepare": [_DEC_P, c_void_p, c_void_p, c_int, c_void_p, c_void_p],
    "sam_decoder_forward": [_DEC_P, c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p, c_int, c_int, c_int,
                            c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_int, c_void_p, c_size_t, c_void_p],
    "sam_decoder_train_workspace_bytes": [_DEC_P, c_int, c_int],
    "sam_decoder_train_forward": [_DEC_P, c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p,
                                  c_int, c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_size_t, C.POINTER(c_void_p), c_void_p],
    "sam_decoder_backward": [c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p],
    "sam_decoder_tape_free": [c_void_p],
    "sam_linear_f32_scratch_bytes": [c_int, c_int, c_int],
    "sam_linear_f32_forward": [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_size_t, c_void_p],
    "sam_linear_f32_backward": [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int,
                                c_void_p, c_size_t, c_void_p],
    "sam_postprocess_masks_backward": [c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p],
    "sam_postprocess_masks": [c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p,
                              c_float, c_void_p],
    "sam_gemm_set_tile_split": [c_int],
    "sam_gemm_schedule": [c_int, c_int, c_int, c_int, c_void_p, c_int],
    "sam_launch_count": [],
    "sam_profile_enable": [c_int],
    "sam_profile_reset": [],
    "sam_profile_collect": [],
    "sam_profile_get": [c_int, C.POINTER(C.c_double), C.POINTER(C.c_longlong), C.POINTER(C.c_double),
                        C.POINTER(C.c_double)],
    "sam_dense_pe": [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p],
    "sam_postprocess_masks_iou": [c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p,
                                  c_float, c_void_p, c_void_p, c_void_p],
    "sam_postprocess_masks_packed": [c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p, c_float,
                                     c_void_p, c_void_p, c_void_p],
    "sam_iou_finalize": [c_void_p, c_int, c_void_p, c_void_p],
    "sam_prompt_sparse": [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_int,
                          c_int, c_int, c_void_p],
    "sam_prompt_mask_blob_elems": [c_int, c_int],
    "sam_prompt_mask_embed": [c_void_p, c_int, c_void_p, c_int, c_void_p, c_int, c_int, c_int, c_int, c_void_p],
    "sam_resize_u8": [c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p, c_int, c_void_p,
                      c_void_p, c_int, c_void_p],
    "sam_preprocess": [c_void_p, c_int, c_void_p, c_int, c_int, c_int, c_int, c_int, C.POINTER(C.c_float),
                       C.POINTER(C.c_float), c_void_p],
}
_RESTYPES = {"sam_last_error": c_char_p, "sam_encoder_w16_elems": c_size_t, "sam_encoder_w32_elems": c_size_t,
             "sam_encoder_workspace_bytes": c_size_t, "sam_decoder_weight_elems": c_size_t,
             "sam_decoder_workspace_bytes": c_size_t, "sam_decoder_derived_bytes": c_size_t,
             "sam_decoder_train_workspace_bytes": c_size_t, "sam_linear_f32_scratch_bytes": c_size_t, "sam_decoder_tape_free": None, "sam_prompt_mask_blob_elems": c_size_t, "sam_gemm_set_tile_split": None, "sam_launch_count": C.c_longlong, "sam_profile_enable": None,
             "sam_profile_reset": None, "sam_profile_get": None}


def exported_symbols() -> list[str]:
    return sorted(_PROTOS)


def load() -> C.CDLL:
    """Load the shared library (building it is the job of anyref_b200.build / __graft_entry__.build)."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise RuntimeError(
            f"{LIB_PATH} not found: run `python -m anyref_b200.build` (there is no CPU or PyTorch fallback)")
    lib = C.CDLL(os.fspath(LIB_PATH))
    for name, argtypes in _PROTOS.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is missing
        fn.argtypes = argtypes
        fn.restype = _RESTYPES.get(name, c_int)
    _lib = lib
    return lib


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = load().sam_last_error().decode("utf-8", "replace")
        raise RuntimeError(f"{what} failed (code {rc}): {msg}")


def ptr(t) -> int | None:
    """Device pointer of a torch tensor (None -> NULL)."""
    if t is None:
        return None
    return t.data_ptr()


def stream_ptr(device=None) -> int:
    import torch

    return torch.cuda.current_stream(device).cuda_stream


def _first_cuda_device(objs, depth: int = 0):
    import torch

    for o in objs:
        if isinstance(o, torch.Tensor):
            if o.is_cuda:
                return o.device
        elif isinstance(o, torch.nn.Module):
            for t in o.parameters():
                return t.device if t.is_cuda else None
            for t in o.buffers():
                return t.device if t.is_cuda else None
        elif isinstance(o, (list, tuple)) and depth < 2:
            d = _first_cuda_device(o, depth + 1)
            if d is not None:
                return d
    return None


def device_scoped(fn):
    """Decorator for every Python entry point that ends in a C-ABI call: the library launches on the CURRENT device
    (kernel attributes, tensor maps and streams are per device), so the device of the first CUDA tensor argument (or
    of the module's parameters) is made current for the duration of the call -- a model on cuda:1 works while cuda:0 is
    current, as with the reference's PyTorch modules."""

    @functools.wraps(fn)
    def wrapper(*args, **kwargs):
        import torch

        dev = _first_cuda_device(list(args[1:]) + list(kwargs.values()) + list(args[:1]))
        if dev is None or dev.index is None or dev.index == torch.cuda.current_device():
            return fn(*args, **kwargs)
        with torch.cuda.device(dev):
            return fn(*args, **kwargs)

    return wrapper


FMT_F16, FMT_BF16, FMT_F32 = 0, 1, 2


def fmt_of(dtype) -> int:
    import torch

    if dtype == torch.float16:
        return FMT_F16
    if dtype == torch.bfloat16:
        return FMT_BF16
    if dtype == torch.float32:
        return FMT_F32
    raise TypeError(f"unsupported dtype {dtype}")


KERNEL_CLASSES = ("gemm", "attn_window", "attn_global", "layernorm", "layout", "decoder", "postprocess")


def gemm_set_tile_split(mode: int) -> None:
    """Half-tile items in the last round of the 2-CTA GEMM (include/anyref_sam.h): -1 policy, 0 never, 1 whenever possible."""
    load().sam_gemm_set_tile_split(int(mode))


def launch_count() -> int:
    return int(load().sam_launch_count())


def profile_enable(on: bool) -> None:
    load().sam_profile_enable(1 if on else 0)


def profile_reset() -> None:
    load().sam_profile_reset()


def profile_read() -> dict:
    """Synchronises the recorded event pairs and returns {class: {ms, launches, flops, bytes}} since the last reset."""
    lib = load()
    check(lib.sam_profile_collect(), "sam_profile_collect")
    out = {}
    for i, name in enumerate(KERNEL_CLASSES):
        ms, n, fl, by = C.c_double(), C.c_longlong(), C.c_double(), C.c_double()
        lib.sam_profile_get(i, C.byref(ms), C.byref(n), C.byref(fl), C.byref(by))
        out[name] = {"ms": ms.value, "launches": n.value, "flops": fl.value, "bytes": by.value}
    return out
