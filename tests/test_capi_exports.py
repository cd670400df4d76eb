"""The C-ABI library must build, load and export every symbol include/anyref_sam.h declares (no compute: no GPU here)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions():
    text = open(os.path.join(ROOT, "include", "anyref_sam.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(sam_[a-z0-9_]+)\s*\(", text)))


@pytest.fixture(scope="module")
def lib():
    from anyref_b200 import build

    path = build.build()
    return ctypes.CDLL(os.fspath(path))


def test_header_declares_the_path_entry_points():
    names = declared_functions()
    for need in ("sam_encoder_forward", "sam_decoder_forward", "sam_postprocess_masks", "sam_dense_pe", "sam_gemm",
                 "sam_attn_window", "sam_attn_global", "sam_layernorm", "sam_last_error"):
        assert need in names


def test_every_declared_symbol_is_exported(lib):
    for name in declared_functions():
        assert hasattr(lib, name), f"{name} declared in include/anyref_sam.h but not exported"


def test_python_binding_covers_the_header(lib):
    from anyref_b200 import _lib

    assert set(_lib.exported_symbols()) == set(declared_functions())
    _lib.load()


def test_sizes_and_errors_without_gpu(lib):
    from anyref_b200 import _lib

    L = _lib.load()
    assert L.sam_abi_version() == 4
    enc = _lib.SamEncoderShape(embed_dim=1280, depth=32, heads=16, mlp_dim=5120, img=1024, patch=16, window=14,
                               out_chans=256, fmt=1, global_mask=0, tap_block=-1, tap_out=None)
    n16 = L.sam_encoder_w16_elems(ctypes.byref(enc))
    # all encoder matrices + per-block operand bias and rel-pos slot
    mats = 1280 * 768 + 32 * (3 * 1280 * 1280 + 1280 * 1280 + 2 * 5120 * 1280) + 256 * 1280 + 256 * 2304
    assert n16 == mats + 32 * (3 * 1280 + 2 * 128 * 80)
    assert L.sam_encoder_workspace_bytes(ctypes.byref(enc), 16) > 16 * 4096 * 1280 * 4
    dec = _lib.SamDecoderShape(C=256, heads=8, depth=2, mlp_dim=2048, num_mask_tokens=4, iou_hidden=256, grid=64)
    # reference decoder parameter count + the 3 extra copies of the first ConvTranspose bias (one per output sub-pixel)
    assert L.sam_decoder_weight_elems(ctypes.byref(dec)) == 4058340 + 3 * 64
    # argument validation happens before any CUDA call
    rc = L.sam_gemm(None, 0, None, 0, 0, 0, 0, 7, None, 0, 0, None, 0, None, 0, 0, None)
    assert rc != 0 and b"fmt" in L.sam_last_error()
    rc = L.sam_postprocess_masks(None, 2, 0, 256, 1024, 1024, 1024, 1024, 1024, None, None, 0.0, None)
    assert rc != 0 and b"postprocess" in L.sam_last_error()
    # mask decoder: sizes grow with the problem, NULL arguments and unsupported shapes are refused before any launch
    small = L.sam_decoder_workspace_bytes(ctypes.byref(dec), 1, 1, 1)
    assert small > 4096 * 256 * 4 * 3
    assert L.sam_decoder_workspace_bytes(ctypes.byref(dec), 1, 16, 1) > 8 * small // 2
    assert L.sam_decoder_workspace_bytes(ctypes.byref(dec), 16, 4, 1) >= L.sam_decoder_workspace_bytes(ctypes.byref(dec), 4, 4, 1)
    derived = L.sam_decoder_derived_bytes(ctypes.byref(dec))
    assert derived > 2 * (2 * 4096 * 128 * 4) and derived % 256 == 0      # the pe.W^T tables alone are 2 MB each
    rc = L.sam_decoder_prepare(ctypes.byref(dec), None, None, 2, None, None)
    assert rc != 0 and b"NULL" in L.sam_last_error()
    rc = L.sam_decoder_forward(ctypes.byref(dec), None, None, None, 2, 1, None, None, 2, 1, 1, None, None, 2, None, None, 2,
                               None, 0, None)
    assert rc != 0 and b"NULL" in L.sam_last_error()
    rc = L.sam_resize_u8(None, 10, 10, 3, None, None, 20, 20, None, None, 0, None, None, 0, None)
    assert rc != 0 and b"resize" in L.sam_last_error()
    # training path: the workspace covers activations + gradients + scratch and grows with the number of prompts
    t1 = L.sam_decoder_train_workspace_bytes(ctypes.byref(dec), 1, 1)
    t4 = L.sam_decoder_train_workspace_bytes(ctypes.byref(dec), 4, 1)
    assert t1 > 64 << 20 and 3 * t1 < t4 < 5 * t1
    tape = ctypes.c_void_p()
    rc = L.sam_decoder_train_forward(ctypes.byref(dec), None, None, 2, 1, None, None, 1, 1, None, None, 2, None, 2, None, None, None,
                                     0, ctypes.byref(tape), None)
    assert rc != 0 and b"NULL" in L.sam_last_error() and not tape.value
    rc = L.sam_decoder_backward(None, None, 0, 4, None, None, None, None)
    assert rc != 0 and b"tape" in L.sam_last_error()
    L.sam_decoder_tape_free(None)
    rc = L.sam_postprocess_masks_backward(None, 1, 256, 1024, 1024, 1024, 1024, 1024, None, None, None)
    assert rc != 0 and b"postprocess_backward" in L.sam_last_error()
    rc = L.sam_linear_f32_backward(None, None, None, None, None, None, None, 1, 1, 1, None, 0, None)
    assert rc != 0 and b"linear_f32_backward" in L.sam_last_error()
    bad = _lib.SamDecoderShape(C=192, heads=8, depth=2, mlp_dim=2048, num_mask_tokens=4, iou_hidden=256, grid=64)
    buf = ctypes.create_string_buffer(1 << 12)                       # any non-NULL host pointer: refused before use
    ptr = ctypes.cast(buf, ctypes.c_void_p)
    rc = L.sam_decoder_prepare(ctypes.byref(bad), ptr, ptr, 2, ptr, None)
    assert rc != 0 and b"transformer_dim" in L.sam_last_error()


def test_product_path_has_no_cpu_fallback():
    import torch

    from anyref_b200.segment_anything import build_sam_from_config
    from anyref_b200.synthetic import CONFIGS

    sam = build_sam_from_config(CONFIGS["vit_tiny80"])
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        sam.image_encoder(torch.zeros(1, 3, 1024, 1024))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        sam.postprocess_masks(torch.zeros(1, 1, 256, 256), (1024, 1024), (1024, 1024))


def test_product_does_not_import_the_oracle():
    pkg = os.path.join(ROOT, "anyref_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cpp", ".h", ".cuh")):
                src = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in src and "from oracle" not in src, f


def test_integration_md_struct_stub_matches_the_binding():
    """The ctypes struct shown to a maintainer in INTEGRATION.md must have the fields of the real binding (a missing
    trailing field yields a short struct and undefined behaviour in sam_encoder_forward)."""
    import re
    from anyref_b200 import _lib

    text = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    block = text[text.index("class SamEncoderShape(C.Structure)"):]
    block = block[:block.index("]") + 1]
    doc_fields = re.findall(r'\("(\w+)",', block)
    assert doc_fields == [f[0] for f in _lib.SamEncoderShape._fields_]


def test_gemm_schedule_covers_every_tile_exactly_once(lib):
    """The persistent 2-CTA GEMM's work list (csrc/gemm2.cu, Sched; host copy of the function the kernel calls): whole
    256 x 256 tiles round-robin over the CTA pairs and, when the last round would leave at least half of the pairs idle,
    that round as 256 x 128 half items -- one per pair, always a pair's LAST item (the residual producers' cross-warp
    statistics hand-off relies on that).  Every tile must be covered exactly once: as a whole, or as both of its halves."""
    from anyref_b200 import _lib

    L = _lib.load()
    cap = 1024
    buf = (ctypes.c_int * (4 * cap))()

    def items(tiles, pairs, pair, split):
        n = L.sam_gemm_schedule(tiles, pairs, pair, split, ctypes.cast(buf, ctypes.c_void_p), cap)
        assert n >= 0
        return [(buf[4 * i], buf[4 * i + 1], buf[4 * i + 2]) for i in range(n)]

    # ViT-H linears at 1 / 2 / 16 images on 74 pairs, a smaller device, tiny launches, and a sweep
    cases = [(t, p) for p in (74, 72, 37, 8) for t in (1, 2, 5, 6, 16, 37, 38, 80, 160, 240, 320, 325, 1280, 3840, 5120)]
    cases += [(t, 74) for t in range(1, 400)]
    for tiles, pairs in cases:
        for split in (0, 1):
            # the host launches min(tiles, pairs) pairs, or 2 * tiles when all of them fit as half items
            grid = min(tiles, pairs)
            if split and 2 * tiles <= pairs:
                grid = 2 * tiles
            whole, halves, loads = {}, {}, []
            for pair in range(grid):
                its = items(tiles, grid, pair, split)
                loads.append(sum(0.5 if hf else 1.0 for _, hf, _ in its))
                for i, (t, hf, h) in enumerate(its):
                    assert 0 <= t < tiles
                    if hf:
                        assert split and i == len(its) - 1 and h in (0, 1)
                        halves.setdefault(t, []).append(h)
                    else:
                        assert t == pair + i * grid
                        whole[t] = whole.get(t, 0) + 1
            assert all(v == 1 for v in whole.values()) and all(sorted(v) == [0, 1] for v in halves.values())
            assert not (set(whole) & set(halves)) and len(whole) + len(halves) == tiles
            rem = tiles % grid
            if split and 0 < rem <= grid // 2 or (split and grid == 2 * tiles):
                assert len(halves) == (tiles if grid == 2 * tiles else rem)
                assert max(loads) == tiles // grid + 0.5        # the last round costs half a tile period
            else:
                assert not halves and max(loads) == -(-tiles // grid)
