"""Pins the oracle: the fp32 restatement must be BIT-IDENTICAL to the reference modules imported from
/root/reference (skipped on machines without the reference tree, e.g. the GPU box)."""
import pytest
import torch

from anyref_b200.synthetic import CONFIGS, synthetic_images, synthetic_seg_embeddings, synthetic_state_dict
from oracle import sam_oracle as O
from tests.refutil import build_reference_sam


@pytest.fixture(scope="module")
def tiny(ref_sa):
    cfg = CONFIGS["vit_tiny80"]
    sd = synthetic_state_dict(cfg, seed=1234)
    ref = build_reference_sam(ref_sa, cfg)
    ref.load_state_dict(sd, strict=True)
    return cfg, sd, ref


@torch.no_grad()
def test_encoder_bit_identical(tiny):
    cfg, sd, ref = tiny
    x = synthetic_images(2, seed=0)
    want = ref.image_encoder(x)
    got = O.image_encoder(sd, x, cfg)
    assert got.shape == (2, 256, 64, 64)
    assert torch.equal(got, want)


@torch.no_grad()
@pytest.mark.parametrize("multimask", [False, True])
@pytest.mark.parametrize("n", [1, 4])
def test_decoder_and_postprocess_bit_identical(tiny, n, multimask):
    cfg, sd, ref = tiny
    emb = torch.randn(1, 256, 64, 64, generator=torch.Generator().manual_seed(3))
    seg = synthetic_seg_embeddings(1, n, seed=1)[0]
    sparse_r, dense_r = ref.prompt_encoder(points=None, boxes=None, masks=None, text_embeds=seg)
    sparse_o, dense_o = O.prompt_encoder(sd, cfg, text_embeds=seg)
    assert torch.equal(sparse_r, sparse_o) and torch.equal(dense_r, dense_o)
    pe_r = ref.prompt_encoder.get_dense_pe()
    pe_o = O.dense_pe(sd, cfg)
    assert torch.equal(pe_r, pe_o)
    low_r, iou_r = ref.mask_decoder(image_embeddings=emb, image_pe=pe_r, sparse_prompt_embeddings=sparse_r,
                                    dense_prompt_embeddings=dense_r, multimask_output=multimask)
    low_o, iou_o = O.mask_decoder(sd, cfg, emb, pe_o, sparse_o, dense_o, multimask)
    assert low_r.shape == (n, 3 if multimask else 1, 256, 256)
    assert torch.equal(low_r, low_o) and torch.equal(iou_r, iou_o)
    for inp, orig in [((1024, 1024), (1024, 1024)), ((1024, 683), (640, 427)), ((768, 1024), (480, 640))]:
        pr = ref.postprocess_masks(low_r, input_size=inp, original_size=orig)
        po = O.postprocess_masks(low_o, inp, orig)
        assert torch.equal(pr, po)
        pt = O.postprocess_masks_taps(low_o, inp, orig)
        assert (pt - po).abs().max().item() < 5e-6
        assert torch.equal(pt > 0, po > 0) or ((pt > 0) != (po > 0)).float().mean().item() < 1e-5


@torch.no_grad()
def test_other_prompt_types_bit_identical(tiny):
    cfg, sd, ref = tiny
    g = torch.Generator().manual_seed(11)
    coords = torch.rand(2, 3, 2, generator=g) * 1024
    labels = torch.tensor([[1, 0, 1], [0, 1, -1]])
    boxes = torch.rand(2, 4, generator=g) * 1024
    masks = torch.randn(2, 1, 256, 256, generator=g)
    for kw in (dict(points=(coords, labels)), dict(boxes=boxes), dict(points=(coords, labels), boxes=boxes),
               dict(masks=masks)):
        full = dict(points=None, boxes=None, masks=None, text_embeds=None)
        full.update(kw)
        sr, dr = ref.prompt_encoder(**full)
        so, do = O.prompt_encoder(sd, cfg, **full)
        assert torch.equal(sr, so) and torch.equal(dr, do)


def test_integer_maps_known_answers():
    m = O.window_token_map(64, 14)
    assert m.shape == (25, 196)
    assert O.window_pad(64, 14) == 6
    assert (m == -1).float().mean().item() == pytest.approx(1 - 4096 / 4900)
    valid = m[m >= 0]
    assert torch.equal(valid.sort().values, torch.arange(4096))
    # window 6 = (wy=1, wx=1), token (iy=2, ix=3) -> pixel (16, 17)
    assert m[6, 2 * 14 + 3].item() == 16 * 64 + 17
    r = O.rel_pos_index(14)
    assert r.min().item() == 0 and r.max().item() == 26 and r[0, 13].item() == 0 and r[13, 0].item() == 26
    assert O.rel_pos_index(64).max().item() == 126


def test_window_map_matches_reference_partition(ref_sa):
    from segment_anything.modeling.image_encoder import window_partition, window_unpartition

    x = torch.arange(4096, dtype=torch.float32).reshape(1, 64, 64, 1) + 1.0
    win, pad_hw = window_partition(x, 14)
    assert pad_hw == (70, 70)
    m = O.window_token_map(64, 14)
    want = torch.where(m >= 0, m.float() + 1.0, torch.zeros(()))
    assert torch.equal(win.reshape(25, 196), want)
    assert torch.equal(window_unpartition(win, 14, pad_hw, (64, 64)), x)


def test_rel_pos_matches_reference(ref_sa):
    from segment_anything.modeling.image_encoder import get_rel_pos

    for s in (14, 64):
        t = torch.randn(2 * s - 1, 8)
        assert torch.equal(get_rel_pos(s, s, t), t[O.rel_pos_index(s)])


def test_resize_longest_side_matches_reference(ref_sa):
    """The predictor's host-side geometry (utils/transforms.py:17-113): shapes, coordinate and box maps are identical to
    the reference's ResizeLongestSide, and the resize oracle reproduces the reference's apply_image bit for bit."""
    import numpy as np
    from segment_anything.utils.transforms import ResizeLongestSide as RefResize

    from anyref_b200.segment_anything.utils import ResizeLongestSide

    ref, mine = RefResize(1024), ResizeLongestSide(1024)
    rng = np.random.default_rng(0)
    for (h, w) in ((480, 640), (683, 1024), (1333, 800), (1024, 1024), (37, 1999)):
        assert mine.get_preprocess_shape(h, w, 1024) == ref.get_preprocess_shape(h, w, 1024)
        pts = rng.uniform(0, max(h, w), size=(7, 2))
        assert np.array_equal(mine.apply_coords(pts, (h, w)), ref.apply_coords(pts, (h, w)))
        box = rng.uniform(0, max(h, w), size=(3, 4))
        assert np.array_equal(mine.apply_boxes(box, (h, w)), ref.apply_boxes(box, (h, w)))
        assert torch.equal(mine.apply_boxes_torch(torch.from_numpy(box), (h, w)),
                           ref.apply_boxes_torch(torch.from_numpy(box), (h, w)))
    # the image resize itself runs on the GPU (sam_resize_u8; bit-exactness vs this oracle is a GPU test): here the
    # oracle's restatement of Pillow's resampler is pinned against the REFERENCE's own apply_image
    from oracle import resize_oracle as R

    for (h, w) in ((120, 200), (480, 640), (1365, 2048), (1024, 683), (50, 37)):
        img = rng.integers(0, 256, size=(h, w, 3), dtype=np.uint8)
        assert np.array_equal(R.apply_image(img, 1024), ref.apply_image(img))


@torch.no_grad()
def test_encoder_bit_identical_head_dim_64(ref_sa):
    """The ViT-L / ViT-B family (head_dim 64): the restatement stays bit-identical to the reference encoder."""
    cfg = CONFIGS["vit_tiny64"]
    sd = synthetic_state_dict(cfg, seed=1234)
    ref = build_reference_sam(ref_sa, cfg)
    ref.load_state_dict(sd, strict=True)
    x = synthetic_images(1, seed=3)
    assert torch.equal(O.image_encoder(sd, x, cfg), ref.image_encoder(x))
