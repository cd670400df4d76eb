"""Drop-in boundary: same constructors, attribute tree and state_dict layout as the reference package."""
import pytest
import torch

from anyref_b200.segment_anything import build_sam_from_config, sam_model_registry
from anyref_b200.synthetic import CONFIGS, sam_tensor_specs, synthetic_state_dict
from tests.refutil import build_reference_sam


@pytest.mark.parametrize("name", ["vit_tiny80", "vit_b"])
def test_state_dict_matches_spec(name):
    cfg = CONFIGS[name]
    with torch.device("meta"):
        sam = build_sam_from_config(cfg)
    mine = {k: tuple(v.shape) for k, v in sam.state_dict().items()}
    want = {n: tuple(s) for n, s, _, _ in sam_tensor_specs(cfg)}
    assert list(mine) == list(want)
    assert mine == want


def test_state_dict_round_trip_strict():
    cfg = CONFIGS["vit_tiny80"]
    sd = synthetic_state_dict(cfg)
    sam = build_sam_from_config(cfg)
    res = sam.load_state_dict(sd, strict=True)
    assert not res.missing_keys and not res.unexpected_keys
    back = sam.state_dict()
    assert all(torch.equal(back[k], sd[k]) for k in sd)


def test_attribute_contract_used_by_anyref():
    """model/anyref.py:106-113, :368, :413-429: attributes and toggles the caller relies on."""
    sam = build_sam_from_config(CONFIGS["vit_tiny80"])
    assert not sam.training
    assert sam.image_encoder.img_size == 1024
    assert sam.mask_threshold == 0.0
    for p in sam.parameters():
        p.requires_grad = False
    sam.mask_decoder.train()
    for p in sam.mask_decoder.parameters():
        p.requires_grad = True
    assert sam.mask_decoder.training and callable(sam.postprocess_masks) and callable(sam.prompt_encoder.get_dense_pe)
    assert set(sam_model_registry) == {"default", "vit_h", "vit_l", "vit_b"}
    # non-persistent buffers exactly as the reference (sam.py:46-49)
    assert "pixel_mean" not in sam.state_dict() and sam.pixel_mean.shape == (3, 1, 1)


def test_prompt_encoder_text_embeds_plumbing():
    """prompt_encoder.py:164-186 with text_embeds only: fp32 promotion of the sparse part, stride-0 dense broadcast."""
    sam = build_sam_from_config(CONFIGS["vit_tiny80"])
    text = torch.randn(3, 1, 256).to(torch.bfloat16)
    sparse, dense = sam.prompt_encoder(points=None, boxes=None, masks=None, text_embeds=text)
    assert sparse.dtype == torch.float32 and sparse.shape == (3, 1, 256)
    assert torch.equal(sparse, text.float())
    assert dense.shape == (3, 256, 64, 64) and dense.stride() == (0, 1, 0, 0)
    assert torch.equal(dense[1, :, 5, 7], sam.prompt_encoder.no_mask_embed.weight[0])
    # point / box / mask prompts are CUDA kernels: on CPU tensors they fail loudly instead of falling back
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        sam.prompt_encoder(points=(torch.zeros(1, 1, 2), torch.zeros(1, 1)), boxes=None, masks=None, text_embeds=None)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        sam.prompt_encoder(points=None, boxes=None, masks=torch.zeros(1, 1, 256, 256), text_embeds=None)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        sam.preprocess(torch.zeros(3, 8, 8))


def test_decoder_training_route_and_frozen_inputs():
    """model/anyref.py:108-113, :413-429 fine-tunes the decoder (train() + requires_grad) and back-propagates into the
    [SEG] embedding: those calls take the training path (csrc/decoder_train.cu), everything else the fused inference
    path.  Both are CUDA-only (no CPU fallback), and the training path refuses to pretend it has a gradient for the
    inputs the reference keeps frozen."""
    sam = build_sam_from_config(CONFIGS["vit_tiny80"])
    emb, pe = torch.zeros(1, 256, 64, 64), torch.zeros(1, 256, 64, 64)
    text = torch.randn(2, 1, 256)
    sparse, dense = sam.prompt_encoder(points=None, boxes=None, masks=None, text_embeds=text)
    sparse, dense = sparse.detach(), dense.detach()
    call = dict(image_pe=pe, dense_prompt_embeddings=dense, multimask_output=False)
    sam.mask_decoder.train()
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        sam.mask_decoder(image_embeddings=emb, sparse_prompt_embeddings=sparse, **call)
    with pytest.raises(NotImplementedError, match="no gradient is produced for image_embeddings"):
        sam.mask_decoder(image_embeddings=emb.clone().requires_grad_(True), sparse_prompt_embeddings=sparse, **call)
    sam.mask_decoder.eval()
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        sam.mask_decoder(image_embeddings=emb, sparse_prompt_embeddings=sparse.clone().requires_grad_(True), **call)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        sam.mask_decoder(image_embeddings=emb, sparse_prompt_embeddings=sparse, **call)
    sam.mask_decoder.train()
    with torch.no_grad(), pytest.raises(RuntimeError, match="no CPU fallback"):
        sam.mask_decoder(image_embeddings=emb, sparse_prompt_embeddings=sparse, **call)


@pytest.mark.parametrize("name", ["vit_tiny80", "vit_h"])
def test_layout_matches_reference_modules(ref_sa, name):
    cfg = CONFIGS[name]
    with torch.device("meta"):
        ref = build_reference_sam(ref_sa, cfg)
        mine = build_sam_from_config(cfg)
    a, b = ref.state_dict(), mine.state_dict()
    assert list(a) == list(b)
    assert all(a[k].shape == b[k].shape and a[k].dtype == b[k].dtype for k in a)


def test_gradient_blob_unpacking_is_the_adjoint_of_weight_packing():
    """sam_decoder_backward returns d(loss)/d(weight blob); unpack_decoder_grads must map it to parameter gradients by
    the transpose of pack_decoder (a linear map: ConvTranspose2d permutations, the first ConvTranspose2d bias repeated
    four times): <pack(P), B> == sum_p <p, unpack(B)[p]> for random P and B."""
    from anyref_b200.segment_anything import _pack

    sam = build_sam_from_config(CONFIGS["vit_tiny80"])
    dec = sam.mask_decoder
    g = torch.Generator().manual_seed(0)
    with torch.no_grad():
        for p in dec.parameters():
            p.copy_(torch.randn(p.shape, generator=g))
    _, blob = _pack.pack_decoder(dec, 64)
    B = torch.randn(blob.shape, generator=g)
    grads = _pack.unpack_decoder_grads(dec, B)
    lhs = float((blob.double() * B.double()).sum())
    rhs = sum(float((p.detach().double() * grads[id(p)].double()).sum()) for p in dec.parameters())
    assert len(grads) == len(list(dec.parameters()))
    assert all(grads[id(p)].shape == p.shape for p in dec.parameters())
    assert abs(lhs - rhs) <= 1e-9 * max(1.0, abs(lhs))

