"""The synthetic checkpoint generator must reproduce the reference state_dict layout exactly."""
import pytest
import torch

from anyref_b200.synthetic import CONFIGS, sam_tensor_specs, synthetic_state_dict
from tests.refutil import build_reference_sam


def test_vit_h_tensor_count_and_params():
    specs = list(sam_tensor_specs(CONFIGS["vit_h"]))
    assert len(specs) == 594
    names = [s[0] for s in specs]
    assert len(set(names)) == len(names)
    import math
    enc = sum(math.prod(s[1]) for s in specs if s[0].startswith("image_encoder."))
    dec = sum(math.prod(s[1]) for s in specs if s[0].startswith("mask_decoder."))
    prm = sum(math.prod(s[1]) for s in specs if s[0].startswith("prompt_encoder.") and "gaussian" not in s[0])
    assert (enc, prm, dec) == (637026048, 6220, 4058340)


def test_synthetic_is_deterministic():
    a = synthetic_state_dict("vit_tiny80", seed=5)
    b = synthetic_state_dict("vit_tiny80", seed=5)
    c = synthetic_state_dict("vit_tiny80", seed=6)
    assert all(torch.equal(a[k], b[k]) for k in a)
    assert any(not torch.equal(a[k], c[k]) for k in a)


@pytest.mark.parametrize("name", ["vit_tiny80", "vit_b", "vit_h"])
def test_layout_matches_reference(ref_sa, name):
    cfg = CONFIGS[name]
    with torch.device("meta"):
        ref = build_reference_sam(ref_sa, cfg)
    ref_sd = ref.state_dict()
    mine = {n: tuple(s) for n, s, _, _ in sam_tensor_specs(cfg)}
    assert list(ref_sd.keys()) == list(mine.keys()) or set(ref_sd.keys()) == set(mine.keys())
    for k, v in ref_sd.items():
        assert tuple(v.shape) == mine[k], k
