"""ResizeLongestSide.apply_image: the numpy restatement of Pillow's 8-bit resampler (oracle/resize_oracle.py) is pinned
against PIL itself -- the reference's own code path (utils/transforms.py:27-34 -> torchvision resize of a PIL image) --
and the product's coefficient tables (host logic of the CUDA path) against the oracle's."""
import numpy as np
import pytest

from oracle import resize_oracle as R

SIZES = [(480, 640), (1365, 2048), (333, 517), (1024, 1024), (2000, 1500), (50, 37), (1024, 683), (7, 1024), (427, 640)]


@pytest.mark.parametrize("hw", SIZES)
def test_oracle_equals_pil(hw):
    Image = pytest.importorskip("PIL.Image")
    rng = np.random.default_rng(hw[0] * 7 + hw[1])
    img = rng.integers(0, 256, (hw[0], hw[1], 3), dtype=np.uint8)
    nh, nw = R.get_preprocess_shape(hw[0], hw[1], 1024)
    want = np.array(Image.fromarray(img).resize((nw, nh), Image.BILINEAR))
    assert np.array_equal(R.apply_image(img, 1024), want)


def test_known_answers():
    assert R.get_preprocess_shape(480, 640, 1024) == (768, 1024)          # SURVEY 8c known-answer fact
    assert R.get_preprocess_shape(1365, 2048, 1024) == (683, 1024)
    const = np.full((300, 200, 3), 77, dtype=np.uint8)
    assert (R.apply_image(const, 1024) == 77).all()                        # a constant image stays constant
    b, k = R.pil_bilinear_coeffs(512, 1024)
    assert k.shape[1] == 3 and (k.sum(1) - (1 << 22)).__abs__().max() <= 2    # weights sum to 1 in 22-bit fixed point


@pytest.mark.parametrize("pair", [(640, 1024), (2048, 1024), (1365, 683), (37, 758), (1500, 768), (1024, 1024)])
def test_product_tables_equal_oracle_tables(pair):
    from anyref_b200.segment_anything.utils.transforms import _pil_bilinear_tables

    b1, k1 = _pil_bilinear_tables(*pair)
    b2, k2 = R.pil_bilinear_coeffs(*pair)
    assert np.array_equal(b1, b2) and np.array_equal(k1, k2)
