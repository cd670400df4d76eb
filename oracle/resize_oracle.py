"""CPU ORACLE (test infrastructure, NOT a product path) -- ResizeLongestSide.apply_image.

The reference resizes on the host: utils/transforms.py:27-34 calls torchvision's `resize(to_pil_image(image), size)`,
i.e. PIL.Image.resize(..., BILINEAR) of an RGB uint8 image.  The arithmetic lives in a third-party dependency absent
from /root/reference: Pillow (pinned Pillow==9.4.0, requirements.txt:16; the resampler is unchanged in the installed 12.x), whose 8-bit
resampler (src/libImaging/Resample.c: precompute_coeffs, normalize_coeffs_8bpc, ImagingResampleHorizontal_8bpc /
Vertical_8bpc) is restated here in numpy:

  * per output index: a triangle filter of support max(scale, 1) around centre (i + 0.5) * scale (antialiasing when
    down-scaling), float64 weights normalised to sum 1, converted to 22-bit fixed point with round-half-away;
  * horizontal pass first, into a uint8 intermediate; then the vertical pass; each output = clip8((2^21 + sum w*p) >> 22).

Pinned by tests/test_resize_oracle.py against PIL itself (the installed Pillow is the reference's own code path).
"""
from __future__ import annotations

import math

import numpy as np

PRECISION_BITS = 32 - 8 - 2


def pil_bilinear_coeffs(in_size: int, out_size: int):
    """-> (bounds int32 [out, 2] = (first source index, tap count), coefficients int32 [out, ksize])."""
    scale = float(np.float32(in_size) - np.float32(0.0)) / out_size
    filterscale = scale if scale >= 1.0 else 1.0
    support = 1.0 * filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    bounds = np.zeros((out_size, 2), dtype=np.int32)
    kk = np.zeros((out_size, ksize), dtype=np.int32)
    ss = 1.0 / filterscale
    for xx in range(out_size):
        center = 0.0 + (xx + 0.5) * scale
        xmin = int(center - support + 0.5)
        if xmin < 0:
            xmin = 0
        xmax = int(center + support + 0.5)
        if xmax > in_size:
            xmax = in_size
        xmax -= xmin
        w = np.zeros(ksize, dtype=np.float64)
        ww = 0.0
        for x in range(xmax):
            a = (x + xmin - center + 0.5) * ss
            if a < 0.0:
                a = -a
            v = 1.0 - a if a < 1.0 else 0.0
            w[x] = v
            ww += v
        for x in range(xmax):
            if ww != 0.0:
                w[x] /= ww
        for x in range(ksize):
            p = w[x] * float(1 << PRECISION_BITS)
            kk[xx, x] = int(-0.5 + p) if w[x] < 0 else int(0.5 + p)
        bounds[xx] = (xmin, xmax)
    return bounds, kk


def _resample_axis0(img: np.ndarray, bounds: np.ndarray, kk: np.ndarray) -> np.ndarray:
    """One separable pass along axis 0 of a uint8 array [n, ...]."""
    out = np.empty((bounds.shape[0],) + img.shape[1:], dtype=np.uint8)
    src = img.astype(np.int64)
    for i in range(bounds.shape[0]):
        x0, n = int(bounds[i, 0]), int(bounds[i, 1])
        acc = np.full(img.shape[1:], 1 << (PRECISION_BITS - 1), dtype=np.int64)
        for t in range(n):
            acc += src[x0 + t] * int(kk[i, t])
        out[i] = np.clip(acc >> PRECISION_BITS, 0, 255).astype(np.uint8)
    return out


def pil_resize_bilinear(image: np.ndarray, new_h: int, new_w: int) -> np.ndarray:
    """HxWxC uint8 -> new_h x new_w x C uint8, as PIL.Image.resize((new_w, new_h), BILINEAR)."""
    h, w = image.shape[:2]
    out = image
    if new_w != w:
        b, k = pil_bilinear_coeffs(w, new_w)
        out = np.ascontiguousarray(_resample_axis0(np.ascontiguousarray(out.transpose(1, 0, 2)), b, k).transpose(1, 0, 2))
    if new_h != h:
        b, k = pil_bilinear_coeffs(h, new_h)
        out = _resample_axis0(out, b, k)
    return np.ascontiguousarray(out)


def get_preprocess_shape(oldh: int, oldw: int, long_side_length: int):
    """utils/transforms.py:102-113."""
    scale = long_side_length * 1.0 / max(oldh, oldw)
    return int(oldh * scale + 0.5), int(oldw * scale + 0.5)


def apply_image(image: np.ndarray, target_length: int = 1024) -> np.ndarray:
    """ResizeLongestSide.apply_image (utils/transforms.py:27-34)."""
    nh, nw = get_preprocess_shape(image.shape[0], image.shape[1], target_length)
    return pil_resize_bilinear(image, nh, nw)
