"""ResizeLongestSide (reference: utils/transforms.py:17-113).

`apply_image` resizes ON THE DEVICE, bit-exactly as the reference's PIL path does (utils/transforms.py:27-34:
torchvision `resize(to_pil_image(image), size)` = PIL BILINEAR with antialiasing): the host only builds Pillow's integer
coefficient tables (a few thousand numbers, cached per size pair), the resample itself is `sam_resize_u8`
(csrc/prompt.cu).  Prompts need the coordinate maps below, which are a few scalar multiplications."""
from __future__ import annotations

import math
from functools import lru_cache
from typing import Tuple

import numpy as np
import torch

from ... import _lib

_PRECISION_BITS = 32 - 8 - 2      # Pillow, Resample.c


@lru_cache(maxsize=64)
def _pil_bilinear_tables(in_size: int, out_size: int):
    """Pillow's precompute_coeffs + normalize_coeffs_8bpc for the bilinear (triangle) filter, in the same float64
    operation order: -> (bounds int32 [out, 2], coefficients int32 [out, ksize])."""
    scale = float(in_size) / out_size
    filterscale = scale if scale >= 1.0 else 1.0
    support = filterscale                     # bilinear support 1.0 * filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    bounds = np.zeros((out_size, 2), dtype=np.int32)
    coeff = np.zeros((out_size, ksize), dtype=np.int32)
    ss = 1.0 / filterscale
    one = float(1 << _PRECISION_BITS)
    for xx in range(out_size):
        center = (xx + 0.5) * scale
        xmin = max(int(center - support + 0.5), 0)
        xmax = min(int(center + support + 0.5), in_size) - xmin
        w = []
        ww = 0.0
        for x in range(xmax):
            a = abs((x + xmin - center + 0.5) * ss)
            v = 1.0 - a if a < 1.0 else 0.0
            w.append(v)
            ww += v
        for x in range(xmax):
            v = w[x] / ww if ww != 0.0 else w[x]
            coeff[xx, x] = int(0.5 + v * one)            # triangle weights are never negative
        bounds[xx, 0], bounds[xx, 1] = xmin, xmax
    return bounds, coeff


class ResizeLongestSide:
    """Scale so that the longest side equals `target_length`; coordinates and boxes follow the same map."""

    def __init__(self, target_length: int) -> None:
        self.target_length = target_length

    @staticmethod
    def get_preprocess_shape(oldh: int, oldw: int, long_side_length: int) -> Tuple[int, int]:
        """utils/transforms.py:102-113: round-half-up of the scaled sides."""
        scale = long_side_length * 1.0 / max(oldh, oldw)
        return int(oldh * scale + 0.5), int(oldw * scale + 0.5)

    def _device_tables(self, in_size: int, out_size: int, device):
        key = (in_size, out_size, str(device))
        cache = self.__dict__.setdefault("_tab", {})
        if key not in cache:
            if len(cache) > 32:
                cache.clear()
            b, k = _pil_bilinear_tables(in_size, out_size)
            cache[key] = (torch.from_numpy(b).to(device), torch.from_numpy(k).to(device), k.shape[1])
        return cache[key]

    @_lib.device_scoped
    def apply_image_cuda(self, image: torch.Tensor) -> torch.Tensor:
        """HxWxC uint8 CUDA tensor -> resized HxWxC uint8 CUDA tensor, bit-exact with the reference's PIL resize."""
        if not (image.is_cuda and image.dtype == torch.uint8 and image.dim() == 3):
            raise ValueError("apply_image_cuda expects an HxWxC uint8 CUDA tensor")
        img = image.contiguous()
        H, W, Cc = img.shape
        newh, neww = self.get_preprocess_shape(H, W, self.target_length)
        out = torch.empty((newh, neww, Cc), dtype=torch.uint8, device=img.device)
        xb = xk = yb = yk = None
        kx = ky = 0
        if neww != W:
            xb, xk, kx = self._device_tables(W, neww, img.device)
        if newh != H:
            yb, yk, ky = self._device_tables(H, newh, img.device)
        tmp = torch.empty((H, neww, Cc), dtype=torch.uint8, device=img.device) if (neww != W and newh != H) else None
        rc = _lib.load().sam_resize_u8(img.data_ptr(), H, W, Cc, _lib.ptr(tmp), out.data_ptr(), newh, neww, _lib.ptr(xb),
                                       _lib.ptr(xk), kx, _lib.ptr(yb), _lib.ptr(yk), ky, _lib.stream_ptr(img.device))
        _lib.check(rc, "sam_resize_u8")
        return out

    def apply_image(self, image, device=None):
        """HxWxC uint8 -> resized HxWxC uint8 (utils/transforms.py:27-34), computed by the CUDA kernel.  A numpy array
        (the reference's argument type) is uploaded, resized on `device` (default: the current CUDA device) and returned
        as a numpy array; a CUDA tensor stays on its device.  There is no host implementation."""
        if isinstance(image, torch.Tensor):
            return self.apply_image_cuda(image)
        if not torch.cuda.is_available():
            raise RuntimeError("ResizeLongestSide.apply_image: anyref_b200 resizes on the GPU -- there is no CPU fallback")
        dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        t = torch.from_numpy(np.ascontiguousarray(image)).to(dev)
        return self.apply_image_cuda(t).cpu().numpy()

    def apply_coords(self, coords: np.ndarray, original_size: Tuple[int, ...]) -> np.ndarray:
        old_h, old_w = original_size
        new_h, new_w = self.get_preprocess_shape(old_h, old_w, self.target_length)
        out = np.array(coords, dtype=float, copy=True)
        out[..., 0] = out[..., 0] * (new_w / old_w)
        out[..., 1] = out[..., 1] * (new_h / old_h)
        return out

    def apply_boxes(self, boxes: np.ndarray, original_size: Tuple[int, ...]) -> np.ndarray:
        return self.apply_coords(np.asarray(boxes).reshape(-1, 2, 2), original_size).reshape(-1, 4)

    def apply_coords_torch(self, coords: torch.Tensor, original_size: Tuple[int, ...]) -> torch.Tensor:
        old_h, old_w = original_size
        new_h, new_w = self.get_preprocess_shape(old_h, old_w, self.target_length)
        out = coords.clone().to(torch.float)
        out[..., 0] = out[..., 0] * (new_w / old_w)
        out[..., 1] = out[..., 1] * (new_h / old_h)
        return out

    def apply_boxes_torch(self, boxes: torch.Tensor, original_size: Tuple[int, ...]) -> torch.Tensor:
        return self.apply_coords_torch(boxes.reshape(-1, 2, 2), original_size).reshape(-1, 4)
