"""Host-side geometry helpers of the predictor (reference: utils/transforms.py:17-113).

The image resize itself is CPU data preparation in AnyRef (PIL bilinear, utils/transforms.py:27-34) and stays on the
host; prompts only need the coordinate maps below, which are a few scalar multiplications."""
from __future__ import annotations

from typing import Tuple

import numpy as np
import torch


class ResizeLongestSide:
    """Scale so that the longest side equals `target_length`; coordinates and boxes follow the same map."""

    def __init__(self, target_length: int) -> None:
        self.target_length = target_length

    @staticmethod
    def get_preprocess_shape(oldh: int, oldw: int, long_side_length: int) -> Tuple[int, int]:
        """utils/transforms.py:102-113: round-half-up of the scaled sides."""
        scale = long_side_length * 1.0 / max(oldh, oldw)
        return int(oldh * scale + 0.5), int(oldw * scale + 0.5)

    def apply_image(self, image: np.ndarray) -> np.ndarray:
        """HxWxC uint8 -> resized uint8 (PIL bilinear, as torchvision's resize of a PIL image does)."""
        from PIL import Image

        newh, neww = self.get_preprocess_shape(image.shape[0], image.shape[1], self.target_length)
        return np.array(Image.fromarray(image).resize((neww, newh), Image.BILINEAR))

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
