"""SamPredictor on the B200 kernels (reference: predictor.py:16-285) -- the caller of the point / box / mask prompt
rows: convert_avs_masks.py:29-58 sets an image once and asks for a box-prompted, multimask prediction.

Same public surface and argument meaning as the reference class: `set_image`, `set_torch_image`, `predict`,
`predict_torch`, `get_image_embedding`, `reset_image`, `device`.  Everything on the device side goes through the fused
kernels: Sam.preprocess (normalise + pad + cast), the image encoder, sam_prompt_sparse / sam_prompt_mask_embed, the
batched mask decoder and the fused postprocess (which also produces the thresholded mask in the same pass).
"""
from __future__ import annotations

from typing import Optional, Tuple

import numpy as np
import torch

from .modeling import Sam
from .utils.transforms import ResizeLongestSide


class SamPredictor:
    def __init__(self, sam_model: Sam) -> None:
        self.model = sam_model
        self.transform = ResizeLongestSide(sam_model.image_encoder.img_size)
        self.reset_image()

    # ------------------------------------------------------------------------------------------------ image
    def set_image(self, image: np.ndarray, image_format: str = "RGB") -> None:
        """HWC uint8 image in [0, 255] (predictor.py:33-62)."""
        if image_format not in ("RGB", "BGR"):
            raise AssertionError(f"image_format must be in ['RGB', 'BGR'], is {image_format}.")
        if image_format != self.model.image_format:
            image = image[..., ::-1]
        # upload the raw uint8 image once; resize (PIL-exact), normalise, pad and cast all run on the device
        raw = torch.from_numpy(np.ascontiguousarray(image)).to(self.device)
        resized = self.transform.apply_image_cuda(raw)
        t = resized.permute(2, 0, 1).contiguous()[None]
        self.set_torch_image(t, image.shape[:2])

    @torch.no_grad()
    def set_torch_image(self, transformed_image: torch.Tensor, original_image_size: Tuple[int, ...]) -> None:
        """1x3xHxW image already resized with ResizeLongestSide (predictor.py:64-91); uint8 or float, 0..255."""
        S = self.model.image_encoder.img_size
        if not (transformed_image.dim() == 4 and transformed_image.shape[1] == 3
                and max(*transformed_image.shape[2:]) == S):
            raise AssertionError(f"set_torch_image input must be BCHW with long side {S}.")
        self.reset_image()
        self.original_size = tuple(int(v) for v in original_image_size)
        self.input_size = tuple(int(v) for v in transformed_image.shape[-2:])
        x = self.model.preprocess(transformed_image.to(self.device))
        self.features = self.model.image_encoder(x)
        self.is_image_set = True

    # ------------------------------------------------------------------------------------------------ prompts
    def predict(self, point_coords: Optional[np.ndarray] = None, point_labels: Optional[np.ndarray] = None,
                box: Optional[np.ndarray] = None, mask_input: Optional[np.ndarray] = None,
                multimask_output: bool = True, return_logits: bool = False):
        """numpy front end (predictor.py:93-176): prompts in ORIGINAL image pixels -> (masks CxHxW, iou C, low-res
        logits Cx256x256)."""
        if not self.is_image_set:
            raise RuntimeError("An image must be set with .set_image(...) before mask prediction.")
        coords_t = labels_t = box_t = mask_t = None
        if point_coords is not None:
            if point_labels is None:
                raise AssertionError("point_labels must be supplied if point_coords is supplied.")
            pc = self.transform.apply_coords(point_coords, self.original_size)
            coords_t = torch.as_tensor(pc, dtype=torch.float, device=self.device)[None]
            labels_t = torch.as_tensor(point_labels, dtype=torch.int, device=self.device)[None]
        if box is not None:
            b = self.transform.apply_boxes(box, self.original_size)
            box_t = torch.as_tensor(b, dtype=torch.float, device=self.device)[None, :]   # [1, 1, 4] as the reference
        if mask_input is not None:
            mask_t = torch.as_tensor(mask_input, dtype=torch.float, device=self.device)[None]
        masks, iou, low = self.predict_torch(coords_t, labels_t, box_t, mask_t, multimask_output,
                                             return_logits=return_logits)
        return masks[0].cpu().numpy(), iou[0].cpu().numpy(), low[0].cpu().numpy()

    @torch.no_grad()
    def predict_torch(self, point_coords: Optional[torch.Tensor], point_labels: Optional[torch.Tensor],
                      boxes: Optional[torch.Tensor] = None, mask_input: Optional[torch.Tensor] = None,
                      multimask_output: bool = True, return_logits: bool = False):
        """Batched torch front end (predictor.py:178-257); prompts already in the resized frame."""
        if not self.is_image_set:
            raise RuntimeError("An image must be set with .set_image(...) before mask prediction.")
        points = (point_coords, point_labels) if point_coords is not None else None
        sparse, dense = self.model.prompt_encoder(points=points, boxes=boxes, masks=mask_input, text_embeds=None)
        low, iou = self.model.mask_decoder(image_embeddings=self.features,
                                           image_pe=self.model.prompt_encoder.get_dense_pe(),
                                           sparse_prompt_embeddings=sparse, dense_prompt_embeddings=dense,
                                           multimask_output=multimask_output)
        if return_logits:
            masks = self.model.postprocess_masks(low, self.input_size, self.original_size)
        else:
            _, binary = self.model.postprocess_masks(low, self.input_size, self.original_size, return_binary=True)
            masks = binary.bool()   # logits > mask_threshold, produced by the same kernel pass
        return masks, iou, low

    # ------------------------------------------------------------------------------------------------ state
    def get_image_embedding(self) -> torch.Tensor:
        if not self.is_image_set:
            raise RuntimeError("An image must be set with .set_image(...) to generate an embedding.")
        return self.features

    @property
    def device(self) -> torch.device:
        return self.model.device

    def reset_image(self) -> None:
        self.is_image_set = False
        self.features = None
        self.orig_h = self.orig_w = self.input_h = self.input_w = None
        self.original_size = self.input_size = None
