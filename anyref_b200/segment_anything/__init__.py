"""Drop-in replacement for the reference's `model/segment_anything` package on AnyRef's grounding path
(model/anyref.py:9 imports build_sam_vit_{h,l,b} from it), plus SamPredictor (predictor.py; used by
convert_avs_masks.py) on the same kernels.  SamAutomaticMaskGenerator / ONNX export are outside the path (SURVEY 2 rows
9-10) and are not provided."""
from .build_sam import (build_sam, build_sam_from_config, build_sam_vit_b, build_sam_vit_h, build_sam_vit_l,
                        sam_model_registry)
from .modeling import ImageEncoderViT, MaskDecoder, PromptEncoder, Sam, TwoWayTransformer
from .predictor import SamPredictor

__all__ = ["build_sam", "build_sam_vit_h", "build_sam_vit_l", "build_sam_vit_b", "build_sam_from_config",
           "sam_model_registry", "ImageEncoderViT", "MaskDecoder", "PromptEncoder", "Sam", "TwoWayTransformer", "SamPredictor"]
