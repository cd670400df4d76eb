from .image_encoder import ImageEncoderViT
from .mask_decoder import MaskDecoder
from .prompt_encoder import PromptEncoder
from .sam import Sam
from .transformer import TwoWayTransformer

__all__ = ["ImageEncoderViT", "MaskDecoder", "PromptEncoder", "Sam", "TwoWayTransformer"]
