from .transforms import ResizeLongestSide

__all__ = ["ResizeLongestSide"]
