"""`music_gan.networks.layers` import path (reference networks/layers.py:5-23)."""
from .progan import PixelNorm

__all__ = ["PixelNorm"]
