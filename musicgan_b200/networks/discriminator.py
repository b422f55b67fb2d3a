"""`music_gan.networks.discriminator` import path (reference networks/discriminator.py): ConvBlock :8-34,
MagPhaseLayer :37-50, Discriminator :53-191 -- implemented in progan.py on the tcgen05 kernels."""
from .progan import ConvBlock, Discriminator, MagPhaseLayer

__all__ = ["ConvBlock", "MagPhaseLayer", "Discriminator"]
