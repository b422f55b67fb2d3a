"""`music_gan.networks.generator` import path (reference networks/generator.py): Block :9-40, ToMagnPhaseLayer :43-52,
Generator :55-171 -- implemented in progan.py on the tcgen05 kernels."""
from .progan import Block, Generator, ToMagnPhaseLayer

__all__ = ["Block", "ToMagnPhaseLayer", "Generator"]
