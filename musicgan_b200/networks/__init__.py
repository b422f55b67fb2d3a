from .criterion import (
    generator_loss,
    discriminator_loss,
    wasserstein_generator_loss,
    wasserstein_discriminator_loss,
)
from .progan import Generator, Discriminator, PixelNorm
