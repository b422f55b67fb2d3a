"""musicgan_b200 -- B200 (sm_100a) implementation of the MusicGAN hot path behind the reference's
Python API (`import musicgan_b200 as music_gan`): audio transform, ProGAN networks, and the three entry points."""
from . import audio  # noqa: F401
from . import networks  # noqa: F401
from .create_dataset import create_dataset  # noqa: F401
from .generate import generate  # noqa: F401
from .train import train  # noqa: F401
