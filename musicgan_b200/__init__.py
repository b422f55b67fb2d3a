"""musicgan_b200 -- B200 (sm_100a) implementation of the MusicGAN hot path behind the reference's
Python API (`import musicgan_b200 as music_gan`)."""
from . import audio  # noqa: F401
