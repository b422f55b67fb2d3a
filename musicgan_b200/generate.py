"""`generate(output_dir, rand_channels, gen_dict_state, nb_vec, nb_music)` with the reference's signature
(generate.py:12-65): load a generator state dict, run G on a (nb_music, C, 2, 2*nb_vec) latent, turn every clip into
audio.  G and the inverse transform both run on the GPU; under torchrun the clips are sharded over the ranks."""
from os import mkdir
from os.path import exists, isdir, join

import torch as th

from . import audio, parallel
from .audio import wavio
from .networks import Generator


def generate(output_dir: str, rand_channels: int, gen_dict_state: str, nb_vec: int, nb_music: int,
             sub_batch: int = 4) -> None:
    parallel.ensure_dir(output_dir, f"\"{output_dir}\" is not a directory")      # every rank of a torchrun launch gets here

    print("Load model...")
    gen = Generator(rand_channels, end_layer=7)
    gen.load_state_dict(th.load(gen_dict_state, map_location="cpu"))
    gen.eval().cuda()

    rank, ws = parallel.world()
    begin, end = parallel.shard_bounds(nb_music, rank, ws)
    with th.no_grad():
        print("Pass rand data to generator...")
        # the latent of ALL clips is drawn once from the default generator, like the reference; a rank keeps its rows
        z = th.randn(nb_music, rand_channels, 2, 2 * nb_vec)[begin:end]
        print("Saving sound...")
        for lo in range(0, z.size(0), sub_batch):
            gen_sound = gen(z[lo:lo + sub_batch].cuda(), 1.0)                 # (n, 2, 512, 512 * nb_vec)
            wavs = audio.magn_phase_to_wave_batch(gen_sound, imgs_per_clip=1).cpu()      # one clip per image (generate.py:58-65)
            for i in range(wavs.size(0)):
                wavio.save(join(output_dir, f"sound_{begin + lo + i}.wav"), wavs[i][None, :], audio.SAMPLE_RATE)
