"""`create_dataset(audio_path, dataset_output_dir)` with the reference's signature and on-disk result
(create_dataset.py:13-64): one float64 (2, 512, 512) `magn_phase_<idx>.pt` per 512-frame chunk, idx counting
through the files in glob order.  The two transform calls of the reference loop are one fused GPU call per file;
under torchrun the FILES are sharded over the ranks (no communication) and every rank writes the idx range the
sequential loop would have used for its files.

`packed=True` (keyword-only extra) writes the packed float32 format of audio/dataset.py instead: one
`shard_<first idx>.f32` per input file + `index.json`, same chunk numbering, half the bytes and no pickle;
`audio.dataset.export_pt` converts it back to the reference's files."""
import glob
from os import mkdir
from os.path import exists, isdir, join

import torch as th
from tqdm import tqdm

from . import audio, parallel
from .audio import dataset, wavio


def _wav_sample_count(path: str) -> int:
    wav, _sr = wavio.load(path)
    return wav.size(1)


def create_dataset(audio_path: str, dataset_output_dir: str, *, packed: bool = False) -> None:
    w_p = glob.glob(audio_path)
    parallel.ensure_dir(dataset_output_dir, f"\"{dataset_output_dir}\" is not a directory")      # every rank of a torchrun launch gets here

    nb_vec = audio.N_VEC
    rank, ws = parallel.world()
    begin, end, idx = 0, len(w_p), 0
    if ws > 1:
        counts = [_wav_sample_count(p) for p in w_p]
        begin, end, idx = parallel.dataset_shard_plan(counts, rank, ws, audio.STFT_STRIDE, nb_vec)

    for wav_p in tqdm(w_p[begin:end]):
        raw_audio, sr = wavio.load(wav_p)
        assert sr == audio.SAMPLE_RATE, \
            f"Audio sample rate must be {audio.SAMPLE_RATE}Hz, " \
            f"file \"{wav_p}\" is {sr}Hz"
        n_frames = 1 + raw_audio.size(1) // audio.STFT_STRIDE
        if n_frames < nb_vec:
            continue
        if n_frames == nb_vec and packed:
            raise ValueError(f"\"{wav_p}\": exactly {nb_vec} frames gives the reference's EMPTY chunk (create_dataset.py:41-64), "
                             "which the packed format cannot hold -- use packed=False for this file")
        if n_frames == nb_vec:
            # reference quirk: the guard passes, the split of an empty tensor yields one EMPTY chunk
            th.save(th.zeros(2, audio.N_FFT // 2, 0, dtype=th.float64), join(dataset_output_dir, f"magn_phase_{idx}.pt"))
            idx += 1
            continue
        magn, phase = audio.wav_to_magn_phase_batch(raw_audio[None].cuda())
        if packed:
            pair = th.stack([magn[0], phase[0]], dim=1).cpu()                   # (n_chunks, 2, 512, 512) float32
            dataset.write_packed_shard(dataset_output_dir, idx, pair)
            idx += pair.size(0)
            continue
        pair = th.stack([magn[0], phase[0]], dim=1).to(th.float64).cpu()        # (n_chunks, 2, 512, 512)
        for s_idx in range(pair.size(0)):
            th.save(pair[s_idx].clone(), join(dataset_output_dir, f"magn_phase_{idx}.pt"))
            idx += 1
    if packed:
        if ws > 1:
            import torch.distributed as dist
            dist.barrier()                 # every rank's shards are on disk
        if rank == 0:
            dataset.write_packed_index(dataset_output_dir)
