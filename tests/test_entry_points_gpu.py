"""GPU: the drop-in entry points create_dataset / generate / train (reference create_dataset.py, generate.py,
train.py) end to end on small synthetic inputs."""
import os

import numpy as np
import pytest
import torch

from oracle import audio_oracle as ao
from oracle import cases

pytestmark = pytest.mark.gpu


def test_create_dataset_matches_reference_numbering_and_content(tmp_path):
    import musicgan_b200 as mg
    from musicgan_b200.audio import wavio
    src, out = tmp_path / "wav", tmp_path / "ds"
    src.mkdir()
    lens = {"a": 140_000, "b": 100_000, "c": 270_001}      # b is too short (T < 512): skipped, consumes no idx
    for name, n in lens.items():
        g = torch.Generator().manual_seed(n)
        wavio.save(str(src / f"{name}.wav"), (torch.rand(2 if name == "c" else 1, n, generator=g) - 0.5), 44100)
    mg.create_dataset(str(src / "*.wav"), str(out))
    import glob
    order = [os.path.basename(p)[0] for p in glob.glob(str(src / "*.wav"))]
    expected = sum(ao.chunk_plan(lens[k])[2] for k in order if ao.chunk_plan(lens[k])[0] >= 512)
    files = sorted(os.listdir(out))
    assert files == sorted(f"magn_phase_{i}.pt" for i in range(expected))
    idx = 0
    for k in order:
        if ao.chunk_plan(lens[k])[0] < 512:
            continue
        wav, _ = wavio.load(str(src / f"{k}.wav"))
        ref_m, ref_p = ao.wav_to_magn_phase(wav.mean(0))
        for c in range(ref_m.size(0)):
            t = torch.load(str(out / f"magn_phase_{idx}.pt"))
            assert t.dtype == torch.float64 and tuple(t.shape) == (2, 512, 512)
            np.testing.assert_allclose(t[0].numpy(), ref_m[c].double().numpy(), rtol=1e-4, atol=2e-5)
            frac = (np.abs(t[1].numpy() - ref_p[c].double().numpy()) <= 1e-4 + 1e-4 * np.abs(ref_p[c].numpy())).mean()
            assert frac >= 0.98
            idx += 1
    with pytest.raises(NotADirectoryError):
        mg.create_dataset(str(src / "*.wav"), str(src / "a.wav"))


def test_create_dataset_packed_equals_reference_format_and_trains(tmp_path):
    """`packed=True` writes one float32 shard per input file + index.json with the reference's chunk numbering;
    the values are the float32 originals of the reference-format float64 files, and `train` reads it like any dataset."""
    import musicgan_b200 as mg
    from musicgan_b200.audio import wavio
    from musicgan_b200.audio.dataset import AudioDataset
    src, ref_out, packed_out = tmp_path / "wav", tmp_path / "ds_pt", tmp_path / "ds_packed"
    src.mkdir()
    for name, n in {"a": 270_001, "b": 100_000, "c": 140_000}.items():      # b is too short: skipped
        g = torch.Generator().manual_seed(n)
        wavio.save(str(src / f"{name}.wav"), torch.rand(1, n, generator=g) - 0.5, 44100)
    mg.create_dataset(str(src / "*.wav"), str(ref_out))
    mg.create_dataset(str(src / "*.wav"), str(packed_out), packed=True)
    legacy, packed = AudioDataset(str(ref_out)), AudioDataset(str(packed_out))
    assert len(legacy) == len(packed) == 3 and "index.json" in os.listdir(packed_out)
    for i in range(3):           # < 10 chunks: the reference's lexicographic file order is the numeric order
        assert torch.equal(packed[i].double(), legacy[i])
    mg.train("t", str(packed_out), str(tmp_path / "run"), batch_size=2, nb_epoch=2, num_workers=2, save_every=2,
             max_iterations=2, seed=0)
    assert "gen_0.pt" in os.listdir(tmp_path / "run")


def test_generate_writes_clips_of_the_reference_length(tmp_path):
    import musicgan_b200 as mg
    from musicgan_b200 import networks
    from musicgan_b200.audio import wavio
    torch.manual_seed(0)
    gen = networks.Generator(32, end_layer=7)
    ckpt = str(tmp_path / "gen.pt")
    torch.save(gen.state_dict(), ckpt)
    out = tmp_path / "sounds"
    mg.generate(str(out), 32, ckpt, nb_vec=2, nb_music=3)
    for i in range(3):
        w, sr = wavio.load(str(out / f"sound_{i}.wav"))
        assert sr == 44100 and tuple(w.shape) == (1, 256 * (512 * 2 - 1)) and torch.isfinite(w).all()


@pytest.mark.parametrize("cuda_graphs", [True, False])
def test_train_runs_grows_and_checkpoints(tmp_path, cuda_graphs):
    """`train` on the CUDA-graph path (the object bench.py times) through TWO growth events -- re-capture at each
    next_layer(), alpha ramping as a device scalar -- and on the eager path; checkpoints follow the reference's schedule
    (utils.py:209-233: first files at call #save_every)."""
    import musicgan_b200 as mg
    ds, out = tmp_path / "ds", tmp_path / "run"
    ds.mkdir()
    g = torch.Generator().manual_seed(1)
    for i in range(8):
        torch.save(torch.rand(2, 512, 512, generator=g, dtype=torch.float64) * 2 - 1, str(ds / f"magn_phase_{i}.pt"))
    mg.train("t", str(ds), str(out), batch_size=4, nb_epoch=8, num_workers=0, save_every=3, max_iterations=9, seed=0,
             cuda_graphs=cuda_graphs, log_every=2, train_lengths=(8, 12, 150000, 200000, 250000, 300000, 350000),
             fadein_lengths=(1, 16, 16, 50000, 62500, 75000, 87500, 100000))
    saved = sorted(os.listdir(out))
    assert {"gen_0.pt", "disc_1.pt", "optim_gen_2.pt"} <= set(saved) and "gen_3.pt" not in saved
    sd0, sd2 = torch.load(str(out / "gen_0.pt")), torch.load(str(out / "gen_2.pt"))
    assert "_Generator__gen_blocks.0.0.weight" in sd0 and all(torch.isfinite(v).all() for v in sd2.values())
    # grown twice by iteration 9 (12 and 24 samples seen): the end block now maps the 96 channels of block 2
    assert tuple(sd2["_Generator__end_block.0.weight"].shape) == (2, 96, 1, 1)
    assert not torch.equal(sd0["_Generator__gen_blocks.0.0.weight"], sd2["_Generator__gen_blocks.0.0.weight"])      # it trains


def test_device_prefetcher_order_and_content():
    """utils.DevicePrefetcher (the loader wrapper of `train` and of bench.py's end-to-end loop): batches arrive on the
    device in order and intact although their uploads run one or two steps ahead on a copy stream."""
    import torch
    from musicgan_b200.utils import DevicePrefetcher
    g = torch.Generator().manual_seed(0)
    host = [torch.randn(4, 2, 64, 64, generator=g).pin_memory() for _ in range(7)]
    got = []
    for x in DevicePrefetcher(iter(host), "cuda:0", depth=2):
        assert x.is_cuda
        got.append((x * 2.0).cpu())          # consume on the current stream
    assert len(got) == len(host)
    for a, b in zip(got, host):
        assert torch.equal(a, b * 2.0)
