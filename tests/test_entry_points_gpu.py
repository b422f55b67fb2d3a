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


def test_train_runs_grows_and_checkpoints(tmp_path):
    import musicgan_b200 as mg
    ds, out = tmp_path / "ds", tmp_path / "run"
    ds.mkdir()
    g = torch.Generator().manual_seed(1)
    for i in range(8):
        torch.save(torch.rand(2, 512, 512, generator=g, dtype=torch.float64) * 2 - 1, str(ds / f"magn_phase_{i}.pt"))
    mg.train("t", str(ds), str(out), batch_size=4, nb_epoch=2, num_workers=0, save_every=3, max_iterations=4, seed=0)
    saved = sorted(os.listdir(out))
    assert "gen_0.pt" in saved and "disc_1.pt" in saved and "optim_gen_0.pt" in saved
    sd = torch.load(str(out / "gen_0.pt"))
    assert "_Generator__gen_blocks.0.0.weight" in sd and all(torch.isfinite(v).all() for v in sd.values())


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
