"""GPU parity of the inverse transform (C ABI mg_istft_from_magif_f32) against the CPU oracle and the
reference's golden vectors: reconstructed audio SNR >= 60 dB (BASELINE north_star)."""
import os

import numpy as np
import pytest
import torch

from oracle import audio_oracle as ao
from oracle import cases

pytestmark = pytest.mark.gpu


def snr_db(got, ref):
    got, ref = np.asarray(got, dtype=np.float64), np.asarray(ref, dtype=np.float64)
    return 10 * np.log10((ref ** 2).sum() / max(((got - ref) ** 2).sum(), 1e-300))


@pytest.mark.parametrize("name", ["rand_w512", "batch2_w512", "coherent_w2048"])
def test_inverse_vs_oracle(name):
    from musicgan_b200 import audio
    mp = cases.inverse_input(name)
    ref = ao.magn_phase_to_wav(mp.clone())
    got = audio.magn_phase_to_wave_batch(mp.cuda(), imgs_per_clip=mp.size(0))
    assert tuple(got.shape) == (1, ref.numel())            # 256 * (N*W - 1): integer, bit exact
    s = snr_db(got[0].cpu().numpy(), ref.numpy())
    print(f"{name}: SNR {s:.1f} dB")
    assert s >= 60.0, s


@pytest.mark.parametrize("name", list(cases.INVERSE_CASES))
def test_inverse_vs_reference_golden(golden_dir, name):
    from musicgan_b200 import audio
    g = np.load(os.path.join(golden_dir, f"audio_inverse_{name}.npz"))
    mp = cases.inverse_input(name)
    got = audio.magn_phase_to_wave_batch(mp.cuda(), imgs_per_clip=mp.size(0))[0]
    assert got.numel() == int(g["n_out"])
    s = snr_db(got[:: int(g["stride"])].cpu().numpy(), g["wav"])
    print(f"{name}: SNR vs reference golden {s:.1f} dB")
    assert s >= 60.0, s


def test_clips_are_independent_units():
    """generate.py:58-65 calls the inverse once per clip: a batched call must equal per-clip calls bit for bit."""
    from musicgan_b200 import audio
    g = torch.Generator().manual_seed(3)
    x = (torch.rand(3, 2, 512, 96, generator=g) * 2 - 1).cuda()
    both = audio.magn_phase_to_wave_batch(x, imgs_per_clip=1)
    for i in range(3):
        one = audio.magn_phase_to_wave_batch(x[i:i + 1].contiguous(), imgs_per_clip=1)
        assert torch.equal(one[0], both[i])


def test_forward_inverse_round_trip():
    """Size-independent property: a (magn, IF) image produced by the forward transform, put back through the
    inverse, must give audio whose forward transform reproduces the magnitude image (the IF image is defined
    up to the unwrap constant, magnitude is not)."""
    from musicgan_b200 import audio
    wav = cases.forward_wav("noise_3s")
    m, p = audio.wav_to_magn_phase_batch(wav[None].cuda())
    img = torch.stack([m[0, 0], p[0, 0]], 0)[None]
    rec = audio.magn_phase_to_wave_batch(img, 1)
    assert rec.shape[1] == 256 * 511 and torch.isfinite(rec).all()
    assert rec.abs().max().item() > 0


def test_wav_file_drop_in(tmp_path):
    """magn_phase_to_wav writes a float32 WAVE that our reader (and the forward path) reads back."""
    from musicgan_b200 import audio
    from musicgan_b200.audio import wavio
    mp = cases.inverse_input("rand_w512")
    path = str(tmp_path / "o.wav")
    audio.magn_phase_to_wav(mp, path, 44100)
    w, sr = wavio.load(path)
    ref = ao.magn_phase_to_wav(mp.clone())
    assert sr == 44100 and tuple(w.shape) == (1, ref.numel())
    assert snr_db(w[0].numpy(), ref.numpy()) >= 60.0
    cv = audio.wav_to_stft(path)
    assert tuple(cv.shape) == (512, 1 + ref.numel() // 256) and not cv.is_cuda
