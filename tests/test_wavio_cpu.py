"""CPU: audio/wavio.py (the torchcodec-free stand-in for torchaudio.load / save, reference audio/functions.py:43,139)
against an independent RIFF implementation, scipy.io.wavfile, in both directions and for every sample format."""
import numpy as np
import pytest
import torch
from scipy.io import wavfile

from musicgan_b200.audio import wavio


def test_files_written_by_wavio_are_read_identically_by_scipy(tmp_path):
    g = torch.Generator().manual_seed(0)
    wav = torch.rand(2, 4411, generator=g) * 2 - 1
    p = str(tmp_path / "a.wav")
    wavio.save(p, wav, 44100)
    sr, data = wavfile.read(p)
    assert sr == 44100 and data.dtype == np.float32 and data.shape == (4411, 2)
    assert np.array_equal(data.T, wav.numpy())
    back, sr2 = wavio.load(p)
    assert sr2 == 44100 and torch.equal(back, wav)


@pytest.mark.parametrize("dtype,scale", [(np.int16, 32768.0), (np.int32, 2147483648.0), (np.uint8, None), (np.float32, 1.0), (np.float64, 1.0)])
@pytest.mark.parametrize("channels", [1, 2])
def test_files_written_by_scipy_are_decoded_like_torchaudio(tmp_path, dtype, scale, channels):
    """torchaudio.load(normalize=True) semantics: integer PCM / 2^(bits-1), 8-bit unsigned (x - 128) / 128, floats as they are."""
    rng = np.random.default_rng(1)
    n = 1000
    if dtype == np.uint8:
        data = rng.integers(0, 256, size=(n, channels), dtype=np.uint8)
        want = (data.astype(np.float32) - 128.0) / 128.0
    elif np.issubdtype(dtype, np.integer):
        info = np.iinfo(dtype)
        data = rng.integers(info.min, info.max, size=(n, channels), dtype=dtype)
        want = (data.astype(np.float64) / scale).astype(np.float32)
    else:
        data = (rng.random((n, channels)) * 2 - 1).astype(dtype)
        want = data.astype(np.float32)
    p = str(tmp_path / "b.wav")
    wavfile.write(p, 22050, data if channels > 1 else data[:, 0])
    wav, sr = wavio.load(p)
    assert sr == 22050 and tuple(wav.shape) == (channels, n) and wav.dtype == torch.float32
    assert np.array_equal(wav.numpy(), want.T)


def test_rejects_non_wave(tmp_path):
    p = tmp_path / "c.wav"
    p.write_bytes(b"not a wave file at all")
    with pytest.raises(ValueError):
        wavio.load(str(p))
