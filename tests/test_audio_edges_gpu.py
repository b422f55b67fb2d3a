"""GPU: edge cases of the audio path the reference's behaviour defines: ragged lengths, the T == 512
empty-chunk quirk, the largest head (511 dropped frames), assertion messages of the drop-in functions."""
import numpy as np
import pytest
import torch

from oracle import audio_oracle as ao

pytestmark = pytest.mark.gpu


def _noise(n, seed):
    return (torch.rand(1, n, generator=torch.Generator().manual_seed(seed)) * 2 - 1) * 0.5


@pytest.mark.parametrize("n", [131_072 + 1, 131_072 + 255, 131_072 + 256, 262_143, 2 * 131_072 - 256 + 511 * 256])
def test_ragged_lengths_integer_plan_and_values(n):
    from musicgan_b200 import audio, _lib
    wav = _noise(n, n % 1000)
    ref_m, ref_p = ao.wav_to_magn_phase(wav[0])
    t, head, n_chunks = _lib.chunk_plan(n)
    assert (t, head, n_chunks) == ao.chunk_plan(n) and ref_m.size(0) == n_chunks
    m, p = audio.wav_to_magn_phase_batch(wav[None].cuda())
    assert tuple(m.shape) == (1, n_chunks, 512, 512)
    np.testing.assert_allclose(m[0].cpu().numpy(), ref_m.numpy(), rtol=1e-4, atol=2e-5)
    frac = (np.abs(p[0].cpu().numpy() - ref_p.numpy()) <= 1e-4 + 1e-4 * np.abs(ref_p.numpy())).mean()
    assert frac >= 0.98


def test_t_equal_512_gives_one_empty_chunk_like_the_reference():
    from musicgan_b200 import audio
    wav = _noise(511 * 256 + 17, 3)                 # T = 512: passes create_dataset's guard, yields an EMPTY chunk
    cv = ao.stft_c64(wav[0])
    assert cv.size(1) == 512
    ref_m, ref_p = ao.stft_to_phase_magn(cv)
    m, p = audio.stft_to_phase_magn(cv)
    assert tuple(m.shape) == tuple(ref_m.shape) == (1, 512, 0) and tuple(p.shape) == tuple(ref_p.shape)


def test_short_clip_below_one_chunk_still_runs_the_kernels():
    from musicgan_b200 import audio
    wav = _noise(40_000, 4)                         # T = 157 < 512: no chunk, min/max still defined
    plan = audio.ForwardPlan(40_000, 1, 1)
    m, p = plan.run(wav[None].cuda().contiguous())
    assert tuple(m.shape) == (1, 0, 512, 512)
    raw_m, raw_p = ao.phase_magn_raw(ao.stft_c64(wav[0]))
    torch.cuda.synchronize()


def test_minmax_output_matches_reference_extrema():
    from musicgan_b200 import audio
    wav = _noise(200_000, 5)
    plan = audio.ForwardPlan(200_000, 1, 1)
    plan.run(wav[None].cuda().contiguous())
    raw_m, raw_p = ao.phase_magn_raw(ao.stft_c64(wav[0]))
    mm = plan.minmax[0].cpu().numpy()
    np.testing.assert_allclose(mm[:2], [raw_m.min().item(), raw_m.max().item()], rtol=1e-5, atol=1e-6)   # the min is a near-zero bin
    np.testing.assert_allclose(mm[2:], [raw_p.min().item(), raw_p.max().item()], rtol=1e-3, atol=1e-3)


def test_assertion_messages_are_the_references(tmp_path):
    from musicgan_b200 import audio
    from musicgan_b200.audio import wavio
    with pytest.raises(AssertionError, match=r"\(N, 2, H, W\), actual"):
        audio.magn_phase_to_wav(torch.zeros(2, 512, 16), "x.wav", 44100)
    with pytest.raises(AssertionError, match="Channels must be equal to 2"):
        audio.magn_phase_to_wav(torch.zeros(1, 3, 512, 16), "x.wav", 44100)
    with pytest.raises(AssertionError, match="Frequency size must be equal to 512"):
        audio.magn_phase_to_wav(torch.zeros(1, 2, 256, 16), "x.wav", 44100)
    with pytest.raises(AssertionError, match=r"\(STFT, TIME\), actual"):
        audio.bark_magn_scale(torch.zeros(4, 4, 4))
    p = str(tmp_path / "w.wav")
    wavio.save(p, torch.zeros(1, 4000), 22050)
    with pytest.raises(AssertionError, match="Audio sample rate must be 44100Hz"):
        audio.wav_to_stft(p)
    with pytest.raises(NotImplementedError):
        audio.stft_from_wave(torch.zeros(1, 4000), nperseg=2048)


def test_helper_functions_match_reference_semantics():
    from musicgan_b200.audio import functions as fn
    g = torch.Generator().manual_seed(9)
    phi = (torch.rand(16, 300, generator=g) * 2 - 1) * 3.14159
    assert torch.equal(fn.diff(phi), ao.diff(phi))
    assert torch.equal(fn.unwrap(phi), ao.unwrap(phi))
    assert torch.equal(fn.unwrap(phi.cuda()).cpu(), ao.unwrap(phi))
    m = torch.rand(512, 40, generator=g)
    assert torch.equal(fn.bark_magn_scale(m), m * ao.bark_gain(512))
    assert torch.equal(fn.bark_magn_scale(m, unscale=True), m / ao.bark_gain(512))
