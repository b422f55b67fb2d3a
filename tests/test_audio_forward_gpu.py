"""GPU parity of the forward transform (through the C ABI) against the CPU oracle and the
golden vectors of the reference.  Protocol = SURVEY Appendix B.4:
  stage 1  FFT:        complex STFT vs oracle, error relative to the spectrum scale
  stage 2  post-FFT:   oracle's own STFT fed to mg_phase_magn_from_stft -> magn allclose, IF >= 99.9 %
                       within rtol 1e-4 (+ atol 1e-6), outliers must be 1-ulp-of-unwrapped-phase or wrap flips
  stage 3  end to end: fused mg_stft_magif_f32 vs oracle; integers bit exact; magn rtol 1e-4; IF
                       fraction within tolerance reported and bounded (ill-conditioned at low |X|)
"""
import os

import numpy as np
import pytest
import torch

from oracle import audio_oracle as ao
from oracle import cases

pytestmark = pytest.mark.gpu

CASES = ["noise_3s", "exact_513", "tonal_6s", "gated_4s"]


def _frac_within(got, ref, rtol, atol):
    return float((np.abs(got - ref) <= atol + rtol * np.abs(ref)).mean())


@pytest.mark.parametrize("name", CASES)
def test_stage1_stft(name):
    from musicgan_b200 import audio
    wav = cases.forward_wav(name)
    ref = ao.stft_c64(wav.mean(0))
    got = audio.stft_from_wave(wav.cuda()).cpu()
    assert got.shape == ref.shape and got.dtype == torch.complex64
    assert got.stride(0) == 1 and ref.stride(0) == 1     # frame-major memory like the reference
    scale = ref.abs().max().item()
    err = (got - ref).abs().max().item()
    assert err <= 1e-4 * scale, (err, scale)
    assert err <= 3e-6 * scale, (err, scale)     # fp32 FFT rounding level


@pytest.mark.parametrize("name", CASES)
def test_stage2_post_fft(name):
    from musicgan_b200 import audio
    wav = cases.forward_wav(name)
    cv = ao.stft_c64(wav.mean(0))
    ref_m, ref_p = ao.stft_to_phase_magn(cv)
    got_m, got_p = audio.stft_to_phase_magn(cv)              # CPU in -> CPU out, like the reference
    assert got_m.shape == ref_m.shape and got_p.shape == ref_p.shape
    np.testing.assert_allclose(got_m.numpy(), ref_m.numpy(), rtol=1e-4, atol=1e-6)
    g, r = got_p.numpy(), ref_p.numpy()
    frac = _frac_within(g, r, 1e-4, 1e-6)
    assert frac >= 0.999, frac
    # every outlier: a few ulp of the unwrapped phase (<= ~4e-3 rad of a ~2pi range -> <= 2e-3 normalised)
    # or a +-2pi wrap flip (normalised difference ~ 2 * 2pi / range ~ 2)
    bad = np.abs(g - r) > 1e-6 + 1e-4 * np.abs(r)
    if bad.any():
        d = np.abs(g - r)[bad]
        assert np.all((d < 4e-3) | (d > 0.9)), d.max()


@pytest.mark.parametrize("name", CASES)
def test_stage3_end_to_end(name):
    from musicgan_b200 import audio, _lib
    wav = cases.forward_wav(name)
    ref_m, ref_p = ao.wav_to_magn_phase(wav.mean(0))
    t, head, n_chunks = _lib.chunk_plan(wav.size(1))
    assert (t, head, n_chunks) == ao.chunk_plan(wav.size(1))
    got_m, got_p = audio.wav_to_magn_phase_batch(wav[None].cuda())
    assert tuple(got_m.shape) == (1, n_chunks, 512, 512)
    gm, gp = got_m[0].cpu().numpy(), got_p[0].cpu().numpy()
    np.testing.assert_allclose(gm, ref_m.numpy(), rtol=1e-4, atol=2e-5)
    frac = _frac_within(gp, ref_p.numpy(), 1e-4, 1e-4)
    print(f"{name}: IF within rtol1e-4+atol1e-4: {frac:.5f}; within 1e-3: {_frac_within(gp, ref_p.numpy(), 0, 1e-3):.5f}")
    assert frac >= 0.98, frac
    assert _frac_within(gp, ref_p.numpy(), 0, 1e-3) >= 0.995


def test_if_threshold_is_the_references_own_cpu_vs_cuda_discrepancy():
    """the end-to-end IF agreement required of the CUDA path is 'no worse than reference-CPU vs
    reference-through-torch-CUDA on the same box'.  The reference's transform restated with stock torch ops (cuFFT, ATen,
    cumsum in double: baseline/torch_b200.py) is run on this GPU and compared with the CPU oracle; our kernels must agree
    with the oracle at least as well (minus 0.2 % slack for run-to-run wrap flips).  This is where the fixed 0.98 of the
    other end-to-end tests comes from (measured 0.99+ for both on B200)."""
    from musicgan_b200 import audio
    from baseline import torch_b200 as tb
    for name in ("noise_3s", "tonal_6s"):
        wav = cases.forward_wav(name)
        mono = wav.mean(0)
        ref_m, ref_p = ao.wav_to_magn_phase(mono)
        t_m, t_p = tb.transform(mono[None].cuda())
        o_m, o_p = audio.wav_to_magn_phase_batch(wav[None].cuda())
        f_torch = _frac_within(t_p[0].cpu().numpy(), ref_p.numpy(), 1e-4, 1e-4)
        f_ours = _frac_within(o_p[0].cpu().numpy(), ref_p.numpy(), 1e-4, 1e-4)
        print(f"{name}: IF within rtol 1e-4 + atol 1e-4 of the CPU reference: stock torch-CUDA {f_torch:.5f}, this library {f_ours:.5f}")
        assert f_ours >= min(f_torch, 0.999) - 2e-3, (f_ours, f_torch)
        assert f_ours >= 0.98


def test_batch_equals_single():
    """Clips of a batch are independent units: every clip of a batch == the clip run alone (bit exact)."""
    from musicgan_b200 import audio
    wavs = cases.batch_wavs(3, 140_000, seed=3).cuda()
    bm, bp = audio.wav_to_magn_phase_batch(wavs)
    bm, bp = bm.clone(), bp.clone()
    for i in range(3):
        m, p = audio.wav_to_magn_phase_batch(wavs[i:i + 1].contiguous())
        assert torch.equal(m[0], bm[i]) and torch.equal(p[0], bp[i])


def test_outputs_span_unit_range_and_stereo_mean():
    from musicgan_b200 import audio
    wav = cases.forward_wav("tonal_6s")            # stereo
    m, p = audio.wav_to_magn_phase_batch(wav[None].cuda())
    for x in (m, p):
        assert x.min().item() == -1.0 and x.max().item() <= 1.0
    m2, p2 = audio.wav_to_magn_phase_batch(wav.mean(0)[None, None].cuda())
    assert torch.equal(m, m2) and torch.equal(p, p2)


@pytest.mark.parametrize("name", ["noise_60s", "tonal_60s"])
def test_config1_full_size_against_golden(golden_dir, name):
    """BASELINE config 1 (60 s, T = 10336, 20 chunks, 95 head frames) against the reference's
    golden subsample -- no oracle run needed at this size."""
    from musicgan_b200 import audio
    g = np.load(os.path.join(golden_dir, f"audio_forward_{name}.npz"))
    wav = cases.forward_wav(name)
    m, p = audio.wav_to_magn_phase_batch(wav[None].cuda())
    assert tuple(m.shape) == (1, int(g["n_chunks"]), 512, 512) and int(g["head"]) == 95 and int(g["T"]) == 10336
    s = int(g["stride"])
    gm = m[0].contiguous().view(-1)[::s].cpu().numpy()
    gp = p[0].contiguous().view(-1)[::s].cpu().numpy()
    np.testing.assert_allclose(gm, g["magn"], rtol=1e-4, atol=2e-5)
    frac = _frac_within(gp, g["phase"], 1e-4, 1e-4)
    print(f"{name}: IF within tol {frac:.5f}")
    assert frac >= 0.98
    # size-independent property: normalised outputs hit both ends of [-1, 1] (min over kept+dropped head may
    # sit in the head, so only the upper/lower bounds are checked)
    assert m.max().item() <= 1.0 and m.min().item() >= -1.0
