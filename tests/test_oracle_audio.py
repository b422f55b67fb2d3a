"""CPU: the oracle restatement against the golden vectors generated from the reference
(oracle/gen_golden.py), and the integer frame / chunk / dataset-index plan (bit exact)."""
import os

import numpy as np
import pytest
import torch

from oracle import audio_oracle as ao
from oracle import cases

FAST_FORWARD = ["noise_3s", "exact_513", "tonal_6s", "gated_4s"]


def _load(golden_dir, name):
    return np.load(os.path.join(golden_dir, name))


@pytest.mark.parametrize("name", FAST_FORWARD + ["noise_60s"])
def test_forward_oracle_matches_reference_golden(golden_dir, name):
    g = _load(golden_dir, f"audio_forward_{name}.npz")
    assert bool(g["oracle_bit_exact_in_authoring_container"])
    wav = cases.forward_wav(name)
    assert wav.size(1) == int(g["n_samples"])
    t, head, n_chunks = ao.chunk_plan(wav.size(1))
    assert (t, head, n_chunks) == (int(g["T"]), int(g["head"]), int(g["n_chunks"]))      # integers: bit exact
    cv = ao.stft_c64(wav.mean(0))
    magn, phase = ao.stft_to_phase_magn(cv)
    s = int(g["stride"])
    assert tuple(magn.shape) == (n_chunks, 512, 512)
    # same library FFT, possibly another CPU: allow FFT-rounding-level differences only
    scale = float(np.abs(g["stft_re"]).max())
    np.testing.assert_allclose(cv.real.contiguous().view(-1)[::s].numpy(), g["stft_re"], atol=2e-6 * scale, rtol=0)
    np.testing.assert_allclose(magn.contiguous().view(-1)[::s].numpy(), g["magn"], atol=2e-5, rtol=1e-4)
    ph = phase.contiguous().view(-1)[::s].numpy()
    frac = np.mean(np.abs(ph - g["phase"]) <= 1e-4 + 1e-4 * np.abs(g["phase"]))
    assert frac >= 0.99, frac          # IF is ill conditioned at low |X|; identical CPUs give 1.0


@pytest.mark.parametrize("name", list(cases.INVERSE_CASES)[:3])
def test_inverse_oracle_matches_reference_golden(golden_dir, name):
    g = _load(golden_dir, f"audio_inverse_{name}.npz")
    assert bool(g["oracle_bit_exact_in_authoring_container"])
    wav = ao.magn_phase_to_wav(cases.inverse_input(name))
    assert wav.numel() == int(g["n_out"])
    got = wav[:: int(g["stride"])].double().numpy()
    ref = g["wav"].astype(np.float64)
    snr = 10 * np.log10((ref ** 2).sum() / max(((got - ref) ** 2).sum(), 1e-300))
    assert snr >= 60.0, snr


def test_dataset_index_plan_matches_reference(golden_dir):
    rows = _load(golden_dir, "index_plan.npz")["rows"]
    plan = ao.dataset_index_plan([int(r[0]) for r in rows])
    for r, (first, written) in zip(rows, plan):
        assert ao.n_frames(int(r[0])) == int(r[1])
        assert (first, written) == (int(r[2]), int(r[3]))


def test_unwrap_scalar_semantics():
    # float32 pi constants, remainder = fmodf + sign fix, cumsum accumulates in float64
    x = torch.full((1, 1 << 12), 0.1)
    assert torch.equal(x.cumsum(1)[0, -1], torch.tensor(np.float32(np.float64(np.float32(0.1)) * (1 << 12))))
    phi = torch.tensor([[0.0, 3.0, -3.0, 3.1, -3.1, 0.5]])
    u = ao.unwrap(phi)
    d = (u[:, 1:] - u[:, :-1]).abs()
    assert float(d.max()) <= np.pi + 1e-5
