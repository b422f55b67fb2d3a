"""Seeded synthetic inputs shared by the golden generator, the tests, smoke() and bench.py
(TEST INFRASTRUCTURE; torch CPU RNG is deterministic across machines)."""
import math

import torch

SR = 44100

# name -> spec.  `stride` = subsampling stride of the stored golden outputs.
FORWARD_CASES = {
    "noise_3s":   dict(n=140_000, kind="noise", seed=1234, channels=1, stride=13),
    "exact_513":  dict(n=131_072, kind="noise", seed=5, channels=1, stride=29),
    "tonal_6s":   dict(n=270_001, kind="tonal", seed=0, channels=2, stride=31),
    "gated_4s":   dict(n=176_400, kind="gated", seed=11, channels=1, stride=17),
    "noise_60s":  dict(n=2_646_000, kind="noise", seed=1234, channels=1, stride=997),   # BASELINE config 1
    "tonal_60s":  dict(n=2_646_000, kind="tonal", seed=0, channels=1, stride=1009),
}

INVERSE_CASES = {
    "rand_w512":      dict(n=1, w=512, kind="rand", seed=7, stride=7),
    "batch2_w512":    dict(n=2, w=512, kind="rand", seed=8, stride=11),
    "coherent_w2048": dict(n=1, w=2048, kind="coherent", seed=9, stride=23),
    "coherent_w5120": dict(n=1, w=5120, kind="coherent", seed=10, stride=101),
}

# create_dataset idx bookkeeping: T<512 skip, T==512 empty chunk, T==513 one chunk, ragged.
INDEX_PLAN_SAMPLE_COUNTS = [100_000, 511 * 256, 130_900, 512 * 256, 512 * 256 + 255, 300_000, 1024 * 256 + 7]


def forward_wav(name: str) -> torch.Tensor:
    """(C, N) fp32 waveform in [-1, 1] for a FORWARD_CASES entry."""
    c = FORWARD_CASES[name]
    g = torch.Generator().manual_seed(c["seed"])
    n, ch = c["n"], c["channels"]
    if c["kind"] == "noise":
        return (torch.rand(ch, n, generator=g) * 2 - 1) * 0.5
    t = torch.arange(n, dtype=torch.float64) / SR
    if c["kind"] == "tonal":
        base = (0.3 * torch.sin(2 * math.pi * 440 * t) + 0.3 * torch.sin(2 * math.pi * 9000 * t)
                + 0.2 * torch.sin(2 * math.pi * 18000 * t)).float()
        wav = base[None, :].repeat(ch, 1) + 0.01 * torch.randn(ch, n, generator=g)
        if ch == 2:
            wav[1] = wav[1] * 0.5
        return wav
    if c["kind"] == "gated":   # noise bursts separated by exact digital silence (atan2(0,0), magn 0)
        wav = (torch.rand(ch, n, generator=g) * 2 - 1) * 0.25
        gate = ((torch.arange(n) // 22050) % 2 == 0).float()
        return wav * gate
    raise KeyError(c["kind"])


def inverse_input(name: str) -> torch.Tensor:
    """(N, 2, 512, W) fp32 in [-1, 1] for an INVERSE_CASES entry."""
    c = INVERSE_CASES[name]
    g = torch.Generator().manual_seed(c["seed"])
    x = torch.rand(c["n"], 2, 512, c["w"], generator=g) * 2 - 1
    if c["kind"] == "coherent":   # phase channel ~0.9 +- 0.05: coherent phase growth
        x[:, 1] = 0.9 + 0.05 * x[:, 1]
    return x


def batch_wavs(batch: int, n: int, seed: int = 2024) -> torch.Tensor:
    """(batch, n) mono noise clips for throughput runs (each clip differently seeded)."""
    g = torch.Generator().manual_seed(seed)
    return (torch.rand(batch, n, generator=g) * 2 - 1) * 0.5
