"""CPU restatement of ``music_gan.audio.functions`` (TEST INFRASTRUCTURE ONLY).

Every function follows the reference op by op, on torch CPU tensors in fp32, so
that the float32 rounding sequence of the reference is reproduced (python-float
scalars are rounded to fp32 by torch's scalar promotion, ``%`` is
``torch.remainder`` == fmodf + sign fix, CPU ``cumsum`` accumulates in fp64 and
rounds every output to fp32 -- SURVEY Appendix B).

Pinned against the imported reference by ``oracle/gen_golden.py`` (bit-exact in
the authoring container) -> fixtures in ``tests/golden/audio_*.npz``.

Citations are into /root/reference/music_gan/audio/functions.py.
"""
from __future__ import annotations

import math
from typing import Tuple

import torch

N_FFT = 1024          # constant.py:1
N_VEC = 512           # constant.py:2
STFT_STRIDE = 256     # constant.py:3
SAMPLE_RATE = 44100   # constant.py:4


# ----------------------------------------------------------------------------
# integer / index arithmetic (must be bit exact)
# ----------------------------------------------------------------------------
def n_frames(n_samples: int, hop: int = STFT_STRIDE) -> int:
    """Centred STFT frame count: T = 1 + N // hop (functions.py:53-59 via torch.stft)."""
    return 1 + n_samples // hop


def chunk_plan(n_samples: int, nb_vec: int = N_VEC, hop: int = STFT_STRIDE) -> Tuple[int, int, int]:
    """(T, head, n_chunks) of functions.py:76-92.

    After the time difference there are T-1 columns; the leading ``(T-1) % nb_vec``
    are dropped (:89-90) and the rest split in ``nb_vec`` wide chunks (:91-92).
    """
    t = n_frames(n_samples, hop)
    cols = t - 1
    head = cols % nb_vec
    return t, head, (cols - head) // nb_vec


def dataset_index_plan(sample_counts, nb_vec: int = N_VEC, hop: int = STFT_STRIDE):
    """File -> first ``magn_phase_<idx>.pt`` index, create_dataset.py:32-64.

    A file with T < nb_vec is skipped (:41); a file with T == nb_vec passes the guard
    but ``split`` of an empty tensor still yields ONE empty chunk, so it consumes an idx.
    Returns a list of (first_idx, n_written) per file.
    """
    out, idx = [], 0
    for n in sample_counts:
        t, _head, n_chunks = chunk_plan(n, nb_vec, hop)
        if t < nb_vec:
            out.append((idx, 0))
            continue
        written = n_chunks if (t - 1) - (t - 1) % nb_vec > 0 else 1
        out.append((idx, written))
        idx += written
    return out


# ----------------------------------------------------------------------------
# forward transform
# ----------------------------------------------------------------------------
def stft_c64(mono: torch.Tensor, n_fft: int = N_FFT, hop: int = STFT_STRIDE) -> torch.Tensor:
    """functions.py:49-62: periodic Hann, centred reflect-padded one-sided STFT divided
    by sqrt(sum(w^2)) (torchaudio ``normalized=True``), Nyquist row dropped.

    mono: (N,) fp32.  Returns (n_fft/2, T) complex64, frame-major memory like the reference.
    """
    assert mono.dim() == 1
    window = torch.hann_window(n_fft)
    padded = torch.nn.functional.pad(mono[None, None, :], (n_fft // 2, n_fft // 2), mode="reflect")[0, 0]
    frames = padded.unfold(0, n_fft, hop)                      # (T, n_fft)
    spec = torch.fft.rfft(frames * window, dim=1)              # (T, n_fft/2+1)
    spec = spec / window.pow(2.0).sum().sqrt()
    return spec.transpose(0, 1)[:-1, :]


def diff(x: torch.Tensor) -> torch.Tensor:
    """functions.py:13-14."""
    d = x[:, 1:] - x[:, :-1]
    return torch.cat([torch.zeros_like(x[:, :1]), d], dim=1)


def unwrap(phi: torch.Tensor) -> torch.Tensor:
    """functions.py:17-23 (np.pi scalars act as float32 constants on fp32 tensors)."""
    pi = math.pi
    dphi = diff(phi)
    dphi_m = torch.remainder(dphi + pi, 2 * pi) - pi
    dphi_m = torch.where((dphi_m == -pi) & (dphi > 0), torch.full_like(dphi_m, pi), dphi_m)
    adj = dphi_m - dphi
    adj = torch.where(dphi.abs() < pi, torch.zeros_like(adj), adj)
    return phi + adj.cumsum(1)


def bark_gain(n_bins: int) -> torch.Tensor:
    """functions.py:29-33: 6*asinh(linspace(20, 22050, F)/600), L2-normalised, shape (F, 1)."""
    scale = 6.0 * torch.arcsinh(torch.linspace(20.0, float(44100 // 2), n_bins) / 600.0)[:, None]
    return scale / scale.norm()


def phase_magn_raw(cv: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """functions.py:69-77 up to (not including) the min/max normalisation.
    Returns (magn, ifreq), each (F, T-1)."""
    magn = torch.abs(cv) * bark_gain(cv.size(0))
    phase = unwrap(torch.angle(cv))
    return magn[:, 1:], phase[:, 1:] - phase[:, :-1]


def stft_to_phase_magn(cv: torch.Tensor, nb_vec: int = N_VEC) -> Tuple[torch.Tensor, torch.Tensor]:
    """functions.py:65-94.  Returns (magn, phase), each (n_chunks, F, nb_vec) in [-1, 1]."""
    magn, ifr = phase_magn_raw(cv)
    mx_m, mn_m, mx_p, mn_p = magn.max(), magn.min(), ifr.max(), ifr.min()
    magn = (magn - mn_m) / (mx_m - mn_m)
    ifr = (ifr - mn_p) / (mx_p - mn_p)
    magn, ifr = magn * 2.0 - 1.0, ifr * 2.0 - 1.0
    head = magn.size(1) % nb_vec
    magn, ifr = magn[:, head:], ifr[:, head:]
    return (torch.stack(magn.split(nb_vec, dim=1), dim=0),
            torch.stack(ifr.split(nb_vec, dim=1), dim=0))


def wav_to_magn_phase(mono: torch.Tensor, nb_vec: int = N_VEC):
    """create_dataset.py:35-47 on an in-memory mono waveform."""
    return stft_to_phase_magn(stft_c64(mono), nb_vec)


# ----------------------------------------------------------------------------
# inverse transform
# ----------------------------------------------------------------------------
def magn_phase_to_complex(magn_phase: torch.Tensor) -> torch.Tensor:
    """functions.py:108-128: (N,2,512,W) -> complex (513, N*W) spectrum incl. zero Nyquist."""
    assert magn_phase.dim() == 4 and magn_phase.size(1) == 2 and magn_phase.size(2) == N_FFT // 2
    flat = magn_phase.permute(1, 2, 0, 3).flatten(2, 3)
    magn, phase = flat[0], flat[1].clone()
    magn = (magn + 1.0) / 2.0
    magn = magn / bark_gain(magn.size(0))
    magn = magn / (magn.max() - magn.min())
    phase = (phase + 1.0) / 2.0 * 2.0 * math.pi - math.pi
    # functions.py:117-118 -- strictly sequential float32 running sum along time
    acc = phase[:, 0].clone()
    cols = [acc]
    for i in range(1, phase.size(1)):
        acc = acc + phase[:, i]
        cols.append(acc)
    phase = torch.stack(cols, dim=1)
    phase = torch.remainder(phase, 2 * math.pi)
    real, imag = magn * torch.cos(phase), magn * torch.sin(phase)
    zero = torch.zeros(1, real.size(1))
    return torch.complex(torch.cat([real, zero], 0), torch.cat([imag, zero], 0))


def istft(z: torch.Tensor) -> torch.Tensor:
    """functions.py:130-137: torchaudio inverse_spectrogram(normalized=True, length=None)
    == istft(z * sqrt(sum w^2)) with Hann / 1024 / 256, centred.  Returns (256*(T-1),)."""
    window = torch.hann_window(N_FFT)
    z = z * window.pow(2.0).sum().sqrt()
    return torch.istft(z, n_fft=N_FFT, hop_length=STFT_STRIDE, win_length=N_FFT,
                       window=window, center=True, normalized=False, onesided=True, length=None)


def magn_phase_to_wav(magn_phase: torch.Tensor) -> torch.Tensor:
    """functions.py:97-137 without the file write (:139)."""
    return istft(magn_phase_to_complex(magn_phase))
