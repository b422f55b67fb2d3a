"""Host-side mirror of the reference's ``music_gan/audio/functions.py`` on top of the C ABI.

Same names, argument meaning, return layout and assertion messages as the reference
(functions.py:13-139); the arithmetic runs in the sm_100a kernels of libmusicgan_b200.so.
Tensors may live on the CPU (as in the reference: they are copied to the current CUDA device and
the results copied back) or already on the GPU (results stay there).  There is no CPU fallback.

Extra, batch-oriented entry points (not in the reference) used by create_dataset / generate /
bench: :func:`wav_to_magn_phase_batch`, :func:`magn_phase_to_wav_batch`.
"""
from __future__ import annotations

import math
from typing import Optional, Tuple

import torch as th

from . import constant, wavio
from .. import _lib

_CONST_CACHE = {}


def _dev(device=None) -> th.device:
    if device is not None and th.device(device).type == "cuda":
        return th.device(device)
    if not th.cuda.is_available():
        raise RuntimeError("musicgan_b200 needs a CUDA (sm_100a) device: the transform has no CPU fallback")
    return th.device("cuda", th.cuda.current_device())


def hann_window(n_fft: int, device) -> th.Tensor:
    key = ("hann", n_fft, str(device))
    if key not in _CONST_CACHE:
        _CONST_CACHE[key] = th.hann_window(n_fft).to(device)           # functions.py:51 (computed as the reference does)
    return _CONST_CACHE[key]


def bark_gain(n_bins: int, device="cpu") -> th.Tensor:
    """functions.py:29-33 -- per-bin gain 6*asinh(f/600)/||.||, f = linspace(20, 22050, F); shape (F,)."""
    key = ("bark", n_bins, str(device))
    if key not in _CONST_CACHE:
        scale = 6. * th.arcsinh(th.linspace(20., float(44100 // 2), n_bins) / 600.)
        _CONST_CACHE[key] = (scale / scale.norm()).to(device)
    return _CONST_CACHE[key]


def _stream() -> int:
    return th.cuda.current_stream().cuda_stream


# --------------------------------------------------------------------------------------------
# module-level helpers the reference exposes (functions.py:13-35); thin torch expressions that
# run on whatever device the argument lives on.  The fused kernels do NOT call these.
# --------------------------------------------------------------------------------------------
def diff(x: th.Tensor) -> th.Tensor:
    return th.nn.functional.pad(x[:, 1:] - x[:, :-1], (1, 0, 0, 0), "constant", 0)


def unwrap(phi: th.Tensor) -> th.Tensor:
    dphi = diff(phi)
    dphi_m = ((dphi + math.pi) % (2 * math.pi)) - math.pi
    dphi_m[(dphi_m == -math.pi) & (dphi > 0)] = math.pi
    phi_adj = dphi_m - dphi
    phi_adj[dphi.abs() < math.pi] = 0
    # the reference runs on CPU where cumsum accumulates in float64; spell it out so
    # that a CUDA tensor gets the same values
    return phi + phi_adj.double().cumsum(1).to(phi.dtype)


def bark_magn_scale(magn: th.Tensor, unscale: bool = False) -> th.Tensor:
    assert len(magn.size()) == 2, f"(STFT, TIME), actual = {magn.size()}"
    scale_norm = bark_gain(magn.size()[0], magn.device)[:, None]
    return magn / scale_norm if unscale else magn * scale_norm


# --------------------------------------------------------------------------------------------
# forward transform
# --------------------------------------------------------------------------------------------
def _check_geometry(nperseg: int, stride: int) -> None:
    if nperseg != constant.N_FFT or stride != constant.STFT_STRIDE:
        raise NotImplementedError(
            f"the sm_100a transform kernels are built for n_fft={constant.N_FFT}, hop={constant.STFT_STRIDE} "
            f"(the reference's constants); got n_fft={nperseg}, hop={stride}")


def stft_from_wave(raw_audio: th.Tensor, nperseg: int = constant.N_FFT, stride: int = constant.STFT_STRIDE,
                   device=None) -> th.Tensor:
    """functions.py:49-62 on an in-memory (channels, samples) waveform.

    Returns the complex64 (n_fft/2, T) STFT (Nyquist dropped) as a CUDA tensor whose memory is
    frame major, like the reference's."""
    _check_geometry(nperseg, stride)
    dev = _dev(device if device is not None else raw_audio.device)
    wav = raw_audio.to(dev, th.float32).contiguous()
    if wav.dim() == 1:
        wav = wav[None]
    ch, n = wav.shape
    t = 1 + n // stride
    out = th.empty(t, nperseg // 2, dtype=th.complex64, device=dev)
    with th.cuda.device(dev):
        _lib.check(_lib.lib().mg_stft_c64(wav.data_ptr(), n, ch, ch * n, 1, hann_window(nperseg, dev).data_ptr(),
                                          out.data_ptr(), _stream()), "mg_stft_c64")
    return out.transpose(0, 1)


def wav_to_stft(wav_p: str, nperseg: int = constant.N_FFT, stride: int = constant.STFT_STRIDE) -> th.Tensor:
    """Drop-in for functions.py:38-62; returns a CPU complex64 tensor (512, T) like the reference."""
    raw_audio, sr = wavio.load(wav_p)
    assert sr == constant.SAMPLE_RATE, \
        f"Audio sample rate must be {constant.SAMPLE_RATE}Hz, " \
        f"file \"{wav_p}\" is {sr}Hz"
    return stft_from_wave(raw_audio, nperseg, stride).cpu()


def stft_to_phase_magn(complex_values: th.Tensor, nb_vec: int = constant.N_VEC) -> Tuple[th.Tensor, th.Tensor]:
    """Drop-in for functions.py:65-94: (F=512, T) complex64 -> (magn, phase), each (n_chunks, 512, nb_vec)."""
    if nb_vec != constant.N_VEC or complex_values.size(0) != constant.N_FFT // 2:
        raise NotImplementedError("the sm_100a kernels are built for 512 bins and nb_vec=512 (reference constants)")
    assert complex_values.dim() == 2 and complex_values.dtype == th.complex64
    src_dev = complex_values.device
    dev = _dev(src_dev)
    cv = complex_values.to(dev)
    f, t = cv.shape
    assert t >= 2, "need at least two STFT frames"
    n_chunks = ((t - 1) - (t - 1) % nb_vec) // nb_vec
    magn = th.empty(n_chunks, f, nb_vec, dtype=th.float32, device=dev)
    phase = th.empty(n_chunks, f, nb_vec, dtype=th.float32, device=dev)
    minmax = th.empty(4, dtype=th.float32, device=dev)
    l = _lib.lib()
    ws_bytes = l.mg_phase_magn_workspace_bytes(t, 1)
    ws = th.empty(ws_bytes, dtype=th.uint8, device=dev)
    with th.cuda.device(dev):
        _lib.check(l.mg_phase_magn_from_stft(cv.data_ptr(), t, cv.stride(0), cv.stride(1), 0, 1,
                                             bark_gain(f, dev).data_ptr(), magn.data_ptr(), phase.data_ptr(),
                                             minmax.data_ptr(), ws.data_ptr(), ws_bytes, _stream()),
                   "mg_phase_magn_from_stft")
    if n_chunks == 0:
        # the reference's split of an empty tensor yields ONE empty chunk (SURVEY 3.1 edge case)
        magn, phase = magn.new_empty(1, f, 0), phase.new_empty(1, f, 0)
    return magn.to(src_dev), phase.to(src_dev)


class ForwardPlan:
    """Reusable buffers for the fused transform of `batch` equally long clips (device resident)."""

    def __init__(self, n_samples: int, batch: int, channels: int = 1, device=None):
        self.dev = _dev(device)
        self.n_samples, self.batch, self.channels = n_samples, batch, channels
        self.T, self.head, self.n_chunks = _lib.chunk_plan(n_samples, constant.STFT_STRIDE, constant.N_VEC)
        l = _lib.lib()
        self.ws_bytes = l.mg_stft_magif_workspace_bytes(n_samples, batch)
        self.ws = th.empty(self.ws_bytes, dtype=th.uint8, device=self.dev)
        shape = (batch, self.n_chunks, constant.N_FFT // 2, constant.N_VEC)
        self.magn = th.empty(shape, dtype=th.float32, device=self.dev)
        self.phase = th.empty(shape, dtype=th.float32, device=self.dev)
        self.minmax = th.empty(batch, 4, dtype=th.float32, device=self.dev)
        self.window = hann_window(constant.N_FFT, self.dev)
        self.bark = bark_gain(constant.N_FFT // 2, self.dev)

    def run(self, wav: th.Tensor) -> Tuple[th.Tensor, th.Tensor]:
        """wav: CUDA fp32 (batch, channels, n_samples) or (batch, n_samples), contiguous."""
        assert wav.is_cuda and wav.dtype == th.float32 and wav.is_contiguous()
        assert wav.numel() == self.batch * self.channels * self.n_samples, tuple(wav.shape)
        with th.cuda.device(self.dev):
            _lib.check(_lib.lib().mg_stft_magif_f32(
                wav.data_ptr(), self.n_samples, self.channels, self.channels * self.n_samples, self.batch,
                self.window.data_ptr(), self.bark.data_ptr(), self.magn.data_ptr(), self.phase.data_ptr(),
                self.minmax.data_ptr(), self.ws.data_ptr(), self.ws_bytes, _stream()), "mg_stft_magif_f32")
        return self.magn, self.phase


def wav_to_magn_phase_batch(wav: th.Tensor, plan: Optional[ForwardPlan] = None) -> Tuple[th.Tensor, th.Tensor]:
    """Fused wav_to_stft + stft_to_phase_magn (create_dataset.py:35-47) for a batch of equally long
    clips: (batch, [channels,] n_samples) -> magn, phase (batch, n_chunks, 512, 512) on the GPU."""
    dev = _dev(wav.device)
    w = wav.to(dev, th.float32).contiguous()
    if w.dim() == 2:
        w = w[:, None, :]
    b, ch, n = w.shape
    if plan is None:
        plan = ForwardPlan(n, b, ch, dev)
    return plan.run(w)


# --------------------------------------------------------------------------------------------
# inverse transform
# --------------------------------------------------------------------------------------------
def magn_phase_to_wave_batch(magn_phase: th.Tensor, imgs_per_clip: int = 1) -> th.Tensor:
    """functions.py:97-137 for many clips at once: (n_clips*imgs_per_clip, 2, 512, W) ->
    (n_clips, 256*(imgs_per_clip*W-1)) fp32 on the GPU.  The `imgs_per_clip` images of a clip are
    concatenated along time and de-normalised together, exactly like one reference call."""
    assert len(magn_phase.size()) == 4, \
        f"(N, 2, H, W), actual = {magn_phase.size()}"
    assert magn_phase.size()[1] == 2, \
        f"Channels must be equal to 2, actual = {magn_phase.size()[1]}"
    assert magn_phase.size()[2] == constant.N_FFT // 2, \
        f"Frequency size must be equal to {constant.N_FFT // 2}, " \
        f"actual = {magn_phase.size()[2]}"
    dev = _dev(magn_phase.device)
    x = magn_phase.to(dev, th.float32).contiguous()
    n_img, _, f, w = x.shape
    assert n_img % imgs_per_clip == 0
    n_clips = n_img // imgs_per_clip
    assert imgs_per_clip * w >= 2, "need at least two frames"
    out = th.empty(n_clips, constant.STFT_STRIDE * (imgs_per_clip * w - 1), dtype=th.float32, device=dev)
    l = _lib.lib()
    ws_bytes = l.mg_istft_workspace_bytes(n_clips, imgs_per_clip, w)
    ws = th.empty(ws_bytes, dtype=th.uint8, device=dev)
    with th.cuda.device(dev):
        _lib.check(l.mg_istft_from_magif_f32(x.data_ptr(), n_clips, imgs_per_clip, w,
                                             hann_window(constant.N_FFT, dev).data_ptr(), bark_gain(f, dev).data_ptr(),
                                             out.data_ptr(), ws.data_ptr(), ws_bytes, _stream()),
                   "mg_istft_from_magif_f32")
    return out


def magn_phase_to_wav(magn_phase: th.Tensor, wav_path: str, sample_rate: int):
    """Drop-in for functions.py:97-139: all N images form ONE clip, written to `wav_path`."""
    raw_audio = magn_phase_to_wave_batch(magn_phase, imgs_per_clip=magn_phase.size()[0])
    wavio.save(wav_path, raw_audio.cpu(), sample_rate)
