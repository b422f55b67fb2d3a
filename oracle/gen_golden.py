"""Generate tests/golden/*.npz from the UNMODIFIED reference (authoring container only).

Run:  python oracle/gen_golden.py            (needs /root/reference; never run on the GPU box)

For every case it (1) runs the reference imported from /root/reference with the three shims of
SURVEY Appendix A (mlflow / matplotlib stubs, torchaudio.load/save patched to memory), (2) runs
the oracle restatement on the same seeded input and records whether the two are bit-identical
here, and (3) stores a strided subsample of the reference outputs + float64 checksums, so the
fixtures stay small.  Inputs are regenerated from the seed by ``oracle/cases.py``.
"""
import os
import sys
import types
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
sys.dont_write_bytecode = True
warnings.filterwarnings("ignore")

for _n in ("mlflow", "matplotlib", "matplotlib.pyplot"):
    sys.modules[_n] = types.ModuleType(_n)
sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
sys.path.insert(0, "/root/reference")

import torch as th            # noqa: E402
import torchaudio             # noqa: E402
from music_gan import audio as ref_audio          # noqa: E402  (reference, unmodified)
from music_gan import networks as ref_networks    # noqa: E402

from oracle import audio_oracle as ao             # noqa: E402
from oracle import cases                          # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")
os.makedirs(GOLD, exist_ok=True)


def sub(t: th.Tensor, stride: int) -> np.ndarray:
    return t.detach().contiguous().view(-1)[::stride].numpy().copy()


def digest(t: th.Tensor) -> np.ndarray:
    d = t.detach().double()
    return np.array([d.sum().item(), d.abs().sum().item(), d.min().item(), d.max().item()], dtype=np.float64)


def audio_forward():
    for name in cases.FORWARD_CASES:
        wav = cases.forward_wav(name)                       # (C, N)
        torchaudio.load = lambda p, w=wav: (w, 44100)
        cv = ref_audio.wav_to_stft("mem.wav")
        magn, phase = ref_audio.stft_to_phase_magn(cv)
        mono = wav.mean(0)
        o_cv = ao.stft_c64(mono)
        o_m, o_p = ao.stft_to_phase_magn(o_cv)
        raw_m, raw_p = ao.phase_magn_raw(cv)
        exact = bool(th.equal(th.view_as_real(cv), th.view_as_real(o_cv)) and th.equal(magn, o_m) and th.equal(phase, o_p))
        t, head, n_chunks = ao.chunk_plan(wav.size(1))
        assert cv.size(1) == t and magn.size(0) == n_chunks, (cv.shape, t, magn.shape, n_chunks)
        stride = cases.FORWARD_CASES[name]["stride"]
        np.savez_compressed(
            os.path.join(GOLD, f"audio_forward_{name}.npz"),
            n_samples=np.int64(wav.size(1)), T=np.int64(cv.size(1)), head=np.int64(head),
            n_chunks=np.int64(magn.size(0)), stride=np.int64(stride),
            oracle_bit_exact_in_authoring_container=np.bool_(exact),
            stft_re=sub(cv.real, stride), stft_im=sub(cv.imag, stride),
            magn=sub(magn, stride), phase=sub(phase, stride),
            magn_digest=digest(magn), phase_digest=digest(phase),
            raw_minmax=np.array([raw_m.min().item(), raw_m.max().item(), raw_p.min().item(), raw_p.max().item()], dtype=np.float32),
        )
        print(f"forward/{name}: T={t} head={head} chunks={n_chunks} oracle==reference: {exact}")


def audio_inverse():
    saved = {}
    torchaudio.save = lambda p, w, sr: saved.__setitem__(p, (w.clone(), sr))
    for name in cases.INVERSE_CASES:
        mp = cases.inverse_input(name)
        ref_audio.magn_phase_to_wav(mp.clone(), "o.wav", 44100)
        wav = saved["o.wav"][0][0]
        o_wav = ao.magn_phase_to_wav(mp.clone())
        exact = bool(th.equal(wav, o_wav))
        stride = cases.INVERSE_CASES[name]["stride"]
        np.savez_compressed(
            os.path.join(GOLD, f"audio_inverse_{name}.npz"),
            n_out=np.int64(wav.numel()), stride=np.int64(stride),
            oracle_bit_exact_in_authoring_container=np.bool_(exact),
            wav=sub(wav, stride), wav_digest=digest(wav),
        )
        print(f"inverse/{name}: out={wav.numel()} oracle==reference: {exact}")


def index_plan():
    # create_dataset.py:32-64 idx bookkeeping incl. the T<512 skip and T==512 empty-chunk quirk,
    # measured by running the reference transform itself on zero-information (random) clips.
    counts = cases.INDEX_PLAN_SAMPLE_COUNTS
    rows, idx = [], 0
    for n in counts:
        th.manual_seed(n)
        wav = th.rand(1, n) - 0.5
        torchaudio.load = lambda p, w=wav: (w, 44100)
        cv = ref_audio.wav_to_stft("mem.wav")
        if cv.size(1) < ref_audio.N_VEC:
            rows.append((n, cv.size(1), idx, 0, 0))
            continue
        m, _ = ref_audio.stft_to_phase_magn(cv)
        rows.append((n, cv.size(1), idx, m.size(0), m.size(2)))
        idx += m.size(0)
    np.savez_compressed(os.path.join(GOLD, "index_plan.npz"), rows=np.array(rows, dtype=np.int64))
    print("index_plan rows (n_samples, T, first_idx, n_written, chunk_width):")
    for r in rows:
        print("   ", r)


def scale_transform():
    """Real-batch transform of the training loop (utils.py:70-82, train.py:139-140): ChannelMinMaxNorm ->
    ChangeRange(-1, 1) -> torchvision Resize(512 / 2^k), run through the reference's own Grower at several growth stages."""
    from music_gan import utils as ref_utils
    g = th.Generator().manual_seed(4242)
    x = th.rand(3, 2, 512, 512, generator=g, dtype=th.float64).float() * 3.0 - 1.0      # a batch of dataset chunks
    out = {"seed": np.int64(4242)}
    fade, lens = [1, 2, 2, 2, 2, 2, 2, 2], [1, 2, 3, 4, 5, 6, 7]
    grower = ref_utils.Grower(7, fade, lens)
    stage = 0
    while True:
        y = grower.scale_transform(x)
        size = 4 * 2 ** stage
        assert tuple(y.shape) == (3, 2, size, size), y.shape
        stride = max(1, y.numel() // 8192)
        out[f"stage{stage}"] = y.contiguous().view(-1)[::stride].numpy().copy()
        out[f"stage{stage}_stride"] = np.int64(stride)
        out[f"stage{stage}_digest"] = digest(y)
        if stage == 7:
            break
        while not grower.grow(1):
            pass
        stage += 1
    np.savez_compressed(os.path.join(GOLD, "scale_transform.npz"), **out)
    print("scale_transform: stages 0..7 written")


if __name__ == "__main__":
    which = sys.argv[1:] or ["forward", "inverse", "index", "networks", "scale"]
    if "scale" in which:
        scale_transform()
    if "forward" in which:
        audio_forward()
    if "inverse" in which:
        audio_inverse()
    if "index" in which:
        index_plan()
    if "networks" in which:
        from oracle import gen_golden_networks
        gen_golden_networks.main(ref_networks, GOLD)
