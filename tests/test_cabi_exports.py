"""CPU: the C-ABI library builds, loads and exports every symbol include/musicgan_b200.h declares
(no compute calls here: there is no GPU in the authoring container)."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared(header):
    src = open(os.path.join(ROOT, "include", header)).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(mg_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from musicgan_b200 import build, _lib
    build.build(verbose=False)
    l = ctypes.CDLL(_lib.LIB_PATH)
    declared = _declared("musicgan_b200.h")
    assert len(declared) >= 30
    missing = [n for n in declared if not hasattr(l, n)]
    assert not missing, missing
    assert l.mg_version() >= 100
    # the test-only probes live in their own library, not in the product
    dbg = ctypes.CDLL(build.DEBUG_LIB)
    probes = _declared("musicgan_b200_debug.h")
    assert probes and all(hasattr(dbg, n) for n in probes)
    assert not any(hasattr(l, n) for n in probes)
    assert sorted(os.listdir(os.path.join(ROOT, "include"))) == ["musicgan_b200.h", "musicgan_b200_debug.h"]


def test_host_only_entry_points():
    from musicgan_b200 import _lib
    from oracle import audio_oracle as ao
    for n in [0, 255, 256, 100_000, 130_816, 131_072, 131_327, 2_646_000, 7_777_777]:
        assert _lib.chunk_plan(n) == ao.chunk_plan(n)
    l = _lib.lib()
    assert l.mg_error_string(-3).decode().startswith("workspace")
    # argument validation happens before any CUDA call
    assert l.mg_stft_magif_f32(None, 1000, 1, 1000, 1, None, None, None, None, None, None, 0, None) == -1
    import numpy as np
    import torch
    w = (ctypes.c_float * 1024)()
    l.mg_fill_hann_host(w, 1024)
    np.testing.assert_allclose(np.frombuffer(w, dtype=np.float32), torch.hann_window(1024).numpy(), atol=3e-7)   # torch builds it in fp32
    b = (ctypes.c_float * 512)()
    l.mg_fill_bark_gain_host(b, 512)
    np.testing.assert_allclose(np.frombuffer(b, dtype=np.float32), ao.bark_gain(512)[:, 0].numpy(), rtol=2e-6)


def test_fft_lane_emulation(tmp_path):
    """Warp-level FFT phases stepped on the host: FFT/split correctness and bank-conflict freedom."""
    import subprocess
    exe = str(tmp_path / "fft_emu")
    subprocess.run(["g++", "-O2", "-std=c++17", "-o", exe, os.path.join(ROOT, "tests", "emu", "fft_emu.cpp")], check=True)
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
