"""ctypes binding of libmusicgan_b200.so (the C ABI declared in include/musicgan_b200.h).

There is deliberately no fallback: if the library is missing or a call fails, the caller gets an
exception.  torch is used only for device memory and streams (``tensor.data_ptr()``,
``torch.cuda.current_stream().cuda_stream``).
"""
from __future__ import annotations

import ctypes
import os
from ctypes import c_char_p, c_float, c_int, c_int64, c_size_t, c_void_p, POINTER

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libmusicgan_b200.so")

_lib = None


class MgError(RuntimeError):
    """A C-ABI call returned a negative mgError code."""


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} not found: build it with `python -m musicgan_b200.build` "
                "(there is no CPU / PyTorch fallback for the hot path)")
        _lib = ctypes.CDLL(LIB_PATH)
        _declare(_lib)
    return _lib


def _declare(l: ctypes.CDLL) -> None:
    l.mg_version.restype = c_int
    l.mg_error_string.restype = c_char_p
    l.mg_error_string.argtypes = [c_int]
    l.mg_last_cuda_error.restype = c_char_p
    l.mg_device_info.argtypes = [POINTER(c_int)] * 3
    l.mg_chunk_plan.argtypes = [c_int64, c_int, c_int, POINTER(c_int64), POINTER(c_int64), POINTER(c_int64)]
    l.mg_fill_hann_host.argtypes = [POINTER(c_float), c_int]
    l.mg_fill_hann_host.restype = None
    l.mg_fill_bark_gain_host.argtypes = [POINTER(c_float), c_int]
    l.mg_fill_bark_gain_host.restype = None
    l.mg_stft_magif_workspace_bytes.restype = c_size_t
    l.mg_stft_magif_workspace_bytes.argtypes = [c_int64, c_int]
    l.mg_stft_magif_f32.argtypes = [c_void_p, c_int64, c_int, c_int64, c_int, c_void_p, c_void_p,
                                    c_void_p, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]
    l.mg_stft_c64.argtypes = [c_void_p, c_int64, c_int, c_int64, c_int, c_void_p, c_void_p, c_void_p]
    l.mg_phase_magn_workspace_bytes.restype = c_size_t
    l.mg_phase_magn_workspace_bytes.argtypes = [c_int64, c_int]
    l.mg_phase_magn_from_stft.argtypes = [c_void_p, c_int64, c_int64, c_int64, c_int64, c_int, c_void_p,
                                          c_void_p, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]
    l.mg_profile_enable.argtypes = [c_int]
    l.mg_profile_enable.restype = None
    l.mg_profile_collect.argtypes = [c_int, POINTER(c_char_p), POINTER(c_float), POINTER(c_int)]
    if hasattr(l, "mg_istft_workspace_bytes"):
        l.mg_istft_workspace_bytes.restype = c_size_t
        l.mg_istft_workspace_bytes.argtypes = [c_int, c_int, c_int]
        l.mg_istft_from_magif_f32.argtypes = [c_void_p, c_int, c_int, c_int, c_void_p, c_void_p,
                                              c_void_p, c_void_p, c_size_t, c_void_p]


def check(rc: int, what: str) -> None:
    if rc != 0:
        l = lib()
        msg = l.mg_error_string(rc).decode()
        cuda = l.mg_last_cuda_error().decode()
        raise MgError(f"{what}: {msg} (code {rc})" + (f" [{cuda}]" if cuda else ""))


def chunk_plan(n_samples: int, hop: int = 256, nb_vec: int = 512):
    """(T, head, n_chunks) -- integer plan of audio/functions.py:53-62,76-92."""
    t, h, c = c_int64(), c_int64(), c_int64()
    check(lib().mg_chunk_plan(n_samples, hop, nb_vec, ctypes.byref(t), ctypes.byref(h), ctypes.byref(c)), "mg_chunk_plan")
    return t.value, h.value, c.value


def profile_enable(on: bool) -> None:
    lib().mg_profile_enable(1 if on else 0)


def profile_collect(max_entries: int = 64):
    """{kernel name: (total ms, launches)} of the launches recorded since the last collect."""
    names = (c_char_p * max_entries)()
    ms = (c_float * max_entries)()
    cnt = (c_int * max_entries)()
    n = lib().mg_profile_collect(max_entries, names, ms, cnt)
    return {names[i].decode(): (float(ms[i]), int(cnt[i])) for i in range(n)}
