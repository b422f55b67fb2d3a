"""Minimal RIFF/WAVE reader and writer (PCM 8/16/24/32 and IEEE float32/64).

The reference calls ``torchaudio.load`` / ``torchaudio.save`` (audio/functions.py:43,139), which in
torchaudio >= 2.9 need the torchcodec package.  This module gives the same in-memory contract --
``load(path) -> (float32 tensor (channels, samples) in [-1, 1], sample_rate)`` and
``save(path, tensor (channels, samples), sample_rate)`` -- without that dependency.
"""
from __future__ import annotations

import struct

import numpy as np
import torch

_PCM, _FLOAT, _EXTENSIBLE = 1, 3, 0xFFFE


def load(path: str):
    with open(path, "rb") as fh:
        data = fh.read()
    if data[:4] != b"RIFF" or data[8:12] != b"WAVE":
        raise ValueError(f"{path}: not a RIFF/WAVE file")
    pos, fmt, payload = 12, None, None
    while pos + 8 <= len(data):
        cid, size = data[pos:pos + 4], struct.unpack("<I", data[pos + 4:pos + 8])[0]
        body = data[pos + 8:pos + 8 + size]
        if cid == b"fmt ":
            tag, ch, sr, _br, _ba, bits = struct.unpack("<HHIIHH", body[:16])
            if tag == _EXTENSIBLE and len(body) >= 26:
                tag = struct.unpack("<H", body[24:26])[0]
            fmt = (tag, ch, sr, bits)
        elif cid == b"data":
            payload = body
        pos += 8 + size + (size & 1)
    if fmt is None or payload is None:
        raise ValueError(f"{path}: missing fmt or data chunk")
    tag, ch, sr, bits = fmt
    if tag == _PCM:
        if bits == 8:
            x = (np.frombuffer(payload, dtype=np.uint8).astype(np.float32) - 128.0) / 128.0
        elif bits == 16:
            x = np.frombuffer(payload, dtype="<i2").astype(np.float32) / 32768.0
        elif bits == 24:
            b = np.frombuffer(payload[: len(payload) // 3 * 3], dtype=np.uint8).reshape(-1, 3).astype(np.int32)
            v = b[:, 0] | (b[:, 1] << 8) | (b[:, 2] << 16)
            v = np.where(v >= 1 << 23, v - (1 << 24), v)
            x = v.astype(np.float32) / float(1 << 23)
        elif bits == 32:
            x = (np.frombuffer(payload, dtype="<i4").astype(np.float64) / float(1 << 31)).astype(np.float32)
        else:
            raise ValueError(f"{path}: unsupported PCM width {bits}")
    elif tag == _FLOAT:
        x = np.frombuffer(payload, dtype="<f4" if bits == 32 else "<f8").astype(np.float32)
    else:
        raise ValueError(f"{path}: unsupported WAVE format tag {tag}")
    n = x.size // ch
    return torch.from_numpy(np.ascontiguousarray(x[: n * ch].reshape(n, ch).T)), sr


def save(path: str, wav: torch.Tensor, sample_rate: int) -> None:
    """Write float32 IEEE WAVE (what torchaudio.save does for a float32 tensor)."""
    assert wav.dim() == 2, f"(channels, samples), actual = {tuple(wav.size())}"
    x = wav.detach().to("cpu", torch.float32).numpy().T.astype("<f4", order="C")
    ch = wav.size(0)
    payload = x.tobytes()
    fmt = struct.pack("<HHIIHH", _FLOAT, ch, sample_rate, sample_rate * ch * 4, ch * 4, 32)
    fact = struct.pack("<I", wav.size(1))
    body = b"WAVE" + b"fmt " + struct.pack("<I", len(fmt)) + fmt + b"fact" + struct.pack("<I", 4) + fact \
        + b"data" + struct.pack("<I", len(payload)) + payload
    with open(path, "wb") as fh:
        fh.write(b"RIFF" + struct.pack("<I", len(body)) + body)
