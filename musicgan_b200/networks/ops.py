"""Raw (non-autograd) wrappers of the convolution entry points of the C ABI.

Activations are logical (B, C, H, W) bf16 tensors in torch.channels_last memory format, i.e. NHWC in
memory -- exactly what the kernels read and write.  Weights are the fp32 master tensors
(Cout, Cin, 3, 3) of the drop-in nn.Modules (same shapes as the reference's nn.Conv2d)."""
from __future__ import annotations

import ctypes
from ctypes import c_int, c_size_t, c_void_p

import torch as th

from .. import _lib

FLAG_LRELU, FLAG_PIXELNORM, FLAG_UPSAMPLE_IN, FLAG_DGRAD = 1, 2, 4, 8

_declared = False
_ws_cache = {}


def _l():
    global _declared
    l = _lib.lib()
    if not _declared:
        l.mg_conv3x3_workspace_bytes.restype = c_size_t
        l.mg_conv3x3_workspace_bytes.argtypes = [c_int, c_int]
        l.mg_conv3x3_bf16.argtypes = [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                      c_int, c_int, c_int, c_int, c_int, c_int, c_void_p, c_size_t, c_void_p]
        if hasattr(l, "mg_conv3x3_wgrad_bf16"):
            l.mg_conv3x3_wgrad_bf16.argtypes = [c_void_p, c_void_p, c_void_p, c_void_p,
                                                c_int, c_int, c_int, c_int, c_int, c_int, c_void_p]
        _declared = True
    return l


def _workspace(dev, nbytes: int) -> th.Tensor:
    key = (dev.index, th.cuda.current_stream(dev).cuda_stream)
    ws = _ws_cache.get(key)
    if ws is None or ws.numel() < nbytes:
        ws = th.empty(max(nbytes, 1 << 20), dtype=th.uint8, device=dev)
        _ws_cache[key] = ws
    return ws


def _check_act(x: th.Tensor, name: str):
    if not (x.is_cuda and x.dtype == th.bfloat16 and x.dim() == 4):
        raise TypeError(f"{name}: expected a CUDA bf16 (B, C, H, W) tensor, got {x.dtype} {tuple(x.shape)} on {x.device}")
    if not x.is_contiguous(memory_format=th.channels_last):
        raise ValueError(f"{name}: activations must be channels_last (NHWC in memory)")


def as_act(x: th.Tensor) -> th.Tensor:
    """bf16 + channels_last (no copy if already so)."""
    return x.to(dtype=th.bfloat16).contiguous(memory_format=th.channels_last)


def conv3x3(x: th.Tensor, w: th.Tensor, bias=None, *, lrelu=False, pixelnorm=False, upsample_in=False,
            dgrad=False, want_inv_norm=False):
    """y = conv3x3(x) (+bias) (+LeakyReLU 0.2) (+PixelNorm); `upsample_in` reads x through a nearest x2
    upsampling; `dgrad` computes the data gradient of the forward conv with weight `w` for x = dL/dy."""
    _check_act(x, "conv3x3 x")
    assert w.is_cuda and w.dtype == th.float32 and w.is_contiguous() and w.dim() == 4 and w.shape[2:] == (3, 3)
    B, C, Hin, Win = x.shape
    H, W = (2 * Hin, 2 * Win) if upsample_in else (Hin, Win)
    if dgrad:
        assert w.shape[0] == C, (w.shape, C)
        cin, cout = C, w.shape[1]
    else:
        assert w.shape[1] == C, (w.shape, C)
        cin, cout = C, w.shape[0]
    flags = (FLAG_LRELU if lrelu else 0) | (FLAG_PIXELNORM if pixelnorm else 0) | \
            (FLAG_UPSAMPLE_IN if upsample_in else 0) | (FLAG_DGRAD if dgrad else 0)
    y = th.empty((B, cout, H, W), dtype=th.bfloat16, device=x.device, memory_format=th.channels_last)
    inv = th.empty((B, H, W), dtype=th.float32, device=x.device) if (pixelnorm and want_inv_norm) else None
    l = _l()
    nbytes = l.mg_conv3x3_workspace_bytes(cin, cout)
    ws = _workspace(x.device, nbytes)
    if bias is not None:
        assert bias.dtype == th.float32 and bias.numel() == cout and bias.is_cuda
    with th.cuda.device(x.device):
        _lib.check(l.mg_conv3x3_bf16(x.data_ptr(), w.data_ptr(), bias.data_ptr() if bias is not None else None,
                                     y.data_ptr(), inv.data_ptr() if inv is not None else None,
                                     B, H, W, cin, cout, flags, ws.data_ptr(), ws.numel(),
                                     th.cuda.current_stream().cuda_stream), "mg_conv3x3_bf16")
    return (y, inv) if want_inv_norm else y


def conv3x3_wgrad(dy: th.Tensor, x: th.Tensor, *, upsample_in=False) -> th.Tensor:
    """dw[co][ci][ky][kx] = sum_{b,y,x} dy[b,co,y,x] * xin[b,ci,y+ky-1,x+kx-1]  (fp32), xin = x or its nearest
    x2 upsampling."""
    _check_act(dy, "wgrad dy"); _check_act(x, "wgrad x")
    B, cout, H, W = dy.shape
    cin = x.shape[1]
    assert x.shape[0] == B and (x.shape[2] * (2 if upsample_in else 1), x.shape[3] * (2 if upsample_in else 1)) == (H, W)
    dw = th.zeros((cout, cin, 3, 3), dtype=th.float32, device=dy.device)
    l = _l()
    with th.cuda.device(dy.device):
        _lib.check(l.mg_conv3x3_wgrad_bf16(dy.data_ptr(), x.data_ptr(), dw.data_ptr(), None,
                                           B, H, W, cin, cout, 1 if upsample_in else 0,
                                           th.cuda.current_stream().cuda_stream), "mg_conv3x3_wgrad_bf16")
    return dw
