"""Differentiable convolution blocks on top of the tcgen05 kernels (autograd glue only).

Three bilinear primitives, each a torch.autograd.Function whose backward is written with the other two,
so the family is closed under differentiation and `autograd.grad(..., create_graph=True)` (the WGAN-GP
penalty, reference networks/discriminator.py:157-184) works through them:

    fprop(x, w)  = conv3x3(x, w)                      d/dx -> dgrad(g, w)    d/dw -> wgrad(g, x)
    dgrad(g, w)  = conv3x3_transposed(g, w)           d/dg -> fprop(gg, w)   d/dw -> wgrad(g, gg)
    wgrad(g, x)  = sum_pixels g (x) shifted x         d/dg -> fprop(x, gw)   d/dx -> dgrad(g, gw)

Activations: bf16, channels_last.  Weights / weight gradients: fp32.
"""
from __future__ import annotations

import torch as th
import torch.nn.functional as F
from torch.autograd import Function

from . import ops

LRELU_SLOPE = 0.2
PN_EPS = 1e-8


def _act(t: th.Tensor) -> th.Tensor:
    return ops.as_act(t)


class ConvFprop(Function):
    """y = conv3x3(x, w), no bias / activation (linear in x and in w)."""

    @staticmethod
    def forward(ctx, x, w):
        ctx.save_for_backward(x, w)
        return ops.conv3x3(_act(x), w.float().contiguous())

    @staticmethod
    def backward(ctx, gy):
        x, w = ctx.saved_tensors
        gx = ConvDgrad.apply(gy, w) if ctx.needs_input_grad[0] else None
        gw = ConvWgrad.apply(gy, x) if ctx.needs_input_grad[1] else None
        return gx, gw


class ConvDgrad(Function):
    """dx = data gradient of conv3x3(., w) for output gradient g."""

    @staticmethod
    def forward(ctx, g, w):
        ctx.save_for_backward(g, w)
        return ops.conv3x3(_act(g), w.float().contiguous(), dgrad=True)

    @staticmethod
    def backward(ctx, gdx):
        g, w = ctx.saved_tensors
        gg = ConvFprop.apply(gdx, w) if ctx.needs_input_grad[0] else None
        gw = ConvWgrad.apply(g, gdx) if ctx.needs_input_grad[1] else None
        return gg, gw


class ConvWgrad(Function):
    """dw = weight gradient for output gradient g and input x (fp32, shape (Cout, Cin, 3, 3))."""

    @staticmethod
    def forward(ctx, g, x):
        ctx.save_for_backward(g, x)
        return ops.conv3x3_wgrad(_act(g), _act(x))

    @staticmethod
    def backward(ctx, gdw):
        g, x = ctx.saved_tensors
        gg = ConvFprop.apply(x, gdw) if ctx.needs_input_grad[0] else None
        gx = ConvDgrad.apply(g, gdw) if ctx.needs_input_grad[1] else None
        return gg, gx


def _lrelu_mask(y: th.Tensor) -> th.Tensor:
    # sign(lrelu(z)) == sign(z): the mask is recovered from the saved OUTPUT, no pre-activation is kept
    one = th.ones((), dtype=y.dtype, device=y.device)
    return th.where(y > 0, one, one * LRELU_SLOPE)


class ConvBiasLReLU(Function):
    """y = LeakyReLU_0.2(conv3x3(x, w) + b), bias and activation fused in the kernel epilogue.
    Backward is composed of differentiable pieces (mask multiply, ConvDgrad, ConvWgrad), so double backward works.
    Reference: discriminator.py:15-22,26-33 (Conv2d + LeakyReLU pairs of ConvBlock)."""

    @staticmethod
    def forward(ctx, x, w, b):
        y = ops.conv3x3(_act(x), w.float().contiguous(), b.float().contiguous(), lrelu=True)
        ctx.save_for_backward(x, w, y)
        return y

    @staticmethod
    def backward(ctx, gy):
        x, w, y = ctx.saved_tensors
        gz = gy * _lrelu_mask(y)
        gx = ConvDgrad.apply(gz, w) if ctx.needs_input_grad[0] else None
        gw = ConvWgrad.apply(gz, x) if ctx.needs_input_grad[1] else None
        gb = gz.float().sum(dim=(0, 2, 3)) if ctx.needs_input_grad[2] else None
        return gx, gw, gb


class GenConv(Function):
    """One generator half-block: [nearest x2 upsample ->] conv3x3 + bias -> LeakyReLU(0.2) -> PixelNorm, all in one
    kernel (generator.py:16-24 / :26-40, layers.py:11-17).  First-order backward only (the generator is never
    differentiated twice: the gradient penalty graph lives in the discriminator)."""

    @staticmethod
    def forward(ctx, x, w, b, upsample_in: bool):
        cout, cin = w.shape[0], w.shape[1]
        fused_pn = 9 * cin * cout * 2 <= 120 * 1024      # weights of all Cout resident in smem (conv_igemm.cu plan)
        xa = _act(x)
        wf, bf = w.float().contiguous(), b.float().contiguous()
        if fused_pn:
            o, inv = ops.conv3x3(xa, wf, bf, lrelu=True, pixelnorm=True, upsample_in=upsample_in, want_inv_norm=True)
        else:
            t = ops.conv3x3(xa, wf, bf, lrelu=True, upsample_in=upsample_in).float()
            inv = th.rsqrt(t.pow(2).mean(dim=1) + PN_EPS)
            o = _act(t * inv[:, None])
        ctx.save_for_backward(xa, w, o, inv)
        ctx.upsample_in = upsample_in
        return o

    @staticmethod
    @th.autograd.function.once_differentiable
    def backward(ctx, go):
        xa, w, o, inv = ctx.saved_tensors
        of, gf = o.float(), go.float()
        # PixelNorm backward: t = o * n ; g_t = (g_o - o * mean_c(g_o * o)) / n
        gt = (gf - of * (gf * of).mean(dim=1, keepdim=True)) * inv[:, None]
        gz = _act(gt * _lrelu_mask(of))
        gx = gw = gb = None
        wf = w.float().contiguous()
        if ctx.needs_input_grad[0]:
            gx = ops.conv3x3(gz, wf, dgrad=True)
            if ctx.upsample_in:      # backward of the nearest upsampling folded into the read: sum of each 2x2 block
                gx = _act(F.avg_pool2d(gx.float(), 2) * 4.0)
        if ctx.needs_input_grad[1]:
            gw = ops.conv3x3_wgrad(gz, xa, upsample_in=ctx.upsample_in)
        if ctx.needs_input_grad[2]:
            gb = gz.float().sum(dim=(0, 2, 3))
        return gx, gw, gb, None


def conv1x1(x: th.Tensor, w: th.Tensor, b: th.Tensor) -> th.Tensor:
    """1x1 convolutions of the to/from magnitude-phase layers (generator.py:46-50, discriminator.py:43-48):
    K = 2 or N = 2, memory bound, not tensor-core work; bf16 operands, fp32 accumulate, fp32 bias add."""
    y = F.conv2d(x, w.to(x.dtype))
    return y.float() + b.float()[None, :, None, None]
