"""Differentiable convolution blocks on top of the tcgen05 kernels (autograd glue only).

Three bilinear primitives, each a torch.autograd.Function whose backward is written with the other two,
so the family is closed under differentiation and `autograd.grad(..., create_graph=True)` (the WGAN-GP
penalty, reference networks/discriminator.py:157-184) works through them:

    fprop(x, w)  = conv3x3(x, w)                      d/dx -> dgrad(g, w)    d/dw -> wgrad(g, x)
    dgrad(g, w)  = conv3x3_transposed(g, w)           d/dg -> fprop(gg, w)   d/dw -> wgrad(g, gg)
    wgrad(g, x)  = sum_pixels g (x) shifted x         d/dg -> fprop(x, gw)   d/dx -> dgrad(g, gw)

Activations: channels_last; bf16 on the layers above 32 x 32, fp32 on the low-resolution layers (the precise path,
split-bf16 operands: ops.PRECISE_MAX_RES, conv_split.cu).  Every function follows the dtype of the activation it is given;
the modules (progan.py) choose it per layer with `prep`.  Weights / weight gradients: fp32.
"""
from __future__ import annotations

import torch as th
import torch.nn.functional as F
from torch.autograd import Function

from . import ops

import os
_UNFUSED_LRELU_BWD = bool(os.environ.get("MG_UNFUSED_LRELU_BWD"))
LRELU_SLOPE = 0.2
PN_EPS = 1e-8


def _act(t: th.Tensor) -> th.Tensor:
    return ops.keep_act(t)


def prep(x: th.Tensor, h_out: int) -> th.Tensor:
    """Input of a layer whose output height is `h_out`, in the activation dtype that layer runs in (a differentiable
    cast where a bf16 layer feeds a precise one or vice versa)."""
    return ops.as_act(x, th.float32 if ops.is_precise(h_out) else th.bfloat16)


# ctx.needs_input_grad is fixed when the forward runs (does the INPUT require grad?), it does not know which
# gradients a particular backward call was asked for.  The gradient penalty's inner autograd.grad(out, x_hat,
# create_graph=True) wants the input gradient only, yet every block would also launch its weight- and bias-gradient
# kernels (a quarter of all weight-gradient launches of a critic step).  `input_grads_only()` switches them off for
# the duration of such a call.
_param_grads = [True]


class input_grads_only:
    def __enter__(self):
        self.prev, _param_grads[0] = _param_grads[0], False

    def __exit__(self, *exc):
        _param_grads[0] = self.prev
        return False


def _lane_bias(w, b, act: th.Tensor) -> bool:
    """True if the open weight-gradient lane will also produce this convolution's bias gradient (from the same launch:
    no column sums in the LeakyReLU / PixelNorm backward kernel, no reduce launch).  Only on the bf16 layers: there the
    launch sums exactly the values the column-sum kernel would; on the fp32 layers of the precise path it would sum
    their bf16 roundings (measured: the critic's bias updates then differ by 2e-3 from the eager step's)."""
    lane = ops.wgrad_lane()
    return lane is not None and act.dtype == th.bfloat16 and lane.accepts_bias(w, b)


def _wgrad(w, g, x, upsample_in: bool = False, bias=None, scale=None):
    """Weight gradient of a block: handed to the step's weight-gradient lane when one is open and takes this parameter
    (ops.WgradLane: computed on a second stream, collected by the step; autograd sees None), else computed in line.
    `bias`: only after _lane_bias(w, bias) said yes.  `scale`: the layer ran with the weight w * scale (0-dim tensor)."""
    lane = ops.wgrad_lane()
    if lane is not None and lane.accepts(w):
        lane.submit(w, g, x, upsample_in, bias=bias, scale=scale)
        return None
    assert bias is None
    if upsample_in:
        gw = ops.conv3x3_wgrad(g, x, upsample_in=True)
    else:
        gw = ConvWgrad.apply(g, x)
    return gw if scale is None else gw * scale


def _scaled(w, scale):
    """The weight a layer runs with: w, or w * scale for the output-scaled layers of the critic's fade-in (progan.py)."""
    w = w.float()
    return (w if scale is None else w * scale).contiguous()


# The gradient penalty's double backward reaches the FORWARD graph of the critic only through the saved activations of
# the mask kernels (LReLUBwd / UnpoolLReLUBwd / the 1x1 maps), whose gradient is None -- LeakyReLU masks are piecewise
# constant.  The autograd engine still runs every forward node it can reach, and a Function that lets torch materialise
# its missing output gradient is then called with ZEROS: the whole critic backward (18 data gradients, 18 weight gradients
# + their reduce kernels and casts, the mask / un-pool / 1x1 kernels) ran once per critic step on zeros.  Every Function
# here therefore switches materialisation off and answers a None gradient with Nones, without a launch.
def _nones(ctx):
    return (None,) * len(ctx.needs_input_grad)


class ConvFprop(Function):
    """y = conv3x3(x, w [* scale]), no bias / activation (linear in x and in w)."""

    @staticmethod
    def forward(ctx, x, w, scale=None):
        ctx.set_materialize_grads(False)
        ctx.save_for_backward(x, w, scale)
        return ops.conv3x3(_act(x), _scaled(w, scale))

    @staticmethod
    def backward(ctx, gy):
        if gy is None:
            return _nones(ctx)
        x, w, scale = ctx.saved_tensors
        gx = ConvDgrad.apply(gy, w, scale) if ctx.needs_input_grad[0] else None
        gw = _wgrad(w, gy, x, scale=scale) if ctx.needs_input_grad[1] and _param_grads[0] else None
        return gx, gw, None


class ConvDgrad(Function):
    """dx = data gradient of conv3x3(., w [* scale]) for output gradient g."""

    @staticmethod
    def forward(ctx, g, w, scale=None):
        ctx.set_materialize_grads(False)
        ctx.save_for_backward(g, w, scale)
        return ops.conv3x3(_act(g), _scaled(w, scale), dgrad=True)

    @staticmethod
    def backward(ctx, gdx):
        if gdx is None:
            return _nones(ctx)
        g, w, scale = ctx.saved_tensors
        gg = ConvFprop.apply(gdx, w, scale) if ctx.needs_input_grad[0] else None
        gw = _wgrad(w, g, gdx, scale=scale) if ctx.needs_input_grad[1] and _param_grads[0] else None
        return gg, gw, None


class ConvWgrad(Function):
    """dw = weight gradient for output gradient g and input x (fp32, shape (Cout, Cin, 3, 3))."""

    @staticmethod
    def forward(ctx, g, x):
        ctx.set_materialize_grads(False)
        ctx.save_for_backward(g, x)
        return ops.conv3x3_wgrad(_act(g), _act(x))

    @staticmethod
    def backward(ctx, gdw):
        if gdw is None:
            return _nones(ctx)
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
    Reference: discriminator.py:15-22,26-33 (Conv2d + LeakyReLU pairs of ConvBlock).
    `scale` (0-dim tensor >= 0, no gradient): the layer runs with w * scale and b * scale, i.e. y = scale * LeakyReLU(conv(x, w)
    + b) -- LeakyReLU is positively homogeneous -- which is how the critic's fade-in weights its two paths without a pass
    over the blended activations (progan.Discriminator.forward)."""

    @staticmethod
    def forward(ctx, x, w, b, scale=None):
        ctx.set_materialize_grads(False)
        y = ops.conv3x3(_act(x), _scaled(w, scale), _scaled(b, scale), lrelu=True, split_w=True, exact_w=True)
        ctx.save_for_backward(x, w, y, b, scale)
        return y

    @staticmethod
    def backward(ctx, gy):
        if gy is None:
            return _nones(ctx)
        x, w, y, b, scale = ctx.saved_tensors
        want_gw = ctx.needs_input_grad[1] and _param_grads[0]
        want_gb = ctx.needs_input_grad[2] and _param_grads[0]
        lane_b = want_gw and want_gb and scale is None and _lane_bias(w, b, y)
        if _UNFUSED_LRELU_BWD:
            gz = gy * _lrelu_mask(y)
            gb = gz.float().sum(dim=(0, 2, 3))
            lane_b = False
        else:
            gz, gb = LReLUBwd.apply(gy, y, want_gb and not lane_b)
        if gb is not None and scale is not None:
            gb = gb * scale
        gx = ConvDgrad.apply(gz, w, scale) if ctx.needs_input_grad[0] else None
        gw = _wgrad(w, gz, x, bias=b if lane_b else None, scale=scale) if want_gw else None
        return gx, gw, (gb if ctx.needs_input_grad[2] and not lane_b else None), None


class LReLUBwd(Function):
    """(gz, gb) = (gy * mask(y), sum_pixels gz): one kernel; linear in gy, so its own backward is the same mask
    multiply (plus the broadcast of the bias-gradient cotangent)."""

    @staticmethod
    def forward(ctx, gy, y, want_bias_grad=True):
        ctx.save_for_backward(y)
        ctx.set_materialize_grads(False)       # an unused bias-gradient output arrives as None, not as zeros
        gz, gb = ops.lrelu_bwd(gy, y, want_bias_grad=bool(want_bias_grad))
        return gz, gb

    @staticmethod
    def backward(ctx, ggz, ggb):
        (y,) = ctx.saved_tensors
        if ggz is None and ggb is None:
            return None, None, None
        if ggb is None:                        # the usual double-backward case: the same masked multiply, same kernel
            return LReLUBwd.apply(ggz, y, False)[0], None, None
        g = ggb.float()[None, :, None, None].expand(y.shape)
        if ggz is not None:
            g = g + ggz.float()
        return (ops.as_act(g, y.dtype) * _lrelu_mask(y)), None, None


class UnpoolLReLUBwd(Function):
    """(gz, gb) = (0.25 * up2(gp) * mask(h), sum_pixels gz): the backward of LeakyReLU -> AvgPool2d(2, 2) in one kernel.
    Linear in gp; its own backward is the masked multiply followed by the pooling (existing kernels)."""

    @staticmethod
    def forward(ctx, gp, h, want_bias_grad=True):
        ctx.save_for_backward(h)
        ctx.set_materialize_grads(False)
        return ops.unpool_lrelu_bwd(gp, h, want_bias_grad=bool(want_bias_grad))

    @staticmethod
    def backward(ctx, ggz, ggb):
        (h,) = ctx.saved_tensors
        if ggz is None and ggb is None:
            return None, None, None
        if ggb is None:
            return Pool2.apply(LReLUBwd.apply(ggz, h, False)[0]), None, None
        g = ggb.float()[None, :, None, None].expand(h.shape)
        if ggz is not None:
            g = g + ggz.float()
        return Pool2.apply(ops.as_act(g, h.dtype) * _lrelu_mask(h)), None, None


class ConvBiasLReLUPool(Function):
    """p = AvgPool2d(2,2)(LeakyReLU_0.2(conv3x3(x, w) + b)): first half of the discriminator's ConvBlock
    (discriminator.py:15-24).  Forward = the fused conv kernel + the pooling kernel; backward starts with ONE kernel for
    un-pooling, LeakyReLU mask and bias gradient (UnpoolLReLUBwd), then ConvDgrad / ConvWgrad."""

    @staticmethod
    def forward(ctx, x, w, b):
        ctx.set_materialize_grads(False)
        h = ops.conv3x3(_act(x), w.float().contiguous(), b.float().contiguous(), lrelu=True, split_w=True, exact_w=True)
        ctx.save_for_backward(x, w, h, b)
        return ops.pool2(h)

    @staticmethod
    def backward(ctx, gp):
        if gp is None:
            return _nones(ctx)
        x, w, h, b = ctx.saved_tensors
        want_gw = ctx.needs_input_grad[1] and _param_grads[0]
        want_gb = ctx.needs_input_grad[2] and _param_grads[0]
        lane_b = want_gw and want_gb and _lane_bias(w, b, h)
        gz, gb = UnpoolLReLUBwd.apply(gp, h, want_gb and not lane_b)
        gx = ConvDgrad.apply(gz, w) if ctx.needs_input_grad[0] else None
        gw = _wgrad(w, gz, x, bias=b if lane_b else None) if want_gw else None
        return gx, gw, (gb if ctx.needs_input_grad[2] and not lane_b else None)


class GenConv(Function):
    """One generator half-block: [nearest x2 upsample ->] conv3x3 + bias -> LeakyReLU(0.2) -> PixelNorm, all in one
    kernel (generator.py:16-24 / :26-40, layers.py:11-17).  First-order backward only (the generator is never
    differentiated twice: the gradient penalty graph lives in the discriminator)."""

    @staticmethod
    def forward(ctx, x, w, b, upsample_in: bool):
        cout, cin = w.shape[0], w.shape[1]
        xa = _act(x)
        # PixelNorm needs all Cout of a pixel in one CTA: the precise kernel streams its weights (Cout <= 128), the
        # bf16 kernel keeps them (hi + lo) resident in shared memory
        fused_pn = cout <= 128 if xa.dtype == th.float32 else 9 * cin * cout * 4 <= 184 * 1024
        wf, bf = w.float().contiguous(), b.float().contiguous()
        if fused_pn:
            o, inv = ops.conv3x3(xa, wf, bf, lrelu=True, pixelnorm=True, upsample_in=upsample_in, want_inv_norm=True, split_w=True)
        else:                # not reached by the reference's channel plan (generator.py:67-76)
            t = ops.conv3x3(xa, wf, bf, lrelu=True, upsample_in=upsample_in, split_w=True).float()
            inv = th.rsqrt(t.pow(2).mean(dim=1) + PN_EPS)
            o = ops.as_act(t * inv[:, None], xa.dtype)
        ctx.save_for_backward(xa, w, o, inv, b)
        ctx.upsample_in = upsample_in
        return o

    @staticmethod
    @th.autograd.function.once_differentiable
    def backward(ctx, go):
        xa, w, o, inv, b = ctx.saved_tensors
        # PixelNorm + LeakyReLU backward and the bias gradient in one kernel:
        #   g_t = (g_o - o * mean_c(g_o * o)) / n ;  g_z = g_t * mask(o) ;  g_b = sum_pixels g_z
        lane_b = ctx.needs_input_grad[1] and ctx.needs_input_grad[2] and _lane_bias(w, b, o)
        gz, gb = ops.pixelnorm_lrelu_bwd(go, o, inv, want_bias_grad=ctx.needs_input_grad[2] and not lane_b)
        gx = gw = None
        wf = w.float().contiguous()
        if ctx.needs_input_grad[0]:
            gx = ops.conv3x3(gz, wf, dgrad=True)
            if ctx.upsample_in:      # backward of the nearest upsampling folded into the read: sum of each 2x2 block
                gx = ops.pool2(gx, sum_pool=True)
        if ctx.needs_input_grad[1]:
            gw = _wgrad(w, gz, xa, ctx.upsample_in, bias=b if lane_b else None)
        return gx, gw, gb, None


class RgbExpand(Function):
    """y = (W x + b) -> LeakyReLU(0.2) or mask multiply; x (B,2,H,W) fp32, w (C,2,1,1)|(C,2), y bf16 NHWC.
    One of the three mutually-differentiating 1x1 maps (RgbExpand / RgbProject / RgbWgrad)."""

    @staticmethod
    def forward(ctx, x, w, b, mask_src, lrelu: bool, out_dtype=th.bfloat16):
        ctx.set_materialize_grads(False)
        y = ops.rgb_expand(x, w, b, mask_src=mask_src, lrelu=lrelu, out_dtype=out_dtype)
        ctx.save_for_backward(x, w, y if lrelu else mask_src)
        ctx.has_mask = lrelu or mask_src is not None
        ctx.w_shape = w.shape
        return y

    @staticmethod
    def backward(ctx, gy):
        if gy is None:
            return _nones(ctx)
        x, w, m = ctx.saved_tensors
        m = m if ctx.has_mask else None
        gx = RgbProject.apply(gy, w, m) if ctx.needs_input_grad[0] else None
        gw = gb = None
        if (ctx.needs_input_grad[1] or ctx.needs_input_grad[2]) and _param_grads[0]:
            gw, gb = RgbWgrad.apply(gy, m, x)
            gw = gw.reshape(ctx.w_shape)
        return gx, gw, gb, None, None, None


class RgbProject(Function):
    """out = W^T (a * mask): a bf16 NHWC (B,C,H,W), w (C,2,...) -> out fp32 (B,2,H,W)."""

    @staticmethod
    def forward(ctx, a, w, mask_src):
        ctx.set_materialize_grads(False)
        ctx.save_for_backward(a, w, mask_src)
        ctx.w_shape = w.shape
        return ops.rgb_project(_act(a), w, mask_src=mask_src, w_is_c_by_2=True)

    @staticmethod
    def backward(ctx, gout):
        if gout is None:
            return _nones(ctx)
        a, w, m = ctx.saved_tensors
        ga = RgbExpand.apply(gout, w, None, m, False, a.dtype) if ctx.needs_input_grad[0] else None
        gw = None
        if ctx.needs_input_grad[1] and _param_grads[0]:
            gw, _ = RgbWgrad.apply(a, m, gout)
            gw = gw.reshape(ctx.w_shape)
        return ga, gw, None


class RgbWgrad(Function):
    """(gw, gb) = (sum_p (g*mask)[p,c] x[k,p], sum_p (g*mask)[p,c])."""

    @staticmethod
    def forward(ctx, g, mask_src, x):
        ctx.save_for_backward(g, mask_src, x)
        return ops.rgb_wgrad(_act(g), mask_src, x)

    @staticmethod
    def backward(ctx, ggw, ggb):
        g, m, x = ctx.saved_tensors
        gg = RgbExpand.apply(x, ggw, ggb, m, False, g.dtype) if ctx.needs_input_grad[0] else None
        gx = RgbProject.apply(g, ggw, m) if ctx.needs_input_grad[2] else None
        return gg, None, gx


class ToRgbTanh(Function):
    """out = tanh(W a + b): a bf16 NHWC, W (2,C,1,1) -> fp32 (B,2,H,W)  (generator.py:46-51).  First order only."""

    @staticmethod
    def forward(ctx, a, w, b):
        aa = _act(a)
        out = ops.rgb_project(aa, w, bias=b, tanh=True)
        ctx.save_for_backward(aa, w, out)
        return out

    @staticmethod
    @th.autograd.function.once_differentiable
    def backward(ctx, gout):
        aa, w, out = ctx.saved_tensors
        gpre = (gout.float() * (1.0 - out * out)).contiguous()
        ga = gw = gb = None
        C = aa.shape[1]
        if ctx.needs_input_grad[0]:
            ga = ops.rgb_expand(gpre, w.float().reshape(2, C).t().contiguous(), out_dtype=aa.dtype)
        if ctx.needs_input_grad[1]:
            gwt, _ = ops.rgb_wgrad(aa, None, gpre)          # (C, 2)
            gw = gwt.t().reshape(w.shape).contiguous()
        if ctx.needs_input_grad[2]:
            gb = gpre.sum(dim=(0, 2, 3))
        return ga, gw, gb


class Pool2(Function):
    """AvgPool2d(2, 2) on bf16 NHWC; its backward is the adjoint kernel, whose backward is the pooling again."""

    @staticmethod
    def forward(ctx, x):
        ctx.set_materialize_grads(False)
        return ops.pool2(_act(x))

    @staticmethod
    def backward(ctx, g):
        if g is None:
            return _nones(ctx)
        return Unpool2.apply(g)


class Unpool2(Function):
    @staticmethod
    def forward(ctx, g):
        ctx.set_materialize_grads(False)
        return ops.pool2(_act(g), adjoint=True)

    @staticmethod
    def backward(ctx, gg):
        if gg is None:
            return _nones(ctx)
        return Pool2.apply(gg)


class PoolPlanes(Function):
    """AvgPool2d(2, 2) on the fp32 NCHW network input (fade-in path of the discriminator); backward = UnpoolPlanes."""

    @staticmethod
    def forward(ctx, x):
        ctx.set_materialize_grads(False)
        return ops.pool2_planes(x)

    @staticmethod
    def backward(ctx, g):
        if g is None:
            return _nones(ctx)
        return UnpoolPlanes.apply(g)


class UnpoolPlanes(Function):
    @staticmethod
    def forward(ctx, g):
        ctx.set_materialize_grads(False)
        return ops.pool2_planes(g, adjoint=True)

    @staticmethod
    def backward(ctx, gg):
        if gg is None:
            return _nones(ctx)
        return PoolPlanes.apply(gg)
