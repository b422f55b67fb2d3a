"""Drop-in ``Generator`` / ``Discriminator`` (progressive-growing GAN) whose 3x3 convolutions run on the
tcgen05 kernels of libmusicgan_b200.so.

Public surface, constructor arguments, properties, growth behaviour and -- importantly -- the state-dict keys
(``_Generator__gen_blocks.3.4.weight`` ...) are those of reference networks/generator.py:55-171 and
networks/discriminator.py:53-191, so checkpoints move both ways.  Parameters are fp32 masters with
nn.Conv2d's default initialisation; activations are bf16 NHWC between layers; network inputs and outputs are
fp32 NCHW like the reference's.  CUDA only: there is no CPU forward.
"""
from __future__ import annotations

from typing import Iterator, List, Tuple

import torch as th
import torch.nn as nn
import torch.nn.functional as F

from . import functional as fn

G_CHANNELS: List[Tuple[int, int]] = [(0, 128), (128, 112), (112, 96), (96, 80), (80, 64), (64, 48), (48, 32), (32, 16)]
D_CHANNELS: List[Tuple[int, int]] = [(16, 32), (32, 48), (48, 64), (64, 80), (80, 96), (96, 112), (112, 128), (128, 144), (144, 160)]


def _conv3(cin: int, cout: int) -> nn.Conv2d:
    return nn.Conv2d(cin, cout, kernel_size=(3, 3), stride=(1, 1), padding=(1, 1))


def _blend(old: th.Tensor, new: th.Tensor, alpha) -> th.Tensor:
    """alpha * new + (1 - alpha) * old (generator.py:124, discriminator.py:113) as ONE kernel: old + alpha * (new - old).
    `alpha` is the reference's Python float or a 0-dim device tensor (graphed.py feeds it to captured graphs that way)."""
    if th.is_tensor(alpha):
        alpha = alpha.to(dtype=old.dtype)
    elif old.dtype == th.bfloat16:
        # a tensor weight of a bf16 blend is itself bf16: round the float the same way, so that the eager steps and the
        # captured graphs (alpha on the device) compute the same blend (2^-9 relative on alpha, below the rounding of the result)
        alpha = float(th.tensor(float(alpha)).bfloat16())
    return th.lerp(old, new.to(old.dtype), alpha)


def _require_cuda(t: th.Tensor, who: str) -> None:
    if not t.is_cuda:
        raise RuntimeError(f"{who}: musicgan_b200 networks run on a CUDA (sm_100a) device only -- call .cuda() "
                           "on the module and its inputs (no CPU fallback)")


class PixelNorm(nn.Module):
    """x / sqrt(mean_c(x^2) + eps) (layers.py:5-23).  Kept as a parameter-free module so that the Sequential
    indices (and therefore the state-dict keys) of a generator block match the reference; inside the blocks the
    normalisation itself is fused into the convolution epilogue."""

    def __init__(self, epsilon: float = 1e-8):
        super().__init__()
        self.epsilon = epsilon

    def forward(self, x: th.Tensor) -> th.Tensor:
        return x / th.sqrt(x.pow(2.).mean(dim=1, keepdim=True) + self.epsilon)

    def __repr__(self):
        return f"PixelNorm(eps={self.epsilon})"


class Block(nn.Sequential):
    """Generator block (generator.py:9-40): conv(cin,cin) LReLU PN | up x2 | conv(cin,cout) LReLU PN,
    executed as TWO fused kernels (the upsampled tensor is never materialised)."""

    def __init__(self, in_channels: int, out_channels: int):
        super().__init__(
            _conv3(in_channels, in_channels), nn.LeakyReLU(2e-1), PixelNorm(),
            nn.Upsample(scale_factor=2., mode="nearest"),
            _conv3(in_channels, out_channels), nn.LeakyReLU(2e-1), PixelNorm())

    def forward(self, x: th.Tensor) -> th.Tensor:
        h = fn.GenConv.apply(fn.prep(x, x.shape[2]), self[0].weight, self[0].bias, False)
        return fn.GenConv.apply(fn.prep(h, 2 * h.shape[2]), self[4].weight, self[4].bias, True)


class ToMagnPhaseLayer(nn.Sequential):
    """1x1 conv C -> 2 + tanh (generator.py:43-52); returns fp32."""

    def __init__(self, in_channels: int):
        super().__init__(nn.Conv2d(in_channels, 2, kernel_size=(1, 1), stride=(1, 1)), nn.Tanh())

    def forward(self, x: th.Tensor) -> th.Tensor:
        return fn.ToRgbTanh.apply(fn.prep(x, x.shape[2]), self[0].weight, self[0].bias)


class _UpsampledToMagnPhase(nn.Sequential):
    """(old end block, nearest x2): generator.py:132-138 -- index 0 holds the previous ToMagnPhaseLayer."""

    def forward(self, x: th.Tensor) -> th.Tensor:
        return F.interpolate(self[0](x), scale_factor=2., mode="nearest")


class Generator(nn.Module):
    def __init__(self, rand_channels: int, end_layer: int = 0):
        super().__init__()
        chans = [(rand_channels, G_CHANNELS[0][1])] + G_CHANNELS[1:]
        assert 0 <= end_layer < len(chans), f"0 <= {end_layer} < {len(chans)}"
        if rand_channels % 16 != 0:
            raise NotImplementedError("the tcgen05 convolution kernels need channel counts that are multiples of 16 "
                                      f"(rand_channels={rand_channels}; the reference trains with 32)")
        self.__curr_layer = end_layer
        self.__nb_downsample = 7
        self.__channels = chans
        self.__gen_blocks = nn.ModuleList([Block(ci, co) for ci, co in chans])
        self.__end_block = ToMagnPhaseLayer(chans[end_layer][1])
        self.__last_end_block = None if end_layer == 0 else _UpsampledToMagnPhase(
            ToMagnPhaseLayer(chans[end_layer - 1][1]), nn.Upsample(scale_factor=2., mode="nearest"))

    def forward(self, z: th.Tensor, alpha: float) -> th.Tensor:
        _require_cuda(z, "Generator.forward")
        with fn.ops.forward_only(not self.training and not th.is_grad_enabled()):      # generate.py:38,54: eval(), no graph
            return self._forward(z, alpha)

    def _forward(self, z: th.Tensor, alpha: float) -> th.Tensor:
        out = z
        for i in range(self.curr_layer):
            out = self.__gen_blocks[i](out)
        top = self.__gen_blocks[self.curr_layer](out)
        new_mp = self.__end_block(top)
        if self.__last_end_block is None:
            return new_mp
        # alpha * new + (1 - alpha) * old (generator.py fade-in) as ONE kernel: old + alpha * (new - old)
        return _blend(self.__last_end_block(out), new_mp, alpha)

    def next_layer(self) -> bool:
        if not self.growing:
            return False
        self.__curr_layer += 1
        self.__last_end_block = _UpsampledToMagnPhase(self.__end_block, nn.Upsample(scale_factor=2., mode="nearest"))
        self.__end_block = ToMagnPhaseLayer(self.__channels[self.curr_layer][1])
        self.__end_block.to(next(self.__gen_blocks.parameters()).device)
        return True

    @property
    def down_sample(self) -> int:
        return self.__nb_downsample

    @property
    def curr_layer(self) -> int:
        return self.__curr_layer

    @property
    def growing(self) -> bool:
        return self.curr_layer < len(self.__gen_blocks) - 1

    def end_block_params(self) -> Iterator[nn.Parameter]:
        return self.__end_block.parameters()

    def zero_grad(self, set_to_none: bool = False) -> None:
        for p in self.parameters():       # generator.py:169-171: always drops the gradients
            p.grad = None


class ConvBlock(nn.Sequential):
    """Discriminator block (discriminator.py:8-34): conv(cin,cout) LReLU | avgpool 2 | conv(cout,cout) LReLU;
    each conv + bias + LeakyReLU is one kernel."""

    def __init__(self, in_channels: int, out_channels: int):
        super().__init__(_conv3(in_channels, out_channels), nn.LeakyReLU(2e-1), nn.AvgPool2d(2, 2),
                         _conv3(out_channels, out_channels), nn.LeakyReLU(2e-1))

    def forward(self, x: th.Tensor, out_scale=None) -> th.Tensor:
        """`out_scale` (0-dim tensor >= 0): the block's output times that factor, folded into the second convolution's weight
        and bias (LeakyReLU is positively homogeneous) -- the fade-in weight of the critic's new path."""
        h = fn.ConvBiasLReLUPool.apply(fn.prep(x, x.shape[2]), self[0].weight, self[0].bias)      # conv + LReLU, pool; fused backward
        return fn.ConvBiasLReLU.apply(fn.prep(h, h.shape[2]), self[3].weight, self[3].bias, out_scale)


class MagPhaseLayer(nn.Sequential):
    """1x1 conv 2 -> C + LeakyReLU (discriminator.py:37-50); bf16 NHWC out."""

    def __init__(self, out_channels: int):
        super().__init__(nn.Conv2d(2, out_channels, kernel_size=(1, 1), stride=(1, 1)), nn.LeakyReLU(2e-1))

    def forward(self, x: th.Tensor, out_scale=None) -> th.Tensor:
        w, b = self[0].weight, self[0].bias
        if out_scale is not None:          # scale * LeakyReLU(W x + b) = LeakyReLU((scale W) x + scale b) for scale >= 0
            w, b = w * out_scale, b * out_scale
        return fn.RgbExpand.apply(x, w, b, None, True, th.float32 if fn.ops.is_precise(x.shape[2]) else th.bfloat16)


class _PooledMagPhase(nn.Sequential):
    """(avgpool 2, old start block): discriminator.py:130-133 -- index 1 holds the previous MagPhaseLayer."""

    def forward(self, x: th.Tensor, out_scale=None) -> th.Tensor:
        return self[1](fn.PoolPlanes.apply(x), out_scale)      # == F.avg_pool2d(x, 2, 2) bit for bit; torch's backward kernel is 10x slower


class Discriminator(nn.Module):
    def __init__(self, start_layer: int = 7):
        super().__init__()
        self.__channels = D_CHANNELS
        self.__curr_layer = start_layer
        self.__nb_layer = len(D_CHANNELS)
        assert 0 <= start_layer <= len(D_CHANNELS)
        self.__conv_blocks = nn.ModuleList([ConvBlock(ci, co) for ci, co in D_CHANNELS])
        self.__last_start_block = None
        self.__start_block = MagPhaseLayer(D_CHANNELS[self.curr_layer][0])
        out_size = D_CHANNELS[-1][1] * 512 // 2 ** self.__nb_layer * 512 // 2 ** self.__nb_layer      # = 160
        self.__clf = nn.Sequential(nn.Linear(out_size, 1))

    def forward(self, x: th.Tensor, alpha: float) -> th.Tensor:
        _require_cuda(x, "Discriminator.forward")
        if self.__last_start_block is None:
            out = self.__conv_blocks[self.__curr_layer](self.__start_block(x))
        else:
            # alpha * new + (1 - alpha) * old (discriminator.py:113) with both factors folded into the LAST layer of each
            # path (LeakyReLU(s z) = s LeakyReLU(z) for s >= 0): the blend is a plain sum, and its backward hands the same
            # gradient to both paths -- no multiply passes over the C x H/2 x W/2 activations and their gradients (six
            # 45 us kernels per critic step at 512 x 512 / batch 8)
            a = alpha if th.is_tensor(alpha) else th.full((), float(alpha), dtype=th.float32, device=x.device)
            a = a.to(dtype=th.float32)
            new = self.__conv_blocks[self.__curr_layer](self.__start_block(x), out_scale=a)
            old = self.__last_start_block(x, out_scale=1.0 - a)
            out = old + new.to(old.dtype)
        for i in range(self.__curr_layer + 1, len(self.__conv_blocks)):
            out = self.__conv_blocks[i](out)
        return self.__clf(out.flatten(1, -1).float())

    def next_layer(self) -> bool:
        if not self.growing:
            return False
        self.__curr_layer -= 1
        self.__last_start_block = _PooledMagPhase(nn.AvgPool2d(2, 2), self.__start_block)
        self.__start_block = MagPhaseLayer(self.__channels[self.curr_layer][0])
        self.__start_block.to(next(self.__conv_blocks.parameters()).device)
        return True

    @property
    def curr_layer(self) -> int:
        return self.__curr_layer

    @property
    def growing(self) -> bool:
        return self.__curr_layer > 0

    def gradient_penalty(self, x_real: th.Tensor, x_gen: th.Tensor, alpha: float, eps: th.Tensor = None) -> th.Tensor:
        """WGAN-GP term, discriminator.py:157-184 (lambda = 10).  Unlike the reference this also works when
        neither input carries a gradient (a detached fake batch): the interpolate is marked as requiring grad."""
        batch_size = x_real.size()[0]
        if eps is None:       # extra keyword (not in the reference) so tests can inject the uniform sample
            eps = th.rand(batch_size, 1, 1, 1, device=x_real.device)
        x_interpolated = eps * x_real + (1 - eps) * x_gen
        if not x_interpolated.requires_grad:
            x_interpolated.requires_grad_(True)
        out_interpolated = self(x_interpolated, alpha)
        with fn.input_grads_only():      # this call wants d out / d x_hat only: no weight / bias gradient kernels
            gradients = th.autograd.grad(out_interpolated, x_interpolated,
                                         grad_outputs=th.ones(out_interpolated.size(), device=x_real.device),
                                         create_graph=True, retain_graph=True)
        gradients = gradients[0].reshape(batch_size, -1)
        gradients_norm = gradients.norm(2, dim=1)
        return 10. * ((gradients_norm - 1.) ** 2.).mean()

    def start_block_parameters(self) -> Iterator[nn.Parameter]:
        return self.__start_block.parameters()

    def zero_grad(self, set_to_none: bool = False) -> None:
        for p in self.parameters():
            p.grad = None
