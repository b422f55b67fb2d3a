"""The bar on the same box: the reference's hot path run through STOCK torch on the B200 (cuDNN convolutions, cuFFT,
ATen elementwise kernels) -- what a user gets today by calling `.cuda()` on the reference modules
(reference train.py:61-62).  COMPARISON ARM ONLY: nothing in musicgan_b200/ imports this file; it launches none of this
repo's kernels.  bench.py times it beside the product (`--impl torch_cuda`, and the `torch_b200` key of the default line).

The reference tree itself cannot travel to the GPU box, so its modules are restated here with the same torch layers in
the same order: Block generator.py:9-40, ToMagnPhaseLayer :43-52, Generator.forward :106-126; ConvBlock
discriminator.py:8-34, MagPhaseLayer :37-50, Discriminator.forward :107-124, gradient_penalty :157-184; step body
train.py:143-214 (fake batch NOT detached, critic gradients computed and dropped in the generator step -- exactly the
work the reference does); transform audio/functions.py:38-94 with the cumsum spelled in double as SURVEY B.2 prescribes
for CUDA.
"""
from __future__ import annotations

import math
import time

import torch as th
import torch.nn as nn
import torch.nn.functional as F

G_CH = [(32, 128), (128, 112), (112, 96), (96, 80), (80, 64), (64, 48), (48, 32), (32, 16)]
D_CH = [(16, 32), (32, 48), (48, 64), (64, 80), (80, 96), (96, 112), (112, 128), (128, 144), (144, 160)]


class PixelNorm(nn.Module):
    def forward(self, x):
        return x / th.sqrt(x.pow(2.).mean(dim=1, keepdim=True) + 1e-8)


def g_block(ci, co):
    return nn.Sequential(nn.Conv2d(ci, ci, 3, 1, 1), nn.LeakyReLU(0.2), PixelNorm(), nn.Upsample(scale_factor=2., mode="nearest"),
                         nn.Conv2d(ci, co, 3, 1, 1), nn.LeakyReLU(0.2), PixelNorm())


def d_block(ci, co):
    return nn.Sequential(nn.Conv2d(ci, co, 3, 1, 1), nn.LeakyReLU(0.2), nn.AvgPool2d(2, 2), nn.Conv2d(co, co, 3, 1, 1), nn.LeakyReLU(0.2))


class Generator(nn.Module):
    def __init__(self, stage: int):
        super().__init__()
        self.stage = stage
        self.blocks = nn.ModuleList([g_block(ci, co) for ci, co in G_CH])
        self.end = nn.Sequential(nn.Conv2d(G_CH[stage][1], 2, 1), nn.Tanh())
        self.last_end = nn.Sequential(nn.Conv2d(G_CH[stage - 1][1], 2, 1), nn.Tanh(), nn.Upsample(scale_factor=2., mode="nearest")) if stage else None

    def forward(self, z, alpha):
        out = z
        for i in range(self.stage):
            out = self.blocks[i](out)
        new = self.end(self.blocks[self.stage](out))
        if self.last_end is None:
            return new
        return alpha * new + (1. - alpha) * self.last_end(out)


class Discriminator(nn.Module):
    def __init__(self, stage: int):
        super().__init__()
        self.cur = 7 - stage
        self.blocks = nn.ModuleList([d_block(ci, co) for ci, co in D_CH])
        self.start = nn.Sequential(nn.Conv2d(2, D_CH[self.cur][0], 1), nn.LeakyReLU(0.2))
        self.last_start = nn.Sequential(nn.AvgPool2d(2, 2), nn.Conv2d(2, D_CH[self.cur + 1][0], 1), nn.LeakyReLU(0.2)) if stage else None
        self.clf = nn.Linear(160, 1)

    def forward(self, x, alpha):
        out = self.blocks[self.cur](self.start(x))
        if self.last_start is not None:
            out = alpha * out + (1. - alpha) * self.last_start(x)
        for i in range(self.cur + 1, len(self.blocks)):
            out = self.blocks[i](out)
        return self.clf(out.flatten(1, -1))

    def gradient_penalty(self, x_real, x_gen, alpha):
        eps = th.rand(x_real.size(0), 1, 1, 1, device=x_real.device)
        x_hat = eps * x_real + (1 - eps) * x_gen
        out = self(x_hat, alpha)
        (g,) = th.autograd.grad(out, x_hat, grad_outputs=th.ones_like(out), create_graph=True, retain_graph=True)
        return 10. * ((g.reshape(g.size(0), -1).float().norm(2, dim=1) - 1.) ** 2.).mean()


class Trainer:
    """One reference iteration: critic step every call, generator step every 5th (train.py:189)."""

    def __init__(self, stage: int, batch: int, mode: str, device="cuda"):
        th.manual_seed(0)
        self.gen, self.disc = Generator(stage).to(device), Discriminator(stage).to(device)
        self.mode, self.batch, self.alpha, self.dev = mode, batch, 0.5, device
        if mode == "bf16_autocast_channels_last":
            self.gen.to(memory_format=th.channels_last); self.disc.to(memory_format=th.channels_last)
        self.og = th.optim.Adam(self.gen.parameters(), lr=1e-3, betas=(0., 0.9))
        self.od = th.optim.Adam(self.disc.parameters(), lr=1e-3, betas=(0., 0.9))
        self.it = 0

    def _ctx(self):
        return th.autocast("cuda", dtype=th.bfloat16) if self.mode == "bf16_autocast_channels_last" else th.autocast("cuda", enabled=False)

    def iteration(self, x_real):
        gen, disc, a = self.gen, self.disc, self.alpha
        if self.mode == "bf16_autocast_channels_last":
            x_real = x_real.contiguous(memory_format=th.channels_last)
        z = th.randn(self.batch, 32, 2, 2, device=self.dev)
        with self._ctx():
            x_fake = gen(z, a)
            loss = -(disc(x_real, a).mean() - disc(x_fake, a).mean())
            gp = disc.gradient_penalty(x_real, x_fake, a)
        gen.zero_grad(set_to_none=True); disc.zero_grad(set_to_none=True)
        (loss + gp).backward()
        self.od.step()
        out = loss.detach() + gp.detach()
        if self.it % 5 == 0:
            z = th.randn(self.batch, 32, 2, 2, device=self.dev)
            with self._ctx():
                g_loss = -disc(gen(z, a), a).mean()
            gen.zero_grad(set_to_none=True); disc.zero_grad(set_to_none=True)
            g_loss.backward()
            self.og.step()
        self.it += 1
        return out


def time_train(stage: int, batch: int, mode: str, steps: int, warmup: int, device="cuda"):
    """steps/s of the stock-torch iteration (CUDA events, inputs resident)."""
    res = 4 * 2 ** stage
    tr = Trainer(stage, batch, mode, device)
    reals = [th.rand(batch, 2, res, res, device=device) * 2 - 1 for _ in range(4)]
    for i in range(warmup):
        tr.iteration(reals[i % 4])
    tr.it = 0
    th.cuda.synchronize()
    a, b = th.cuda.Event(enable_timing=True), th.cuda.Event(enable_timing=True)
    a.record()
    for i in range(steps):
        tr.iteration(reals[i % 4])
    b.record()
    th.cuda.synchronize()
    ms = a.elapsed_time(b)
    return steps / (ms * 1e-3), ms / steps


# ---- transform ------------------------------------------------------------------------------------------------------
def bark_gain(n_bins, device):
    scale = 6. * th.asinh(th.linspace(20., 44100 // 2, n_bins, device=device) / 600.)[:, None]
    return scale / th.norm(scale, p=2)


def transform(wav: th.Tensor, nb_vec: int = 512):
    """(B, N) mono clips -> magn, phase (B, n_chunks, 512, nb_vec): functions.py:38-94 batched over clips with torch-CUDA ops."""
    pi = math.pi
    win = th.hann_window(1024, device=wav.device)
    cv = th.stft(wav, 1024, 256, 1024, win, center=True, pad_mode="reflect", normalized=False, onesided=True, return_complex=True)
    cv = (cv / win.pow(2.).sum().sqrt())[:, :-1, :]
    magn, phase = th.abs(cv) * bark_gain(512, wav.device), th.angle(cv)
    dphi = F.pad(phase[:, :, 1:] - phase[:, :, :-1], (1, 0))
    dphi_m = ((dphi + pi) % (2 * pi)) - pi
    dphi_m[(dphi_m == -pi) & (dphi > 0)] = pi
    adj = dphi_m - dphi
    adj[dphi.abs() < pi] = 0
    phase = phase + adj.double().cumsum(2).float()
    phase = phase[:, :, 1:] - phase[:, :, :-1]
    magn = magn[:, :, 1:]
    mx, mn = magn.amax((1, 2), keepdim=True), magn.amin((1, 2), keepdim=True)
    magn = (magn - mn) / (mx - mn) * 2. - 1.
    mx, mn = phase.amax((1, 2), keepdim=True), phase.amin((1, 2), keepdim=True)
    phase = (phase - mn) / (mx - mn) * 2. - 1.
    head = magn.size(2) % nb_vec
    B = wav.size(0)
    to_chunks = lambda t: t[:, :, head:].reshape(B, 512, -1, nb_vec).permute(0, 2, 1, 3).contiguous()
    return to_chunks(magn), to_chunks(phase)


def time_transform(clips: int, n_samples: int, steps: int, warmup: int, device="cuda"):
    g = th.Generator().manual_seed(7)
    wav = ((th.rand(clips, n_samples, generator=g) * 2 - 1) * 0.5).to(device)
    for _ in range(warmup):
        transform(wav)
    th.cuda.synchronize()
    a, b = th.cuda.Event(enable_timing=True), th.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(steps):
        transform(wav)
    b.record()
    th.cuda.synchronize()
    ms = a.elapsed_time(b)
    frames = clips * (1 + n_samples // 256)
    return frames * steps / (ms * 1e-3), ms / steps
