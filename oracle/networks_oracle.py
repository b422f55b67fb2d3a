"""CPU fp32 restatement of the reference's ProGAN generator / discriminator and of the WGAN-GP step body
(TEST INFRASTRUCTURE ONLY).  Functional style: every function takes the state dict (reference key names) and
runs plain torch fp32 ops in the order the reference modules run them.

Pinned against the imported reference modules by oracle/gen_golden_networks.py -> tests/golden/networks_*.npz.
Citations: networks/generator.py (G), networks/discriminator.py (D), networks/layers.py, networks/criterion.py,
train.py:143-214 (step body).
"""
from __future__ import annotations

from typing import Dict, List, Tuple

import torch
import torch.nn.functional as F

G_CHANNELS = [(None, 128), (128, 112), (112, 96), (96, 80), (80, 64), (64, 48), (48, 32), (32, 16)]     # generator.py:67-76
D_CHANNELS = [(16, 32), (32, 48), (48, 64), (64, 80), (80, 96), (96, 112), (112, 128), (128, 144), (144, 160)]  # discriminator.py:60-70

SD = Dict[str, torch.Tensor]


def state_shapes(kind: str, stage: int, rand_channels: int = 32) -> List[Tuple[str, Tuple[int, ...]]]:
    """(key, shape) list of the reference module's state dict after `stage` calls of next_layer() starting from
    Generator(rand_channels, 0) / Discriminator(7), in registration order."""
    out: List[Tuple[str, Tuple[int, ...]]] = []
    if kind == "gen":
        chans = [(rand_channels, 128)] + G_CHANNELS[1:]
        for i, (ci, co) in enumerate(chans):
            out += [(f"_Generator__gen_blocks.{i}.0.weight", (ci, ci, 3, 3)), (f"_Generator__gen_blocks.{i}.0.bias", (ci,)),
                    (f"_Generator__gen_blocks.{i}.4.weight", (co, ci, 3, 3)), (f"_Generator__gen_blocks.{i}.4.bias", (co,))]
        out += [("_Generator__end_block.0.weight", (2, chans[stage][1], 1, 1)), ("_Generator__end_block.0.bias", (2,))]
        if stage >= 1:
            out += [("_Generator__last_end_block.0.0.weight", (2, chans[stage - 1][1], 1, 1)),
                    ("_Generator__last_end_block.0.0.bias", (2,))]
    else:
        for j, (ci, co) in enumerate(D_CHANNELS):
            out += [(f"_Discriminator__conv_blocks.{j}.0.weight", (co, ci, 3, 3)), (f"_Discriminator__conv_blocks.{j}.0.bias", (co,)),
                    (f"_Discriminator__conv_blocks.{j}.3.weight", (co, co, 3, 3)), (f"_Discriminator__conv_blocks.{j}.3.bias", (co,))]
        cur = 7 - stage
        out += [("_Discriminator__start_block.0.weight", (D_CHANNELS[cur][0], 2, 1, 1)),
                ("_Discriminator__start_block.0.bias", (D_CHANNELS[cur][0],)),
                ("_Discriminator__clf.0.weight", (1, 160)), ("_Discriminator__clf.0.bias", (1,))]
        if stage >= 1:
            out += [("_Discriminator__last_start_block.1.0.weight", (D_CHANNELS[cur + 1][0], 2, 1, 1)),
                    ("_Discriminator__last_start_block.1.0.bias", (D_CHANNELS[cur + 1][0],))]
    return out


def make_state(kind: str, stage: int, seed: int, rand_channels: int = 32) -> SD:
    """Seeded synthetic weights with nn.Conv2d-like scale U(-1/sqrt(fan_in), 1/sqrt(fan_in)); identical on
    every machine (torch CPU generator)."""
    g = torch.Generator().manual_seed(seed)
    sd = {}
    last_fan = 1
    for key, shape in state_shapes(kind, stage, rand_channels):
        if key.endswith("weight"):
            fan_in = 1
            for s in shape[1:]:
                fan_in *= s
            last_fan = fan_in
        bound = 1.0 / (last_fan ** 0.5)
        sd[key] = (torch.rand(shape, generator=g) * 2 - 1) * bound
    return sd


def pixel_norm(x):                                               # layers.py:11-17
    return x / torch.sqrt(x.pow(2.0).mean(dim=1, keepdim=True) + 1e-8)


def gen_forward(sd: SD, z: torch.Tensor, alpha: float, stage: int) -> torch.Tensor:
    """generator.py:106-126 at curr_layer == stage."""
    def block(x, i):                                             # generator.py:15-40
        p = f"_Generator__gen_blocks.{i}."
        x = pixel_norm(F.leaky_relu(F.conv2d(x, sd[p + "0.weight"], sd[p + "0.bias"], padding=1), 0.2))
        x = F.interpolate(x, scale_factor=2.0, mode="nearest")
        return pixel_norm(F.leaky_relu(F.conv2d(x, sd[p + "4.weight"], sd[p + "4.bias"], padding=1), 0.2))

    out = z
    for i in range(stage):
        out = block(out, i)
    top = block(out, stage)
    new = torch.tanh(F.conv2d(top, sd["_Generator__end_block.0.weight"], sd["_Generator__end_block.0.bias"]))
    if stage == 0:
        return new
    old = torch.tanh(F.conv2d(out, sd["_Generator__last_end_block.0.0.weight"], sd["_Generator__last_end_block.0.0.bias"]))
    old = F.interpolate(old, scale_factor=2.0, mode="nearest")
    return alpha * new + (1.0 - alpha) * old


def disc_forward(sd: SD, x: torch.Tensor, alpha: float, stage: int, return_features: bool = False) -> torch.Tensor:
    """discriminator.py:107-124 at curr_layer == 7 - stage."""
    def block(h, j):                                             # discriminator.py:14-34
        p = f"_Discriminator__conv_blocks.{j}."
        h = F.leaky_relu(F.conv2d(h, sd[p + "0.weight"], sd[p + "0.bias"], padding=1), 0.2)
        h = F.avg_pool2d(h, 2, 2)
        return F.leaky_relu(F.conv2d(h, sd[p + "3.weight"], sd[p + "3.bias"], padding=1), 0.2)

    cur = 7 - stage
    h = F.leaky_relu(F.conv2d(x, sd["_Discriminator__start_block.0.weight"], sd["_Discriminator__start_block.0.bias"]), 0.2)
    h = block(h, cur)
    if stage >= 1:
        old = F.leaky_relu(F.conv2d(F.avg_pool2d(x, 2, 2), sd["_Discriminator__last_start_block.1.0.weight"],
                                    sd["_Discriminator__last_start_block.1.0.bias"]), 0.2)
        h = alpha * h + (1 - alpha) * old
    for j in range(cur + 1, len(D_CHANNELS)):
        h = block(h, j)
    if return_features:
        return h.flatten(1, -1)
    return F.linear(h.flatten(1, -1), sd["_Discriminator__clf.0.weight"], sd["_Discriminator__clf.0.bias"])


def gradient_penalty(sd_d: SD, x_real, x_gen, alpha: float, stage: int, eps: torch.Tensor):
    """discriminator.py:157-184 with the uniform sample `eps` (B,1,1,1) supplied by the caller."""
    x_hat = eps * x_real + (1 - eps) * x_gen
    if not x_hat.requires_grad:
        x_hat.requires_grad_(True)
    out = disc_forward(sd_d, x_hat, alpha, stage)
    (g,) = torch.autograd.grad(out, x_hat, grad_outputs=torch.ones_like(out), create_graph=True, retain_graph=True)
    n = g.view(g.size(0), -1).norm(2, dim=1)
    return 10.0 * ((n - 1.0) ** 2.0).mean()


def _leaf(sd: SD) -> SD:
    return {k: v.clone().requires_grad_(True) for k, v in sd.items()}


def d_step(sd_g: SD, sd_d: SD, z, x_real, eps, alpha: float, stage: int, attached: bool = False):
    """train.py:143-174: losses and discriminator parameter gradients of one critic step (fake batch detached:
    SURVEY B.7 shows D's gradients are bit-identical to the non-detached reference).  attached=True keeps the fake
    batch in the graph like the reference does (the generator is back-propagated for nothing, train.py:152-174): same
    gradients, the reference's full cost -- used by the timed CPU baseline."""
    d = _leaf(sd_d)
    if attached:
        x_fake = gen_forward(_leaf(sd_g), z, alpha, stage)
    else:
        with torch.no_grad():
            x_fake = gen_forward(sd_g, z, alpha, stage)
    out_real, out_fake = disc_forward(d, x_real, alpha, stage), disc_forward(d, x_fake, alpha, stage)
    loss = -(out_real.mean() - out_fake.mean())                  # criterion.py:12-14
    gp = gradient_penalty(d, x_real, x_fake, alpha, stage, eps)
    (loss + gp).backward()
    return dict(loss=loss.detach(), gp=gp.detach(), x_fake=x_fake.detach(), out_real=out_real.detach(), out_fake=out_fake.detach(),
                grads={k: v.grad for k, v in d.items()})


def g_step(sd_g: SD, sd_d: SD, z, alpha: float, stage: int, critic_grads: bool = False):
    """train.py:191-213: generator loss and generator parameter gradients.  critic_grads=True also computes (and drops)
    the critic's parameter gradients like the reference's backward() does -- the timed CPU baseline's full cost."""
    g = _leaf(sd_g)
    x_fake = gen_forward(g, z, alpha, stage)
    out_fake = disc_forward(_leaf(sd_d) if critic_grads else sd_d, x_fake, alpha, stage)
    loss = -out_fake.mean()                                      # criterion.py:17-18
    loss.backward()
    return dict(loss=loss.detach(), x_fake=x_fake.detach(), out_fake=out_fake.detach(), grads={k: v.grad for k, v in g.items()})
