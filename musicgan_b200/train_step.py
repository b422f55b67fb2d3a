"""The two optimisation steps of the reference training loop (train.py:143-175 critic step, :191-214
generator step) as functions of the drop-in modules; used by `train`, bench.py and the parity tests.

Differences from the reference that do not change any parameter update: the fake batch is
detached in the critic step (the reference back-propagates into G and then throws those gradients away).
"""
from __future__ import annotations

import torch as th

from . import networks


def critic_step(gen, disc, optim_disc, z, x_real, alpha: float, eps=None, step: bool = True):
    with th.no_grad():
        x_fake = gen(z, alpha)
    # the critic has no batch-coupled layer, so D(real) and D(fake) are one pass over the concatenated batch
    n = x_real.size(0)
    out_both = disc(th.cat([x_real, x_fake], dim=0), alpha)
    out_real, out_fake = out_both[:n], out_both[n:]
    disc_loss = networks.wasserstein_discriminator_loss(out_real, out_fake)
    grad_pen = disc.gradient_penalty(x_real, x_fake, alpha, eps=eps)
    gen.zero_grad()
    disc.zero_grad()
    (disc_loss + grad_pen).backward()
    if step and optim_disc is not None:
        optim_disc.step()
        networks.ops.invalidate_pack_cache()      # parameters changed (fused optimisers do not bump Tensor._version)
    return disc_loss.detach(), grad_pen.detach(), out_real.detach(), out_fake.detach()


class frozen:
    """The critic's parameters do not require grad inside the block: the generator step differentiates THROUGH the
    critic but uses none of its parameter gradients (the reference computes them and throws them away at the next
    zero_grad, train.py:203-214) -- 18 weight-gradient launches per generator step that are simply not issued here."""

    def __init__(self, module):
        self.params = [p for p in module.parameters() if p.requires_grad]

    def __enter__(self):
        for p in self.params:
            p.requires_grad_(False)
            p._mg_frozen = True          # still a parameter for the packed-weight cache (networks/ops.py)

    def __exit__(self, *exc):
        for p in self.params:
            p.requires_grad_(True)
            p._mg_frozen = False
        return False


def generator_step(gen, disc, optim_gen, z, alpha: float, step: bool = True):
    gen.zero_grad()
    disc.zero_grad()
    with frozen(disc):
        x_fake = gen(z, alpha)
        out_fake = disc(x_fake, alpha)
        gen_loss = networks.wasserstein_generator_loss(out_fake)
        gen_loss.backward()
    if step and optim_gen is not None:
        optim_gen.step()
        networks.ops.invalidate_pack_cache()
    return gen_loss.detach(), out_fake.detach()
