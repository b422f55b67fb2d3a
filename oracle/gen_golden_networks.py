"""Golden vectors for the networks from the UNMODIFIED reference modules (authoring container only; called by
oracle/gen_golden.py).  Seeded synthetic state dicts (oracle.networks_oracle.make_state) are loaded into the
reference's Generator / Discriminator, which also proves that the key names / shapes are the reference's."""
import os

import numpy as np
import torch as th

from oracle import networks_oracle as no

CASES = {   # name -> (stage, batch, alpha)
    "stage0_b2": (0, 2, 1.0),
    "stage2_b3": (2, 3, 0.5),
    "stage4_b2": (4, 2, 0.25),
}


def case_inputs(name):
    stage, batch, alpha = CASES[name]
    g = th.Generator().manual_seed(1000 + stage)
    r = 4 * 2 ** stage
    z = th.randn(batch, 32, 2, 2, generator=g)
    z2 = th.randn(batch, 32, 2, 2, generator=g)
    x_real = th.rand(batch, 2, r, r, generator=g) * 2 - 1
    eps = th.rand(batch, 1, 1, 1, generator=g)
    return stage, batch, alpha, z, z2, x_real, eps


def _digest(t):
    d = t.detach().double()
    return np.array([d.sum().item(), d.abs().sum().item(), d.norm().item()], dtype=np.float64)


def main(ref_networks, gold_dir):
    for name in CASES:
        stage, batch, alpha, z, z2, x_real, eps = case_inputs(name)
        sd_g, sd_d = no.make_state("gen", stage, 11), no.make_state("disc", stage, 12)
        gen, disc = ref_networks.Generator(32, 0), ref_networks.Discriminator(7)
        for _ in range(stage):
            gen.next_layer(); disc.next_layer()
        gen.load_state_dict(sd_g, strict=True)
        disc.load_state_dict(sd_d, strict=True)
        assert [k for k, _ in no.state_shapes("gen", stage)] == list(gen.state_dict().keys())
        assert [k for k, _ in no.state_shapes("disc", stage)] == list(disc.state_dict().keys())

        # ---- reference D step (train.py:143-174), eps injected by patching th.rand for the one call ----
        x_fake = gen(z, alpha)
        out_real, out_fake = disc(x_real, alpha), disc(x_fake, alpha)
        d_loss = ref_networks.wasserstein_discriminator_loss(out_real, out_fake)
        real_rand = th.rand
        th.rand = lambda *a, **k: eps
        try:
            gp = disc.gradient_penalty(x_real, x_fake, alpha)
        finally:
            th.rand = real_rand
        gen.zero_grad(); disc.zero_grad()
        (d_loss + gp).backward()
        d_grads = {k: p.grad.clone() for k, p in disc.named_parameters() if p.grad is not None}
        none_d = [k for k, p in disc.named_parameters() if p.grad is None]

        # ---- reference G step (train.py:191-213) ----
        x_fake2 = gen(z2, alpha)
        out_fake2 = disc(x_fake2, alpha)
        g_loss = ref_networks.wasserstein_generator_loss(out_fake2)
        gen.zero_grad(); disc.zero_grad()
        g_loss.backward()
        g_grads = {k: p.grad.clone() for k, p in gen.named_parameters() if p.grad is not None}
        none_g = [k for k, p in gen.named_parameters() if p.grad is None]

        # ---- oracle on the same inputs ----
        o_d = no.d_step(sd_g, sd_d, z, x_real, eps, alpha, stage)
        o_g = no.g_step(sd_g, sd_d, z2, alpha, stage)
        ex = th.equal(o_d["x_fake"], x_fake.detach()) and th.equal(o_d["out_real"], out_real.detach()) and th.equal(o_g["out_fake"], out_fake2.detach())
        rel = lambda a, b: ((a - b).norm() / b.norm().clamp_min(1e-30)).item()
        worst_d = max(rel(o_d["grads"][k], d_grads[k]) for k in d_grads)
        worst_g = max(rel(o_g["grads"][k], g_grads[k]) for k in g_grads)
        print(f"networks/{name}: forward bit-exact {ex}; oracle-vs-reference grad rel-L2 worst D {worst_d:.2e} G {worst_g:.2e}; "
              f"gp {gp.item():.6f} vs {o_d['gp'].item():.6f}; params without grad: D {len(none_d)} G {len(none_g)}")
        assert worst_d < 1e-4 and worst_g < 1e-4

        arrays = dict(stage=np.int64(stage), batch=np.int64(batch), alpha=np.float64(alpha),
                      forward_bit_exact=np.bool_(ex),
                      x_fake=x_fake.detach().numpy(), out_real=out_real.detach().numpy(), out_fake=out_fake.detach().numpy(),
                      d_loss=np.float64(d_loss.item()), gp=np.float64(gp.item()), g_loss=np.float64(g_loss.item()),
                      out_fake2=out_fake2.detach().numpy(),
                      none_d=np.array(none_d), none_g=np.array(none_g))
        for k, v in d_grads.items():
            arrays["dgrad_digest/" + k] = _digest(v)
            arrays["dgrad/" + k] = v.numpy() if v.numel() <= 4096 else v.contiguous().view(-1)[::37].numpy().copy()
        for k, v in g_grads.items():
            arrays["ggrad_digest/" + k] = _digest(v)
            arrays["ggrad/" + k] = v.numpy() if v.numel() <= 4096 else v.contiguous().view(-1)[::37].numpy().copy()
        np.savez_compressed(os.path.join(gold_dir, f"networks_{name}.npz"), **arrays)
