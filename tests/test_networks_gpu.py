"""GPU parity of the drop-in Generator / Discriminator (tcgen05 convolutions, bf16 activations) against the fp32
CPU oracle and the reference's golden vectors.  Tolerance (BASELINE north_star): relative L2 <= 1e-2 for G/D
outputs and parameter gradients."""
import os

import numpy as np
import pytest
import torch

from oracle import networks_oracle as no
from oracle.gen_golden_networks import CASES, case_inputs

pytestmark = pytest.mark.gpu
TOL = 1e-2


def rel(a, b):
    a, b = a.detach().double().cpu().flatten(), b.detach().double().cpu().flatten()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


def build(stage, sd_g, sd_d):
    from musicgan_b200 import networks
    gen, disc = networks.Generator(32, 0), networks.Discriminator(7)
    for _ in range(stage):
        assert gen.next_layer() and disc.next_layer()
    gen.load_state_dict(sd_g, strict=True)
    disc.load_state_dict(sd_d, strict=True)
    return gen.cuda(), disc.cuda()


def cat_grads(named, ref):
    keys = [k for k in ref if ref[k] is not None]
    for k in list(keys):
        if k not in named:      # no graph path in ours: only legitimate when the reference gradient is exactly zero
            assert float(ref[k].abs().max()) == 0.0, k      # (e.g. biases under the gradient-penalty term alone)
            keys.remove(k)
    return torch.cat([named[k].detach().float().cpu().flatten() for k in keys]), torch.cat([ref[k].flatten() for k in keys])


@pytest.mark.parametrize("name", list(CASES))
def test_forward_vs_oracle(name):
    stage, batch, alpha, z, z2, x_real, eps = case_inputs(name)
    sd_g, sd_d = no.make_state("gen", stage, 11), no.make_state("disc", stage, 12)
    gen, disc = build(stage, sd_g, sd_d)
    with torch.no_grad():
        xf = gen(z.cuda(), alpha)
        ref_xf = no.gen_forward(sd_g, z, alpha, stage)
        assert xf.shape == ref_xf.shape and xf.dtype == torch.float32
        assert rel(xf, ref_xf) <= TOL, rel(xf, ref_xf)
        out = disc(x_real.cuda(), alpha)
        ref_out = no.disc_forward(sd_d, x_real, alpha, stage)
        assert out.shape == ref_out.shape == (batch, 1)
        assert rel(out, ref_out) <= TOL, rel(out, ref_out)


def _param_grads(module):
    return {k: p.grad for k, p in module.named_parameters() if p.grad is not None}


def _oracle_term_grads(sd_d, fn_of_leaf):
    d = no._leaf(sd_d)
    fn_of_leaf(d).backward()
    return {k: v.grad for k, v in d.items()}


@pytest.mark.parametrize("name", list(CASES))
def test_discriminator_first_order_gradients(name):
    """Gradients of mean(D(x)) w.r.t. every active D parameter (the well-conditioned building block of the
    critic loss, train.py:155-159): rel-L2 <= 1e-2 against the fp32 oracle."""
    stage, batch, alpha, z, z2, x_real, eps = case_inputs(name)
    sd_g, sd_d = no.make_state("gen", stage, 11), no.make_state("disc", stage, 12)
    _, disc = build(stage, sd_g, sd_d)
    disc.zero_grad()
    disc(x_real.cuda(), alpha).mean().backward()
    ref = _oracle_term_grads(sd_d, lambda d: no.disc_forward(d, x_real, alpha, stage).mean())
    g, r = cat_grads(_param_grads(disc), ref)
    e = rel(g, r)
    print(f"{name}: grad of mean D(x_real) rel-L2 {e:.2e}")
    assert e <= TOL, e


@pytest.mark.parametrize("name", list(CASES))
def test_critic_and_generator_step_gradients(golden_dir, name):
    """One critic step and one generator step (train.py:143-214) against the oracle and the reference goldens.
    The critic gradient is a DIFFERENCE of two nearly equal terms (real minus fake) plus the penalty term, so its
    error is measured against the summed norms of the three terms (each term is checked on its own elsewhere)."""
    from musicgan_b200 import train_step
    stage, batch, alpha, z, z2, x_real, eps = case_inputs(name)
    sd_g, sd_d = no.make_state("gen", stage, 11), no.make_state("disc", stage, 12)
    gen, disc = build(stage, sd_g, sd_d)
    gold = np.load(os.path.join(golden_dir, f"networks_{name}.npz"))

    d_loss, gp, out_real, out_fake = train_step.critic_step(gen, disc, None, z.cuda(), x_real.cuda(), alpha, eps=eps.cuda(), step=False)
    ref = no.d_step(sd_g, sd_d, z, x_real, eps, alpha, stage)
    named = dict(disc.named_parameters())
    got_none = sorted(k for k, p in named.items() if p.grad is None)
    assert got_none == sorted(k for k, v in ref["grads"].items() if v is None) == sorted(gold["none_d"].tolist())
    g, r = cat_grads(_param_grads(disc), ref["grads"])
    x_fake = ref["x_fake"]
    t_real = _oracle_term_grads(sd_d, lambda d: no.disc_forward(d, x_real, alpha, stage).mean())
    t_fake = _oracle_term_grads(sd_d, lambda d: no.disc_forward(d, x_fake, alpha, stage).mean())
    t_gp = _oracle_term_grads(sd_d, lambda d: no.gradient_penalty(d, x_real, x_fake, alpha, stage, eps))
    denom = sum(torch.cat([v.flatten() for v in t.values() if v is not None]).double().norm().item() for t in (t_real, t_fake, t_gp))
    e_terms = (g.double() - r.double()).norm().item() / denom
    # noise floor in the same normalisation: the fp32 oracle's own critic gradient when ONLY the weights of G and D are
    # rounded to bf16 (the penalty term is ill conditioned, see test_gradient_penalty_double_backward)
    rb = lambda sd: {k: v.bfloat16().float() for k, v in sd.items()}
    refb = no.d_step(rb(sd_g), rb(sd_d), z, x_real, eps, alpha, stage)
    keys = [k for k, v in ref["grads"].items() if v is not None]
    floor = (torch.cat([refb["grads"][k].flatten() for k in keys]).double()
             - torch.cat([ref["grads"][k].flatten() for k in keys]).double()).norm().item() / denom
    print(f"{name}: critic-step grads: error / (|g_real|+|g_fake|+|g_gp|) = {e_terms:.2e} (oracle floor under bf16 weight "
          f"rounding {floor:.2e}; plain rel-L2 of the difference {rel(g, r):.2e}); "
          f"d_loss {d_loss.item():.6f} vs {ref['loss'].item():.6f}; gp {gp.item():.5f} vs {ref['gp'].item():.5f}")
    assert e_terms <= max(TOL, 2.5 * floor), (e_terms, floor)        # same factor as the penalty term on its own
    assert abs(gp.item() - float(gold["gp"])) <= 1e-2 * abs(float(gold["gp"]))
    assert rel(out_real, torch.from_numpy(gold["out_real"])) <= TOL
    for p in gen.parameters():
        assert p.grad is None                       # fake batch detached: G untouched by the critic step

    g_loss, out_fake2 = train_step.generator_step(gen, disc, None, z2.cuda(), alpha, step=False)
    refg = no.g_step(sd_g, sd_d, z2, alpha, stage)
    named = dict(gen.named_parameters())
    assert sorted(k for k, p in named.items() if p.grad is None) == sorted(gold["none_g"].tolist())
    g, r = cat_grads(_param_grads(gen), refg["grads"])
    e = rel(g, r)
    # noise floor: the fp32 oracle's own G gradient when ONLY the weights of G and D are rounded to bf16
    rb = lambda sd: {k: v.bfloat16().float() for k, v in sd.items()}
    refb = no.g_step(rb(sd_g), rb(sd_d), z2, alpha, stage)
    keys = [k for k, v in refg["grads"].items() if v is not None]
    floor = rel(torch.cat([refb["grads"][k].flatten() for k in keys]), torch.cat([refg["grads"][k].flatten() for k in keys]))
    print(f"{name}: G-step grads rel-L2 {e:.2e} (oracle floor under bf16 weight rounding {floor:.2e}); "
          f"g_loss {g_loss.item():.5f} vs {refg['loss'].item():.5f}")
    assert e <= max(TOL, 2.0 * floor), (e, floor)
    assert abs(g_loss.item() - float(gold["g_loss"])) <= 1e-2 * max(abs(float(gold["g_loss"])), 1e-3)
    worst = 0.0
    for k, p in named.items():
        if p.grad is None:
            continue
        ref_norm = float(gold["ggrad_digest/" + k][2])
        worst = max(worst, abs(p.grad.double().norm().item() - ref_norm) / max(ref_norm, 1e-30))
    print(f"{name}: worst per-tensor G grad norm deviation vs reference golden {worst:.2e}")
    assert worst <= max(3e-2, 2.0 * floor)


def _bf16_weight_floor(sd_d, term):
    """How much the fp32 ORACLE's own gradient moves when only its weights are rounded to bf16 (all arithmetic
    still fp32): the noise floor any bf16-operand implementation inherits (cf. SURVEY B.4 for the IF path)."""
    a = _oracle_term_grads(sd_d, term)
    sd_b = {k: v.bfloat16().float() for k, v in sd_d.items()}
    b = _oracle_term_grads(sd_b, term)
    keys = [k for k in a if a[k] is not None]
    return rel(torch.cat([b[k].flatten() for k in keys]), torch.cat([a[k].flatten() for k in keys]))


@pytest.mark.parametrize("stage,batch,scale", [(1, 2, 3.0), (3, 2, 3.0), (4, 2, 1.0)])
def test_gradient_penalty_double_backward(stage, batch, scale):
    """The GP term alone: its parameter gradients exist only through the double-backward graph
    (fprop <-> dgrad <-> wgrad closure).  The GP gradient of a piecewise-linear critic is ill conditioned: rounding
    just the WEIGHTS of the fp32 oracle to bf16 moves it by several percent, so the bound is that measured floor."""
    g = torch.Generator().manual_seed(77 + stage)
    r = 4 * 2 ** stage
    x_real = torch.rand(batch, 2, r, r, generator=g) * 2 - 1
    x_fake = torch.rand(batch, 2, r, r, generator=g) * 2 - 1
    eps = torch.rand(batch, 1, 1, 1, generator=g)
    sd_d = {k: v * scale for k, v in no.make_state("disc", stage, 5).items()}
    _, disc = build(stage, no.make_state("gen", stage, 4), sd_d)
    gp = disc.gradient_penalty(x_real.cuda(), x_fake.cuda(), 0.5, eps=eps.cuda())
    disc.zero_grad()
    gp.backward()
    term = lambda d: no.gradient_penalty(d, x_real, x_fake, 0.5, stage, eps)
    ref = _oracle_term_grads(sd_d, term)
    ref_gp = term(sd_d)
    gg, rr = cat_grads(_param_grads(disc), ref)
    e = rel(gg, rr)
    floor = _bf16_weight_floor(sd_d, term)
    print(f"stage {stage} x{scale}: gp {gp.item():.5f} vs {ref_gp.item():.5f}; GP-only grads rel-L2 {e:.2e}; "
          f"oracle floor under bf16 weight rounding {floor:.2e}")
    assert abs(gp.item() - ref_gp.item()) <= 1e-2 * abs(ref_gp.item())
    assert e <= max(TOL, 2.5 * floor), (e, floor)


def test_full_resolution_stage7():
    """BASELINE config 2 geometry (512 x 512, both fade paths alive, alpha = 0.5) at batch 1: outputs and
    first-order gradients against the fp32 CPU oracle."""
    stage, alpha = 7, 0.5
    sd_g, sd_d = no.make_state("gen", stage, 21), no.make_state("disc", stage, 22)
    gen, disc = build(stage, sd_g, sd_d)
    g = torch.Generator().manual_seed(5)
    z = torch.randn(1, 32, 2, 2, generator=g)
    x_real = torch.rand(1, 2, 512, 512, generator=g) * 2 - 1
    with torch.no_grad():
        xf = gen(z.cuda(), alpha)
        ref_xf = no.gen_forward(sd_g, z, alpha, stage)
    assert tuple(xf.shape) == (1, 2, 512, 512)
    e_g = rel(xf, ref_xf)
    disc.zero_grad()
    out = disc(x_real.cuda(), alpha)
    out.mean().backward()
    d = no._leaf(sd_d)
    ref_out = no.disc_forward(d, x_real, alpha, stage)
    ref_out.mean().backward()
    gg, rr = cat_grads(_param_grads(disc), {k: v.grad for k, v in d.items()})
    # the critic output is clf_w . features + clf_b; measure its error against |clf_w| . |features| + |clf_b| (the
    # scale of the terms being summed), not against the possibly cancelling sum
    feat = no.disc_forward(sd_d, x_real, alpha, stage, return_features=True)
    scale = (sd_d["_Discriminator__clf.0.weight"].abs() @ feat.abs().t()).max().item() + sd_d["_Discriminator__clf.0.bias"].abs().item()
    e_o, e_d = (out.detach().cpu() - ref_out.detach()).abs().max().item() / scale, rel(gg, rr)
    print(f"stage 7: G output rel-L2 {e_g:.2e}; D output error / term scale {e_o:.2e}; D first-order grads {e_d:.2e}")
    assert e_g <= 1.2 * TOL and e_o <= TOL and e_d <= TOL


def test_growth_and_shapes():
    """networks/test_networks.py:4-38 of the reference: the only 'known answers' it has -- output sizes per stage."""
    from musicgan_b200 import networks
    gen, disc = networks.Generator(32), networks.Discriminator(7)
    gen.cuda(); disc.cuda()
    z = torch.randn(2, 32, 2, 2).cuda()
    for s in range(8):
        with torch.no_grad():
            out = gen(z, 0.5)
            assert tuple(out.shape) == (2, 2, 4 * 2 ** s, 4 * 2 ** s)
            assert tuple(disc(out, 0.5).shape) == (2, 1)
        grew = gen.next_layer()
        assert grew == disc.next_layer() == (s < 7)
    assert not gen.growing and not disc.growing


def test_graph_replays_equal_eager_steps():
    """bench.py's headline path replays CUDA graphs (graphed.py): the same latent vectors, penalty samples and real
    batches must move the parameters exactly like the eager steps of train_step.py (same kernels, same order; only the
    1x1-layer weight gradients use atomics, so agreement is to fp32 summation-order noise, amplified by Adam's
    normalisation for near-zero gradients -> compared as update directions)."""
    from musicgan_b200 import train_step
    from musicgan_b200.graphed import GraphedSteps
    stage, batch, alpha = 3, 4, 0.5
    res = 4 * 2 ** stage
    sd_g, sd_d = no.make_state("gen", stage, 21), no.make_state("disc", stage, 22)
    pairs = []
    for _ in range(2):
        gen, disc = build(stage, sd_g, sd_d)
        og = torch.optim.Adam(gen.parameters(), lr=1e-3, betas=(0.0, 0.9), capturable=True, fused=True)
        od = torch.optim.Adam(disc.parameters(), lr=1e-3, betas=(0.0, 0.9), capturable=True, fused=True)
        pairs.append((gen, disc, og, od))
    (gen_e, disc_e, og_e, od_e), (gen_g, disc_g, og_g, od_g) = pairs
    start = {k: v.detach().clone() for k, v in list(gen_e.named_parameters()) + list(disc_e.named_parameters())}
    gs = GraphedSteps(gen_g, disc_g, og_g, od_g, batch, 32, res, alpha, warmup=1, static_noise=True)
    # the warm-up before capture took optimiser steps: put parameters and Adam state back IN PLACE (the graphs hold
    # the storages)
    gen_g.load_state_dict({k: v.cuda() for k, v in sd_g.items()}, strict=True)
    disc_g.load_state_dict({k: v.cuda() for k, v in sd_d.items()}, strict=True)
    for opt in (og_g, od_g):
        for st in opt.state.values():
            st["exp_avg"].zero_(); st["exp_avg_sq"].zero_(); st["step"].zero_()
    for k, p in list(gen_g.named_parameters()) + list(disc_g.named_parameters()):
        assert torch.equal(p.detach(), start[k]), k
    g = torch.Generator().manual_seed(5)
    for it in range(3):
        x_real = (torch.rand(batch, 2, res, res, generator=g) * 2 - 1).cuda()
        z, z2 = torch.randn(batch, 32, 2, 2, generator=g).cuda(), torch.randn(batch, 32, 2, 2, generator=g).cuda()
        eps = torch.rand(batch, 1, 1, 1, generator=g).cuda()
        d_e = train_step.critic_step(gen_e, disc_e, od_e, z, x_real, alpha, eps=eps)
        gs.z.copy_(z); gs.eps.copy_(eps)
        stats = gs.critic_step(x_real).clone()
        # d_loss = mean D(fake) - mean D(real) is a difference of nearly equal numbers: tolerance on their scale
        scale = abs(d_e[2].mean().item()) + abs(d_e[3].mean().item())
        print(f"it {it}: d_loss {stats[0].item():.3e} vs {d_e[0].item():.3e} (scale {scale:.3e}); gp {stats[1].item():.6f} vs {d_e[1].item():.6f}")
        assert abs(stats[0].item() - d_e[0].item()) <= 1e-3 * scale
        assert abs(stats[1].item() - d_e[1].item()) <= 1e-3 * abs(d_e[1].item())
        g_e = train_step.generator_step(gen_e, disc_e, og_e, z2, alpha)
        gs.z.copy_(z2)
        gstats = gs.generator_step().clone()
        print(f"it {it}: g_loss {gstats[0].item():.3e} vs {g_e[0].item():.3e}")
        assert abs(gstats[0].item() - g_e[0].item()) <= 1e-3 * max(abs(g_e[0].item()), scale)
    torch.cuda.synchronize()
    named_g = dict(list(gen_g.named_parameters()) + list(disc_g.named_parameters()))
    moved = 0
    for k, pe in list(gen_e.named_parameters()) + list(disc_e.named_parameters()):
        de, dg = (pe - start[k]).flatten().double(), (named_g[k] - start[k]).flatten().double()
        if de.norm() == 0:
            assert dg.norm() == 0, k              # blocks not in the active path stay untouched in both
            continue
        moved += 1
        cos = (de @ dg / (de.norm() * dg.norm())).item()
        assert cos >= 0.999, (k, cos)
    assert moved > 10


def test_repeated_steps_are_bitwise_reproducible():
    """Race / hazard detector in place of compute-sanitizer (closed on this pool): with programmatic dependent launch,
    early accumulator release, deterministic two-step reductions and the two-stream critic graph, the SAME critic step
    on the SAME inputs must give bit-identical convolution weight / bias gradients, losses and penalties every time
    (only the 1x1-layer gradients use atomics and may differ in the last bits)."""
    from musicgan_b200 import train_step
    stage, batch, alpha = 4, 4, 0.5
    res = 4 * 2 ** stage
    gen, disc = build(stage, no.make_state("gen", stage, 31), no.make_state("disc", stage, 32))
    g = torch.Generator().manual_seed(9)
    x_real = (torch.rand(batch, 2, res, res, generator=g) * 2 - 1).cuda()
    z = torch.randn(batch, 32, 2, 2, generator=g).cuda()
    eps = torch.rand(batch, 1, 1, 1, generator=g).cuda()
    ref = None
    for rep in range(12):
        d_loss, gp, out_real, out_fake = train_step.critic_step(gen, disc, None, z, x_real, alpha, eps=eps, step=False)
        snap = {"d_loss": d_loss.clone(), "gp": gp.clone(), "out_real": out_real.clone(), "out_fake": out_fake.clone()}
        for k, p in disc.named_parameters():
            if p.grad is not None and p.dim() == 4 and tuple(p.shape[2:]) == (3, 3):
                snap[k] = p.grad.clone()
            elif p.grad is not None and p.dim() == 1 and "conv_blocks" in k:
                snap[k] = p.grad.clone()
        if ref is None:
            ref = snap
            assert len(ref) > 20
            continue
        for k, v in snap.items():
            assert torch.equal(v, ref[k]), (rep, k, (v.float() - ref[k].float()).abs().max().item())
