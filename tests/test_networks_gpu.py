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


BIG = {"stage7_b8"}                          # oracle takes tens of seconds on the host cores: run in the dedicated test
STEP_CASES = [k for k in CASES if k not in BIG]


@pytest.mark.parametrize("name", STEP_CASES)
def test_forward_vs_oracle(name):
    stage, batch, alpha, z, z2, x_real, eps = case_inputs(name)
    sd_g, sd_d = no.make_state("gen", stage, 11), no.make_state("disc", stage, 12)
    gen, disc = build(stage, sd_g, sd_d)
    with torch.no_grad():
        xf = gen(z.cuda(), alpha)
        ref_xf = no.gen_forward(sd_g, z, alpha, stage)
        assert xf.shape == ref_xf.shape and xf.dtype == torch.float32
        e_g = rel(xf, ref_xf)
        out = disc(x_real.cuda(), alpha)
        ref_out = no.disc_forward(sd_d, x_real, alpha, stage)
        assert out.shape == ref_out.shape == (batch, 1)
        e_d = rel(out, ref_out)
        print(f"{name}: G output rel-L2 {e_g:.2e}; D output rel-L2 {e_d:.2e}")
        assert e_g <= TOL, e_g
        assert e_d <= TOL, e_d


@pytest.mark.parametrize("name", ["stage5_b2", "stage7_b1"])
def test_generator_inference_mode(name):
    """generate.py:38,54 runs the generator in eval mode without a graph: the forward-only mode (plain bf16 weights above
    32 x 32, no gradient masks to protect) must still give the reference's images within rel-L2 1e-2."""
    stage, batch, alpha, z, z2, x_real, eps = case_inputs(name)
    sd_g, sd_d = no.make_state("gen", stage, 11), no.make_state("disc", stage, 12)
    gen, _ = build(stage, sd_g, sd_d)
    gen.eval()
    with torch.no_grad():
        xf = gen(z.cuda(), 1.0)
    ref = no.gen_forward(sd_g, z, 1.0, stage)
    e = rel(xf, ref)
    print(f"{name}: inference-mode G output rel-L2 {e:.2e}")
    assert e <= TOL, e


def _param_grads(module):
    return {k: p.grad for k, p in module.named_parameters() if p.grad is not None}


def _oracle_term_grads(sd_d, fn_of_leaf):
    d = no._leaf(sd_d)
    fn_of_leaf(d).backward()
    return {k: v.grad for k, v in d.items()}


def _norm(grads):
    return torch.cat([v.flatten() for v in grads.values() if v is not None]).double().norm().item()


def _our_term(disc, fn_of_disc):
    disc.zero_grad()
    fn_of_disc(disc).backward()
    return _param_grads(disc)


@pytest.mark.parametrize("name", STEP_CASES)
def test_discriminator_first_order_gradients(name):
    """Gradients of mean(D(x)) w.r.t. every active D parameter (the building block of the critic loss,
    train.py:155-159): rel-L2 <= 1e-2 against the fp32 oracle."""
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


def _check_steps(name, golden_dir):
    """One critic step and one generator step (train.py:143-214) against the fp32 oracle and the reference goldens; plain
    relative L2 <= 1e-2, no floor, for: the generator-step gradient, and each of the three terms of the critic gradient
    (mean D(real), mean D(fake), gradient penalty).  The critic-step TOTAL  g_fake - g_real + g_gp  is printed with its
    conditioning |terms| / |total| (15 at stage 0 ... 5e6 at stage 7, stored in the fixtures: with a random-init critic
    g_real ~ g_fake, and at 1e-6 of the terms the reference's own fp32 summation order shows); its error is asserted on
    the scale it is computed at, 1e-2 * (|g_real| + |g_fake| + |g_gp|), which the three term bounds imply."""
    from musicgan_b200 import train_step
    stage, batch, alpha, z, z2, x_real, eps = case_inputs(name)
    sd_g, sd_d = no.make_state("gen", stage, 11), no.make_state("disc", stage, 12)
    gen, disc = build(stage, sd_g, sd_d)
    gold = np.load(os.path.join(golden_dir, f"networks_{name}.npz"))
    xr, ec = x_real.cuda(), eps.cuda()

    d_loss, gp, out_real, out_fake = train_step.critic_step(gen, disc, None, z.cuda(), xr, alpha, eps=ec, step=False)
    ref = no.d_step(sd_g, sd_d, z, x_real, eps, alpha, stage)
    named = dict(disc.named_parameters())
    got_none = sorted(k for k, p in named.items() if p.grad is None)
    assert got_none == sorted(k for k, v in ref["grads"].items() if v is None) == sorted(gold["none_d"].tolist())
    g_tot, r_tot = cat_grads(_param_grads(disc), ref["grads"])
    for p in gen.parameters():
        assert p.grad is None                       # fake batch detached: G untouched by the critic step
    x_fake = ref["x_fake"]
    with torch.no_grad():
        our_fake = gen(z.cuda(), alpha)
    terms = {
        "real": (lambda d: d(xr, alpha).mean(), lambda d: no.disc_forward(d, x_real, alpha, stage).mean()),
        "fake": (lambda d: d(our_fake, alpha).mean(), lambda d: no.disc_forward(d, x_fake, alpha, stage).mean()),
        "gp": (lambda d: d.gradient_penalty(xr, our_fake, alpha, eps=ec), lambda d: no.gradient_penalty(d, x_real, x_fake, alpha, stage, eps)),
    }
    denom, errs = 0.0, {}
    for key, (ours, oracle) in terms.items():
        t_ref = _oracle_term_grads(sd_d, oracle)
        g, r = cat_grads(_our_term(disc, ours), t_ref)
        errs[key] = rel(g, r)
        denom += _norm(t_ref)
    e_plain = rel(g_tot, r_tot)
    e_terms = (g_tot.double() - r_tot.double()).norm().item() / denom
    print(f"{name}: critic-step gradient terms rel-L2: real {errs['real']:.2e} fake {errs['fake']:.2e} gp {errs['gp']:.2e}; "
          f"total: {e_terms:.2e} of the term scale, plain rel-L2 {e_plain:.2e} at conditioning {float(gold['critic_cond']):.1e}; "
          f"d_loss {d_loss.item():.6f} vs {ref['loss'].item():.6f}; gp {gp.item():.5f} vs {ref['gp'].item():.5f}")
    assert max(errs.values()) <= TOL, errs
    assert e_terms <= TOL, e_terms
    assert abs(gp.item() - float(gold["gp"])) <= 1e-2 * abs(float(gold["gp"]))
    assert rel(out_real, torch.from_numpy(gold["out_real"])) <= TOL

    g_loss, out_fake2 = train_step.generator_step(gen, disc, None, z2.cuda(), alpha, step=False)
    refg = no.g_step(sd_g, sd_d, z2, alpha, stage)
    named = dict(gen.named_parameters())
    assert sorted(k for k, p in named.items() if p.grad is None) == sorted(gold["none_g"].tolist())
    g, r = cat_grads(_param_grads(gen), refg["grads"])
    e = rel(g, r)
    worst = 0.0
    for k, p in named.items():
        if p.grad is None:
            continue
        ref_norm = float(gold["ggrad_digest/" + k][2])
        worst = max(worst, abs(p.grad.double().norm().item() - ref_norm) / max(ref_norm, 1e-30))
    print(f"{name}: G-step grads rel-L2 {e:.2e}; worst per-tensor norm deviation vs reference golden {worst:.2e}; "
          f"g_loss {g_loss.item():.5f} vs {refg['loss'].item():.5f}")
    assert e <= TOL, e
    assert worst <= 3 * TOL, worst
    assert abs(g_loss.item() - float(gold["g_loss"])) <= 1e-2 * max(abs(float(gold["g_loss"])), 1e-3)


@pytest.mark.parametrize("name", STEP_CASES)
def test_critic_and_generator_step_gradients(golden_dir, name):
    _check_steps(name, golden_dir)


def test_benchmarked_shape_stage7_batch8(golden_dir):
    """BASELINE config 2 itself: 512 x 512, batch 8, both fade paths alive -- the step bench.py times."""
    _check_steps("stage7_b8", golden_dir)


@pytest.mark.parametrize("stage,batch,scale", [(1, 2, 3.0), (3, 2, 3.0), (4, 2, 1.0), (6, 1, 1.0)])
def test_gradient_penalty_double_backward(stage, batch, scale):
    """The GP term alone: its parameter gradients exist only through the double-backward graph
    (fprop <-> dgrad <-> wgrad closure).  Plain rel-L2 <= 1e-2 against the fp32 oracle."""
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
    print(f"stage {stage} x{scale}: gp {gp.item():.5f} vs {ref_gp.item():.5f}; GP-only grads rel-L2 {e:.2e}")
    assert abs(gp.item() - ref_gp.item()) <= 1e-2 * abs(ref_gp.item())
    assert e <= TOL, e


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


@pytest.mark.parametrize("name", ["stage2_b3", "stage4_b2", "stage7_b8"])
def test_graphed_steps_vs_oracle(name):
    """The path bench.py and train() run -- graphed.GraphedSteps: two parallel branches, weight gradients on side streams
    (ops.WgradLane), bias gradients of the bf16 layers from the weight-gradient launch, one-launch weight packing --
    measured DIRECTLY against the fp32 oracle (train.py:143-214), not only through its equality with the eager steps:
    critic-step gradient on the scale of its terms and generator-step gradient, rel-L2 <= 1e-2.  The optimisers are SGD
    with lr 0 so that the warm-up and the replays leave the golden parameters where they are."""
    from musicgan_b200.graphed import GraphedSteps
    stage, batch, alpha, z, z2, x_real, eps = case_inputs(name)
    sd_g, sd_d = no.make_state("gen", stage, 11), no.make_state("disc", stage, 12)
    gen, disc = build(stage, sd_g, sd_d)
    og, od = torch.optim.SGD(gen.parameters(), lr=0.0), torch.optim.SGD(disc.parameters(), lr=0.0)
    gs = GraphedSteps(gen, disc, og, od, batch, 32, 4 * 2 ** stage, alpha, warmup=1, static_noise=True)
    gs.z.copy_(z.cuda()); gs.eps.copy_(eps.cuda())
    stats = gs.critic_step(x_real.cuda()).clone()
    ref = no.d_step(sd_g, sd_d, z, x_real, eps, alpha, stage)
    assert sorted(k for k, p in disc.named_parameters() if p.grad is None) == sorted(k for k, v in ref["grads"].items() if v is None)
    g_tot, r_tot = cat_grads(_param_grads(disc), ref["grads"])
    x_fake = ref["x_fake"]
    denom = 0.0
    for oracle in (lambda d: no.disc_forward(d, x_real, alpha, stage).mean(), lambda d: no.disc_forward(d, x_fake, alpha, stage).mean(),
                   lambda d: no.gradient_penalty(d, x_real, x_fake, alpha, stage, eps)):
        denom += _norm(_oracle_term_grads(sd_d, oracle))
    e_terms = (g_tot.double() - r_tot.double()).norm().item() / denom
    gs.z.copy_(z2.cuda())
    gstats = gs.generator_step().clone()
    refg = no.g_step(sd_g, sd_d, z2, alpha, stage)
    assert sorted(k for k, p in gen.named_parameters() if p.grad is None) == sorted(k for k, v in refg["grads"].items() if v is None)
    g, r = cat_grads(_param_grads(gen), refg["grads"])
    e_g = rel(g, r)
    print(f"{name} (graph replay): critic-step gradient {e_terms:.2e} of the term scale, gp {stats[1].item():.5f} vs {ref['gp'].item():.5f}; "
          f"G-step gradient rel-L2 {e_g:.2e}, g_loss {gstats[0].item():.5f} vs {refg['loss'].item():.5f}")
    assert e_terms <= TOL, e_terms
    assert abs(stats[1].item() - ref["gp"].item()) <= 1e-2 * abs(ref["gp"].item())
    assert e_g <= TOL, e_g
    assert abs(gstats[0].item() - refg["loss"].item()) <= 1e-2 * max(abs(refg["loss"].item()), 1e-3)


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


def test_several_graph_captures_in_one_process():
    """`train` re-captures at every growth event and bench.py's sweep captures one GraphedSteps per stage: kernel
    scratch requested during a capture lives in that graph's private memory pool and must die with it (round 2: the
    process-wide workspace cache handed a destroyed graph's addresses to the next one -> illegal memory access)."""
    import gc
    from musicgan_b200.graphed import GraphedSteps
    batch = 4
    for stage in (0, 1, 2, 3):
        gen, disc = build(stage, no.make_state("gen", stage, 41), no.make_state("disc", stage, 42))
        res = 4 * 2 ** stage
        og = torch.optim.Adam(gen.parameters(), lr=1e-3, betas=(0.0, 0.9), capturable=True, fused=True)
        od = torch.optim.Adam(disc.parameters(), lr=1e-3, betas=(0.0, 0.9), capturable=True, fused=True)
        gs = GraphedSteps(gen, disc, og, od, batch, 32, res, 0.5, warmup=1)
        x_real = torch.rand(batch, 2, res, res, device="cuda") * 2 - 1
        for it in range(6):
            stats = gs.critic_step(x_real)
            if it % 5 == 0:
                gs.generator_step()
        torch.cuda.synchronize()
        assert torch.isfinite(stats).all()
        del gs, gen, disc, og, od
        gc.collect()
        torch.cuda.empty_cache()


def test_one_launch_weight_packing_equals_per_tensor_packing():
    """ops.prepack (mg_pack_weights_multi: every packed copy of a network in one launch) writes the same bytes as the
    per-tensor pack kernels the eager path uses, for every kind of copy (bf16 / hi + lo / split parts, both orientations)."""
    from musicgan_b200.networks import ops
    stage = 5
    gen, disc = build(stage, no.make_state("gen", stage, 51), no.make_state("disc", stage, 52))
    x = (torch.rand(2, 2, 128, 128) * 2 - 1).cuda().requires_grad_(True)
    z = torch.randn(2, 32, 2, 2).cuda()
    (disc(gen(z, 0.5), 0.5).mean() + disc.gradient_penalty(x.detach(), gen(z, 0.5).detach(), 0.5)).backward()      # asks for every kind
    torch.cuda.synchronize()
    n = 0
    for module in (gen, disc):
        single = {(id(p), key): buf.clone() for p in module.parameters() for key, (_, buf) in getattr(p, "_mg_packed", {}).items()}
        for p in module.parameters():
            for key, (_, buf) in getattr(p, "_mg_packed", {}).items():
                buf.zero_()
        ops.invalidate_pack_cache()
        ops.prepack(module)
        torch.cuda.synchronize()
        for p in module.parameters():
            for key, (stamp, buf) in getattr(p, "_mg_packed", {}).items():
                kind, cin, cout = key
                used = 9 * cin * cout * 2 * ((3 if kind[1] & 2 else 2) if kind[0] == "split" else (2 if kind[1] & 2 else 1))
                assert torch.equal(buf[:used], single[(id(p), key)][:used]), key
                n += 1
    assert n >= 60
