"""CPU emulation of storage / operand rounding in the ProGAN step (authoring tool, not product, not a test).

Every tensor the CUDA path stores between kernels (activations, activation gradients) and every MMA operand (weights)
is rounded here with a chosen mode, all arithmetic stays fp32, and the resulting parameter gradients are compared with
the fp32 oracle.  Used to decide which layers need which precision to meet north_star's rel-L2 1e-2.

    python scripts/precision_study.py
"""
import os
import sys
import itertools

import torch
import torch.nn.functional as F
from torch.autograd import Function

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import networks_oracle as no                     # noqa: E402
from oracle.gen_golden_networks import CASES, case_inputs    # noqa: E402


def rnd(x, mode):
    """straight-through rounding (the gradient passes unrounded: weight gradients are fp32 in the product)"""
    if mode == "fp32":
        return x
    with torch.no_grad():
        r = _rnd(x.detach(), mode)
    return x + (r - x.detach())


def _rnd(x, mode):
    if mode == "bf16":
        return x.bfloat16().float()
    if mode == "split2":                    # hi + lo bf16 pair
        hi = x.bfloat16().float()
        return hi + (x - hi).bfloat16().float()
    if mode == "split3":
        hi = x.bfloat16().float()
        mid = (x - hi).bfloat16().float()
        return hi + mid + (x - hi - mid).bfloat16().float()
    if mode == "fp16":
        return x.half().float()
    if mode == "tf32":
        i = x.contiguous().view(torch.int32)
        i = (i + 0x1000) & ~0x1FFF
        return i.view(torch.float32)
    raise ValueError(mode)


class Q(Function):
    @staticmethod
    def forward(ctx, x, f, b):
        ctx.f, ctx.b = f, b
        return rnd(x, f)

    @staticmethod
    def backward(ctx, g):
        return Q.apply(g, ctx.b, ctx.f), None, None


def q(x, f, b):
    if f == "fp32" and b == "fp32":
        return x
    return Q.apply(x, f, b)


class Policy:
    """modes per layer, decided from the layer's OUTPUT resolution (H of the conv output)."""

    def __init__(self, w_hi="bf16", a_hi="bf16", g_hi="bf16", w_lo="bf16", a_lo="bf16", g_lo="bf16", thr=0):
        self.hi = (w_hi, a_hi, g_hi)
        self.lo = (w_lo, a_lo, g_lo)
        self.thr = thr

    def __call__(self, res):
        return self.lo if res <= self.thr else self.hi

    def __repr__(self):
        return f"hi(w,a,g)={self.hi} lo={self.lo} thr={self.thr}"


def pixel_norm(x):
    return x / torch.sqrt(x.pow(2.0).mean(dim=1, keepdim=True) + 1e-8)


def gen_forward(sd, z, alpha, stage, pol):
    def half(x, w, b, up):
        res = x.shape[2] * (2 if up else 1)
        wm, am, gm = pol(res)
        if up:
            x = F.interpolate(x, scale_factor=2.0, mode="nearest")
            x = q(x, "fp32", gm)           # dgrad output stored before the 2x2 sum
        x = q(x, am, "fp32")               # the MMA reads its input in this layer's operand precision
        y = F.conv2d(x, rnd(w, wm), None, padding=1) + b[None, :, None, None]
        y = q(y, "fp32", gm)               # gz stored
        return q(pixel_norm(F.leaky_relu(y, 0.2)), am, gm)

    def block(x, i):
        p = f"_Generator__gen_blocks.{i}."
        x = half(x, sd[p + "0.weight"], sd[p + "0.bias"], False)
        return half(x, sd[p + "4.weight"], sd[p + "4.bias"], True)

    out = q(z, pol(2)[1], pol(2)[2])
    for i in range(stage):
        out = block(out, i)
    top = block(out, stage)
    new = torch.tanh(F.conv2d(top, sd["_Generator__end_block.0.weight"], sd["_Generator__end_block.0.bias"]))
    if stage == 0:
        return new
    old = torch.tanh(F.conv2d(out, sd["_Generator__last_end_block.0.0.weight"], sd["_Generator__last_end_block.0.0.bias"]))
    old = F.interpolate(old, scale_factor=2.0, mode="nearest")
    return alpha * new + (1.0 - alpha) * old


def disc_forward(sd, x, alpha, stage, pol):
    def conv(h, w, b):
        wm, am, gm = pol(h.shape[2])
        h = q(h, am, gm)
        y = F.conv2d(h, rnd(w, wm), None, padding=1) + b[None, :, None, None]
        y = q(y, "fp32", gm)
        return q(F.leaky_relu(y, 0.2), am, gm)

    def block(h, j):
        p = f"_Discriminator__conv_blocks.{j}."
        h = conv(h, sd[p + "0.weight"], sd[p + "0.bias"])
        _, am, gm = pol(h.shape[2] // 2)
        h = q(F.avg_pool2d(h, 2, 2), am, gm)
        return conv(h, sd[p + "3.weight"], sd[p + "3.bias"])

    cur = 7 - stage
    _, am, gm = pol(x.shape[2])
    h = q(F.leaky_relu(F.conv2d(x, sd["_Discriminator__start_block.0.weight"], sd["_Discriminator__start_block.0.bias"]), 0.2), am, gm)
    h = block(h, cur)
    if stage >= 1:
        _, am, gm = pol(x.shape[2] // 2)
        old = q(F.leaky_relu(F.conv2d(F.avg_pool2d(x, 2, 2), sd["_Discriminator__last_start_block.1.0.weight"],
                                      sd["_Discriminator__last_start_block.1.0.bias"]), 0.2), am, gm)
        h = q(alpha * h + (1 - alpha) * old, am, gm)
    for j in range(cur + 1, len(no.D_CHANNELS)):
        h = block(h, j)
    return F.linear(h.flatten(1, -1), sd["_Discriminator__clf.0.weight"], sd["_Discriminator__clf.0.bias"])


def gradient_penalty(sd_d, x_real, x_gen, alpha, stage, eps, pol):
    x_hat = eps * x_real + (1 - eps) * x_gen
    if not x_hat.requires_grad:
        x_hat.requires_grad_(True)
    out = disc_forward(sd_d, x_hat, alpha, stage, pol)
    (g,) = torch.autograd.grad(out, x_hat, grad_outputs=torch.ones_like(out), create_graph=True, retain_graph=True)
    n = g.view(g.size(0), -1).norm(2, dim=1)
    return 10.0 * ((n - 1.0) ** 2.0).mean()


def rel(a, b):
    a, b = a.double().flatten(), b.double().flatten()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


def catg(grads, ref):
    keys = [k for k in ref if ref[k] is not None and grads.get(k) is not None]
    return torch.cat([grads[k].flatten() for k in keys]), torch.cat([ref[k].flatten() for k in keys])


def g_step(sd_g, sd_d, z, alpha, stage, pol):
    g = no._leaf(sd_g)
    out = disc_forward(sd_d, gen_forward(g, z, alpha, stage, pol), alpha, stage, pol)
    (-out.mean()).backward()
    return {k: v.grad for k, v in g.items()}


def d_step(sd_g, sd_d, z, x_real, eps, alpha, stage, pol):
    d = no._leaf(sd_d)
    with torch.no_grad():
        x_fake = gen_forward(sd_g, z, alpha, stage, pol)
    o_r, o_f = disc_forward(d, x_real, alpha, stage, pol), disc_forward(d, x_fake, alpha, stage, pol)
    loss = -(o_r.mean() - o_f.mean())
    gp = gradient_penalty(d, x_real, x_fake, alpha, stage, eps, pol)
    (loss + gp).backward()
    return {k: v.grad for k, v in d.items()}


def gp_only(sd_d, x_real, x_fake, eps, stage, pol):
    d = no._leaf(sd_d)
    gradient_penalty(d, x_real, x_fake, 0.5, stage, eps, pol).backward()
    return {k: v.grad for k, v in d.items()}


def study(policies, cases=None, gp_cases=((1, 2, 3.0), (3, 2, 3.0), (4, 2, 1.0))):
    cases = cases or list(CASES)
    for name in cases:
        stage, batch, alpha, z, z2, x_real, eps = case_inputs(name)
        sd_g, sd_d = no.make_state("gen", stage, 11), no.make_state("disc", stage, 12)
        ref_d = no.d_step(sd_g, sd_d, z, x_real, eps, alpha, stage)["grads"]
        ref_g = no.g_step(sd_g, sd_d, z2, alpha, stage)["grads"]
        for pname, pol in policies.items():
            eg = rel(*catg(g_step(sd_g, sd_d, z2, alpha, stage, pol), ref_g))
            ed = rel(*catg(d_step(sd_g, sd_d, z, x_real, eps, alpha, stage, pol), ref_d))
            print(f"{name:10s} {pname:28s} G-step {eg:.2e}  critic-step {ed:.2e}", flush=True)
    for stage, batch, scale in gp_cases:
        g = torch.Generator().manual_seed(77 + stage)
        r = 4 * 2 ** stage
        x_real = torch.rand(batch, 2, r, r, generator=g) * 2 - 1
        x_fake = torch.rand(batch, 2, r, r, generator=g) * 2 - 1
        eps = torch.rand(batch, 1, 1, 1, generator=g)
        sd_d = {k: v * scale for k, v in no.make_state("disc", stage, 5).items()}
        ref = gp_only(sd_d, x_real, x_fake, eps, stage, Policy("fp32", "fp32", "fp32", "fp32", "fp32", "fp32"))
        for pname, pol in policies.items():
            e = rel(*catg(gp_only(sd_d, x_real, x_fake, eps, stage, pol), ref))
            print(f"gp stage{stage} x{scale} {pname:28s} GP-only {e:.2e}", flush=True)


if __name__ == "__main__":
    P = Policy
    policies = {
        "all bf16": P(),
        "w split2, a/g bf16": P(w_hi="split2", w_lo="split2"),
        "w fp32, a/g bf16": P(w_hi="fp32", w_lo="fp32"),
        "w bf16, a/g fp32": P(a_hi="fp32", g_hi="fp32", a_lo="fp32", g_lo="fp32"),
        "w bf16, a fp32, g bf16": P(a_hi="fp32", a_lo="fp32"),
        "w split2, a fp32, g bf16": P(w_hi="split2", w_lo="split2", a_hi="fp32", a_lo="fp32"),
        "w split2, a bf16, g fp32": P(w_hi="split2", w_lo="split2", g_hi="fp32", g_lo="fp32"),
        "<=32 all split2; w split2": P(w_hi="split2", w_lo="split2", a_lo="split2", g_lo="split2", thr=32),
        "<=32 all fp32; w split2": P(w_hi="split2", w_lo="fp32", a_lo="fp32", g_lo="fp32", thr=32),
        "<=16 all fp32; w split2": P(w_hi="split2", w_lo="fp32", a_lo="fp32", g_lo="fp32", thr=16),
        "<=64 all fp32; w split2": P(w_hi="split2", w_lo="fp32", a_lo="fp32", g_lo="fp32", thr=64),
        "all tf32": P("tf32", "tf32", "tf32", "tf32", "tf32", "tf32"),
        "all split2": P("split2", "split2", "split2", "split2", "split2", "split2"),
    }
    study(policies)
