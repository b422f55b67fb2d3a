"""GPU parity of the tcgen05 implicit-GEMM 3x3 convolutions (C ABI mg_conv3x3_bf16 / mg_conv3x3_wgrad_bf16)
against torch fp32 conv2d on the SAME bf16-rounded operands (so only accumulation order and the final
bf16 rounding differ): rel-L2 <= 4e-3 (one bf16 rounding of the output is 2^-9 = 2e-3 per element)."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def rel_l2(a, b):
    a, b = a.float(), b.float()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


def _mk(B, C, H, W, seed):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(B, C, H, W, generator=g).cuda().bfloat16().contiguous(memory_format=torch.channels_last)


SHAPES = [  # (B, Cin, Cout, H, W)
    (1, 16, 16, 16, 8), (2, 32, 32, 32, 32), (2, 16, 32, 64, 64), (1, 48, 32, 40, 24), (3, 64, 48, 16, 16),
    (2, 128, 112, 8, 8), (2, 160, 160, 4, 4), (4, 144, 160, 2, 2), (2, 32, 128, 4, 4), (1, 32, 16, 128, 128),
    (1, 16, 32, 72, 20), (2, 32, 16, 40, 24), (1, 16, 16, 200, 12),      # ragged heights with 4 / 2 stacked blocks per step
]


@pytest.mark.parametrize("B,Cin,Cout,H,W", SHAPES)
def test_fprop(B, Cin, Cout, H, W):
    from musicgan_b200.networks import ops
    x = _mk(B, Cin, H, W, 1)
    g = torch.Generator().manual_seed(2)
    w = (torch.randn(Cout, Cin, 3, 3, generator=g) / (3 * Cin ** 0.5)).cuda()
    b = torch.randn(Cout, generator=g).cuda()
    y = ops.conv3x3(x, w, b)
    ref = F.conv2d(x.float(), w.bfloat16().float(), b, padding=1)
    assert y.shape == ref.shape and y.is_contiguous(memory_format=torch.channels_last)
    assert rel_l2(y, ref) <= 4e-3, rel_l2(y, ref)


@pytest.mark.parametrize("B,Cin,Cout,H,W", SHAPES[:7] + SHAPES[10:])
def test_dgrad(B, Cin, Cout, H, W):
    from musicgan_b200.networks import ops
    dy = _mk(B, Cout, H, W, 3)
    g = torch.Generator().manual_seed(4)
    w = (torch.randn(Cout, Cin, 3, 3, generator=g) / (3 * Cin ** 0.5)).cuda()
    dx = ops.conv3x3(dy, w, None, dgrad=True)
    ref = F.conv_transpose2d(dy.float(), w.bfloat16().float(), padding=1)
    assert dx.shape == ref.shape
    assert rel_l2(dx, ref) <= 4e-3, rel_l2(dx, ref)


@pytest.mark.parametrize("B,Cin,Cout,H,W", [(2, 16, 32, 64, 64), (3, 32, 16, 40, 24), (2, 48, 32, 40, 24), (8, 160, 160, 4, 4),
                                             (2, 32, 48, 64, 32), (1, 16, 16, 200, 12)])
def test_ring_wraparound(B, Cin, Cout, H, W, monkeypatch):
    """Few CTAs -> every CTA walks many pipeline steps, so halo slots, accumulators and their barrier phases wrap
    around several times (the default grids of these small shapes give each CTA a single step)."""
    from musicgan_b200.networks import ops
    monkeypatch.setenv("MG_CONV_MAX_CTAS", "2")
    x = _mk(B, Cin, H, W, 11)
    dy = _mk(B, Cout, H, W, 12)
    g = torch.Generator().manual_seed(13)
    w = (torch.randn(Cout, Cin, 3, 3, generator=g) / (3 * Cin ** 0.5)).cuda()
    b = torch.randn(Cout, generator=g).cuda()
    y = ops.conv3x3(x, w, b, lrelu=True)
    ref = F.leaky_relu(F.conv2d(x.float(), w.bfloat16().float(), b, padding=1), 0.2)
    assert rel_l2(y, ref) <= 4e-3, rel_l2(y, ref)
    dx = ops.conv3x3(dy, w, None, dgrad=True)
    ref = F.conv_transpose2d(dy.float(), w.bfloat16().float(), padding=1)
    assert rel_l2(dx, ref) <= 4e-3, rel_l2(dx, ref)
    torch.cuda.synchronize()


@pytest.mark.parametrize("Cin,Cout", [(16, 32), (64, 112), (160, 160)])
def test_bias_enters_through_the_accumulator(Cin, Cout):
    """The bias is added by one extra K = 16 MMA as a bf16 pair hi + lo (conv_igemm.cu): with zero weights the output
    must be bf16(LeakyReLU(bias)) -- i.e. the pair carries the fp32 bias to well below one bf16 ulp -- in every N slice."""
    from musicgan_b200.networks import ops
    x = _mk(2, Cin, 16, 16, 21)
    w = torch.zeros(Cout, Cin, 3, 3, device="cuda")
    b = (torch.randn(Cout, generator=torch.Generator().manual_seed(22)) * 3).cuda()
    y = ops.conv3x3(x, w, b, lrelu=True).float()
    ref = F.leaky_relu(b, 0.2).view(1, -1, 1, 1).expand_as(y)
    ulp = ref.abs() * 2.0 ** -8
    assert ((y - ref).abs() <= ulp).all()                                   # within one bf16 rounding everywhere
    assert (y == ref.bfloat16().float()).float().mean().item() >= 0.99      # and the correctly rounded value nearly always


@pytest.mark.parametrize("B,Cin,Cout,H,W", [(2, 32, 32, 16, 16), (1, 48, 32, 32, 16), (2, 128, 128, 4, 4), (2, 32, 16, 64, 64)])
def test_fused_epilogue_and_upsample(B, Cin, Cout, H, W):
    """conv3x3(nearest_up2(x)) + bias -> LeakyReLU(0.2) -> PixelNorm == generator.py:26-40 second half of Block."""
    from musicgan_b200.networks import ops
    x = _mk(B, Cin, H // 2, W // 2, 5)
    g = torch.Generator().manual_seed(6)
    w = (torch.randn(Cout, Cin, 3, 3, generator=g) / (3 * Cin ** 0.5)).cuda()
    b = torch.randn(Cout, generator=g).cuda()
    fits = 9 * Cin * Cout * 2 <= 184 * 1024
    if not fits:
        y = ops.conv3x3(x, w, b, lrelu=True, upsample_in=True)
        ref = F.leaky_relu(F.conv2d(F.interpolate(x.float(), scale_factor=2.0, mode="nearest"), w.bfloat16().float(), b, padding=1), 0.2)
        assert rel_l2(y, ref) <= 4e-3
        return
    y, inv = ops.conv3x3(x, w, b, lrelu=True, pixelnorm=True, upsample_in=True, want_inv_norm=True)
    z = F.leaky_relu(F.conv2d(F.interpolate(x.float(), scale_factor=2.0, mode="nearest"), w.bfloat16().float(), b, padding=1), 0.2)
    ref = z / torch.sqrt(z.pow(2.0).mean(dim=1, keepdim=True) + 1e-8)
    assert rel_l2(y, ref) <= 4e-3, rel_l2(y, ref)
    ref_inv = 1.0 / torch.sqrt(z.pow(2.0).mean(dim=1) + 1e-8)
    assert rel_l2(inv, ref_inv) <= 1e-4


@pytest.mark.parametrize("B,Cin,Cout,H,W", SHAPES)
def test_wgrad(B, Cin, Cout, H, W):
    from musicgan_b200.networks import ops
    x = _mk(B, Cin, H, W, 7)
    dy = _mk(B, Cout, H, W, 8)
    dw = ops.conv3x3_wgrad(dy, x)
    ref = torch.nn.grad.conv2d_weight(x.float(), (Cout, Cin, 3, 3), dy.float(), padding=1)
    assert dw.shape == ref.shape and dw.dtype == torch.float32
    assert rel_l2(dw, ref) <= 1e-4, rel_l2(dw, ref)       # fp32 accumulate of exact bf16 products


@pytest.mark.parametrize("B,Cin,Cout,H,W", SHAPES + [(3, 128, 32, 4, 4), (2, 256, 48, 8, 8), (8, 16, 32, 40, 24)])
def test_wgrad_with_bias_row(B, Cin, Cout, H, W):
    """mg_conv3x3_wgrad_bias_bf16: the bias gradient as one more GEMM row (ones), written / accumulated together with dw
    (also where 9 * Cin is a multiple of 128 and the row opens a block of its own)."""
    from musicgan_b200.networks import ops
    x = _mk(B, Cin, H, W, 17)
    dy = _mk(B, Cout, H, W, 18)
    ref_w = torch.nn.grad.conv2d_weight(x.float(), (Cout, Cin, 3, 3), dy.float(), padding=1)
    ref_b = dy.float().sum(dim=(0, 2, 3))
    dw = torch.full((Cout, Cin, 3, 3), float("nan"), device="cuda")
    db = torch.full((Cout,), float("nan"), device="cuda")
    ops.conv3x3_wgrad(dy, x, out=dw, bias_out=db)
    assert rel_l2(dw, ref_w) <= 1e-4 and rel_l2(db, ref_b) <= 1e-5, (rel_l2(dw, ref_w), rel_l2(db, ref_b))
    assert torch.equal(dw, ops.conv3x3_wgrad(dy, x))          # the extra row does not disturb the others
    ops.conv3x3_wgrad(dy, x, out=dw, accumulate=True, bias_out=db, accumulate_bias=True)
    assert rel_l2(dw, 2 * ref_w) <= 1e-4 and rel_l2(db, 2 * ref_b) <= 1e-5
    ops.conv3x3_wgrad(dy, x, out=dw, accumulate=True, bias_out=db)      # dw += , db =
    assert rel_l2(dw, 3 * ref_w) <= 1e-4 and rel_l2(db, ref_b) <= 1e-5
    ops.conv3x3_wgrad(dy, x, out=dw, accumulate=True)                   # no bias row: db untouched
    assert rel_l2(dw, 4 * ref_w) <= 1e-4 and rel_l2(db, ref_b) <= 1e-5


def test_wgrad_upsampled_input():
    from musicgan_b200.networks import ops
    x = _mk(2, 48, 16, 8, 9)
    dy = _mk(2, 32, 32, 16, 10)
    dw = ops.conv3x3_wgrad(dy, x, upsample_in=True)
    xin = F.interpolate(x.float(), scale_factor=2.0, mode="nearest")
    ref = torch.nn.grad.conv2d_weight(xin, (32, 48, 3, 3), dy.float(), padding=1)
    assert rel_l2(dw, ref) <= 1e-4


def test_rejects_bad_arguments():
    from musicgan_b200.networks import ops
    from musicgan_b200._lib import MgError
    x = _mk(1, 24, 8, 8, 0)       # channels not a multiple of 16
    w = torch.randn(16, 24, 3, 3).cuda()
    with pytest.raises(MgError):
        ops.conv3x3(x, w)
    with pytest.raises(ValueError):
        ops.conv3x3(torch.randn(1, 16, 8, 8).cuda().bfloat16(), torch.randn(16, 16, 3, 3).cuda())   # not channels_last


def test_rgb_layers_and_pooling():
    """1x1 to/from magnitude-phase layers and 2x2 average pooling against torch fp32."""
    from musicgan_b200.networks import ops
    torch.backends.cudnn.allow_tf32 = False          # the torch fp32 reference must really be fp32
    torch.backends.cuda.matmul.allow_tf32 = False
    g = torch.Generator().manual_seed(11)
    B, C, H, W = 2, 48, 24, 40
    x = (torch.rand(B, 2, H, W, generator=g) * 2 - 1).cuda()
    w = torch.randn(C, 2, 1, 1, generator=g).cuda()
    b = torch.randn(C, generator=g).cuda()
    y = ops.rgb_expand(x, w, b, lrelu=True)
    ref = F.leaky_relu(F.conv2d(x, w, b), 0.2)
    assert rel_l2(y, ref) <= 4e-3
    a = _mk(B, C, H, W, 12)
    m = _mk(B, C, H, W, 13)
    w2 = torch.randn(2, C, 1, 1, generator=g).cuda()
    b2 = torch.randn(2, generator=g).cuda()
    out = ops.rgb_project(a, w2, bias=b2, tanh=True)
    assert rel_l2(out, torch.tanh(F.conv2d(a.float(), w2, b2))) <= 1e-5
    mask = torch.where(m.float() > 0, 1.0, 0.2)
    out2 = ops.rgb_project(a, w, mask_src=m, w_is_c_by_2=True)
    assert rel_l2(out2, F.conv_transpose2d(a.float() * mask, w)) <= 1e-5
    gw, gb = ops.rgb_wgrad(a, m, x)
    am = a.float() * mask
    assert rel_l2(gw, torch.einsum("bchw,bkhw->ck", am, x)) <= 1e-4 and rel_l2(gb, am.sum((0, 2, 3))) <= 1e-4
    y2 = ops.rgb_expand(x, w, None, mask_src=m)
    assert rel_l2(y2, F.conv2d(x, w) * mask) <= 4e-3
    p = ops.pool2(a)
    assert rel_l2(p, F.avg_pool2d(a.float(), 2, 2)) <= 4e-3
    u = ops.pool2(p, adjoint=True)
    assert rel_l2(u, F.interpolate(p.float(), scale_factor=2.0, mode="nearest") * 0.25) <= 4e-3


def test_pixelnorm_lrelu_backward_kernel():
    """Fused backward of LeakyReLU -> PixelNorm against torch autograd (fp32) on the same bf16 inputs."""
    from musicgan_b200.networks import ops
    B, C, H, W = 2, 48, 12, 20
    g = torch.Generator().manual_seed(21)
    z = torch.randn(B, C, H, W, generator=g).cuda().requires_grad_(True)
    t = F.leaky_relu(z, 0.2)
    n = torch.sqrt(t.pow(2.0).mean(dim=1, keepdim=True) + 1e-8)
    o = t / n
    go = torch.randn(B, C, H, W, generator=g).cuda().bfloat16().contiguous(memory_format=torch.channels_last)
    ob = o.detach().bfloat16().contiguous(memory_format=torch.channels_last)
    inv = (1.0 / n.detach())[:, 0].contiguous()
    gz, gb = ops.pixelnorm_lrelu_bwd(go, ob, inv)
    # reference: differentiate with o, n replaced by their bf16 / fp32 saved values (the kernel's inputs)
    of, gf = ob.float(), go.float()
    ref = (gf - of * (gf * of).mean(dim=1, keepdim=True)) * inv[:, None] * torch.where(of > 0, 1.0, 0.2)
    assert rel_l2(gz, ref) <= 4e-3
    assert rel_l2(gb, ref.sum((0, 2, 3))) <= 4e-3
    (auto,) = torch.autograd.grad(o, z, go.float())
    assert rel_l2(gz, auto) <= 2e-2          # vs exact autograd: only the bf16 rounding of the saved output differs
    s = ops.pool2(go, sum_pool=True)
    assert rel_l2(s, F.avg_pool2d(go.float(), 2) * 4) <= 4e-3


@pytest.mark.parametrize("B,C,H,W", [(2, 32, 16, 16), (3, 48, 40, 24), (1, 160, 4, 4), (8, 16, 64, 32)])
def test_unpool_lrelu_backward_kernel(B, C, H, W):
    """mg_unpool2_lrelu_bwd_bf16 == mg_pool2_bf16(adjoint) followed by mg_lrelu_bwd_bf16: identical bf16 values
    (0.25 * g is exact), bias sums equal up to fp32 summation order; both bias-gradient variants."""
    from musicgan_b200.networks import ops
    gp = _mk(B, C, H // 2, W // 2, 31)
    h = _mk(B, C, H, W, 32)
    gz_ref, gb_ref = ops.lrelu_bwd(ops.pool2(gp, adjoint=True), h)
    gz, gb = ops.unpool_lrelu_bwd(gp, h)
    assert torch.equal(gz, gz_ref)
    assert torch.allclose(gb, gb_ref, rtol=1e-5, atol=1e-4 * gb_ref.abs().max().item())
    exact = (0.25 * F.interpolate(gp.float(), scale_factor=2.0, mode="nearest") * torch.where(h.float() > 0, 1.0, 0.2)).sum(dim=(0, 2, 3))
    assert torch.allclose(gb, exact, rtol=1e-4, atol=1e-4 * exact.abs().max().item())
    gz2, none = ops.unpool_lrelu_bwd(gp, h, want_bias_grad=False)
    assert none is None and torch.equal(gz2, gz_ref)


def test_pool2_planes_matches_torch_bitwise():
    """mg_pool2_planes_f32 (fp32 NCHW input planes of the fade-in path) == F.avg_pool2d and its backward, bit for bit,
    including through autograd (double backward = the pooling again)."""
    from musicgan_b200.networks import ops, functional as fn
    g = torch.Generator().manual_seed(41)
    x = torch.randn(3, 2, 64, 48, generator=g).cuda()
    assert torch.equal(ops.pool2_planes(x), F.avg_pool2d(x, 2, 2))
    gsmall = torch.randn(3, 2, 32, 24, generator=g).cuda()
    xr = x.clone().requires_grad_(True)
    ref = torch.autograd.grad(F.avg_pool2d(xr, 2, 2), xr, gsmall)[0]
    assert torch.equal(ops.pool2_planes(gsmall, adjoint=True), ref)
    xa = x.clone().requires_grad_(True)
    got = torch.autograd.grad(fn.PoolPlanes.apply(xa), xa, gsmall, create_graph=True)[0]
    assert torch.equal(got, ref)
    gs2 = gsmall.clone().requires_grad_(True)
    gg = torch.autograd.grad(fn.UnpoolPlanes.apply(gs2), gs2, x)[0]
    assert torch.equal(gg, F.avg_pool2d(x, 2, 2))


# ---- precise path: fp32 activations, split-bf16 operands (conv_split.cu) ---------------------------------------------
def _mk32(B, C, H, W, seed):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(B, C, H, W, generator=g).cuda().contiguous(memory_format=torch.channels_last)


SPLIT_SHAPES = [  # (B, Cin, Cout, H, W): the layers with output height <= 32 of both networks + ragged sizes
    (2, 32, 32, 2, 2), (2, 32, 128, 4, 4), (3, 128, 112, 8, 8), (2, 112, 96, 16, 16), (2, 96, 80, 32, 32),
    (2, 80, 96, 32, 32), (4, 144, 160, 2, 2), (5, 160, 160, 1, 1), (1, 16, 16, 20, 12), (1, 48, 64, 32, 72),
]


@pytest.mark.parametrize("B,Cin,Cout,H,W", SPLIT_SHAPES)
def test_split_fprop_and_dgrad(B, Cin, Cout, H, W):
    """x = hi + lo, w = hi + lo, three MMAs per K step: the result must sit within a few 2^-17 of the fp64 convolution
    of the fp32 operands (a bf16-operand kernel is at 2^-9)."""
    from musicgan_b200.networks import ops
    torch.backends.cudnn.allow_tf32 = False
    x = _mk32(B, Cin, H, W, 51)
    g = torch.Generator().manual_seed(52)
    w = (torch.randn(Cout, Cin, 3, 3, generator=g) / (3 * Cin ** 0.5)).cuda()
    b = torch.randn(Cout, generator=g).cuda()
    y = ops.conv3x3(x, w, b, lrelu=True)
    assert y.dtype == torch.float32 and y.is_contiguous(memory_format=torch.channels_last)
    ref = F.leaky_relu(F.conv2d(x.double(), w.double(), b.double(), padding=1), 0.2)
    e = ((y.double() - ref).norm() / ref.norm()).item()
    assert e <= 2e-5, e
    dy = _mk32(B, Cout, H, W, 53)
    dx = ops.conv3x3(dy, w, None, dgrad=True)
    ref = F.conv_transpose2d(dy.double(), w.double(), padding=1)
    e = ((dx.double() - ref).norm() / ref.norm()).item()
    assert dx.shape == (B, Cin, H, W) and e <= 2e-5, e


def test_split_exact_weights():
    """flag 32: w = hi + mid + lo (the fp32 weight exactly, five MMAs per K step).  On a constant image a centre tap of
    1 + 2^-9 + 2^-19 and a left tap of -1 leave 2^-9 + 2^-19 in the interior; two parts (hi + mid) stop at 2^-9."""
    from musicgan_b200.networks import ops
    for C in (16, 96, 160):
        x = torch.ones(2, C, 8, 16, device="cuda").contiguous(memory_format=torch.channels_last)
        w = torch.zeros(C, C, 3, 3, device="cuda")
        w[torch.arange(C), torch.arange(C), 1, 1] = 1.0 + 2.0 ** -9 + 2.0 ** -19
        w[torch.arange(C), torch.arange(C), 1, 0] = -1.0
        y3 = ops.conv3x3(x, w, None, exact_w=True)
        y2 = ops.conv3x3(x, w, None)
        assert torch.equal(y3[:, :, :, 1:], torch.full_like(y3[:, :, :, 1:], 2.0 ** -9 + 2.0 ** -19))
        assert torch.equal(y2[:, :, :, 1:], torch.full_like(y2[:, :, :, 1:], 2.0 ** -9))
    # random data: same accuracy class as the two-part kernel (the activations still carry 16 bits), all N slices
    x = _mk32(3, 144, 4, 4, 57)
    g = torch.Generator().manual_seed(58)
    w = (torch.randn(160, 144, 3, 3, generator=g) / 36).cuda()
    b = torch.randn(160, generator=g).cuda()
    y = ops.conv3x3(x, w, b, lrelu=True, exact_w=True)
    ref = F.leaky_relu(F.conv2d(x.double(), w.double(), b.double(), padding=1), 0.2)
    assert ((y.double() - ref).norm() / ref.norm()).item() <= 2e-5


@pytest.mark.parametrize("B,Cin,Cout,H,W", [(2, 32, 128, 4, 4), (2, 128, 112, 8, 8), (1, 96, 80, 32, 32), (2, 32, 16, 24, 40)])
def test_split_fused_epilogue_and_upsample(B, Cin, Cout, H, W):
    from musicgan_b200.networks import ops
    x = _mk32(B, Cin, H // 2, W // 2, 55)
    g = torch.Generator().manual_seed(56)
    w = (torch.randn(Cout, Cin, 3, 3, generator=g) / (3 * Cin ** 0.5)).cuda()
    b = torch.randn(Cout, generator=g).cuda()
    y, inv = ops.conv3x3(x, w, b, lrelu=True, pixelnorm=True, upsample_in=True, want_inv_norm=True)
    z = F.leaky_relu(F.conv2d(F.interpolate(x.double(), scale_factor=2.0, mode="nearest"), w.double(), b.double(), padding=1), 0.2)
    ref = z / torch.sqrt(z.pow(2.0).mean(dim=1, keepdim=True) + 1e-8)
    e = ((y.double() - ref).norm() / ref.norm()).item()
    assert e <= 2e-5, e
    ref_inv = 1.0 / torch.sqrt(z.pow(2.0).mean(dim=1) + 1e-8)
    assert ((inv.double() - ref_inv).norm() / ref_inv.norm()).item() <= 2e-5


def test_split_weights_in_the_bf16_kernel():
    """flag 16: weights enter as hi + lo bf16 pairs.  On a constant image a centre tap of 256.5 (hi = 256, lo = 0.5) and a
    left tap of -256 must leave 0.5 in the interior -- the low halves alone; with plain bf16 weights the result is 0."""
    from musicgan_b200.networks import ops
    for C in (16, 48, 80):
        x = torch.ones(2, C, 32, 24, device="cuda").bfloat16().contiguous(memory_format=torch.channels_last)
        w = torch.zeros(C, C, 3, 3, device="cuda")
        w[torch.arange(C), torch.arange(C), 1, 1] = 256.5
        w[torch.arange(C), torch.arange(C), 1, 0] = -256.0
        y = ops.conv3x3(x, w, None, split_w=True).float()
        assert torch.equal(y[:, :, :, 1:], torch.full_like(y[:, :, :, 1:], 0.5))
        assert torch.equal(y[:, :, :, 0], torch.full_like(y[:, :, :, 0], 256.0))      # left tap in the padding: bf16(256.5)
        y1 = ops.conv3x3(x, w, None).float()                       # without the flag: bf16(w) = 256
        assert torch.equal(y1[:, :, :, 1:], torch.zeros_like(y1[:, :, :, 1:]))
    # and through the PixelNorm epilogue of the widest such layer (all Cout in one slice, weights 2 x 92 KB)
    x = _mk(1, 80, 32, 32, 62)
    g = torch.Generator().manual_seed(63)
    w = (torch.randn(64, 80, 3, 3, generator=g) / (3 * 80 ** 0.5)).cuda()
    b = torch.randn(64, generator=g).cuda()
    y = ops.conv3x3(x, w, b, lrelu=True, pixelnorm=True, upsample_in=True, split_w=True)
    z = F.leaky_relu(F.conv2d(F.interpolate(x.float(), scale_factor=2.0, mode="nearest"), w, b, padding=1), 0.2)
    ref = z / torch.sqrt(z.pow(2.0).mean(dim=1, keepdim=True) + 1e-8)
    assert rel_l2(y, ref) <= 4e-3, rel_l2(y, ref)


def test_pointwise_kernels_on_fp32_activations():
    """The fp32-activation instantiations of the memory-bound kernels against torch fp32."""
    from musicgan_b200.networks import ops
    torch.backends.cudnn.allow_tf32 = False
    g = torch.Generator().manual_seed(71)
    B, C, H, W = 3, 48, 12, 20
    a, m, gy = _mk32(B, C, H, W, 72), _mk32(B, C, H, W, 73), _mk32(B, C, H, W, 74)
    mask = torch.where(m > 0, 1.0, 0.2)
    gz, gb = ops.lrelu_bwd(gy, m)
    assert gz.dtype == torch.float32 and torch.equal(gz, gy * mask)
    assert torch.allclose(gb, (gy * mask).sum((0, 2, 3)), rtol=1e-5, atol=1e-4)
    gp = _mk32(B, C, H // 2, W // 2, 75)
    gz2, gb2 = ops.unpool_lrelu_bwd(gp, m)
    ref = 0.25 * F.interpolate(gp, scale_factor=2.0, mode="nearest") * mask
    assert torch.equal(gz2, ref) and torch.allclose(gb2, ref.sum((0, 2, 3)), rtol=1e-5, atol=1e-4)
    assert torch.allclose(ops.pool2(a), F.avg_pool2d(a, 2, 2), rtol=1e-6, atol=1e-6)
    assert torch.equal(ops.pool2(gp, adjoint=True), F.interpolate(gp, scale_factor=2.0, mode="nearest") * 0.25)
    assert torch.allclose(ops.pool2(a, sum_pool=True), F.avg_pool2d(a, 2, 2) * 4, rtol=1e-6, atol=1e-6)
    x = (torch.rand(B, 2, H, W, generator=g) * 2 - 1).cuda()
    w = torch.randn(C, 2, 1, 1, generator=g).cuda()
    b = torch.randn(C, generator=g).cuda()
    y = ops.rgb_expand(x, w, b, lrelu=True, out_dtype=torch.float32)
    assert y.dtype == torch.float32 and rel_l2(y, F.leaky_relu(F.conv2d(x, w, b), 0.2)) <= 1e-6
    assert rel_l2(ops.rgb_expand(x, w, None, mask_src=m), F.conv2d(x, w) * mask) <= 1e-6
    w2 = torch.randn(2, C, 1, 1, generator=g).cuda()
    b2 = torch.randn(2, generator=g).cuda()
    assert rel_l2(ops.rgb_project(a, w2, bias=b2, tanh=True), torch.tanh(F.conv2d(a, w2, b2))) <= 1e-5
    assert rel_l2(ops.rgb_project(a, w, mask_src=m, w_is_c_by_2=True), F.conv_transpose2d(a * mask, w)) <= 1e-5
    gw, gbb = ops.rgb_wgrad(a, m, x)
    assert rel_l2(gw, torch.einsum("bchw,bkhw->ck", a * mask, x)) <= 1e-4 and rel_l2(gbb, (a * mask).sum((0, 2, 3))) <= 1e-4
    # PixelNorm + LeakyReLU backward against autograd on fp32
    z = _mk32(B, C, H, W, 76).requires_grad_(True)
    t = F.leaky_relu(z, 0.2)
    n = torch.sqrt(t.pow(2.0).mean(dim=1, keepdim=True) + 1e-8)
    o = t / n
    gz3, gb3 = ops.pixelnorm_lrelu_bwd(gy, o.detach().contiguous(memory_format=torch.channels_last), (1.0 / n.detach())[:, 0].contiguous())
    (auto,) = torch.autograd.grad(o, z, gy)
    assert rel_l2(gz3, auto) <= 1e-5 and rel_l2(gb3, auto.sum((0, 2, 3))) <= 1e-4
    # weight gradient from fp32 activations (bf16 operands inside: terminal product)
    dw = ops.conv3x3_wgrad(gy, a)
    ref = torch.nn.grad.conv2d_weight(a.bfloat16().float(), (C, C, 3, 3), gy.bfloat16().float(), padding=1)
    assert rel_l2(dw, ref) <= 1e-4
